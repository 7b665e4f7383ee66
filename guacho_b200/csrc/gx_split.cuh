// gx_split.cuh — HLLE with every variable split into background + fluctuation (SOLVER_HLLE_SPLIT_ALL):
// src/hlle_split_all.f90:51-85 (prim2fhlleSplitAll) and the split branches of src/hydro_core.f90 (u2primSplitAll :143-229,
// prim2u :340-368, prim2f :404-427,457-461).  The state arrays hold FLUCTUATIONS about the background primitives primit0,
// which the host sets (gx_set_background).  MHD only (the split flux exists under `if (mhd)` alone), no passive scalars.
// Same expression order as the reference in both build flavours; the production build differs by FMA contraction and the
// shared-reciprocal division policy of gx_physics.cuh.
#pragma once
#include "gx_physics.cuh"

namespace gxp {

__device__ __forceinline__ void u2prim_split(const Phys& P, const double (&u)[8], const double (&w0)[8], double (&w)[8], double& T) {
  w[0] = u[0];
  const double r = w[0] + w0[0];
  const Den dr(r);
  w[1] = dr.div(u[1]);
  w[2] = dr.div(u[2]);
  w[3] = dr.div(u[3]);
  w[5] = u[5]; w[6] = u[6]; w[7] = u[7];
  w[4] = Den(P.cv).div(u[4] - 0.5 * r * (w[1] * w[1] + w[2] * w[2] + w[3] * w[3])
                            - 0.5 * (u[5] * u[5] + u[6] * u[6] + u[7] * u[7])
                            - w0[5] * u[5] - w0[6] * u[6] - w0[7] * u[7]);
  T = 0.0;
  if (P.eos == GX_EOS_ADIABATIC) T = dr.div(w[4] + w0[4]) * P.Tempsc;
  else if (P.eos == GX_EOS_SINGLE_SPECIE) T = gx_max(1., Den(gx_max(r, 1e-15)).div(w[4] + w0[4]) * P.Tempsc);
}

__device__ __forceinline__ void prim2u_split(const Phys& P, const double (&w)[8], const double (&w0)[8], double (&uu)[8]) {
  const double v2 = w[1] * w[1] + w[2] * w[2] + w[3] * w[3];
  uu[0] = w[0];
  uu[1] = w[1] * (w[0] + w0[0]);
  uu[2] = w[2] * (w[0] + w0[0]);
  uu[3] = w[3] * (w[0] + w0[0]);
  uu[4] = 0.5 * w[0] * v2 + P.cv * w[4];
  uu[4] = uu[4] + 0.5 * (w[5] * w[5] + w[6] * w[6] + w[7] * w[7]) + 0.5 * w0[0] * v2 + w0[5] * w[5] + w0[6] * w[6] + w0[7] * w[7];
  uu[5] = w[5]; uu[6] = w[6]; uu[7] = w[7];
}

__device__ __forceinline__ void prim2f_split(const Phys& P, const double (&w)[8], const double (&w0)[8], double (&ff)[8]) {
  const double rt = w[0] + w0[0];
  const double etot = 0.5 * (rt * (w[1] * w[1] + w[2] * w[2] + w[3] * w[3]) + w[5] * w[5] + w[6] * w[6] + w[7] * w[7])
                      + P.cv * w[4] + w0[5] * w[5] + w0[6] * w[6] + w0[7] * w[7];
  ff[0] = rt * w[1];
  ff[1] = rt * w[1] * w[1] + w[4] + 0.5 * (w[6] * w[6] + w[7] * w[7] - w[5] * w[5]) - w0[5] * w[5] + w0[6] * w[6] + w0[7] * w[7];
  ff[2] = rt * w[1] * w[2] - w[5] * w[6] - w0[6] * w[5] - w0[5] * w[6];
  ff[3] = rt * w[1] * w[3] - w[5] * w[7] - w0[7] * w[5] - w0[5] * w[7];
  ff[4] = w[1] * (etot + w[4] + 0.5 * ((w[5] + w0[5]) * (w[5] + w0[5]) + (w[6] + w0[6]) * (w[6] + w0[6]) + (w[7] + w0[7]) * (w[7] + w0[7]))
                  + P.cv * w0[4] + w0[4] + 0.5 * (w0[5] * w0[5] + w0[6] * w0[6] + w0[7] * w0[7]))
          - (w[5] + w0[5]) * (w[1] * (w[5] + w0[5]) + w[2] * (w[6] + w0[6]) + w[3] * (w[7] + w0[7]));
  ff[5] = 0.;
  ff[6] = w[1] * (w0[6] + w[6]) - w[2] * (w0[5] + w[5]);
  ff[7] = w[1] * (w0[7] + w[7]) - w[3] * (w0[5] + w[5]);
}

__device__ __forceinline__ void riemann_hlle_split_all(const Phys& P, const double (&wl)[8], const double (&wr)[8],
                                                       const double (&w0l)[8], const double (&w0r)[8], double (&ff)[8]) {
  double tl[8], tr[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) { tl[q] = wl[q] + w0l[q]; tr[q] = wr[q] + w0r[q]; }
  const double csl = cfastX(P, tl), csr = cfastX(P, tr);
  const double sr = gx_max(wl[1] + w0l[1] + csl, wr[1] + w0r[1] + csr);
  const double sl = gx_min(wl[1] + w0l[1] - csl, wr[1] + w0r[1] - csr);
  if (sl > 0) { prim2f_split(P, wl, w0l, ff); return; }
  if (sr < 0) { prim2f_split(P, wr, w0r, ff); return; }
  double fL[8], fR[8], uL[8], uR[8];
  prim2f_split(P, wl, w0l, fL); prim2f_split(P, wr, w0r, fR);
  prim2u_split(P, wl, w0l, uL); prim2u_split(P, wr, w0r, uR);
  const Den ds(sr - sl);
#pragma unroll
  for (int q = 0; q < 8; ++q) ff[q] = ds.div(sr * fL[q] - sl * fR[q] + sl * sr * (uR[q] - uL[q]));
}

}  // namespace gxp
