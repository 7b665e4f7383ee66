// gx_thermal.cuh — thermal conduction (src/thermal_cond.f90), the operator the reference applies at the end of tstep
// (hydro_solver.f90:227).  Included by gx_api.cu only, i.e. compiled with -fmad=false and IEEE div/sqrt: the expressions
// below are evaluated in the reference's order; pow() is CUDA's FP64 routine (<= 2 ulp against libm's).
//
// Reference structure per substep: heatfluxes / MHD_heatfluxes fill f,g,h(5,...) over 0..n from primit and Temp, u(5) is
// updated over the physical cells, thermal_bounds exchanges one layer of u(5), calcprim(u) refreshes primit and Temp.
// Here: no flux arrays.  A substep is ONE kernel over the physical cells that evaluates the (up to) six face fluxes of its
// cell from a two-variable scratch array (pressure, temperature; the density and B come straight from u) and updates u(5)
// in place (the fluxes do not read u(5)), followed by the ghost layer of u(5) and k_tc_prim (the calcprim of the reference,
// reduced to the two variables the operator reads).
#pragma once
#include "gx_kernels.cuh"

namespace gxtc {

struct TcPar {
  int mode;                  // GX_TC_ISOTROPIC | GX_TC_ANISOTROPIC
  int sat;                   // tc_saturation
  int mhd;
  double dxr, dyr, dzr;      // dx*rsc, dy*rsc, dz*rsc (the reference divides by the product)
  double dx, dy, dz;
  double vsc, sqrt_vsc2;     // parameters.f90:167 ; sqrt(vsc2) as heatfluxes spells it (:213)
  double Psc, rhosc, bsc2;   // Psc = rhosc*vsc2 (:168) ; bsc**2
};

__device__ __forceinline__ double Ksp(double T) { return 6.e-7 * pow(T, 2.5); }                    // thermal_cond.f90:142-149
__device__ __forceinline__ double Ksp_parl(double T) { return 9.2181e-7 * pow(T, 2.5); }           // :157-164
__device__ __forceinline__ double Ksp_perp(double T, double dens, double B2) { return 0.30089e+33 * dens / (B2 * sqrt(T)) * dens; }   // :172-178

// calcprim reduced to what the operator reads: PT[0] = primit(5), PT[1] = Temp over 0..n+1 (every cell a heat flux touches);
// want_dt: the Spitzer time-scale candidates primit(1)/Ksp(Temp) of get_dt_cond (:90-99) over the physical cells
__global__ void __launch_bounds__(128) k_tc_prim(gx::Grid g, gxp::Phys P, int mhd, const double* __restrict__ U, double* __restrict__ PT,
                                                 unsigned long long* __restrict__ dt_bits, int want_dt) {
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x), j = (int)blockIdx.y, k = (int)blockIdx.z;
  double cand = 1.7976931348623157e308;
  if (i <= g.nx + 1) {
    const long long c = g.idx(i, j, k), vs = g.vs;
    double uu[8], prim[8], T;
#pragma unroll
    for (int q = 0; q < 8; ++q) uu[q] = (q < g.neqdyn) ? U[q * vs + c] : 0.0;
    const double un = g.npas > 0 ? U[(long long)g.neqdyn * vs + c] : 0.0;
    if (mhd) gxp::u2prim<true, true>(P, uu, prim, un, T); else gxp::u2prim<false, true>(P, uu, prim, un, T);
    PT[c] = prim[4];
    PT[vs + c] = T;
    if (want_dt && i >= 1 && i <= g.nx && j >= 1 && j <= g.ny && k >= 1 && k <= g.nz) cand = prim[0] / Ksp(T);
  }
  if (want_dt) {                                      // positive doubles order like their bit patterns
    for (int o = 16; o > 0; o >>= 1) cand = fmin(cand, __shfl_xor_sync(0xffffffffu, cand, o));
    if ((threadIdx.x & 31) == 0 && cand < 1.7976931348623157e308) atomicMin(dt_bits, (unsigned long long)__double_as_longlong(cand));
  }
}

// heatfluxes (:189-267): flux through the face between cell a (low side) and cell b
__device__ __forceinline__ double flux_iso(const TcPar& t, const gxp::Phys& P, double Ta, double Tb, double pa, double pb, double ra, double rb, double dxr) {
  if (Ta == Tb) return 0.;
  const double meanP = 0.5 * (pa + pb);
  const double meanDens = 0.5 * (ra + rb);
  const double meanT = 0.5 * (Ta + Tb);
  const double dT = (Tb - Ta) / dxr;
  double coef;
  if (t.sat) {
    double cs = gxp::csound(P, meanP, meanDens);
    cs = fmin(cs * t.sqrt_vsc2, 3.E10);
    coef = fmin(Ksp(meanT), 5. * 0.4 * cs * meanP * t.Psc / fabs(dT));
  } else coef = Ksp(meanT);
  return -coef * dT * 1.;
}

// MHD_heatfluxes (:277-487): component D of the flux triplet the reference stores at cell c — B of cell c, forward
// temperature differences of cell c in all three directions
template <int D>
__device__ __forceinline__ double flux_aniso(const TcPar& t, const gxp::Phys& P, const gx::Grid& g, const double* __restrict__ U, const double* __restrict__ PT, long long c) {
  const long long vs = g.vs;
  const long long nb[3] = {c + 1, c + g.px, c + (long long)g.px * g.py};
  const double dr[3] = {t.dxr, t.dyr, t.dzr};
  double bx = U[5 * vs + c], by = U[6 * vs + c], bz = U[7 * vs + c];
  const double B2 = bx * bx + by * by + bz * bz;
  const double modB = sqrt(B2);
  bx = bx / modB; by = by / modB; bz = bz / modB;
  const double Tc = PT[vs + c], rc = fmax(U[c], 1e-15), pc = PT[c];
  double grad[3], Kparl = 0.0, Kperp = 0.0, coefSat = 0.0;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double Tn = PT[vs + nb[d]];
    if (fabs(Tc - Tn) < 1.0e-14) { grad[d] = 0.0; continue; }
    grad[d] = (Tn - Tc) / dr[d];
    if (d == D) {
      const double meanDens = 0.5 * (rc + fmax(U[nb[d]], 1e-15));
      const double meanTemp = 0.5 * (Tc + Tn);
      if (t.sat) {
        const double meanPres = 0.5 * (pc + PT[nb[d]]);
        double cs = gxp::csound(P, meanPres, meanDens);
        cs = fmin(cs * t.vsc, 3.E10);
        coefSat = (5.0 * 0.3) * meanDens * (cs * cs * cs);
      }
      Kparl = Ksp_parl(meanTemp);
      Kperp = Ksp_perp(meanTemp, meanDens * t.rhosc, B2 * t.bsc2);
    }
  }
  const double bgradT = bx * grad[0] + by * grad[1] + bz * grad[2];
  const double parl[3] = {bgradT * bx, bgradT * by, bgradT * bz};
  const double perp[3] = {grad[0] - parl[0], grad[1] - parl[1], grad[2] - parl[2]};
  if (!t.sat) return -Kparl * parl[D] - Kperp * perp[D];
  const double gradT_parl = bgradT;
  const double gradT_perp = sqrt(perp[0] * perp[0] + perp[1] * perp[1] + perp[2] * perp[2]);
  return -1. / (1. / (Kparl + 1.e-14) + gradT_parl / (coefSat + 1.e-14)) * parl[D]
         - 1. / (1. / (Kperp + 1.e-14) + gradT_perp / (coefSat + 1.e-14)) * perp[D];
}

// one substep over the physical cells (:749-757), fluxes evaluated in place of f,g,h(5,...)
__global__ void __launch_bounds__(128) k_tc_update(gx::Grid g, gxp::Phys P, TcPar t, const double* __restrict__ PT, double* __restrict__ U, double dts) {
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x) + 1, j = (int)blockIdx.y + 1, k = (int)blockIdx.z + 1;
  if (i > g.nx) return;
  const long long c = g.idx(i, j, k), vs = g.vs, sy = g.px, sz = (long long)g.px * g.py;
  double fc, fm, gc, gm, hc, hm;
  if (t.mode == GX_TC_ISOTROPIC) {
    const double Tc = PT[vs + c], pc = PT[c], rc = fmax(U[c], 1e-15);
    auto lo = [&](long long n, double dxr) { return flux_iso(t, P, PT[vs + n], Tc, PT[n], pc, fmax(U[n], 1e-15), rc, dxr); };
    auto hi = [&](long long n, double dxr) { return flux_iso(t, P, Tc, PT[vs + n], pc, PT[n], rc, fmax(U[n], 1e-15), dxr); };
    fc = hi(c + 1, t.dxr); fm = lo(c - 1, t.dxr);
    gc = hi(c + sy, t.dyr); gm = lo(c - sy, t.dyr);
    hc = hi(c + sz, t.dzr); hm = lo(c - sz, t.dzr);
  } else {
    fc = flux_aniso<0>(t, P, g, U, PT, c); fm = flux_aniso<0>(t, P, g, U, PT, c - 1);
    gc = flux_aniso<1>(t, P, g, U, PT, c); gm = flux_aniso<1>(t, P, g, U, PT, c - sy);
    hc = flux_aniso<2>(t, P, g, U, PT, c); hm = flux_aniso<2>(t, P, g, U, PT, c - sz);
  }
  U[4 * vs + c] = U[4 * vs + c] - dts * ((fc - fm) / t.dx + (gc - gm) / t.dy + (hc - hm) / t.dz);
}

// ---- host side of thermal_conduction (:625-681): super-time-stepping schedule ----
// integer powers are gfortran's __builtin_powi (binary exponentiation)
inline double powi(double x, int m) {
  unsigned n = m < 0 ? (unsigned)(-m) : (unsigned)m;
  double y = (n % 2) ? x : 1.0;
  while (n >>= 1) { x = x * x; if (n % 2) y *= x; }
  return m < 0 ? 1.0 / y : y;
}
inline double superstep(int N, double snu) {
  return (double)N / (2. * snu) * (powi(1 + snu, 2 * N) - powi(1 - snu, 2 * N)) / (powi(1 + snu, 2 * N) + powi(1 - snu, 2 * N));
}
inline double substep(int j, int N, double nu) {
  const double pi = acos(-1.);
  return 1. / ((nu - 1.) * cos(pi * (double)(2 * j - 1) / (2. * (double)N)) + nu + 1.);
}
inline void ST_steps(double fs, int& Ns, double& fstep) {
  const double snu = sqrt(0.01);
  int j;
  for (j = 1; j <= 199; ++j) if (superstep(j, snu) > fs) break;
  Ns = j;
  fstep = fs / superstep(Ns, snu);
}

}  // namespace gxtc
