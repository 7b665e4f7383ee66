// gx_kernels.cu — hydro/MHD step kernels for sm_100a (FP64 stencils, SoA layout).
// Compiled twice: -DGX_FLAVOUR_STRICT with -fmad=false (bit-comparison build) and
// -DGX_FLAVOUR_FAST with -fmad=true.
//
// Reference passes covered (file:line in the reference tree):
//   k_calcprim  : calcprim/u2prim  src/hydro_core.f90:245-320, 46-129  (+ CFL candidates :644-675)
//   k_flux      : hll?fluxes(choice) sweeps  src/hlld.f90:331-432 (hll/hllc/hlle identical)
//   k_efield    : get_efield  src/flux_cd_module.f90:245-273
//   k_update    : step + flux_cd_update + source  src/hydro_solver.f90:77-127,
//                 src/flux_cd_module.f90:285-323, src/sources.f90:124-220
//   k_viscous   : viscous_copy  src/hydro_solver.f90:47-65
#include <algorithm>
#include "gx_kernels.cuh"
#include "gx_split.cuh"

#if defined(GX_FLAVOUR_STRICT)
#define GX_NS strict_ns
#elif defined(GX_FLAVOUR_FAST)
#define GX_NS fast_ns
#else
#error "define GX_FLAVOUR_STRICT or GX_FLAVOUR_FAST"
#endif

namespace gx {
namespace GX_NS {

using gxp::Phys;

// ---------------------------------------------------------------------------
// block-wide min of positive doubles -> one atomicMin on the ordered bit pattern.
// INVERSE: the threads hold 1/dt candidates (max is taken, one reciprocal per CTA).
template <bool INVERSE = false>
__device__ __forceinline__ void block_atomic_min(double v, unsigned long long* dst) {
  auto red = [](double a, double b) { return INVERSE ? fmax(a, b) : fmin(a, b); };
  for (int o = 16; o > 0; o >>= 1) v = red(v, __shfl_xor_sync(0xffffffffu, v, o));
  __shared__ double smin[32];
  const int tl = (int)((threadIdx.z * blockDim.y + threadIdx.y) * blockDim.x + threadIdx.x);   // linear thread id (the flux-CD kernel uses 3-D CTAs)
  const int lane = tl & 31, wid = tl >> 5;
  if (lane == 0) smin[wid] = v;
  __syncthreads();
  if (wid == 0) {
    const int nw = (int)((blockDim.x * blockDim.y * blockDim.z + 31) >> 5);
    v = lane < nw ? smin[lane] : (INVERSE ? 0.0 : 1.e30);
    for (int o = 16; o > 0; o >>= 1) v = red(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (INVERSE) v = v > 0.0 ? 1.0 / v : 1.e30;
    if (lane == 0) atomicMin(dst, (unsigned long long)__double_as_longlong(v));   // v > 0: bit order == value order
  }
}

// ---------------------------------------------------------------------------
template <bool MHD>
__global__ void __launch_bounds__(128) k_calcprim(const StepArgs A, const double* __restrict__ U, double* __restrict__ W,
                                                  double* __restrict__ Temp, unsigned long long* dtmin_bits, int want_cfl) {
  const Grid& g = A.g;
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x) - 1;   // Fortran i = -1 .. nx+2
  const int j = (int)blockIdx.y - 1, k = (int)blockIdx.z - 1;
  double dtp = 1.e30;
  if (i <= g.nx + 2) {
    const long long c = g.idx(i, j, k);
    double u[8], w[8], T;
#pragma unroll
    for (int q = 0; q < (MHD ? 8 : 5); ++q) u[q] = U[q * g.vs + c];
    const double pas0 = g.npas > 0 ? U[(long long)g.neqdyn * g.vs + c] : 0.0;
    gxp::u2prim<MHD>(A.phys, u, w, pas0, T);
    if (W) {
#pragma unroll
      for (int q = 0; q < (MHD ? 8 : 5); ++q) W[q * g.vs + c] = w[q];
      for (int q = g.neqdyn; q < g.neq; ++q) W[q * g.vs + c] = U[q * g.vs + c];
    }
    if (Temp) Temp[c] = T;
    if (want_cfl && i >= 1 && i <= g.nx && j >= 1 && j <= g.ny && k >= 1 && k <= g.nz) {
      if (MHD) {
        double cx, cy, cz;
        gxp::cfast3(A.phys, w[4], w[0], w[5], w[6], w[7], cx, cy, cz);
        dtp = fmin(dtp, g.dx / (fabs(w[1]) + cx));
        dtp = fmin(dtp, g.dy / (fabs(w[2]) + cy));
        dtp = fmin(dtp, g.dz / (fabs(w[3]) + cz));
      } else {
        double cs = gxp::csound(A.phys, w[4], w[0]);
        dtp = fmin(dtp, g.dx / (fabs(w[1]) + cs));
        dtp = fmin(dtp, g.dy / (fabs(w[2]) + cs));
        dtp = fmin(dtp, g.dz / (fabs(w[3]) + cs));
      }
    }
  }
  if (want_cfl) block_atomic_min(dtp, dtmin_bits);
}

// calcprim of the split-all solver: u2primSplitAll (src/hydro_core.f90:143-229, called at :263-309) and the CFL candidates of
// the TOTAL state (:650-654).  u, W: fluctuations; W0: background primitives.
__global__ void __launch_bounds__(128) k_calcprim_split(const StepArgs A, const double* __restrict__ U, double* __restrict__ W,
                                                        double* __restrict__ Temp, unsigned long long* dtmin_bits, int want_cfl) {
  const Grid& g = A.g;
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x) - 1;
  const int j = (int)blockIdx.y - 1, k = (int)blockIdx.z - 1;
  double dtp = 1.e30;
  if (i <= g.nx + 2) {
    const long long c = g.idx(i, j, k);
    double u[8], w0[8], w[8], T;
#pragma unroll
    for (int q = 0; q < 8; ++q) { u[q] = U[q * g.vs + c]; w0[q] = A.W0[q * g.vs + c]; }
    gxp::u2prim_split(A.phys, u, w0, w, T);
    if (W) {
#pragma unroll
      for (int q = 0; q < 8; ++q) W[q * g.vs + c] = w[q];
    }
    if (Temp) Temp[c] = T;
    if (want_cfl && i >= 1 && i <= g.nx && j >= 1 && j <= g.ny && k >= 1 && k <= g.nz) {
      double cx, cy, cz;
      gxp::cfast3(A.phys, w[4] + w0[4], w[0] + w0[0], w[5] + w0[5], w[6] + w0[6], w[7] + w0[7], cx, cy, cz);
      dtp = fmin(dtp, g.dx / (fabs(w[1]) + cx));
      dtp = fmin(dtp, g.dy / (fabs(w[2]) + cy));
      dtp = fmin(dtp, g.dz / (fabs(w[3]) + cz));
    }
  }
  if (want_cfl) block_atomic_min(dtp, dtmin_bits);
}

// ---------------------------------------------------------------------------
// storage component of rotated slot c for sweep direction D (swapy / swapz as an index map)
template <int D> __device__ __forceinline__ constexpr int comp(int c) {
  return (c == 1) ? 1 + D : (c == 1 + D) ? 1 : (c == 5) ? 5 + D : (c == 5 + D) ? 5 : c;
}

// One thread = one interface (upper face of cell i,j,k in direction D).
// Faces the update never reads are skipped.
#ifndef GX_FLUX_MINBLOCKS
#define GX_FLUX_MINBLOCKS 4
#endif
template <int SOLVER, int LIM, int ORDER, int D>
__global__ void __launch_bounds__(128, GX_FLUX_MINBLOCKS) k_flux(const StepArgs A, const double* __restrict__ W, double* __restrict__ F, int* errflag) {
  constexpr bool MHD = (SOLVER == GX_SOLVER_HLLE || SOLVER == GX_SOLVER_HLLD);
  constexpr int NQ = MHD ? 8 : 5;
  const Grid& g = A.g;
  // face index ranges: normal index 0..n, transverse 1..n
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x) + (D == 0 ? 0 : 1);
  const int j = (int)blockIdx.y + (D == 1 ? 0 : 1);
  const int k = (int)blockIdx.z + (D == 2 ? 0 : 1);
  if (i > g.nx) return;
  const long long st = (D == 0) ? 1 : (D == 1) ? (long long)g.px : (long long)g.px * g.py;
  const long long c = g.idx(i, j, k);
  double wl[8], wr[8], ff[8];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const double* Wq = W + (long long)comp<D>(q) * g.vs + c;
    double pl = Wq[0], pr = Wq[st];
    if (ORDER == 2) gxp::reconstruct<LIM>(Wq[-st], pl, pr, Wq[2 * st]);
    wl[q] = pl; wr[q] = pr;
  }
  gxp::PasInfo I;
  int err = gxp::riemann<SOLVER>(A.phys, wl, wr, ff, I);
  if (err) atomicOr(errflag, 1);
  double* Fd = F + (long long)D * g.neq * g.vs + c;
#pragma unroll
  for (int q = 0; q < NQ; ++q) Fd[(long long)comp<D>(q) * g.vs] = ff[q];
  for (int q = g.neqdyn; q < g.neq; ++q) {           // passive scalars (limited like every primitive)
    const double* Wq = W + (long long)q * g.vs + c;
    double pl = Wq[0], pr = Wq[st];
    if (ORDER == 2) gxp::reconstruct<LIM>(Wq[-st], pl, pr, Wq[2 * st]);
    Fd[(long long)q * g.vs] = gxp::passive_flux(I, pl, pr);
  }
}

// hllEfluxesSplitAll(choice) (src/hlle_split_all.f90:97-241): the sweep of k_flux with the background states alongside; the
// limiter acts on the fluctuations and on the background separately (:154-155).
template <int LIM, int ORDER, int D>
__global__ void __launch_bounds__(128) k_flux_split(const StepArgs A, const double* __restrict__ W, double* __restrict__ F) {
  const Grid& g = A.g;
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x) + (D == 0 ? 0 : 1);
  const int j = (int)blockIdx.y + (D == 1 ? 0 : 1);
  const int k = (int)blockIdx.z + (D == 2 ? 0 : 1);
  if (i > g.nx) return;
  const long long st = (D == 0) ? 1 : (D == 1) ? (long long)g.px : (long long)g.px * g.py;
  const long long c = g.idx(i, j, k);
  double wl[8], wr[8], w0l[8], w0r[8], ff[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const double* Wq = W + (long long)comp<D>(q) * g.vs + c;
    const double* Bq = A.W0 + (long long)comp<D>(q) * g.vs + c;
    double pl = Wq[0], pr = Wq[st], bl = Bq[0], br = Bq[st];
    if (ORDER == 2) { gxp::reconstruct<LIM>(Wq[-st], pl, pr, Wq[2 * st]); gxp::reconstruct<LIM>(Bq[-st], bl, br, Bq[2 * st]); }
    wl[q] = pl; wr[q] = pr; w0l[q] = bl; w0r[q] = br;
  }
  gxp::riemann_hlle_split_all(A.phys, wl, wr, w0l, w0r, ff);
  double* Fd = F + (long long)D * g.neq * g.vs + c;
#pragma unroll
  for (int q = 0; q < 8; ++q) Fd[(long long)comp<D>(q) * g.vs] = ff[q];
}

// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_efield(const StepArgs A, const double* __restrict__ F, double* __restrict__ E) {
  const Grid& g = A.g;
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x) + 1, j = (int)blockIdx.y + 1, k = (int)blockIdx.z + 1;
  if (i > g.nx) return;
  const long long c = g.idx(i, j, k), sy = g.px, sz = (long long)g.px * g.py, vs = g.vs;
  const double* f = F;
  const double* gg = F + (long long)g.neq * vs;
  const double* h = F + 2LL * g.neq * vs;
  // e(1) = 1/4(-g8(j-1) - g8(j) + h7(k-1) + h7(k)) etc.  (components are 0-based here: B = 5,6,7)
  E[0 * vs + c] = 0.25 * (-gg[7 * vs + c - sy] - gg[7 * vs + c] + h[6 * vs + c - sz] + h[6 * vs + c]);
  E[1 * vs + c] = 0.25 * (+f[7 * vs + c - 1] + f[7 * vs + c] - h[5 * vs + c - sz] - h[5 * vs + c]);
  E[2 * vs + c] = 0.25 * (-f[6 * vs + c - 1] - f[6 * vs + c] + gg[5 * vs + c - sy] + gg[5 * vs + c]);
}

// ---------------------------------------------------------------------------
template <bool FLUXCD>
__global__ void __launch_bounds__(128) k_update(const StepArgs A, double dt, const double* __restrict__ U, const double* __restrict__ F,
                                                const double* __restrict__ E, const double* __restrict__ W, double* __restrict__ dst) {
  const Grid& g = A.g;
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x) + 1, j = (int)blockIdx.y + 1, k = (int)blockIdx.z + 1;
  if (i > g.nx) return;
  const long long c = g.idx(i, j, k), sy = g.px, sz = (long long)g.px * g.py, vs = g.vs;
  const double dtdx = dt / g.dx, dtdy = dt / g.dy, dtdz = dt / g.dz;
  const double* f = F;
  const double* gg = F + (long long)g.neq * vs;
  const double* h = F + 2LL * g.neq * vs;
  const bool src = A.eight_wave || (A.user_src && A.grav.n > 0);
  double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (src) {
    // source(i,j,k,primit(:,i,j,k),s): src/sources.f90:190-220
    double pp[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) pp[q] = (q < g.neqdyn) ? W[q * vs + c] : 0.0;
    if (A.user_src) {
      const double xc = ((double)(i + g.cx * g.nx - g.nxtot / 2) - 0.5) * g.dx;
      const double yc = ((double)(j + g.cy * g.ny - g.nytot / 2) - 0.5) * g.dy;
      const double zc = ((double)(k + g.cz * g.nz - g.nztot / 2) - 0.5) * g.dz;
      for (int l = 0; l < A.grav.n; ++l) {
        const double x = xc - A.grav.x[l], y = yc - A.grav.y[l], z = zc - A.grav.z[l];
        const double rad2 = x * x + y * y + z * z;
        const double r15 = pow(rad2, 1.5);
        s[1] = s[1] - pp[0] * A.grav.gm[l] * x / r15;
        s[2] = s[2] - pp[0] * A.grav.gm[l] * y / r15;
        s[3] = s[3] - pp[0] * A.grav.gm[l] * z / r15;
        s[4] = s[4] - pp[0] * A.grav.gm[l] * (pp[1] * x + pp[2] * y + pp[3] * z) / r15;
      }
    }
    if (A.eight_wave && g.neqdyn == 8) {
      const double d = (W[5 * vs + c + 1] - W[5 * vs + c - 1]) / (2. * g.dx)
                     + (W[6 * vs + c + sy] - W[6 * vs + c - sy]) / (2. * g.dy)
                     + (W[7 * vs + c + sz] - W[7 * vs + c - sz]) / (2. * g.dz);
      s[1] = s[1] - d * pp[5];
      s[2] = s[2] - d * pp[6];
      s[3] = s[3] - d * pp[7];
      s[4] = s[4] - d * (pp[1] * pp[5] + pp[2] * pp[6] + pp[3] * pp[7]);
      s[5] = s[5] - d * pp[1];
      s[6] = s[6] - d * pp[2];
      s[7] = s[7] - d * pp[3];
    }
  }
  for (int q = 0; q < g.neq; ++q) {
    const long long o = q * vs + c;
    double v;
    if (FLUXCD && q >= 5 && q < 8) {
      // evolution of B with flux-CD: src/flux_cd_module.f90:311-321
      if (q == 5) v = U[o] - 0.5 * dtdy * (E[2 * vs + c + sy] - E[2 * vs + c - sy]) + 0.5 * dtdz * (E[1 * vs + c + sz] - E[1 * vs + c - sz]);
      else if (q == 6) v = U[o] + 0.5 * dtdx * (E[2 * vs + c + 1] - E[2 * vs + c - 1]) - 0.5 * dtdz * (E[0 * vs + c + sz] - E[0 * vs + c - sz]);
      else v = U[o] - 0.5 * dtdx * (E[1 * vs + c + 1] - E[1 * vs + c - 1]) + 0.5 * dtdy * (E[0 * vs + c + sy] - E[0 * vs + c - sy]);
    } else {
      v = U[o] - dtdx * (f[o] - f[o - 1]) - dtdy * (gg[o] - gg[o - sy]) - dtdz * (h[o] - h[o - sz]);
    }
    if (src) v = v + dt * (q < 8 ? s[q] : 0.0);
    dst[o] = v;
  }
}

// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_viscous(const StepArgs A, double eta, const double* __restrict__ UP, double* __restrict__ U) {
  const Grid& g = A.g;
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x) + 1, j = (int)blockIdx.y + 1, k = (int)blockIdx.z + 1;
  if (i > g.nx) return;
  const long long c = g.idx(i, j, k), sy = g.px, sz = (long long)g.px * g.py;
  for (int q = 0; q < g.neq; ++q) {
    const double* p = UP + q * g.vs + c;
    U[q * g.vs + c] = p[0] + eta * (p[1] + p[-1] + p[sy] + p[-sy] + p[sz] + p[-sz] - 6. * p[0]);
  }
}

// viscous_copy when the full-step state was written to T instead of up (fused path): the reference reads up(i+-1, ...) whose ghost
// cells still hold the half-step halo of boundaryII — a neighbour outside the block is therefore read from UP
__global__ void __launch_bounds__(128) k_viscous2(const StepArgs A, double eta, const double* __restrict__ T, const double* __restrict__ UP, double* __restrict__ U) {
  const Grid& g = A.g;
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x) + 1, j = (int)blockIdx.y + 1, k = (int)blockIdx.z + 1;
  if (i > g.nx) return;
  const long long c = g.idx(i, j, k), sy = g.px, sz = (long long)g.px * g.py;
  for (int q = 0; q < g.neq; ++q) {
    const double* p = T + q * g.vs + c;
    const double* h = UP + q * g.vs + c;
    const double xp = i < g.nx ? p[1] : h[1], xm = i > 1 ? p[-1] : h[-1];
    const double yp = j < g.ny ? p[sy] : h[sy], ym = j > 1 ? p[-sy] : h[-sy];
    const double zp = k < g.nz ? p[sz] : h[sz], zm = k > 1 ? p[-sz] : h[-sz];
    U[q * g.vs + c] = p[0] + eta * (xp + xm + yp + ym + zp + zm - 6. * p[0]);
  }
}

// ---------------------------------------------------------------------------
// flux-CD evolution of B from the cell-centred E (flux_cd_update, src/flux_cd_module.f90:311-321),
// the companion of the fused stage kernel; with want_cfl also the CFL candidates of the
// finished state (get_timestep, src/hydro_core.f90:644-675).
// A CTA is a 32 x BUY x BUZ brick of cells: the y and z neighbours of the E stencil are then read by threads of the
// same CTA and hit L1 (read-only path) instead of being fetched again from L2 — the kernel is bound by the L2 -> SM
// traffic of the twelve stencil reads per cell, not by DRAM (72 / 112 algorithmic bytes per cell).
#ifndef GX_BU_Y
#define GX_BU_Y 4
#endif
#ifndef GX_BU_Z
#define GX_BU_Z 2
#endif
// One thread = TWO x-adjacent cells (i, i+1), i odd: every centre / y / z neighbour access is one aligned 16-byte load (rows start
// on a 128-byte line and i = 1 sits on an even element), the x neighbours i-1 and i+2 are two scalar loads: half the load
// instructions and twice the bytes per request of the one-cell form.
template <bool CFL>
__global__ void __launch_bounds__(32 * GX_BU_Y * GX_BU_Z) k_bupdate(const StepArgs A, const double dtdx, const double dtdy, const double dtdz,
                                                                    const double* Ub, const double* __restrict__ E, double* dst, unsigned long long* dtmin_bits) {
  const Grid& g = A.g;
  const int i = 2 * (int)(blockIdx.x * 32 + threadIdx.x) + 1, j = (int)(blockIdx.y * GX_BU_Y + threadIdx.y) + 1;
  const int nz1 = (A.klast - A.kbeg + GX_BU_Z) / GX_BU_Z;              // z blocks of the first plane range; then the second (see StepArgs)
  const bool second = (int)blockIdx.z >= nz1;
  const int k = second ? (int)((blockIdx.z - nz1) * GX_BU_Z + threadIdx.z) + A.kbeg2 : (int)(blockIdx.z * GX_BU_Z + threadIdx.z) + A.kbeg;
  const int klast = second ? A.klast2 : A.klast;
#if defined(GX_FLAVOUR_FAST)
  double inv_dtp = 0.0;                           // max over cells of (|v| + c) / dx: one reciprocal per CTA instead of three divisions per cell
#else
  double dtp = 1.e30;
#endif
  if (i <= g.nx && j <= g.ny && k <= klast) {
    const bool two = i + 1 <= g.nx;               // (odd nx: the last thread of a row owns one cell; its second lane of work is discarded)
    const long long c = g.idx(i, j, k), sy = g.px, sz = (long long)g.px * g.py, vs = g.vs;
    // neighbours of the E stencil; a self-periodic direction wraps instead of reading a ghost cell
    const long long xm = (A.wrap[0] && i == 1) ? c + (g.nx - 1) : c - 1;                         // left of cell i
    const long long xp = (A.wrap[0] && i + 1 >= g.nx) ? c + 1 - (g.nx - 1) : c + 2;              // right of cell i+1
    const long long ym = (A.wrap[1] && j == 1) ? c + (g.ny - 1) * sy : c - sy, yp = (A.wrap[1] && j == g.ny) ? c - (g.ny - 1) * sy : c + sy;
    const long long zm = (A.wrap[2] && k == 1) ? c + (g.nz - 1) * sz : c - sz, zp = (A.wrap[2] && k == g.nz) ? c - (g.nz - 1) * sz : c + sz;
    auto ld2 = [](const double* p) { return *reinterpret_cast<const double2*>(p); };
    const double2 b5 = ld2(Ub + 5 * vs + c), b6 = ld2(Ub + 6 * vs + c), b7 = ld2(Ub + 7 * vs + c);
    const double2 e2yp = ld2(E + 2 * vs + yp), e2ym = ld2(E + 2 * vs + ym), e1zp = ld2(E + 1 * vs + zp), e1zm = ld2(E + 1 * vs + zm);
    const double2 e0zp = ld2(E + 0 * vs + zp), e0zm = ld2(E + 0 * vs + zm), e0yp = ld2(E + 0 * vs + yp), e0ym = ld2(E + 0 * vs + ym);
    double2 e2c = ld2(E + 2 * vs + c), e1c = ld2(E + 1 * vs + c);
    const double e2l = E[2 * vs + xm], e2r = E[2 * vs + xp], e1l = E[1 * vs + xm], e1r = E[1 * vs + xp];
    if (!two && A.wrap[0]) {                      // odd nx, last cell of a wrapped row: its right neighbour is cell 1, not the ghost cell
      e2c.y = E[2 * vs + c - (g.nx - 1)];
      e1c.y = E[1 * vs + c - (g.nx - 1)];
    }
    // flux_cd_update, src/flux_cd_module.f90:311-321 (cell i: .x, cell i+1: .y); x differences: (E(i+1) - E(i-1)) and (E(i+2) - E(i))
    double2 bx, by, bz;
    bx.x = b5.x - 0.5 * dtdy * (e2yp.x - e2ym.x) + 0.5 * dtdz * (e1zp.x - e1zm.x);
    bx.y = b5.y - 0.5 * dtdy * (e2yp.y - e2ym.y) + 0.5 * dtdz * (e1zp.y - e1zm.y);
    by.x = b6.x + 0.5 * dtdx * (e2c.y - e2l) - 0.5 * dtdz * (e0zp.x - e0zm.x);
    by.y = b6.y + 0.5 * dtdx * (e2r - e2c.x) - 0.5 * dtdz * (e0zp.y - e0zm.y);
    bz.x = b7.x - 0.5 * dtdx * (e1c.y - e1l) + 0.5 * dtdy * (e0yp.x - e0ym.x);
    bz.y = b7.y - 0.5 * dtdx * (e1r - e1c.x) + 0.5 * dtdy * (e0yp.y - e0ym.y);
    if (two) {
      *reinterpret_cast<double2*>(dst + 5 * vs + c) = bx; *reinterpret_cast<double2*>(dst + 6 * vs + c) = by; *reinterpret_cast<double2*>(dst + 7 * vs + c) = bz;
    } else {
      dst[5 * vs + c] = bx.x; dst[6 * vs + c] = by.x; dst[7 * vs + c] = bz.x;
    }
    if (CFL) {
      double2 d5[5];
#pragma unroll
      for (int q = 0; q < 5; ++q) d5[q] = ld2(dst + q * vs + c);
      const double2 pas = g.npas > 0 ? ld2(dst + 8 * vs + c) : make_double2(0.0, 0.0);          // EOS_H_RATE reads the first passive
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        if (m == 1 && !two) break;
        double u[8], w[8], T;
#pragma unroll
        for (int q = 0; q < 5; ++q) u[q] = m ? d5[q].y : d5[q].x;
        u[5] = m ? bx.y : bx.x; u[6] = m ? by.y : by.x; u[7] = m ? bz.y : bz.x;
        gxp::u2prim<true>(A.phys, u, w, m ? pas.y : pas.x, T);
        double cx, cy, cz;
        gxp::cfast3(A.phys, w[4], w[0], w[5], w[6], w[7], cx, cy, cz);
#if defined(GX_FLAVOUR_FAST)
        inv_dtp = gxp::gx_max(inv_dtp, (fabs(w[1]) + cx) * A.idx3[0]);
        inv_dtp = gxp::gx_max(inv_dtp, (fabs(w[2]) + cy) * A.idx3[1]);
        inv_dtp = gxp::gx_max(inv_dtp, (fabs(w[3]) + cz) * A.idx3[2]);
#else
        dtp = fmin(dtp, g.dx / (fabs(w[1]) + cx));
        dtp = fmin(dtp, g.dy / (fabs(w[2]) + cy));
        dtp = fmin(dtp, g.dz / (fabs(w[3]) + cz));
#endif
      }
    }
  }
  if (CFL) {
#if defined(GX_FLAVOUR_FAST)
    block_atomic_min<true>(inv_dtp, dtmin_bits);
#else
    block_atomic_min(dtp, dtmin_bits);
#endif
  }
}

// ---------------------------------------------------------------------------
// Per-interface Riemann flux for n independent state pairs (prim2fhll / prim2fhllc / prim2fhlle / prim2fhlld called
// directly: src/hll.f90:47, src/hllc.f90:44, src/hlle.f90:48, src/hlld.f90:48): the SAME device functions the sweeps use,
// on caller-supplied (rotated) primitive states.  wl, wr, ff: [n][8] (hydro solvers use the first 5 slots).
template <int SOLVER>
__global__ void __launch_bounds__(128) k_riemann_points(const Phys P, int n, const double* __restrict__ wl, const double* __restrict__ wr,
                                                        double* __restrict__ ff, int* __restrict__ err) {
  const int t = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (t >= n) return;
  double l[8], r[8], f[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) { l[q] = wl[8 * t + q]; r[q] = wr[8 * t + q]; f[q] = 0.0; }
  gxp::PasInfo I;
  err[t] = gxp::riemann<SOLVER>(P, l, r, f, I);
#pragma unroll
  for (int q = 0; q < 8; ++q) ff[8 * t + q] = f[q];
}
static int l_riemann_points(const Phys& P, int solver, int n, const double* wl, const double* wr, double* ff, int* err, cudaStream_t s) {
  const int nb = (n + 127) / 128;
  switch (solver) {
    case GX_SOLVER_HLL:  k_riemann_points<GX_SOLVER_HLL><<<nb, 128, 0, s>>>(P, n, wl, wr, ff, err); return 0;
    case GX_SOLVER_HLLC: k_riemann_points<GX_SOLVER_HLLC><<<nb, 128, 0, s>>>(P, n, wl, wr, ff, err); return 0;
    case GX_SOLVER_HLLE: k_riemann_points<GX_SOLVER_HLLE><<<nb, 128, 0, s>>>(P, n, wl, wr, ff, err); return 0;
    case GX_SOLVER_HLLD: k_riemann_points<GX_SOLVER_HLLD><<<nb, 128, 0, s>>>(P, n, wl, wr, ff, err); return 0;
  }
  return GX_EUNSUPPORTED;
}

// ---------------------------------------------------------------------------
#include "gx_thermal.cuh"          // thermal conduction kernels, compiled in this flavour
#include "gx_cooling.cuh"          // COOL_H

// ---------------------------------------------------------------------------
// launchers
static inline dim3 grid_for(int nxr, int nyr, int nzr, int bx) { return dim3((unsigned)((nxr + bx - 1) / bx), (unsigned)nyr, (unsigned)nzr); }

static void l_calcprim(const StepArgs& A, const double* U, double* W, double* Temp, unsigned long long* dtmin_bits, int want_cfl, cudaStream_t s) {
  const Grid& g = A.g;
  dim3 grid = grid_for(g.nx + 4, g.ny + 4, g.nz + 4, 128);
  if (A.solver == GX_SOLVER_HLLE_SPLIT_ALL) k_calcprim_split<<<grid, 128, 0, s>>>(A, U, W, Temp, dtmin_bits, want_cfl);
  else if (A.phys.neqdyn == 8) k_calcprim<true><<<grid, 128, 0, s>>>(A, U, W, Temp, dtmin_bits, want_cfl);
  else k_calcprim<false><<<grid, 128, 0, s>>>(A, U, W, Temp, dtmin_bits, want_cfl);
}

template <int SOLVER, int LIM, int ORDER>
static void l_flux3(const StepArgs& A, const double* W, double* F, int* errflag, cudaStream_t s) {
  const Grid& g = A.g;
  k_flux<SOLVER, LIM, ORDER, 0><<<grid_for(g.nx + 1, g.ny, g.nz, 128), 128, 0, s>>>(A, W, F, errflag);
  k_flux<SOLVER, LIM, ORDER, 1><<<grid_for(g.nx, g.ny + 1, g.nz, 128), 128, 0, s>>>(A, W, F, errflag);
  k_flux<SOLVER, LIM, ORDER, 2><<<grid_for(g.nx, g.ny, g.nz + 1, 128), 128, 0, s>>>(A, W, F, errflag);
}

template <int SOLVER>
static int l_flux_solver(const StepArgs& A, int order, const double* W, double* F, int* errflag, cudaStream_t s) {
  if (order == 1) { l_flux3<SOLVER, GX_LIMITER_NO_AVERAGE, 1>(A, W, F, errflag, s); return 0; }
  switch (A.limiter) {
    case GX_LIMITER_NO_AVERAGE: l_flux3<SOLVER, GX_LIMITER_NO_AVERAGE, 2>(A, W, F, errflag, s); return 0;
    case GX_LIMITER_NO_LIMIT:   l_flux3<SOLVER, GX_LIMITER_NO_LIMIT, 2>(A, W, F, errflag, s); return 0;
    case GX_LIMITER_MINMOD:     l_flux3<SOLVER, GX_LIMITER_MINMOD, 2>(A, W, F, errflag, s); return 0;
    case GX_LIMITER_VAN_LEER:   l_flux3<SOLVER, GX_LIMITER_VAN_LEER, 2>(A, W, F, errflag, s); return 0;
    case GX_LIMITER_VAN_ALBADA: l_flux3<SOLVER, GX_LIMITER_VAN_ALBADA, 2>(A, W, F, errflag, s); return 0;
    case GX_LIMITER_UMIST:      l_flux3<SOLVER, GX_LIMITER_UMIST, 2>(A, W, F, errflag, s); return 0;
    case GX_LIMITER_WOODWARD:   l_flux3<SOLVER, GX_LIMITER_WOODWARD, 2>(A, W, F, errflag, s); return 0;
    case GX_LIMITER_SUPERBEE:   l_flux3<SOLVER, GX_LIMITER_SUPERBEE, 2>(A, W, F, errflag, s); return 0;
  }
  return GX_EINVAL;
}

template <int LIM, int ORDER>
static void l_flux3_split(const StepArgs& A, const double* W, double* F, cudaStream_t s) {
  const Grid& g = A.g;
  k_flux_split<LIM, ORDER, 0><<<grid_for(g.nx + 1, g.ny, g.nz, 128), 128, 0, s>>>(A, W, F);
  k_flux_split<LIM, ORDER, 1><<<grid_for(g.nx, g.ny + 1, g.nz, 128), 128, 0, s>>>(A, W, F);
  k_flux_split<LIM, ORDER, 2><<<grid_for(g.nx, g.ny, g.nz + 1, 128), 128, 0, s>>>(A, W, F);
}
static int l_flux_split(const StepArgs& A, int order, const double* W, double* F, cudaStream_t s) {
  if (!A.W0) return GX_ESTATE;
  if (order == 1) { l_flux3_split<GX_LIMITER_NO_AVERAGE, 1>(A, W, F, s); return 0; }
  switch (A.limiter) {
    case GX_LIMITER_NO_AVERAGE: l_flux3_split<GX_LIMITER_NO_AVERAGE, 2>(A, W, F, s); return 0;
    case GX_LIMITER_NO_LIMIT:   l_flux3_split<GX_LIMITER_NO_LIMIT, 2>(A, W, F, s); return 0;
    case GX_LIMITER_MINMOD:     l_flux3_split<GX_LIMITER_MINMOD, 2>(A, W, F, s); return 0;
    case GX_LIMITER_VAN_LEER:   l_flux3_split<GX_LIMITER_VAN_LEER, 2>(A, W, F, s); return 0;
    case GX_LIMITER_VAN_ALBADA: l_flux3_split<GX_LIMITER_VAN_ALBADA, 2>(A, W, F, s); return 0;
    case GX_LIMITER_UMIST:      l_flux3_split<GX_LIMITER_UMIST, 2>(A, W, F, s); return 0;
    case GX_LIMITER_WOODWARD:   l_flux3_split<GX_LIMITER_WOODWARD, 2>(A, W, F, s); return 0;
    case GX_LIMITER_SUPERBEE:   l_flux3_split<GX_LIMITER_SUPERBEE, 2>(A, W, F, s); return 0;
  }
  return GX_EINVAL;
}

static int l_fluxes(const StepArgs& A, int order, const double* W, double* F, int* errflag, cudaStream_t s) {
  switch (A.solver) {
    case GX_SOLVER_HLLE_SPLIT_ALL: return l_flux_split(A, order, W, F, s);
    case GX_SOLVER_HLL:  return l_flux_solver<GX_SOLVER_HLL>(A, order, W, F, errflag, s);
    case GX_SOLVER_HLLC: return l_flux_solver<GX_SOLVER_HLLC>(A, order, W, F, errflag, s);
    case GX_SOLVER_HLLE: return l_flux_solver<GX_SOLVER_HLLE>(A, order, W, F, errflag, s);
    case GX_SOLVER_HLLD: return l_flux_solver<GX_SOLVER_HLLD>(A, order, W, F, errflag, s);
  }
  return GX_EUNSUPPORTED;
}

static void l_efield(const StepArgs& A, const double* F, double* E, cudaStream_t s) {
  const Grid& g = A.g;
  k_efield<<<grid_for(g.nx, g.ny, g.nz, 128), 128, 0, s>>>(A, F, E);
}

static void l_update(const StepArgs& A, double dt, const double* U, const double* F, const double* E, const double* W, double* dst, cudaStream_t s) {
  const Grid& g = A.g;
  dim3 grid = grid_for(g.nx, g.ny, g.nz, 128);
  if (A.flux_cd) k_update<true><<<grid, 128, 0, s>>>(A, dt, U, F, E, W, dst);
  else k_update<false><<<grid, 128, 0, s>>>(A, dt, U, F, E, W, dst);
}

static void l_viscous(const StepArgs& A, double eta, const double* UP, double* U, cudaStream_t s) {
  const Grid& g = A.g;
  k_viscous<<<grid_for(g.nx, g.ny, g.nz, 128), 128, 0, s>>>(A, eta, UP, U);
}

static void l_viscous2(const StepArgs& A, double eta, const double* T, const double* UP, double* U, cudaStream_t s) {
  const Grid& g = A.g;
  k_viscous2<<<grid_for(g.nx, g.ny, g.nz, 128), 128, 0, s>>>(A, eta, T, UP, U);
}

// fused stage kernels live in gx_stage.cu, one translation unit per solver
int l_stage_1(const StepArgs&, int, double, const double*, const double*, double*, double*, int, unsigned long long*, int, int*, cudaStream_t);
int l_stage_2(const StepArgs&, int, double, const double*, const double*, double*, double*, int, unsigned long long*, int, int*, cudaStream_t);
int l_stage_3(const StepArgs&, int, double, const double*, const double*, double*, double*, int, unsigned long long*, int, int*, cudaStream_t);
int l_stage_4(const StepArgs&, int, double, const double*, const double*, double*, double*, int, unsigned long long*, int, int*, cudaStream_t);
static int l_stage(const StepArgs& A, int order, double dt, const double* S, const double* Ub, double* dst, double* E, int kz,
                   unsigned long long* dtmin_bits, int want_cfl, int* errflag, cudaStream_t s) {
  switch (A.solver) {
    case GX_SOLVER_HLL:  return l_stage_1(A, order, dt, S, Ub, dst, E, kz, dtmin_bits, want_cfl, errflag, s);
    case GX_SOLVER_HLLC: return l_stage_2(A, order, dt, S, Ub, dst, E, kz, dtmin_bits, want_cfl, errflag, s);
    case GX_SOLVER_HLLE: return l_stage_3(A, order, dt, S, Ub, dst, E, kz, dtmin_bits, want_cfl, errflag, s);
    case GX_SOLVER_HLLD: return l_stage_4(A, order, dt, S, Ub, dst, E, kz, dtmin_bits, want_cfl, errflag, s);
  }
  return GX_EUNSUPPORTED;
}

static void l_bupdate(const StepArgs& A, double dt, const double* Ub, const double* E, double* dst, unsigned long long* dtmin_bits, int want_cfl, cudaStream_t s) {
  const Grid& g = A.g;
  const dim3 block(32, GX_BU_Y, GX_BU_Z);
  const int n2 = A.klast2 >= A.kbeg2 ? A.klast2 - A.kbeg2 + 1 : 0;
  const dim3 grid((g.nx + 63) / 64, (g.ny + GX_BU_Y - 1) / GX_BU_Y, (A.klast - A.kbeg + 1 + GX_BU_Z - 1) / GX_BU_Z + (n2 + GX_BU_Z - 1) / GX_BU_Z);
  const double dtdx = dt / g.dx, dtdy = dt / g.dy, dtdz = dt / g.dz;
  if (want_cfl) k_bupdate<true><<<grid, block, 0, s>>>(A, dtdx, dtdy, dtdz, Ub, E, dst, dtmin_bits);
  else k_bupdate<false><<<grid, block, 0, s>>>(A, dtdx, dtdy, dtdz, Ub, E, dst, dtmin_bits);
}

static void l_coolingh(const StepArgs& A, int mhd, double dt_seconds, double* U, cudaStream_t s) {
  const Grid& g = A.g;
  k_coolingh<<<grid_for(g.nx, g.ny, g.nz, 128), 128, 0, s>>>(g, A.phys, mhd, dt_seconds, U);
}
static void l_tc_prim(const StepArgs& A, int mhd, const double* U, double* PT, unsigned long long* dt_bits, int want_dt, cudaStream_t s) {
  const Grid& g = A.g;
  // want_dt: the candidates are collected in 256 words (16 doubles apart) behind the two scratch variables, then folded into *dt_bits
  unsigned long long* slots = reinterpret_cast<unsigned long long*>(PT + 2 * g.vs);
  if (want_dt) cudaMemsetAsync(slots, 0x7f, 256 * 16 * sizeof(unsigned long long), s);
  k_tc_prim<<<dim3((g.nx + 2 + 127) / 128, g.ny + 2, g.nz + 2), 128, 0, s>>>(g, A.phys, mhd, U, PT, slots, want_dt);
  if (want_dt) k_tc_slots_min<<<1, 256, 0, s>>>(slots, dt_bits);
}
static void l_tc_update(const StepArgs& A, const TcPar& t, int fill, const double* PT, double* U, double dts, cudaStream_t s) {
  const Grid& g = A.g;
  const dim3 grid = grid_for(g.nx, g.ny, g.nz, 128);
  if (fill) k_tc_update<true><<<grid, 128, 0, s>>>(g, A.phys, t, PT, U, dts);
  else k_tc_update<false><<<grid, 128, 0, s>>>(g, A.phys, t, PT, U, dts);
}
static void l_tc_march(const StepArgs& A, const TcPar& t, int mhd, const double* U, const double* E5in, double* E5out, double dts, cudaStream_t s) {
  const Grid& g = A.g;
  const int tiles = ((g.nx + TCM_TX - 1) / TCM_TX) * ((g.ny + TCM_TY - 1) / TCM_TY);
  // planes per CTA: long marches (the prologue of a chunk costs two planes of conversions), but at least ~3 waves of resident CTAs
  int chunks = std::max(1, std::min(g.nz / 16, (148 * 3 * 3 + tiles - 1) / tiles));
  const int kz = (g.nz + chunks - 1) / chunks;
  chunks = (g.nz + kz - 1) / kz;
  const dim3 grid((g.nx + TCM_TX - 1) / TCM_TX, (g.ny + TCM_TY - 1) / TCM_TY, chunks), block(TCM_TX, TCM_TY);
  if (mhd) k_tc_march<true><<<grid, block, 0, s>>>(g, A.phys, t, U, E5in, E5out, dts, kz);
  else k_tc_march<false><<<grid, block, 0, s>>>(g, A.phys, t, U, E5in, E5out, dts, kz);
}
static void l_tc_fill(const StepArgs& A, double* Aq, int edge, cudaStream_t s) {
  const Grid& g = A.g;
  k_tc_fill<<<dim3((std::max(g.nx + 2, g.ny) + 127) / 128, 2 * (g.ny + 2) + 4 * g.nz), 128, 0, s>>>(g, Aq, edge);
}

static const KernelTable table = {l_calcprim, l_fluxes, l_efield, l_update, l_viscous, l_viscous2, l_stage, l_bupdate, l_tc_prim, l_tc_update, l_tc_fill, l_tc_march, l_coolingh, l_riemann_points};

}  // namespace GX_NS

#if defined(GX_FLAVOUR_STRICT)
const KernelTable* kernels_strict() { return &strict_ns::table; }
#else
const KernelTable* kernels_fast() { return &fast_ns::table; }
#endif

}  // namespace gx
