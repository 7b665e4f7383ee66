// gx_kernels.cuh — device layout + kernel declarations shared by the API layer and
// the two kernel builds (strict: -fmad=false, fast: -fmad=true).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "gx_physics.cuh"

namespace gx {

// Device layout: SoA, x fastest.  Variable v of cell (i,j,k) [Fortran indices,
// i = 1-ng .. nx+ng] lives at  v*vs + ((k+1)*py + (j+1))*px + (i + xo).
// xo = 15 puts the first physical cell (i=1) on a 128-byte boundary of every row
// (px is a multiple of 16 doubles, cudaMalloc bases are 256-byte aligned), so a warp
// reading 32 consecutive physical cells touches exactly two 128-byte lines.
struct Grid {
  int nx, ny, nz;          // physical cells of this block
  int px, py, pz;          // padded extents (px: row pitch in doubles)
  int xo;                  // x offset of Fortran index 0
  long long vs;            // variable stride = px*py*pz
  int neq, neqdyn, npas;
  int cx, cy, cz;          // block coords (for positions in source terms)
  int nxtot, nytot, nztot;
  double dx, dy, dz;
  __host__ __device__ __forceinline__ long long idx(int i, int j, int k) const {
    return ((long long)(k + 1) * py + (j + 1)) * px + (i + xo);
  }
};

struct GravityPoints {     // get_user_source_terms functor (EXO/user_mod.f90:158-206)
  int n;
  double gm[4], x[4], y[4], z[4];
};

struct StepArgs {
  Grid g;
  int wrap[3];             // direction is periodic with the block as its own neighbour: the fused
                           // kernels read the wrapped cell instead of a ghost cell (ghosts are
                           // then only materialised when the host asks for the arrays)
  int tma;                 // TMA tile loads (one per staged plane) instead of per-thread cp.async: 0 off, 1 second-order stage, 2 both stages
  int kbeg, klast;         // planes (Fortran k) the fused stage / B-update launch covers; 1..nz unless the step is split into
                           // boundary-first and interior launches to overlap the halo exchange (multi-GPU)
  int kbeg2, klast2;       // second plane range of the same launch (the two boundary slabs of a block go out as ONE launch);
                           // empty when klast2 < kbeg2
  gxp::Phys phys;
  double idx3[3];          // 1/dx, 1/dy, 1/dz (production kernels: CFL candidates without divisions)
  int solver, limiter;
  int flux_cd, eight_wave, user_src;
  GravityPoints grav;
  const double* W0;        // background primitives of the split-all solver (src/globals.f90:42 primit0; gx_set_background), else null
};

// thermal conduction (src/thermal_cond.f90): what the substep kernels read besides the grid
struct TcPar {
  int mode;                  // GX_TC_ISOTROPIC | GX_TC_ANISOTROPIC
  int sat;                   // tc_saturation
  int mhd;
  double dxr, dyr, dzr;      // dx*rsc, dy*rsc, dz*rsc (the reference divides by the product)
  double idxr, idyr, idzr;   // their reciprocals (production kernels)
  double dx, dy, dz, idx, idy, idz;
  double vsc, sqrt_vsc2;     // parameters.f90:167 ; sqrt(vsc2) as heatfluxes spells it (:213)
  double Psc, rhosc, bsc2;   // Psc = rhosc*vsc2 (:168) ; bsc**2
};

// classes for the per-kernel timing table (gx_kernel_time_ms)
enum { KC_FLUX = 0, KC_UPDATE = 1, KC_EFIELD = 2, KC_PRIM = 3, KC_BC = 4, KC_XPOSE = 5, KC_VISC = 6, KC_STAGE1 = 7, KC_STAGE2 = 8, KC_BUPDATE = 9, KC_TCOND = 10, KC_COUNT = 11 };

// One set of launchers per build flavour.
struct KernelTable {
  // primitives (+Temp, + CFL min into *dtmin_bits when want_cfl) over the whole padded array
  void (*calcprim)(const StepArgs&, const double* U, double* W, double* Temp, unsigned long long* dtmin_bits, int want_cfl, cudaStream_t);
  // face fluxes, order 1|2, all three directions (3 launches)
  int (*fluxes)(const StepArgs&, int order, const double* W, double* F, int* errflag, cudaStream_t);
  void (*efield)(const StepArgs&, const double* F, double* E, cudaStream_t);
  // dst = U - dt*div(F) [flux-CD for B] [+ dt*S(W)]
  void (*update)(const StepArgs&, double dt, const double* U, const double* F, const double* E, const double* W, double* dst, cudaStream_t);
  void (*viscous)(const StepArgs&, double eta, const double* UP, double* U, cudaStream_t);
  // viscous_copy of the fused path: the full-step state lives in T (physical cells), the ghost cells it reads are up's
  void (*viscous2)(const StepArgs&, double eta, const double* T, const double* UP, double* U, cudaStream_t);
  // fused stage (gx_stage.cu): dst = Ub - dt*div F(prim(S)) for the non-B variables (all variables
  // without flux-CD), E = cell-centred electric field of the same fluxes (flux-CD only);
  // without flux-CD and with want_cfl the CFL minimum of the new state goes to *dtmin_bits.
  int (*stage)(const StepArgs&, int order, double dt, const double* S, const double* Ub, double* dst, double* E, int kz /* planes per CTA; <= 0: chosen by the launcher */,
               unsigned long long* dtmin_bits, int want_cfl, int* errflag, cudaStream_t);
  // flux-CD: dst(B) = Ub(B) - dt*curl E (central differences); want_cfl: CFL minimum of dst
  void (*bupdate)(const StepArgs&, double dt, const double* Ub, const double* E, double* dst,
                  unsigned long long* dtmin_bits, int want_cfl, cudaStream_t);
  // thermal conduction (gx_thermal.cuh): pressure + temperature over 0..n+1 (+ Spitzer time-scale candidates); one substep of
  // u(5) over the physical cells (fill: the block owns the whole domain, ghost copies written by the update); the zero-gradient
  // ghost layer of u(5) after an exchange
  void (*tc_prim)(const StepArgs&, int mhd, const double* U, double* PT, unsigned long long* dt_bits, int want_dt, cudaStream_t);
  void (*tc_update)(const StepArgs&, const TcPar&, int fill, const double* PT, double* U, double dts, cudaStream_t);
  void (*tc_fill)(const StepArgs&, double* A, int edge, cudaStream_t);
  // isotropic conduction, block without neighbours: one marching kernel per substep, u(5) from E5in to E5out (k_tc_march)
  void (*tc_march)(const StepArgs&, const TcPar&, int mhd, const double* U, const double* E5in, double* E5out, double dts, cudaStream_t);
  // COOL_H (gx_cooling.cuh): atomic(dt, uu) over the physical cells (src/cooling_h.f90:41-67, 259-371)
  void (*coolingh)(const StepArgs&, int mhd, double dt_seconds, double* U, cudaStream_t);
  // per-interface flux of n (rotated) primitive state pairs [n][8] (gx_riemann_flux)
  int (*riemann_points)(const gxp::Phys&, int solver, int n, const double* wl, const double* wr, double* ff, int* err, cudaStream_t);
};
const KernelTable* kernels_strict();
const KernelTable* kernels_fast();

}  // namespace gx
