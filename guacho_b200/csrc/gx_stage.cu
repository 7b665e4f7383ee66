// gx_stage.cu — fused stage kernel of the hydro/MHD step for sm_100a.
//
// One launch does what the reference does in five full-array passes per stage
// (src/hydro_solver.f90:155-184): calcprim (u2prim, src/hydro_core.f90:46-129), the
// hll?fluxes(choice) sweep in x, y and z (src/hlld.f90:331-432 and twins), get_efield
// (src/flux_cd_module.f90:245-273) and the conservative update of step()
// (src/hydro_solver.f90:99-107).  Primitives and face fluxes never touch HBM: the kernel
// reads the conserved state once and writes the updated state (and, with flux-CD, the
// cell-centred electric field) once.
//
// Structure (2.5-D marching, FP64 stencil):
//   * a CTA owns a TX x TY = 32 x 11 column of cells (12 warps = 3 per scheduler, one CTA per SM at <= 168
//     registers) and marches along z through KZ planes;
//   * conserved planes are staged into shared memory one plane ahead of the compute — ONE TMA tile load per plane
//     (cp.async.bulk.tensor, a 4-D box x, y, plane, variable of the SoA array, completion on an mbarrier) where the
//     loader need not wrap indices, per-thread cp.async (LDGSTS, 8-byte elements) otherwise — converted in place to
//     primitives, and kept in a ring of 2*ORDER planes (the z stencil) + 1 in flight.  The first-order stage also keeps the
//     per-cell signal speeds of the Riemann solver in the ring (cell_speeds): there the states of
//     a face ARE cell states, so the speeds are evaluated once per cell and direction, not per face and side;
//   * warp r (< TY) owns row r of the tile, lane l owns cell i0+l.  Each thread solves the
//     LOWER x face, the LOWER y face of its cell and then the UPPER z face, so every interface is
//     solved exactly once inside the tile; warp TY solves the tile's closing faces (the x faces
//     right of the last column, then the y faces above the last row);
//   * the three solves of a thread run through ONE copy of the Riemann solver in the instruction stream
//     (a run-time loop over the face direction); what differs per direction — where the 2*ORDER states
//     come from and where the flux goes — is compile-time specialised code either side of it (rotation of
//     the components, swapy/swapz of src/hydro_core.f90:485-534, as constant offsets);
//   * x and y face fluxes are exchanged through shared memory behind split mbarrier arrive/wait pairs;
//     the z flux never leaves registers: the update of the cell is the epilogue of its z solve.  The
//     update is written with fully coalesced 256-byte row segments.
// Compiled per (flavour, solver): -DGX_FLAVOUR_STRICT|-DGX_FLAVOUR_FAST, -DGX_STAGE_SOLVER=n.
#define GX_SOLVE_MASK 0xffffffffu   // every interface solve of this kernel is executed by all 32 lanes of a warp
#include <cuda.h>                    // CUtensorMap (the TMA descriptor of a staged array)
#include <mutex>
#include <vector>
#include "gx_kernels.cuh"

#if defined(GX_FLAVOUR_STRICT)
#define GX_NS strict_ns
#elif defined(GX_FLAVOUR_FAST)
#define GX_NS fast_ns
#else
#error "define GX_FLAVOUR_STRICT or GX_FLAVOUR_FAST"
#endif
#ifndef GX_STAGE_SOLVER
#error "define GX_STAGE_SOLVER (1..4)"
#endif
#ifndef GX_STAGE_PRESPEED            // 1: per-cell signal speeds in the ring of the first-order stage
#define GX_STAGE_PRESPEED 1
#endif
#ifndef GX_STAGE_UB_EARLY            // 1: the second-order stage loads the base state of the update (a cold array) before the z solve,
#define GX_STAGE_UB_EARLY 1          //    so it is in flight during the solve; the first-order stage (base state = the staged array,
#endif                               //    hot in L2) loads it in the epilogue.  Measured at 256^3: 1.81 vs 1.85 ms (stage 2)

namespace gx {
namespace GX_NS {

// storage component of rotated slot c for sweep direction D (swapy / swapz as an index map,
// src/hydro_core.f90:485-534)
template <int D> __device__ __forceinline__ constexpr int rot(int c) {
  return (c == 1) ? 1 + D : (c == 1 + D) ? 1 : (c == 5) ? 5 + D : (c == 5 + D) ? 5 : c;
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async8(unsigned smem_dst, const double* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
#ifndef GX_STAGE_UB_PREFETCH         // where the base state of the update is prefetched to at the top of a plane: 0 nowhere, 1 L1, 2 L2
#define GX_STAGE_UB_PREFETCH 2
#endif
__device__ __forceinline__ void prefetch_ub(const double* p) {
#if GX_STAGE_UB_PREFETCH == 1
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#elif GX_STAGE_UB_PREFETCH == 2
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}
// Split CTA barriers on mbarriers (arrive early, wait late): with one CTA per SM a full
// __syncthreads idles the SM, so every hand-over in the plane loop is an arrive followed,
// as late as the data dependence allows, by a parity wait.  All NT threads arrive once per
// plane on each barrier; phase parity = plane counter & 1.  Barriers are addressed by their
// 32-bit shared-window address, computed once.
__device__ __forceinline__ void mbar_init(unsigned bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// TMA tile load of one staged plane: box (CX, RY, 1, NU) of the 4-D tensor (x, y, plane, variable) -> dense [variable][row][column]
// in shared memory; the bytes are counted on `bar` (expect_tx by the issuing thread, complete_tx by the copy engine).
__device__ __forceinline__ void tma_load_plane(unsigned smem_dst, unsigned long long tm, int c0, int c1, int c2, unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
               ::"r"(smem_dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(0), "r"(bar) : "memory");
}
// generic-proxy accesses (the in-place conversion, the solves' reads) before the async proxy (TMA) overwrites the slot
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned bar, int parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "MBAR_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra MBAR_WAIT_%=;\n\t}"
      ::"r"(bar), "r"(parity) : "memory");
}

#ifndef GX_STAGE_TY1                 // tile rows (= main warps) of the first-order / second-order stage kernel
#define GX_STAGE_TY1 11
#endif
#ifndef GX_STAGE_TY2
#define GX_STAGE_TY2 11
#endif
#ifndef GX_STAGE_TYP1                // tile rows of the kernels that carry passive scalars (more variables per staged cell): first-order stage
#define GX_STAGE_TYP1 11             // (3 ring slots of 13 variables: 12 warps fit.  EXO 400x100x400: 1.81 ms against 1.89 with 7 rows, 1.97 with 10)
#endif
#ifndef GX_STAGE_TYP2                // ... second-order stage (5 ring slots of 10 variables: 8 rows would fit, but 9 warps are capped at 168 registers
#define GX_STAGE_TYP2 7              //     like 12 and the kernel spills: 3.83 ms against 2.56)
#endif
#ifndef GX_STAGE_CONVERT_EARLY       // 1: the next plane is converted to primitives BEFORE the z solve (between the arrive and the wait of the
#define GX_STAGE_CONVERT_EARLY 0     //    XY hand-over: more slack for warps that run ahead) instead of at the end of the plane.  Measured at 256^3:
                                     //    stage 1 1.313 vs 1.301 ms, stage 2 1.742 vs 1.694 ms — rejected
#endif
#ifndef GX_STAGE_MINB1               // resident CTAs per SM the first-order headline kernel is compiled for (register cap = 64 K / threads)
#define GX_STAGE_MINB1 1
#endif
#ifndef GX_STAGE_NPAS                // passive scalars of the second set of fused kernels (EXO: neutral H density + tracer)
#define GX_STAGE_NPAS 2
#endif

template <int NQ_, int ORDER_, int NCF_, int NPAS_>
struct StageGeom {
  static constexpr int NQ = NQ_, H = ORDER_, NCF = NCF_, NPAS = NPAS_;
  static constexpr int NU = NQ + NPAS;           // advected variables: dynamic + passive scalars
  static constexpr int NV = NU + NCF;            // staged per cell: primitives, passives (+ signal speeds, first-order stage)
  static constexpr int TX = 32, TY = NPAS_ ? ((ORDER_ == 1) ? GX_STAGE_TYP1 : GX_STAGE_TYP2) : ((ORDER_ == 1) ? GX_STAGE_TY1 : GX_STAGE_TY2);
  static constexpr int NW = TY + 1, NT = NW * 32;
  static constexpr int HX = 2;                   // x halo of the staged frame: 2 for both orders, so that every staged row starts on an
                                                 // even element (16-byte aligned box start: a TMA tile load faults on less)
  static constexpr int CX = TX + 2 * HX;         // staged columns  i0-HX .. i0+TX+HX-1
  static constexpr int RY = TY + 2 * H;          // staged rows     j0-H .. j0+TY+H-1
  static constexpr int NSLOT = 2 * H + 1;        // z ring: 2H planes of stencil + 1 in flight
  static constexpr int PCELLS = CX * RY;
  static constexpr int PLANE = (NV * PCELLS + 15) / 16 * 16;   // doubles per ring slot (128-byte multiple: TMA destination alignment)
  static constexpr int XBV = TY * (TX + 1);      // x-face flux exchange, per variable
  static constexpr int YBV = (TY + 1) * TX;      // y-face flux exchange, per variable
  static constexpr int XB = NU * XBV, YB = NU * YBV;
  static constexpr int SCR = 32;                 // scratch: block reduction, dead-lane stores
  // TMA loader with a direction wrapped in the loader: the staged cells that are periodic images (two ghost columns of an x-edge
  // tile, two ghost rows of a y-edge tile) cannot come out of the box; the thread that converts such a cell fetches it itself
  // (cp.async from the far side of the block) into its own entry of this patch
  static constexpr int NPATCH = 2 * RY + 2 * CX;
  static constexpr int PATCH = (NPAS_ == 0) ? NU * NPATCH : 0;
  static constexpr size_t SMEM = sizeof(double) * ((size_t)NSLOT * PLANE + XB + YB + SCR + PATCH) + 128;   // + slack to align the ring to 128 bytes
};

// block-wide min of positive doubles -> one atomicMin on the ordered bit pattern
template <int NT>
__device__ __forceinline__ void stage_block_min(double v, unsigned long long* dst, double* scratch) {
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  if (wid == 0) {
    v = lane < NT / 32 ? scratch[lane] : 1.e30;
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) atomicMin(dst, (unsigned long long)__double_as_longlong(v));
  }
}

// The 2*ORDER states of one interface of sweep direction D, gathered from the staged primitive planes and reconstructed
// (limiter, src/hydro_core.f90:712-798).  p_m2 .. p_p1 point at variable 0 of cells l-1, l | r, r+1; rotated slot q
// is storage component rot<D>(q): constant offsets.  PRE: the signal speeds of the two states come from the ring.
template <int D, int LIM, int ORDER, int NQ, int NU, int PC, bool PRE>
__device__ __forceinline__ void gather(const double* p_m2, const double* p_m1, const double* p_p0, const double* p_p1,
                                       double (&wl)[8], double (&wr)[8], double& csl, double& csr) {
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int c = rot<D>(q) * PC;
    double pl = p_m1[c], pr = p_p0[c];
    if (ORDER == 2) gxp::reconstruct<LIM>(p_m2[c], pl, pr, p_p1[c]);
    wl[q] = pl; wr[q] = pr;
  }
  if (PRE) {
    const int c = (NU + (NQ == 8 ? D : 0)) * PC;   // MHD: one fast speed per direction; hydro: the sound speed
    csl = p_m1[c]; csr = p_p0[c];
  }
}

struct StageDt { double dtdx, dtdy, dtdz, dt; };

template <int SOLVER, int LIM, int ORDER, bool FLUXCD, int NPAS>
struct StageTraits {
  static constexpr bool MHD = (SOLVER == GX_SOLVER_HLLE || SOLVER == GX_SOLVER_HLLD);
  static constexpr int NQ = MHD ? 8 : 5;
  static constexpr int NCF = (ORDER == 1 && GX_STAGE_PRESPEED) ? (MHD ? 3 : 1) : 0;
  using G = StageGeom<NQ, ORDER, NCF, NPAS>;
};

// NPAS = 0: dynamic variables, adiabatic equation of state, no sources (the headline path).  NPAS > 0: the kernel also advects
// NPAS passive scalars (limited like every primitive; fluxes from the closed forms of passive_flux), evaluates
// u2prim with the run-time equation of state (EOS_H_RATE reads the first passive) and adds the point-mass gravity functor
// of get_user_source_terms in the update — EXO as shipped leaves the pass-per-routine path.
template <int SOLVER, int LIM, int ORDER, bool FLUXCD, int NPAS, bool TMA>
__global__ void __launch_bounds__((StageTraits<SOLVER, LIM, ORDER, FLUXCD, NPAS>::G::NT), (ORDER == 1 && NPAS == 0) ? GX_STAGE_MINB1 : 1)
k_stage(const StepArgs A, const StageDt sdt, const double* __restrict__ S, const double* Ub, double* dst,
        double* __restrict__ E, const int kz, unsigned long long* dtmin_bits, const int want_cfl, int* errflag,
        const __grid_constant__ CUtensorMap tmS) {
  using T = StageTraits<SOLVER, LIM, ORDER, FLUXCD, NPAS>;
  using G = typename T::G;
  constexpr bool MHD = T::MHD;
  constexpr int NQ = T::NQ, NCF = T::NCF, NU = G::NU;
  constexpr bool PRE = NCF > 0;
  constexpr bool UB_EARLY = GX_STAGE_UB_EARLY && ORDER == 2;
  constexpr int H = G::H, HX = G::HX, TX = G::TX, TY = G::TY, CX = G::CX, NSLOT = G::NSLOT, NT = G::NT;
  constexpr int PC = G::PCELLS;
  constexpr int XBV = G::XBV, YBV = G::YBV;
  extern __shared__ double sm[];
  double* const ring = sm + (((smem_u32(sm) + 127u) & ~127u) - smem_u32(sm)) / 8u;   // 128-byte aligned (TMA destination)
  double* const xb = ring + (size_t)NSLOT * G::PLANE;    // [q][TY][TX+1]
  double* const yb = xb + G::XB;                         // [q][TY+1][TX]
  double* const scr = yb + G::YB;                        // [32]
  double* const patch = scr + G::SCR;                    // [NU][NPATCH] (TMA loader only)
  __shared__ unsigned long long bars[2 + NSLOT];         // XY, FREE, and (TMA) one "plane landed" barrier per ring slot

  const Grid& g = A.g;
  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  // z chunk of this CTA: the chunks of the first plane range, then those of the second (boundary slabs of the overlapped step)
  const int nch1 = (A.klast - A.kbeg + kz) / kz;
  const bool second = (int)blockIdx.z >= nch1;
  const int i0 = 1 + (int)blockIdx.x * TX, j0 = 1 + (int)blockIdx.y * TY;
  const int k0 = second ? A.kbeg2 + ((int)blockIdx.z - nch1) * kz : A.kbeg + (int)blockIdx.z * kz;
  const int kend = min(k0 + kz - 1, second ? A.klast2 : A.klast);
  const long long vs = g.vs;
  const bool main_warp = wrp < TY;                       // warp-uniform

  // ---- plane staging: conserved -> shared (cp.async), converted in place to primitives ----
  // Ownership: a main-warp thread stages and converts ITS OWN centre cell (so its z solves, which
  // only read its own column, never wait for another thread), and the halo frame of the plane
  // (consumed ORDER+1 planes later by the x/y solves) is dealt one cell per thread from the LAST thread
  // down, so the closing warp (two solves per plane instead of three) takes its share first.
  const int cidx = (min(wrp, TY - 1) + H) * CX + (lane + HX);  // this thread's cell inside a staged plane
  constexpr int WN = TX + 2 * H;                         // columns the stencil needs (the frame may be wider: HX >= H)
  constexpr int HC = 2 * H * WN + 2 * H * TY;            // halo cells of a staged plane that are ever read
  static_assert(HC <= NT, "one halo cell per thread");
  auto halo_cell = [&](int h) {                          // h-th halo cell -> plane-local index
    if (h < H * WN) { const int r = h / WN; return r * CX + (HX - H) + (h - r * WN); }                 // rows below the tile
    if (h < H * WN + 2 * H * TY) {
      const int t = h - H * WN, r = t / (2 * H), sx = t - r * (2 * H);
      return (H + r) * CX + (HX - H) + (sx < H ? sx : TX + sx);                                        // left | right columns
    }
    const int t = h - (H * WN + 2 * H * TY), r = t / WN;
    return (H + TY + r) * CX + (HX - H) + (t - r * WN);                                                // rows above the tile
  };
  const bool has_halo = (NT - 1 - tid) < HC;
  const int hcell = halo_cell(has_halo ? NT - 1 - tid : 0);
  // ring bookkeeping: plane k0-H sits in slot 0; `sk` (slot of the current plane k) advances by one per plane
  auto slot_of = [&](int p) { return (p - (k0 - H)) % NSLOT; };                 // prologue only
  auto slot_add = [&](int s, int d) { const int t = s + d; return t >= NSLOT ? t - NSLOT : t; };   // 0 <= d <= NSLOT
  // The (i, j) offsets of a thread's two cells inside a global plane — clamped to the array, wrapped where the
  // block is its own periodic neighbour — are fixed for the whole march.
  // patch entry of a staged cell that is a periodic image (TMA loader): -1 if it is not, or if nobody ever reads it (corner
  // cells, rows / columns more than two cells outside the block)
  constexpr int NPATCH = G::NPATCH;
  auto ij_off = [&](int c, int& pid) {
    const int rr = c / CX, cc = c - rr * CX;
    const int iu = i0 - HX + cc, ju = j0 - H + rr;
    int i = min(iu, g.nx + 2), j = min(ju, g.ny + 2);
    if (A.wrap[0]) i = i < 1 ? i + g.nx : (i > g.nx ? i - g.nx : i);
    if (A.wrap[1]) j = j < 1 ? j + g.ny : (j > g.ny ? j - g.ny : j);
    const bool gx_ = A.wrap[0] && (iu < 1 || iu > g.nx), gy_ = A.wrap[1] && (ju < 1 || ju > g.ny);
    pid = -1;
    if (gx_ && !gy_ && iu >= -1 && iu <= g.nx + 2) pid = rr * 2 + (iu < 1 ? iu + 1 : iu - g.nx - 1);                      // two ghost columns, every row
    if (gy_ && !gx_ && ju >= -1 && ju <= g.ny + 2) pid = 2 * G::RY + (ju < 1 ? ju + 1 : ju - g.ny - 1) * CX + cc;         // two ghost rows, every column
    return (j + 1) * g.px + (i + g.xo);
  };
  int own_pid, halo_pid;
  const int own_off = ij_off(cidx, own_pid), halo_off = ij_off(hcell, halo_pid);
  if (!TMA || !main_warp) own_pid = -1;
  if (!TMA || !has_halo) halo_pid = -1;
  const int gplane = g.px * g.py;                         // cells per plane of one variable (fits 32 bits)
  const unsigned own_u32 = smem_u32(ring + cidx), halo_u32 = smem_u32(ring + hcell);
  const double* const S_own = S + own_off;                // variable 0 of my cells in plane index 0 of the padded array
  const double* const S_halo = S + halo_off;
  auto stage_cell = [&](unsigned d, const double* src) {
#pragma unroll
    for (int q = 0; q < NU; ++q) { cp_async8(d + (unsigned)(q * PC) * 8u, src); src += vs; }
  };
  const unsigned ring_u32 = smem_u32(ring), bar_full0 = smem_u32(&bars[2]);
  const bool issuer = TMA && tid == NT - 32;              // lane 0 of the closing warp issues the tile loads
  // address of the descriptor IN THE PARAMETER SPACE (taken here, not inside a lambda: a by-reference capture may copy the
  // parameter to local memory, and a descriptor there is an illegal instruction for the copy engine)
  const unsigned long long tm_addr = reinterpret_cast<unsigned long long>(&tmS);
  auto issue_load = [&](int p, int slot) {
    int kk = min(p, g.nz + 2);
    if (A.wrap[2]) kk = kk < 1 ? kk + g.nz : (kk > g.nz ? kk - g.nz : kk);
    const unsigned so = (unsigned)(slot * G::PLANE) * 8u;
    const long long po = (long long)(kk + 1) * gplane;
    if (TMA) {
      if (issuer) tma_load_plane(ring_u32 + so, tm_addr, i0 - HX + g.xo, j0 - H + 1, kk + 1, bar_full0 + 8u * (unsigned)slot, (unsigned)(NU * PC * 8));
      return;
    }
    if (main_warp) stage_cell(own_u32 + so, S_own + po);
    if (has_halo) stage_cell(halo_u32 + so, S_halo + po);
    cp_async_commit();
  };
  // periodic-image cells of plane p -> this thread's patch entries (TMA loader; the entries are private to the thread, which
  // reads them when it converts the plane and only then fetches the next plane's)
  const unsigned patch_u32 = smem_u32(patch);
  auto issue_patch = [&](int p) {
    if (!TMA || (own_pid < 0 && halo_pid < 0)) return;
    int kk = min(p, g.nz + 2);
    if (A.wrap[2]) kk = kk < 1 ? kk + g.nz : (kk > g.nz ? kk - g.nz : kk);
    const long long po = (long long)(kk + 1) * gplane;
    if (own_pid >= 0) {
      const double* src = S_own + po;
#pragma unroll
      for (int q = 0; q < NU; ++q) { cp_async8(patch_u32 + (unsigned)(q * NPATCH + own_pid) * 8u, src); src += vs; }
    }
    if (halo_pid >= 0) {
      const double* src = S_halo + po;
#pragma unroll
      for (int q = 0; q < NU; ++q) { cp_async8(patch_u32 + (unsigned)(q * NPATCH + halo_pid) * 8u, src); src += vs; }
    }
    cp_async_commit();
  };
  auto wait_load = [&](int slot, int parity) {            // the plane staged into `slot` has landed
    if (TMA) { mbar_wait(bar_full0 + 8u * (unsigned)slot, parity); if (own_pid >= 0 || halo_pid >= 0) cp_async_wait_all(); }
    else cp_async_wait_all();
  };
  auto convert_cell = [&](double* sl, int c, int pid) {
    double u[8], w[8], Tk;
    if (TMA && pid >= 0) {
#pragma unroll
      for (int q = 0; q < NQ; ++q) u[q] = patch[q * NPATCH + pid];
    } else {
#pragma unroll
      for (int q = 0; q < NQ; ++q) u[q] = sl[q * PC + c];
    }
    gxp::u2prim<MHD, false, NPAS == 0>(A.phys, u, w, NPAS ? sl[NQ * PC + c] : 0.0, Tk);   // passives are their own primitives
#pragma unroll
    for (int q = 0; q < NQ; ++q) sl[q * PC + c] = w[q];
    if (PRE) {
      double cs[3];
      gxp::cell_speeds<MHD>(A.phys, w, cs);
#pragma unroll
      for (int d = 0; d < NCF; ++d) sl[(NU + d) * PC + c] = cs[d];
    }
  };
  auto convert = [&](int slot) {   // each thread converts the same cells it stages on the cp.async path: its centre cell and one halo cell
    double* sl = ring + slot * G::PLANE;
    if (main_warp) convert_cell(sl, cidx, own_pid);
    if (has_halo) convert_cell(sl, hcell, halo_pid);
  };

  // Barriers of the plane loop (all NT threads arrive once per plane on each):
  //   XY   : x and y face fluxes of this plane are in the exchange buffers
  //   FREE : this thread has consumed the exchange buffers (they may be overwritten)
  const unsigned bar_xy = smem_u32(&bars[0]), bar_free = smem_u32(&bars[1]);
  if (tid == 0) {
    mbar_init(bar_xy, NT); mbar_init(bar_free, NT);
    if (TMA) {
      for (int sl = 0; sl < NSLOT; ++sl) mbar_init(bar_full0 + 8u * (unsigned)sl, 1);
    }
  }
  if (TMA) __syncthreads();                              // barriers initialised before the first tile load counts on them
#pragma unroll 1
  for (int p = k0 - H; p <= k0 + H - 1; ++p) issue_load(p, slot_of(p));
#pragma unroll 1
  for (int p = k0 - H; p <= k0 + H - 1; ++p) { issue_patch(p); wait_load(slot_of(p), 0); convert(slot_of(p)); }   // (one patch entry per cell: plane by plane)
  __syncthreads();

  const int i = i0 + lane, j = j0 + wrp;
  const bool cell_ok = main_warp && i <= g.nx && j <= g.ny;
  // plane-invariant parts of this thread's face jobs (x: the extra warp closes right of the last column, row = lane;
  // y: the extra warp closes the row above the tile)
  const int row_x = main_warp ? wrp : min(lane, TY - 1), col_x = main_warp ? lane : TX;
  const int c0x = (row_x + H) * CX + (col_x + HX);       // right cell of my x face
  const int c0y = cidx + (main_warp ? 0 : CX);           // upper cell of my y face
  // lanes of the closing warp that own no x face write to a scratch word instead of branching around the stores
  double* const out_x = (main_warp || lane < TY) ? xb + row_x * (TX + 1) + col_x : scr + lane;
  const int ovs_x = (main_warp || lane < TY) ? XBV : 0;
  double* const out_y = yb + wrp * TX + lane;
  const bool check_x = main_warp ? (i <= g.nx + 1 && j <= g.ny) : (lane < TY && i0 + TX <= g.nx + 1 && j0 + lane <= g.ny);
  const bool check_y = (i <= g.nx && j <= g.ny + 1);
  const int njobs = main_warp ? 3 : 2;
  // hprev: flux through the lower z face of my cell (the previous plane's z solve), storage components
  double hprev[8 + NPAS];
#pragma unroll
  for (int q = 0; q < 8 + NPAS; ++q) hprev[q] = 0.0;
  double dtp = 1.e30;
  int err = 0;
  int it = 0;                                             // x/y plane counter (barrier phase)

  int sk = H - 1;                                         // ring slot of plane k (plane k0-H is slot 0)
  // my cell (clamped into the block so that every lane forms a valid address) in plane index 0; the plane offset is added per plane
  const long long cg0 = g.idx(min(i, g.nx), min(j, g.ny), -1);
  int lpar = 0;                                           // phase parity of the "landed" barrier of the slot being loaded
#pragma unroll 1
  for (int k = k0 - 1; k <= kend; ++k, sk = slot_add(sk, 1)) {
    const int sload = slot_add(sk, H + 1);                // slot of plane k-H = slot of plane k+H+1
    if (sload == 0) lpar ^= 1;                            // the ring has gone round once more
    const bool xy = k >= k0;                              // the chunk's leading plane only supplies the first z flux
    // cp.async: every thread overwrites only cells whose last readers were its own z solve (centre) and x/y solves >= 1 XY
    // barrier ago (halo), so the load goes out at once.  TMA: ONE thread overwrites the whole slot, so it waits until every
    // thread is past the z solve of the previous plane — the FREE barrier it waits for anyway before its first flux store
    // (the chunk's leading plane, which has no x/y faces, loads into a slot nobody has used yet).
    if (k < kend && (!TMA || !xy)) issue_load(k + H + 1, sload);
    if (k < kend) issue_patch(k + H + 1);
    const double* const pk = ring + sk * G::PLANE;
    const long long cg = cg0 + (long long)(k + 1) * gplane;
    if (!UB_EARLY && GX_STAGE_UB_PREFETCH && xy && cell_ok && Ub != S) {                       // second stage: the base state is not the staged array; start it on its way
#pragma unroll
      for (int q = 0; q < NU; ++q) if (!(FLUXCD && q >= 5 && q < 8)) prefetch_ub(Ub + q * vs + cg);
    }
    // jobs of this thread: x face (0), y face (1), upper z face (2; main warps only)
#pragma unroll 1
    for (int jt = (xy ? 0 : 2); jt < njobs; ++jt) {
      double wl[8], wr[8], fr[8], csl = 0.0, csr = 0.0;
      const double *pa, *pb, *pc, *pd;                    // variable 0 of cells l-1, l | r, r+1 of this interface (passive scalars)
      if (jt == 0) {
        const double* c = pk + c0x;
        gather<0, LIM, ORDER, NQ, NU, PC, PRE>(c - 2, c - 1, c, c + 1, wl, wr, csl, csr);
        if (NPAS) { pa = c - 2; pb = c - 1; pc = c; pd = c + 1; }
      } else if (jt == 1) {
        const double* c = pk + c0y;
        gather<1, LIM, ORDER, NQ, NU, PC, PRE>(c - 2 * CX, c - CX, c, c + CX, wl, wr, csl, csr);
        if (NPAS) { pa = c - 2 * CX; pb = c - CX; pc = c; pd = c + CX; }
      } else {
        if (GX_STAGE_CONVERT_EARLY && k < kend) { wait_load(sload, lpar); convert(sload); }
        const double* cm = (ORDER == 2) ? ring + slot_add(sk, NSLOT - 1) * G::PLANE + cidx : pk + cidx;
        const double* cp1 = ring + slot_add(sk, 1) * G::PLANE + cidx;
        const double* cp2 = (ORDER == 2) ? ring + slot_add(sk, 2) * G::PLANE + cidx : cp1;
        gather<2, LIM, ORDER, NQ, NU, PC, PRE>(cm, pk + cidx, cp1, cp2, wl, wr, csl, csr);
        if (NPAS) { pa = cm; pb = pk + cidx; pc = cp1; pd = cp2; }
      }
      double ub[8 + NPAS];
      if (UB_EARLY && jt == 2 && xy) {                    // base state for the update: in flight during the z solve
#pragma unroll
        for (int q = 0; q < NU; ++q) if (!(FLUXCD && q >= 5 && q < 8)) ub[q] = Ub[q * vs + cg];
      }
      gxp::PasInfo I;
      const int e = gxp::riemann<SOLVER, PRE>(A.phys, wl, wr, fr, I, csl, csr);
      double frp[NPAS ? NPAS : 1];                        // passive scalars: limited like every primitive, flux from the
      if (NPAS) {                                         // closed form of the branch the solver took (passive_flux)
#if defined(GX_FLAVOUR_FAST)
        double cl, cr;
        gxp::passive_coeffs(I, cl, cr);
#endif
#pragma unroll
        for (int m = 0; m < NPAS; ++m) {
          const int c = (NQ + m) * PC;
          double ql = pb[c], qr = pc[c];
          if (ORDER == 2) gxp::reconstruct<LIM>(pa[c], ql, qr, pd[c]);
#if defined(GX_FLAVOUR_FAST)
          frp[m] = cl * ql + cr * qr;
#else
          frp[m] = gxp::passive_flux(I, ql, qr);
#endif
        }
      }
      if (jt == 0) {
        err |= check_x ? e : 0;
        if (it > 0) mbar_wait(bar_free, (it - 1) & 1);    // every thread has finished reading the previous plane's fluxes
        if (TMA && k < kend) issue_load(k + H + 1, sload);                    // ... and the z stencil of the previous plane
#pragma unroll
        for (int q = 0; q < NQ; ++q) out_x[rot<0>(q) * ovs_x] = fr[q];
#pragma unroll
        for (int m = 0; m < NPAS; ++m) out_x[(NQ + m) * ovs_x] = frp[m];
      } else if (jt == 1) {
        err |= check_y ? e : 0;
#pragma unroll
        for (int q = 0; q < NQ; ++q) out_y[rot<1>(q) * YBV] = fr[q];
#pragma unroll
        for (int m = 0; m < NPAS; ++m) out_y[(NQ + m) * YBV] = frp[m];
        mbar_arrive(bar_xy);                              // my x and y fluxes are written
      } else {
        err |= cell_ok ? e : 0;
        // ---- epilogue of the z solve: the update of my cell in plane k (fr = flux through its upper z face) ----
        double h[8 + NPAS];
#pragma unroll
        for (int q = 0; q < NQ; ++q) h[rot<2>(q)] = fr[q];
#pragma unroll
        for (int m = 0; m < NPAS; ++m) h[NQ + m] = frp[m];
        if (xy) {
          if (!UB_EARLY && cell_ok) {
#pragma unroll
            for (int q = 0; q < NU; ++q) if (!(FLUXCD && q >= 5 && q < 8)) ub[q] = Ub[q * vs + cg];
          }
          mbar_wait(bar_xy, it & 1);                      // all x/y face fluxes of this plane visible
          if (cell_ok) {
            const long long c = cg;                       // cell_ok: no clamping took place
            const double* xr = xb + wrp * (TX + 1) + lane;
            const double* yr = yb + wrp * TX + lane;
            double un[8 + NPAS];
            // get_user_source_terms as the point-mass gravity functor (EXO/user_mod.f90:158-206; same expressions as k_update),
            // evaluated on the primitives of the staged state: prim(u^n) in the first stage, prim(up) in the second
            double src[5] = {0., 0., 0., 0., 0.};
            const bool with_src = NPAS && A.user_src && A.grav.n > 0;
            if (with_src) {
              const double* wc = pk + cidx;
              const double r0 = wc[0], v1 = wc[PC], v2 = wc[2 * PC], v3 = wc[3 * PC];
              const double xc = ((double)(i + g.cx * g.nx - g.nxtot / 2) - 0.5) * g.dx;
              const double yc = ((double)(j + g.cy * g.ny - g.nytot / 2) - 0.5) * g.dy;
              const double zc = ((double)(k + g.cz * g.nz - g.nztot / 2) - 0.5) * g.dz;
              for (int l = 0; l < A.grav.n; ++l) {
                const double x = xc - A.grav.x[l], y = yc - A.grav.y[l], z = zc - A.grav.z[l];
                const double rad2 = x * x + y * y + z * z;
#if defined(GX_FLAVOUR_FAST)
                const double ir = gxp::fast_rsqrt(rad2);                       // rad2**(-1.5) without pow and without divisions
                const double k3 = r0 * A.grav.gm[l] * (ir * ir * ir);
                src[1] = src[1] - k3 * x;
                src[2] = src[2] - k3 * y;
                src[3] = src[3] - k3 * z;
                src[4] = src[4] - k3 * (v1 * x + v2 * y + v3 * z);
#else
                const double r15 = pow(rad2, 1.5);
                src[1] = src[1] - r0 * A.grav.gm[l] * x / r15;
                src[2] = src[2] - r0 * A.grav.gm[l] * y / r15;
                src[3] = src[3] - r0 * A.grav.gm[l] * z / r15;
                src[4] = src[4] - r0 * A.grav.gm[l] * (v1 * x + v2 * y + v3 * z) / r15;
#endif
              }
            }
#pragma unroll
            for (int q = 0; q < NU; ++q) {
              if (FLUXCD && q >= 5 && q < 8) continue;    // B is advanced from E by k_bupdate
              const double flo = xr[q * XBV], fup = xr[q * XBV + 1];
              const double glo = yr[q * YBV], gup = yr[q * YBV + TX];
              // step(): up = u - dt/dx (f(i)-f(i-1)) - dt/dy (g(j)-g(j-1)) - dt/dz (h(k)-h(k-1))   hydro_solver.f90:105-107
              double v = ub[q] - sdt.dtdx * (fup - flo) - sdt.dtdy * (gup - glo) - sdt.dtdz * (h[q] - hprev[q]);
              if (with_src && q < 5) v = v + sdt.dt * src[q];                  // up = up + dt*s  hydro_solver.f90:115-121
              un[q] = v;
              dst[q * vs + c] = v;
            }
            if (FLUXCD) {                                 // get_efield, flux_cd_module.f90:258-265
              const double f6l = xr[6 * XBV], f6u = xr[6 * XBV + 1];
              const double f7l = xr[7 * XBV], f7u = xr[7 * XBV + 1];
              const double g5l = yr[5 * YBV], g5u = yr[5 * YBV + TX];
              const double g7l = yr[7 * YBV], g7u = yr[7 * YBV + TX];
              E[0 * vs + c] = 0.25 * (-g7l - g7u + hprev[6] + h[6]);
              E[1 * vs + c] = 0.25 * (+f7l + f7u - hprev[5] - h[5]);
              E[2 * vs + c] = 0.25 * (-f6l - f6u + g5l + g5u);
            } else if (want_cfl) {                        // get_timestep candidates of the new state, hydro_core.f90:644-675
              double w[8], Tk;
              double u8[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) u8[q] = q < NQ ? un[q] : 0.0;
              gxp::u2prim<MHD, false, NPAS == 0>(A.phys, u8, w, NPAS ? un[NQ] : 0.0, Tk);
              if (MHD) {
                double cx, cy, cz;
                gxp::cfast3(A.phys, w[4], w[0], w[5], w[6], w[7], cx, cy, cz);
                dtp = fmin(dtp, g.dx / (fabs(w[1]) + cx));
                dtp = fmin(dtp, g.dy / (fabs(w[2]) + cy));
                dtp = fmin(dtp, g.dz / (fabs(w[3]) + cz));
              } else {
                const double cs = gxp::csound(A.phys, w[4], w[0]);
                dtp = fmin(dtp, g.dx / (fabs(w[1]) + cs));
                dtp = fmin(dtp, g.dy / (fabs(w[2]) + cs));
                dtp = fmin(dtp, g.dz / (fabs(w[3]) + cs));
              }
            }
          }
          if (TMA) fence_proxy_async();
          mbar_arrive(bar_free);                          // done reading the exchange buffers (and the oldest plane of the z stencil)
        }
#pragma unroll
        for (int q = 0; q < NU; ++q) hprev[q] = h[q];
      }
    }
    if (GX_STAGE_CONVERT_EARLY && !main_warp && k < kend) { wait_load(sload, lpar); convert(sload); }
    if (!main_warp && xy) {                               // the closing warp reads no fluxes; it keeps in step with the planes
      mbar_wait(bar_xy, it & 1);
      if (TMA) fence_proxy_async();
      mbar_arrive(bar_free);
    }
    if (xy) ++it;
    if (!GX_STAGE_CONVERT_EARLY && k < kend) { wait_load(sload, lpar); convert(sload); }    // next plane -> primitives (own cells)
    if (TMA && !xy) { fence_proxy_async(); __syncthreads(); }    // leading plane: its z solves are done before the next tile load (no FREE phase yet)
  }
  __syncthreads();
  if (err) atomicOr(errflag, 1);
  if (!FLUXCD && want_cfl) stage_block_min<NT>(dtp, dtmin_bits, scr);
}

// ---------------------------------------------------------------------------
// planes per CTA: as long as possible (the z prologue of a chunk costs 2*ORDER plane loads and one extra z solve) while the
// grid still fills whole waves of SMs (one CTA per SM)
static int auto_kz(long long tiles, int nplanes) {
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148; }
  int best_kz = nplanes; double best_eff = -1.0;
  for (int chunks = 1; chunks <= nplanes; ++chunks) {
    const int kzc = (nplanes + chunks - 1) / chunks;
    const long long total = tiles * ((nplanes + kzc - 1) / kzc);
    const double eff = (double)total / (double)(((total + sms - 1) / sms) * sms);
    if (eff > best_eff + 1e-9) { best_eff = eff; best_kz = kzc; }
    if (eff >= 0.95 && total >= 2LL * sms) { best_kz = kzc; break; }
    if (kzc <= 4) break;
  }
  return best_kz;
}

// ---- TMA descriptors of the staged arrays: 4-D tensor (x, y, plane, variable) over the SoA layout, box = one staged plane ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
struct TmapKey { const void* base; int px, py, pz, nu, cx, ry; long long vs; };
struct TmapEntry { TmapKey k; CUtensorMap m; };
static const CUtensorMap* stage_tensor_map(const double* S, const Grid& g, int nu, int cx, int ry) {
  static std::mutex mu;
  static std::vector<TmapEntry*> cache;
  static EncodeTiledFn encode = nullptr;
  static bool tried = false;
  std::lock_guard<std::mutex> lock(mu);
  if (!tried) {
    tried = true;
    cudaDriverEntryPointQueryResult qr;
    void* fn = nullptr;
    if (!getenv("GX_NO_TMA") && cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
      encode = (EncodeTiledFn)fn;
    else cudaGetLastError();
  }
  if (!encode) return nullptr;
  for (TmapEntry* e : cache)
    if (e->k.base == S && e->k.px == g.px && e->k.py == g.py && e->k.pz == g.pz && e->k.nu == nu && e->k.cx == cx && e->k.ry == ry && e->k.vs == g.vs) return &e->m;
  TmapEntry* e = new TmapEntry();
  e->k = TmapKey{S, g.px, g.py, g.pz, nu, cx, ry, g.vs};
  const cuuint64_t dims[4] = {(cuuint64_t)g.px, (cuuint64_t)g.py, (cuuint64_t)g.pz, (cuuint64_t)nu};
  const cuuint64_t strides[3] = {(cuuint64_t)g.px * 8, (cuuint64_t)g.px * g.py * 8, (cuuint64_t)g.vs * 8};
  const cuuint32_t box[4] = {(cuuint32_t)cx, (cuuint32_t)ry, 1, (cuuint32_t)nu};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = encode(&e->m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<double*>(S), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { delete e; return nullptr; }
  if (cache.size() > 64) { for (TmapEntry* o : cache) delete o; cache.clear(); }     // arrays come and go with their solvers
  cache.push_back(e);
  return &e->m;
}

template <int SOLVER, int LIM, int ORDER, bool FLUXCD, int NPAS>
static int launch_one(const StepArgs& A, double dt, const double* S, const double* Ub, double* dst, double* E, int kz,
                      unsigned long long* dtmin_bits, int want_cfl, int* errflag, cudaStream_t st) {
  using G = typename StageTraits<SOLVER, LIM, ORDER, FLUXCD, NPAS>::G;
  static_assert(G::SMEM + 64 <= 227 * 1024, "stage kernel tile does not fit the shared memory of one SM");
  const Grid& g = A.g;
  // TMA tile loads (one per plane) on the headline kernels.  TMA reads ghost cells as they are; where x / y are wrapped in the
  // loaders the periodic-image cells of the edge tiles come through the per-thread patch, z wraps through the box coordinate
  const CUtensorMap* tm = nullptr;
  const int tiles_x = (g.nx + G::TX - 1) / G::TX, tiles_y = (g.ny + G::TY - 1) / G::TY;
  // (a tile that holds BOTH periodic ends of a direction would need two patches: such small blocks keep the cp.async loader)
  // Measured at 256^3 (B200): second-order stage 1.69 ms with TMA against 1.79 with cp.async; first-order stage 1.32 against 1.30
  // (it stages 3 planes of 13 rows, where the per-thread loads are already cheap) — so A.tma = 1 means the second-order stage only.
  if (NPAS == 0 && (A.tma >= 2 || (A.tma == 1 && ORDER == 2)) && (!A.wrap[0] || tiles_x >= 2) && (!A.wrap[1] || tiles_y >= 2))
    tm = stage_tensor_map(S, g, G::NU, G::CX, G::RY);
  static const CUtensorMap no_map = {};
  auto kern = tm ? k_stage<SOLVER, LIM, ORDER, FLUXCD, NPAS, NPAS == 0> : k_stage<SOLVER, LIM, ORDER, FLUXCD, NPAS, false>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM) != cudaSuccess) return GX_ECUDA;
  const StageDt sdt = {dt / g.dx, dt / g.dy, dt / g.dz, dt};
  const int tx = (g.nx + G::TX - 1) / G::TX, ty = (g.ny + G::TY - 1) / G::TY, nplanes = A.klast - A.kbeg + 1;
  const int nplanes2 = A.klast2 >= A.kbeg2 ? A.klast2 - A.kbeg2 + 1 : 0;
  if (kz <= 0) kz = auto_kz((long long)tx * ty * (nplanes2 ? 2 : 1), nplanes);            // kz > 0: the caller's choice (GX_KZ)
  dim3 grid(tx, ty, (nplanes + kz - 1) / kz + (nplanes2 + kz - 1) / kz);
  kern<<<grid, G::NT, G::SMEM, st>>>(A, sdt, S, Ub, dst, E, kz, dtmin_bits, want_cfl, errflag, tm ? *tm : no_map);
  return GX_OK;
}

template <int SOLVER, int LIM, int ORDER>
static int launch_cd(const StepArgs& A, double dt, const double* S, const double* Ub, double* dst, double* E, int kz,
                     unsigned long long* dtmin_bits, int want_cfl, int* errflag, cudaStream_t st) {
  constexpr bool MHD = (SOLVER == GX_SOLVER_HLLE || SOLVER == GX_SOLVER_HLLD);
  if (A.g.npas == GX_STAGE_NPAS) {               // the kernels that carry passive scalars, the run-time equation of state and the gravity functor
    if (MHD && A.flux_cd) return launch_one<SOLVER, LIM, ORDER, MHD, GX_STAGE_NPAS>(A, dt, S, Ub, dst, E, kz, dtmin_bits, want_cfl, errflag, st);
    return launch_one<SOLVER, LIM, ORDER, false, GX_STAGE_NPAS>(A, dt, S, Ub, dst, E, kz, dtmin_bits, want_cfl, errflag, st);
  }
  if (A.g.npas != 0) return GX_EUNSUPPORTED;
  if (MHD && A.flux_cd) return launch_one<SOLVER, LIM, ORDER, MHD, 0>(A, dt, S, Ub, dst, E, kz, dtmin_bits, want_cfl, errflag, st);
  return launch_one<SOLVER, LIM, ORDER, false, 0>(A, dt, S, Ub, dst, E, kz, dtmin_bits, want_cfl, errflag, st);
}

#define GX_CAT2(a, b) a##b
#define GX_CAT(a, b) GX_CAT2(a, b)
// exported per (flavour, solver): l_stage_<solver>
int GX_CAT(l_stage_, GX_STAGE_SOLVER)(const StepArgs& A, int order, double dt, const double* S, const double* Ub, double* dst,
                                      double* E, int kz, unsigned long long* dtmin_bits, int want_cfl, int* errflag, cudaStream_t st) {
  constexpr int SV = GX_STAGE_SOLVER;
  if (order == 1) return launch_cd<SV, GX_LIMITER_NO_AVERAGE, 1>(A, dt, S, Ub, dst, E, kz, dtmin_bits, want_cfl, errflag, st);
  switch (A.limiter) {
#ifndef GX_DEV_MINMOD_ONLY
    case GX_LIMITER_NO_AVERAGE: return launch_cd<SV, GX_LIMITER_NO_AVERAGE, 2>(A, dt, S, Ub, dst, E, kz, dtmin_bits, want_cfl, errflag, st);
    case GX_LIMITER_NO_LIMIT:   return launch_cd<SV, GX_LIMITER_NO_LIMIT, 2>(A, dt, S, Ub, dst, E, kz, dtmin_bits, want_cfl, errflag, st);
    case GX_LIMITER_VAN_LEER:   return launch_cd<SV, GX_LIMITER_VAN_LEER, 2>(A, dt, S, Ub, dst, E, kz, dtmin_bits, want_cfl, errflag, st);
    case GX_LIMITER_VAN_ALBADA: return launch_cd<SV, GX_LIMITER_VAN_ALBADA, 2>(A, dt, S, Ub, dst, E, kz, dtmin_bits, want_cfl, errflag, st);
    case GX_LIMITER_UMIST:      return launch_cd<SV, GX_LIMITER_UMIST, 2>(A, dt, S, Ub, dst, E, kz, dtmin_bits, want_cfl, errflag, st);
    case GX_LIMITER_WOODWARD:   return launch_cd<SV, GX_LIMITER_WOODWARD, 2>(A, dt, S, Ub, dst, E, kz, dtmin_bits, want_cfl, errflag, st);
    case GX_LIMITER_SUPERBEE:   return launch_cd<SV, GX_LIMITER_SUPERBEE, 2>(A, dt, S, Ub, dst, E, kz, dtmin_bits, want_cfl, errflag, st);
#endif
    case GX_LIMITER_MINMOD:     return launch_cd<SV, GX_LIMITER_MINMOD, 2>(A, dt, S, Ub, dst, E, kz, dtmin_bits, want_cfl, errflag, st);
  }
  return GX_EUNSUPPORTED;
}

}  // namespace GX_NS
}  // namespace gx
