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
//   * a CTA owns a TX x TY = 32 x 7 column of cells (8 warps, one CTA per SM: 296 CTAs = 2 per SM at 256^3) and
//     marches along z through KZ planes;
//   * conserved planes are staged into shared memory with cp.async (LDGSTS, 8-byte
//     elements: tile rows start on odd 8-byte offsets once the halo is included) one
//     plane ahead of the compute, converted in place to primitives, and kept in a ring
//     of 2*ORDER planes (the z stencil) + 1 in flight;
//   * warp r (< TY) owns row r of the tile, lane l owns cell i0+l.  Each thread solves the
//     LOWER x face, the LOWER y face of its cell and then the UPPER z face (whose flux is
//     carried in registers to the next plane), so every interface is solved exactly once
//     inside the tile; warp TY solves the tile's closing faces (the y faces above the
//     last row, then the x faces right of the last column);
//   * face fluxes are exchanged through shared memory and the update is written with
//     fully coalesced 256-byte row segments.
// Compiled per (flavour, solver): -DGX_FLAVOUR_STRICT|-DGX_FLAVOUR_FAST, -DGX_STAGE_SOLVER=n.
#define GX_SOLVE_MASK 0xffffffffu   // every interface solve of this kernel is executed by all 32 lanes of a warp
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
#ifndef GX_STAGE2_MINBLOCKS
#define GX_STAGE2_MINBLOCKS 1
#endif
#ifndef GX_STAGE1_MINBLOCKS
#define GX_STAGE1_MINBLOCKS 1
#endif

namespace gx {
namespace GX_NS {

// storage component of rotated slot c for sweep direction D (swapy / swapz as an index map,
// src/hydro_core.f90:485-534)
template <int D> __device__ __forceinline__ constexpr int rot(int c) {
  return (c == 1) ? 1 + D : (c == 1 + D) ? 1 : (c == 5) ? 5 + D : (c == 5 + D) ? 5 : c;
}

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// Split CTA barriers on mbarriers (arrive early, wait late): with one CTA per SM a full
// __syncthreads idles the SM, so every hand-over in the plane loop is an arrive followed,
// as late as the data dependence allows, by a parity wait.  All NT threads arrive once per
// plane on each barrier; phase parity = plane counter & 1.
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, int parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "MBAR_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra MBAR_WAIT_%=;\n\t}"
      ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}

template <int NQ_, int ORDER_>
struct StageGeom {
  static constexpr int NQ = NQ_, H = ORDER_;
  static constexpr int TX = GX_STAGE_TX, TY = GX_STAGE_TY;
  static constexpr int NW = TY + 1, NT = NW * 32;
  static constexpr int CX = TX + 2 * H;          // staged columns  i0-H .. i0+TX+H-1
  static constexpr int RY = TY + 2 * H;          // staged rows     j0-H .. j0+TY+H-1
  static constexpr int NSLOT = 2 * H + 1;        // z ring: 2H planes of stencil + 1 in flight
  static constexpr int PCELLS = CX * RY;
  static constexpr int PLANE = NQ * PCELLS;      // doubles per ring slot
  static constexpr int XB = NQ * TY * (TX + 1);  // x-face flux exchange
  static constexpr int YB = NQ * (TY + 1) * TX;  // y-face flux exchange
  static constexpr int ZB = NQ * TY * TX;        // z-face flux hand-over (thread-private slots)
  static constexpr size_t SMEM = sizeof(double) * ((size_t)NSLOT * PLANE + XB + YB + ZB);
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

// One interface with a RUN-TIME sweep direction: the 2*ORDER states are gathered from the
// staged primitive planes through per-direction variable offsets (swapy/swapz,
// src/hydro_core.f90:485-534, as an index map), reconstructed (src/hydro_core.f90:712-798),
// solved (prim2fhll*) and the flux is scattered back through the same map.  Keeping the
// direction a run-time value means ONE copy of the solver in the instruction stream for
// all x, y and z faces (the kernel is instruction-cache bound otherwise).
//   o_m2..o_p1 : offsets (doubles, inside the ring) of variable 0 of cells  l-1, l | r, r+1
//   vn,vt1,vt2 : offsets of the normal / transverse velocity components (bn.. = vn.. + 4 planes)
struct FaceJob {
  int o_m2, o_m1, o_p0, o_p1;      // ring offsets of the four cells
  int vn, vt1, vt2;                // rotated velocity components, in doubles (component * PCELLS)
  int out, ovs;                    // exchange-buffer offset of variable 0, variable stride
  int on, ot1, ot2;                // rotated components of the output (component index)
  bool store, check;
};

template <int SOLVER, int LIM, int ORDER, int NQ, int PC>
__device__ __forceinline__ int solve_job(const gxp::Phys& P, const double* ring, double* xch, const FaceJob& J,
                                         unsigned long long* bar_free, int free_parity) {
  double wl[8], wr[8], fr[8];
  {
    const int off[8] = {0, J.vn, J.vt1, J.vt2, 4 * PC, J.vn + 4 * PC, J.vt1 + 4 * PC, J.vt2 + 4 * PC};
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      double pl = ring[J.o_m1 + off[q]], pr = ring[J.o_p0 + off[q]];
      if (ORDER == 2) gxp::reconstruct<LIM>(ring[J.o_m2 + off[q]], pl, pr, ring[J.o_p1 + off[q]]);
      wl[q] = pl; wr[q] = pr;
    }
  }
  gxp::PasInfo I;
  const int err = gxp::riemann<SOLVER>(P, wl, wr, fr, I);
  if (free_parity >= 0) mbar_wait(bar_free, free_parity);   // every warp has finished reading the previous plane's fluxes
  {   // lanes of the closing warp that own no face write to a scratch word instead of branching around the stores
    __shared__ double s_dummy[32];
    double* o = J.store ? xch + J.out : s_dummy + (threadIdx.x & 31);
    const int ovs = J.store ? J.ovs : 0;
    const int oc[8] = {0, J.on, J.ot1, J.ot2, 4, J.on + 4, J.ot1 + 4, J.ot2 + 4};
#pragma unroll
    for (int q = 0; q < NQ; ++q) o[oc[q] * ovs] = fr[q];
  }
  return J.check ? err : 0;
}

template <int SOLVER, int LIM, int ORDER, bool FLUXCD>
__global__ void __launch_bounds__(StageGeom<(SOLVER == GX_SOLVER_HLLE || SOLVER == GX_SOLVER_HLLD) ? 8 : 5, ORDER>::NT, (ORDER == 1 ? GX_STAGE1_MINBLOCKS : GX_STAGE2_MINBLOCKS))
k_stage(const StepArgs A, const double dt, const double* __restrict__ S, const double* Ub, double* dst,
        double* __restrict__ E, const int kz, unsigned long long* dtmin_bits, const int want_cfl, int* errflag) {
  constexpr bool MHD = (SOLVER == GX_SOLVER_HLLE || SOLVER == GX_SOLVER_HLLD);
  constexpr int NQ = MHD ? 8 : 5;
  using G = StageGeom<NQ, ORDER>;
  constexpr int H = G::H, TX = G::TX, TY = G::TY, CX = G::CX, NSLOT = G::NSLOT, NT = G::NT;
  constexpr int PC = G::PCELLS;
  constexpr int XBV = TY * (TX + 1), YBV = (TY + 1) * TX, ZBV = TY * TX;
  extern __shared__ double sm[];
  double* const ring = sm;
  double* const xch = sm + (size_t)NSLOT * G::PLANE;     // exchange buffers: x | y | z
  constexpr int XB0 = 0, YB0 = G::XB, ZB0 = G::XB + G::YB;
  const double* const xb = xch + XB0;                    // [q][TY][TX+1]
  const double* const yb = xch + YB0;                    // [q][TY+1][TX]
  const double* const zb = xch + ZB0;                    // [q][TY][TX]

  const Grid& g = A.g;
  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  const int i0 = 1 + (int)blockIdx.x * TX, j0 = 1 + (int)blockIdx.y * TY, k0 = A.kbeg + (int)blockIdx.z * kz;
  const int kend = min(k0 + kz - 1, A.klast);
  const long long vs = g.vs;

  // ---- plane staging: conserved -> shared (cp.async), converted in place to primitives ----
  // Ownership: a main-warp thread stages and converts ITS OWN centre cell (so its z solves, which
  // only read its own column, never wait for another thread), and the halo frame of the plane
  // (consumed ORDER+1 planes later by the x/y solves) is dealt round-robin.
  // warp-uniform by construction; in the first-order kernel telling the compiler so (a vote) pays, in the second-order one it does not
  const bool main_warp = (ORDER == 1) ? (bool)__all_sync(0xffffffffu, wrp < TY) : (wrp < TY);
  const int cidx = (min(wrp, TY - 1) + H) * CX + (lane + H);   // this thread's cell inside a staged plane
  constexpr int HC = PC - TY * TX;                       // halo cells of a staged plane
  auto halo_cell = [&](int h) {                          // h-th halo cell -> plane-local index
    if (h < H * CX) return h;                                                   // rows below the tile
    if (h < H * CX + 2 * H * TY) {
      const int t = h - H * CX, r = t / (2 * H), sx = t - r * (2 * H);
      return (H + r) * CX + (sx < H ? sx : TX + sx);                            // left | right columns
    }
    return h - (H * CX + 2 * H * TY) + (H + TY) * CX;                           // rows above the tile
  };
  auto slot_off = [&](int p) { return ((p - (k0 - H)) % NSLOT) * G::PLANE; };
  // Every thread stages (and later converts) at most two cells of a plane: its own centre cell and one cell of
  // the halo frame.  Their plane-local indices and their (i, j) offsets inside a global plane — clamped to the
  // array, wrapped where the block is its own periodic neighbour — are fixed for the whole march: compute them once.
  static_assert(HC <= NT, "one halo cell per thread");
  const bool has_halo = tid < HC;
  const int hcell = halo_cell(has_halo ? tid : 0);
  auto ij_off = [&](int c) {
    const int rr = c / CX, cc = c - rr * CX;
    int i = min(i0 - H + cc, g.nx + 2), j = min(j0 - H + rr, g.ny + 2);
    if (A.wrap[0]) i = i < 1 ? i + g.nx : (i > g.nx ? i - g.nx : i);
    if (A.wrap[1]) j = j < 1 ? j + g.ny : (j > g.ny ? j - g.ny : j);
    return (j + 1) * g.px + (i + g.xo);
  };
  const int own_off = ij_off(cidx), halo_off = ij_off(hcell);
  const long long gplane = (long long)g.px * g.py;
  auto stage_cell = [&](double* sl, int c, const double* src) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) cp_async8(sl + q * PC + c, src + q * vs);
  };
  auto issue_load = [&](int p) {
    double* sl = ring + slot_off(p);
    int kk = min(p, g.nz + 2);
    if (A.wrap[2]) kk = kk < 1 ? kk + g.nz : (kk > g.nz ? kk - g.nz : kk);
    const double* pl = S + (long long)(kk + 1) * gplane;
    if (main_warp) stage_cell(sl, cidx, pl + own_off);
    if (has_halo) stage_cell(sl, hcell, pl + halo_off);
    cp_async_commit();
  };
  auto convert_cell = [&](double* sl, int c) {
    double u[8], w[8], T;
#pragma unroll
    for (int q = 0; q < NQ; ++q) u[q] = sl[q * PC + c];
    gxp::u2prim<MHD, false, true>(A.phys, u, w, 0.0, T);
#pragma unroll
    for (int q = 0; q < NQ; ++q) sl[q * PC + c] = w[q];
  };
  auto convert = [&](int p) {      // each thread converts exactly the cells it staged itself
    double* sl = ring + slot_off(p);
    if (main_warp) convert_cell(sl, cidx);
    if (has_halo) convert_cell(sl, hcell);
  };

  // Barriers of the plane loop (all NT threads arrive once per plane on each):
  //   bars[0] XY   : x and y face fluxes of this plane are in the exchange buffers
  //   bars[1] FREE : this thread has consumed the exchange buffers (they may be overwritten)
  // The z flux is handed over in thread-private slots and the newest plane's centre cell is
  // converted by its consumer, so neither needs a CTA-wide barrier.
  __shared__ unsigned long long bars[2];
  if (tid == 0) { mbar_init(&bars[0], NT); mbar_init(&bars[1], NT); }
#pragma unroll 1
  for (int p = k0 - H; p <= k0 + H - 1; ++p) issue_load(p);
  cp_async_wait_all();
#pragma unroll 1
  for (int p = k0 - H; p <= k0 + H - 1; ++p) convert(p);
  __syncthreads();

  const int i = i0 + lane, j = j0 + wrp;
  const bool cell_ok = main_warp && i <= g.nx && j <= g.ny;
  const double dtdx = dt / g.dx, dtdy = dt / g.dy, dtdz = dt / g.dz;
  // plane-invariant parts of this thread's three face jobs (x: extra warp closes right of the last column, row = lane;
  // y: extra warp closes the row above the tile)
  const int row_x = main_warp ? wrp : min(lane, TY - 1), col_x = main_warp ? lane : TX;
  const int cx_const = (row_x + H) * CX + (col_x + H), cy_const = cidx + (main_warp ? 0 : CX);
  const int out_x = XB0 + row_x * (TX + 1) + col_x;
  const bool store_x = main_warp || lane < TY;
  const bool check_x = main_warp ? (i <= g.nx + 1 && j <= g.ny) : (lane < TY && i0 + TX <= g.nx + 1 && j0 + lane <= g.ny);
  const bool check_y = (i <= g.nx && j <= g.ny + 1);
  double hprev[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) hprev[q] = 0.0;
  double dtp = 1.e30;
  int err = 0;
  int it = 0;                                             // plane counter (barrier phase)

  for (int k = k0 - 1; k <= kend; ++k, ++it) {
    if (k < kend) issue_load(k + H + 1);                  // into the slot of plane k-H: its last readers were this thread's
                                                          // own z solve (centre) and x/y solves >= 1 XY barrier ago (halo)
    const int sk = slot_off(k);
    const bool xy = k >= k0;
    // jobs of this thread: main warps solve the lower y face, the lower x face and the upper z face
    // of their cell; warp TY closes the tile (y faces above the last row, x faces right of the last column)
    const int njobs = main_warp ? 3 : 2;
    const int sp1 = slot_off(k + 1), sm1 = (ORDER == 2) ? slot_off(k - 1) : sk, sp2 = (ORDER == 2) ? slot_off(k + 2) : sp1;
    double ub[8];
    const long long cg = g.idx(min(i, g.nx), min(j, g.ny), max(k, 1));
#pragma unroll 1
    for (int jb = (xy ? 0 : 2); jb < njobs; ++jb) {
      // branch-free job descriptor: the job type only steers selects (no divergent if/else chain per solve)
      FaceJob J;
      const int jt = jb == 0 ? 1 : (jb == 1 ? 0 : 2);     // job order: x face, y face, z face
      const bool isx = jt == 1, isz = jt == 2;
      const int cxy = isx ? sk + cx_const : sk + cy_const;
      const int st = isx ? 1 : CX;
      J.o_p0 = isz ? sp1 + cidx : cxy;
      J.o_m1 = isz ? sk + cidx : cxy - st;
      J.o_m2 = isz ? ((ORDER == 2) ? sm1 + cidx : sk + cidx) : cxy - 2 * st;
      J.o_p1 = isz ? ((ORDER == 2) ? sp2 + cidx : sp1 + cidx) : cxy + st;
      J.on = isx ? 1 : (isz ? 3 : 2); J.ot1 = isx ? 2 : (isz ? 2 : 1); J.ot2 = isz ? 1 : 3;
      J.out = isx ? out_x : (isz ? ZB0 : YB0) + wrp * TX + lane;
      J.ovs = isx ? XBV : (isz ? ZBV : YBV);
      J.store = isx ? store_x : true;
      J.check = isx ? check_x : (isz ? cell_ok : check_y);
      if (isz && xy) {                                    // base state for the update: in flight during the z solve
#pragma unroll
        for (int q = 0; q < NQ; ++q) if (!(FLUXCD && q >= 5)) ub[q] = Ub[q * vs + cg];
      }
      J.vn = J.on * PC; J.vt1 = J.ot1 * PC; J.vt2 = J.ot2 * PC;
      // the first x/y store of a plane waits until every thread has consumed the previous plane's fluxes
      err |= solve_job<SOLVER, LIM, ORDER, NQ, PC>(A.phys, ring, xch, J, &bars[1], (jb == 0 && it > 0) ? ((it - 1) & 1) : -1);
      if (jb == 1) mbar_arrive(&bars[0]);                 // my x and y fluxes are written
    }
    if (!xy) mbar_arrive(&bars[0]);                       // (no x/y faces on the chunk's leading plane)
    double h[8];
    if (main_warp) {
#pragma unroll
      for (int q = 0; q < NQ; ++q) h[q] = zb[q * ZBV + wrp * TX + lane];
    }
    mbar_wait(&bars[0], it & 1);                          // all x/y face fluxes of this plane visible
    if (xy && cell_ok) {
      const long long c = g.idx(i, j, k);
      double un[8];
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        if (FLUXCD && q >= 5) continue;                   // B is advanced from E by k_bupdate
        const double flo = xb[q * XBV + wrp * (TX + 1) + lane], fup = xb[q * XBV + wrp * (TX + 1) + lane + 1];
        const double glo = yb[q * YBV + wrp * TX + lane], gup = yb[q * YBV + (wrp + 1) * TX + lane];
        // step(): up = u - dt/dx (f(i)-f(i-1)) - dt/dy (g(j)-g(j-1)) - dt/dz (h(k)-h(k-1))   hydro_solver.f90:105-107
        const double v = ub[q] - dtdx * (fup - flo) - dtdy * (gup - glo) - dtdz * (h[q] - hprev[q]);
        un[q] = v;
        dst[q * vs + c] = v;
      }
      if (FLUXCD) {                                       // get_efield, flux_cd_module.f90:258-265
        const double f6l = xb[6 * XBV + wrp * (TX + 1) + lane], f6u = xb[6 * XBV + wrp * (TX + 1) + lane + 1];
        const double f7l = xb[7 * XBV + wrp * (TX + 1) + lane], f7u = xb[7 * XBV + wrp * (TX + 1) + lane + 1];
        const double g5l = yb[5 * YBV + wrp * TX + lane], g5u = yb[5 * YBV + (wrp + 1) * TX + lane];
        const double g7l = yb[7 * YBV + wrp * TX + lane], g7u = yb[7 * YBV + (wrp + 1) * TX + lane];
        E[0 * vs + c] = 0.25 * (-g7l - g7u + hprev[6] + h[6]);
        E[1 * vs + c] = 0.25 * (+f7l + f7u - hprev[5] - h[5]);
        E[2 * vs + c] = 0.25 * (-f6l - f6u + g5l + g5u);
      } else if (want_cfl) {                              // get_timestep candidates of the new state, hydro_core.f90:644-675
        double w[8], T;
        gxp::u2prim<MHD, false, true>(A.phys, un, w, 0.0, T);
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
    mbar_arrive(&bars[1]);                                // done reading the exchange buffers
#pragma unroll
    for (int q = 0; q < NQ; ++q) hprev[q] = h[q];
    if (k < kend) { cp_async_wait_all(); convert(k + H + 1); }   // next plane -> primitives (own cells)
  }
  __syncthreads();
  if (err) atomicOr(errflag, 1);
  if (!FLUXCD && want_cfl) stage_block_min<NT>(dtp, dtmin_bits, xch);
}

// ---------------------------------------------------------------------------
template <int SOLVER, int LIM, int ORDER, bool FLUXCD>
static int launch_one(const StepArgs& A, double dt, const double* S, const double* Ub, double* dst, double* E, int kz,
                      unsigned long long* dtmin_bits, int want_cfl, int* errflag, cudaStream_t st) {
  constexpr bool MHD = (SOLVER == GX_SOLVER_HLLE || SOLVER == GX_SOLVER_HLLD);
  using G = StageGeom<MHD ? 8 : 5, ORDER>;
  auto kern = k_stage<SOLVER, LIM, ORDER, FLUXCD>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM) != cudaSuccess) return GX_ECUDA;
  const Grid& g = A.g;
  dim3 grid((g.nx + G::TX - 1) / G::TX, (g.ny + G::TY - 1) / G::TY, (A.klast - A.kbeg + 1 + kz - 1) / kz);
  kern<<<grid, G::NT, G::SMEM, st>>>(A, dt, S, Ub, dst, E, kz, dtmin_bits, want_cfl, errflag);
  return GX_OK;
}

template <int SOLVER, int LIM, int ORDER>
static int launch_cd(const StepArgs& A, double dt, const double* S, const double* Ub, double* dst, double* E, int kz,
                     unsigned long long* dtmin_bits, int want_cfl, int* errflag, cudaStream_t st) {
  constexpr bool MHD = (SOLVER == GX_SOLVER_HLLE || SOLVER == GX_SOLVER_HLLD);
  if (MHD && A.flux_cd) return launch_one<SOLVER, LIM, ORDER, MHD>(A, dt, S, Ub, dst, E, kz, dtmin_bits, want_cfl, errflag, st);
  return launch_one<SOLVER, LIM, ORDER, false>(A, dt, S, Ub, dst, E, kz, dtmin_bits, want_cfl, errflag, st);
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
