// gx_api.cu — C-ABI of libguacho_gx.so (see include/guacho_gx.h): solver object,
// reference-layout <-> device-SoA conversion, ghost-cell boundaries, step driver,
// NCCL halo exchange and CFL reduction.  No CPU fallback exists on any path.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>
#include <vector>

#include "../../include/guacho_gx.h"
#include "gx_kernels.cuh"

using gx::Grid;
using gx::StepArgs;

static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  g_err = buf;
  return code;
}
#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(GX_ECUDA, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

// ---------------------------------------------------------------------------
// NCCL through dlopen: a single-GPU host never needs libnccl to be present.
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
static NcclApi g_nccl;
static int nccl_load() {
  if (g_nccl.ok) return GX_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) { g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (g_nccl.lib) break; }
  if (!g_nccl.lib) return fail(GX_ECOMM, "cannot dlopen libnccl.so.2: %s", dlerror());
#define LD(f) *(void**)(&g_nccl.f) = dlsym(g_nccl.lib, "nccl" #f); if (!g_nccl.f) return fail(GX_ECOMM, "libnccl lacks nccl" #f)
  LD(GetUniqueId); LD(CommInitRank); LD(CommDestroy); LD(Send); LD(Recv); LD(AllReduce); LD(GroupStart); LD(GroupEnd); LD(GetErrorString);
#undef LD
  g_nccl.ok = true;
  return GX_OK;
}
#define NCCL_TRY(x) do { ncclResult_t r_ = (x); if (r_ != ncclSuccess) return fail(GX_ECOMM, "%s failed: %s", #x, g_nccl.GetErrorString(r_)); } while (0)

// layout of the boundary struct (include/guacho_gx.h): no implicit padding, the same image in C, ctypes and bind(C)
static_assert(offsetof(gx_config, pad_) == 35 * 4 && offsetof(gx_config, dx) == 36 * 4 && sizeof(gx_config) == 36 * 4 + 14 * 8,
              "gx_config layout changed: update include/guacho_gx.h, guacho_b200/config.py and guacho_b200/fortran/guacho_gpu.f90 together");

// ---------------------------------------------------------------------------
struct TimedLaunch { int cls; cudaEvent_t a, b; };

struct gx_solver {
  gx_config cfg;
  StepArgs A;
  const gx::KernelTable* K = nullptr;
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t cstream = nullptr;                             // halo push stream (overlaps the interior launches)
  cudaStream_t xstream = nullptr;                             // PCIe copies of the layout conversion (upload_aos / download_aos)
  cudaEvent_t ev_copy[2] = {nullptr, nullptr}, ev_xpose[2] = {nullptr, nullptr};
  cudaEvent_t ev_bnd = nullptr, ev_comm = nullptr;
  bool overlap = false;                                       // z slabs + peer push: boundary-first launches, exchange on cstream
  double *U = nullptr, *UP = nullptr, *W = nullptr, *F = nullptr, *E = nullptr, *Temp = nullptr, *T = nullptr;
  double* W0 = nullptr;                                       // split-all solver: background primitives (gx_set_background)
  double* PT = nullptr;                                       // thermal conduction: pressure and temperature (2 variables)
  double tc_dt_cond = 0.0; int tc_nsteps = 0;                 // what the reference logs per call (thermal_cond.f90:725)
  double* stage = nullptr; size_t stage_doubles = 0;         // AoS staging for layout conversion
  double* halo_send[6] = {0, 0, 0, 0, 0, 0};                  // packed faces (multi-GPU)
  double* halo_recv[6] = {0, 0, 0, 0, 0, 0};
  size_t halo_doubles[3] = {0, 0, 0};
  struct DevScalars { unsigned long long dtmin_bits; int err; int pad; unsigned long long tc_bits; }* dscal = nullptr;   // device
  DevScalars* hscal = nullptr;                                // pinned host mirror
  bool have_state = false;
  bool ghosts_stale = false;   // self-periodic ghost layers of u/up not materialised since the last fused step
  bool fused = false;      // fused stage kernels (gx_stage.cu); otherwise the pass-per-routine kernels
  int kz = 0;              // planes one CTA of the fused stage kernel marches through (0: the launcher fills whole waves of SMs; GX_KZ overrides)
  double time = 0.0;
  // block topology (mpi_cart_shift results; -1 = MPI_PROC_NULL)
  int nb[3] = {1, 1, 1}, co[3] = {0, 0, 0};
  int nbr[3][2] = {{-1, -1}, {-1, -1}, {-1, -1}};             // neighbour ranks [dir][low/high]
  bool periodic[3] = {false, false, false};
  int bc[3][2];
  // comm
  ncclComm_t comm = nullptr; int rank = 0, nranks = 1;
  // peer-memory halo push for z slabs (CUDA IPC mappings of the neighbours' arrays over NVLink)
  bool p2p = false;
  uint32_t* flags = nullptr;                                  // device: READY/ARRIVE words written by the neighbours
  struct Peer { double *U = nullptr, *UP = nullptr, *E = nullptr; uint32_t* flags = nullptr; } peer[2];   // [lo, hi]
  uint32_t xseq = 0;                                          // exchange sequence number (same on every rank)
  // user functors
  std::vector<gx_wind_sphere> spheres;                        // passed to k_wind_spheres by value (<= GX_MAX_SPHERES)
  gx_host_bc_fn host_bc = nullptr; void* host_bc_user = nullptr;
  gx_bc_hook_fn bc_hook = nullptr; void* bc_hook_user = nullptr;
  gx_host_source_fn host_src = nullptr; void* host_src_user = nullptr;
  // diagnostics
  long long launches = 0;
  bool profiling = false;
  std::vector<TimedLaunch> timed; std::vector<cudaEvent_t> evpool;
  double cls_ms[gx::KC_COUNT] = {0}; long long cls_n[gx::KC_COUNT] = {0};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr; double last_ms = 0.0;
};

// profiling brackets: CUDA events on the solver's own stream around each launch group
struct LaunchScope {
  gx_solver* s; int cls; cudaEvent_t a = nullptr, b = nullptr;
  LaunchScope(gx_solver* s_, int cls_, int n = 1) : s(s_), cls(cls_) {
    s->launches += n;
    if (s->profiling) {
      auto get = [&]() { cudaEvent_t e; if (!s->evpool.empty()) { e = s->evpool.back(); s->evpool.pop_back(); } else cudaEventCreate(&e); return e; };
      a = get(); b = get();
      cudaEventRecord(a, s->stream);
    }
  }
  ~LaunchScope() { if (s->profiling) { cudaEventRecord(b, s->stream); s->timed.push_back({cls, a, b}); } }
};
static void collect_timed(gx_solver* s) {
  if (s->timed.empty()) return;
  cudaStreamSynchronize(s->stream);
  if (s->cstream) cudaStreamSynchronize(s->cstream);
  for (auto& t : s->timed) {
    float ms = 0; cudaEventElapsedTime(&ms, t.a, t.b);
    s->cls_ms[t.cls] += ms; s->cls_n[t.cls] += 1;
    s->evpool.push_back(t.a); s->evpool.push_back(t.b);
  }
  s->timed.clear();
}

// ---------------------------------------------------------------------------
// layout conversion kernels: reference AoS (var fastest) <-> device SoA, z-chunked
__global__ void k_aos_to_soa(Grid g, int nvar, const double* __restrict__ aos, double* __restrict__ soa, int k0, int nk) {
  // aos chunk: (nvar, nx+4, ny+4, nk) column-major, planes k0..k0+nk-1 (0-based padded index)
  const int ii = blockIdx.x * blockDim.x + threadIdx.x;      // 0..nx+3
  const int jj = blockIdx.y, kk = blockIdx.z;
  if (ii >= g.nx + 4) return;
  const long long a = (long long)nvar * (ii + (long long)(g.nx + 4) * (jj + (long long)(g.ny + 4) * kk));
  const long long c = ((long long)(k0 + kk) * g.py + jj) * g.px + (ii - 1 + g.xo);
  for (int q = 0; q < nvar; ++q) soa[q * g.vs + c] = aos[a + q];
}
__global__ void k_soa_to_aos(Grid g, int nvar, const double* __restrict__ soa, double* __restrict__ aos, int k0, int nk) {
  const int ii = blockIdx.x * blockDim.x + threadIdx.x;
  const int jj = blockIdx.y, kk = blockIdx.z;
  if (ii >= g.nx + 4) return;
  const long long a = (long long)nvar * (ii + (long long)(g.nx + 4) * (jj + (long long)(g.ny + 4) * kk));
  const long long c = ((long long)(k0 + kk) * g.py + jj) * g.px + (ii - 1 + g.xo);
  for (int q = 0; q < nvar; ++q) aos[a + q] = soa[q * g.vs + c];
}

// ---------------------------------------------------------------------------
// ghost-cell kernels.  A face "job" copies nl layers across one face of the block.
//   mode 0: periodic wrap inside the block   dst -> dst +/- n
//   mode 1: mirror (outflow / closed)        low: src = 1 - dst ; high: src = 2n+1 - dst
// negvar: variable whose sign flips (closed walls), -1 for none.
// Transverse extent: 1-nl .. n+nl in both transverse directions, which is 0..n+1 for the
// one-layer boundaries (src/boundaries.f90:101-242) and the full array for boundaryII (:316-505).
__global__ void k_bc_face(Grid g, int nvar, double* __restrict__ A, int dir, int side, int mode, int nl, int negvar) {
  const int n[3] = {g.nx, g.ny, g.nz};
  const int ta = dir == 0 ? 1 : 0, tb = dir == 2 ? 1 : 2;     // transverse axes (a fast)
  const int a = (int)(blockIdx.x * blockDim.x + threadIdx.x) + 1 - nl;
  const int b = (int)blockIdx.y + 1 - nl;
  if (a > n[ta] + nl) return;
  const int nd = n[dir];
  for (int l = 0; l < nl; ++l) {
    const int dst = side == 0 ? (1 - nl + l) : (nd + 1 + l);
    const int src = mode == 0 ? (side == 0 ? dst + nd : dst - nd) : (side == 0 ? 1 - dst : 2 * nd + 1 - dst);
    int id[3], is[3];
    id[dir] = dst; is[dir] = src; id[ta] = is[ta] = a; id[tb] = is[tb] = b;
    const long long cd = g.idx(id[0], id[1], id[2]), cs = g.idx(is[0], is[1], is[2]);
    for (int q = 0; q < nvar; ++q) {
      const double v = A[q * g.vs + cs];
      A[q * g.vs + cd] = (q == negvar) ? -v : v;
    }
  }
}

// All faces of a block that has no neighbours in ONE launch over its ghost shell (nl layers).  The reference fills the faces one
// after the other (periodic copies, then closed walls, then outflow: boundaries.f90:101-242, 316-505), every pass over the full
// transverse extent including the ghost layers already filled, so edge and corner ghosts end up with the COMPOSITION of the
// per-direction maps (wrap: dst -/+ n; mirror: 1 - dst | 2n + 1 - dst) whatever the order — which is what a thread evaluates here.
// mode[2*dir+side]: 0 leave alone (user boundary, or a direction the fused kernels wrap in their loaders), 1 periodic wrap,
// 2 mirror; neg[2*dir+side]: variable whose sign flips across that face (closed walls), -1 for none.  Sources are never
// destinations: a source has every index either inside the block or in a direction that is left alone.
struct ShellBc { int mode[6], neg[6]; };
__global__ void __launch_bounds__(128) k_bc_shell(Grid g, int nvar, double* __restrict__ A, int nl, ShellBc bc) {
  const int t = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  const int ex = g.nx + 2 * nl, ey = g.ny + 2 * nl;
  int r = (int)blockIdx.y, i, j, k;
  if (r < 2 * nl * ey) {                                   // the 2 nl z planes: rows along x
    const int pl = r / ey; j = r - pl * ey + 1 - nl; k = pl < nl ? pl + 1 - nl : g.nz + 1 + (pl - nl); i = t + 1 - nl; if (t >= ex) return;
  } else if ((r -= 2 * nl * ey) < 2 * nl * g.nz) {         // the 2 nl y planes between them: rows along x
    const int pl = r / g.nz; k = r - pl * g.nz + 1; j = pl < nl ? pl + 1 - nl : g.ny + 1 + (pl - nl); i = t + 1 - nl; if (t >= ex) return;
  } else {                                                 // the 2 nl x planes: rows along y
    r -= 2 * nl * g.nz; const int pl = r / g.nz; k = r - pl * g.nz + 1; i = pl < nl ? pl + 1 - nl : g.nx + 1 + (pl - nl); j = t + 1; if (t >= g.ny) return;
  }
  const int n[3] = {g.nx, g.ny, g.nz};
  int d[3] = {i, j, k}, sidx[3] = {i, j, k}, flip0 = -1, flip1 = -1, flip2 = -1;
  bool moved = false;
#pragma unroll
  for (int dir = 0; dir < 3; ++dir) {
    const int side = d[dir] < 1 ? 0 : (d[dir] > n[dir] ? 1 : -1);
    if (side < 0) continue;
    const int m = bc.mode[2 * dir + side];
    if (m == 1) { sidx[dir] = side == 0 ? d[dir] + n[dir] : d[dir] - n[dir]; moved = true; }
    else if (m == 2) {
      sidx[dir] = side == 0 ? 1 - d[dir] : 2 * n[dir] + 1 - d[dir]; moved = true;
      const int nv = bc.neg[2 * dir + side];
      if (dir == 0) flip0 = nv; else if (dir == 1) flip1 = nv; else flip2 = nv;
    }
  }
  if (!moved) return;
  const long long cd = g.idx(i, j, k), cs = g.idx(sidx[0], sidx[1], sidx[2]);
  for (int q = 0; q < nvar; ++q) {
    double v = A[q * g.vs + cs];
    if (q == flip0) v = -v;
    if (q == flip1) v = -v;
    if (q == flip2) v = -v;
    A[q * g.vs + cd] = v;
  }
}

// pack / unpack a box [lo,hi] (Fortran indices, inclusive) of nvar variables to/from a contiguous buffer
struct Box { int lo[3], hi[3]; };
__global__ void k_pack(Grid g, int nvar, double* A, double* buf, Box bx, int unpack_flag) {
  const int ex = bx.hi[0] - bx.lo[0] + 1, ey = bx.hi[1] - bx.lo[1] + 1;
  const int ii = blockIdx.x * blockDim.x + threadIdx.x;
  if (ii >= ex) return;
  const int jj = blockIdx.y, kk = blockIdx.z;
  const long long c = g.idx(bx.lo[0] + ii, bx.lo[1] + jj, bx.lo[2] + kk);
  const long long per = (long long)ex * ey * (bx.hi[2] - bx.lo[2] + 1);
  const long long p = ii + (long long)ex * (jj + (long long)ey * kk);
  for (int q = 0; q < nvar; ++q) {
    if (unpack_flag) A[q * g.vs + c] = buf[q * per + p];
    else buf[q * per + p] = A[q * g.vs + c];
  }
}

// impose_user_bc functor: wind spheres (EXO/exoplanet.f90:125-266).  First matching sphere wins
// (the reference tests the star first, then `else if` the planet).
// One launch per sphere over the sphere's bounding box (clipped to the block with its ghost cells); the spheres are applied LAST to
// FIRST, so where two overlap the earlier one ends up in the cells — the reference's `if star ... else if planet`.
__global__ void k_wind_sphere(Grid g, gxp::Phys P, int mhd, const gx_wind_sphere S, double* __restrict__ A, int ilo, int jlo, int klo, int ihi) {
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x) + ilo;
  const int j = (int)blockIdx.y + jlo, k = (int)blockIdx.z + klo;
  if (i > ihi) return;
  const double x = ((double)(i + g.cx * g.nx - g.nxtot / 2) + 0.5) * g.dx;
  const double y = ((double)(j + g.cy * g.ny - g.nytot / 2) + 0.5) * g.dy;
  const double z = ((double)(k + g.cz * g.nz - g.nztot / 2) + 0.5) * g.dz;
  const long long c = g.idx(i, j, k);
  const double xl = x - S.xc, yl = y - S.yc, zl = z - S.zc;
  double rad = sqrt(xl * xl + yl * yl + zl * zl);
  if (rad <= S.radius) {
    if (rad == 0.) rad = g.dx * 0.10;
    const double velx = S.vbx + S.vwind * xl / rad, vely = S.vby + S.vwind * yl / rad, velz = S.vbz + S.vwind * zl / rad;
    const double dens = S.dens;
    A[0 * g.vs + c] = dens;
    A[1 * g.vs + c] = dens * velx;
    A[2 * g.vs + c] = dens * vely;
    A[3 * g.vs + c] = dens * velz;
    double b2h = 0.0;
    if (g.neqdyn == 8) {
      const double q3 = S.radius / rad;
      const double cpi = S.bdip * (q3 * q3 * q3) / (2. * (rad * rad));
      const double bx = 3. * yl * xl * cpi, by = (3. * (yl * yl) - rad * rad) * cpi, bz = 3. * yl * zl * cpi;
      A[5 * g.vs + c] = bx; A[6 * g.vs + c] = by; A[7 * g.vs + c] = bz;
      b2h = 0.5 * (bx * bx + by * by + bz * bz);
    }
    double e = 0.5 * dens * (velx * velx + vely * vely + velz * velz) + P.cv * dens * S.tfac * S.temp;   // exoplanet.f90:181-188,236-243
    if (mhd) e = e + b2h;
    A[4 * g.vs + c] = e;
    for (int q = 0; q < g.npas && q < 4; ++q) A[(long long)(g.neqdyn + q) * g.vs + c] = S.pas[q] * dens;
  }
}
static void launch_wind_spheres(gx_solver* s, double* A) {
  const Grid& g = s->A.g;
  for (int m = (int)s->spheres.size() - 1; m >= 0; --m) {
    const gx_wind_sphere& S = s->spheres[m];
    // cell i has x = (i + cx*nx - nxtot/2 + 0.5) dx: indices whose centre can lie within radius (+1 cell of slack), clipped to -1 .. n+2
    auto range = [](double c, double r, double d, int off, int n, int& lo, int& hi) {
      lo = std::max(-1, (int)floor((c - r) / d - 0.5) - off - 1);
      hi = std::min(n + 2, (int)ceil((c + r) / d - 0.5) - off + 1);
    };
    int ilo, ihi, jlo, jhi, klo, khi;
    range(S.xc, S.radius, g.dx, g.cx * g.nx - g.nxtot / 2, g.nx, ilo, ihi);
    range(S.yc, S.radius, g.dy, g.cy * g.ny - g.nytot / 2, g.ny, jlo, jhi);
    range(S.zc, S.radius, g.dz, g.cz * g.nz - g.nztot / 2, g.nz, klo, khi);
    if (ilo > ihi || jlo > jhi || klo > khi) continue;
    LaunchScope ls(s, gx::KC_BC);
    k_wind_sphere<<<dim3((ihi - ilo + 1 + 63) / 64, jhi - jlo + 1, khi - klo + 1), 64, 0, s->stream>>>(g, s->A.phys, s->cfg.mhd, S, A, ilo, jlo, klo, ihi);
  }
}

// ---------------------------------------------------------------------------
// COOL_H (src/cooling_h.f90): kernel in gx_cooling.cuh, compiled in both flavours (gx_kernels.cu)
static void launch_coolingh(gx_solver* s, double dt_cfl) {
  LaunchScope ls(s, gx::KC_UPDATE);
  s->K->coolingh(s->A, s->cfg.mhd, dt_cfl * s->cfg.tsc, s->U, s->stream);
}

// ---------------------------------------------------------------------------
// Layout conversion pipeline: the staging area is used as two halves; PCIe copies run on a copy stream, the transposes on the
// solver's stream, chained by events, so the copy of chunk c+1 overlaps the transpose of chunk c (round 1 synchronised the host
// after every chunk).  The PCIe copy (1.1 GB at ~50 GB/s for 256^3) remains what the conversion costs.
static int xfer_setup(gx_solver* s) {
  if (s->xstream) return GX_OK;
  CUDA_TRY(cudaStreamCreateWithFlags(&s->xstream, cudaStreamNonBlocking));
  for (int h = 0; h < 2; ++h) {
    CUDA_TRY(cudaEventCreateWithFlags(&s->ev_copy[h], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&s->ev_xpose[h], cudaEventDisableTiming));
  }
  return GX_OK;
}
static int upload_aos(gx_solver* s, const double* host, double* soa, int nvar) {
  const Grid& g = s->A.g;
  int rc = xfer_setup(s); if (rc) return rc;
  const size_t plane = (size_t)nvar * (g.nx + 4) * (g.ny + 4);
  const size_t half = s->stage_doubles / 2;
  const int chunk = (int)std::max<size_t>(1, std::min<size_t>(g.nz + 4, half / plane));
  CUDA_TRY(cudaEventRecord(s->ev_xpose[0], s->stream));          // the copy stream starts after whatever the solver's stream was doing
  CUDA_TRY(cudaStreamWaitEvent(s->xstream, s->ev_xpose[0], 0));
  int c = 0;
  for (int k0 = 0; k0 < g.nz + 4; k0 += chunk, ++c) {
    const int nk = std::min(chunk, g.nz + 4 - k0), h = c & 1;
    double* st = s->stage + (size_t)h * half;
    if (c >= 2) CUDA_TRY(cudaStreamWaitEvent(s->xstream, s->ev_xpose[h], 0));      // the transpose that read this half (chunk c-2) is done
    CUDA_TRY(cudaMemcpyAsync(st, host + plane * k0, plane * nk * sizeof(double), cudaMemcpyHostToDevice, s->xstream));
    CUDA_TRY(cudaEventRecord(s->ev_copy[h], s->xstream));
    CUDA_TRY(cudaStreamWaitEvent(s->stream, s->ev_copy[h], 0));
    {
      LaunchScope ls(s, gx::KC_XPOSE);
      k_aos_to_soa<<<dim3((g.nx + 4 + 127) / 128, g.ny + 4, nk), 128, 0, s->stream>>>(g, nvar, st, soa, k0, nk);
    }
    CUDA_TRY(cudaEventRecord(s->ev_xpose[h], s->stream));
  }
  CUDA_TRY(cudaGetLastError());
  return GX_OK;
}
static int download_aos(gx_solver* s, const double* soa, double* host, int nvar) {
  const Grid& g = s->A.g;
  int rc = xfer_setup(s); if (rc) return rc;
  const size_t plane = (size_t)nvar * (g.nx + 4) * (g.ny + 4);
  const size_t half = s->stage_doubles / 2;
  const int chunk = (int)std::max<size_t>(1, std::min<size_t>(g.nz + 4, half / plane));
  int c = 0;
  for (int k0 = 0; k0 < g.nz + 4; k0 += chunk, ++c) {
    const int nk = std::min(chunk, g.nz + 4 - k0), h = c & 1;
    double* st = s->stage + (size_t)h * half;
    if (c >= 2) CUDA_TRY(cudaStreamWaitEvent(s->stream, s->ev_copy[h], 0));        // the copy that read this half (chunk c-2) is done
    {
      LaunchScope ls(s, gx::KC_XPOSE);
      k_soa_to_aos<<<dim3((g.nx + 4 + 127) / 128, g.ny + 4, nk), 128, 0, s->stream>>>(g, nvar, soa, st, k0, nk);
    }
    CUDA_TRY(cudaEventRecord(s->ev_xpose[h], s->stream));
    CUDA_TRY(cudaStreamWaitEvent(s->xstream, s->ev_xpose[h], 0));
    CUDA_TRY(cudaMemcpyAsync(host + plane * k0, st, plane * nk * sizeof(double), cudaMemcpyDeviceToHost, s->xstream));
    CUDA_TRY(cudaEventRecord(s->ev_copy[h], s->xstream));
  }
  CUDA_TRY(cudaStreamSynchronize(s->xstream));                    // the host array is complete; the staging halves are free again
  CUDA_TRY(cudaGetLastError());
  return GX_OK;
}

// ---- halo exchange between blocks (replaces the six mpi_sendrecv of each boundary routine) ----
static Box face_box(const Grid& g, int dir, int side, int nl, bool ghost) {
  // ghost=false: the nl physical layers adjacent to the face; ghost=true: the nl ghost layers beyond it
  const int n[3] = {g.nx, g.ny, g.nz};
  Box b;
  for (int d = 0; d < 3; ++d) { b.lo[d] = 1 - nl; b.hi[d] = n[d] + nl; }
  if (!ghost) { b.lo[dir] = side == 0 ? 1 : n[dir] - nl + 1; b.hi[dir] = side == 0 ? nl : n[dir]; }
  else { b.lo[dir] = side == 0 ? 1 - nl : n[dir] + 1; b.hi[dir] = side == 0 ? 0 : n[dir] + nl; }
  return b;
}
static size_t box_cells(const Box& b) { return (size_t)(b.hi[0] - b.lo[0] + 1) * (b.hi[1] - b.lo[1] + 1) * (b.hi[2] - b.lo[2] + 1); }
static void launch_pack(gx_solver* s, int nvar, double* A, double* buf, const Box& b, int unpack) {
  LaunchScope ls(s, gx::KC_BC);
  dim3 grid((b.hi[0] - b.lo[0] + 1 + 63) / 64, b.hi[1] - b.lo[1] + 1, b.hi[2] - b.lo[2] + 1);
  k_pack<<<grid, 64, 0, s->stream>>>(s->A.g, nvar, A, buf, b, unpack);
}

// ---- peer-memory halo push (z slabs): replaces pack -> ncclSend/ncclRecv -> unpack ----
// In the SoA layout the nl boundary planes of one variable are one contiguous run and so are the neighbour's
// ghost planes, so a z face is ONE strided device-to-device copy (rows = variables, pitch = variable stride)
// straight into the neighbour's array, mapped here through CUDA IPC: copy engines over NVLink, no kernel, no
// staging buffer.  Ordering uses stream memory operations on 32-bit words in the RECEIVER's memory:
//   READY_FROM_x  : written by neighbour x when its stream reaches the exchange ("my ghost planes may be written")
//   ARRIVE_FROM_x : written by neighbour x after its copy into my ghost planes has completed
// Every rank first posts READY to both neighbours, then waits for each neighbour's READY before pushing, so the
// protocol cannot deadlock and a fast rank can never overwrite ghosts its neighbour still reads or re-uploads.
enum { FL_READY_FROM_LO = 0, FL_READY_FROM_HI = 16, FL_ARRIVE_FROM_LO = 32, FL_ARRIVE_FROM_HI = 48, FL_WORDS = 64 };
typedef int (*StreamOp32)(cudaStream_t, unsigned long long, unsigned int, unsigned int);
static StreamOp32 g_write32 = nullptr, g_wait32 = nullptr;
static bool load_stream_memops() {
  if (g_write32 && g_wait32) return true;
  cudaDriverEntryPointQueryResult qr;
  void *w = nullptr, *t = nullptr;
  if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &w, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) { cudaGetLastError(); return false; }
  if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &t, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) { cudaGetLastError(); return false; }
  g_write32 = (StreamOp32)w; g_wait32 = (StreamOp32)t;
  return true;
}
static double* peer_array(const gx_solver* s, int side, const double* A) {
  // `A` may point at one variable inside an array (thermal_bounds exchanges u(5) alone)
  const long long nu = s->A.g.vs * s->A.g.neq, ne = s->A.g.vs * 3;
  if (s->peer[side].U && A >= s->U && A < s->U + nu) return s->peer[side].U + (A - s->U);
  if (s->peer[side].UP && A >= s->UP && A < s->UP + nu) return s->peer[side].UP + (A - s->UP);
  if (s->peer[side].E && s->E && A >= s->E && A < s->E + ne) return s->peer[side].E + (A - s->E);
  return nullptr;
}
static int exchange_z_p2p(gx_solver* s, double* A, int nvar, int nl, cudaStream_t st) {
  const Grid& g = s->A.g;
  const int lo = s->nbr[2][0], hi = s->nbr[2][1];
  const long long plane = (long long)g.px * g.py;
  const size_t pitch = (size_t)g.vs * sizeof(double), width = (size_t)nl * plane * sizeof(double);
  const unsigned seq = ++s->xseq;
  const unsigned GEQ = 0;                               // CU_STREAM_WAIT_VALUE_GEQ (the words only ever grow)
  auto W = [&](uint32_t* base, int word) { return g_write32(st, (unsigned long long)(uintptr_t)(base + word), seq, 0); };
  auto T = [&](int word) { return g_wait32(st, (unsigned long long)(uintptr_t)(s->flags + word), seq, GEQ); };
  int e = 0;
  if (lo >= 0) e |= W(s->peer[0].flags, FL_READY_FROM_HI);          // I am lo's high neighbour
  if (hi >= 0) e |= W(s->peer[1].flags, FL_READY_FROM_LO);          // I am hi's low neighbour
  if (hi >= 0) {                                                    // my top planes nz-nl+1..nz -> hi's ghost planes 1-nl..0
    e |= T(FL_READY_FROM_HI);
    CUDA_TRY(cudaMemcpy2DAsync(peer_array(s, 1, A) + (long long)(2 - nl) * plane, pitch, A + (long long)(g.nz - nl + 2) * plane, pitch,
                               width, (size_t)nvar, cudaMemcpyDeviceToDevice, st));
    e |= W(s->peer[1].flags, FL_ARRIVE_FROM_LO);
  }
  if (lo >= 0) {                                                    // my bottom planes 1..nl -> lo's ghost planes nz+1..nz+nl
    e |= T(FL_READY_FROM_LO);
    CUDA_TRY(cudaMemcpy2DAsync(peer_array(s, 0, A) + (long long)(g.nz + 2) * plane, pitch, A + (long long)2 * plane, pitch,
                               width, (size_t)nvar, cudaMemcpyDeviceToDevice, st));
    e |= W(s->peer[0].flags, FL_ARRIVE_FROM_HI);
  }
  if (lo >= 0) e |= T(FL_ARRIVE_FROM_LO);
  if (hi >= 0) e |= T(FL_ARRIVE_FROM_HI);
  if (e) return fail(GX_ECUDA, "stream memory operation failed in the peer halo push (CUresult %d)", e);
  return GX_OK;
}

static int exchange_dir(gx_solver* s, double* A, int nvar, int nl, int dir, cudaStream_t st = nullptr) {
  if (!st) st = s->stream;
  // all faces are packed before anything is received, like the reference (boundaries.f90:70-75)
  const Grid& g = s->A.g;
  const int lo = s->nbr[dir][0], hi = s->nbr[dir][1];
  if (lo < 0 && hi < 0) return GX_OK;
  if (!s->comm) return fail(GX_ECOMM, "block has neighbours but no communicator is attached (gx_comm_attach)");
  if (dir == 2 && s->p2p && peer_array(s, lo >= 0 ? 0 : 1, A)) return exchange_z_p2p(s, A, nvar, nl, st);
  const Box sb_lo = face_box(g, dir, 0, nl, false), sb_hi = face_box(g, dir, 1, nl, false);
  const size_t cnt = box_cells(sb_lo) * nvar;
  if (cnt > s->halo_doubles[dir]) return fail(GX_ESTATE, "halo buffer too small");
  double *send_lo = s->halo_send[2 * dir], *send_hi = s->halo_send[2 * dir + 1];
  double *recv_lo = s->halo_recv[2 * dir], *recv_hi = s->halo_recv[2 * dir + 1];
  if (lo >= 0) launch_pack(s, nvar, A, send_lo, sb_lo, 0);
  if (hi >= 0) launch_pack(s, nvar, A, send_hi, sb_hi, 0);
  // Message order matters when lo == hi (two blocks, periodic): NCCL pairs the sends and receives of
  // one peer in issue order, and my upward message must land in the peer's LOW ghost layers.
  NCCL_TRY(g_nccl.GroupStart());
  if (hi >= 0) NCCL_TRY(g_nccl.Send(send_hi, cnt, ncclDouble, hi, s->comm, s->stream));
  if (lo >= 0) NCCL_TRY(g_nccl.Recv(recv_lo, cnt, ncclDouble, lo, s->comm, s->stream));
  if (lo >= 0) NCCL_TRY(g_nccl.Send(send_lo, cnt, ncclDouble, lo, s->comm, s->stream));
  if (hi >= 0) NCCL_TRY(g_nccl.Recv(recv_hi, cnt, ncclDouble, hi, s->comm, s->stream));
  NCCL_TRY(g_nccl.GroupEnd());
  if (lo >= 0) launch_pack(s, nvar, A, recv_lo, face_box(g, dir, 0, nl, true), 1);
  if (hi >= 0) launch_pack(s, nvar, A, recv_hi, face_box(g, dir, 1, nl, true), 1);
  return GX_OK;
}

static void launch_bc_face(gx_solver* s, double* A, int nvar, int dir, int side, int mode, int nl, int negvar) {
  const Grid& g = s->A.g;
  const int n[3] = {g.nx, g.ny, g.nz};
  const int ta = dir == 0 ? 1 : 0, tb = dir == 2 ? 1 : 2;
  LaunchScope ls(s, gx::KC_BC);
  dim3 grid((n[ta] + 2 * nl + 63) / 64, n[tb] + 2 * nl);
  k_bc_face<<<grid, 64, 0, s->stream>>>(g, nvar, A, dir, side, mode, nl, negvar);
}

// kind 0: conserved/primitive array (closed wall flips normal momentum, boundaries.f90:146-199, 361-438)
// kind 1: electric field (closed wall flips e(1) on x walls, e(2) on y walls, nothing on z walls,
//         flux_cd_module.f90:143-192)
// `st`: only the overlapped step passes the halo stream, and only when every direction but z is wrapped in the loaders
// and z goes through the peer push (no kernels of this function then run besides the physical fills at the z ends)
// `local_only`: no exchange with other blocks (their ghost layers were filled by the last step); this is what the
// download entry points use, so that gx_get_state / gx_get_up are NOT collective calls.
static int apply_boundaries(gx_solver* s, double* A, int nvar, int nl, int kind, bool skip_wrapped = false, cudaStream_t st = nullptr,
                            bool local_only = false) {
  if (s->nb[0] * s->nb[1] * s->nb[2] == 1 && !getenv("GX_NO_BC_SHELL")) {       // no neighbours: the whole ghost shell in one launch
    ShellBc bc;
    bool any = false;
    for (int dir = 0; dir < 3; ++dir)
      for (int side = 0; side < 2; ++side) {
        int m = 0, neg = -1;
        if (s->periodic[dir]) m = (skip_wrapped && s->A.wrap[dir]) ? 0 : 1;
        else if (s->bc[dir][side] == GX_BC_OUTFLOW) m = 2;
        else if (s->bc[dir][side] == GX_BC_CLOSED) { m = 2; neg = kind == 0 ? 1 + dir : (dir == 2 ? -1 : dir); }
        bc.mode[2 * dir + side] = m; bc.neg[2 * dir + side] = neg;
        any = any || m != 0;
      }
    if (any) {
      const Grid& g = s->A.g;
      LaunchScope ls(s, gx::KC_BC);
      const dim3 grid((std::max(g.nx + 2 * nl, g.ny) + 127) / 128, 2 * nl * (g.ny + 2 * nl) + 4 * nl * g.nz);
      k_bc_shell<<<grid, 128, 0, st ? st : s->stream>>>(g, nvar, A, nl, bc);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return fail(GX_ECUDA, "boundary kernel launch: %s", cudaGetErrorString(e));
    }
    return GX_OK;
  }
  for (int dir = 0; dir < 3; ++dir) {
    if (s->nb[dir] == 1 && s->periodic[dir]) {            // neighbour is the block itself
      if (skip_wrapped && s->A.wrap[dir]) continue;       // the fused kernels wrap their reads instead
      launch_bc_face(s, A, nvar, dir, 0, 0, nl, -1);
      launch_bc_face(s, A, nvar, dir, 1, 0, nl, -1);
    } else if (!local_only) {
      int rc = exchange_dir(s, A, nvar, nl, dir, st);
      if (rc) return rc;
    }
  }
  for (int pass = 0; pass < 2; ++pass) {                   // closed first, then outflow (reference order)
    const int want = pass == 0 ? GX_BC_CLOSED : GX_BC_OUTFLOW;
    for (int dir = 0; dir < 3; ++dir)
      for (int side = 0; side < 2; ++side) {
        if (s->bc[dir][side] != want) continue;
        if (s->co[dir] != (side == 0 ? 0 : s->nb[dir] - 1)) continue;   // only blocks on the domain edge
        int negvar = -1;
        if (want == GX_BC_CLOSED) negvar = kind == 0 ? 1 + dir : (dir == 2 ? -1 : dir);
        launch_bc_face(s, A, nvar, dir, side, 1, nl, negvar);
      }
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(GX_ECUDA, "boundary kernel launch: %s", cudaGetErrorString(e));
  return GX_OK;
}

static int apply_user_bc(gx_solver* s, double* A, int order) {
  if (!s->cfg.bc_user) return GX_OK;
  const Grid& g = s->A.g;
  // the reference's impose_user_bc is also where problem modules move their state (the orbiting planet,
  // EXO/exoplanet.f90:137-144): the host hook may re-position the device functors before they are applied
  if (s->bc_hook) s->bc_hook(order, s->time, s->bc_hook_user);
  if (!s->spheres.empty()) launch_wind_spheres(s, A);
  if (s->host_bc) {   // slow path: device -> host -> callback -> device
    std::vector<double> h((size_t)g.neq * (g.nx + 4) * (g.ny + 4) * (g.nz + 4));
    int rc = download_aos(s, A, h.data(), g.neq); if (rc) return rc;
    s->host_bc(h.data(), order, s->host_bc_user);
    rc = upload_aos(s, h.data(), A, g.neq); if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(s->stream));
  }
  return GX_OK;
}

// ---- slow path of get_user_source_terms for arbitrary user code (src/sources.f90:205): primitives to the host, the
// user's callback fills s(neq, ...) in reference layout, s back to the device (into the free face-flux array), then
// up = up + dt * s over the interior (hydro_solver.f90:115-121).  Not on any timed path.
__global__ void k_add_source(Grid g, int nvar, double dt, const double* __restrict__ S, double* __restrict__ dst) {
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x) + 1, j = (int)blockIdx.y + 1, k = (int)blockIdx.z + 1;
  if (i > g.nx) return;
  const long long c = g.idx(i, j, k);
  for (int q = 0; q < nvar; ++q) dst[q * g.vs + c] = dst[q * g.vs + c] + dt * S[q * g.vs + c];
}
static int apply_host_source(gx_solver* s, double dt, double* dst) {
  if (!s->host_src) return GX_OK;
  const Grid& g = s->A.g;
  const size_t n = (size_t)g.neq * (g.nx + 4) * (g.ny + 4) * (g.nz + 4);
  std::vector<double> hw(n), hs(n, 0.0);
  int rc = download_aos(s, s->W, hw.data(), g.neq); if (rc) return rc;
  s->host_src(hw.data(), hs.data(), s->host_src_user);
  rc = upload_aos(s, hs.data(), s->F, g.neq); if (rc) return rc;
  { LaunchScope ls(s, gx::KC_UPDATE); k_add_source<<<dim3((g.nx + 127) / 128, g.ny, g.nz), 128, 0, s->stream>>>(g, g.neq, dt, s->F, dst); }
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  return GX_OK;
}

// ---------------------------------------------------------------------------
extern "C" {

const char* gx_last_error(void) { return g_err.c_str(); }

// prim2fhll* called directly (src/hll.f90:47, hllc.f90:44, hlle.f90:48, hlld.f90:48) for n state pairs
int gx_riemann_flux(const gx_config* c, int32_t n, const double* wl, const double* wr, double* ff, int32_t* err) {
  if (!c || !wl || !wr || !ff || n < 0) return fail(GX_EINVAL, "null argument");
  if (c->struct_bytes != (int)sizeof(gx_config)) return fail(GX_EINVAL, "gx_config size mismatch");
  if (c->neqdyn != 5 && c->neqdyn != 8) return fail(GX_EINVAL, "neqdyn must be 5 or 8");
  const bool mhd_solver = c->riemann_solver == GX_SOLVER_HLLE || c->riemann_solver == GX_SOLVER_HLLD;
  if (mhd_solver != (c->neqdyn == 8)) return fail(GX_EINVAL, "HLLE/HLLD need neqdyn = 8, HLL/HLLC neqdyn = 5 (SURVEY Q12)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(GX_ENODEVICE, "no CUDA device: the Riemann solvers have no CPU fallback"); }
  if (c->device >= 0) { if (c->device >= ndev) return fail(GX_ENODEVICE, "device %d of %d", c->device, ndev); cudaSetDevice(c->device); }
  if (n == 0) return GX_OK;
  gxp::Phys P{};
  P.cv = c->cv; P.gamma = c->gamma; P.Tempsc = c->Tempsc; P.inv_cv = 1.0 / c->cv; P.m4gamma = -4.0 * c->gamma; P.inv_Tempsc = 1.0 / c->Tempsc;
  P.eos = c->eq_of_state; P.neqdyn = c->neqdyn; P.npas = 0;
  const int nq = c->neqdyn;
  std::vector<double> hl((size_t)n * 8, 0.0), hr((size_t)n * 8, 0.0), hf((size_t)n * 8, 0.0);
  std::vector<int> he((size_t)n, 0);
  for (int t = 0; t < n; ++t) for (int q = 0; q < nq; ++q) { hl[(size_t)t * 8 + q] = wl[(size_t)t * nq + q]; hr[(size_t)t * 8 + q] = wr[(size_t)t * nq + q]; }
  double *dl = nullptr, *dr = nullptr, *df = nullptr; int* de = nullptr;
  const size_t b = (size_t)n * 8 * sizeof(double);
  auto cleanup = [&]() { cudaFree(dl); cudaFree(dr); cudaFree(df); cudaFree(de); };
  if (cudaMalloc((void**)&dl, b) != cudaSuccess || cudaMalloc((void**)&dr, b) != cudaSuccess || cudaMalloc((void**)&df, b) != cudaSuccess ||
      cudaMalloc((void**)&de, (size_t)n * sizeof(int)) != cudaSuccess) { cleanup(); cudaGetLastError(); return fail(GX_ENOMEM, "cudaMalloc failed"); }
  cudaMemcpy(dl, hl.data(), b, cudaMemcpyHostToDevice);
  cudaMemcpy(dr, hr.data(), b, cudaMemcpyHostToDevice);
  const gx::KernelTable* K = c->strict_fp ? gx::kernels_strict() : gx::kernels_fast();
  int rc = K->riemann_points(P, c->riemann_solver, n, dl, dr, df, de, 0);
  if (rc) { cleanup(); return fail(rc, "riemann solver %d not implemented", c->riemann_solver); }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { cleanup(); return fail(GX_ECUDA, "riemann kernel: %s", cudaGetErrorString(e)); }
  cudaMemcpy(hf.data(), df, b, cudaMemcpyDeviceToHost);
  cudaMemcpy(he.data(), de, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost);
  cleanup();
  for (int t = 0; t < n; ++t) { for (int q = 0; q < nq; ++q) ff[(size_t)t * nq + q] = hf[(size_t)t * 8 + q]; if (err) err[t] = he[t]; }
  return GX_OK;
}

const char* gx_build_info(void) {
  return "libguacho_gx sm_100a; kernels: strict(-fmad=false) + fast(-fmad=true); FP64; built " __DATE__;
}

int gx_create(const gx_config* c, gx_solver** out) {
  if (!c || !out) return fail(GX_EINVAL, "null argument");
  *out = nullptr;
  if (c->struct_bytes != (int)sizeof(gx_config)) return fail(GX_EINVAL, "gx_config size mismatch: caller %d, library %d", c->struct_bytes, (int)sizeof(gx_config));
  if (c->nghost != 2) return fail(GX_EINVAL, "nghost must be 2 (parameters.f90:193)");
  if (c->nbx < 1 || c->nby < 1 || c->nbz < 1) return fail(GX_EINVAL, "bad block grid");
  if (c->nxtot % c->nbx || c->nytot % c->nby || c->nztot % c->nbz) return fail(GX_EINVAL, "grid not divisible by block grid");
  const int nx = c->nxtot / c->nbx, ny = c->nytot / c->nby, nz = c->nztot / c->nbz;
  if (nx < 2 || ny < 2 || nz < 2) return fail(GX_EINVAL, "each block needs >= 2 cells per direction");
  if (c->cx < 0 || c->cx >= c->nbx || c->cy < 0 || c->cy >= c->nby || c->cz < 0 || c->cz >= c->nbz) return fail(GX_EINVAL, "block coords outside block grid");
  if (c->neqdyn != 5 && c->neqdyn != 8) return fail(GX_EINVAL, "neqdyn must be 5 or 8");
  if (c->neq != c->neqdyn + c->npas || c->npas < 0) return fail(GX_EINVAL, "neq != neqdyn + npas");
  if (c->pmhd) return fail(GX_EUNSUPPORTED, "passive-MHD (pmhd) is not implemented on the device path");
  const bool mhd_solver = c->riemann_solver == GX_SOLVER_HLLE || c->riemann_solver == GX_SOLVER_HLLD;
  const bool hd_solver = c->riemann_solver == GX_SOLVER_HLL || c->riemann_solver == GX_SOLVER_HLLC;
  const bool split_solver = c->riemann_solver == GX_SOLVER_HLLE_SPLIT_ALL;
  if (!mhd_solver && !hd_solver && !split_solver)
    return fail(GX_EUNSUPPORTED, "riemann_solver %d not implemented (the SPLIT_B variants and HLLD_SPLIT_ALL have no reference implementation either: "
                                 "hydro_solver.f90:159-162 calls none of them)", c->riemann_solver);
  if (split_solver && !(c->mhd && c->neqdyn == 8 && c->npas == 0 && (c->eq_of_state == GX_EOS_ADIABATIC || c->eq_of_state == GX_EOS_SINGLE_SPECIE)))
    return fail(GX_EINVAL, "HLLE_SPLIT_ALL: mhd = 1, neqdyn = 8, no passive scalars (the split flux exists under `if (mhd)` only, hydro_core.f90:404)");
  if (split_solver && c->cooling != GX_COOL_NONE) return fail(GX_EUNSUPPORTED, "HLLE_SPLIT_ALL with cooling");
  if (split_solver && c->th_cond != GX_TC_OFF) return fail(GX_EUNSUPPORTED, "HLLE_SPLIT_ALL with thermal conduction");
  if (mhd_solver && !(c->mhd && c->neqdyn == 8)) return fail(GX_EINVAL, "HLLE/HLLD need mhd=1, neqdyn=8");
  if (hd_solver && (c->mhd || c->neqdyn != 5)) return fail(GX_EINVAL, "HLL/HLLC use hydro wave speeds: run with mhd=0, neqdyn=5 (SURVEY Q12)");
  if (c->enable_flux_cd && !c->mhd) return fail(GX_EINVAL, "flux-CD without B field updates nothing (hydro_solver.f90:103-113)");
  if (c->slope_limiter < -1 || c->slope_limiter > 6) return fail(GX_EINVAL, "unknown slope limiter");
  if (c->cooling != GX_COOL_NONE && c->cooling != GX_COOL_H) return fail(GX_EUNSUPPORTED, "cooling %d stays in the host (only COOL_NONE / COOL_H are device operators)", c->cooling);
  if (c->cooling == GX_COOL_H && (c->npas < 1 || !(c->tsc > 0))) return fail(GX_EINVAL, "COOL_H needs npas >= 1 (neutral H density in u(neqdyn+1)) and tsc > 0");
  if (c->th_cond < GX_TC_OFF || c->th_cond > GX_TC_ANISOTROPIC) return fail(GX_EINVAL, "unknown th_cond %d", c->th_cond);
  if (c->th_cond == GX_TC_ANISOTROPIC && !c->mhd) return fail(GX_EINVAL, "anisotropic thermal conduction needs the B field (mhd = 1)");
  if (c->th_cond != GX_TC_OFF && !(c->tsc > 0 && c->rsc > 0 && c->rhosc > 0 && c->vsc2 > 0 && c->mu > 0 && (c->th_cond == GX_TC_ISOTROPIC || c->bsc > 0)))
    return fail(GX_EINVAL, "thermal conduction needs the cgs scalings tsc, rsc, rhosc, vsc2, mu (and bsc when anisotropic) > 0");
  if (c->eq_of_state == GX_EOS_CHEM) return fail(GX_EUNSUPPORTED, "EOS_CHEM needs the chemistry network (out of scope)");
  const int bcs[6] = {c->bc_left, c->bc_right, c->bc_bottom, c->bc_top, c->bc_out, c->bc_in};
  for (int b : bcs) if (b < GX_BC_OUTFLOW || b > GX_BC_OTHER) return fail(GX_EINVAL, "unknown boundary condition %d", b);
  if (!(c->dx > 0 && c->dy > 0 && c->dz > 0 && c->cv > 0 && c->gamma > 0)) return fail(GX_EINVAL, "dx, dy, dz, cv, gamma must be positive");

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(GX_ENODEVICE, "no CUDA device: the step has no CPU fallback"); }
  gx_solver* s = new gx_solver();
  s->cfg = *c;
  if (c->device >= 0) { if (c->device >= ndev) { delete s; return fail(GX_ENODEVICE, "device %d of %d", c->device, ndev); } s->device = c->device; cudaSetDevice(s->device); }
  else cudaGetDevice(&s->device);

  Grid& g = s->A.g;
  g.nx = nx; g.ny = ny; g.nz = nz;
  g.xo = 15;
  g.px = ((nx + 2 + g.xo + 1) + 15) / 16 * 16;
  g.py = ny + 4; g.pz = nz + 4;
  g.vs = (long long)g.px * g.py * g.pz;
  g.neq = c->neq; g.neqdyn = c->neqdyn; g.npas = c->npas;
  g.cx = c->cx; g.cy = c->cy; g.cz = c->cz;
  g.nxtot = c->nxtot; g.nytot = c->nytot; g.nztot = c->nztot;
  g.dx = c->dx; g.dy = c->dy; g.dz = c->dz;
  s->A.phys.cv = c->cv; s->A.phys.gamma = c->gamma; s->A.phys.Tempsc = c->Tempsc; s->A.phys.inv_cv = 1.0 / c->cv; s->A.phys.m4gamma = -4.0 * c->gamma; s->A.phys.inv_Tempsc = 1.0 / c->Tempsc;
  s->A.phys.eos = c->eq_of_state; s->A.phys.neqdyn = c->neqdyn; s->A.phys.npas = c->npas;
  s->A.idx3[0] = 1.0 / c->dx; s->A.idx3[1] = 1.0 / c->dy; s->A.idx3[2] = 1.0 / c->dz;
  s->A.solver = c->riemann_solver; s->A.limiter = c->slope_limiter;
  s->A.flux_cd = c->enable_flux_cd; s->A.eight_wave = c->eight_wave; s->A.user_src = c->user_source_terms;
  s->A.grav.n = 0; s->A.W0 = nullptr;
  s->A.kbeg = 1; s->A.klast = nz; s->A.kbeg2 = 1; s->A.klast2 = 0;
  s->K = c->strict_fp ? gx::kernels_strict() : gx::kernels_fast();

  s->nb[0] = c->nbx; s->nb[1] = c->nby; s->nb[2] = c->nbz;
  s->co[0] = c->cx; s->co[1] = c->cy; s->co[2] = c->cz;
  for (int d = 0; d < 3; ++d) { s->bc[d][0] = bcs[2 * d]; s->bc[d][1] = bcs[2 * d + 1]; s->periodic[d] = (bcs[2 * d] == GX_BC_PERIODIC && bcs[2 * d + 1] == GX_BC_PERIODIC); }
  auto rank_of = [&](int x, int y, int z) { return (x * c->nby + y) * c->nbz + z; };
  for (int d = 0; d < 3; ++d)
    for (int side = 0; side < 2; ++side) {
      int cc[3] = {c->cx, c->cy, c->cz};
      cc[d] += side == 0 ? -1 : 1;
      if (cc[d] < 0 || cc[d] >= s->nb[d]) { if (!s->periodic[d]) { s->nbr[d][side] = -1; continue; } cc[d] = (cc[d] + s->nb[d]) % s->nb[d]; }
      s->nbr[d][side] = (s->nb[d] == 1) ? -1 : rank_of(cc[0], cc[1], cc[2]);   // self-neighbour handled locally
    }
  s->rank = rank_of(c->cx, c->cy, c->cz);

#define ALLOC(p, n) do { cudaError_t e_ = cudaMalloc((void**)&(p), (n)); if (e_ != cudaSuccess) { std::string m = cudaGetErrorString(e_); gx_destroy(s); return fail(GX_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", (size_t)(n), m.c_str()); } cudaMemsetAsync((p), 0, (n), 0); } while (0)
  if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) { s->stream = nullptr; gx_destroy(s); cudaGetLastError(); return fail(GX_ECUDA, "cudaStreamCreate failed"); }
  const size_t var_bytes = (size_t)g.vs * sizeof(double);
  // Fused stage kernels (gx_stage.cu): (a) the dynamic variables with the adiabatic equation of state, no sources — the headline path;
  // (b) two passive scalars, any equation of state but EOS_CHEM, user sources as the point-mass gravity functor — EXO as shipped.
  // eta != 0 adds the viscous_copy pass (its stale half-step ghosts are reproduced).  The 8-wave source, other passive
  // counts and a host-callback source (gx_register_host_source switches this off) take the pass-per-routine kernels.
  const bool fuse_a = c->npas == 0 && !c->user_source_terms && c->eq_of_state == GX_EOS_ADIABATIC;
  const bool fuse_b = c->npas == 2;
  s->fused = (fuse_a || fuse_b) && !c->eight_wave && !split_solver && !getenv("GX_NO_FUSED");
  s->kz = 0;                                        // planes per CTA of the fused stage kernels: chosen by their launcher
  if (const char* e = getenv("GX_KZ")) s->kz = std::max(1, atoi(e));
  // a user boundary functor may write ghost cells, so ghosts must be real arrays then
  // (eta != 0: viscous_copy reads up's ghost cells — the half-step halo, SURVEY Q5 — so they must be real arrays as well;
  //  thermal conduction reads the ghost layer of u)
  for (int d = 0; d < 3; ++d) s->A.wrap[d] = (s->fused && s->periodic[d] && s->nb[d] == 1 && !c->bc_user && c->eta == 0.0 && c->th_cond == GX_TC_OFF && !getenv("GX_NO_WRAP")) ? 1 : 0;
  // the loaders of the fused headline kernels are ONE TMA tile load per plane (GX_TMA=0: per-thread cp.async everywhere)
  // GX_TMA = 0: cp.async everywhere; 1 (default): TMA for the second-order stage; 2: both stages
  s->A.tma = (fuse_a && s->fused) ? (getenv("GX_TMA") ? atoi(getenv("GX_TMA")) : 1) : 0;
  ALLOC(s->U, var_bytes * g.neq);
  ALLOC(s->UP, var_bytes * g.neq);
  if (!s->fused) {                                   // primitives and face fluxes only exist in HBM on the unfused path
    ALLOC(s->W, var_bytes * g.neq);
    ALLOC(s->F, var_bytes * g.neq * 3);
  } else if (c->eta != 0.0) {
    ALLOC(s->T, var_bytes * g.neq);                  // full-step state before viscous_copy (up keeps its half-step ghosts)
  }
  if (c->enable_flux_cd) ALLOC(s->E, var_bytes * 3);
  if (c->th_cond != GX_TC_OFF) ALLOC(s->PT, var_bytes * 2 + 256 * 16 * sizeof(double));    // p, T + the reduction slots of get_dt_cond
  if (split_solver) ALLOC(s->W0, var_bytes * g.neq);
  ALLOC(s->dscal, sizeof(gx_solver::DevScalars));
  if (cudaMallocHost((void**)&s->hscal, sizeof(gx_solver::DevScalars)) != cudaSuccess) { s->hscal = nullptr; gx_destroy(s); cudaGetLastError(); return fail(GX_ENOMEM, "cudaMallocHost failed (pinned scalars)"); }
  // staging: up to 64 MiB or 4 planes, whichever is larger
  const size_t plane = (size_t)g.neq * (g.nx + 4) * (g.ny + 4);
  s->stage_doubles = std::max(plane * 4, std::min(plane * (size_t)(g.nz + 4), (size_t)(64u << 20) / sizeof(double)));
  s->stage_doubles = std::min(s->stage_doubles, plane * (size_t)(g.nz + 4));
  ALLOC(s->stage, s->stage_doubles * sizeof(double));
  // halo buffers only where a real neighbour exists
  const int n3[3] = {nx, ny, nz};
  for (int d = 0; d < 3; ++d) {
    if (s->nbr[d][0] < 0 && s->nbr[d][1] < 0) continue;
    size_t cells = 2;   // nl = 2 layers, full transverse extent
    for (int t = 0; t < 3; ++t) if (t != d) cells *= (size_t)(n3[t] + 4);
    s->halo_doubles[d] = cells * g.neq;
    for (int side = 0; side < 2; ++side) { ALLOC(s->halo_send[2 * d + side], s->halo_doubles[d] * sizeof(double)); ALLOC(s->halo_recv[2 * d + side], s->halo_doubles[d] * sizeof(double)); }
  }
#undef ALLOC
  if (cudaEventCreate(&s->ev0) != cudaSuccess || cudaEventCreate(&s->ev1) != cudaSuccess) { gx_destroy(s); cudaGetLastError(); return fail(GX_ECUDA, "cudaEventCreate failed"); }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { gx_destroy(s); return fail(GX_ECUDA, "device sync after allocation: %s", cudaGetErrorString(e)); }
  *out = s;
  return GX_OK;
}

int gx_destroy(gx_solver* s) {
  if (!s) return GX_OK;
  cudaSetDevice(s->device);
  if (s->stream) cudaStreamSynchronize(s->stream);
  if (s->cstream) cudaStreamSynchronize(s->cstream);
  if (s->xstream) cudaStreamSynchronize(s->xstream);
  if (s->p2p && s->comm && g_nccl.ok) {          // nobody may unmap or free while a neighbour can still push into these arrays
    g_nccl.AllReduce(s->flags + FL_WORDS - 1, s->flags + FL_WORDS - 1, 1, ncclUint32, ncclSum, s->comm, s->stream);
    cudaStreamSynchronize(s->stream);
  }
  for (int side = 0; side < 2; ++side) {
    if (side == 1 && s->peer[1].U == s->peer[0].U) break;
    void* ptrs[] = {s->peer[side].U, s->peer[side].UP, s->peer[side].E, s->peer[side].flags};
    for (void* p : ptrs) if (p) cudaIpcCloseMemHandle(p);
  }
  if (s->flags) cudaFree(s->flags);
  if (s->comm && g_nccl.ok) g_nccl.CommDestroy(s->comm);
  double* ptrs[] = {s->U, s->UP, s->W, s->F, s->E, s->Temp, s->T, s->PT, s->W0, s->stage};
  for (double* p : ptrs) if (p) cudaFree(p);
  for (int q = 0; q < 6; ++q) { if (s->halo_send[q]) cudaFree(s->halo_send[q]); if (s->halo_recv[q]) cudaFree(s->halo_recv[q]); }
  if (s->dscal) cudaFree(s->dscal);
  if (s->hscal) cudaFreeHost(s->hscal);
  for (auto e : s->evpool) cudaEventDestroy(e);
  if (s->ev0) cudaEventDestroy(s->ev0);
  if (s->ev1) cudaEventDestroy(s->ev1);
  if (s->ev_bnd) cudaEventDestroy(s->ev_bnd);
  if (s->ev_comm) cudaEventDestroy(s->ev_comm);
  if (s->cstream) cudaStreamDestroy(s->cstream);
  for (int h = 0; h < 2; ++h) { if (s->ev_copy[h]) cudaEventDestroy(s->ev_copy[h]); if (s->ev_xpose[h]) cudaEventDestroy(s->ev_xpose[h]); }
  if (s->xstream) cudaStreamDestroy(s->xstream);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
  return GX_OK;
}

static int reset_scalars(gx_solver* s) {          // CFL minimum := +inf, solver error flag := 0 (new state)
  gx_solver::DevScalars init; init.dtmin_bits = 0x7FF0000000000000ull; init.err = 0; init.pad = 0; init.tc_bits = 0x7FF0000000000000ull;
  *s->hscal = init;
  CUDA_TRY(cudaMemcpyAsync(s->dscal, s->hscal, sizeof init, cudaMemcpyHostToDevice, s->stream));
  return GX_OK;
}
static int reset_dtmin(gx_solver* s) {            // CFL minimum := +inf; the error flag stays sticky
  CUDA_TRY(cudaMemsetAsync(&s->dscal->dtmin_bits, 0x7f, sizeof(unsigned long long), s->stream));   // 0x7f7f.. ~ 1.4e306
  return GX_OK;
}
static int ensure_array(gx_solver* s, double** p, size_t nvar) {
  if (*p) return GX_OK;
  const size_t bytes = (size_t)s->A.g.vs * sizeof(double) * nvar;
  cudaError_t e = cudaMalloc((void**)p, bytes);
  if (e != cudaSuccess) return fail(GX_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
  CUDA_TRY(cudaMemsetAsync(*p, 0, bytes, s->stream));
  return GX_OK;
}

// ---------------------------------------------------------------------------
namespace gxtc {
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

// thermal_conduction (src/thermal_cond.f90:690-768), called at the end of tstep on u with its ghost layer filled
// (hydro_solver.f90:216-227).  One host round trip per call: the conduction time scale decides the number of substeps,
// exactly where the reference does its mpi_allreduce (:104).
static int thermal_bounds(gx_solver* s) {
  // :496-616 (MPI branch): one layer of u(5) between blocks, then zero-gradient copies on every face of the DOMAIN whatever
  // its boundary type — so a periodic direction owned by one block needs no wrap copy at all (it would be overwritten).
  const Grid& g = s->A.g;
  double* A = s->U + 4 * g.vs;
  int edge = 0;
  for (int dir = 0; dir < 3; ++dir) {
    for (int side = 0; side < 2; ++side) if (s->co[dir] == (side == 0 ? 0 : s->nb[dir] - 1)) edge |= 1 << (2 * dir + side);
    if (s->nb[dir] == 1) continue;
    int rc = exchange_dir(s, A, 1, 1, dir); if (rc) return rc;
  }
  LaunchScope ls(s, gx::KC_BC);
  s->K->tc_fill(s->A, A, edge, s->stream);
  return GX_OK;
}
static int thermal_conduction(gx_solver* s, double dt_cfl) {
  const Grid& g = s->A.g;
  const gx_config& c = s->cfg;
  gx::TcPar t;
  t.mode = c.th_cond; t.sat = c.tc_saturation; t.mhd = c.mhd;
  t.dxr = c.dx * c.rsc; t.dyr = c.dy * c.rsc; t.dzr = c.dz * c.rsc;
  t.idxr = 1.0 / t.dxr; t.idyr = 1.0 / t.dyr; t.idzr = 1.0 / t.dzr;
  t.dx = c.dx; t.dy = c.dy; t.dz = c.dz; t.idx = 1.0 / c.dx; t.idy = 1.0 / c.dy; t.idz = 1.0 / c.dz;
  t.vsc = sqrt(c.vsc2); t.sqrt_vsc2 = sqrt(c.vsc2);
  t.Psc = c.rhosc * c.vsc2; t.rhosc = c.rhosc; t.bsc2 = c.bsc * c.bsc;
  const bool alone = s->nb[0] * s->nb[1] * s->nb[2] == 1;      // the block owns every face of the domain: ghost copies written by the update itself
  auto prim = [&](int want_dt) { LaunchScope ls(s, gx::KC_TCOND); s->K->tc_prim(s->A, c.mhd, s->U, s->PT, &s->dscal->tc_bits, want_dt, s->stream); };
  auto update = [&](double dts) { LaunchScope ls(s, gx::KC_TCOND); s->K->tc_update(s->A, t, alone ? 1 : 0, s->PT, s->U, dts, s->stream); };
  // get_dt_cond (:78-110)
  prim(1);
  if (s->comm && s->nranks > 1) NCCL_TRY(g_nccl.AllReduce(&s->dscal->tc_bits, &s->dscal->tc_bits, 1, ncclUint64, ncclMin, s->comm, s->stream));
  CUDA_TRY(cudaMemcpyAsync(&s->hscal->tc_bits, &s->dscal->tc_bits, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->stream));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  double dtp; memcpy(&dtp, &s->hscal->tc_bits, sizeof dtp);
  double ddx = std::min(c.dx, c.dy);
  ddx = std::min(ddx, c.dz);
  // the scaling is a chain of products with positive factors, monotonic under rounding: min first or last is the same number
  const double dt_cond = 0.25 * 0.5 * ((ddx * c.rsc) * (ddx * c.rsc)) * c.cv * 8.3145e7 * dtp * c.rhosc / c.mu;          // :101
  const double dt_hydro = dt_cfl * c.tsc;
  bool SuperStep = true;
  int Nsteps; double fstep;
  if (dt_cond < dt_hydro) gxtc::ST_steps(dt_hydro / dt_cond, Nsteps, fstep);
  else { SuperStep = false; fstep = dt_hydro / dt_cond; Nsteps = 1; }
  s->tc_dt_cond = dt_cond; s->tc_nsteps = Nsteps;
  // isotropic conduction on a block without neighbours: one marching kernel per substep (k_tc_march), u(5) alternating between
  // the energy array of u and the first scratch variable; everything else: k_tc_update / thermal_bounds / k_tc_prim
  const bool march = alone && c.th_cond == GX_TC_ISOTROPIC && !getenv("GX_NO_TC_MARCH");
  double* const e5u = s->U + 4 * g.vs;
  double* e5in = e5u;
  for (int n = 1; n <= Nsteps; ++n) {
    double dts;
    if (SuperStep) dts = dt_cond * fstep * gxtc::substep(n, Nsteps, 0.01) / t.Psc / c.rsc;                                 // :732
    else dts = dt_hydro / (double)Nsteps / t.Psc / c.rsc;
    if (march) {
      double* e5out = e5in == e5u ? s->PT : e5u;
      { LaunchScope ls(s, gx::KC_TCOND); s->K->tc_march(s->A, t, c.mhd, s->U, e5in, e5out, dts, s->stream); }
      e5in = e5out;
      if (n == Nsteps && e5in != e5u)                  // odd number of substeps: the result sits in the scratch variable (ghost layer included)
        CUDA_TRY(cudaMemcpyAsync(e5u, e5in, (size_t)g.vs * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
      continue;
    }
    update(dts);
    if (!alone) { int rc = thermal_bounds(s); if (rc) return rc; }
    if (n < Nsteps) prim(0);                         // calcprim (:764); after the last substep the caller's calcprim pass does it
  }
  CUDA_TRY(cudaGetLastError());
  return GX_OK;
}

// boundaryI + calcprim(u, primit) (+ CFL candidates for the next get_timestep)
static int finish_u(gx_solver* s) {
  int rc = apply_boundaries(s, s->U, s->A.g.neq, 1, 0); if (rc) return rc;
  rc = apply_user_bc(s, s->U, 1); if (rc) return rc;
  rc = reset_dtmin(s); if (rc) return rc;
  { LaunchScope ls(s, gx::KC_PRIM); s->K->calcprim(s->A, s->U, s->W, nullptr, &s->dscal->dtmin_bits, 1, s->stream); }
  CUDA_TRY(cudaGetLastError());
  return GX_OK;
}

int gx_set_background(gx_solver* s, const double* primit0) {
  if (!s || !primit0) return fail(GX_EINVAL, "null argument");
  if (s->cfg.riemann_solver != GX_SOLVER_HLLE_SPLIT_ALL) return fail(GX_EINVAL, "gx_set_background: riemann_solver is not HLLE_SPLIT_ALL");
  cudaSetDevice(s->device);
  int rc = upload_aos(s, primit0, s->W0, s->A.g.neq); if (rc) return rc;
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  collect_timed(s);
  s->A.W0 = s->W0;
  return GX_OK;
}

int gx_set_state(gx_solver* s, const double* u) {
  if (!s || !u) return fail(GX_EINVAL, "null argument");
  if (s->cfg.riemann_solver == GX_SOLVER_HLLE_SPLIT_ALL && !s->A.W0)
    return fail(GX_ESTATE, "HLLE_SPLIT_ALL: call gx_set_background (primit0) before gx_set_state — u holds fluctuations about it");
  cudaSetDevice(s->device);
  int rc = upload_aos(s, u, s->U, s->A.g.neq); if (rc) return rc;
  rc = reset_scalars(s); if (rc) return rc;
  rc = finish_u(s); if (rc) return rc;
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  collect_timed(s);
  s->ghosts_stale = false;
  s->have_state = true;
  return GX_OK;
}

int gx_set_time(gx_solver* s, double time) { if (!s) return fail(GX_EINVAL, "null"); s->time = time; return GX_OK; }

int gx_get_timestep(gx_solver* s, int32_t current_iter, int32_t n_iter, double current_time, double tprint, double* dt, int32_t* dump_flag) {
  if (!s || !dt) return fail(GX_EINVAL, "null argument");
  if (!s->have_state) return fail(GX_ESTATE, "gx_set_state has not been called");
  cudaSetDevice(s->device);
  if (s->comm && s->nranks > 1) {   // mpi_allreduce(MIN) (hydro_core.f90:685); min of positive doubles == min of bit patterns
    NCCL_TRY(g_nccl.AllReduce(&s->dscal->dtmin_bits, &s->dscal->dtmin_bits, 1, ncclUint64, ncclMin, s->comm, s->stream));
  }
  CUDA_TRY(cudaMemcpyAsync(s->hscal, s->dscal, sizeof(gx_solver::DevScalars), cudaMemcpyDeviceToHost, s->stream));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  if (s->hscal->err) return fail(GX_ENUMERIC, "Riemann solver fell through every branch (NaN state): the reference prints 'Error in HLLD routine' and stops");
  double dtp; memcpy(&dtp, &s->hscal->dtmin_bits, sizeof dtp);
  dtp = std::min(dtp, 1.e30);                                            // hydro_core.f90:642
  // the ramp multiplies by exact powers of two, so applying it after the global min is
  // bit-identical to the reference's per-rank application before mpi_allreduce (:677-685)
  if (current_iter <= n_iter) dtp = s->cfg.cfl * pow(2., -(double)(n_iter + 1 - current_iter)) * dtp;
  else dtp = s->cfg.cfl * dtp;
  if ((current_time + dtp) >= tprint) { dtp = tprint - current_time; if (dump_flag) *dump_flag = 1; }
  *dt = dtp;
  return GX_OK;
}

// tstep with the fused stage kernels: per stage one k_stage launch (prim + 3 sweeps + E + update)
// and, with flux-CD, the E ghost layer + k_bupdate.  Same sequence as hydro_solver.f90:134-229.
static int tstep_enqueue_fused(gx_solver* s, double dt_cfl) {
  const gx::KernelTable* K = s->K;
  const StepArgs& A = s->A;
  const int neq = A.g.neq;
  const double dtm = dt_cfl / 2.;
  const bool cfl_in_step = !s->cfg.bc_user;          // a user BC may overwrite physical cells after the update
  int rc;
  { LaunchScope ls(s, gx::KC_STAGE1); rc = K->stage(A, 1, dtm, s->U, s->U, s->UP, s->E, s->kz, nullptr, 0, &s->dscal->err, s->stream); } if (rc) return fail(rc, "stage-1 launch");
  if (A.flux_cd) {
    rc = apply_boundaries(s, s->E, 3, 1, 1, true); if (rc) return rc;
    { LaunchScope ls(s, gx::KC_BUPDATE); K->bupdate(A, dtm, s->U, s->E, s->UP, nullptr, 0, s->stream); }
  }
  rc = apply_boundaries(s, s->UP, neq, 2, 0, true); if (rc) return rc;          // boundaryII :169
  rc = apply_user_bc(s, s->UP, 2); if (rc) return rc;
  rc = reset_dtmin(s); if (rc) return rc;
  // eta == 0: viscous_copy is u(interior) = up(interior), so the full step goes straight into u.  eta != 0: it goes into T and
  // viscous_copy (hydro_solver.f90:54-63) reads T inside the block and up — the half-step halo of boundaryII — in the ghost cells
  const bool visc = s->cfg.eta != 0.0;
  const bool cfl2 = cfl_in_step && !visc && s->cfg.cooling == GX_COOL_NONE && s->cfg.th_cond == GX_TC_OFF;   // nothing may touch u after the stage for its CFL to stand
  double* full = visc ? s->T : s->U;
  { LaunchScope ls(s, gx::KC_STAGE2); rc = K->stage(A, 2, dt_cfl, s->UP, s->U, full, s->E, s->kz, &s->dscal->dtmin_bits, cfl2 && !A.flux_cd, &s->dscal->err, s->stream); } if (rc) return fail(rc, "stage-2 launch");
  if (A.flux_cd) {
    rc = apply_boundaries(s, s->E, 3, 1, 1, true); if (rc) return rc;
    { LaunchScope ls(s, gx::KC_BUPDATE); K->bupdate(A, dt_cfl, s->U, s->E, full, &s->dscal->dtmin_bits, cfl2, s->stream); }
  }
  if (visc) { LaunchScope ls(s, gx::KC_VISC); K->viscous2(A, s->cfg.eta, s->T, s->UP, s->U, s->stream); }               // :188
  if (s->cfg.cooling == GX_COOL_H) launch_coolingh(s, dt_cfl);                  // coolingh :202-204
  rc = apply_boundaries(s, s->U, neq, 1, 0, true); if (rc) return rc;           // boundaryI :216
  rc = apply_user_bc(s, s->U, 1); if (rc) return rc;
  if (s->cfg.th_cond != GX_TC_OFF) { rc = thermal_conduction(s, dt_cfl); if (rc) return rc; }      // :227
  if (!cfl2) {
    LaunchScope ls(s, gx::KC_PRIM);
    K->calcprim(A, s->U, nullptr, nullptr, &s->dscal->dtmin_bits, 1, s->stream);
  }
  s->ghosts_stale = s->A.wrap[0] || s->A.wrap[1] || s->A.wrap[2];
  CUDA_TRY(cudaGetLastError());
  return GX_OK;
}

// tstep, fused kernels, z slabs with the peer-memory push: every producer of exchanged data (the stage kernels for
// E, k_bupdate for the B part of up / u) is launched boundary planes first; the push of those planes then runs on
// the halo stream — copy engines and stream memory operations only, no SM — while the interior launch runs on the
// solver's stream ("overlapped with interior-cell updates").  Exchange order and contents are the
// reference's (E, up 2 layers, E, u 1 layer); results are bitwise those of the serialized path.
static int tstep_enqueue_fused_overlap(gx_solver* s, double dt_cfl) {
  const gx::KernelTable* K = s->K;
  const int neq = s->A.g.neq, nz = s->A.g.nz;
  const double dtm = dt_cfl / 2.;
  const int kb = 4;                                   // boundary thickness of the stage launches (>= 2: up sends 2 layers)
  // the two boundary slabs of a kernel go out as ONE launch (second plane range of StepArgs): one tail instead of two
  StepArgs bnd = s->A, mid = s->A;
  bnd.kbeg = 1; bnd.klast = kb; bnd.kbeg2 = nz - kb + 1; bnd.klast2 = nz; mid.kbeg = kb + 1; mid.klast = nz - kb;
  StepArgs bbnd = s->A, bmid = s->A;                  // B update: exactly the layers that travel
  bbnd.kbeg = 1; bbnd.klast = 2; bbnd.kbeg2 = nz - 1; bbnd.klast2 = nz; bmid.kbeg = 3; bmid.klast = nz - 2;
  int rc;
  auto fork = [&]() { cudaEventRecord(s->ev_bnd, s->stream); cudaStreamWaitEvent(s->cstream, s->ev_bnd, 0); };
  auto join = [&]() { cudaEventRecord(s->ev_comm, s->cstream); cudaStreamWaitEvent(s->stream, s->ev_comm, 0); };
  auto stage = [&](int cls, int order, double dt, const double* S, const double* Ub, double* dst) -> int {
    {
      LaunchScope ls(s, cls);
      int r = K->stage(bnd, order, dt, S, Ub, dst, s->E, s->kz, nullptr, 0, &s->dscal->err, s->stream); if (r) return r;
    }
    fork();
    int r = apply_boundaries(s, s->E, 3, 1, 1, true, s->cstream); if (r) return r;      // boundaryI_ef
    { LaunchScope ls(s, cls); r = K->stage(mid, order, dt, S, Ub, dst, s->E, s->kz, nullptr, 0, &s->dscal->err, s->stream); if (r) return r; }
    join();
    return GX_OK;
  };
  auto bupdate = [&](double dt, const double* Ub, double* dst, int nl, unsigned long long* dtmin, int want_cfl) -> int {
    { LaunchScope ls(s, gx::KC_BUPDATE); K->bupdate(bbnd, dt, Ub, s->E, dst, dtmin, want_cfl, s->stream); }
    fork();
    int r = apply_boundaries(s, dst, neq, nl, 0, true, s->cstream); if (r) return r;    // boundaryII (nl = 2) / boundaryI (nl = 1)
    { LaunchScope ls(s, gx::KC_BUPDATE); K->bupdate(bmid, dt, Ub, s->E, dst, dtmin, want_cfl, s->stream); }
    join();
    return GX_OK;
  };
  rc = stage(gx::KC_STAGE1, 1, dtm, s->U, s->U, s->UP); if (rc) return fail(rc, "stage-1 launch");
  rc = bupdate(dtm, s->U, s->UP, 2, nullptr, 0); if (rc) return rc;
  rc = reset_dtmin(s); if (rc) return rc;
  rc = stage(gx::KC_STAGE2, 2, dt_cfl, s->UP, s->U, s->U); if (rc) return fail(rc, "stage-2 launch");
  rc = bupdate(dt_cfl, s->U, s->U, 1, &s->dscal->dtmin_bits, 1); if (rc) return rc;
  s->ghosts_stale = true;
  CUDA_TRY(cudaGetLastError());
  return GX_OK;
}

static int tstep_enqueue(gx_solver* s, double dt_cfl) {
  if (s->fused && s->overlap && s->cfg.eta == 0.0 && s->cfg.cooling == GX_COOL_NONE && s->cfg.th_cond == GX_TC_OFF && !s->cfg.bc_user) return tstep_enqueue_fused_overlap(s, dt_cfl);
  if (s->fused) return tstep_enqueue_fused(s, dt_cfl);
  const gx::KernelTable* K = s->K;
  const StepArgs& A = s->A;
  const int neq = A.g.neq;
  const double dtm = dt_cfl / 2.;                                        // hydro_solver.f90:152
  int rc;
  { LaunchScope ls(s, gx::KC_FLUX, 3); rc = K->fluxes(A, 1, s->W, s->F, &s->dscal->err, s->stream); } if (rc) return fail(rc, "flux launch");
  if (A.flux_cd) {
    { LaunchScope ls(s, gx::KC_EFIELD); K->efield(A, s->F, s->E, s->stream); }
    rc = apply_boundaries(s, s->E, 3, 1, 1); if (rc) return rc;
  }
  { LaunchScope ls(s, gx::KC_UPDATE); K->update(A, dtm, s->U, s->F, s->E, s->W, s->UP, s->stream); }      // step(dtm) :165
  rc = apply_host_source(s, dtm, s->UP); if (rc) return rc;                // get_user_source_terms, slow path (sources.f90:205)
  rc = apply_boundaries(s, s->UP, neq, 2, 0); if (rc) return rc;          // boundaryII :169
  rc = apply_user_bc(s, s->UP, 2); if (rc) return rc;
  { LaunchScope ls(s, gx::KC_PRIM); K->calcprim(A, s->UP, s->W, nullptr, nullptr, 0, s->stream); }        // :170
  { LaunchScope ls(s, gx::KC_FLUX, 3); rc = K->fluxes(A, 2, s->W, s->F, &s->dscal->err, s->stream); } if (rc) return fail(rc, "flux launch");
  if (A.flux_cd) {
    { LaunchScope ls(s, gx::KC_EFIELD); K->efield(A, s->F, s->E, s->stream); }
    rc = apply_boundaries(s, s->E, 3, 1, 1); if (rc) return rc;
  }
  if (s->cfg.eta == 0.0) {
    // viscous_copy with eta = 0 is u(interior) = up(interior): write the full step straight into u
    { LaunchScope ls(s, gx::KC_UPDATE); K->update(A, dt_cfl, s->U, s->F, s->E, s->W, s->U, s->stream); }
    rc = apply_host_source(s, dt_cfl, s->U); if (rc) return rc;
  } else {
    { LaunchScope ls(s, gx::KC_UPDATE); K->update(A, dt_cfl, s->U, s->F, s->E, s->W, s->UP, s->stream); } // step(dt) :184
    rc = apply_host_source(s, dt_cfl, s->UP); if (rc) return rc;
    { LaunchScope ls(s, gx::KC_VISC); K->viscous(A, s->cfg.eta, s->UP, s->U, s->stream); }                // :188 (stale half-step ghosts of up)
  }
  if (s->cfg.cooling == GX_COOL_H) launch_coolingh(s, dt_cfl);            // coolingh :202-204
  rc = finish_u(s); if (rc) return rc;                                    // boundaryI :216, calcprim :218-224
  if (s->cfg.th_cond != GX_TC_OFF) {                                      // thermal_conduction :227 (ends with calcprim(u, primit))
    rc = thermal_conduction(s, dt_cfl); if (rc) return rc;
    rc = reset_dtmin(s); if (rc) return rc;
    { LaunchScope ls(s, gx::KC_PRIM); K->calcprim(A, s->U, s->W, nullptr, &s->dscal->dtmin_bits, 1, s->stream); }
  }
  CUDA_TRY(cudaGetLastError());
  return GX_OK;
}

static int check_sources(gx_solver* s) {
  if (s->cfg.user_source_terms && s->A.grav.n == 0 && !s->host_src)
    return fail(GX_ESTATE, "user_source_terms = 1 but no source is attached: call gx_set_gravity_points (device functor) or "
                           "gx_register_host_source (host slow path) before stepping; the reference calls get_user_source_terms (sources.f90:205)");
  return GX_OK;
}

int gx_tstep(gx_solver* s, double dt_cfl) {
  if (!s) return fail(GX_EINVAL, "null argument");
  if (!s->have_state) return fail(GX_ESTATE, "gx_set_state has not been called");
  { int rcs = check_sources(s); if (rcs) return rcs; }
  cudaSetDevice(s->device);
  CUDA_TRY(cudaEventRecord(s->ev0, s->stream));
  int rc = tstep_enqueue(s, dt_cfl); if (rc) return rc;
  CUDA_TRY(cudaEventRecord(s->ev1, s->stream));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  float ms = 0; cudaEventElapsedTime(&ms, s->ev0, s->ev1); s->last_ms = ms;
  collect_timed(s);
  return GX_OK;
}

int gx_run(gx_solver* s, int32_t n_steps, int32_t n_iter_ramp, double* time, int32_t* iter, double* last_dt) {
  if (!s || !time || !iter) return fail(GX_EINVAL, "null argument");
  if (!s->have_state) return fail(GX_ESTATE, "gx_set_state has not been called");
  { int rcs = check_sources(s); if (rcs) return rcs; }
  cudaSetDevice(s->device);
  CUDA_TRY(cudaEventRecord(s->ev0, s->stream));
  for (int n = 0; n < n_steps; ++n) {
    double dt; int32_t dump = 0;
    int rc = gx_get_timestep(s, *iter, n_iter_ramp, *time, 1.e300, &dt, &dump); if (rc) return rc;
    s->time = *time;
    rc = tstep_enqueue(s, dt); if (rc) return rc;
    *time += dt; *iter += 1;
    if (last_dt) *last_dt = dt;
  }
  CUDA_TRY(cudaEventRecord(s->ev1, s->stream));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  float ms = 0; cudaEventElapsedTime(&ms, s->ev0, s->ev1); s->last_ms = ms;
  collect_timed(s);
  return GX_OK;
}

int gx_get_state(gx_solver* s, double* u, double* primit, double* temp) {
  if (!s) return fail(GX_EINVAL, "null argument");
  if (!s->have_state) return fail(GX_ESTATE, "gx_set_state has not been called");
  cudaSetDevice(s->device);
  const Grid& g = s->A.g;
  int rc;
  if (s->ghosts_stale) {                              // self-periodic boundaryI / boundaryII copies the fused step did not need
    rc = apply_boundaries(s, s->U, g.neq, 1, 0, false, nullptr, true); if (rc) return rc;      // (ghost layers owned by OTHER blocks were
    rc = apply_boundaries(s, s->UP, g.neq, 2, 0, false, nullptr, true); if (rc) return rc;     //  exchanged by the step: no collective here)
    s->ghosts_stale = false;
  }
  if (u) { rc = download_aos(s, s->U, u, g.neq); if (rc) return rc; }
  if (primit || temp) {                              // calcprim(u, primit) over the whole array, on demand
    rc = ensure_array(s, &s->W, g.neq); if (rc) return rc;
    if (temp) { rc = ensure_array(s, &s->Temp, 1); if (rc) return rc; }
    { LaunchScope ls(s, gx::KC_PRIM); s->K->calcprim(s->A, s->U, s->W, temp ? s->Temp : nullptr, nullptr, 0, s->stream); }
    if (primit) { rc = download_aos(s, s->W, primit, g.neq); if (rc) return rc; }
    if (temp) { rc = download_aos(s, s->Temp, temp, 1); if (rc) return rc; }
  }
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  collect_timed(s);
  return GX_OK;
}

int gx_get_up(gx_solver* s, double* up) {
  if (!s || !up) return fail(GX_EINVAL, "null argument");
  cudaSetDevice(s->device);
  int rc;
  if (s->ghosts_stale) {
    rc = apply_boundaries(s, s->U, s->A.g.neq, 1, 0, false, nullptr, true); if (rc) return rc;
    rc = apply_boundaries(s, s->UP, s->A.g.neq, 2, 0, false, nullptr, true); if (rc) return rc;
    s->ghosts_stale = false;
  }
  rc = download_aos(s, s->UP, up, s->A.g.neq); if (rc) return rc;
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  return GX_OK;
}

int gx_register_host_source(gx_solver* s, gx_host_source_fn cb, void* user) {
  if (!s) return fail(GX_EINVAL, "null argument");
  if (cb && !s->cfg.user_source_terms) return fail(GX_EINVAL, "gx_register_host_source needs user_source_terms = 1");
  if (cb && s->fused) {          // a host callback needs the primitives in HBM: leave the fused kernels for the pass-per-routine path
    if (s->have_state) return fail(GX_ESTATE, "gx_register_host_source must be called before gx_set_state");
    cudaSetDevice(s->device);
    s->fused = false;
    for (int d = 0; d < 3; ++d) s->A.wrap[d] = 0;
    int rc = ensure_array(s, &s->W, s->A.g.neq); if (rc) return rc;
    rc = ensure_array(s, &s->F, (size_t)s->A.g.neq * 3); if (rc) return rc;
  }
  s->host_src = cb; s->host_src_user = user;
  return GX_OK;
}

int gx_set_gravity_points(gx_solver* s, int32_t n, const double* gm, const double* pos) {
  if (!s || n < 0 || n > 4 || (n > 0 && (!gm || !pos))) return fail(GX_EINVAL, "0 <= n <= 4 point masses");
  s->A.grav.n = n;
  for (int l = 0; l < n; ++l) { s->A.grav.gm[l] = gm[l]; s->A.grav.x[l] = pos[3 * l]; s->A.grav.y[l] = pos[3 * l + 1]; s->A.grav.z[l] = pos[3 * l + 2]; }
  return GX_OK;
}

int gx_set_wind_spheres(gx_solver* s, int32_t n, const gx_wind_sphere* sph) {
  if (!s || n < 0 || n > GX_MAX_SPHERES || (n > 0 && !sph)) return fail(GX_EINVAL, "0 <= n <= %d wind spheres", GX_MAX_SPHERES);
  s->spheres.assign(sph, sph + n);          // takes effect at the next boundary call (kernel argument, no device copy)
  return GX_OK;
}

int gx_register_bc_hook(gx_solver* s, gx_bc_hook_fn cb, void* user) {
  if (!s) return fail(GX_EINVAL, "null argument");
  s->bc_hook = cb; s->bc_hook_user = user;
  return GX_OK;
}

int gx_register_host_bc(gx_solver* s, gx_host_bc_fn cb, void* user) {
  if (!s) return fail(GX_EINVAL, "null argument");
  s->host_bc = cb; s->host_bc_user = user;
  return GX_OK;
}

int gx_comm_unique_id(void* id_out, int32_t nbytes) {
  if (!id_out || nbytes < (int)sizeof(ncclUniqueId)) return fail(GX_EINVAL, "id buffer must hold %d bytes", (int)sizeof(ncclUniqueId));
  int rc = nccl_load(); if (rc) return rc;
  ncclUniqueId id;
  NCCL_TRY(g_nccl.GetUniqueId(&id));
  memcpy(id_out, &id, sizeof id);
  return GX_OK;
}

int gx_comm_attach(gx_solver* s, const void* idp, int32_t nbytes, int32_t rank, int32_t nranks) {
  if (!s || !idp || nbytes < (int)sizeof(ncclUniqueId)) return fail(GX_EINVAL, "bad arguments");
  if (nranks != s->nb[0] * s->nb[1] * s->nb[2]) return fail(GX_EINVAL, "communicator size %d != number of blocks %d", nranks, s->nb[0] * s->nb[1] * s->nb[2]);
  if (rank != s->rank) return fail(GX_EINVAL, "rank %d does not own block coords (%d,%d,%d) = rank %d", rank, s->co[0], s->co[1], s->co[2], s->rank);
  int rc = nccl_load(); if (rc) return rc;
  cudaSetDevice(s->device);
  ncclUniqueId id; memcpy(&id, idp, sizeof id);
  NCCL_TRY(g_nccl.CommInitRank(&s->comm, nranks, id, rank));
  s->nranks = nranks;
  // z slabs: map the neighbours' arrays (CUDA IPC) for the peer-memory halo push; any failure keeps the NCCL path
  if (s->nb[0] == 1 && s->nb[1] == 1 && s->nb[2] > 1 && !getenv("GX_NO_P2P") && load_stream_memops()) {
    struct Pack { cudaIpcMemHandle_t h[4]; int have_e; int pad[3]; };
    Pack mine; memset(&mine, 0, sizeof mine);
    bool ok = true;
    if (!s->flags) { ok = cudaMalloc((void**)&s->flags, FL_WORDS * sizeof(uint32_t)) == cudaSuccess && cudaMemset(s->flags, 0, FL_WORDS * sizeof(uint32_t)) == cudaSuccess; }
    ok = ok && cudaIpcGetMemHandle(&mine.h[0], s->U) == cudaSuccess && cudaIpcGetMemHandle(&mine.h[1], s->UP) == cudaSuccess &&
         cudaIpcGetMemHandle(&mine.h[3], s->flags) == cudaSuccess;
    if (ok && s->E) { ok = cudaIpcGetMemHandle(&mine.h[2], s->E) == cudaSuccess; mine.have_e = 1; }
    // every rank must take the same branch below: agree on `ok` first (min over ranks)
    struct DevTmp { void* p = nullptr; ~DevTmp() { if (p) cudaFree(p); } } t_ok, t_pk;   // freed on every exit path
    if (cudaMalloc(&t_ok.p, sizeof(int)) != cudaSuccess || cudaMalloc(&t_pk.p, 3 * sizeof(Pack)) != cudaSuccess) { cudaGetLastError(); return fail(GX_ENOMEM, "peer link buffers"); }
    int* d_ok = (int*)t_ok.p; Pack* d_pk = (Pack*)t_pk.p;    // d_pk[0] = mine, [1] = from lo, [2] = from hi
    int okv = ok ? 1 : 0;
    CUDA_TRY(cudaMemcpy(d_ok, &okv, sizeof okv, cudaMemcpyHostToDevice));
    NCCL_TRY(g_nccl.AllReduce(d_ok, d_ok, 1, ncclInt32, ncclMin, s->comm, s->stream));
    CUDA_TRY(cudaMemcpyAsync(&okv, d_ok, sizeof okv, cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    if (okv) {
      const int lo = s->nbr[2][0], hi = s->nbr[2][1];
      CUDA_TRY(cudaMemcpy(d_pk, &mine, sizeof mine, cudaMemcpyHostToDevice));
      NCCL_TRY(g_nccl.GroupStart());                          // same pairing order as the halo exchange
      if (hi >= 0) NCCL_TRY(g_nccl.Send(d_pk, sizeof(Pack), ncclUint8, hi, s->comm, s->stream));
      if (lo >= 0) NCCL_TRY(g_nccl.Recv(d_pk + 1, sizeof(Pack), ncclUint8, lo, s->comm, s->stream));
      if (lo >= 0) NCCL_TRY(g_nccl.Send(d_pk, sizeof(Pack), ncclUint8, lo, s->comm, s->stream));
      if (hi >= 0) NCCL_TRY(g_nccl.Recv(d_pk + 2, sizeof(Pack), ncclUint8, hi, s->comm, s->stream));
      NCCL_TRY(g_nccl.GroupEnd());
      Pack got[3];
      CUDA_TRY(cudaMemcpyAsync(got, d_pk, sizeof got, cudaMemcpyDeviceToHost, s->stream));
      CUDA_TRY(cudaStreamSynchronize(s->stream));
      int opened = 1;
      for (int side = 0; side < 2 && opened; ++side) {
        const int nb = side == 0 ? lo : hi;
        if (nb < 0) continue;
        if (side == 1 && hi == lo) { s->peer[1] = s->peer[0]; continue; }       // two blocks, periodic: one neighbour on both sides
        const Pack& pk = got[1 + side];
        void* ptr[4] = {nullptr, nullptr, nullptr, nullptr};
        for (int q = 0; q < 4; ++q) {
          if (q == 2 && !pk.have_e) continue;
          if (cudaIpcOpenMemHandle(&ptr[q], pk.h[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); opened = 0; break; }
        }
        s->peer[side].U = (double*)ptr[0]; s->peer[side].UP = (double*)ptr[1]; s->peer[side].E = (double*)ptr[2]; s->peer[side].flags = (uint32_t*)ptr[3];
      }
      CUDA_TRY(cudaMemcpy(d_ok, &opened, sizeof opened, cudaMemcpyHostToDevice));
      NCCL_TRY(g_nccl.AllReduce(d_ok, d_ok, 1, ncclInt32, ncclMin, s->comm, s->stream));
      CUDA_TRY(cudaMemcpyAsync(&opened, d_ok, sizeof opened, cudaMemcpyDeviceToHost, s->stream));
      CUDA_TRY(cudaStreamSynchronize(s->stream));
      s->p2p = opened != 0;
      // periodic z (no physical fills at the slab ends), x and y wrapped in the loaders, flux-CD: overlap the push
      s->overlap = s->p2p && s->fused && s->A.flux_cd && s->periodic[2] && s->A.wrap[0] && s->A.wrap[1] && s->A.g.nz >= 16 &&
                   !s->cfg.bc_user && !getenv("GX_NO_OVERLAP");
      if (s->overlap && !s->cstream) {
        int lo_p = 0, hi_p = 0; cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p);
        CUDA_TRY(cudaStreamCreateWithPriority(&s->cstream, cudaStreamNonBlocking, hi_p));
        CUDA_TRY(cudaEventCreateWithFlags(&s->ev_bnd, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&s->ev_comm, cudaEventDisableTiming));
      }
    }
  }
  return GX_OK;
}

int gx_tc_info(const gx_solver* s, double* dt_cond, int32_t* nsteps) {
  if (!s) return fail(GX_EINVAL, "null argument");
  if (dt_cond) *dt_cond = s->tc_dt_cond;
  if (nsteps) *nsteps = s->tc_nsteps;
  return GX_OK;
}

int64_t gx_launch_count(const gx_solver* s) { return s ? s->launches : 0; }
double gx_last_elapsed_ms(const gx_solver* s) { return s ? s->last_ms : 0.0; }
int gx_set_profiling(gx_solver* s, int32_t on) {
  if (!s) return fail(GX_EINVAL, "null argument");
  collect_timed(s);
  s->profiling = on != 0;
  for (int c = 0; c < gx::KC_COUNT; ++c) { s->cls_ms[c] = 0; s->cls_n[c] = 0; }
  return GX_OK;
}
int gx_kernel_time_ms(const gx_solver* s, int32_t which, double* total_ms, int64_t* launches) {
  if (!s || which < 0 || which >= gx::KC_COUNT) return fail(GX_EINVAL, "bad kernel class");
  if (total_ms) *total_ms = s->cls_ms[which];
  if (launches) *launches = s->cls_n[which];
  return GX_OK;
}

}  // extern "C"
