// guacho_oracle.cpp — CPU restatement of Guacho-3D's hydro/MHD time step.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product path (guacho_b200/, the
// C-ABI library) may include, link or call this file; only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
//
// PARITY UNPINNED BY THE REFERENCE: the reference ships no tests, golden
// vectors or benchmark outputs, and no Fortran compiler or MPI exists in this
// image, so the reference itself cannot be run here.  This
// file restates the Fortran loop for loop (same AoS layout, same index ranges,
// same operation order, same pack-all-then-exchange halo semantics) and is
// pinned by (a) tests/test_oracle_golden.py: fixtures computed from the
// PUBLISHED form of each algorithm, independently of this file
// (tests/golden/make_golden.py: Miyoshi-Kusano HLLD in jump-condition form,
// HLL/HLLC, the exact Sod solution, the analytic first Orszag-Tang time step)
// and frozen end-to-end outputs; (b) tests/test_oracle_kat.py: known answers
// and invariants that follow from the reference code; (c) tests/test_bin_io.py:
// the dump format through the reference's own Python reader.
//
// Build: g++ -O2 -ffp-contract=off (the reference is built -O3 without FMA
// contraction on x86-64: OT/Makefile:24,115-123).  Every function cites the
// reference file:line it follows (paths relative to the reference root).
//
// Blocks emulate MPI ranks: block r owns coords (cx,cy,cz) with the row-major
// map r = (cx*NBY + cy)*NBZ + cz; "mpi_sendrecv" is a copy between
// per-block send buffers that are ALL packed before any ghost is written,
// exactly like src/boundaries.f90:70-106.

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#include "../include/guacho_gx.h"   // gx_config + named constants only (no product code)

namespace orc {

// ---------------------------------------------------------------------------
// parameters (runtime mirror of OT/parameters.f90)
struct Par {
  int nxtot, nytot, nztot, NBX, NBY, NBZ;
  int nx, ny, nz;                 // parameters.f90:212-216
  int neq, neqdyn, npas;
  bool mhd, pmhd, passives;
  int riemann_solver, slope_limiter, eq_of_state;
  bool enable_flux_cd, eight_wave, user_source_terms;
  int bc_left, bc_right, bc_bottom, bc_top, bc_out, bc_in;
  bool bc_user;
  double dx, dy, dz, cv, gamma, Tempsc, cfl, eta;
  int cooling; double tsc;        // parameters.f90: cooling, tsc
  int th_cond; bool tc_saturation;                 // parameters.f90:117-119
  double rsc, rhosc, vsc2, vsc, Psc, bsc, mu;      // parameters.f90:159-170 (vsc = sqrt(vsc2), Psc = rhosc*vsc2)
  int nxmin, nxmax, nymin, nymax, nzmin, nzmax;   // parameters.f90:222-227
};

// 4-D array with Fortran indexing u(ieq, i, j, k), ieq = 1..neq, i = nxmin..nxmax
struct Arr4 {
  int n1 = 0, NX = 0, NY = 0, NZ = 0;
  std::vector<double> d;
  void alloc(int n1_, int nx, int ny, int nz) {
    n1 = n1_; NX = nx + 4; NY = ny + 4; NZ = nz + 4;
    d.assign((size_t)n1 * NX * NY * NZ, 0.0);   // zero pages
  }
  inline size_t idx(int ieq, int i, int j, int k) const {
    return (size_t)(ieq - 1) + (size_t)n1 * ((size_t)(i + 1) + (size_t)NX * ((size_t)(j + 1) + (size_t)NY * (size_t)(k + 1)));
  }
  inline double& operator()(int ieq, int i, int j, int k) { return d[idx(ieq, i, j, k)]; }
  inline const double& operator()(int ieq, int i, int j, int k) const { return d[idx(ieq, i, j, k)]; }
  inline double* cell(int i, int j, int k) { return &d[idx(1, i, j, k)]; }
  inline const double* cell(int i, int j, int k) const { return &d[idx(1, i, j, k)]; }
};

typedef void (*user_bc_fn)(double* u, int order, const int* coords, double time, void* ctx);
typedef void (*user_src_fn)(const double* pp, double* s, int i, int j, int k, const int* coords, void* ctx);

struct Block {
  int rank;
  int coords[3];
  int left, right, bottom, top, out, in;   // src/init.f90:108-110 ; -1 = MPI_PROC_NULL
  Arr4 u, up, primit, f, g, h, e;          // src/globals.f90:33-39, flux_cd_module.f90:35
  Arr4 primit0;                            // src/globals.f90:42 (split-all solvers only: background primitives, set by the host)
  std::vector<double> Temp;                // (nxmin:nxmax, nymin:nymax, nzmin:nzmax)
  // halo buffers (src/boundaries.f90:56-58, 271-273)
  std::vector<double> sendr, sendl, sendt, sendb, sendi, sendo;
};

struct Oracle {
  Par P;
  std::vector<Block> B;
  double time = 0.0;
  int nthreads = 1;
  user_bc_fn user_bc = nullptr;   void* user_bc_ctx = nullptr;
  user_src_fn user_src = nullptr; void* user_src_ctx = nullptr;
  int builtin_user = 0;           // 0 none, 1 EXO (exoplanet.f90 + EXO/user_mod.f90)
  // EXO module state (EXO/exoplanet.f90:31-49)
  double RSW, TSW, VSW, dsw, RsS, bsw, bpw, RPW, TPW, VPW, dpw, torb, rorb, omegap, MassS, MassP, xp, yp, zp;
  double exo_rsc = 1, exo_vsc2 = 1;
  bool error_flag = false;
  // thermal conduction log (what the reference writes to logs/thermal_conduction.log, thermal_cond.f90:725)
  double tc_dt_cond = 0.0; int tc_nsteps = 0;
};

static inline double sign1(double x) { return std::copysign(1.0, x); }   // Fortran sign(1.,x), SURVEY Q11

// ---------------------------------------------------------------------------
// src/hydro_core.f90:46-129  u2prim
static void u2prim(const Par& P, const double* uu, double* prim, double& T) {
  double r = std::max(uu[0], 1e-15);                                   // :62
  prim[0] = r;
  prim[1] = uu[1] / r;                                                  // :65-67
  prim[2] = uu[2] / r;
  prim[3] = uu[3] / r;
  if (P.mhd) {                                                          // :69-76
    prim[4] = (uu[4] - 0.5 * r * (prim[1] * prim[1] + prim[2] * prim[2] + prim[3] * prim[3])
                     - 0.5 * (uu[5] * uu[5] + uu[6] * uu[6] + uu[7] * uu[7])) / P.cv;
  } else {
    prim[4] = (uu[4] - 0.5 * r * (prim[1] * prim[1] + prim[2] * prim[2] + prim[3] * prim[3])) / P.cv;
  }
  prim[4] = std::max(prim[4], 1e-16);                                   // :78
  if (P.mhd || P.pmhd) { prim[5] = uu[5]; prim[6] = uu[6]; prim[7] = uu[7]; }   // :80-84
  if (P.passives) for (int q = P.neqdyn; q < P.neq; ++q) prim[q] = uu[q];       // :86-90
  T = 0.0;
  if (P.eq_of_state == GX_EOS_ADIABATIC) T = (prim[4] / r) * P.Tempsc;          // :93-95
  if (P.eq_of_state == GX_EOS_SINGLE_SPECIE) {                                  // :97-102
    r = std::max(r, 1e-15);
    T = std::max(1.0, (prim[4] / r) * P.Tempsc);
    prim[4] = r * T / P.Tempsc;
  }
  if (P.passives && P.eq_of_state == GX_EOS_H_RATE) {                           // :107-112
    double dentot = (2.0 * r - prim[P.neqdyn]);
    dentot = std::max(dentot, 1e-15);
    T = std::max(1.0, (prim[4] / dentot) * P.Tempsc);
    prim[4] = dentot * T / P.Tempsc;
  }
}

static inline bool split_all(const Par& P) { return P.riemann_solver == GX_SOLVER_HLLE_SPLIT_ALL || P.riemann_solver == GX_SOLVER_HLLD_SPLIT_ALL; }

// src/hydro_core.f90:143-229  u2primSplitAll: uu and prim are FLUCTUATIONS about the background prim0; no floors
static void u2primSplitAll(const Par& P, const double* uu, double* prim, const double* prim0, double& T) {
  prim[0] = uu[0];                                                      // :159
  double r = prim[0] + prim0[0];                                        // :161
  prim[1] = uu[1] / r;
  prim[2] = uu[2] / r;
  prim[3] = uu[3] / r;
  if (P.mhd || P.pmhd) { prim[5] = uu[5]; prim[6] = uu[6]; prim[7] = uu[7]; }   // :168-170
  if (P.mhd) {                                                          // :174-179
    prim[4] = (uu[4] - 0.5 * r * (prim[1] * prim[1] + prim[2] * prim[2] + prim[3] * prim[3])
                     - 0.5 * (uu[5] * uu[5] + uu[6] * uu[6] + uu[7] * uu[7])
                     - prim0[5] * uu[5] - prim0[6] * uu[6] - prim0[7] * uu[7]) / P.cv;
  } else {
    prim[4] = (uu[4] - 0.5 * r * (prim[1] * prim[1] + prim[2] * prim[2] + prim[3] * prim[3])) / P.cv;
  }
  if (P.passives) for (int q = P.neqdyn; q < P.neq; ++q) prim[q] = uu[q];       // :187-189
  T = 0.0;
  if (P.eq_of_state == GX_EOS_ADIABATIC) T = ((prim[4] + prim0[4]) / r) * P.Tempsc;          // :194-196
  if (P.eq_of_state == GX_EOS_SINGLE_SPECIE) {                                                // :198-203
    r = std::max(r, 1e-15);
    T = std::max(1., ((prim[4] + prim0[4]) / r) * P.Tempsc);
  }
  if (P.passives && P.eq_of_state == GX_EOS_H_RATE) {                                         // :207-212
    double dentot = (2. * r - prim[P.neqdyn] - prim0[P.neqdyn]);
    dentot = std::max(dentot, 1e-15);
    T = std::max(1., ((prim[4] + prim0[4]) / dentot) * P.Tempsc);
  }
}

// src/hydro_core.f90:245-320  calcprim
static void calcprim(const Par& P, const Arr4& u, Arr4& primit, std::vector<double>& Temp, bool only_ghost = false, const Arr4* primit0 = nullptr) {
  const int NX = u.NX, NY = u.NY;
  auto tix = [&](int i, int j, int k) { return (size_t)(i + 1) + (size_t)NX * ((size_t)(j + 1) + (size_t)NY * (size_t)(k + 1)); };
  auto one = [&](int i, int j, int k) {
    if (split_all(P)) u2primSplitAll(P, u.cell(i, j, k), primit.cell(i, j, k), primit0->cell(i, j, k), Temp[tix(i, j, k)]);   // :263-266, 307-309
    else u2prim(P, u.cell(i, j, k), primit.cell(i, j, k), Temp[tix(i, j, k)]);
  };
  if (only_ghost) {                                                     // :258-301
    for (int j = 0; j <= P.ny + 1; ++j) for (int i = 0; i <= P.nx + 1; ++i) { one(i, j, 0); one(i, j, P.nz + 1); }
    for (int k = 0; k <= P.nz + 1; ++k) for (int i = 0; i <= P.nx + 1; ++i) { one(i, 0, k); one(i, P.ny + 1, k); }
    for (int k = 0; k <= P.nz + 1; ++k) for (int j = 0; j <= P.ny + 1; ++j) { one(0, j, k); one(P.nx + 1, j, k); }
  } else {                                                              // :304-316
    for (int k = P.nzmin; k <= P.nzmax; ++k)
      for (int j = P.nymin; j <= P.nymax; ++j)
        for (int i = P.nxmin; i <= P.nxmax; ++i) one(i, j, k);
  }
}

// src/hydro_core.f90:331-383  prim2u (non-split branch)
static void prim2u(const Par& P, const double* prim, double* uu) {
  uu[0] = prim[0];
  uu[1] = prim[0] * prim[1];
  uu[2] = prim[0] * prim[2];
  uu[3] = prim[0] * prim[3];
  uu[4] = 0.5 * prim[0] * (prim[1] * prim[1] + prim[2] * prim[2] + prim[3] * prim[3]) + P.cv * prim[4];   // :356
  if (P.mhd) uu[4] = uu[4] + 0.5 * (prim[5] * prim[5] + prim[6] * prim[6] + prim[7] * prim[7]);        // :366
  if (P.mhd || P.pmhd) { uu[5] = prim[5]; uu[6] = prim[6]; uu[7] = prim[7]; }
  if (P.passives) for (int q = P.neqdyn; q < P.neq; ++q) uu[q] = prim[q];
}

// src/hydro_core.f90:395-476  prim2f (non-split branches)
static void prim2f(const Par& P, const double* prim, double* ff) {
  double etot;
  if (P.mhd) {                                                          // :432-441
    etot = 0.5 * (prim[0] * (prim[1] * prim[1] + prim[2] * prim[2] + prim[3] * prim[3])
                  + prim[5] * prim[5] + prim[6] * prim[6] + prim[7] * prim[7])
           + P.cv * prim[4];
    ff[0] = prim[0] * prim[1];
    ff[1] = prim[0] * prim[1] * prim[1] + prim[4] + 0.5 * (prim[6] * prim[6] + prim[7] * prim[7] - prim[5] * prim[5]);
    ff[2] = prim[0] * prim[1] * prim[2] - prim[5] * prim[6];
    ff[3] = prim[0] * prim[1] * prim[3] - prim[5] * prim[7];
    ff[4] = prim[1] * (etot + prim[4] + 0.5 * (prim[5] * prim[5] + prim[6] * prim[6] + prim[7] * prim[7]))
            - prim[5] * (prim[1] * prim[5] + prim[2] * prim[6] + prim[3] * prim[7]);
  } else {                                                              // :446-452
    etot = 0.5 * prim[0] * (prim[1] * prim[1] + prim[2] * prim[2] + prim[3] * prim[3]) + P.cv * prim[4];
    ff[0] = prim[0] * prim[1];
    ff[1] = prim[0] * prim[1] * prim[1] + prim[4];
    ff[2] = prim[0] * prim[1] * prim[2];
    ff[3] = prim[0] * prim[1] * prim[3];
    ff[4] = prim[1] * (etot + prim[4]);
  }
  if (P.mhd || P.pmhd) {                                                // :463-465
    ff[5] = 0.0;
    ff[6] = prim[1] * prim[6] - prim[5] * prim[2];
    ff[7] = prim[1] * prim[7] - prim[5] * prim[3];
  }
  if (P.passives) for (int q = P.neqdyn; q < P.neq; ++q) ff[q] = prim[q] * prim[1];   // :472
}

// src/hydro_core.f90:485-534  swapy / swapz
static inline void swapy(const Par& P, double* var) {
  std::swap(var[1], var[2]);
  if (P.mhd || P.pmhd) std::swap(var[5], var[6]);
}
static inline void swapz(const Par& P, double* var) {
  std::swap(var[1], var[3]);
  if (P.mhd || P.pmhd) std::swap(var[5], var[7]);
}

// src/hydro_core.f90:544-553  csound
static inline double csound(const Par& P, double p, double d) { return std::sqrt(P.gamma * p / d); }

// src/hydro_core.f90:568-581  cfast
static inline void cfast(const Par& P, double p, double d, double bx, double by, double bz, double& cfx, double& cfy, double& cfz) {
  const double g = P.gamma;
  double b2 = bx * bx + by * by + bz * bz;
  cfx = std::sqrt(0.5 * ((g * p + b2) + std::sqrt((g * p + b2) * (g * p + b2) - 4. * g * p * bx * bx)) / d);
  cfy = std::sqrt(0.5 * ((g * p + b2) + std::sqrt((g * p + b2) * (g * p + b2) - 4. * g * p * by * by)) / d);
  cfz = std::sqrt(0.5 * ((g * p + b2) + std::sqrt((g * p + b2) * (g * p + b2) - 4. * g * p * bz * bz)) / d);
}

// src/hydro_core.f90:591-604  cfastX
static inline double cfastX(const Par& P, const double* prim) {
  double b2 = prim[5] * prim[5] + prim[6] * prim[6] + prim[7] * prim[7];
  double cs2va2 = (P.gamma * prim[4] + b2) / prim[0];
  return std::sqrt(0.5 * (cs2va2 + std::sqrt(cs2va2 * cs2va2 - 4. * P.gamma * prim[4] * (prim[5] * prim[5]) / prim[0] / prim[0])));
}

// src/hydro_core.f90:735-796  average  (the eight slope limiters)
static inline double average(int slope_limiter, double a, double b) {
  const double delta = 1.e-7;                                           // :721
  double avg = 0.0, s, c, d, av1, av2;
  switch (slope_limiter) {
    case GX_LIMITER_NO_AVERAGE: avg = 0.; break;                        // :741-744
    case GX_LIMITER_NO_LIMIT: avg = 0.5 * (a + b); break;               // :746-749
    case GX_LIMITER_MINMOD:                                             // :751-755
      s = sign1(a);
      avg = s * std::max(0., std::min(std::fabs(a), s * b));
      break;
    case GX_LIMITER_VAN_LEER:                                           // :757-764
      if (a * b <= 0.) avg = 0.;
      else avg = a * b * (a + b) / (a * a + b * b);
      break;
    case GX_LIMITER_VAN_ALBADA:                                         // :766-769
      avg = (a * (b * b + delta) + b * (a * a + delta)) / (a * a + b * b + delta);
      break;
    case GX_LIMITER_UMIST:                                              // :771-778
      s = sign1(a);
      c = 0.25 * a + 0.75 * b;
      d = 0.75 * a + 0.25 * b;
      avg = std::min(std::min(2. * std::fabs(a), 2. * s * b), std::min(s * c, s * d));
      avg = s * std::max(0., avg);
      break;
    case GX_LIMITER_WOODWARD:                                           // :780-786
      s = sign1(a);
      c = 0.5 * (a + b);
      avg = std::min(std::min(2. * std::fabs(a), 2. * s * b), s * c);
      avg = s * std::max(0., avg);
      break;
    case GX_LIMITER_SUPERBEE:                                           // :788-794
      s = sign1(b);
      av1 = std::min(2. * std::fabs(b), s * a);
      av2 = std::min(std::fabs(b), 2. * s * a);
      avg = s * std::max(0., std::max(av1, av2));
      break;
  }
  return avg;
}

// src/hydro_core.f90:712-731  limiter
static void limiter(int slope_limiter, const double* pll, double* pl, double* pr, const double* prr, int neq) {
  for (int q = 0; q < neq; ++q) {
    double dl = pl[q] - pll[q];
    double dm = pr[q] - pl[q];
    double dr = prr[q] - pr[q];
    double al = average(slope_limiter, dl, dm);
    double ar = average(slope_limiter, dm, dr);
    pl[q] = pl[q] + al * 0.5;
    pr[q] = pr[q] - ar * 0.5;
  }
}

// ---------------------------------------------------------------------------
// src/hll.f90:47-82 prim2fhll  and  src/hlle.f90:48-83 prim2fhlle
// (identical except csound vs cfastX for the signal speeds)
static void prim2fhll_e(const Par& P, bool fast, const double* priml, const double* primr, double* ff) {
  double csl, csr;
  if (fast) { csl = cfastX(P, priml); csr = cfastX(P, primr); }
  else { csl = csound(P, priml[4], priml[0]); csr = csound(P, primr[4], primr[0]); }
  double sr = std::max(priml[1] + csl, primr[1] + csr);
  double sl = std::min(priml[1] - csl, primr[1] - csr);
  if (sl > 0) { prim2f(P, priml, ff); return; }
  if (sr < 0) { prim2f(P, primr, ff); return; }
  double fL[16], fR[16], uL[16], uR[16];
  prim2f(P, priml, fL); prim2f(P, primr, fR);
  prim2u(P, priml, uL); prim2u(P, primr, uR);
  // which entries prim2f/prim2u define: 1..5, 6..8 if (mhd|pmhd), passives
  for (int q = 0; q < P.neq; ++q) {
    bool defined = q < 5 || (q < 8 && (P.mhd || P.pmhd) && P.neqdyn == 8) || (q >= P.neqdyn && P.passives);
    if (defined) ff[q] = (sr * fL[q] - sl * fR[q] + sl * sr * (uR[q] - uL[q])) / (sr - sl);
  }
}

// src/hllc.f90:44-140 prim2fhllc
static int prim2fhllc(const Par& P, const double* priml, const double* primr, double* ff) {
  double uu[16], uuk[16];
  for (int q = 0; q < 16; ++q) uuk[q] = 0.0;   // reference leaves uuk(6:8) unset unless pmhd
  double csl = csound(P, priml[4], priml[0]);
  double csr = csound(P, primr[4], primr[0]);
  double sr = std::max(priml[1] + csl, primr[1] + csr);
  double sl = std::min(priml[1] - csl, primr[1] - csr);
  if (sl > 0) { prim2f(P, priml, ff); return 0; }
  if (sr < 0) { prim2f(P, primr, ff); return 0; }
  double slmul = sl - priml[1];
  double srmur = sr - primr[1];
  double rholul = priml[0] * priml[1];
  double rhorur = primr[0] * primr[1];
  double sst = (srmur * rhorur - slmul * rholul - primr[4] + priml[4]) / (srmur * primr[0] - slmul * priml[0]);   // :76-77
  const int nset = P.neq;
  if (sst >= 0.) {                                                      // :79-107
    double rhost = priml[0] * (slmul) / (sl - sst);
    double ek = 0.5 * priml[0] * (priml[1] * priml[1] + priml[2] * priml[2] + priml[3] * priml[3]) + P.cv * priml[4];
    uuk[0] = rhost;
    uuk[1] = rhost * sst;
    uuk[2] = rhost * priml[2];
    uuk[3] = rhost * priml[3];
    uuk[4] = rhost * (ek / priml[0] + (sst - priml[1]) * (sst + priml[4] / (priml[0] * slmul)));
    if (P.pmhd) for (int q = 5; q < 8; ++q) uuk[q] = rhost * priml[q] / priml[0];
    if (P.passives) for (int q = P.neqdyn; q < P.neq; ++q) uuk[q] = rhost * priml[q] / priml[0];
    prim2f(P, priml, ff);
    prim2u(P, priml, uu);
    for (int q = 0; q < nset; ++q) ff[q] = ff[q] + sl * (uuk[q] - uu[q]);
    return 0;
  }
  if (sst <= 0.) {                                                      // :109-133
    double rhost = primr[0] * (srmur) / (sr - sst);
    double ek = 0.5 * primr[0] * (primr[1] * primr[1] + primr[2] * primr[2] + primr[3] * primr[3]) + P.cv * primr[4];
    uuk[0] = rhost;
    uuk[1] = rhost * sst;
    uuk[2] = rhost * primr[2];
    uuk[3] = rhost * primr[3];
    uuk[4] = rhost * (ek / primr[0] + (sst - primr[1]) * (sst + primr[4] / (primr[0] * srmur)));
    if (P.pmhd) for (int q = 5; q < 8; ++q) uuk[q] = rhost * primr[q] / primr[0];
    if (P.passives) for (int q = P.neqdyn; q < P.neq; ++q) uuk[q] = rhost * primr[q] / primr[0];
    prim2f(P, primr, ff);
    prim2u(P, primr, uu);
    for (int q = 0; q < nset; ++q) ff[q] = ff[q] + sr * (uuk[q] - uu[q]);
    return 0;
  }
  return 1;   // 'Error in hllc' + stop (:135-138): NaN input
}

// src/hlld.f90:48-319 prim2fhlld
static int prim2fhlld(const Par& P, const double* priml, const double* primr, double* ff) {
  const double cv = P.cv;
  double csl = cfastX(P, priml);                                        // :64-65
  double csr = cfastX(P, primr);
  double sr = std::max(priml[1] + csl, primr[1] + csr);                 // :67-68
  double sl = std::min(priml[1] - csl, primr[1] - csr);
  if (sl > 0) { prim2f(P, priml, ff); return 0; }                       // :71-74
  if (sr < 0) { prim2f(P, primr, ff); return 0; }                       // :77-80

  double bx = 0.5 * (priml[5] + primr[5]);                              // :82
  double signBx = sign1(bx);                                            // :83
  double pTL = priml[4] + 0.5 * (bx * bx + priml[6] * priml[6] + priml[7] * priml[7]);   // :86-87
  double pTR = primr[4] + 0.5 * (bx * bx + primr[6] * primr[6] + primr[7] * primr[7]);
  double slmul = sl - priml[1];                                         // :89-90
  double srmur = sr - primr[1];
  double rholul = priml[0] * priml[1];                                  // :92-93
  double rhorur = primr[0] * primr[1];
  double sM = (srmur * rhorur - slmul * rholul - pTR + pTL) / (srmur * primr[0] - slmul * priml[0]);   // :95
  double srmsM = sr - sM;                                               // :97-98
  double slmsM = sl - sM;
  double rhostl = priml[0] * slmul / slmsM;                             // :100-101
  double rhostr = primr[0] * srmur / srmsM;
  double sstl = sM - std::fabs(bx) / std::sqrt(rhostl);                 // :103-104
  double sstr = sM + std::fabs(bx) / std::sqrt(rhostr);
  double pst = (srmur * primr[0] * pTL - slmul * priml[0] * pTR
                + priml[0] * primr[0] * srmur * slmul * (primr[1] - priml[1]))
               / (srmur * primr[0] - slmul * priml[0]);                 // :106-108

  double el, er, denl, denr, sMmul, sMmur;
  double vstl = 0, wstl = 0, bystl = 0, bzstl = 0, estl, vdotbl, vstdotbstl;
  double vstr = 0, wstr = 0, bystr = 0, bzstr = 0, estr, vdotbr, vstdotbstr;

  if (sstl >= 0) {                                                      // UL*  :111-156
    el = 0.5 * priml[0] * (priml[1] * priml[1] + priml[2] * priml[2] + priml[3] * priml[3]) + cv * priml[4]
         + 0.5 * (bx * bx + priml[6] * priml[6] + priml[7] * priml[7]);
    sMmul = sM - priml[1];
    denl = priml[0] * slmul * slmsM - bx * bx;
    if (denl == 0) {
      vstl = priml[2]; wstl = priml[3]; bystl = 0.; bzstl = 0.;
    } else {
      vstl = priml[2] - bx * priml[6] * sMmul / denl;
      wstl = priml[3] - bx * priml[7] * sMmul / denl;
      bystl = priml[6] * (priml[0] * (slmul * slmul) - bx * bx) / denl;
      bzstl = priml[7] * (priml[0] * (slmul * slmul) - bx * bx) / denl;
    }
    vdotbl = priml[1] * bx + priml[2] * priml[6] + priml[3] * priml[7];
    vstdotbstl = sM * bx + vstl * bystl + wstl * bzstl;
    estl = (slmul * el - pTL * priml[1] + pst * sM + bx * (vdotbl - vstdotbstl)) / slmsM;
    ff[0] = rhostl * sM;
    ff[1] = rhostl * (sM * sM) + pst - bx * bx;
    ff[2] = rhostl * sM * vstl - bx * bystl;
    ff[3] = rhostl * sM * wstl - bx * bzstl;
    ff[4] = sM * (estl + pst) - bx * (vstdotbstl);
    ff[5] = 0.;
    ff[6] = bystl * sM - bx * vstl;
    ff[7] = bzstl * sM - bx * wstl;
    if (P.passives) for (int q = P.neqdyn; q < P.neq; ++q) ff[q] = sM * priml[q] * slmul / slmsM;
    return 0;
  }

  if (sstr <= 0) {                                                      // UR*  :159-204
    er = 0.5 * primr[0] * (primr[1] * primr[1] + primr[2] * primr[2] + primr[3] * primr[3]) + cv * primr[4]
         + 0.5 * (bx * bx + primr[6] * primr[6] + primr[7] * primr[7]);
    sMmur = sM - primr[1];
    denr = primr[0] * srmur * srmsM - bx * bx;
    if (denr == 0) {
      // reference assigns the L variables here; vstr.. stay undefined (0 here)
      vstl = priml[2]; wstl = priml[3]; bystl = 0.; bzstl = 0.;
    } else {
      vstr = primr[2] - bx * primr[6] * sMmur / denr;
      wstr = primr[3] - bx * primr[7] * sMmur / denr;
      bystr = primr[6] * (primr[0] * (srmur * srmur) - bx * bx) / denr;
      bzstr = primr[7] * (primr[0] * (srmur * srmur) - bx * bx) / denr;
    }
    vdotbr = primr[1] * bx + primr[2] * primr[6] + primr[3] * primr[7];
    vstdotbstr = sM * bx + vstr * bystr + wstr * bzstr;
    estr = (srmur * er - pTR * primr[1] + pst * sM + bx * (vdotbr - vstdotbstr)) / srmsM;
    ff[0] = rhostr * sM;
    ff[1] = rhostr * (sM * sM) + pst - bx * bx;
    ff[2] = rhostr * sM * vstr - bx * bystr;
    ff[3] = rhostr * sM * wstr - bx * bzstr;
    ff[4] = sM * (estr + pst) - bx * (vstdotbstr);
    ff[5] = 0.;
    ff[6] = bystr * sM - bx * vstr;
    ff[7] = bzstr * sM - bx * wstr;
    if (P.passives) for (int q = P.neqdyn; q < P.neq; ++q) ff[q] = sM * primr[q] * srmur / srmsM;
    return 0;
  }

  // needed on both UL** and UR**  :207-252
  sMmul = sM - priml[1];
  sMmur = sM - primr[1];
  denl = priml[0] * slmul * slmsM - bx * bx;
  denr = primr[0] * srmur * srmsM - bx * bx;
  if (denl == 0) {
    vstl = priml[2]; wstl = priml[3]; bystl = 0.; bzstl = 0.;
  } else {
    vstl = priml[2] - bx * priml[6] * sMmul / denl;
    wstl = priml[3] - bx * priml[7] * sMmul / denl;
    bystl = priml[6] * (priml[0] * (slmul * slmul) - bx * bx) / denl;
    bzstl = priml[7] * (priml[0] * (slmul * slmul) - bx * bx) / denl;
  }
  if (denr == 0) {
    vstr = primr[2]; wstr = primr[3]; bystr = 0.; bzstr = 0.;
  } else {
    vstr = primr[2] - bx * primr[6] * sMmur / denr;
    wstr = primr[3] - bx * primr[7] * sMmur / denr;
    bystr = primr[6] * (primr[0] * (srmur * srmur) - bx * bx) / denr;
    bzstr = primr[7] * (primr[0] * (srmur * srmur) - bx * bx) / denr;
  }
  double dd = std::sqrt(rhostl) + std::sqrt(rhostr);                    // :241
  double vstst = (std::sqrt(rhostl) * vstl + std::sqrt(rhostr) * vstr + (bystr - bystl) * signBx) / dd;
  double wstst = (std::sqrt(rhostl) * wstl + std::sqrt(rhostr) * wstr + (bzstr - bzstl) * signBx) / dd;
  double bystst = (std::sqrt(rhostl) * bystr + std::sqrt(rhostr) * bystl + std::sqrt(rhostl * rhostr) * (vstr - vstl) * signBx) / dd;
  double bzstst = (std::sqrt(rhostl) * bzstr + std::sqrt(rhostr) * bzstl + std::sqrt(rhostl * rhostr) * (wstr - wstl) * signBx) / dd;
  double vststdotbstst = sM * bx + vstst * bystst + wstst * bzstst;     // :252

  if (sM >= 0) {                                                        // UL**  :255-283
    el = 0.5 * priml[0] * (priml[1] * priml[1] + priml[2] * priml[2] + priml[3] * priml[3]) + cv * priml[4]
         + 0.5 * (bx * bx + priml[6] * priml[6] + priml[7] * priml[7]);
    vdotbl = priml[1] * bx + priml[2] * priml[6] + priml[3] * priml[7];
    vstdotbstl = sM * bx + vstl * bystl + wstl * bzstl;
    estl = (slmul * el - pTL * priml[1] + pst * sM + bx * (vdotbl - vstdotbstl)) / slmsM;
    double eststl = estl - std::sqrt(rhostl) * (vstdotbstl - vststdotbstst) * signBx;
    ff[0] = rhostl * sM;
    ff[1] = rhostl * (sM * sM) + pst - bx * bx;
    ff[2] = rhostl * sM * vstst - bx * bystst;
    ff[3] = rhostl * sM * wstst - bx * bzstst;
    ff[4] = sM * (eststl + pst) - bx * (vststdotbstst);
    ff[5] = 0.;
    ff[6] = bystst * sM - bx * vstst;
    ff[7] = bzstst * sM - bx * wstst;
    if (P.passives) for (int q = P.neqdyn; q < P.neq; ++q) ff[q] = sM * priml[q] * slmul / slmsM;
    return 0;
  }
  if (sM <= 0) {                                                        // UR**  :286-314
    er = 0.5 * primr[0] * (primr[1] * primr[1] + primr[2] * primr[2] + primr[3] * primr[3]) + cv * primr[4]
         + 0.5 * (bx * bx + primr[6] * primr[6] + primr[7] * primr[7]);
    vdotbr = primr[1] * bx + primr[2] * primr[6] + primr[3] * primr[7];
    vstdotbstr = sM * bx + vstr * bystr + wstr * bzstr;
    estr = (srmur * er - pTR * primr[1] + pst * sM + bx * (vdotbr - vstdotbstr)) / srmsM;
    double eststr = estr + std::sqrt(rhostr) * (vstdotbstr - vststdotbstst) * signBx;
    ff[0] = rhostr * sM;
    ff[1] = rhostr * (sM * sM) + pst - bx * bx;
    ff[2] = rhostr * sM * vstst - bx * bystst;
    ff[3] = rhostr * sM * wstst - bx * bzstst;
    ff[4] = sM * (eststr + pst) - bx * (vststdotbstst);
    ff[5] = 0.;
    ff[6] = bystst * sM - bx * vstst;
    ff[7] = bzstst * sM - bx * wstst;
    if (P.passives) for (int q = P.neqdyn; q < P.neq; ++q) ff[q] = sM * primr[q] * srmur / srmsM;
    return 0;
  }
  return 1;   // 'Error in HLLD routine' + stop  :316-317 (NaN)
}

static int riemann(const Par& P, const double* pl, const double* pr, double* ff) {
  switch (P.riemann_solver) {
    case GX_SOLVER_HLL:  prim2fhll_e(P, false, pl, pr, ff); return 0;
    case GX_SOLVER_HLLE: prim2fhll_e(P, true, pl, pr, ff); return 0;
    case GX_SOLVER_HLLC: return prim2fhllc(P, pl, pr, ff);
    case GX_SOLVER_HLLD: return prim2fhlld(P, pl, pr, ff);
  }
  return 2;
}

// src/hll.f90:93-194 / hllc.f90:152-253 / hlle.f90:95-196 / hlld.f90:331-432
// (the four sweeps are textually identical except for the solver call)
static int fluxes(const Par& P, Block& b, int choice) {
  const int neq = P.neq;
  double priml[16], primr[16], primll[16], primrr[16], ff[16];
  int err = 0;
  for (int q = 0; q < 16; ++q) ff[q] = 0.0;
  auto ld = [&](double* dst, int i, int j, int k) { std::memcpy(dst, b.primit.cell(i, j, k), sizeof(double) * neq); };
  for (int k = 0; k <= P.nz; ++k)
    for (int j = 0; j <= P.ny; ++j)
      for (int i = 0; i <= P.nx; ++i) {
        // x
        ld(priml, i, j, k); ld(primr, i + 1, j, k);
        if (choice == 2) { ld(primll, i - 1, j, k); ld(primrr, i + 2, j, k); limiter(P.slope_limiter, primll, priml, primr, primrr, neq); }
        err |= riemann(P, priml, primr, ff);
        std::memcpy(b.f.cell(i, j, k), ff, sizeof(double) * neq);
        // y
        ld(priml, i, j, k); ld(primr, i, j + 1, k);
        swapy(P, priml); swapy(P, primr);
        if (choice == 2) {
          ld(primll, i, j - 1, k); ld(primrr, i, j + 2, k);
          swapy(P, primll); swapy(P, primrr);
          limiter(P.slope_limiter, primll, priml, primr, primrr, neq);
        }
        err |= riemann(P, priml, primr, ff);
        swapy(P, ff);
        std::memcpy(b.g.cell(i, j, k), ff, sizeof(double) * neq);
        // z
        ld(priml, i, j, k); ld(primr, i, j, k + 1);
        swapz(P, priml); swapz(P, primr);
        if (choice == 2) {
          ld(primll, i, j, k - 1); ld(primrr, i, j, k + 2);
          swapz(P, primll); swapz(P, primrr);
          limiter(P.slope_limiter, primll, priml, primr, primrr, neq);
        }
        err |= riemann(P, priml, primr, ff);
        swapz(P, ff);
        std::memcpy(b.h.cell(i, j, k), ff, sizeof(double) * neq);
      }
  return err;
}

// ---------------------------------------------------------------------------
// HLLE with every variable split into background + fluctuation (src/hlle_split_all.f90; marked unfinished upstream: "REVISAR")
// src/hydro_core.f90:340-368  prim2u, split branch (prim0 present)
static void prim2u_split(const Par& P, const double* prim, const double* prim0, double* uu) {
  uu[0] = prim[0];
  uu[1] = prim[1] * (prim[0] + prim0[0]);
  uu[2] = prim[2] * (prim[0] + prim0[0]);
  uu[3] = prim[3] * (prim[0] + prim0[0]);
  uu[4] = 0.5 * prim[0] * (prim[1] * prim[1] + prim[2] * prim[2] + prim[3] * prim[3]) + P.cv * prim[4];   // :356
  if (P.mhd) {                                                          // :362-364
    uu[4] = uu[4] + 0.5 * (prim[5] * prim[5] + prim[6] * prim[6] + prim[7] * prim[7])
                  + 0.5 * prim0[0] * (prim[1] * prim[1] + prim[2] * prim[2] + prim[3] * prim[3])
                  + prim0[5] * prim[5] + prim0[6] * prim[6] + prim0[7] * prim[7];
  }
  if (P.mhd || P.pmhd) { uu[5] = prim[5]; uu[6] = prim[6]; uu[7] = prim[7]; }
  if (P.passives) for (int q = P.neqdyn; q < P.neq; ++q) uu[q] = prim[q];
}
// src/hydro_core.f90:404-427, 457-461  prim2f, split branch (mhd, prim0 present)
static void prim2f_split(const Par& P, const double* prim, const double* prim0, double* ff) {
  const double rt = prim[0] + prim0[0];
  const double etot = 0.5 * (rt * (prim[1] * prim[1] + prim[2] * prim[2] + prim[3] * prim[3])
                             + prim[5] * prim[5] + prim[6] * prim[6] + prim[7] * prim[7])
                      + P.cv * prim[4]
                      + prim0[5] * prim[5] + prim0[6] * prim[6] + prim0[7] * prim[7];
  ff[0] = rt * prim[1];
  ff[1] = rt * prim[1] * prim[1] + prim[4] + 0.5 * (prim[6] * prim[6] + prim[7] * prim[7] - prim[5] * prim[5])
          - prim0[5] * prim[5] + prim0[6] * prim[6] + prim0[7] * prim[7];
  ff[2] = rt * prim[1] * prim[2] - prim[5] * prim[6]
          - prim0[6] * prim[5] - prim0[5] * prim[6];
  ff[3] = rt * prim[1] * prim[3] - prim[5] * prim[7]
          - prim0[7] * prim[5] - prim0[5] * prim[7];
  ff[4] = prim[1] * (etot + prim[4] + 0.5 * ((prim[5] + prim0[5]) * (prim[5] + prim0[5]) + (prim[6] + prim0[6]) * (prim[6] + prim0[6])
                                              + (prim[7] + prim0[7]) * (prim[7] + prim0[7]))
                     + P.cv * prim0[4] + prim0[4] + 0.5 * (prim0[5] * prim0[5] + prim0[6] * prim0[6] + prim0[7] * prim0[7]))
          - (prim[5] + prim0[5]) * (prim[1] * (prim[5] + prim0[5]) + prim[2] * (prim[6] + prim0[6]) + prim[3] * (prim[7] + prim0[7]));
  ff[5] = 0.;
  ff[6] = prim[1] * (prim0[6] + prim[6]) - prim[2] * (prim0[5] + prim[5]);
  ff[7] = prim[1] * (prim0[7] + prim[7]) - prim[3] * (prim0[5] + prim[5]);
  if (P.passives) for (int q = P.neqdyn; q < P.neq; ++q) ff[q] = prim[q] * prim[1];
}
// src/hlle_split_all.f90:51-85  prim2fhlleSplitAll
static void prim2fhlleSplitAll(const Par& P, const double* priml, const double* primr, const double* prim0l, const double* prim0r, double* ff) {
  double tl[16], tr[16];
  for (int q = 0; q < P.neq; ++q) { tl[q] = priml[q] + prim0l[q]; tr[q] = primr[q] + prim0r[q]; }
  const double csl = cfastX(P, tl), csr = cfastX(P, tr);                 // :61-62
  const double sr = std::max(priml[1] + prim0l[1] + csl, primr[1] + prim0r[1] + csr);
  const double sl = std::min(priml[1] + prim0l[1] - csl, primr[1] + prim0r[1] - csr);
  if (sl > 0) { prim2f_split(P, priml, prim0l, ff); return; }
  if (sr < 0) { prim2f_split(P, primr, prim0r, ff); return; }
  double fL[16], fR[16], uL[16], uR[16];
  prim2f_split(P, priml, prim0l, fL);
  prim2f_split(P, primr, prim0r, fR);
  prim2u_split(P, priml, prim0l, uL);
  prim2u_split(P, primr, prim0r, uR);
  for (int q = 0; q < P.neq; ++q) ff[q] = (sr * fL[q] - sl * fR[q] + sl * sr * (uR[q] - uL[q])) / (sr - sl);   // :83
}
// src/hlle_split_all.f90:97-241  hllEfluxesSplitAll(choice): the sweep of `fluxes` with the background states alongside
// (the limiter is applied to the fluctuations and to the background separately, :154-155)
static int fluxes_split_all(const Par& P, Block& b, int choice) {
  const int neq = P.neq;
  double priml[16], primr[16], primll[16], primrr[16], prim0l[16], prim0r[16], prim0ll[16], prim0rr[16], ff[16];
  for (int q = 0; q < 16; ++q) ff[q] = 0.0;
  auto ld = [&](double* dst, const Arr4& A, int i, int j, int k) { std::memcpy(dst, A.cell(i, j, k), sizeof(double) * neq); };
  const int di[3] = {1, 0, 0}, dj[3] = {0, 1, 0}, dk[3] = {0, 0, 1};
  Arr4* out[3] = {&b.f, &b.g, &b.h};
  for (int k = 0; k <= P.nz; ++k)
    for (int j = 0; j <= P.ny; ++j)
      for (int i = 0; i <= P.nx; ++i)
        for (int d = 0; d < 3; ++d) {
          auto sw = [&](double* v) { if (d == 1) swapy(P, v); if (d == 2) swapz(P, v); };
          ld(priml, b.primit, i, j, k); ld(primr, b.primit, i + di[d], j + dj[d], k + dk[d]);
          ld(prim0l, b.primit0, i, j, k); ld(prim0r, b.primit0, i + di[d], j + dj[d], k + dk[d]);
          sw(priml); sw(primr); sw(prim0l); sw(prim0r);
          if (choice == 2) {
            ld(primll, b.primit, i - di[d], j - dj[d], k - dk[d]); ld(primrr, b.primit, i + 2 * di[d], j + 2 * dj[d], k + 2 * dk[d]);
            ld(prim0ll, b.primit0, i - di[d], j - dj[d], k - dk[d]); ld(prim0rr, b.primit0, i + 2 * di[d], j + 2 * dj[d], k + 2 * dk[d]);
            sw(primll); sw(primrr); sw(prim0ll); sw(prim0rr);
            limiter(P.slope_limiter, primll, priml, primr, primrr, neq);
            limiter(P.slope_limiter, prim0ll, prim0l, prim0r, prim0rr, neq);
          }
          prim2fhlleSplitAll(P, priml, primr, prim0l, prim0r, ff);
          sw(ff);
          std::memcpy(out[d]->cell(i, j, k), ff, sizeof(double) * neq);
        }
  return 0;
}

// ---------------------------------------------------------------------------
// halo machinery shared by boundaryI, boundaryII and boundaryI_ef
// pack a sub-box of A into buf (Fortran order: var fastest, then i, j, k)
static void pack(const Arr4& A, std::vector<double>& buf, int i0, int i1, int j0, int j1, int k0, int k1) {
  size_t n = (size_t)A.n1 * (i1 - i0 + 1) * (j1 - j0 + 1) * (k1 - k0 + 1);
  buf.resize(n);
  size_t p = 0;
  for (int k = k0; k <= k1; ++k) for (int j = j0; j <= j1; ++j) for (int i = i0; i <= i1; ++i) {
    const double* c = A.cell(i, j, k);
    for (int q = 0; q < A.n1; ++q) buf[p++] = c[q];
  }
}
static void unpack(Arr4& A, const std::vector<double>& buf, int i0, int i1, int j0, int j1, int k0, int k1) {
  size_t p = 0;
  for (int k = k0; k <= k1; ++k) for (int j = j0; j <= j1; ++j) for (int i = i0; i <= i1; ++i) {
    double* c = A.cell(i, j, k);
    for (int q = 0; q < A.n1; ++q) c[q] = buf[p++];
  }
}
// A(q, dst) = sgn * A(q, src) over a plane range; dir 0/1/2 = x/y/z index being copied
static void copy_plane(Arr4& A, int dir, int dst, int src, int a0, int a1, int b0, int b1, int q0, int q1, double sgn) {
  for (int bb = b0; bb <= b1; ++bb) for (int aa = a0; aa <= a1; ++aa) {
    double *cd, *cs;
    if (dir == 0) { cd = A.cell(dst, aa, bb); cs = A.cell(src, aa, bb); }
    else if (dir == 1) { cd = A.cell(aa, dst, bb); cs = A.cell(aa, src, bb); }
    else { cd = A.cell(aa, bb, dst); cs = A.cell(aa, bb, src); }
    for (int q = q0; q <= q1; ++q) cd[q - 1] = sgn * cs[q - 1];
  }
}

template <class F> static void for_blocks(Oracle& O, F fn) {
  int nb = (int)O.B.size();
  int nt = std::min(O.nthreads, nb);
  if (nt <= 1) { for (int b = 0; b < nb; ++b) fn(O.B[b]); return; }
  std::vector<std::thread> th;
  std::atomic<int> next(0);
  for (int t = 0; t < nt; ++t) th.emplace_back([&]() { int b; while ((b = next.fetch_add(1)) < nb) fn(O.B[b]); });
  for (auto& t : th) t.join();
}

static void user_bc_dispatch(Oracle& O, Block& b, Arr4& A, int order);

// src/boundaries.f90:44-247 boundaryI  (MPI branch)
static void boundaryI(Oracle& O) {
  const Par& P = O.P;
  const int nx = P.nx, ny = P.ny, nz = P.nz, nxp1 = nx + 1, nyp1 = ny + 1, nzp1 = nz + 1, neq = P.neq;
  for_blocks(O, [&](Block& b) {                                         // :70-75
    pack(b.u, b.sendr, nx, nx, 0, nyp1, 0, nzp1);
    pack(b.u, b.sendl, 1, 1, 0, nyp1, 0, nzp1);
    pack(b.u, b.sendt, 0, nxp1, ny, ny, 0, nzp1);
    pack(b.u, b.sendb, 0, nxp1, 1, 1, 0, nzp1);
    pack(b.u, b.sendi, 0, nxp1, 0, nyp1, nz, nz);
    pack(b.u, b.sendo, 0, nxp1, 0, nyp1, 1, 1);
  });
  for_blocks(O, [&](Block& b) {                                         // :77-106
    if (b.left != -1)   unpack(b.u, O.B[b.left].sendr, 0, 0, 0, nyp1, 0, nzp1);
    if (b.right != -1)  unpack(b.u, O.B[b.right].sendl, nxp1, nxp1, 0, nyp1, 0, nzp1);
    if (b.bottom != -1) unpack(b.u, O.B[b.bottom].sendt, 0, nxp1, 0, 0, 0, nzp1);
    if (b.top != -1)    unpack(b.u, O.B[b.top].sendb, 0, nxp1, nyp1, nyp1, 0, nzp1);
    if (b.out != -1)    unpack(b.u, O.B[b.out].sendi, 0, nxp1, 0, nyp1, 0, 0);
    if (b.in != -1)     unpack(b.u, O.B[b.in].sendo, 0, nxp1, 0, nyp1, nzp1, nzp1);
    // reflecting  :146-199
    if (P.bc_left == GX_BC_CLOSED && b.coords[0] == 0) {
      copy_plane(b.u, 0, 0, 1, 0, nyp1, 0, nzp1, 1, 1, 1.); copy_plane(b.u, 0, 0, 1, 0, nyp1, 0, nzp1, 2, 2, -1.); copy_plane(b.u, 0, 0, 1, 0, nyp1, 0, nzp1, 3, neq, 1.);
    }
    if (P.bc_right == GX_BC_CLOSED && b.coords[0] == P.NBX - 1) {
      copy_plane(b.u, 0, nxp1, nx, 0, nyp1, 0, nzp1, 1, 1, 1.); copy_plane(b.u, 0, nxp1, nx, 0, nyp1, 0, nzp1, 2, 2, -1.); copy_plane(b.u, 0, nxp1, nx, 0, nyp1, 0, nzp1, 3, neq, 1.);
    }
    if (P.bc_bottom == GX_BC_CLOSED && b.coords[1] == 0) {
      copy_plane(b.u, 1, 0, 1, 0, nxp1, 0, nzp1, 1, 2, 1.); copy_plane(b.u, 1, 0, 1, 0, nxp1, 0, nzp1, 3, 3, -1.); copy_plane(b.u, 1, 0, 1, 0, nxp1, 0, nzp1, 4, neq, 1.);
    }
    if (P.bc_top == GX_BC_CLOSED && b.coords[1] == P.NBY - 1) {
      copy_plane(b.u, 1, nyp1, ny, 0, nxp1, 0, nzp1, 1, 2, 1.); copy_plane(b.u, 1, nyp1, ny, 0, nxp1, 0, nzp1, 3, 3, -1.); copy_plane(b.u, 1, nyp1, ny, 0, nxp1, 0, nzp1, 4, neq, 1.);
    }
    if (P.bc_out == GX_BC_CLOSED && b.coords[2] == 0) {
      copy_plane(b.u, 2, 0, 1, 0, nxp1, 0, nyp1, 1, 3, 1.); copy_plane(b.u, 2, 0, 1, 0, nxp1, 0, nyp1, 4, 4, -1.); copy_plane(b.u, 2, 0, 1, 0, nxp1, 0, nyp1, 5, neq, 1.);
    }
    if (P.bc_in == GX_BC_CLOSED && b.coords[2] == P.NBZ - 1) {
      copy_plane(b.u, 2, nzp1, nz, 0, nxp1, 0, nyp1, 1, 3, 1.); copy_plane(b.u, 2, nzp1, nz, 0, nxp1, 0, nyp1, 4, 4, -1.); copy_plane(b.u, 2, nzp1, nz, 0, nxp1, 0, nyp1, 5, neq, 1.);
    }
    // outflow  :201-242
    if (P.bc_left == GX_BC_OUTFLOW && b.coords[0] == 0) copy_plane(b.u, 0, 0, 1, 0, nyp1, 0, nzp1, 1, neq, 1.);
    if (P.bc_right == GX_BC_OUTFLOW && b.coords[0] == P.NBX - 1) copy_plane(b.u, 0, nxp1, nx, 0, nyp1, 0, nzp1, 1, neq, 1.);
    if (P.bc_bottom == GX_BC_OUTFLOW && b.coords[1] == 0) copy_plane(b.u, 1, 0, 1, 0, nxp1, 0, nzp1, 1, neq, 1.);
    if (P.bc_top == GX_BC_OUTFLOW && b.coords[1] == P.NBY - 1) copy_plane(b.u, 1, nyp1, ny, 0, nxp1, 0, nzp1, 1, neq, 1.);
    if (P.bc_out == GX_BC_OUTFLOW && b.coords[2] == 0) copy_plane(b.u, 2, 0, 1, 0, nxp1, 0, nyp1, 1, neq, 1.);
    if (P.bc_in == GX_BC_OUTFLOW && b.coords[2] == P.NBZ - 1) copy_plane(b.u, 2, nzp1, nz, 0, nxp1, 0, nyp1, 1, neq, 1.);
    if (P.bc_user) user_bc_dispatch(O, b, b.u, 1);                      // :245
  });
}

// src/boundaries.f90:256-512 boundaryII  (MPI branch)
static void boundaryII(Oracle& O) {
  const Par& P = O.P;
  const int nx = P.nx, ny = P.ny, nz = P.nz, ng = 2, neq = P.neq;
  const int nxmg = nx - ng + 1, nymg = ny - ng + 1, nzmg = nz - ng + 1, nxp = nx + 1, nyp = ny + 1, nzp = nz + 1;
  for_blocks(O, [&](Block& b) {                                         // :285-290
    pack(b.up, b.sendr, nxmg, nx, P.nymin, P.nymax, P.nzmin, P.nzmax);
    pack(b.up, b.sendl, 1, ng, P.nymin, P.nymax, P.nzmin, P.nzmax);
    pack(b.up, b.sendt, P.nxmin, P.nxmax, nymg, ny, P.nzmin, P.nzmax);
    pack(b.up, b.sendb, P.nxmin, P.nxmax, 1, ng, P.nzmin, P.nzmax);
    pack(b.up, b.sendi, P.nxmin, P.nxmax, P.nymin, P.nymax, nzmg, nz);
    pack(b.up, b.sendo, P.nxmin, P.nxmax, P.nymin, P.nymax, 1, ng);
  });
  for_blocks(O, [&](Block& b) {                                         // :316-321
    if (b.left != -1)   unpack(b.up, O.B[b.left].sendr, P.nxmin, 0, P.nymin, P.nymax, P.nzmin, P.nzmax);
    if (b.right != -1)  unpack(b.up, O.B[b.right].sendl, nxp, P.nxmax, P.nymin, P.nymax, P.nzmin, P.nzmax);
    if (b.bottom != -1) unpack(b.up, O.B[b.bottom].sendt, P.nxmin, P.nxmax, P.nymin, 0, P.nzmin, P.nzmax);
    if (b.top != -1)    unpack(b.up, O.B[b.top].sendb, P.nxmin, P.nxmax, nyp, P.nymax, P.nzmin, P.nzmax);
    if (b.out != -1)    unpack(b.up, O.B[b.out].sendi, P.nxmin, P.nxmax, P.nymin, P.nymax, P.nzmin, 0);
    if (b.in != -1)     unpack(b.up, O.B[b.in].sendo, P.nxmin, P.nxmax, P.nymin, P.nymax, nzp, P.nzmax);
    // mirror-type closed (:361-438) and outflow (:440-505); full transverse ranges (`:`)
    auto mirror = [&](int dir, bool low, int nneg /*1-based var to negate, 0 = none*/) {
      int n = dir == 0 ? nx : dir == 1 ? ny : nz;
      int a0 = dir == 0 ? P.nymin : P.nxmin, a1 = dir == 0 ? P.nymax : P.nxmax;
      int b0 = dir == 2 ? P.nymin : P.nzmin, b1 = dir == 2 ? P.nymax : P.nzmax;
      for (int l = 0; l < ng; ++l) {
        int dst = low ? (1 - ng + l) : (n + 1 + l);
        int src = low ? (ng - l) : (n - l);
        if (nneg == 0) copy_plane(b.up, dir, dst, src, a0, a1, b0, b1, 1, neq, 1.);
        else {
          if (nneg > 1) copy_plane(b.up, dir, dst, src, a0, a1, b0, b1, 1, nneg - 1, 1.);
          copy_plane(b.up, dir, dst, src, a0, a1, b0, b1, nneg, nneg, -1.);
          copy_plane(b.up, dir, dst, src, a0, a1, b0, b1, nneg + 1, neq, 1.);
        }
      }
    };
    if (P.bc_left == GX_BC_CLOSED && b.coords[0] == 0) mirror(0, true, 2);
    if (P.bc_right == GX_BC_CLOSED && b.coords[0] == P.NBX - 1) mirror(0, false, 2);
    if (P.bc_bottom == GX_BC_CLOSED && b.coords[1] == 0) mirror(1, true, 3);
    if (P.bc_top == GX_BC_CLOSED && b.coords[1] == P.NBY - 1) mirror(1, false, 3);
    if (P.bc_out == GX_BC_CLOSED && b.coords[2] == 0) mirror(2, true, 4);
    if (P.bc_in == GX_BC_CLOSED && b.coords[2] == P.NBZ - 1) mirror(2, false, 4);
    if (P.bc_left == GX_BC_OUTFLOW && b.coords[0] == 0) mirror(0, true, 0);
    if (P.bc_right == GX_BC_OUTFLOW && b.coords[0] == P.NBX - 1) mirror(0, false, 0);
    if (P.bc_bottom == GX_BC_OUTFLOW && b.coords[1] == 0) mirror(1, true, 0);
    if (P.bc_top == GX_BC_OUTFLOW && b.coords[1] == P.NBY - 1) mirror(1, false, 0);
    if (P.bc_out == GX_BC_OUTFLOW && b.coords[2] == 0) mirror(2, true, 0);
    if (P.bc_in == GX_BC_OUTFLOW && b.coords[2] == P.NBZ - 1) mirror(2, false, 0);
    if (P.bc_user) user_bc_dispatch(O, b, b.up, 2);                     // :508
  });
}

// src/flux_cd_module.f90:45-237 boundaryI_ef  (MPI branch; the serial branch is broken)
static void boundaryI_ef(Oracle& O) {
  const Par& P = O.P;
  const int nx = P.nx, ny = P.ny, nz = P.nz, nxp1 = nx + 1, nyp1 = ny + 1, nzp1 = nz + 1;
  for_blocks(O, [&](Block& b) {                                         // :68-73
    pack(b.e, b.sendr, nx, nx, 0, nyp1, 0, nzp1);
    pack(b.e, b.sendl, 1, 1, 0, nyp1, 0, nzp1);
    pack(b.e, b.sendt, 0, nxp1, ny, ny, 0, nzp1);
    pack(b.e, b.sendb, 0, nxp1, 1, 1, 0, nzp1);
    pack(b.e, b.sendi, 0, nxp1, 0, nyp1, nz, nz);
    pack(b.e, b.sendo, 0, nxp1, 0, nyp1, 1, 1);
  });
  for_blocks(O, [&](Block& b) {                                         // :99-104
    if (b.left != -1)   unpack(b.e, O.B[b.left].sendr, 0, 0, 0, nyp1, 0, nzp1);
    if (b.right != -1)  unpack(b.e, O.B[b.right].sendl, nxp1, nxp1, 0, nyp1, 0, nzp1);
    if (b.bottom != -1) unpack(b.e, O.B[b.bottom].sendt, 0, nxp1, 0, 0, 0, nzp1);
    if (b.top != -1)    unpack(b.e, O.B[b.top].sendb, 0, nxp1, nyp1, nyp1, 0, nzp1);
    if (b.out != -1)    unpack(b.e, O.B[b.out].sendi, 0, nxp1, 0, nyp1, 0, 0);
    if (b.in != -1)     unpack(b.e, O.B[b.in].sendo, 0, nxp1, 0, nyp1, nzp1, nzp1);
    // reflecting ("not tested" in the reference)  :143-192
    if (P.bc_left == GX_BC_CLOSED && b.coords[0] == 0) { copy_plane(b.e, 0, 0, 1, 0, nyp1, 0, nzp1, 1, 1, -1.); copy_plane(b.e, 0, 0, 1, 0, nyp1, 0, nzp1, 2, 3, 1.); }
    if (P.bc_right == GX_BC_CLOSED && b.coords[0] == P.NBX - 1) { copy_plane(b.e, 0, nxp1, nx, 0, nyp1, 0, nzp1, 1, 1, -1.); copy_plane(b.e, 0, nxp1, nx, 0, nyp1, 0, nzp1, 2, 3, 1.); }
    if (P.bc_bottom == GX_BC_CLOSED && b.coords[1] == 0) { copy_plane(b.e, 1, 0, 1, 0, nxp1, 0, nzp1, 1, 1, 1.); copy_plane(b.e, 1, 0, 1, 0, nxp1, 0, nzp1, 2, 2, -1.); copy_plane(b.e, 1, 0, 1, 0, nxp1, 0, nzp1, 3, 3, 1.); }
    if (P.bc_top == GX_BC_CLOSED && b.coords[1] == P.NBY - 1) { copy_plane(b.e, 1, nyp1, ny, 0, nxp1, 0, nzp1, 1, 1, 1.); copy_plane(b.e, 1, nyp1, ny, 0, nxp1, 0, nzp1, 2, 2, -1.); copy_plane(b.e, 1, nyp1, ny, 0, nxp1, 0, nzp1, 3, 3, 1.); }
    if (P.bc_out == GX_BC_CLOSED && b.coords[2] == 0) copy_plane(b.e, 2, 0, 1, 0, nxp1, 0, nyp1, 1, 3, 1.);      // no sign flip (:179-184)
    if (P.bc_in == GX_BC_CLOSED && b.coords[2] == P.NBZ - 1) copy_plane(b.e, 2, nzp1, nz, 0, nxp1, 0, nyp1, 1, 3, 1.);
    // outflow  :194-235
    if (P.bc_left == GX_BC_OUTFLOW && b.coords[0] == 0) copy_plane(b.e, 0, 0, 1, 0, nyp1, 0, nzp1, 1, 3, 1.);
    if (P.bc_right == GX_BC_OUTFLOW && b.coords[0] == P.NBX - 1) copy_plane(b.e, 0, nxp1, nx, 0, nyp1, 0, nzp1, 1, 3, 1.);
    if (P.bc_bottom == GX_BC_OUTFLOW && b.coords[1] == 0) copy_plane(b.e, 1, 0, 1, 0, nxp1, 0, nzp1, 1, 3, 1.);
    if (P.bc_top == GX_BC_OUTFLOW && b.coords[1] == P.NBY - 1) copy_plane(b.e, 1, nyp1, ny, 0, nxp1, 0, nzp1, 1, 3, 1.);
    if (P.bc_out == GX_BC_OUTFLOW && b.coords[2] == 0) copy_plane(b.e, 2, 0, 1, 0, nxp1, 0, nyp1, 1, 3, 1.);
    if (P.bc_in == GX_BC_OUTFLOW && b.coords[2] == P.NBZ - 1) copy_plane(b.e, 2, nzp1, nz, 0, nxp1, 0, nyp1, 1, 3, 1.);
  });
}

// src/flux_cd_module.f90:245-273 get_efield
static void get_efield(Oracle& O) {
  const Par& P = O.P;
  for_blocks(O, [&](Block& b) {
    for (int k = 1; k <= P.nz; ++k) for (int j = 1; j <= P.ny; ++j) for (int i = 1; i <= P.nx; ++i) {
      b.e(1, i, j, k) = 0.25 * (-b.g(8, i, j - 1, k) - b.g(8, i, j, k) + b.h(7, i, j, k - 1) + b.h(7, i, j, k));
      b.e(2, i, j, k) = 0.25 * (+b.f(8, i - 1, j, k) + b.f(8, i, j, k) - b.h(6, i, j, k - 1) - b.h(6, i, j, k));
      b.e(3, i, j, k) = 0.25 * (-b.f(7, i - 1, j, k) - b.f(7, i, j, k) + b.g(6, i, j - 1, k) + b.g(6, i, j, k));
    }
  });
  boundaryI_ef(O);                                                      // :271
}

// src/flux_cd_module.f90:285-323 flux_cd_update
static void flux_cd_update(const Par& P, Block& b, int i, int j, int k, double dt) {
  double dtdx = dt / P.dx, dtdy = dt / P.dy, dtdz = dt / P.dz;
  auto upw = [&](int q) {
    b.up(q, i, j, k) = b.u(q, i, j, k) - dtdx * (b.f(q, i, j, k) - b.f(q, i - 1, j, k))
                                       - dtdy * (b.g(q, i, j, k) - b.g(q, i, j - 1, k))
                                       - dtdz * (b.h(q, i, j, k) - b.h(q, i, j, k - 1));
  };
  for (int q = 1; q <= 5; ++q) upw(q);                                  // :299-301
  if (P.passives) for (int q = P.neqdyn + 1; q <= P.neq; ++q) upw(q);   // :304-308
  b.up(6, i, j, k) = b.u(6, i, j, k) - 0.5 * dtdy * (b.e(3, i, j + 1, k) - b.e(3, i, j - 1, k))
                                     + 0.5 * dtdz * (b.e(2, i, j, k + 1) - b.e(2, i, j, k - 1));   // :311-313
  b.up(7, i, j, k) = b.u(7, i, j, k) + 0.5 * dtdx * (b.e(3, i + 1, j, k) - b.e(3, i - 1, j, k))
                                     - 0.5 * dtdz * (b.e(1, i, j, k + 1) - b.e(1, i, j, k - 1));   // :315-317
  b.up(8, i, j, k) = b.u(8, i, j, k) - 0.5 * dtdx * (b.e(2, i + 1, j, k) - b.e(2, i - 1, j, k))
                                     + 0.5 * dtdy * (b.e(1, i, j + 1, k) - b.e(1, i, j - 1, k));   // :319-321
}

// EXO/user_mod.f90:158-206 get_user_source_terms (built-in problem module)
static void exo_user_source(const Oracle& O, const Block& b, const double* pp, double* s, int i, int j, int k) {
  const Par& P = O.P;
  const double Ggrav = 6.67259e-8;
  const int nb = 2;
  double x[2], y[2], z[2], GM[2], rad2[2];
  GM[0] = 0.3 * Ggrav * O.MassS / O.exo_rsc / O.exo_vsc2;
  GM[1] = Ggrav * O.MassP / O.exo_rsc / O.exo_vsc2;
  double xc = ((double)(i + b.coords[0] * P.nx - P.nxtot / 2) - 0.5) * P.dx;
  double yc = ((double)(j + b.coords[1] * P.ny - P.nytot / 2) - 0.5) * P.dy;
  double zc = ((double)(k + b.coords[2] * P.nz - P.nztot / 2) - 0.5) * P.dz;
  x[0] = xc; y[0] = yc; z[0] = zc;
  rad2[0] = x[0] * x[0] + y[0] * y[0] + z[0] * z[0];
  x[1] = xc - O.xp; y[1] = yc; z[1] = zc - O.zp;
  rad2[1] = x[1] * x[1] + y[1] * y[1] + z[1] * z[1];
  for (int l = 0; l < nb; ++l) {
    double r15 = std::pow(rad2[l], 1.5);
    s[1] = s[1] - pp[0] * GM[l] * x[l] / r15;
    s[2] = s[2] - pp[0] * GM[l] * y[l] / r15;
    s[3] = s[3] - pp[0] * GM[l] * z[l] / r15;
    s[4] = s[4] - pp[0] * GM[l] * (pp[1] * x[l] + pp[2] * y[l] + pp[3] * z[l]) / r15;
  }
}

// src/sources.f90:124-134, 151-175, 190-220  source (+ divergence_B, divbcorr_8w_source)
static void source(const Oracle& O, const Block& b, int i, int j, int k, const double* prim, double* s) {
  const Par& P = O.P;
  for (int q = 0; q < P.neq; ++q) s[q] = 0.;
  if (P.user_source_terms) {
    if (O.builtin_user == 1) exo_user_source(O, b, prim, s, i, j, k);
    else if (O.user_src) O.user_src(prim, s, i, j, k, b.coords, O.user_src_ctx);
  }
  // radiation_pressure needs difrad::ph — out of scope
  if (P.eight_wave && (P.mhd || P.pmhd)) {
    double d = (b.primit(6, i + 1, j, k) - b.primit(6, i - 1, j, k)) / (2. * P.dx)
             + (b.primit(7, i, j + 1, k) - b.primit(7, i, j - 1, k)) / (2. * P.dy)
             + (b.primit(8, i, j, k + 1) - b.primit(8, i, j, k - 1)) / (2. * P.dz);
    s[1] = s[1] - d * prim[5];
    s[2] = s[2] - d * prim[6];
    s[3] = s[3] - d * prim[7];
    s[4] = s[4] - d * (prim[1] * prim[5] + prim[2] * prim[6] + prim[3] * prim[7]);
    s[5] = s[5] - d * prim[1];
    s[6] = s[6] - d * prim[2];
    s[7] = s[7] - d * prim[3];
  }
}

// src/hydro_solver.f90:77-127 step
static void step(Oracle& O, double dt) {
  const Par& P = O.P;
  const double dtdx = dt / P.dx, dtdy = dt / P.dy, dtdz = dt / P.dz;
  const bool bfield = (P.neqdyn == 8);
  if (bfield && P.enable_flux_cd) get_efield(O);                        // :95-97
  for_blocks(O, [&](Block& b) {
    double s[16];
    for (int k = 1; k <= P.nz; ++k) for (int j = 1; j <= P.ny; ++j) for (int i = 1; i <= P.nx; ++i) {
      if (!P.enable_flux_cd) {                                          // :103-107
        for (int q = 1; q <= P.neq; ++q)
          b.up(q, i, j, k) = b.u(q, i, j, k) - dtdx * (b.f(q, i, j, k) - b.f(q, i - 1, j, k))
                                             - dtdy * (b.g(q, i, j, k) - b.g(q, i, j - 1, k))
                                             - dtdz * (b.h(q, i, j, k) - b.h(q, i, j, k - 1));
      } else if (bfield) {
        flux_cd_update(P, b, i, j, k, dt);                              // :110-112 (nothing happens without BFIELD)
      }
      if (P.user_source_terms || P.eight_wave) {                        // :115-121
        source(O, b, i, j, k, b.primit.cell(i, j, k), s);
        for (int q = 1; q <= P.neq; ++q) b.up(q, i, j, k) = b.up(q, i, j, k) + dt * s[q - 1];
      }
    }
  });
}

// src/hydro_solver.f90:47-65 viscous_copy
static void viscous_copy(Oracle& O) {
  const Par& P = O.P;
  for_blocks(O, [&](Block& b) {
    for (int k = 1; k <= P.nz; ++k) for (int j = 1; j <= P.ny; ++j) for (int i = 1; i <= P.nx; ++i)
      for (int q = 1; q <= P.neq; ++q)
        b.u(q, i, j, k) = b.up(q, i, j, k) + P.eta * (b.up(q, i + 1, j, k) + b.up(q, i - 1, j, k)
                                                    + b.up(q, i, j + 1, k) + b.up(q, i, j - 1, k)
                                                    + b.up(q, i, j, k + 1) + b.up(q, i, j, k - 1)
                                                    - 6. * b.up(q, i, j, k));
  });
}

// ---------------------------------------------------------------------------
// src/cooling_h.f90 — parametrised hydrogen cooling (COOL_H), a cell-local operator applied to u after
// viscous_copy (src/hydro_solver.f90:202-204).  Arithmetic note: the reference is built with
// -fdefault-real-8, under which gfortran promotes the `d`-exponent literals below (2.55d-13, 1.d4, 5.83d-11,
// 1.133D-24, xi, boltzm) to 16-byte reals, so those sub-expressions are evaluated in quad precision and then
// rounded to real(8).  This restatement evaluates them in FP64: the results agree with the correctly rounded
// values to an ulp or two, far inside the parity tolerance stated in tests/test_exo_gpu.py.
static double cool_alpha(double T) { return 2.55e-13 * std::pow(1.e4 / T, 0.79); }                 // cooling_h.f90:75-84
static double cool_colf(double T) { return 5.83e-11 * std::sqrt(T) * std::exp(-157828. / T); }      // :109-118
static double cool_betah(double T) {                                                              // :126-137
  double a = 157890. / T;
  return 1.133e-24 / std::sqrt(a) * (-0.0713 + 0.5 * std::log(a) + 0.640 * std::pow(a, -0.33333));
}
// cooling_h.f90:159-246 ALOSS(X1,X2,DT,DEN,DH0,TE0)
static double cool_aloss(double X1, double X2, double DT, double DEN, double DH0, double TE0) {
  const double XION = 2.179e-11, XO = 1.e-3;
  const double C0 = 0.5732, C1 = 1.8288e-5, C2 = -1.15822e-10, C3 = 9.4288e-16;
  const double D0 = 0.5856, D1 = 1.55083e-5, D2 = -9.669e-12, D3 = 5.716e-19;
  const double ENK = 118409., EN = 1.634E-11;
  double TE = std::max(TE0, 10.);
  double DH = DEN;
  double DHP = (1. - X1) * DH;
  double DE = DHP + 1.E-4 * DH;
  double DOI = XO * DH0;
  double DOII = XO * DHP;
  (void)X2; (void)DT;                                  // SHP is computed but never used (:183)
  if (TE <= 1e4) return 0.;
  double OMEGA = 0.;
  if (TE <= 55000.) OMEGA = C0 + TE * (C1 + TE * (C2 + TE * C3));
  if (TE >= 72000.) OMEGA = D0 + TE * (D1 + TE * (D2 + TE * D3));
  if (TE > 55000. && TE < 72000.) {
    double OMEGAL = C0 + TE * (C1 + TE * (C2 + TE * C3));
    double OMEGAH = D0 + TE * (D1 + TE * (D2 + TE * D3));
    double FRAC = (TE - 55000.) / 17000.;
    OMEGA = (1. - FRAC) * OMEGAL + FRAC * OMEGAH;
  }
  double QLA = 8.6287E-6 / (2. * std::sqrt(TE)) * OMEGA * std::exp(-ENK / TE);
  double ECOLL = DE * DH0 * QLA * EN;
  ECOLL = std::max(ECOLL, 0.);
  double CION = 5.834E-11 * std::sqrt(TE) * std::exp(-1.579E5 / TE);
  double EION = DE * DH0 * CION * XION;
  double EREC = DE * DHP * (cool_betah(TE));
  EREC = std::max(EREC, 0.);
  double TM = 1. / TE;
  double T2 = TM * TM;
  double EOI = DE * DOI * std::pow(10., 1381465 * T2 - 12328.69 * TM - 19.82621);
  double EOII = DE * DOII * std::pow(10., -2061075. * T2 - 14596.24 * TM - 19.01402);
  EOI = std::max(EOI, 0.);
  EOII = std::max(EOII, 0.);
  double BETAF = 1.3 * 1.42E-27 * std::pow(TE, 0.5);
  double HIICOOL = DE * DHP * BETAF;
  double EQUIL = (1.0455E-18 / std::pow(TE, 0.63)) * (1. - std::exp(-std::pow(TE * 1.E-5, 1.63))) * DE * DEN + HIICOOL;
  double FR = 0.;
  if (TE <= 44770.) FR = 0.;
  if (TE >= 54770.) FR = 1.;
  if (TE > 44770. && TE < 54770.) {
    double EX2 = std::exp(-2. * (TE - 49770.) / 500.);
    double TANH = (1. - EX2) / (1. + EX2);
    FR = 0.5 * (1. + TANH);
  }
  return ECOLL + EION + (EREC + 7.033 * (EOI + EOII)) * (1. - FR) + EQUIL * FR;
}
// cooling_h.f90:259-371 atomic(dt,uu,tau,radphi)  (dif_rad = .false. branch: radphi unused)
static void cool_atomic(const Par& P, double dt, double* uu) {
  const double xi = 1.e-4, boltzm = 1.3807e-16;
  double prim[16], T;
  u2prim(P, uu, prim, T);
  double col = cool_colf(T);
  double rec = cool_alpha(T);
  double y0 = uu[P.neqdyn] / uu[0];
  double dh = uu[0];
  double a = rec + col;
  double b = -((2. + xi) * rec + (1. + xi) * col);
  double c = (1. + xi) * rec;
  double d = std::sqrt(b * b - 4. * a * c);
  double g0 = (2. * a * y0 + b + d) / (2. * a * y0 + b - d);
  double e = std::exp(-d * dh * dt);
  double y1 = (-b - d * (1. + g0 * e) / (1. - g0 * e)) / (2. * a);
  y1 = std::min(y1, 0.9999);
  y1 = std::max(y1, 0.);
  double dh0 = uu[P.neqdyn];
  double al = cool_aloss(y0, y1, dt, dh, dh0, T) / (dh * dh);
  double tprime = 10.;
  double ce = (2. * dh * al) / (3. * boltzm * T);
  double t1 = tprime + (T - tprime) * std::exp(-ce * dt);
  t1 = std::max(t1, 0.1 * T);
  t1 = std::min(t1, 10. * T);
  uu[P.neqdyn] = y1 * uu[0];
  if (P.mhd) uu[4] = P.cv * (2. * uu[0] - uu[P.neqdyn]) * t1 / P.Tempsc + 0.5 * prim[0] * (prim[1] * prim[1] + prim[2] * prim[2] + prim[3] * prim[3])
                     + 0.5 * (prim[5] * prim[5] + prim[6] * prim[6] + prim[7] * prim[7]);
  else uu[4] = P.cv * (2. * uu[0] - uu[P.neqdyn]) * t1 / P.Tempsc + 0.5 * prim[0] * (prim[1] * prim[1] + prim[2] * prim[2] + prim[3] * prim[3]);
}
// cooling_h.f90:41-67 coolingh
static void coolingh(Oracle& O, double dt_CFL) {
  const Par& P = O.P;
  const double dt_seconds = dt_CFL * P.tsc;
  for_blocks(O, [&](Block& b) {
    for (int k = 1; k <= P.nz; ++k) for (int j = 1; j <= P.ny; ++j) for (int i = 1; i <= P.nx; ++i)
      cool_atomic(P, dt_seconds, b.u.cell(i, j, k));
  });
}

// ---------------------------------------------------------------------------
// src/thermal_cond.f90 — thermal conduction, operator-split at the end of tstep (hydro_solver.f90:227).
// Arithmetic notes: `T**(2.5)` is a call to pow(); integer powers `x**n` are gfortran's __builtin_powi
// (libgcc __powidf2: binary exponentiation, restated in tc_powi); `**2`/`**3` expand to products.
// Heat fluxes live in the 5th component of the global f, g, h arrays, in cgs.
static const double TC_ph = 0.4, TC_nu = 0.01, TC_tstep_red_factor = 0.25;      // thermal_cond.f90:35-40
static const double TC_Rg = 8.3145e7;                                            // constants.f90:35
static const double TC_clight = 3.E10;                                           // local parameter, thermal_cond.f90:193,282
static double tc_powi(double x, int m) {
  unsigned n = m < 0 ? (unsigned)(-m) : (unsigned)m;
  double y = (n % 2) ? x : 1.0;
  while (n >>= 1) { x = x * x; if (n % 2) y *= x; }
  return m < 0 ? 1.0 / y : y;
}
static double tc_Ksp(double T) { return 6.e-7 * std::pow(T, 2.5); }                                   // :142-149
static double tc_Ksp_parl(double T) { return 9.2181e-7 * std::pow(T, 2.5); }                          // :157-164
static double tc_Ksp_perp(double T, double dens, double B2) { return 0.30089e+33 * dens / (B2 * std::sqrt(T)) * dens; }   // :172-178
static inline size_t tc_tix(const Arr4& u, int i, int j, int k) { return (size_t)(i + 1) + (size_t)u.NX * ((size_t)(j + 1) + (size_t)u.NY * (size_t)(k + 1)); }

// thermal_cond.f90:78-110 get_dt_cond
static double tc_get_dt_cond(Oracle& O) {
  const Par& P = O.P;
  std::vector<double> dts(O.B.size(), 0.0);
  double ddx = std::min(P.dx, P.dy);
  ddx = std::min(ddx, P.dz);
  for_blocks(O, [&](Block& b) {
    double dtp = 1.7976931348623157e308;                                // huge(1.)
    for (int k = 1; k <= P.nz; ++k) for (int j = 1; j <= P.ny; ++j) for (int i = 1; i <= P.nx; ++i)
      dtp = std::min(dtp, b.primit(1, i, j, k) / tc_Ksp(b.Temp[tc_tix(b.u, i, j, k)]));     // :95
    dtp = TC_tstep_red_factor * 0.5 * ((ddx * P.rsc) * (ddx * P.rsc)) * P.cv * TC_Rg * dtp * P.rhosc / P.mu;   // :101
    dts[b.rank] = dtp;
  });
  double dt = dts[0];
  for (double v : dts) dt = std::min(dt, v);                             // :104 mpi_allreduce(MIN)
  return dt;
}

// thermal_cond.f90:189-267 heatfluxes
static void tc_heatfluxes(const Par& P, Block& b) {
  const double yhp = 1.;                                                  // :200
  auto T = [&](int i, int j, int k) { return b.Temp[tc_tix(b.u, i, j, k)]; };
  auto one = [&](int i, int j, int k, int i2, int j2, int k2, double dxx) -> double {
    if (T(i, j, k) == T(i2, j2, k2)) return 0.;
    const double meanP = 0.5 * (b.primit(5, i, j, k) + b.primit(5, i2, j2, k2));
    const double meanDens = 0.5 * (b.primit(1, i, j, k) + b.primit(1, i2, j2, k2));
    const double meanT = 0.5 * (T(i, j, k) + T(i2, j2, k2));
    const double dT = (T(i2, j2, k2) - T(i, j, k)) / (dxx * P.rsc);
    double coef;
    if (P.tc_saturation) {
      double cs = csound(P, meanP, meanDens);
      cs = std::min(cs * std::sqrt(P.vsc2), TC_clight);
      coef = std::min(tc_Ksp(meanT), 5. * TC_ph * cs * meanP * P.Psc / std::fabs(dT));
    } else coef = tc_Ksp(meanT);
    return -coef * dT * yhp;
  };
  for (int k = 0; k <= P.nz; ++k) for (int j = 0; j <= P.ny; ++j) for (int i = 0; i <= P.nx; ++i) {
    b.f(5, i, j, k) = one(i, j, k, i + 1, j, k, P.dx);
    b.g(5, i, j, k) = one(i, j, k, i, j + 1, k, P.dy);
    b.h(5, i, j, k) = one(i, j, k, i, j, k + 1, P.dz);
  }
}

// thermal_cond.f90:277-487 MHD_heatfluxes
static void tc_mhd_heatfluxes(const Par& P, Block& b) {
  const double phi = 0.3, alpha = 5.0 * phi;
  auto T = [&](int i, int j, int k) { return b.Temp[tc_tix(b.u, i, j, k)]; };
  for (size_t n = 4; n < b.f.d.size(); n += (size_t)b.f.n1) { b.f.d[n] = 0.0; b.g.d[n] = 0.0; b.h.d[n] = 0.0; }     // :298
  for (int k = 0; k <= P.nz; ++k) for (int j = 0; j <= P.ny; ++j) for (int i = 0; i <= P.nx; ++i) {
    double bx = b.primit(6, i, j, k), by = b.primit(7, i, j, k), bz = b.primit(8, i, j, k);
    const double B2 = bx * bx + by * by + bz * bz;
    const double modB = std::sqrt(B2);
    bx = bx / modB; by = by / modB; bz = bz / modB;
    double grad[3], Kparl[3], Kperp[3], coefSat[3];
    const int nb[3][3] = {{i + 1, j, k}, {i, j + 1, k}, {i, j, k + 1}};
    const double dd[3] = {P.dx, P.dy, P.dz};
    for (int d = 0; d < 3; ++d) {
      const int i2 = nb[d][0], j2 = nb[d][1], k2 = nb[d][2];
      if (std::fabs(T(i, j, k) - T(i2, j2, k2)) < 1.0e-14) { grad[d] = 0.0; Kparl[d] = 0.0; Kperp[d] = 0.0; coefSat[d] = 0.0; continue; }
      const double meanDens = 0.5 * (b.primit(1, i, j, k) + b.primit(1, i2, j2, k2));
      const double meanTemp = 0.5 * (T(i, j, k) + T(i2, j2, k2));
      if (P.tc_saturation) {                                               // :398-403
        const double meanPres = 0.5 * (b.primit(5, i, j, k) + b.primit(5, i2, j2, k2));
        double cs = csound(P, meanPres, meanDens);
        cs = std::min(cs * P.vsc, TC_clight);
        coefSat[d] = alpha * meanDens * (cs * cs * cs);
      } else coefSat[d] = 0.0;
      grad[d] = (T(i2, j2, k2) - T(i, j, k)) / (dd[d] * P.rsc);
      Kparl[d] = tc_Ksp_parl(meanTemp);
      Kperp[d] = tc_Ksp_perp(meanTemp, meanDens * P.rhosc, B2 * (P.bsc * P.bsc));
    }
    const double bgradT = bx * grad[0] + by * grad[1] + bz * grad[2];     // :370, :458
    const double parl[3] = {bgradT * bx, bgradT * by, bgradT * bz};
    const double perp[3] = {grad[0] - parl[0], grad[1] - parl[1], grad[2] - parl[2]};
    double fl[3];
    if (!P.tc_saturation) {
      for (int d = 0; d < 3; ++d) fl[d] = -Kparl[d] * parl[d] - Kperp[d] * perp[d];        // :380-384
    } else {
      const double gradT_parl = bgradT;                                    // :464 (signed, as in the reference)
      const double gradT_perp = std::sqrt(perp[0] * perp[0] + perp[1] * perp[1] + perp[2] * perp[2]);
      for (int d = 0; d < 3; ++d)                                          // :472-479
        fl[d] = -1. / (1. / (Kparl[d] + 1.e-14) + gradT_parl / (coefSat[d] + 1.e-14)) * parl[d]
                - 1. / (1. / (Kperp[d] + 1.e-14) + gradT_perp / (coefSat[d] + 1.e-14)) * perp[d];
    }
    b.f(5, i, j, k) = fl[0]; b.g(5, i, j, k) = fl[1]; b.h(5, i, j, k) = fl[2];
  }
}

// thermal_cond.f90:496-616 thermal_bounds (MPI branch): one layer of u(5) between blocks, all six faces packed before any is
// written; then zero-gradient copies on EVERY face of the domain, whatever its boundary type (:589-614) — with periodic
// boundaries they overwrite what the periodic neighbour sent.
static void tc_thermal_bounds(Oracle& O) {
  const Par& P = O.P;
  const int nx = P.nx, ny = P.ny, nz = P.nz, nxp1 = nx + 1, nyp1 = ny + 1, nzp1 = nz + 1;
  auto pack5 = [](const Arr4& A, std::vector<double>& buf, int i0, int i1, int j0, int j1, int k0, int k1) {
    buf.resize((size_t)(i1 - i0 + 1) * (j1 - j0 + 1) * (k1 - k0 + 1));
    size_t p = 0;
    for (int k = k0; k <= k1; ++k) for (int j = j0; j <= j1; ++j) for (int i = i0; i <= i1; ++i) buf[p++] = A(5, i, j, k);
  };
  auto unpack5 = [](Arr4& A, const std::vector<double>& buf, int i0, int i1, int j0, int j1, int k0, int k1) {
    size_t p = 0;
    for (int k = k0; k <= k1; ++k) for (int j = j0; j <= j1; ++j) for (int i = i0; i <= i1; ++i) A(5, i, j, k) = buf[p++];
  };
  for_blocks(O, [&](Block& b) {                                           // :514-519
    pack5(b.u, b.sendr, nx, nx, 0, nyp1, 0, nzp1);
    pack5(b.u, b.sendl, 1, 1, 0, nyp1, 0, nzp1);
    pack5(b.u, b.sendt, 0, nxp1, ny, ny, 0, nzp1);
    pack5(b.u, b.sendb, 0, nxp1, 1, 1, 0, nzp1);
    pack5(b.u, b.sendi, 0, nxp1, 0, nyp1, nz, nz);
    pack5(b.u, b.sendo, 0, nxp1, 0, nyp1, 1, 1);
  });
  for_blocks(O, [&](Block& b) {                                           // :545-550
    if (b.left != -1)   unpack5(b.u, O.B[b.left].sendr, 0, 0, 0, nyp1, 0, nzp1);
    if (b.right != -1)  unpack5(b.u, O.B[b.right].sendl, nxp1, nxp1, 0, nyp1, 0, nzp1);
    if (b.bottom != -1) unpack5(b.u, O.B[b.bottom].sendt, 0, nxp1, 0, 0, 0, nzp1);
    if (b.top != -1)    unpack5(b.u, O.B[b.top].sendb, 0, nxp1, nyp1, nyp1, 0, nzp1);
    if (b.out != -1)    unpack5(b.u, O.B[b.out].sendi, 0, nxp1, 0, nyp1, 0, 0);
    if (b.in != -1)     unpack5(b.u, O.B[b.in].sendo, 0, nxp1, 0, nyp1, nzp1, nzp1);
    if (b.coords[0] == 0)         copy_plane(b.u, 0, 0, 1, 0, nyp1, 0, nzp1, 5, 5, 1.);        // :592-614
    if (b.coords[0] == P.NBX - 1) copy_plane(b.u, 0, nxp1, nx, 0, nyp1, 0, nzp1, 5, 5, 1.);
    if (b.coords[1] == 0)         copy_plane(b.u, 1, 0, 1, 0, nxp1, 0, nzp1, 5, 5, 1.);
    if (b.coords[1] == P.NBY - 1) copy_plane(b.u, 1, nyp1, ny, 0, nxp1, 0, nzp1, 5, 5, 1.);
    if (b.coords[2] == 0)         copy_plane(b.u, 2, 0, 1, 0, nxp1, 0, nyp1, 5, 5, 1.);
    if (b.coords[2] == P.NBZ - 1) copy_plane(b.u, 2, nzp1, nz, 0, nxp1, 0, nyp1, 5, 5, 1.);
  });
}

// thermal_cond.f90:625-636 superstep, :646-654 substep, :664-681 ST_steps
static double tc_superstep(int N, double snu) {
  return (double)N / (2. * snu) * (tc_powi(1 + snu, 2 * N) - tc_powi(1 - snu, 2 * N)) / (tc_powi(1 + snu, 2 * N) + tc_powi(1 - snu, 2 * N));
}
static double tc_substep(int j, int N, double nu) {
  const double pi = std::acos(-1.);
  return 1. / ((nu - 1.) * std::cos(pi * (double)(2 * j - 1) / (2. * (double)N)) + nu + 1.);
}
static void tc_ST_steps(double fs, int& Ns, double& fstep) {
  const double snu = std::sqrt(TC_nu);
  int j;
  for (j = 1; j <= 199; ++j) if (tc_superstep(j, snu) > fs) break;       // a completed Fortran do loop leaves j = jmax + 1
  Ns = j;
  fstep = fs / tc_superstep(Ns, snu);
}

// thermal_cond.f90:690-768 thermal_conduction
static void thermal_conduction(Oracle& O, double dt_CFL) {
  const Par& P = O.P;
  const double dt_hydro = dt_CFL * P.tsc;
  const double dt_cond = tc_get_dt_cond(O);
  bool SuperStep = true;
  int Nsteps; double fstep;
  if (dt_cond < dt_hydro) tc_ST_steps(dt_hydro / dt_cond, Nsteps, fstep);
  else { SuperStep = false; fstep = dt_hydro / dt_cond; Nsteps = 1; }
  O.tc_dt_cond = dt_cond; O.tc_nsteps = Nsteps;
  for (int n = 1; n <= Nsteps; ++n) {
    double dts;
    if (SuperStep) dts = dt_cond * fstep * tc_substep(n, Nsteps, TC_nu) / P.Psc / P.rsc;   // :732
    else dts = dt_hydro / (double)Nsteps / P.Psc / P.rsc;
    for_blocks(O, [&](Block& b) {
      if (P.th_cond == GX_TC_ANISOTROPIC) tc_mhd_heatfluxes(P, b);
      if (P.th_cond == GX_TC_ISOTROPIC) tc_heatfluxes(P, b);
      for (int k = 1; k <= P.nz; ++k) for (int j = 1; j <= P.ny; ++j) for (int i = 1; i <= P.nx; ++i)      // :749-757
        b.u(5, i, j, k) = b.u(5, i, j, k) - dts * ((b.f(5, i, j, k) - b.f(5, i - 1, j, k)) / P.dx
                                                + (b.g(5, i, j, k) - b.g(5, i, j - 1, k)) / P.dy
                                                + (b.h(5, i, j, k) - b.h(5, i, j, k - 1)) / P.dz);
    });
    tc_thermal_bounds(O);                                                // :761
    for_blocks(O, [&](Block& b) { calcprim(P, b.u, b.primit, b.Temp, false, &b.primit0); });   // :764
  }
}

// src/hydro_solver.f90:134-229 tstep  (hydro/MHD part + COOL_H; the other operator-split add-ons are out of scope)
static int tstep(Oracle& O, double dt_CFL) {
  const Par& P = O.P;
  std::atomic<int> err(0);
  double dtm = dt_CFL / 2.;                                             // :152
  for_blocks(O, [&](Block& b) { err |= split_all(P) ? fluxes_split_all(P, b, 1) : fluxes(P, b, 1); });   // :155-161
  step(O, dtm);                                                         // :165
  boundaryII(O);                                                        // :169
  for_blocks(O, [&](Block& b) { calcprim(P, b.up, b.primit, b.Temp, false, &b.primit0); });   // :170
  for_blocks(O, [&](Block& b) { err |= split_all(P) ? fluxes_split_all(P, b, 2) : fluxes(P, b, 2); });   // :174-180
  step(O, dt_CFL);                                                      // :184
  viscous_copy(O);                                                      // :188
  if (P.cooling == GX_COOL_H) coolingh(O, dt_CFL);                      // :202-204
  boundaryI(O);                                                         // :216
  for_blocks(O, [&](Block& b) { calcprim(P, b.u, b.primit, b.Temp, false, &b.primit0); }); // :218-220 (cooling NONE/H branch)
  if (P.th_cond != 0) thermal_conduction(O, dt_CFL);                    // :227
  if (err) O.error_flag = true;
  return err;
}

// src/hydro_core.f90:623-697 get_timestep
static void get_timestep(Oracle& O, int current_iter, int n_iter, double current_time, double tprint, double& dt, int& dump_flag) {
  const Par& P = O.P;
  std::vector<double> dtps(O.B.size(), 0.0);
  for_blocks(O, [&](Block& b) {
    double dtp = 1.e30;                                                 // :642
    for (int k = 1; k <= P.nz; ++k) for (int j = 1; j <= P.ny; ++j) for (int i = 1; i <= P.nx; ++i) {
      const double* p = b.primit.cell(i, j, k);
      if (P.mhd) {
        double cx, cy, cz;
        if (split_all(P)) {                                             // :650-654
          const double* p0 = b.primit0.cell(i, j, k);
          cfast(P, p[4] + p0[4], p[0] + p0[0], p[5] + p0[5], p[6] + p0[6], p[7] + p0[7], cx, cy, cz);
        } else
        cfast(P, p[4], p[0], p[5], p[6], p[7], cx, cy, cz);
        dtp = std::min(dtp, P.dx / (std::fabs(p[1]) + cx));
        dtp = std::min(dtp, P.dy / (std::fabs(p[2]) + cy));
        dtp = std::min(dtp, P.dz / (std::fabs(p[3]) + cz));
      } else {
        double c = csound(P, p[4], p[0]);
        dtp = std::min(dtp, P.dx / (std::fabs(p[1]) + c));
        dtp = std::min(dtp, P.dy / (std::fabs(p[2]) + c));
        dtp = std::min(dtp, P.dz / (std::fabs(p[3]) + c));
      }
    }
    if (current_iter <= n_iter) dtp = P.cfl * std::pow(2., -(double)(n_iter + 1 - current_iter)) * dtp;   // :677-682
    else dtp = P.cfl * dtp;
    dtps[b.rank] = dtp;
  });
  dt = dtps[0];
  for (double v : dtps) dt = std::min(dt, v);                           // :685 mpi_allreduce(MIN)
  if ((current_time + dt) >= tprint) {                                  // :691-694
    dt = tprint - current_time;
    dump_flag = 1;
  }
}

// ---------------------------------------------------------------------------
// problem modules
// OT/orzag_tang.f90:14-70  init_ot + impose_ot  (rsc = xphys/xmax = 1 in OT/parameters.f90:159)
static void impose_ot(Oracle& O, double rsc) {
  const Par& P = O.P;
  const double pi = std::acos(-1.);
  const double twopi = 2. * pi;
  const double rho = 25. / (36. * pi), p = 5. / (12. * pi);
  for_blocks(O, [&](Block& b) {
    for (int i = P.nxmin; i <= P.nxmax; ++i) for (int j = P.nymin; j <= P.nymax; ++j) for (int k = P.nzmin; k <= P.nzmax; ++k) {
      double x = ((double)(i + b.coords[0] * P.nx) + 0.5) * P.dx * rsc;
      double y = ((double)(j + b.coords[1] * P.ny) + 0.5) * P.dy * rsc;
      double vx = -std::sin(y * twopi), vy = std::sin(x * twopi), vz = 0.;
      b.u(1, i, j, k) = rho;
      b.u(2, i, j, k) = rho * vx;
      b.u(3, i, j, k) = rho * vy;
      b.u(4, i, j, k) = rho * vz;
      if (P.neqdyn == 8) {
        double bx = -std::sin(y * twopi) / std::sqrt(4 * pi);
        double by = std::sin(2. * x * twopi) / std::sqrt(4 * pi);
        double bz = 0.;
        b.u(5, i, j, k) = 0.5 * rho * (vx * vx + vy * vy + vz * vz) + P.cv * p + 0.5 * (bx * bx + by * by + bz * bz);
        b.u(6, i, j, k) = bx; b.u(7, i, j, k) = by; b.u(8, i, j, k) = bz;
      } else {
        b.u(5, i, j, k) = 0.5 * rho * (vx * vx + vy * vy + vz * vz) + P.cv * p;
      }
    }
  });
}

// EXO/exoplanet.f90:59-117 init_exo  (scalings from EXO/parameters.f90:152-166)
static void init_exo(Oracle& O, double rsc, double rhosc, double Tempsc, double vsc2, double tsc, double bsc) {
  const double pi = std::acos(-1.);
  const double msun = 1.99E33, rsun = 6.955e10, yr = 3.1536E7, mjup = 1.898E30, Rjup = 7.1492E9, AU = 1.496e13, day = 86400.;
  O.MassS = 1.1 * msun;
  O.RsS = 1.2 * rsun;
  double AMDOT = 2.E-14 * msun / yr;
  O.TSW = 1.56E6;
  O.RSW = 1.2 * rsun;
  O.VSW = 1.e5;
  O.dsw = ((AMDOT / O.RSW) / (4 * pi * O.RSW * O.VSW));
  O.bsw = 1.0;
  O.MassP = 0.67 * mjup;
  double AMPDOT = 1.E10;
  O.TPW = 1E4;
  O.RPW = 3. * 1.38 * Rjup;
  O.VPW = 10.e5;
  O.dpw = ((AMPDOT / O.RPW) / (4 * pi * O.RPW * O.VPW));
  O.bpw = 0.04;
  O.rorb = .047 * AU;
  O.torb = 3.52 * day;
  O.dsw = O.dsw / rhosc;  O.VSW = O.VSW / std::sqrt(vsc2);  O.TSW = O.TSW / Tempsc;  O.RSW = O.RSW / rsc;  O.RsS = O.RsS / rsc;
  O.bsw = O.bsw / bsc;  O.bpw = O.bpw / bsc;  O.dpw = O.dpw / rhosc;  O.VPW = O.VPW / std::sqrt(vsc2);
  O.TPW = O.TPW / Tempsc;  O.RPW = O.RPW / rsc;  O.rorb = O.rorb / rsc;  O.torb = O.torb / tsc;
  O.omegap = 2. * pi / O.torb;
  O.xp = O.rorb * std::cos(-25. * pi / 180.);
  O.yp = 0.;
  O.zp = O.rorb * std::sin(-25. * pi / 180.);
  O.exo_rsc = rsc; O.exo_vsc2 = vsc2;
}

// EXO/exoplanet.f90:125-266 impose_exo
static void impose_exo(Oracle& O, Block& b, Arr4& u, double time) {
  const Par& P = O.P;
  const double pi = std::acos(-1.);
  const int nq = P.neqdyn;
  double phi = -25. * pi / 180.;
  O.xp = O.rorb * std::cos(O.omegap * time + phi);
  O.zp = O.rorb * std::sin(O.omegap * time + phi);
  double vxorb = -O.omegap * O.rorb * std::sin(O.omegap * time + phi);
  double vzorb = O.omegap * O.rorb * std::cos(O.omegap * time + phi);
  double vyorb = 0.;
  for (int i = P.nxmin; i <= P.nxmax; ++i) for (int j = P.nymin; j <= P.nymax; ++j) for (int k = P.nzmin; k <= P.nzmax; ++k) {
    double x = ((double)(i + b.coords[0] * P.nx - P.nxtot / 2) + 0.5) * P.dx;
    double y = ((double)(j + b.coords[1] * P.ny - P.nytot / 2) + 0.5) * P.dy;
    double z = ((double)(k + b.coords[2] * P.nz - P.nztot / 2) + 0.5) * P.dz;
    double xpl = x - O.xp, ypl = y, zpl = z - O.zp;
    double rads = std::sqrt(x * x + y * y + z * z);
    double radp = std::sqrt(xpl * xpl + ypl * ypl + zpl * zpl);
    if (rads <= O.RSW) {
      if (rads == 0.) rads = P.dx * 0.10;
      double velx = O.VSW * x / rads, vely = O.VSW * y / rads, velz = O.VSW * z / rads, dens = O.dsw;
      u(1, i, j, k) = dens; u(2, i, j, k) = dens * velx; u(3, i, j, k) = dens * vely; u(4, i, j, k) = dens * velz;
      if (P.pmhd || P.mhd) {
        double q3 = O.RSW / rads;
        double cpi = O.bsw * (q3 * q3 * q3) / (2. * (rads * rads));
        u(6, i, j, k) = 3. * y * x * cpi;
        u(7, i, j, k) = (3. * (y * y) - rads * rads) * cpi;
        u(8, i, j, k) = 3. * y * z * cpi;
      }
      if (P.mhd) u(5, i, j, k) = 0.5 * dens * (velx * velx + vely * vely + velz * velz) + P.cv * dens * O.TSW
                                 + 0.5 * (u(6, i, j, k) * u(6, i, j, k) + u(7, i, j, k) * u(7, i, j, k) + u(8, i, j, k) * u(8, i, j, k));
      else u(5, i, j, k) = 0.5 * dens * (velx * velx + vely * vely + velz * velz) + P.cv * dens * 1.9999 * O.TSW;
      if (P.passives) { u(nq + 1, i, j, k) = 0.0001 * dens; u(nq + 2, i, j, k) = dens; }
    } else if (radp <= O.RPW) {
      if (radp == 0.) radp = P.dx * 0.10;
      double velx = vxorb + O.VPW * xpl / radp, vely = vyorb + O.VPW * ypl / radp, velz = vzorb + O.VPW * zpl / radp, dens = O.dpw;
      u(1, i, j, k) = dens; u(2, i, j, k) = dens * velx; u(3, i, j, k) = dens * vely; u(4, i, j, k) = dens * velz;
      if (P.neqdyn == 8) {
        double q3 = O.RPW / radp;
        double cpi = O.bpw * (q3 * q3 * q3) / (2. * (radp * radp));
        u(6, i, j, k) = 3. * ypl * xpl * cpi;
        u(7, i, j, k) = (3. * (ypl * ypl) - radp * radp) * cpi;
        u(8, i, j, k) = 3. * ypl * zpl * cpi;
      }
      if (P.mhd) u(5, i, j, k) = 0.5 * dens * (velx * velx + vely * vely + velz * velz) + P.cv * dens * 1.8 * O.TPW
                                 + 0.5 * (u(6, i, j, k) * u(6, i, j, k) + u(7, i, j, k) * u(7, i, j, k) + u(8, i, j, k) * u(8, i, j, k));
      else u(5, i, j, k) = 0.5 * dens * (velx * velx + vely * vely + velz * velz) + P.cv * dens * 1.8 * O.TPW;
      if (P.passives) { u(nq + 1, i, j, k) = 0.2 * dens; u(nq + 2, i, j, k) = -dens; }
    }
  }
}

// EXO/user_mod.f90:58-121 initial_conditions
static void exo_initial_conditions(Oracle& O) {
  const Par& P = O.P;
  const int nq = P.neqdyn;
  for (auto& b : O.B) {
    for (int i = P.nxmin; i <= P.nxmax; ++i) for (int j = P.nymin; j <= P.nymax; ++j) for (int k = P.nzmin; k <= P.nzmax; ++k) {
      double x = ((double)(i + b.coords[0] * P.nx - P.nxtot / 2) + 0.5) * P.dx;
      double y = ((double)(j + b.coords[1] * P.ny - P.nytot / 2) + 0.5) * P.dy;
      double z = ((double)(k + b.coords[2] * P.nz - P.nztot / 2) + 0.5) * P.dz;
      double rads = std::sqrt(x * x + y * y + z * z);
      double velx = O.VSW * x / rads, vely = O.VSW * y / rads, velz = O.VSW * z / rads;
      double dens = O.dsw * (O.RSW * O.RSW) / (rads * rads);
      b.u(1, i, j, k) = dens; b.u(2, i, j, k) = dens * velx; b.u(3, i, j, k) = dens * vely; b.u(4, i, j, k) = dens * velz;
      if (P.pmhd || P.mhd) {
        double q3 = O.RSW / rads;
        double cpi = O.bsw * (q3 * q3 * q3) / (2. * (rads * rads));
        b.u(6, i, j, k) = 3. * y * x * cpi;
        b.u(7, i, j, k) = (3. * (y * y) - rads * rads) * cpi;
        b.u(8, i, j, k) = 3. * y * z * cpi;
      }
      if (P.mhd) b.u(5, i, j, k) = 0.5 * dens * (O.VSW * O.VSW) + P.cv * dens * O.TSW
                                   + 0.5 * (b.u(6, i, j, k) * b.u(6, i, j, k) + b.u(7, i, j, k) * b.u(7, i, j, k) + b.u(8, i, j, k) * b.u(8, i, j, k));
      else b.u(5, i, j, k) = 0.5 * dens * (velx * velx + vely * vely + velz * velz) + P.cv * dens * 1.9999 * O.TSW;
      if (P.passives) { b.u(nq + 1, i, j, k) = 0.0001 * dens; b.u(nq + 2, i, j, k) = dens; }
    }
    impose_exo(O, b, b.u, 0.);
  }
}

static void user_bc_dispatch(Oracle& O, Block& b, Arr4& A, int order) {
  if (O.builtin_user == 1) { if (order >= 1) impose_exo(O, b, A, O.time); }   // EXO/user_mod.f90:131-144
  else if (O.user_bc) O.user_bc(A.d.data(), order, b.coords, O.time, O.user_bc_ctx);
}

// ---------------------------------------------------------------------------
static Oracle* create(const gx_config& c) {
  Oracle* O = new Oracle();
  Par& P = O->P;
  P.nxtot = c.nxtot; P.nytot = c.nytot; P.nztot = c.nztot;
  P.NBX = std::max(1, c.nbx); P.NBY = std::max(1, c.nby); P.NBZ = std::max(1, c.nbz);
  P.nx = P.nxtot / P.NBX; P.ny = P.nytot / P.NBY; P.nz = P.nztot / P.NBZ;
  P.neq = c.neq; P.neqdyn = c.neqdyn; P.npas = c.npas;
  P.mhd = c.mhd; P.pmhd = c.pmhd; P.passives = c.passives;
  P.riemann_solver = c.riemann_solver; P.slope_limiter = c.slope_limiter; P.eq_of_state = c.eq_of_state;
  P.enable_flux_cd = c.enable_flux_cd; P.eight_wave = c.eight_wave; P.user_source_terms = c.user_source_terms;
  P.bc_left = c.bc_left; P.bc_right = c.bc_right; P.bc_bottom = c.bc_bottom; P.bc_top = c.bc_top; P.bc_out = c.bc_out; P.bc_in = c.bc_in;
  P.bc_user = c.bc_user;
  P.dx = c.dx; P.dy = c.dy; P.dz = c.dz; P.cv = c.cv; P.gamma = c.gamma; P.Tempsc = c.Tempsc; P.cfl = c.cfl; P.eta = c.eta;
  P.cooling = c.cooling; P.tsc = c.tsc;
  P.th_cond = c.th_cond; P.tc_saturation = c.tc_saturation != 0;
  P.rsc = c.rsc; P.rhosc = c.rhosc; P.vsc2 = c.vsc2; P.vsc = std::sqrt(c.vsc2); P.Psc = c.rhosc * c.vsc2; P.bsc = c.bsc; P.mu = c.mu;
  P.nxmin = -1; P.nxmax = P.nx + 2; P.nymin = -1; P.nymax = P.ny + 2; P.nzmin = -1; P.nzmax = P.nz + 2;
  const bool perx = (P.bc_left == GX_BC_PERIODIC && P.bc_right == GX_BC_PERIODIC);    // src/init.f90:64-66
  const bool pery = (P.bc_bottom == GX_BC_PERIODIC && P.bc_top == GX_BC_PERIODIC);
  const bool perz = (P.bc_out == GX_BC_PERIODIC && P.bc_in == GX_BC_PERIODIC);
  auto rank_of = [&](int cx, int cy, int cz) { return (cx * P.NBY + cy) * P.NBZ + cz; };
  auto shift = [&](int cidx, int n, bool per) { if (cidx < 0 || cidx >= n) { if (!per) return -1; return (cidx + n) % n; } return cidx; };
  O->B.resize((size_t)P.NBX * P.NBY * P.NBZ);
  for (int cx = 0; cx < P.NBX; ++cx) for (int cy = 0; cy < P.NBY; ++cy) for (int cz = 0; cz < P.NBZ; ++cz) {
    Block& b = O->B[rank_of(cx, cy, cz)];
    b.rank = rank_of(cx, cy, cz);
    b.coords[0] = cx; b.coords[1] = cy; b.coords[2] = cz;
    int l = shift(cx - 1, P.NBX, perx), r = shift(cx + 1, P.NBX, perx);   // mpi_cart_shift, src/init.f90:108-110
    int bo = shift(cy - 1, P.NBY, pery), t = shift(cy + 1, P.NBY, pery);
    int o = shift(cz - 1, P.NBZ, perz), in = shift(cz + 1, P.NBZ, perz);
    b.left = l < 0 ? -1 : rank_of(l, cy, cz);  b.right = r < 0 ? -1 : rank_of(r, cy, cz);
    b.bottom = bo < 0 ? -1 : rank_of(cx, bo, cz);  b.top = t < 0 ? -1 : rank_of(cx, t, cz);
    b.out = o < 0 ? -1 : rank_of(cx, cy, o);  b.in = in < 0 ? -1 : rank_of(cx, cy, in);
    b.u.alloc(P.neq, P.nx, P.ny, P.nz); b.up.alloc(P.neq, P.nx, P.ny, P.nz); b.primit.alloc(P.neq, P.nx, P.ny, P.nz);
    b.f.alloc(P.neq, P.nx, P.ny, P.nz); b.g.alloc(P.neq, P.nx, P.ny, P.nz); b.h.alloc(P.neq, P.nx, P.ny, P.nz);
    if (P.enable_flux_cd && P.neqdyn == 8) b.e.alloc(3, P.nx, P.ny, P.nz);
    b.Temp.assign((size_t)(P.nx + 4) * (P.ny + 4) * (P.nz + 4), 0.0);
    if (split_all(P)) b.primit0.alloc(P.neq, P.nx, P.ny, P.nz);          // src/init.f90:153-154
  }
  return O;
}

}  // namespace orc

// ===========================================================================
// C interface for ctypes (tests/, bench.py)
using orc::Oracle;
extern "C" {

void* orc_create(const gx_config* cfg) { return orc::create(*cfg); }
void orc_destroy(void* h) { delete (Oracle*)h; }
void orc_set_threads(void* h, int n) { ((Oracle*)h)->nthreads = std::max(1, n); }
int orc_num_blocks(void* h) { return (int)((Oracle*)h)->B.size(); }
void orc_block_dims(void* h, int* nx, int* ny, int* nz) { Oracle* O = (Oracle*)h; *nx = O->P.nx; *ny = O->P.ny; *nz = O->P.nz; }
void orc_block_coords(void* h, int b, int* c) { Oracle* O = (Oracle*)h; for (int q = 0; q < 3; ++q) c[q] = O->B[b].coords[q]; }
void orc_block_neighbors(void* h, int b, int* n6) {
  Oracle* O = (Oracle*)h; const orc::Block& B = O->B[b];
  n6[0] = B.left; n6[1] = B.right; n6[2] = B.bottom; n6[3] = B.top; n6[4] = B.out; n6[5] = B.in;
}
void orc_set_time(void* h, double t) { ((Oracle*)h)->time = t; }
int orc_error(void* h) { return ((Oracle*)h)->error_flag ? 1 : 0; }

// which: 0 u, 1 up, 2 primit, 3 f, 4 g, 5 h, 6 e, 7 Temp, 8 primit0
static orc::Arr4* pick(Oracle* O, int b, int which) {
  orc::Block& B = O->B[b];
  switch (which) { case 0: return &B.u; case 1: return &B.up; case 2: return &B.primit; case 3: return &B.f; case 4: return &B.g; case 5: return &B.h; case 6: return &B.e; case 8: return &B.primit0; }
  return nullptr;
}
int64_t orc_block_array_size(void* h, int which) {
  Oracle* O = (Oracle*)h;
  if (which == 7) return (int64_t)O->B[0].Temp.size();
  orc::Arr4* A = pick(O, 0, which); return A ? (int64_t)A->d.size() : 0;
}
void orc_get_block(void* h, int b, int which, double* out) {
  Oracle* O = (Oracle*)h;
  if (which == 7) { std::memcpy(out, O->B[b].Temp.data(), O->B[b].Temp.size() * sizeof(double)); return; }
  orc::Arr4* A = pick(O, b, which); std::memcpy(out, A->d.data(), A->d.size() * sizeof(double));
}
void orc_set_block(void* h, int b, int which, const double* in) {
  Oracle* O = (Oracle*)h; orc::Arr4* A = pick(O, b, which); std::memcpy(A->d.data(), in, A->d.size() * sizeof(double));
}
// interior of all blocks gathered into a global (n1, nxtot, nytot, nztot) column-major array
void orc_gather_interior(void* h, int which, double* out) {
  Oracle* O = (Oracle*)h; const orc::Par& P = O->P;
  for (auto& B : O->B) {
    orc::Arr4* A = pick(O, B.rank, which);
    for (int k = 1; k <= P.nz; ++k) for (int j = 1; j <= P.ny; ++j) for (int i = 1; i <= P.nx; ++i) {
      size_t gi = (size_t)(i - 1 + B.coords[0] * P.nx), gj = (size_t)(j - 1 + B.coords[1] * P.ny), gk = (size_t)(k - 1 + B.coords[2] * P.nz);
      double* o = out + (size_t)A->n1 * (gi + (size_t)P.nxtot * (gj + (size_t)P.nytot * gk));
      const double* c = A->cell(i, j, k);
      for (int q = 0; q < A->n1; ++q) o[q] = c[q];
    }
  }
}
// scatter a global array WITH 2 ghost layers, shape (neq, nxtot+4, nytot+4, nztot+4), into every block's
// u including the block's own ghosts (what initial_conditions does: it fills nxmin:nxmax, user_mod.f90:58)
void orc_scatter_u_with_ghosts(void* h, const double* g) {
  Oracle* O = (Oracle*)h; const orc::Par& P = O->P;
  const size_t GX = P.nxtot + 4, GY = P.nytot + 4;
  for (auto& B : O->B)
    for (int k = P.nzmin; k <= P.nzmax; ++k) for (int j = P.nymin; j <= P.nymax; ++j) for (int i = P.nxmin; i <= P.nxmax; ++i) {
      size_t gi = (size_t)(i + 1 + B.coords[0] * P.nx), gj = (size_t)(j + 1 + B.coords[1] * P.ny), gk = (size_t)(k + 1 + B.coords[2] * P.nz);
      const double* s = g + (size_t)P.neq * (gi + GX * (gj + GY * gk));
      double* c = B.u.cell(i, j, k);
      for (int q = 0; q < P.neq; ++q) c[q] = s[q];
    }
}

// split-all solvers: scatter a global background WITH ghosts into every block's primit0 (the host's job in the reference)
void orc_scatter_primit0_with_ghosts(void* h, const double* g) {
  Oracle* O = (Oracle*)h; const orc::Par& P = O->P;
  const size_t GX = P.nxtot + 4, GY = P.nytot + 4;
  for (auto& B : O->B)
    for (int k = P.nzmin; k <= P.nzmax; ++k) for (int j = P.nymin; j <= P.nymax; ++j) for (int i = P.nxmin; i <= P.nxmax; ++i) {
      size_t gi = (size_t)(i + 1 + B.coords[0] * P.nx), gj = (size_t)(j + 1 + B.coords[1] * P.ny), gk = (size_t)(k + 1 + B.coords[2] * P.nz);
      const double* s = g + (size_t)P.neq * (gi + GX * (gj + GY * gk));
      double* c = B.primit0.cell(i, j, k);
      for (int q = 0; q < P.neq; ++q) c[q] = s[q];
    }
}
void orc_riemann_split_all(void* h, const double* pl, const double* pr, const double* p0l, const double* p0r, double* ff) {
  orc::prim2fhlleSplitAll(((Oracle*)h)->P, pl, pr, p0l, p0r, ff);
}

void orc_impose_ot(void* h, double rsc) { orc::impose_ot(*(Oracle*)h, rsc); }
void orc_init_exo(void* h, double rsc, double rhosc, double Tempsc, double vsc2, double tsc, double bsc) {
  Oracle* O = (Oracle*)h; orc::init_exo(*O, rsc, rhosc, Tempsc, vsc2, tsc, bsc); O->builtin_user = 1;
}
void orc_exo_initial_conditions(void* h) { orc::exo_initial_conditions(*(Oracle*)h); }
void orc_exo_params(void* h, double* out19) {
  Oracle* O = (Oracle*)h;
  double v[19] = {O->RSW, O->TSW, O->VSW, O->dsw, O->RsS, O->bsw, O->bpw, O->RPW, O->TPW, O->VPW, O->dpw, O->torb, O->rorb, O->omegap, O->MassS, O->MassP, O->xp, O->yp, O->zp};
  std::memcpy(out19, v, sizeof(v));
}
void orc_set_user_bc(void* h, orc::user_bc_fn fn, void* ctx) { Oracle* O = (Oracle*)h; O->user_bc = fn; O->user_bc_ctx = ctx; }
void orc_set_user_source(void* h, orc::user_src_fn fn, void* ctx) { Oracle* O = (Oracle*)h; O->user_src = fn; O->user_src_ctx = ctx; }

// the calls main.f90 makes (src/main.f90:73-79, 97, 106)
void orc_boundaryI(void* h) { orc::boundaryI(*(Oracle*)h); }
void orc_boundaryII(void* h) { orc::boundaryII(*(Oracle*)h); }
void orc_calcprim_u(void* h) { Oracle* O = (Oracle*)h; orc::for_blocks(*O, [&](orc::Block& b) { orc::calcprim(O->P, b.u, b.primit, b.Temp, false, &b.primit0); }); }
void orc_calcprim_up(void* h) { Oracle* O = (Oracle*)h; orc::for_blocks(*O, [&](orc::Block& b) { orc::calcprim(O->P, b.up, b.primit, b.Temp, false, &b.primit0); }); }
void orc_start(void* h) { orc_boundaryI(h); orc_calcprim_u(h); }
void orc_get_timestep(void* h, int current_iter, int n_iter, double current_time, double tprint, double* dt, int* dump_flag) {
  orc::get_timestep(*(Oracle*)h, current_iter, n_iter, current_time, tprint, *dt, *dump_flag);
}
int orc_tstep(void* h, double dt_cfl) { return orc::tstep(*(Oracle*)h, dt_cfl); }
int orc_fluxes(void* h, int choice) { Oracle* O = (Oracle*)h; int e = 0; for (auto& b : O->B) e |= orc::fluxes(O->P, b, choice); return e; }
void orc_step(void* h, double dt) { orc::step(*(Oracle*)h, dt); }
void orc_viscous_copy(void* h) { orc::viscous_copy(*(Oracle*)h); }

// main.f90:94-125 loop body without output; returns wall seconds spent in get_timestep+tstep
double orc_run(void* h, int n_steps, int n_iter_ramp, double* time, int* iter, double* last_dt) {
  Oracle* O = (Oracle*)h;
  auto t0 = std::chrono::steady_clock::now();
  for (int s = 0; s < n_steps; ++s) {
    double dt; int dump = 0;
    orc::get_timestep(*O, *iter, n_iter_ramp, *time, 1.e300, dt, dump);
    O->time = *time;
    orc::tstep(*O, dt);
    *time += dt; *iter += 1; *last_dt = dt;
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// ---- single-cell entry points for known-answer tests ----
void orc_u2prim(void* h, const double* uu, double* prim, double* T) { orc::u2prim(((Oracle*)h)->P, uu, prim, *T); }
void orc_prim2u(void* h, const double* prim, double* uu) { orc::prim2u(((Oracle*)h)->P, prim, uu); }
void orc_prim2f(void* h, const double* prim, double* ff) { orc::prim2f(((Oracle*)h)->P, prim, ff); }
int orc_riemann(void* h, const double* pl, const double* pr, double* ff) { return orc::riemann(((Oracle*)h)->P, pl, pr, ff); }
void orc_limiter(void* h, const double* pll, double* pl, double* pr, const double* prr) {
  Oracle* O = (Oracle*)h; orc::limiter(O->P.slope_limiter, pll, pl, pr, prr, O->P.neq);
}
double orc_average(int slope_limiter, double a, double b) { return orc::average(slope_limiter, a, b); }
void orc_cfast(void* h, double p, double d, double bx, double by, double bz, double* c3) { orc::cfast(((Oracle*)h)->P, p, d, bx, by, bz, c3[0], c3[1], c3[2]); }
double orc_cfastX(void* h, const double* prim) { return orc::cfastX(((Oracle*)h)->P, prim); }
void orc_cool_atomic(void* h, double dt_seconds, double* uu) { orc::cool_atomic(((Oracle*)h)->P, dt_seconds, uu); }
double orc_cool_rate(int which, double T) { return which == 0 ? orc::cool_alpha(T) : which == 1 ? orc::cool_colf(T) : orc::cool_betah(T); }
double orc_cool_aloss(double x1, double x2, double dt, double den, double dh0, double te) { return orc::cool_aloss(x1, x2, dt, den, dh0, te); }
// thermal conduction: what the reference logs per call (dt_cond in seconds, number of substeps), and the pieces for KATs
void orc_tc_info(void* h, double* dt_cond, int* nsteps) { Oracle* O = (Oracle*)h; *dt_cond = O->tc_dt_cond; *nsteps = O->tc_nsteps; }
void orc_thermal_conduction(void* h, double dt_cfl) { orc::thermal_conduction(*(Oracle*)h, dt_cfl); }
double orc_tc_superstep(int n) { return orc::tc_superstep(n, std::sqrt(orc::TC_nu)); }
double orc_tc_substep(int j, int n) { return orc::tc_substep(j, n, orc::TC_nu); }
void orc_tc_st_steps(double fs, int* ns, double* fstep) { orc::tc_ST_steps(fs, *ns, *fstep); }
double orc_tc_ksp(int which, double T, double dens, double B2) { return which == 0 ? orc::tc_Ksp(T) : which == 1 ? orc::tc_Ksp_parl(T) : orc::tc_Ksp_perp(T, dens, B2); }
double orc_csound(void* h, double p, double d) { return orc::csound(((Oracle*)h)->P, p, d); }

}  // extern "C"
