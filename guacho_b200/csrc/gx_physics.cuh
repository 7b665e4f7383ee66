// gx_physics.cuh — per-cell / per-interface device functions of the hydro/MHD step.
//
// Register-resident, branch-uniform-where-possible formulations of the reference's
// cell-level routines.  Slot convention for a state rotated into sweep direction d:
//   w[0]=rho  w[1]=v_n  w[2]=v_t1  w[3]=v_t2  w[4]=p  w[5]=B_n  w[6]=B_t1  w[7]=B_t2
// which is what swapy/swapz (src/hydro_core.f90:485-534) produce; the rotation is done
// by the loader through a component permutation, never by moving data.
// Operation ORDER follows the reference expressions so that a -fmad=false build is
// bit-comparable with the no-FMA x86 reference build.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "../../include/guacho_gx.h"

namespace gxp {

struct Phys {            // the scalar `parameter`s the cell routines read
  double cv, gamma, Tempsc;
  double inv_cv;         // 1/cv (fast build: p = (...)*inv_cv instead of an IEEE division)
  double m4gamma;        // -4 gamma (fast build: discriminant of the fast speed)
  double inv_Tempsc;     // 1/Tempsc (fast build: pressure from the floored temperature)
  int eos;               // GX_EOS_*
  int neqdyn, npas;      // neq = neqdyn + npas
};

__device__ __forceinline__ double sign1(double x) { return copysign(1.0, x); }   // Fortran sign(1.,x)

// max / min policy.  fmax/fmin expand to DSETP + selects + NaN quieting + register moves
// (~6 issue slots each on sm_100a); the fast build uses a bare compare-and-select (3 slots).
// NaNs never reach these in a live run (the reference stops on NaN, hlld.f90:316-317).
#if defined(GX_FLAVOUR_FAST)
__device__ __forceinline__ double gx_max(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double gx_min(double a, double b) { return a < b ? a : b; }
#else
__device__ __forceinline__ double gx_max(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ double gx_min(double a, double b) { return fmin(a, b); }
#endif

// ---- division / square root policy ----
// strict build (GX_FLAVOUR_STRICT, -fmad=false): IEEE a/b and sqrt, one per occurrence in the
// reference expression, so results are bit-comparable with the reference arithmetic.
// fast build: divisions by the same denominator share ONE reciprocal (MUFU.RCP64H seed + a
// third-order Newton step, ~1 ulp) and square roots come from MUFU.RSQ64H + two coupled
// Newton steps that also yield 1/sqrt.  FP64 division is ~12 FP64-pipe instructions on
// sm_100a and HLLD has 25 of them per interface, so this removes most of the FP64 work.
#if defined(GX_FLAVOUR_FAST)
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));      // rel. error <= 2^-23
  double e = fma(-x, r, 1.0);
  double t = fma(e, e, e);                                    // e + e^2
  return fma(r, t, r);                                        // error ~ e^3 = 2^-69
}
// 1/sqrt(x) for x > 0 (callers guarantee x > 0: densities, or a discriminant clamped to a tiny positive number).
// MUFU.RSQ64H seed y (rel. error d <= 2^-22) and ONE third-order step: with e = 1 - x y^2 (= -2d),
// 1/sqrt(x) = y (1 - e)^(-1/2) = y (1 + e/2 + 3 e^2/8 + 5 e^3/16 ...): the truncation error 5/16 e^3 is below 2^-64.
// Five FP64 instructions after the seed (two coupled Newton steps cost eight).
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double t = x * y;
  const double e = fma(-t, y, 1.0);
  const double p = fma(0.375, e, 0.5);
  return fma(y, e * p, y);
}
__device__ __forceinline__ void fast_sqrt_rsqrt(double x, double& s, double& rs) { rs = fast_rsqrt(x); s = x * rs; }
__device__ __forceinline__ double gx_sqrt(double x) { return x * fast_rsqrt(x); }
// discriminant of the fast-speed formula: >= 0 analytically, may round to -eps (or be exactly 0)
__device__ __forceinline__ double gx_sqrt_disc(double x) { return gx_sqrt(gx_max(x, 1e-300)); }
struct Den {                      // a denominator used one or more times
  double inv;
  __device__ __forceinline__ explicit Den(double b) : inv(fast_rcp(b)) {}
  __device__ __forceinline__ double div(double a) const { return a * inv; }
};
struct SqrtDen {                  // sqrt(x) that is also used as a denominator
  double s, rs;
  __device__ __forceinline__ explicit SqrtDen(double x) { fast_sqrt_rsqrt(x, s, rs); }
  __device__ __forceinline__ double div(double a) const { return a * rs; }
};
__device__ __forceinline__ double sqrt_prod(const SqrtDen& a, const SqrtDen& b, double) { return a.s * b.s; }
#else
__device__ __forceinline__ double gx_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ double gx_sqrt_disc(double x) { return sqrt(x); }
struct Den {
  double b;
  __device__ __forceinline__ explicit Den(double b_) : b(b_) {}
  __device__ __forceinline__ double div(double a) const { return a / b; }
};
struct SqrtDen {
  double s;
  __device__ __forceinline__ explicit SqrtDen(double x) : s(sqrt(x)) {}
  __device__ __forceinline__ double div(double a) const { return a / s; }
};
__device__ __forceinline__ double sqrt_prod(const SqrtDen&, const SqrtDen&, double xy) { return sqrt(xy); }
#endif

// Supersonic interfaces (sl > 0 or sr < 0: the flux is the upwind physical flux) are rare.  In the production build the
// test is made warp-uniform with one vote, which keeps the override and its reconvergence bookkeeping out of the common
// path (measured: -6 % on the second-order stage kernel); the strict build keeps the reference's plain per-lane tests.
#ifndef GX_SOLVE_MASK            // lanes that solve together: the fused stage kernel solves with whole warps
#define GX_SOLVE_MASK __activemask()
#endif
#if defined(GX_FLAVOUR_FAST) && !defined(GX_NO_VOTE_SUPERSONIC)
#define GX_ANY_SUPERSONIC(sl, sr) __any_sync(GX_SOLVE_MASK, ((sl) > 0.0) || ((sr) < 0.0))
#else
#define GX_ANY_SUPERSONIC(sl, sr) true
#endif
#if defined(GX_FLAVOUR_FAST) && !defined(GX_NO_VOTE_DSTAR)
#define GX_ANY_LANE(p) __any_sync(GX_SOLVE_MASK, (p))
#else
#define GX_ANY_LANE(p) true
#endif

// ---- u2prim: src/hydro_core.f90:46-129 (dynamic variables only; passives are copies) ----
// `pas0` is the first passive (needed by EOS_H_RATE only).
// ADIABATIC: the caller guarantees eq_of_state == EOS_ADIABATIC (the fused stage kernels: gx_create routes every other
// equation of state to the pass-per-routine kernels), so the temperature-floor branches and their divisions are not compiled.
template <bool MHD, bool WANT_T = true, bool ADIABATIC = false>
__device__ __forceinline__ void u2prim(const Phys& P, const double (&u)[8], double (&w)[8], double pas0, double& T) {
  double r = gx_max(u[0], 1e-15);
  w[0] = r;
  const Den dr(r);
  w[1] = dr.div(u[1]);
  w[2] = dr.div(u[2]);
  w[3] = dr.div(u[3]);
  double ek = 0.5 * r * (w[1] * w[1] + w[2] * w[2] + w[3] * w[3]);
  double p;
#if defined(GX_FLAVOUR_FAST)
  if (MHD) p = (u[4] - ek - 0.5 * (u[5] * u[5] + u[6] * u[6] + u[7] * u[7])) * P.inv_cv;
  else p = (u[4] - ek) * P.inv_cv;
#else
  if (MHD) p = (u[4] - ek - 0.5 * (u[5] * u[5] + u[6] * u[6] + u[7] * u[7])) / P.cv;
  else p = (u[4] - ek) / P.cv;
#endif
  p = gx_max(p, 1e-16);
  if (MHD) { w[5] = u[5]; w[6] = u[6]; w[7] = u[7]; }
  T = 0.0;
  if (ADIABATIC || P.eos == GX_EOS_ADIABATIC) {
    if (WANT_T) T = (p / r) * P.Tempsc;
  } else if (P.eos == GX_EOS_SINGLE_SPECIE) {
    double rr = gx_max(r, 1e-15);
#if defined(GX_FLAVOUR_FAST)
    T = gx_max(1.0, (p * dr.inv) * P.Tempsc);          // rr == r (already floored)
    p = rr * T * P.inv_Tempsc;
#else
    T = gx_max(1.0, (p / rr) * P.Tempsc);
    p = rr * T / P.Tempsc;
#endif
  } else if (P.eos == GX_EOS_H_RATE && P.npas > 0) {
    double dentot = gx_max(2.0 * r - pas0, 1e-15);
#if defined(GX_FLAVOUR_FAST)
    T = gx_max(1.0, (p * fast_rcp(dentot)) * P.Tempsc);
    p = dentot * T * P.inv_Tempsc;
#else
    T = gx_max(1.0, (p / dentot) * P.Tempsc);
    p = dentot * T / P.Tempsc;
#endif
  }
  w[4] = p;
}

// ---- wave speeds: src/hydro_core.f90:544-604 ----
#if defined(GX_FLAVOUR_FAST)
// Production forms: no division at all.  With g = gamma p + B^2 and D = g^2 - 4 gamma p Bn^2 (rho^2 times the
// reference's discriminant), cf^2 = (g + sqrt D) / (2 rho) =: N / (2 rho), so cf = N / sqrt(2 rho N): two
// reciprocal square roots (MUFU seed + one third-order step each) instead of a reciprocal and two square roots.
__device__ __forceinline__ double csound(const Phys& P, double p, double d) {
  const double gp = P.gamma * p;
  return gp * fast_rsqrt(gp * d);                                  // sqrt(gamma p / rho)
}
// g2 = g*g, m4gp = -4 gamma p, rho2 = 2 rho, bn2 = Bn^2
__device__ __forceinline__ double cfast_from(double g, double g2, double m4gp, double rho2, double bn2) {
  const double x = fabs(fma(m4gp, bn2, g2)) + 1e-300;              // D >= 0 analytically; may round to -eps or be exactly 0
  const double n = fma(x, fast_rsqrt(x), g);                       // N = g + sqrt D
  return n * fast_rsqrt(rho2 * n);
}
// bt2 = sum of the squares of the two transverse field components (shared with the total pressure in HLLD)
__device__ __forceinline__ double cfast_dir(const Phys& P, double rho, double p, double bn, double bt2) {
  const double bn2 = bn * bn;
  const double g = fma(P.gamma, p, bn2) + bt2;
  return cfast_from(g, g * g, P.m4gamma * p, rho + rho, bn2);
}
__device__ __forceinline__ double cfastX(const Phys& P, const double (&w)[8]) {
  return cfast_dir(P, w[0], w[4], w[5], fma(w[7], w[7], w[6] * w[6]));
}
// CFL form (src/hydro_core.f90:568-581): fast speed along each axis
__device__ __forceinline__ void cfast3(const Phys& P, double p, double d, double bx, double by, double bz,
                                       double& cx, double& cy, double& cz) {
  const double bx2 = bx * bx, by2 = by * by, bz2 = bz * bz;
  const double g = fma(P.gamma, p, bx2) + (by2 + bz2);
  const double g2 = g * g, m4gp = P.m4gamma * p, rho2 = d + d;
  cx = cfast_from(g, g2, m4gp, rho2, bx2);
  cy = cfast_from(g, g2, m4gp, rho2, by2);
  cz = cfast_from(g, g2, m4gp, rho2, bz2);
}
#else
__device__ __forceinline__ double csound(const Phys& P, double p, double d) { return gx_sqrt(Den(d).div(P.gamma * p)); }

__device__ __forceinline__ double cfastX(const Phys& P, const double (&w)[8]) {
  double b2 = w[5] * w[5] + w[6] * w[6] + w[7] * w[7];
  const Den rho(w[0]);
  double cs2va2 = rho.div(P.gamma * w[4] + b2);
  return gx_sqrt(0.5 * (cs2va2 + gx_sqrt_disc(cs2va2 * cs2va2 - rho.div(rho.div(4. * P.gamma * w[4] * (w[5] * w[5]))))));
}

// CFL form (src/hydro_core.f90:568-581): fast speed along each axis
__device__ __forceinline__ void cfast3(const Phys& P, double p, double d, double bx, double by, double bz,
                                       double& cx, double& cy, double& cz) {
  double b2 = bx * bx + by * by + bz * bz;
  double gpb = P.gamma * p + b2;
  double gp4 = 4. * P.gamma * p;
  const Den dd(d);
  cx = gx_sqrt(dd.div(0.5 * (gpb + gx_sqrt_disc(gpb * gpb - gp4 * bx * bx))));
  cy = gx_sqrt(dd.div(0.5 * (gpb + gx_sqrt_disc(gpb * gpb - gp4 * by * by))));
  cz = gx_sqrt(dd.div(0.5 * (gpb + gx_sqrt_disc(gpb * gpb - gp4 * bz * bz))));
}
#endif

// Signal speeds of one CELL for the first-order stage, where the states either side of a face are cell values: the
// reference evaluates csound / cfastX per side per face (hll.f90:57-58, hlle.f90:58-59, hlld.f90:64-65), i.e. six
// times per cell and stage; the same function of the same cell state gives the same bits, so the fused stage kernel
// evaluates it once per cell and direction.  c[d] = speed along x, y, z (hydro solvers: c[0] only).
template <bool MHD>
__device__ __forceinline__ void cell_speeds(const Phys& P, const double (&w)[8], double (&c)[3]) {
  if (!MHD) { c[0] = csound(P, w[4], w[0]); c[1] = c[2] = 0.0; return; }
#if defined(GX_FLAVOUR_FAST)
  // same instruction sequence as cfast_dir for each direction (bt2 = the other two squares, y and z in storage order)
  c[0] = cfast_dir(P, w[0], w[4], w[5], fma(w[7], w[7], w[6] * w[6]));
  c[1] = cfast_dir(P, w[0], w[4], w[6], fma(w[7], w[7], w[5] * w[5]));
  c[2] = cfast_dir(P, w[0], w[4], w[7], fma(w[5], w[5], w[6] * w[6]));
#else
  { const double r[8] = {w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]}; c[0] = cfastX(P, r); }
  { const double r[8] = {w[0], w[2], w[1], w[3], w[4], w[6], w[5], w[7]}; c[1] = cfastX(P, r); }   // swapy
  { const double r[8] = {w[0], w[3], w[2], w[1], w[4], w[7], w[6], w[5]}; c[2] = cfastX(P, r); }   // swapz
#endif
}

// ---- prim2f / prim2u: src/hydro_core.f90:331-476 (non-split branches) ----
template <bool MHD>
__device__ __forceinline__ void prim2f(const Phys& P, const double (&w)[8], double (&ff)[8]) {
  double v2 = w[1] * w[1] + w[2] * w[2] + w[3] * w[3];
  if (MHD) {
    double b2 = w[5] * w[5] + w[6] * w[6] + w[7] * w[7];
    double etot = 0.5 * (w[0] * v2 + w[5] * w[5] + w[6] * w[6] + w[7] * w[7]) + P.cv * w[4];
    ff[0] = w[0] * w[1];
    ff[1] = w[0] * w[1] * w[1] + w[4] + 0.5 * (w[6] * w[6] + w[7] * w[7] - w[5] * w[5]);
    ff[2] = w[0] * w[1] * w[2] - w[5] * w[6];
    ff[3] = w[0] * w[1] * w[3] - w[5] * w[7];
    ff[4] = w[1] * (etot + w[4] + 0.5 * b2) - w[5] * (w[1] * w[5] + w[2] * w[6] + w[3] * w[7]);
    ff[5] = 0.0;
    ff[6] = w[1] * w[6] - w[5] * w[2];
    ff[7] = w[1] * w[7] - w[5] * w[3];
  } else {
    double etot = 0.5 * w[0] * v2 + P.cv * w[4];
    ff[0] = w[0] * w[1];
    ff[1] = w[0] * w[1] * w[1] + w[4];
    ff[2] = w[0] * w[1] * w[2];
    ff[3] = w[0] * w[1] * w[3];
    ff[4] = w[1] * (etot + w[4]);
  }
}

template <bool MHD>
__device__ __forceinline__ void prim2u(const Phys& P, const double (&w)[8], double (&uu)[8]) {
  uu[0] = w[0];
  uu[1] = w[0] * w[1];
  uu[2] = w[0] * w[2];
  uu[3] = w[0] * w[3];
  uu[4] = 0.5 * w[0] * (w[1] * w[1] + w[2] * w[2] + w[3] * w[3]) + P.cv * w[4];
  if (MHD) {
    uu[4] = uu[4] + 0.5 * (w[5] * w[5] + w[6] * w[6] + w[7] * w[7]);
    uu[5] = w[5]; uu[6] = w[6]; uu[7] = w[7];
  }
}

// ---- slope limiters: src/hydro_core.f90:735-796 ----
template <int LIM>
__device__ __forceinline__ double average(double a, double b) {
  if (LIM == GX_LIMITER_NO_AVERAGE) return 0.;
  if (LIM == GX_LIMITER_NO_LIMIT) return 0.5 * (a + b);
  if (LIM == GX_LIMITER_MINMOD) {
#if defined(GX_FLAVOUR_FAST)
    // same value as the reference expression (exactly a, b or 0), selected instead of computed:
    // one FP64 compare + integer sign test instead of 4 FP64 ops and two fmin/fmax expansions
    const double m = (fabs(a) < fabs(b)) ? a : b;
    return ((__double2hiint(a) ^ __double2hiint(b)) >= 0) ? m : 0.0;
#else
    double s = sign1(a);
    return s * gx_max(0., gx_min(fabs(a), s * b));
#endif
  }
  if (LIM == GX_LIMITER_VAN_LEER) {
    if (a * b <= 0.) return 0.;
    return a * b * (a + b) / (a * a + b * b);
  }
  if (LIM == GX_LIMITER_VAN_ALBADA) {
    const double delta = 1.e-7;
    return (a * (b * b + delta) + b * (a * a + delta)) / (a * a + b * b + delta);
  }
  if (LIM == GX_LIMITER_UMIST) {
    double s = sign1(a);
    double c = 0.25 * a + 0.75 * b;
    double d = 0.75 * a + 0.25 * b;
    double m = gx_min(gx_min(2. * fabs(a), 2. * s * b), gx_min(s * c, s * d));
    return s * gx_max(0., m);
  }
  if (LIM == GX_LIMITER_WOODWARD) {
    double s = sign1(a);
    double c = 0.5 * (a + b);
    double m = gx_min(gx_min(2. * fabs(a), 2. * s * b), s * c);
    return s * gx_max(0., m);
  }
  if (LIM == GX_LIMITER_SUPERBEE) {
    double s = sign1(b);
    double av1 = gx_min(2. * fabs(b), s * a);
    double av2 = gx_min(fabs(b), 2. * s * a);
    return s * gx_max(0., gx_max(av1, av2));
  }
  return 0.;
}

// reconstruct one variable at the interface between pl and pr (src/hydro_core.f90:723-731)
template <int LIM>
__device__ __forceinline__ void reconstruct(double pll, double& pl, double& pr, double prr) {
  double dl = pl - pll;
  double dm = pr - pl;
  double dr = prr - pr;
#if defined(GX_FLAVOUR_FAST) && !defined(GX_NO_MINMOD_HALF)
  if (LIM == GX_LIMITER_MINMOD) {
    // minmod(a, b) * 0.5 as  m * h  with m = the argument of smaller magnitude and h = 0.5 | 0 by the sign test: the same value
    // (m * 0.5 is exact, m * 0 adds a zero), but the "or zero" select acts on one constant word instead of the two words of m
    const double ml = (fabs(dl) < fabs(dm)) ? dl : dm, mr = (fabs(dm) < fabs(dr)) ? dm : dr;
    const double hl = ((__double2hiint(dl) ^ __double2hiint(dm)) >= 0) ? 0.5 : 0.0;
    const double hr = ((__double2hiint(dm) ^ __double2hiint(dr)) >= 0) ? 0.5 : 0.0;
    pl = fma(ml, hl, pl);
    pr = fma(-mr, hr, pr);
    return;
  }
#endif
  double al = average<LIM>(dl, dm);
  double ar = average<LIM>(dm, dr);
  pl = pl + al * 0.5;
  pr = pr - ar * 0.5;
}

// ---- passive-scalar flux bookkeeping ----
// Every solver advects passives with one of a few closed forms; the solver records
// which one applied and the scalars it needs, the sweep then applies it per passive.
enum { PAS_UPL = 0, PAS_UPR = 1, PAS_HLL = 2, PAS_HLLC_L = 3, PAS_HLLC_R = 4, PAS_HLLD_L = 5, PAS_HLLD_R = 6 };
struct PasInfo {
  int mode;
  double ul, ur, sl, sr, a, b, c;   // meaning depends on mode
};
#if defined(GX_FLAVOUR_FAST)
// Every closed form below is linear in (ql, qr): flux = cl*ql + cr*qr with coefficients that depend on the interface only.
// The production build forms them once per interface (one reciprocal) instead of dividing per passive scalar.
__device__ __forceinline__ void passive_coeffs(const PasInfo& I, double& cl, double& cr) {
  cl = 0.0; cr = 0.0;
  switch (I.mode) {
    case PAS_UPL: cl = I.ul; break;
    case PAS_UPR: cr = I.ur; break;
    case PAS_HLL: { const double r = fast_rcp(I.sr - I.sl), ss = I.sl * I.sr; cl = (I.sr * I.ul - ss) * r; cr = (ss - I.sl * I.ur) * r; break; }
    case PAS_HLLC_L: cl = I.ul + I.sl * (I.a * fast_rcp(I.b) - 1.0); break;
    case PAS_HLLC_R: cr = I.ur + I.sr * (I.a * fast_rcp(I.b) - 1.0); break;
    case PAS_HLLD_L: cl = I.a * I.b * fast_rcp(I.c); break;
    case PAS_HLLD_R: cr = I.a * I.b * fast_rcp(I.c); break;
  }
}
#endif
__device__ __forceinline__ double passive_flux(const PasInfo& I, double ql, double qr) {
  switch (I.mode) {
    case PAS_UPL: return ql * I.ul;                                       // prim2f(L): hydro_core.f90:472
    case PAS_UPR: return qr * I.ur;
    case PAS_HLL: return (I.sr * (ql * I.ul) - I.sl * (qr * I.ur) + I.sl * I.sr * (qr - ql)) / (I.sr - I.sl);   // hll.f90:76
    case PAS_HLLC_L: return ql * I.ul + I.sl * (I.a * ql / I.b - ql);     // hllc.f90:94-101: a=rhost, b=rhoL
    case PAS_HLLC_R: return qr * I.ur + I.sr * (I.a * qr / I.b - qr);
    case PAS_HLLD_L: return I.a * ql * I.b / I.c;                         // hlld.f90:151: a=sM, b=slmul, c=slmsM
    case PAS_HLLD_R: return I.a * qr * I.b / I.c;
  }
  return 0.;
}

// Every solver takes the signal speeds of the two states as arguments (csound for HLL/HLLC, cfastX for HLLE/HLLD:
// hll.f90:57-58, hllc.f90:55-56, hlle.f90:58-59, hlld.f90:64-65): the second-order sweeps evaluate them from the
// reconstructed states, the first-order stage of the fused kernel reads the per-cell values (cell_speeds above).

// ---- HLL (hydro speeds) / HLLE (fast speeds): src/hll.f90:47-82, src/hlle.f90:48-83 ----
template <bool MHD>
__device__ __forceinline__ int riemann_hll(const Phys& P, const double (&wl)[8], const double (&wr)[8], double (&ff)[8], PasInfo& I,
                                           double csl, double csr) {
  double sr = gx_max(wl[1] + csl, wr[1] + csr);
  double sl = gx_min(wl[1] - csl, wr[1] - csr);
  I.ul = wl[1]; I.ur = wr[1]; I.sl = sl; I.sr = sr;
  if (GX_ANY_SUPERSONIC(sl, sr)) {
    if (sl > 0) { prim2f<MHD>(P, wl, ff); I.mode = PAS_UPL; return 0; }
    if (sr < 0) { prim2f<MHD>(P, wr, ff); I.mode = PAS_UPR; return 0; }
  }
  double fL[8], fR[8], uL[8], uR[8];
  prim2f<MHD>(P, wl, fL); prim2f<MHD>(P, wr, fR);
  prim2u<MHD>(P, wl, uL); prim2u<MHD>(P, wr, uR);
  const int n = MHD ? 8 : 5;
  const Den ds(sr - sl);
#pragma unroll
  for (int q = 0; q < n; ++q) ff[q] = ds.div(sr * fL[q] - sl * fR[q] + sl * sr * (uR[q] - uL[q]));
  I.mode = PAS_HLL;
  return 0;
}

// ---- HLLC: src/hllc.f90:44-140 (hydro) ----
__device__ __forceinline__ int riemann_hllc_ref(const Phys& P, const double (&wl)[8], const double (&wr)[8], double (&ff)[8], PasInfo& I,
                                                double csl, double csr) {   // csl, csr = csound of either side (hllc.f90:55-56)
  double sr = gx_max(wl[1] + csl, wr[1] + csr);
  double sl = gx_min(wl[1] - csl, wr[1] - csr);
  I.ul = wl[1]; I.ur = wr[1]; I.sl = sl; I.sr = sr;
  if (sl > 0) { prim2f<false>(P, wl, ff); I.mode = PAS_UPL; return 0; }
  if (sr < 0) { prim2f<false>(P, wr, ff); I.mode = PAS_UPR; return 0; }
  double slmul = sl - wl[1];
  double srmur = sr - wr[1];
  double rholul = wl[0] * wl[1];
  double rhorur = wr[0] * wr[1];
  double sst = (srmur * rhorur - slmul * rholul - wr[4] + wl[4]) / (srmur * wr[0] - slmul * wl[0]);
  double uu[8], uuk[8];
  if (sst >= 0.) {
    double rhost = wl[0] * (slmul) / (sl - sst);
    double ek = 0.5 * wl[0] * (wl[1] * wl[1] + wl[2] * wl[2] + wl[3] * wl[3]) + P.cv * wl[4];
    uuk[0] = rhost;
    uuk[1] = rhost * sst;
    uuk[2] = rhost * wl[2];
    uuk[3] = rhost * wl[3];
    uuk[4] = rhost * (ek / wl[0] + (sst - wl[1]) * (sst + wl[4] / (wl[0] * slmul)));
    prim2f<false>(P, wl, ff);
    prim2u<false>(P, wl, uu);
#pragma unroll
    for (int q = 0; q < 5; ++q) ff[q] = ff[q] + sl * (uuk[q] - uu[q]);
    I.mode = PAS_HLLC_L; I.a = rhost; I.b = wl[0];
    return 0;
  }
  if (sst <= 0.) {
    double rhost = wr[0] * (srmur) / (sr - sst);
    double ek = 0.5 * wr[0] * (wr[1] * wr[1] + wr[2] * wr[2] + wr[3] * wr[3]) + P.cv * wr[4];
    uuk[0] = rhost;
    uuk[1] = rhost * sst;
    uuk[2] = rhost * wr[2];
    uuk[3] = rhost * wr[3];
    uuk[4] = rhost * (ek / wr[0] + (sst - wr[1]) * (sst + wr[4] / (wr[0] * srmur)));
    prim2f<false>(P, wr, ff);
    prim2u<false>(P, wr, uu);
#pragma unroll
    for (int q = 0; q < 5; ++q) ff[q] = ff[q] + sr * (uuk[q] - uu[q]);
    I.mode = PAS_HLLC_R; I.a = rhost; I.b = wr[0];
    return 0;
  }
  return 1;   // NaN: the reference prints 'Error in hllc' and stops (hllc.f90:135-138)
}

#if defined(GX_FLAVOUR_FAST)
// Production HLLC: the same expressions as riemann_hllc_ref / src/hllc.f90:44-140 in one basic block — the side K
// that supplies the star state (L for S* >= 0, R otherwise) is chosen by selects, the divisions become
// Newton-refined reciprocals, and the two supersonic cases are a rare override at the end.
__device__ __forceinline__ int riemann_hllc(const Phys& P, const double (&wl)[8], const double (&wr)[8], double (&ff)[8], PasInfo& I,
                                            double csl, double csr) {
  const double sr = gx_max(wl[1] + csl, wr[1] + csr);
  const double sl = gx_min(wl[1] - csl, wr[1] - csr);
  I.ul = wl[1]; I.ur = wr[1]; I.sl = sl; I.sr = sr;
  const double slmul = sl - wl[1], srmur = sr - wr[1];
  const double mL = wl[0] * slmul, mR = wr[0] * srmur;
  const double sst = (mR * wr[1] - mL * wl[1] - wr[4] + wl[4]) * fast_rcp(mR - mL);          // hllc.f90:76-77
  const bool left = sst >= 0.;
  const int err = (!left && !(sst <= 0.)) ? 1 : 0;                                               // NaN: 'Error in hllc' + stop (:135-138)
  const double q0 = left ? wl[0] : wr[0], q1 = left ? wl[1] : wr[1], q2 = left ? wl[2] : wr[2], q3 = left ? wl[3] : wr[3], q4 = left ? wl[4] : wr[4];
  const double sK = left ? sl : sr, mK = left ? mL : mR;
  const double rhost = mK * fast_rcp(sK - sst);                                                  // :80, :108
  const double ek = 0.5 * q0 * (q1 * q1 + q2 * q2 + q3 * q3) + P.cv * q4;
  // ek / rho_K + (S* - u_K)(S* + p_K / m_K) with ONE reciprocal: 1/(rho_K m_K) serves both quotients
  const double irm = fast_rcp(q0 * mK);
  const double uk4 = rhost * ((ek * mK) * irm + (sst - q1) * (sst + (q4 * q0) * irm));           // :86-87
  const double m1 = q0 * q1;                                                                     // prim2f / prim2u of side K
  ff[0] = m1 + sK * (rhost - q0);
  ff[1] = (m1 * q1 + q4) + sK * (rhost * sst - m1);
  ff[2] = m1 * q2 + sK * (rhost * q2 - q0 * q2);
  ff[3] = m1 * q3 + sK * (rhost * q3 - q0 * q3);
  ff[4] = q1 * (ek + q4) + sK * (uk4 - ek);
  I.mode = left ? PAS_HLLC_L : PAS_HLLC_R; I.a = rhost; I.b = q0;
  if (GX_ANY_SUPERSONIC(sl, sr)) {
    if (sl > 0) { prim2f<false>(P, wl, ff); I.mode = PAS_UPL; return 0; }
    if (sr < 0) { prim2f<false>(P, wr, ff); I.mode = PAS_UPR; return 0; }
  }
  return err;
}
#else
__device__ __forceinline__ int riemann_hllc(const Phys& P, const double (&wl)[8], const double (&wr)[8], double (&ff)[8], PasInfo& I,
                                            double csl, double csr) {
  return riemann_hllc_ref(P, wl, wr, ff, I, csl, csr);
}
#endif

// ---- HLLD (Miyoshi & Kusano 2005 five-wave): src/hlld.f90:48-319 ----
// One-sided star state (used for both sides with the roles of L/R exchanged).
struct Star { double v, w, by, bz; };
__device__ __forceinline__ Star hlld_star(const double (&q)[8], double bx, double smu /*S_K - u_K*/, double sms /*S_K - S_M*/, double sMmu /*S_M - u_K*/) {
  Star s;
  double den = q[0] * smu * sms - bx * bx;
  if (den == 0) {                      // hlld.f90:119-126 degenerate guard
    s.v = q[2]; s.w = q[3]; s.by = 0.; s.bz = 0.;
  } else {
    const Den dn(den);
    s.v = q[2] - dn.div(bx * q[6] * sMmu);
    s.w = q[3] - dn.div(bx * q[7] * sMmu);
    double num = q[0] * (smu * smu) - bx * bx;
    s.by = dn.div(q[6] * num);
    s.bz = dn.div(q[7] * num);
  }
  return s;
}
// total energy of side K with the interface-averaged Bx (hlld.f90:113-114)
__device__ __forceinline__ double hlld_energy(const Phys& P, const double (&q)[8], double bx) {
  return 0.5 * q[0] * (q[1] * q[1] + q[2] * q[2] + q[3] * q[3]) + P.cv * q[4] + 0.5 * (bx * bx + q[6] * q[6] + q[7] * q[7]);
}

__device__ __forceinline__ int riemann_hlld_ref(const Phys& P, const double (&wl)[8], const double (&wr)[8], double (&ff)[8], PasInfo& I,
                                                double csl, double csr) {   // csl, csr = cfastX of either side (hlld.f90:64-65)
  double sr = gx_max(wl[1] + csl, wr[1] + csr);
  double sl = gx_min(wl[1] - csl, wr[1] - csr);
  I.ul = wl[1]; I.ur = wr[1]; I.sl = sl; I.sr = sr;
  if (sl > 0) { prim2f<true>(P, wl, ff); I.mode = PAS_UPL; return 0; }
  if (sr < 0) { prim2f<true>(P, wr, ff); I.mode = PAS_UPR; return 0; }

  double bx = 0.5 * (wl[5] + wr[5]);
  double signBx = sign1(bx);
  double pTL = wl[4] + 0.5 * (bx * bx + wl[6] * wl[6] + wl[7] * wl[7]);
  double pTR = wr[4] + 0.5 * (bx * bx + wr[6] * wr[6] + wr[7] * wr[7]);
  double slmul = sl - wl[1];
  double srmur = sr - wr[1];
  double rholul = wl[0] * wl[1];
  double rhorur = wr[0] * wr[1];
  const Den den(srmur * wr[0] - slmul * wl[0]);
  double sM = den.div(srmur * rhorur - slmul * rholul - pTR + pTL);
  double srmsM = sr - sM;
  double slmsM = sl - sM;
  const Den dslm(slmsM), dsrm(srmsM);
  double rhostl = dslm.div(wl[0] * slmul);
  double rhostr = dsrm.div(wr[0] * srmur);
  const SqrtDen SQL(rhostl), SQR(rhostr);
  const double sql = SQL.s, sqr = SQR.s;
  double sstl = sM - SQL.div(fabs(bx));
  double sstr = sM + SQR.div(fabs(bx));
  double pst = den.div(srmur * wr[0] * pTL - slmul * wl[0] * pTR + wl[0] * wr[0] * srmur * slmul * (wr[1] - wl[1]));

  // Which side supplies the outer state: L for UL*, UL**; R for UR*, UR**.
  // Region order as in the reference: sstl>=0, sstr<=0, sM>=0, sM<=0.
  const bool starL = (sstl >= 0), starR = !starL && (sstr <= 0);
  const bool dstar = !starL && !starR;
  if (dstar && !(sM >= 0) && !(sM <= 0)) return 1;          // NaN: 'Error in HLLD routine' + stop
  const bool left = starL || (dstar && sM >= 0);

  Star SL, SR;
  if (left || dstar) SL = hlld_star(wl, bx, slmul, slmsM, sM - wl[1]);
  if (!left || dstar) SR = hlld_star(wr, bx, srmur, srmsM, sM - wr[1]);

  double vs, ws, bys, bzs, vdb_ss = 0.;     // state used in the flux (star or double-star)
  if (dstar) {
    const Den dd(sql + sqr);
    double sq2 = sqrt_prod(SQL, SQR, rhostl * rhostr);
    vs = dd.div(sql * SL.v + sqr * SR.v + (SR.by - SL.by) * signBx);
    ws = dd.div(sql * SL.w + sqr * SR.w + (SR.bz - SL.bz) * signBx);
    bys = dd.div(sql * SR.by + sqr * SL.by + sq2 * (SR.v - SL.v) * signBx);
    bzs = dd.div(sql * SR.bz + sqr * SL.bz + sq2 * (SR.w - SL.w) * signBx);
    vdb_ss = sM * bx + vs * bys + ws * bzs;
  }

  double rhost, es;
  if (left) {
    double el = hlld_energy(P, wl, bx);
    double vdotb = wl[1] * bx + wl[2] * wl[6] + wl[3] * wl[7];
    double vsdotbs = sM * bx + SL.v * SL.by + SL.w * SL.bz;
    double estl = dslm.div(slmul * el - pTL * wl[1] + pst * sM + bx * (vdotb - vsdotbs));
    rhost = rhostl;
    if (dstar) { es = estl - sql * (vsdotbs - vdb_ss) * signBx; }
    else { es = estl; vs = SL.v; ws = SL.w; bys = SL.by; bzs = SL.bz; vdb_ss = vsdotbs; }
    I.mode = PAS_HLLD_L; I.a = sM; I.b = slmul; I.c = slmsM;
  } else {
    double er = hlld_energy(P, wr, bx);
    double vdotb = wr[1] * bx + wr[2] * wr[6] + wr[3] * wr[7];
    double vsdotbs = sM * bx + SR.v * SR.by + SR.w * SR.bz;
    double estr = dsrm.div(srmur * er - pTR * wr[1] + pst * sM + bx * (vdotb - vsdotbs));
    rhost = rhostr;
    if (dstar) { es = estr + sqr * (vsdotbs - vdb_ss) * signBx; }
    else { es = estr; vs = SR.v; ws = SR.w; bys = SR.by; bzs = SR.bz; vdb_ss = vsdotbs; }
    I.mode = PAS_HLLD_R; I.a = sM; I.b = srmur; I.c = srmsM;
  }
  ff[0] = rhost * sM;
  ff[1] = rhost * (sM * sM) + pst - bx * bx;
  ff[2] = rhost * sM * vs - bx * bys;
  ff[3] = rhost * sM * ws - bx * bzs;
  ff[4] = sM * (es + pst) - bx * (vdb_ss);
  ff[5] = 0.;
  ff[6] = bys * sM - bx * vs;
  ff[7] = bzs * sM - bx * ws;
  return 0;
}

#if defined(GX_FLAVOUR_FAST)
// x * sign(1, s) without touching the FP64 pipe: flip the sign bit of x by the sign bit of s
__device__ __forceinline__ double mul_sign(double x, double s) {
  return __hiloint2double(__double2hiint(x) ^ (__double2hiint(s) & 0x80000000), __double2loint(x));
}
// Straight-line (select-based) HLLD for the production build: the same quantities as riemann_hlld_ref /
// src/hlld.f90:48-319 — both star states and the double-star blend are always formed and the region
// (UL*, UL**, UR**, UR*) is chosen by selects, so the whole solve is one basic block whose independent L/R chains
// ptxas interleaves; the two supersonic cases are a rare override at the end.  Division- and square-root-free:
//   * rho*_K = m_K / (S_K - S_M), m_K = rho_K (S_K - u_K).  With x_K = m_K (S_K - S_M) > 0 and r_K = 1/sqrt(x_K):
//     sqrt(rho*_K) = |m_K| r_K,  1/sqrt(rho*_K) = |S_K - S_M| r_K,  rho*_K = (|m_K| r_K)^2, and x_K - Bx^2 is the
//     denominator of the star state (hlld.f90:117, 165) — one reciprocal square root per side instead of a
//     reciprocal plus a square root;
//   * 1/(S_K - S_M) is only needed for the energy of the side that supplies the flux: one reciprocal after the select.
__device__ __forceinline__ int riemann_hlld(const Phys& P, const double (&wl)[8], const double (&wr)[8], double (&ff)[8], PasInfo& I,
                                            double csl, double csr, double bt2L, double bt2R) {   // bt2_K = By_K^2 + Bz_K^2
  const double sr = gx_max(wl[1] + csl, wr[1] + csr);
  const double sl = gx_min(wl[1] - csl, wr[1] - csr);
  I.ul = wl[1]; I.ur = wr[1]; I.sl = sl; I.sr = sr;

  const double bx = 0.5 * (wl[5] + wr[5]);
  const double bx2 = bx * bx;
  const double pTL = fma(0.5, bx2 + bt2L, wl[4]), pTR = fma(0.5, bx2 + bt2R, wr[4]);
  const double slmul = sl - wl[1], srmur = sr - wr[1];
  const double mL = wl[0] * slmul, mR = wr[0] * srmur;            // rho_K (S_K - u_K)
  const double iden = fast_rcp(mR - mL);
  const double sM = (mR * wr[1] - mL * wl[1] - pTR + pTL) * iden;
  const double slmsM = sl - sM, srmsM = sr - sM;
  const double xL = mL * slmsM, xR = mR * srmsM;                  // rho*_K (S_K - S_M)^2
  const double rL = fast_rsqrt(xL), rR = fast_rsqrt(xR);
  const double sql = fabs(mL) * rL, sqr = fabs(mR) * rR;          // sqrt(rho*_K)
  const double abx = fabs(bx);
  const double sstl = fma(-(abx * fabs(slmsM)), rL, sM), sstr = fma(abx * fabs(srmsM), rR, sM);   // S*_K = S_M -+ |Bx| / sqrt(rho*_K)
  const double pst = (mR * pTL - mL * pTR + mL * mR * (wr[1] - wl[1])) * iden;

  // star states of both sides (hlld.f90:116-135, 164-183), degenerate guard by select
  const double dL = xL - bx2, dR = xR - bx2;
  const double rL_ = fast_rcp(dL), rR_ = fast_rcp(dR);
  // hlld.f90:119-126: den == 0 -> v* = v, w* = w, By* = Bz* = 0, which is what 1/den := 0 produces below
  const double idL = (dL != 0.0) ? rL_ : 0.0, idR = (dR != 0.0) ? rR_ : 0.0;
  const double cL = bx * (sM - wl[1]) * idL, cR = bx * (sM - wr[1]) * idR;          // Bx (S_M - u_K) / den_K
  const double nL = fma(mL, slmul, -bx2) * idL, nR = fma(mR, srmur, -bx2) * idR;    // (rho_K (S_K - u_K)^2 - Bx^2) / den_K
  const double vL = fma(-wl[6], cL, wl[2]), wLs = fma(-wl[7], cL, wl[3]);
  const double vR = fma(-wr[6], cR, wr[2]), wRs = fma(-wr[7], cR, wr[3]);
  const double byL = wl[6] * nL, bzL = wl[7] * nL;
  const double byR = wr[6] * nR, bzR = wr[7] * nR;

  // region: hlld.f90 tests sstl>=0, sstr<=0, sM>=0, sM<=0 in this order
  const bool starL = sstl >= 0.0, starR = !starL && (sstr <= 0.0);
  const bool dstar = !(starL || starR);
  const bool left = starL || (dstar && sM >= 0.0);
  const int err = (dstar && !(sM >= 0.0) && !(sM <= 0.0)) ? 1 : 0;     // NaN: 'Error in HLLD routine' + stop

  // double-star state (hlld.f90:207-253): formed for the whole warp when any of its interfaces lies between the
  // two Alfven waves, skipped otherwise (faces with Bn = 0, super-Alfvenic flow)
  double vss = 0.0, wss = 0.0, byss = 0.0, bzss = 0.0, vdb_ss = 0.0;
  if (GX_ANY_LANE(dstar)) {
    const double idd = fast_rcp(sql + sqr), sq2 = sql * sqr;
    vss = (sql * vL + sqr * vR + mul_sign(byR - byL, bx)) * idd;
    wss = (sql * wLs + sqr * wRs + mul_sign(bzR - bzL, bx)) * idd;
    byss = (sql * byR + sqr * byL + sq2 * mul_sign(vR - vL, bx)) * idd;
    bzss = (sql * bzR + sqr * bzL + sq2 * mul_sign(wRs - wLs, bx)) * idd;
    vdb_ss = sM * bx + vss * byss + wss * bzss;
  }

  // outer state K = L | R by select, then one energy evaluation (hlld.f90:113-114, 137-156)
  const double q0 = left ? wl[0] : wr[0], q1 = left ? wl[1] : wr[1], q2 = left ? wl[2] : wr[2], q3 = left ? wl[3] : wr[3];
  const double q4 = left ? wl[4] : wr[4], q6 = left ? wl[6] : wr[6], q7 = left ? wl[7] : wr[7], bt2K = left ? bt2L : bt2R;
  const double sKmu = left ? slmul : srmur, sKmsM = left ? slmsM : srmsM, pTK = left ? pTL : pTR;
  const double vK = left ? vL : vR, wK = left ? wLs : wRs, byK = left ? byL : byR, bzK = left ? bzL : bzR;
  const double sqK = left ? -sql : sqr;
  const double iK = fast_rcp(sKmsM);
  const double eK = 0.5 * q0 * (q1 * q1 + q2 * q2 + q3 * q3) + P.cv * q4 + 0.5 * (bx2 + bt2K);
  const double vdotb = q1 * bx + q2 * q6 + q3 * q7;
  const double vsdotbs = sM * bx + vK * byK + wK * bzK;
  const double estK = (sKmu * eK - pTK * q1 + pst * sM + bx * (vdotb - vsdotbs)) * iK;
  const double es = dstar ? estK + sqK * mul_sign(vsdotbs - vdb_ss, bx) : estK;
  const double vs = dstar ? vss : vK, ws = dstar ? wss : wK, bys = dstar ? byss : byK, bzs = dstar ? bzss : bzK;
  const double vdb = dstar ? vdb_ss : vsdotbs;
  I.mode = left ? PAS_HLLD_L : PAS_HLLD_R; I.a = sM; I.b = sKmu; I.c = sKmsM;

  const double rsm = (sqK * sqK) * sM;                              // rho*_K S_M
  ff[0] = rsm;
  ff[1] = rsm * sM + pst - bx2;
  ff[2] = rsm * vs - bx * bys;
  ff[3] = rsm * ws - bx * bzs;
  ff[4] = sM * (es + pst) - bx * vdb;
  ff[5] = 0.;
  ff[6] = bys * sM - bx * vs;
  ff[7] = bzs * sM - bx * ws;
  // supersonic interfaces are rare: one warp-uniform test keeps the override and its
  // reconvergence bookkeeping out of the common path
  if (GX_ANY_SUPERSONIC(sl, sr)) {
    if (sl > 0.0) { prim2f<true>(P, wl, ff); I.mode = PAS_UPL; return 0; }
    if (sr < 0.0) { prim2f<true>(P, wr, ff); I.mode = PAS_UPR; return 0; }
  }
  return err;
}
#endif

// Riemann flux of solver SOLVER.  PRE: csl / csr hold the signal speeds of the two states (first-order stage, cell_speeds);
// otherwise they are evaluated here from wl / wr.
template <int SOLVER, bool PRE = false>
__device__ __forceinline__ int riemann(const Phys& P, const double (&wl)[8], const double (&wr)[8], double (&ff)[8], PasInfo& I,
                                       double csl = 0.0, double csr = 0.0) {
  constexpr bool FASTSPEED = (SOLVER == GX_SOLVER_HLLE || SOLVER == GX_SOLVER_HLLD);
#if defined(GX_FLAVOUR_FAST)
  if (SOLVER == GX_SOLVER_HLLD) {
    const double bt2L = fma(wl[7], wl[7], wl[6] * wl[6]), bt2R = fma(wr[7], wr[7], wr[6] * wr[6]);
    if (!PRE) { csl = cfast_dir(P, wl[0], wl[4], wl[5], bt2L); csr = cfast_dir(P, wr[0], wr[4], wr[5], bt2R); }
    return riemann_hlld(P, wl, wr, ff, I, csl, csr, bt2L, bt2R);
  }
#endif
  if (!PRE) {
    if (FASTSPEED) { csl = cfastX(P, wl); csr = cfastX(P, wr); }
    else { csl = csound(P, wl[4], wl[0]); csr = csound(P, wr[4], wr[0]); }
  }
  if (SOLVER == GX_SOLVER_HLL) return riemann_hll<false>(P, wl, wr, ff, I, csl, csr);
  if (SOLVER == GX_SOLVER_HLLE) return riemann_hll<true>(P, wl, wr, ff, I, csl, csr);
  if (SOLVER == GX_SOLVER_HLLC) return riemann_hllc(P, wl, wr, ff, I, csl, csr);
#if defined(GX_FLAVOUR_FAST)
  return 1;
#else
  return riemann_hlld_ref(P, wl, wr, ff, I, csl, csr);
#endif
}

}  // namespace gxp
