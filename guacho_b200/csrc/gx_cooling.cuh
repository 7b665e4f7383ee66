// gx_cooling.cuh — COOL_H (src/cooling_h.f90): cell-local operator on u after viscous_copy (hydro_solver.f90:202-204).
// Included by gx_kernels.cu INSIDE its flavour namespace.  Bit-comparison build (-fmad=false, IEEE div/sqrt): the expressions
// are evaluated in the reference's order; exp/log/pow are CUDA's FP64 routines (<= 2 ulp).  Production build: the operator is
// bound by FP64 transcendentals (seven pow, seven exp, a log and five square roots per cell: 1.54 ms per step for EXO's 16 M
// cells), so the general powers become exp(y log x) (relative error <= |y ln x| ulp ~ 5e-15 for the temperatures involved),
// 10**x becomes exp10 and x**0.5 a square root.  The `d`-exponent literals of the reference are quad under -fdefault-real-8
// (see oracle/guacho_oracle.cpp); FP64 here.
#if defined(GX_FLAVOUR_FAST)
#define GX_POW(x, y) exp((y) * log(x))
#define GX_POW10(x) exp10(x)
#define GX_POWHALF(x) sqrt(x)
#else
#define GX_POW(x, y) pow((x), (y))
#define GX_POW10(x) pow(10., (x))
#define GX_POWHALF(x) pow((x), 0.5)
#endif
__device__ __forceinline__ double cool_alpha(double T) { return 2.55e-13 * GX_POW(1.e4 / T, 0.79); }            // :75-84
__device__ __forceinline__ double cool_colf(double T) { return 5.83e-11 * sqrt(T) * exp(-157828. / T); }      // :109-118
__device__ __forceinline__ double cool_betah(double T) {                                                     // :126-137
  const double a = 157890. / T;
  return 1.133e-24 / sqrt(a) * (-0.0713 + 0.5 * log(a) + 0.640 * GX_POW(a, -0.33333));
}
__device__ double cool_aloss(double X1, double DEN, double DH0, double TE0) {                               // :159-246
  const double XION = 2.179e-11, XO = 1.e-3;
  const double C0 = 0.5732, C1 = 1.8288e-5, C2 = -1.15822e-10, C3 = 9.4288e-16;
  const double D0 = 0.5856, D1 = 1.55083e-5, D2 = -9.669e-12, D3 = 5.716e-19;
  const double ENK = 118409., EN = 1.634E-11;
  const double TE = fmax(TE0, 10.);
  const double DH = DEN;
  const double DHP = (1. - X1) * DH;
  const double DE = DHP + 1.E-4 * DH;
  const double DOI = XO * DH0;
  const double DOII = XO * DHP;
  if (TE <= 1e4) return 0.;
  double OMEGA = 0.;
  if (TE <= 55000.) OMEGA = C0 + TE * (C1 + TE * (C2 + TE * C3));
  if (TE >= 72000.) OMEGA = D0 + TE * (D1 + TE * (D2 + TE * D3));
  if (TE > 55000. && TE < 72000.) {
    const double OMEGAL = C0 + TE * (C1 + TE * (C2 + TE * C3));
    const double OMEGAH = D0 + TE * (D1 + TE * (D2 + TE * D3));
    const double FRAC = (TE - 55000.) / 17000.;
    OMEGA = (1. - FRAC) * OMEGAL + FRAC * OMEGAH;
  }
  const double QLA = 8.6287E-6 / (2. * sqrt(TE)) * OMEGA * exp(-ENK / TE);
  double ECOLL = DE * DH0 * QLA * EN;
  ECOLL = fmax(ECOLL, 0.);
  const double CION = 5.834E-11 * sqrt(TE) * exp(-1.579E5 / TE);
  const double EION = DE * DH0 * CION * XION;
  double EREC = DE * DHP * (cool_betah(TE));
  EREC = fmax(EREC, 0.);
  const double TM = 1. / TE;
  const double T2 = TM * TM;
  double EOI = DE * DOI * GX_POW10(1381465 * T2 - 12328.69 * TM - 19.82621);
  double EOII = DE * DOII * GX_POW10(-2061075. * T2 - 14596.24 * TM - 19.01402);
  EOI = fmax(EOI, 0.);
  EOII = fmax(EOII, 0.);
  const double BETAF = 1.3 * 1.42E-27 * GX_POWHALF(TE);
  const double HIICOOL = DE * DHP * BETAF;
  const double EQUIL = (1.0455E-18 / GX_POW(TE, 0.63)) * (1. - exp(-GX_POW(TE * 1.E-5, 1.63))) * DE * DEN + HIICOOL;
  double FR = 0.;
  if (TE >= 54770.) FR = 1.;
  if (TE > 44770. && TE < 54770.) {
    const double EX2 = exp(-2. * (TE - 49770.) / 500.);
    const double TANH = (1. - EX2) / (1. + EX2);
    FR = 0.5 * (1. + TANH);
  }
  return ECOLL + EION + (EREC + 7.033 * (EOI + EOII)) * (1. - FR) + EQUIL * FR;
}
// atomic(dt,uu,tau,radphi), dif_rad = .false. (:259-371) on one cell: uu = the dynamic variables, un = the neutral H density
// uu(neqdyn+1); returns the new energy uu(5) and neutral density
__device__ __forceinline__ void cool_cell(const gxp::Phys& P, int mhd, double dt, const double (&uu)[8], double un, double& e5_out, double& un_out) {
  const double xi = 1.e-4, boltzm = 1.3807e-16;
  double prim[8], T;
  if (mhd) gxp::u2prim<true, true>(P, uu, prim, un, T); else gxp::u2prim<false, true>(P, uu, prim, un, T);
  const double col = cool_colf(T);
  const double rec = cool_alpha(T);
  const double y0 = un / uu[0];
  const double dh = uu[0];
  const double a = rec + col;
  const double b = -((2. + xi) * rec + (1. + xi) * col);
  const double cc = (1. + xi) * rec;
  const double d = sqrt(b * b - 4. * a * cc);
  const double g0 = (2. * a * y0 + b + d) / (2. * a * y0 + b - d);
  const double e = exp(-d * dh * dt);
  double y1 = (-b - d * (1. + g0 * e) / (1. - g0 * e)) / (2. * a);
  y1 = fmin(y1, 0.9999);
  y1 = fmax(y1, 0.);
  const double al = cool_aloss(y0, dh, un, T) / (dh * dh);
  const double tprime = 10.;
  const double ce = (2. * dh * al) / (3. * boltzm * T);
  double t1 = tprime + (T - tprime) * exp(-ce * dt);
  t1 = fmax(t1, 0.1 * T);
  t1 = fmin(t1, 10. * T);
  const double un1 = y1 * uu[0];
  double e5 = P.cv * (2. * uu[0] - un1) * t1 / P.Tempsc + 0.5 * prim[0] * (prim[1] * prim[1] + prim[2] * prim[2] + prim[3] * prim[3]);
  if (mhd) e5 = e5 + 0.5 * (prim[5] * prim[5] + prim[6] * prim[6] + prim[7] * prim[7]);
  e5_out = e5; un_out = un1;
}
// coolingh (:41-67): atomic over the physical cells
__global__ void __launch_bounds__(128) k_coolingh(Grid g, gxp::Phys P, int mhd, double dt, double* __restrict__ U) {
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x) + 1, j = (int)blockIdx.y + 1, k = (int)blockIdx.z + 1;
  if (i > g.nx) return;
  const long long c = g.idx(i, j, k), vs = g.vs;
  double uu[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) uu[q] = (q < g.neqdyn) ? U[q * vs + c] : 0.0;
  const double un = U[(long long)g.neqdyn * vs + c];            // neutral H density, uu(neqdyn+1)
  double e5, un1;
  cool_cell(P, mhd, dt, uu, un, e5, un1);
  U[(long long)g.neqdyn * vs + c] = un1;
  U[4 * vs + c] = e5;
}
// (Measured and rejected: coolingh as the epilogue of the viscous_copy stencil pass — 2.12 ms against 0.59 + 1.01 ms for the two
// kernels at EXO's 400x100x400: 70 stencil loads in front of a long dependent transcendental chain at 72 registers hide nothing.)
#undef GX_POW
#undef GX_POW10
#undef GX_POWHALF
