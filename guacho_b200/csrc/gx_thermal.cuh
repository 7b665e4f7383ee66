// gx_thermal.cuh — thermal conduction (src/thermal_cond.f90), the operator the reference applies at the end of tstep
// (hydro_solver.f90:227).  Included by gx_kernels.cu INSIDE its flavour namespace, so it is compiled twice like every other
// kernel: the bit-comparison build (-fmad=false, IEEE div/sqrt, pow(): the reference's expressions in the reference's order)
// and the production build (FMA, shared reciprocals and MUFU-seeded square roots of gx_physics.cuh, T**2.5 as T*T*sqrt(T)).
//
// Reference structure per substep: heatfluxes / MHD_heatfluxes fill f,g,h(5,...) over 0..n from primit and Temp, u(5) is
// updated over the physical cells, thermal_bounds exchanges one layer of u(5), calcprim(u) refreshes primit and Temp.
// Here: no flux arrays.  A substep is ONE kernel over the physical cells that evaluates the (up to) six face fluxes of its
// cell from a two-variable scratch array (pressure, temperature; the density and B come straight from u) and updates u(5)
// in place (the fluxes do not read u(5)), followed by the ghost layer of u(5) and k_tc_prim (the calcprim of the reference,
// reduced to the two variables the operator reads).
// (no include guard: one inclusion per flavour namespace)


// Production build: T**2.5 as T*T*sqrt(T) — three roundings instead of pow()'s ~150 FP64 instructions (six conductivities per
// cell and substep made the operator FP64-bound: 1.66 ms per substep at 256^3).
__device__ __forceinline__ double pow25(double T) {
#if defined(GX_FLAVOUR_FAST)
  return T * T * gxp::gx_sqrt(T);
#else
  return pow(T, 2.5);
#endif
}
__device__ __forceinline__ double Ksp(double T) { return 6.e-7 * pow25(T); }                    // thermal_cond.f90:142-149
__device__ __forceinline__ double Ksp_parl(double T) { return 9.2181e-7 * pow25(T); }           // :157-164
__device__ __forceinline__ double Ksp_perp(double T, double dens, double B2) { return gxp::Den(B2 * gxp::gx_sqrt(T)).div(0.30089e+33 * dens) * dens; }   // :172-178

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
    if (want_dt && i >= 1 && i <= g.nx && j >= 1 && j <= g.ny && k >= 1 && k <= g.nz) cand = gxp::Den(Ksp(T)).div(prim[0]);
  }
  if (want_dt) {                                      // positive doubles order like their bit patterns
    for (int o = 16; o > 0; o >>= 1) cand = fmin(cand, __shfl_xor_sync(0xffffffffu, cand, o));
    // half a million warps on ONE word serialise (0.78 ms against 0.24 for the same pass without the reduction): the candidates go
    // to 256 words on separate 128-byte lines, chosen by the row of the CTA; k_tc_slots_min folds them into the result
    if ((threadIdx.x & 31) == 0 && cand < 1.7976931348623157e308)
      atomicMin(dt_bits + 16 * (((unsigned)blockIdx.z * gridDim.y + blockIdx.y) & 255u), (unsigned long long)__double_as_longlong(cand));
  }
}

__global__ void k_tc_slots_min(const unsigned long long* __restrict__ slots, unsigned long long* __restrict__ out) {   // <<<1, 256>>>
  unsigned long long v = slots[16 * threadIdx.x];
  for (int o = 16; o > 0; o >>= 1) { const unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o); v = w < v ? w : v; }
  __shared__ unsigned long long sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) { for (int w = 1; w < 8; ++w) v = sm[w] < v ? sm[w] : v; *out = v; }
}

// heatfluxes (:189-267): flux through the face between cell a (low side) and cell b
__device__ __forceinline__ double flux_iso(const gx::TcPar& t, const gxp::Phys& P, double Ta, double Tb, double pa, double pb, double ra, double rb, double dxr, double idxr) {
  if (Ta == Tb) return 0.;
  const double meanP = 0.5 * (pa + pb);
  const double meanDens = 0.5 * (ra + rb);
  const double meanT = 0.5 * (Ta + Tb);
#if defined(GX_FLAVOUR_FAST)
  const double dT = (Tb - Ta) * idxr;
#else
  const double dT = (Tb - Ta) / dxr;
#endif
  double coef;
  if (t.sat) {
    double cs = gxp::csound(P, meanP, meanDens);
    cs = fmin(cs * t.sqrt_vsc2, 3.E10);
    coef = fmin(Ksp(meanT), gxp::Den(fabs(dT)).div(5. * 0.4 * cs * meanP * t.Psc));
  } else coef = Ksp(meanT);
  return -coef * dT * 1.;
}

// MHD_heatfluxes (:277-487): component D of the flux triplet the reference stores at cell c — B of cell c, forward
// temperature differences of cell c in all three directions
template <int D>
__device__ __forceinline__ double flux_aniso(const gx::TcPar& t, const gxp::Phys& P, const gx::Grid& g, const double* __restrict__ U, const double* __restrict__ PT, long long c) {
  const long long vs = g.vs;
  const long long nb[3] = {c + 1, c + g.px, c + (long long)g.px * g.py};
  const double dr[3] = {t.dxr, t.dyr, t.dzr}, idr[3] = {t.idxr, t.idyr, t.idzr};
  double bx = U[5 * vs + c], by = U[6 * vs + c], bz = U[7 * vs + c];
  const double B2 = bx * bx + by * by + bz * bz;
  const gxp::SqrtDen modB(B2);
  bx = modB.div(bx); by = modB.div(by); bz = modB.div(bz);
  const double Tc = PT[vs + c], rc = fmax(U[c], 1e-15), pc = PT[c];
  double grad[3], Kparl = 0.0, Kperp = 0.0, coefSat = 0.0;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double Tn = PT[vs + nb[d]];
    if (fabs(Tc - Tn) < 1.0e-14) { grad[d] = 0.0; continue; }
#if defined(GX_FLAVOUR_FAST)
    grad[d] = (Tn - Tc) * idr[d];
#else
    grad[d] = (Tn - Tc) / dr[d];
#endif
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
  const double gradT_perp = gxp::gx_sqrt_disc(perp[0] * perp[0] + perp[1] * perp[1] + perp[2] * perp[2]);
  const gxp::Den dsat(coefSat + 1.e-14);
  return -gxp::Den(gxp::Den(Kparl + 1.e-14).div(1.) + dsat.div(gradT_parl)).div(1.) * parl[D]
         - gxp::Den(gxp::Den(Kperp + 1.e-14).div(1.) + dsat.div(gradT_perp)).div(1.) * perp[D];
}

// one substep over the physical cells (:749-757), fluxes evaluated in place of f,g,h(5,...)
// FILL: the block owns the whole domain — every face gets thermal_bounds' zero-gradient copy (:589-614), which the thread of
// a boundary cell writes itself (the six sequential plane copies leave every ghost cell of the layer with the value of the
// physical cell nearest to it: faces, edges and corners alike), so a substep needs no boundary launch at all.
#ifndef GX_TC_MINB                   // resident CTAs per SM the substep kernel is compiled for: it is latency-bound (21 scattered loads, then
#define GX_TC_MINB 8                 // dependent div / sqrt chains); 64 registers with a few spills ran 1.7x faster than 96 without (256^3)
#endif
template <bool FILL>
__global__ void __launch_bounds__(128, GX_TC_MINB) k_tc_update(gx::Grid g, gxp::Phys P, gx::TcPar t, const double* __restrict__ PT, double* __restrict__ U, double dts) {
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x) + 1, j = (int)blockIdx.y + 1, k = (int)blockIdx.z + 1;
  if (i > g.nx) return;
  const long long c = g.idx(i, j, k), vs = g.vs, sy = g.px, sz = (long long)g.px * g.py;
  double fc, fm, gc, gm, hc, hm;
  if (t.mode == GX_TC_ISOTROPIC) {
    const double Tc = PT[vs + c], pc = PT[c], rc = fmax(U[c], 1e-15);
    auto lo = [&](long long n, double dxr, double idxr) { return flux_iso(t, P, PT[vs + n], Tc, PT[n], pc, fmax(U[n], 1e-15), rc, dxr, idxr); };
    auto hi = [&](long long n, double dxr, double idxr) { return flux_iso(t, P, Tc, PT[vs + n], pc, PT[n], rc, fmax(U[n], 1e-15), dxr, idxr); };
    fc = hi(c + 1, t.dxr, t.idxr); fm = lo(c - 1, t.dxr, t.idxr);
    gc = hi(c + sy, t.dyr, t.idyr); gm = lo(c - sy, t.dyr, t.idyr);
    hc = hi(c + sz, t.dzr, t.idzr); hm = lo(c - sz, t.dzr, t.idzr);
  } else {
    fc = flux_aniso<0>(t, P, g, U, PT, c); fm = flux_aniso<0>(t, P, g, U, PT, c - 1);
    gc = flux_aniso<1>(t, P, g, U, PT, c); gm = flux_aniso<1>(t, P, g, U, PT, c - sy);
    hc = flux_aniso<2>(t, P, g, U, PT, c); hm = flux_aniso<2>(t, P, g, U, PT, c - sz);
  }
#if defined(GX_FLAVOUR_FAST)
  const double v = U[4 * vs + c] - dts * ((fc - fm) * t.idx + (gc - gm) * t.idy + (hc - hm) * t.idz);
#else
  const double v = U[4 * vs + c] - dts * ((fc - fm) / t.dx + (gc - gm) / t.dy + (hc - hm) / t.dz);
#endif
  double* const E5 = U + 4 * vs;
  E5[c] = v;
  if (FILL) {
    const int ex = i == 1 ? -1 : (i == g.nx ? 1 : 0), ey = j == 1 ? -1 : (j == g.ny ? 1 : 0), ez = k == 1 ? -1 : (k == g.nz ? 1 : 0);
    // (gx_create demands >= 2 cells per direction: no cell touches both faces of a direction)
    if (ex | ey | ez) {
      for (int a = 0; a <= (ez != 0); ++a)
        for (int b = 0; b <= (ey != 0); ++b)
          for (int d = 0; d <= (ex != 0); ++d)
            if (a | b | d) E5[c + (long long)(a * ez) * sz + (long long)(b * ey) * sy + d * ex] = v;
    }
  }
}

// thermal_bounds' zero-gradient copies as ONE launch over the ghost shell (layer 1) when the block has neighbours: after the
// exchange, a ghost cell takes the value found by clamping its indices into the block in every direction whose face lies on
// the DOMAIN boundary (`edge` bit 2*dir+side) — what the six sequential plane copies of :592-614 leave behind; cells that
// are ghosts only across internal faces keep what the neighbour sent.  Sources are never destinations.
// Rows of the shell (blockIdx.y): 2 (ny+2) x-rows of the two z planes, 2 nz x-rows of the two y planes, 2 nz y-rows of the two x planes.
__global__ void __launch_bounds__(128) k_tc_fill(gx::Grid g, double* __restrict__ A, int edge) {
  const int t = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  int r = (int)blockIdx.y, i, j, k;
  if (r < 2 * (g.ny + 2)) { k = r < g.ny + 2 ? 0 : g.nz + 1; j = r < g.ny + 2 ? r : r - (g.ny + 2); i = t; if (i > g.nx + 1) return; }
  else if ((r -= 2 * (g.ny + 2)) < 2 * g.nz) { j = r < g.nz ? 0 : g.ny + 1; k = 1 + (r < g.nz ? r : r - g.nz); i = t; if (i > g.nx + 1) return; }
  else { r -= 2 * g.nz; i = r < g.nz ? 0 : g.nx + 1; k = 1 + (r < g.nz ? r : r - g.nz); j = 1 + t; if (j > g.ny) return; }
  int si = i, sj = j, sk = k;
  if (i == 0 && (edge & 1)) si = 1;
  if (i == g.nx + 1 && (edge & 2)) si = g.nx;
  if (j == 0 && (edge & 4)) sj = 1;
  if (j == g.ny + 1 && (edge & 8)) sj = g.ny;
  if (k == 0 && (edge & 16)) sk = 1;
  if (k == g.nz + 1 && (edge & 32)) sk = g.nz;
  if (si == i && sj == j && sk == k) return;
  A[g.idx(i, j, k)] = A[g.idx(si, sj, sk)];
}
