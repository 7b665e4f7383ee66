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

// ---------------------------------------------------------------------------
// One substep of ISOTROPIC conduction for a block that owns the whole domain, as ONE marching kernel: no pressure /
// temperature arrays in HBM and every face flux evaluated once.  A CTA owns a 32 x 8 column of cells and marches along z.
// Per plane a thread loads the conserved variables of its cell in the NEXT plane (+ one cell of that plane's halo ring),
// converts them (u2prim: the reference's calcprim, restricted to what the fluxes read) and parks pressure, temperature and
// density of the plane in shared memory; it evaluates the upper x and y face of its cell from the CURRENT plane's tile and
// the upper z face from registers (the lower z face is the previous plane's upper one), hands the x / y fluxes over through
// shared memory and updates u(5).  The 40 faces on the low x / y side of the tile are shared out to the first 40 threads.
// u(5) is read from E5in and written to E5out (the halo cells of a tile belong to other CTAs, which may already have
// updated them): the caller alternates the energy array of u and a scratch array.  Ghost copies of thermal_bounds as in
// k_tc_update<FILL>.  Algorithmic traffic: neqdyn reads + 1 write per zone and substep (72 B; the two-kernel form moves 120).
// Same arithmetic per face and per cell as heatfluxes / the update loop of thermal_conduction (:189-267, :749-757).
constexpr int TCM_TX = 32, TCM_TY = 8, TCM_SX = TCM_TX + 2, TCM_SY = TCM_TY + 2, TCM_NT = TCM_TX * TCM_TY;
#ifndef GX_TCM_MINB                  // resident CTAs per SM: latency-bound like k_tc_update — 4 (64 registers, a few spills): 4.47 ms for 9 substeps
#define GX_TCM_MINB 4                // at 256^3, 3 (80 registers): 4.85, 2 (100): 6.42
#endif
template <bool MHD>
__global__ void __launch_bounds__(TCM_NT, GX_TCM_MINB) k_tc_march(gx::Grid g, gxp::Phys P, gx::TcPar t, const double* __restrict__ U, const double* __restrict__ E5in,
                                                         double* __restrict__ E5out, double dts, int kz) {
  constexpr int TX = TCM_TX, TY = TCM_TY, SX = TCM_SX, SY = TCM_SY, NT = TCM_NT;
  __shared__ double sT[2][SY][SX], sp[2][SY][SX], sr[2][SY][SX];
  __shared__ double fx[TY][TX + 1], fy[TY + 1][TX];
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TX + tx;
  const int i0 = 1 + (int)blockIdx.x * TX, j0 = 1 + (int)blockIdx.y * TY;
  const int k0 = 1 + (int)blockIdx.z * kz, k1 = min(k0 + kz - 1, g.nz);
  const int i = i0 + tx, j = j0 + ty;
  const bool ok = i <= g.nx && j <= g.ny;
  const long long vs = g.vs, sz = (long long)g.px * g.py;
  // this thread's halo cell of a plane (tile-local coordinates; the four corners are never read): 2 x 32 + 2 x 8 = 80 cells
  int hx = -1, hy = -1;
  if (tid < TX) { hx = tid + 1; hy = 0; }
  else if (tid < 2 * TX) { hx = tid - TX + 1; hy = SY - 1; }
  else if (tid < 2 * TX + TY) { hx = 0; hy = tid - 2 * TX + 1; }
  else if (tid < 2 * TX + 2 * TY) { hx = SX - 1; hy = tid - 2 * TX - TY + 1; }
  // addresses clamped into the ghost layer (tiles may overhang the block; such cells are never used by a live update)
  const long long own0 = g.idx(min(i, g.nx + 1), min(j, g.ny + 1), 0);
  const long long halo0 = g.idx(min(i0 - 1 + max(hx, 0), g.nx + 1), min(j0 - 1 + max(hy, 0), g.ny + 1), 0);
  auto prim_of = [&](long long c, double& pT, double& pp, double& pr, double& e5) {
    double uu[8], w[8], T;
#pragma unroll
    for (int q = 0; q < 8; ++q) uu[q] = (q < (MHD ? 8 : 5) && q != 4) ? U[q * vs + c] : 0.0;
    uu[4] = e5 = E5in[c];
    const double un = g.npas > 0 ? U[(long long)g.neqdyn * vs + c] : 0.0;
    gxp::u2prim<MHD, true>(P, uu, w, un, T);
    pT = T; pp = w[4]; pr = w[0];
  };
  auto stage_plane = [&](int k, int b, double& Tn, double& pn, double& rn, double& en) {      // plane k -> tile buffer b (+ own values)
    const long long off = (long long)k * sz;
    prim_of(own0 + off, Tn, pn, rn, en);
    sT[b][ty + 1][tx + 1] = Tn; sp[b][ty + 1][tx + 1] = pn; sr[b][ty + 1][tx + 1] = rn;
    if (hx >= 0) {
      double hT, hp, hr, he;
      prim_of(halo0 + off, hT, hp, hr, he);
      sT[b][hy][hx] = hT; sp[b][hy][hx] = hp; sr[b][hy][hx] = hr;
    }
  };
  // prologue: plane k0-1 (own column only: the first lower z flux) and plane k0 (tile)
  double Tc, pc, rc, ec, Tn, pn, rn, en, zlo;
  {
    double Tm, pm, rm, em;
    prim_of(own0 + (long long)(k0 - 1) * sz, Tm, pm, rm, em);
    stage_plane(k0, 0, Tc, pc, rc, ec);
    zlo = flux_iso(t, P, Tm, Tc, pm, pc, rm, rc, t.dzr, t.idzr);
  }
  __syncthreads();
  int b = 0;
#pragma unroll 1
  for (int k = k0; k <= k1; ++k, b ^= 1) {
    stage_plane(k + 1, b ^ 1, Tn, pn, rn, en);                       // next plane: conversion + tile for the next iteration
    // upper x and y face of my cell, from the current plane's tile
    fx[ty][tx + 1] = flux_iso(t, P, Tc, sT[b][ty + 1][tx + 2], pc, sp[b][ty + 1][tx + 2], rc, sr[b][ty + 1][tx + 2], t.dxr, t.idxr);
    fy[ty + 1][tx] = flux_iso(t, P, Tc, sT[b][ty + 2][tx + 1], pc, sp[b][ty + 2][tx + 1], rc, sr[b][ty + 2][tx + 1], t.dyr, t.idyr);
    if (tid < TY) {                                                  // low x side of the tile: face (i0-1 | i0) of row tid
      fx[tid][0] = flux_iso(t, P, sT[b][tid + 1][0], sT[b][tid + 1][1], sp[b][tid + 1][0], sp[b][tid + 1][1], sr[b][tid + 1][0], sr[b][tid + 1][1], t.dxr, t.idxr);
    } else if (tid < TY + TX) {                                      // low y side: face (j0-1 | j0) of column tid - TY
      const int cx = tid - TY + 1;
      fy[0][cx - 1] = flux_iso(t, P, sT[b][0][cx], sT[b][1][cx], sp[b][0][cx], sp[b][1][cx], sr[b][0][cx], sr[b][1][cx], t.dyr, t.idyr);
    }
    const double zhi = flux_iso(t, P, Tc, Tn, pc, pn, rc, rn, t.dzr, t.idzr);
    __syncthreads();                                                 // fluxes of this plane and the next plane's tile are in place
    if (ok) {
      const long long c = own0 + (long long)k * sz;
#if defined(GX_FLAVOUR_FAST)
      const double v = ec - dts * ((fx[ty][tx + 1] - fx[ty][tx]) * t.idx + (fy[ty + 1][tx] - fy[ty][tx]) * t.idy + (zhi - zlo) * t.idz);
#else
      const double v = ec - dts * ((fx[ty][tx + 1] - fx[ty][tx]) / t.dx + (fy[ty + 1][tx] - fy[ty][tx]) / t.dy + (zhi - zlo) / t.dz);
#endif
      E5out[c] = v;
      const int ex = i == 1 ? -1 : (i == g.nx ? 1 : 0), ey = j == 1 ? -1 : (j == g.ny ? 1 : 0), ez = k == 1 ? -1 : (k == g.nz ? 1 : 0);
      if (ex | ey | ez) {
        for (int a = 0; a <= (ez != 0); ++a)
          for (int bb = 0; bb <= (ey != 0); ++bb)
            for (int d = 0; d <= (ex != 0); ++d)
              if (a | bb | d) E5out[c + (long long)(a * ez) * sz + (long long)(bb * ey) * g.px + d * ex] = v;
      }
    }
    zlo = zhi; Tc = Tn; pc = pn; rc = rn; ec = en;
    __syncthreads();                                                 // everyone has read fx / fy and tile b before the next plane overwrites them
  }
}
