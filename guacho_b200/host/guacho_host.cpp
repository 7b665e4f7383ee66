// guacho_host.cpp — a compiled host for libguacho_gx.so that mirrors the reference's driver, src/main.f90:50-137,
// for the shipped Orszag-Tang problem: initmain / initflow (OT/user_mod.f90 -> OT/orzag_tang.f90:14-70), boundaryI +
// calcprim, write_output(0), then the time loop get_timestep -> tstep -> output every dtprint.  Everything numerical
// happens behind the C ABI (include/guacho_gx.h); this file only owns the host arrays, the loop and the BIN dumps
// (src/Out_BIN_Module.f90:40-102,129-165), exactly what the Fortran host keeps.  Single block (one GPU).
//
//   guacho_host [-n NX NY NZ] [-tmax T] [-dtprint DT] [-o OUTDIR] [-strict] [-quiet]
//
// Build: g++ -O2 -std=c++17 guacho_host.cpp -I../../include -L.. -lguacho_gx -Wl,-rpath,'$ORIGIN/..' -o guacho_host
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <sys/stat.h>
#include <vector>

#include "guacho_gx.h"

struct Par {                                   // OT/parameters.f90 (the values the shipped run compiles in)
  int nxtot = 512, nytot = 512, nztot = 2;
  double xmax = 1.0, ymax = 1.0, zmax = 2.0 / 512.0;
  double cv = 1.5, cfl = 0.2, eta = 0.0, tmax = 0.5, dtprint = 0.1;
  double rsc = 1.0, vsc = 1.0, rhosc = 1.0, Tempsc = 1.0;
  int neq = 8, nghost = 2;
};

static void die(const char* where, int rc) {   // the reference prints and stops
  std::fprintf(stderr, "guacho_host: %s failed (%d): %s\n", where, rc, gx_last_error());
  std::exit(1);
}

// u(neq, nxmin:nxmax, nymin:nymax, nzmin:nzmax), column-major, Fortran index base 1-nghost
struct Field {
  int neq, NX, NY, NZ;
  std::vector<double> d;
  Field(int neq_, int nx, int ny, int nz) : neq(neq_), NX(nx + 4), NY(ny + 4), NZ(nz + 4), d((size_t)neq_ * NX * NY * NZ, 0.0) {}
  double& at(int q, int i, int j, int k) { return d[(size_t)(q - 1) + (size_t)neq * ((size_t)(i + 1) + (size_t)NX * ((size_t)(j + 1) + (size_t)NY * (size_t)(k + 1)))]; }
};

// OT/orzag_tang.f90:14-70 impose_ot (coords = 0: single block)
static void impose_ot(const Par& p, Field& u) {
  const double pi = std::acos(-1.0), twopi = 2.0 * pi;
  const double rho = 25.0 / (36.0 * pi), pr = 5.0 / (12.0 * pi);
  const double dx = p.xmax / p.nxtot, dy = p.ymax / p.nytot;
  for (int i = -1; i <= p.nxtot + 2; ++i)
    for (int j = -1; j <= p.nytot + 2; ++j)
      for (int k = -1; k <= p.nztot + 2; ++k) {
        const double x = ((double)i + 0.5) * dx * p.rsc, y = ((double)j + 0.5) * dy * p.rsc;
        const double vx = -std::sin(y * twopi), vy = std::sin(x * twopi), vz = 0.0;
        const double bx = -std::sin(y * twopi) / std::sqrt(4 * pi), by = std::sin(2.0 * x * twopi) / std::sqrt(4 * pi), bz = 0.0;
        u.at(1, i, j, k) = rho;
        u.at(2, i, j, k) = rho * vx;
        u.at(3, i, j, k) = rho * vy;
        u.at(4, i, j, k) = rho * vz;
        u.at(5, i, j, k) = 0.5 * rho * (vx * vx + vy * vy + vz * vz) + p.cv * pr + 0.5 * (bx * bx + by * by + bz * bz);
        u.at(6, i, j, k) = bx;
        u.at(7, i, j, k) = by;
        u.at(8, i, j, k) = bz;
      }
}

// src/Out_BIN_Module.f90:40-102 write_header + :152-165 write_BIN, MPI naming with rank 0
static std::string es103(double x) { char b[32]; std::snprintf(b, sizeof b, "%10.3E", x); return b; }
static void write_bin(const Par& p, const std::string& outdir, int itprint, const Field& u) {
  mkdir(outdir.c_str(), 0755);
  mkdir((outdir + "/BIN").c_str(), 0755);
  char name[512];
  std::snprintf(name, sizeof name, "%s/BIN/points%03d.%03d.bin", outdir.c_str(), 0, itprint);
  FILE* f = std::fopen(name, "wb");
  if (!f) { std::perror(name); std::exit(1); }
  const double dx = p.xmax / p.nxtot, dy = p.ymax / p.nytot, dz = p.zmax / p.nztot;
  auto line = [&](const std::string& s) { std::string t = s; while (!t.empty() && t.back() == ' ') t.pop_back(); std::fwrite(t.data(), 1, t.size(), f); std::fputc('\n', f); };
  char b[256];
  line("**************** Output for Guacho v1.3****************");
  std::snprintf(b, sizeof b, "Dimensions    : %d %d %d", p.nxtot, p.nytot, p.nztot); line(b);
  line("Spacings      : " + es103(dx) + es103(dy) + es103(dz));
  line("Block Origin, cells    : 0 0 0");
  line("MPI blocks (X, Y, Z)   : 1 1 1");
  std::snprintf(b, sizeof b, "Number of Equations/dynamical ones  %d/%d", p.neq, 8); line(b);
  std::snprintf(b, sizeof b, "Number of Ghost Cells  %d", p.nghost); line(b);
  line("Scalings");
  line("r_sc: " + es103(p.rsc) + " v_sc: " + es103(p.vsc) + " rho_sc: " + es103(p.rhosc));
  std::snprintf(b, sizeof b, "Specfic heat at constant volume Cv: %7.2f", p.cv); line(b);
  line("Double precision 8 byte floats");
  line("*******************************************************");
  std::fputc(0xFF, f); std::fputc('\n', f); std::fputc('d', f);
  const int32_t n3[3] = {p.nxtot, p.nytot, p.nztot}, o3[3] = {0, 0, 0}, m3[3] = {1, 1, 1}, ne[2] = {p.neq, 8}, ng = p.nghost;
  const double d3[3] = {dx, dy, dz}, sc[3] = {p.rsc, p.vsc, p.rhosc};
  std::fwrite(n3, 4, 3, f); std::fwrite(d3, 8, 3, f); std::fwrite(o3, 4, 3, f); std::fwrite(m3, 4, 3, f);
  std::fwrite(ne, 4, 2, f); std::fwrite(&ng, 4, 1, f); std::fwrite(sc, 8, 3, f); std::fwrite(&p.cv, 8, 1, f);
  std::fwrite(u.d.data(), 8, u.d.size(), f);
  std::fclose(f);
}

int main(int argc, char** argv) {
  Par p;
  std::string outdir = ".";
  bool strict = false, quiet = false, grid_given = false;
  for (int a = 1; a < argc; ++a) {
    if (!std::strcmp(argv[a], "-n") && a + 3 < argc) { p.nxtot = std::atoi(argv[a + 1]); p.nytot = std::atoi(argv[a + 2]); p.nztot = std::atoi(argv[a + 3]); a += 3; grid_given = true; }
    else if (!std::strcmp(argv[a], "-tmax") && a + 1 < argc) p.tmax = std::atof(argv[++a]);
    else if (!std::strcmp(argv[a], "-dtprint") && a + 1 < argc) p.dtprint = std::atof(argv[++a]);
    else if (!std::strcmp(argv[a], "-o") && a + 1 < argc) outdir = argv[++a];
    else if (!std::strcmp(argv[a], "-strict")) strict = true;
    else if (!std::strcmp(argv[a], "-quiet")) quiet = true;
    else { std::fprintf(stderr, "usage: guacho_host [-n NX NY NZ] [-tmax T] [-dtprint DT] [-o OUTDIR] [-strict] [-quiet]\n"); return 2; }
  }
  if (grid_given) p.zmax = p.xmax * p.nztot / p.nxtot;        // cubic cells, like the shipped 512 x 512 x 2 box

  // initmain (src/init.f90:40-205): every `parameter` the step reads goes into gx_config
  gx_config c;
  std::memset(&c, 0, sizeof c);
  c.struct_bytes = (int32_t)sizeof c; c.device = -1;
  c.nxtot = p.nxtot; c.nytot = p.nytot; c.nztot = p.nztot;
  c.nbx = c.nby = c.nbz = 1; c.nghost = 2;
  c.neq = 8; c.neqdyn = 8; c.npas = 0; c.mhd = 1;
  c.riemann_solver = GX_SOLVER_HLLD; c.slope_limiter = GX_LIMITER_MINMOD; c.eq_of_state = GX_EOS_ADIABATIC;
  c.enable_flux_cd = 1;
  c.bc_left = c.bc_right = c.bc_bottom = c.bc_top = c.bc_out = c.bc_in = GX_BC_PERIODIC;
  c.strict_fp = strict ? 1 : 0; c.cooling = GX_COOL_NONE;
  c.dx = p.xmax / p.nxtot; c.dy = p.ymax / p.nytot; c.dz = p.zmax / p.nztot;
  c.cv = p.cv; c.gamma = (p.cv + 1.0) / p.cv; c.Tempsc = p.Tempsc; c.cfl = p.cfl; c.eta = p.eta; c.tsc = 1.0;
  gx_solver* s = nullptr;
  int rc = gx_create(&c, &s); if (rc) die("gx_create", rc);

  Field u(p.neq, p.nxtot, p.nytot, p.nztot);
  double time = 0.0, tprint = p.dtprint, dt_CFL = 0.0;       // init.f90:125-131
  int itprint = 0, currentIteration = 1;
  impose_ot(p, u);                                            // initflow -> initial_conditions(u)
  rc = gx_set_state(s, u.d.data()); if (rc) die("gx_set_state", rc);          // boundaryI + calcprim (main.f90:76-79)
  rc = gx_get_state(s, u.d.data(), nullptr, nullptr); if (rc) die("gx_get_state", rc);
  write_bin(p, outdir, itprint, u); itprint += 1;             // main.f90:84-87

  while (time <= p.tmax) {                                    // main.f90:94
    int32_t dump = 0;
    rc = gx_get_timestep(s, currentIteration, 10, time, tprint, &dt_CFL, &dump); if (rc) die("gx_get_timestep", rc);
    if (!quiet) std::printf("Iteration %d | time:%12.3E | dt:%12.3E | tprint:%12.3E\n", currentIteration, time, dt_CFL, tprint);
    rc = gx_set_time(s, time); if (rc) die("gx_set_time", rc);
    rc = gx_tstep(s, dt_CFL); if (rc) die("gx_tstep", rc);    // main.f90:106
    time += dt_CFL;
    if (dump) {                                               // main.f90:110-121
      rc = gx_get_state(s, u.d.data(), nullptr, nullptr); if (rc) die("gx_get_state", rc);
      write_bin(p, outdir, itprint, u);
      if (!quiet) std::printf("****************** wrote output *************** :%4d\n", itprint);
      tprint += p.dtprint; itprint += 1;
    }
    currentIteration += 1;
  }
  std::printf("--- My work here is done, have a nice day ---  (%d iterations, %lld kernel launches)\n", currentIteration - 1, (long long)gx_launch_count(s));
  gx_destroy(s);
  return 0;
}
