/* guacho_gx.h — C ABI of the B200 hydro/MHD time-step library (libguacho_gx.so).
 *
 * This is the drop-in boundary for Guacho-3D's hot path: the calls that
 * src/main.f90 makes into hydro_core / hydro_solver / boundaries are replaced
 * by the entry points below (plain pointers and sizes, no C++/torch types).
 * Every entry point cites the reference interface it replaces (file:line,
 * relative to the reference tree).  The Fortran side binds them through
 * guacho_b200/fortran/guacho_gpu.f90 (ISO_C_BINDING); see INTEGRATION.md.
 *
 * All arrays passed across this boundary use the REFERENCE layout:
 *   u(neq, nxmin:nxmax, nymin:nymax, nzmin:nzmax), column-major, FP64, ghosts
 *   included (nghost = 2), i.e. element (ieq,i,j,k) [Fortran indices] lives at
 *   (ieq-1) + neq*((i+1) + (nx+4)*((j+1) + (ny+4)*(k+1)))
 * (src/init.f90:145-150, src/globals.f90:33-39).  The library owns all device
 * memory and converts to its own SoA layout internally.
 *
 * Return value of every function: 0 on success, a negative GX_E* code on
 * failure; gx_last_error() returns a human-readable message.  There is no CPU
 * fallback: if no CUDA device is usable gx_create fails with GX_ENODEVICE.
 */
#ifndef GUACHO_GX_H
#define GUACHO_GX_H

#include <stdint.h>

#if defined(__GNUC__)
#define GX_API __attribute__((visibility("default")))
#else
#define GX_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ---- named constants: values identical to src/constants.f90:56-98 ---- */
enum { GX_SOLVER_HLL = 1, GX_SOLVER_HLLC = 2, GX_SOLVER_HLLE = 3, GX_SOLVER_HLLD = 4,
       GX_SOLVER_HLLE_SPLIT_B = 5, GX_SOLVER_HLLD_SPLIT_B = 6,
       GX_SOLVER_HLLE_SPLIT_ALL = 7, GX_SOLVER_HLLD_SPLIT_ALL = 8 };
enum { GX_EOS_ADIABATIC = 1, GX_EOS_SINGLE_SPECIE = 2, GX_EOS_H_RATE = 3, GX_EOS_CHEM = 4 };
enum { GX_BC_OUTFLOW = 1, GX_BC_CLOSED = 2, GX_BC_PERIODIC = 3, GX_BC_OTHER = 4 };
enum { GX_LIMITER_NO_AVERAGE = -1, GX_LIMITER_NO_LIMIT = 0, GX_LIMITER_MINMOD = 1,
       GX_LIMITER_VAN_LEER = 2, GX_LIMITER_VAN_ALBADA = 3, GX_LIMITER_UMIST = 4,
       GX_LIMITER_WOODWARD = 5, GX_LIMITER_SUPERBEE = 6 };

enum { GX_TC_OFF = 0, GX_TC_ISOTROPIC = 1, GX_TC_ANISOTROPIC = 2 };
enum { GX_COOL_NONE = 0, GX_COOL_H = 1, GX_COOL_BBC = 2, GX_COOL_DMC = 3, GX_COOL_CHI = 4, GX_COOL_CHEM = 5 };

/* ---- error codes ---- */
enum { GX_OK = 0, GX_EINVAL = -1, GX_ENODEVICE = -2, GX_ECUDA = -3, GX_ENOMEM = -4,
       GX_EUNSUPPORTED = -5, GX_ESTATE = -6, GX_ECOMM = -7, GX_ENUMERIC = -8 };

/* Every Fortran `parameter` the step reads (OT/parameters.f90:48-227) plus the
 * block decomposition that replaces MPI_NBX/NBY/NBZ and mpi_cart_coords
 * (src/init.f90:103-110).  Plain int32/double, no implicit padding: 36 int32 first
 * (35 parameters + one explicit pad word, so the doubles start on an 8-byte offset in
 * every binding, packed or not), then 14 doubles; gx_api.cu static_asserts the offsets. */
typedef struct gx_config {
  int32_t struct_bytes;      /* = sizeof(gx_config); ABI check                         */
  int32_t device;            /* CUDA device ordinal; -1 = keep the current device     */
  int32_t nxtot, nytot, nztot;   /* parameters.f90: nxtot nytot nztot                 */
  int32_t nbx, nby, nbz;     /* MPI_NBX MPI_NBY MPI_NBZ  -> one block per GPU         */
  int32_t cx, cy, cz;        /* this block's coords(0:2) (mpi_cart_coords)            */
  int32_t nghost;            /* must be 2 (parameters.f90:193)                        */
  int32_t neq, neqdyn, npas; /* neq = neqdyn + npas; neqdyn = 5 (hydro) | 8 (BFIELD)  */
  int32_t mhd, pmhd, passives;
  int32_t riemann_solver;    /* GX_SOLVER_*                                           */
  int32_t slope_limiter;     /* GX_LIMITER_*                                          */
  int32_t eq_of_state;       /* GX_EOS_*                                              */
  int32_t enable_flux_cd, eight_wave, user_source_terms;
  int32_t bc_left, bc_right, bc_bottom, bc_top, bc_out, bc_in;  /* GX_BC_*           */
  int32_t bc_user;
  int32_t strict_fp;         /* 1: kernels built with -fmad=false (bit-comparison mode,
                                matches the reference's no-FMA x86 build); 0: FMA     */
  int32_t cooling;           /* GX_COOL_*: NONE, or H = the parametrised hydrogen cooling operator
                                (src/cooling_h.f90:41-67, applied after viscous_copy, hydro_solver.f90:202-204);
                                it needs EOS_H_RATE-style passives (npas >= 1: neutral H density)   */
  int32_t th_cond;           /* GX_TC_*: thermal conduction operator at the end of tstep (src/thermal_cond.f90:690-768,
                                called from hydro_solver.f90:227); OFF, ISOTROPIC (Spitzer), ANISOTROPIC (needs B)        */
  int32_t tc_saturation;     /* parameters.f90: tc_saturation (saturated heat flux)                          */
  int32_t pad_;              /* explicit padding word (keeps the int32 count even); set to 0                  */
  double dx, dy, dz;         /* globals dx dy dz (src/init.f90:120-122)               */
  double cv, gamma;          /* parameters.f90: cv, gamma=(cv+1)/cv                   */
  double Tempsc;             /* temperature scaling used by u2prim                    */
  double cfl, eta;
  double tsc;                /* time scaling to seconds (parameters.f90: tsc); used by GX_COOL_H and thermal conduction  */
  double rsc, rhosc, vsc2;   /* length, density and velocity^2 scalings to cgs (parameters.f90:163-167); thermal conduction only  */
  double bsc, mu;            /* magnetic field scaling, mean atomic mass (parameters.f90:159,170); thermal conduction only        */
} gx_config;

typedef struct gx_solver gx_solver;   /* opaque; one per block (= per GPU / MPI rank) */

/* Replaces the allocation tail of initmain (src/init.f90:144-159): validates the
 * configuration, selects the device and allocates u, up, fluxes and E on it. */
GX_API int gx_create(const gx_config* cfg, gx_solver** out);

/* Frees all device memory (the reference never deallocates; end of main.f90). */
GX_API int gx_destroy(gx_solver* s);

/* Replaces: the host filling `primit0` (src/globals.f90:42, allocated at src/init.f90:153-154) — the background
 * primitives of SOLVER_HLLE_SPLIT_ALL (src/hlle_split_all.f90:51-241; u2primSplitAll, src/hydro_core.f90:143-229).
 * Same layout as gx_set_state's array; u then holds FLUCTUATIONS about it (totals: src/Out_BIN_Module.f90:144-156).
 * Must precede gx_set_state.  The reference ships no problem that uses this solver and marks its energy flux "REVISAR". */
GX_API int gx_set_background(gx_solver* s, const double* primit0);

/* Replaces initflow -> boundaryI -> calcprim at start-up (src/main.f90:73-79):
 * `u` is the caller-owned conserved array in reference layout WITH ghosts.
 * The library uploads it, converts to SoA, and applies boundaryI semantics. */
GX_API int gx_set_state(gx_solver* s, const double* u);

/* Simulation time seen by user boundary functors (globals::time; used by
 * impose_user_bc, EXO/user_mod.f90:131-144). */
GX_API int gx_set_time(gx_solver* s, double time);

/* Replaces get_timestep (src/hydro_core.f90:623-697): CFL minimum over the
 * block's physical cells (and over all blocks when a communicator is attached),
 * start-up ramp for current_iter <= n_iter, clipping to tprint.  *dump_flag is
 * only ever set to 1 (never cleared), like the reference's intent(out) logical
 * that is assigned only inside the `if`. */
GX_API int gx_get_timestep(gx_solver* s, int32_t current_iter, int32_t n_iter, double current_time,
                    double tprint, double* dt, int32_t* dump_flag);

/* Replaces tstep (src/hydro_solver.f90:134-229) for the hydro/MHD part:
 * first-order half step, boundaryII, second-order full step, viscous_copy,
 * boundaryI, primitives.  `dt_cfl` is globals::dt_CFL. */
GX_API int gx_tstep(gx_solver* s, double dt_cfl);

/* Convenience for benchmarking / long runs: n_steps iterations of
 * (get_timestep, tstep, time += dt) entirely driven from the library, as the
 * loop body of src/main.f90:94-125 without output.  In/out: *time, *iter. */
GX_API int gx_run(gx_solver* s, int32_t n_steps, int32_t n_iter_ramp, double* time, int32_t* iter,
           double* last_dt);

/* Implicit "state is on the host" before write_output (src/main.f90:85,112):
 * fills caller arrays in reference layout.  Any pointer may be NULL.  NOT collective: a rank may call it on its own
 * (ghost layers owned by other blocks hold what the last step exchanged; only self-periodic copies are refreshed).
 *   u      : (neq, nx+4, ny+4, nz+4)   conserved, as after boundaryI
 *   primit : (neq, nx+4, ny+4, nz+4)   calcprim(u, primit) over the whole array
 *   temp   : (nx+4, ny+4, nz+4)        Temp from u2prim                         */
GX_API int gx_get_state(gx_solver* s, double* u, double* primit, double* temp);

/* Same for the half-step array `up` (debug/parity aid; globals::up). */
GX_API int gx_get_up(gx_solver* s, double* up);

/* ---- user_mod plugin surface (OT/user_mod.f90:43-126, EXO/user_mod.f90) ---- */

/* get_user_source_terms as a device functor: N point masses,
 *   s(2:4) -= rho*GM*r_vec/|r|^3,  s(5) -= rho*GM*(v.r_vec)/|r|^3
 * with the reference's cell-centre convention (i - nxtot/2 - 0.5)*dx
 * (EXO/user_mod.f90:158-206).  gm[n], pos[3*n] in code units.  n = 0 disables. */
GX_API int gx_set_gravity_points(gx_solver* s, int32_t n, const double* gm, const double* pos);

/* impose_user_bc as a device functor: conserved state re-imposed inside
 * spheres on every boundaryI/boundaryII call (EXO/exoplanet.f90:125-266).
 * Each sphere: centre, radius, radial wind speed, density, thermal term
 * cv*dens*tfac*temp, bulk velocity, dipole moment amplitude (b0 at radius), passive
 * values per unit density.  See gx_wind_sphere. */
#define GX_MAX_SPHERES 4
typedef struct gx_wind_sphere {
  double xc, yc, zc, radius;
  double vwind, dens;
  double tfac, temp;                 /* u5 thermal part = cv*dens*tfac*temp (EXO: 1|1.9999 x TSW, 1.8 x TPW) */
  double vbx, vby, vbz;              /* bulk (orbital) velocity added to the wind */
  double bdip;                       /* dipole field strength at `radius`        */
  double pas[4];                     /* passive scalars per unit density (npas<=4) */
} gx_wind_sphere;
GX_API int gx_set_wind_spheres(gx_solver* s, int32_t n, const gx_wind_sphere* sph);

/* Host hook run at the top of every impose_user_bc application (boundaryI: order 1, boundaryII: order 2;
 * src/boundaries.f90:245,508) with globals::time as set by gx_set_time.  It exists because the reference's
 * problem modules mutate their own state there (EXO/exoplanet.f90:137-144 moves the planet, and the user
 * SOURCE then sees that position, EXO/user_mod.f90:174-187): the hook may call gx_set_wind_spheres /
 * gx_set_gravity_points, which take effect immediately.  No array crosses the boundary; cheap. */
typedef void (*gx_bc_hook_fn)(int32_t order, double time, void* user);
GX_API int gx_register_bc_hook(gx_solver* s, gx_bc_hook_fn cb, void* user);

/* Slow path for arbitrary user code: host callbacks run on a host copy of the
 * array in reference layout (device -> host -> callback -> device). Excluded
 * from any timed path.  cb(u, order, user) mirrors impose_user_bc(u, order). */
typedef void (*gx_host_bc_fn)(double* u, int32_t order, void* user);
GX_API int gx_register_host_bc(gx_solver* s, gx_host_bc_fn cb, void* user);

/* Slow path of get_user_source_terms(pp, s, i, j, k) (src/sources.f90:205, OT/user_mod.f90:111) for arbitrary user code:
 * once per stage the library hands the callback the primitives of the block in reference layout,
 * primit(neq, nx+4, ny+4, nz+4), and a zero-filled s of the same shape; the callback adds its source terms for the
 * physical cells (a Fortran host loops i, j, k over 1..n and calls its own get_user_source_terms(primit(:,i,j,k),
 * s(:,i,j,k), i, j, k)); the library then applies up = up + dt * s (src/hydro_solver.f90:115-121).  Device -> host ->
 * callback -> device on every stage: excluded from any timed path.  With user_source_terms = 1 and neither this callback
 * nor gx_set_gravity_points, gx_tstep / gx_run fail with GX_ESTATE (never a silent no-op). */
typedef void (*gx_host_source_fn)(const double* primit, double* s, void* user);
GX_API int gx_register_host_source(gx_solver* s, gx_host_source_fn cb, void* user);

/* ---- multi-GPU: replaces mpi_cart_create / mpi_sendrecv / mpi_allreduce
 * (src/init.f90:103-110, src/boundaries.f90:77-99,292-314,
 *  src/flux_cd_module.f90:75-97, src/hydro_core.f90:685) with NCCL over NVLink.
 * Rank r of the communicator owns block coords given by the row-major map
 * r = (cx*nby + cy)*nbz + cz.  The 128-byte id is created on rank 0 and
 * distributed by the host (MPI_Bcast in a Fortran host, torch.distributed in
 * the Python host). */
GX_API int gx_comm_unique_id(void* id_out, int32_t nbytes);
GX_API int gx_comm_attach(gx_solver* s, const void* id, int32_t nbytes, int32_t rank, int32_t nranks);

/* ---- per-interface entry point ----
 * Replaces a direct call of prim2fhll / prim2fhllc / prim2fhlle / prim2fhlld (src/hll.f90:47-82, src/hllc.f90:44-140,
 * src/hlle.f90:48-83, src/hlld.f90:48-319) for n independent interfaces: wl, wr = primitive states either side,
 * already rotated into the sweep direction (swapy/swapz, src/hydro_core.f90:485-534), [n][neqdyn] with neqdyn fastest
 * (HOST memory); ff = the flux, same shape; err[t] = 1 where the reference would print 'Error in HLLD/hllc routine' and
 * stop (NULL: not wanted).  Uses cfg->riemann_solver, neqdyn, cv, gamma, strict_fp and device; runs the same device
 * functions as the sweeps of gx_tstep (no CPU fallback). */
GX_API int gx_riemann_flux(const gx_config* cfg, int32_t n, const double* wl, const double* wr, double* ff, int32_t* err);

/* What the reference writes to logs/thermal_conduction.log per call (src/thermal_cond.f90:723-726): the conduction
 * time scale in seconds (get_dt_cond, :78-110) and the number of substeps of the last gx_tstep.  Zero when th_cond = 0. */
GX_API int gx_tc_info(const gx_solver* s, double* dt_cond, int32_t* nsteps);

/* ---- diagnostics ---- */
GX_API const char* gx_last_error(void);
/* number of kernels of this library launched since gx_create (bench evidence) */
GX_API int64_t gx_launch_count(const gx_solver* s);
/* device time of the last gx_tstep / gx_run in ms, CUDA events on the solver's
 * own stream; per-kernel-class accumulators for the roofline report. */
GX_API double gx_last_elapsed_ms(const gx_solver* s);
GX_API int gx_kernel_time_ms(const gx_solver* s, int32_t which, double* total_ms, int64_t* launches);
GX_API int gx_set_profiling(gx_solver* s, int32_t on);
/* library build info: "sm_100a fmad=on ..." */
GX_API const char* gx_build_info(void);

#ifdef __cplusplus
}
#endif
#endif /* GUACHO_GX_H */
