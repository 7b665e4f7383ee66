!=======================================================================
!> @file guacho_gpu.f90
!> @brief ISO_C_BINDING interface to libguacho_gx.so (include/guacho_gx.h)
!> @details Drop-in replacement of the hydro/MHD step for Guacho-3D's
!> Fortran host.  The host keeps main.f90, init.f90, parameters.f90,
!> user_mod.f90 and the output modules; the calls it makes into
!> hydro_core / hydro_solver / boundaries become calls into this module
!> (see INTEGRATION.md for the exact edits to main.f90).
!> Not compiled in this repository's image (no Fortran compiler here);
!> it is the binding a maintainer adds to the reference's Makefile
!> (OBJECTS += guacho_gpu.o ; LDFLAGS += -lguacho_gx).
!=======================================================================
module guacho_gpu
  use iso_c_binding
  implicit none

  !> image of struct gx_config (all int32 first, then doubles)
  type, bind(C) :: gx_config
    integer(c_int32_t) :: struct_bytes, device
    integer(c_int32_t) :: nxtot, nytot, nztot
    integer(c_int32_t) :: nbx, nby, nbz
    integer(c_int32_t) :: cx, cy, cz
    integer(c_int32_t) :: nghost
    integer(c_int32_t) :: neq, neqdyn, npas
    integer(c_int32_t) :: mhd, pmhd, passives
    integer(c_int32_t) :: riemann_solver, slope_limiter, eq_of_state
    integer(c_int32_t) :: enable_flux_cd, eight_wave, user_source_terms
    integer(c_int32_t) :: bc_left, bc_right, bc_bottom, bc_top, bc_out, bc_in
    integer(c_int32_t) :: bc_user, strict_fp, cooling, th_cond, tc_saturation, pad_
    real(c_double)     :: dx, dy, dz, cv, gamma, Tempsc, cfl, eta, tsc, rsc, rhosc, vsc2, bsc, mu
  end type gx_config

  !> image of struct gx_wind_sphere (impose_user_bc functor, EXO/exoplanet.f90:125-266)
  type, bind(C) :: gx_wind_sphere
    real(c_double) :: xc, yc, zc, radius
    real(c_double) :: vwind, dens
    real(c_double) :: tfac, temp
    real(c_double) :: vbx, vby, vbz
    real(c_double) :: bdip
    real(c_double) :: pas(4)
  end type gx_wind_sphere

  type(c_ptr), save :: gx_handle = c_null_ptr   !< the solver of this MPI rank

  interface
    integer(c_int) function gx_create(cfg, handle) bind(C, name="gx_create")
      import :: c_int, c_ptr, gx_config
      type(gx_config), intent(in) :: cfg
      type(c_ptr), intent(out)    :: handle
    end function
    integer(c_int) function gx_destroy(handle) bind(C, name="gx_destroy")
      import :: c_int, c_ptr
      type(c_ptr), value :: handle
    end function
    !> primit0 of SOLVER_HLLE_SPLIT_ALL (globals.f90:42, init.f90:153-154): background primitives; before gx_set_state
    integer(c_int) function gx_set_background(handle, primit0) bind(C, name="gx_set_background")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value         :: handle
      real(c_double), intent(in) :: primit0(*)
    end function
    !> initflow -> boundaryI -> calcprim  (main.f90:73-79)
    integer(c_int) function gx_set_state(handle, u) bind(C, name="gx_set_state")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value         :: handle
      real(c_double), intent(in) :: u(*)
    end function
    integer(c_int) function gx_set_time(handle, time) bind(C, name="gx_set_time")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value    :: handle
      real(c_double), value :: time
    end function
    !> get_timestep  (hydro_core.f90:623-697)
    integer(c_int) function gx_get_timestep(handle, current_iter, n_iter, current_time, tprint, dt, dump_flag) &
        bind(C, name="gx_get_timestep")
      import :: c_int, c_int32_t, c_ptr, c_double
      type(c_ptr), value           :: handle
      integer(c_int32_t), value    :: current_iter, n_iter
      real(c_double), value        :: current_time, tprint
      real(c_double), intent(out)  :: dt
      integer(c_int32_t), intent(inout) :: dump_flag
    end function
    !> tstep  (hydro_solver.f90:134-229)
    integer(c_int) function gx_tstep(handle, dt_cfl) bind(C, name="gx_tstep")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value    :: handle
      real(c_double), value :: dt_cfl
    end function
    !> state on the host before write_output (main.f90:85,112); pass c_null_ptr to skip an array
    integer(c_int) function gx_get_state(handle, u, primit, temp) bind(C, name="gx_get_state")
      import :: c_int, c_ptr
      type(c_ptr), value :: handle, u, primit, temp
    end function
    integer(c_int) function gx_comm_unique_id(id, nbytes) bind(C, name="gx_comm_unique_id")
      import :: c_int, c_int32_t, c_char
      character(kind=c_char), intent(out) :: id(*)
      integer(c_int32_t), value :: nbytes
    end function
    integer(c_int) function gx_comm_attach(handle, id, nbytes, rank, nranks) bind(C, name="gx_comm_attach")
      import :: c_int, c_int32_t, c_ptr, c_char
      type(c_ptr), value :: handle
      character(kind=c_char), intent(in) :: id(*)
      integer(c_int32_t), value :: nbytes, rank, nranks
    end function
    integer(c_int) function gx_set_gravity_points(handle, n, gm, pos) bind(C, name="gx_set_gravity_points")
      import :: c_int, c_int32_t, c_ptr, c_double
      type(c_ptr), value :: handle
      integer(c_int32_t), value :: n
      real(c_double), intent(in) :: gm(*), pos(*)
    end function
    !> n_steps iterations of the loop body of main.f90:94-125 without output
    integer(c_int) function gx_run(handle, n_steps, n_iter_ramp, time, iter, last_dt) bind(C, name="gx_run")
      import :: c_int, c_int32_t, c_ptr, c_double
      type(c_ptr), value :: handle
      integer(c_int32_t), value :: n_steps, n_iter_ramp
      real(c_double), intent(inout) :: time
      integer(c_int32_t), intent(inout) :: iter
      real(c_double), intent(out) :: last_dt
    end function
    !> the half-step array `up` on the host (debug / parity aid)
    integer(c_int) function gx_get_up(handle, up) bind(C, name="gx_get_up")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: handle
      real(c_double), intent(out) :: up(*)
    end function
    !> thermal conduction log line (src/thermal_cond.f90:723-726): dt_cond in seconds, substeps of the last gx_tstep
    integer(c_int) function gx_tc_info(handle, dt_cond, nsteps) bind(C, name="gx_tc_info")
      import :: c_int, c_int32_t, c_ptr, c_double
      type(c_ptr), value :: handle
      real(c_double), intent(out) :: dt_cond
      integer(c_int32_t), intent(out) :: nsteps
    end function
    !> impose_user_bc as a device functor: wind spheres (EXO/exoplanet.f90:125-266)
    integer(c_int) function gx_set_wind_spheres(handle, n, sph) bind(C, name="gx_set_wind_spheres")
      import :: c_int, c_int32_t, c_ptr, gx_wind_sphere
      type(c_ptr), value :: handle
      integer(c_int32_t), value :: n
      type(gx_wind_sphere), intent(in) :: sph(*)
    end function
    !> host hook at the top of every impose_user_bc application (the planet moves there, exoplanet.f90:137-144)
    integer(c_int) function gx_register_bc_hook(handle, cb, user) bind(C, name="gx_register_bc_hook")
      import :: c_int, c_ptr, c_funptr
      type(c_ptr), value :: handle
      type(c_funptr), value :: cb          !< c_funloc of a bind(C) subroutine (order, time, user)
      type(c_ptr), value :: user
    end function
    !> slow paths for arbitrary user code: impose_user_bc(u, order) / get_user_source_terms on host arrays
    integer(c_int) function gx_register_host_bc(handle, cb, user) bind(C, name="gx_register_host_bc")
      import :: c_int, c_ptr, c_funptr
      type(c_ptr), value :: handle
      type(c_funptr), value :: cb          !< bind(C) subroutine (u, order, user)
      type(c_ptr), value :: user
    end function
    integer(c_int) function gx_register_host_source(handle, cb, user) bind(C, name="gx_register_host_source")
      import :: c_int, c_ptr, c_funptr
      type(c_ptr), value :: handle
      type(c_funptr), value :: cb          !< bind(C) subroutine (primit, s, user)
      type(c_ptr), value :: user
    end function
    !> prim2fhll / prim2fhllc / prim2fhlle / prim2fhlld for n interfaces on the device
    integer(c_int) function gx_riemann_flux(cfg, n, wl, wr, ff, err) bind(C, name="gx_riemann_flux")
      import :: c_int, c_int32_t, c_double, c_ptr, gx_config
      type(gx_config), intent(in) :: cfg
      integer(c_int32_t), value :: n
      real(c_double), intent(in) :: wl(*), wr(*)
      real(c_double), intent(out) :: ff(*)
      type(c_ptr), value :: err            !< c_loc of an integer(c_int32_t) array, or c_null_ptr
    end function
    type(c_ptr) function gx_last_error() bind(C, name="gx_last_error")
      import :: c_ptr
    end function
  end interface

contains

  !> stop with the library's message, like the reference's print + stop (hlld.f90:316-317)
  subroutine gx_check(rc, where)
    integer(c_int), intent(in) :: rc
    character(len=*), intent(in) :: where
    character(kind=c_char), pointer :: msg(:)
    integer :: n
    if (rc == 0) return
    call c_f_pointer(gx_last_error(), msg, [512])
    n = 1
    do while (n < 512 .and. msg(n) /= c_null_char)
      n = n + 1
    end do
    print '(a,a,a,i0,a,512a1)', 'guacho_gx error in ', where, ' (', rc, '): ', msg(1:n-1)
    stop
  end subroutine gx_check

  !> replaces the allocation tail of initmain (init.f90:144-159): fills gx_config from the
  !> compile-time parameters and the MPI cartesian coordinates, one solver per rank
  subroutine gx_initmain()
    use parameters
    use globals, only : dx, dy, dz, coords, rank
    type(gx_config) :: c
    c%struct_bytes = int(c_sizeof(c), c_int32_t)
    c%device = -1                       ! keep the device the launcher selected (one rank per GPU)
    c%nxtot = nxtot;  c%nytot = nytot;  c%nztot = nztot
    c%nbx = MPI_NBX;  c%nby = MPI_NBY;  c%nbz = MPI_NBZ
    c%cx = coords(0); c%cy = coords(1); c%cz = coords(2)
    c%nghost = nghost
    c%neq = neq; c%neqdyn = neqdyn; c%npas = npas
    c%mhd = merge(1, 0, mhd); c%pmhd = merge(1, 0, pmhd); c%passives = merge(1, 0, passives)
    c%riemann_solver = riemann_solver; c%slope_limiter = slope_limiter; c%eq_of_state = eq_of_state
    c%enable_flux_cd = merge(1, 0, enable_flux_cd); c%eight_wave = merge(1, 0, eight_wave)
    c%user_source_terms = merge(1, 0, user_source_terms)
    c%bc_left = bc_left; c%bc_right = bc_right; c%bc_bottom = bc_bottom
    c%bc_top = bc_top;   c%bc_out = bc_out;     c%bc_in = bc_in
    c%bc_user = merge(1, 0, bc_user)
    c%strict_fp = 0
    c%cooling = merge(cooling, 0, cooling == COOL_H)   ! the other cooling modules stay in the host
    c%th_cond = th_cond; c%tc_saturation = merge(1, 0, tc_saturation)
    c%pad_ = 0
    c%dx = dx; c%dy = dy; c%dz = dz
    c%cv = cv; c%gamma = gamma; c%Tempsc = Tempsc; c%cfl = cfl; c%eta = eta; c%tsc = tsc
    c%rsc = rsc; c%rhosc = rhosc; c%vsc2 = vsc2; c%bsc = bsc; c%mu = mu
    call gx_check(gx_create(c, gx_handle), 'gx_create')
  end subroutine gx_initmain

  !> slow path of get_user_source_terms for arbitrary user code: register with
  !>   call gx_check(gx_register_host_source(gx_handle, c_funloc(gx_host_source_tramp), c_null_ptr), 'host source')
  !> The library calls it once per stage with the block's primitives; the loop below is the one of
  !> step() (hydro_solver.f90:99-121) around the user's own get_user_source_terms (user_mod.f90).
  subroutine gx_host_source_tramp(primit_c, s_c, user) bind(C)
    use parameters, only : neq, nx, ny, nz, nxmin, nxmax, nymin, nymax, nzmin, nzmax
    use user_mod, only : get_user_source_terms
    type(c_ptr), value :: primit_c, s_c, user
    real(c_double), pointer :: pp(:,:,:,:), ss(:,:,:,:)
    integer :: i, j, k
    call c_f_pointer(primit_c, pp, [neq, nxmax-nxmin+1, nymax-nymin+1, nzmax-nzmin+1])
    call c_f_pointer(s_c, ss, [neq, nxmax-nxmin+1, nymax-nymin+1, nzmax-nzmin+1])
    do k = 1, nz
      do j = 1, ny
        do i = 1, nx          ! array index = Fortran index - nxmin + 1
          call get_user_source_terms(pp(:, i-nxmin+1, j-nymin+1, k-nzmin+1), ss(:, i-nxmin+1, j-nymin+1, k-nzmin+1), i, j, k)
        end do
      end do
    end do
  end subroutine gx_host_source_tramp

#ifdef MPIP
  !> replaces mpi_cart_create's role for the halo exchange: NCCL communicator over the same ranks
  subroutine gx_attach_comm()
    use globals, only : rank, comm3d
    use parameters, only : np, master
    include "mpif.h"
    character(kind=c_char) :: id(128)
    integer :: err
    if (rank == master) call gx_check(gx_comm_unique_id(id, 128_c_int32_t), 'gx_comm_unique_id')
    call mpi_bcast(id, 128, mpi_character, master, comm3d, err)
    call gx_check(gx_comm_attach(gx_handle, id, 128_c_int32_t, int(rank, c_int32_t), int(np, c_int32_t)), 'gx_comm_attach')
  end subroutine gx_attach_comm
#endif

end module guacho_gpu
