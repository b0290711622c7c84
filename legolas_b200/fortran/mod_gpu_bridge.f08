! =============================================================================
!> iso_c_binding bridge between the Legolas Fortran host and liblegolas_b200.so.
!! NOT compiled in the build container (no Fortran compiler there); this is the
!! binding a maintainer adds to the reference tree (see INTEGRATION.md):
!!   - build_matrices_gpu replaces the body of build_matrices
!!       (src/matrices/mod_matrix_manager.f08:138-266)
!!   - solve_arpack_shift_invert_gpu replaces solve_arpack_shift_invert
!!       (src/solvers/arnoldi/smod_arpack_shift_invert.f08:15-161)
module mod_gpu_bridge
  use, intrinsic :: iso_c_binding
  use mod_global_variables, only: dp
  implicit none

  private

  integer, parameter :: LGPU_N_FIELDS = 38

  type, bind(C) :: lgpu_settings
    integer(c_int32_t) :: gridpts, physics_type, geometry, incompressible
    integer(c_int32_t) :: flow, resistivity, cooling, heating, conduction
    integer(c_int32_t) :: perpendicular_conduction, viscosity, viscous_heating
    integer(c_int32_t) :: hall, electron_inertia, gravity, boundary_type, coaxial
    integer(c_int32_t) :: reserved
    real(c_double) :: k2, k3, gamma, viscosity_value, electron_fraction
    real(c_double) :: gauss_nodes(4), gauss_weights(4)
  end type lgpu_settings

  type, bind(C) :: lgpu_arnoldi
    integer(c_int32_t) :: nev, ncv, maxiter
    character(kind=c_char) :: which(2)
    integer(c_int16_t) :: pad
    real(c_double) :: tol, sigma_re, sigma_im
    integer(c_int32_t) :: refine_steps, reserved
  end type lgpu_arnoldi

  type, bind(C) :: lgpu_stats
    integer(c_int32_t) :: info, nconv, n_op, n_bx, n_reorth, n_restart, lu_info, reserved
    real(c_double) :: t_factor_ms, t_iter_ms, t_extract_ms
  end type lgpu_stats

  interface
    integer(c_int) function lgpu_create(ctx, device, log_level) bind(C, name="lgpu_create")
      import :: c_ptr, c_int, c_int32_t
      type(c_ptr), intent(out) :: ctx
      integer(c_int32_t), value :: device, log_level
    end function lgpu_create

    integer(c_int) function lgpu_destroy(ctx) bind(C, name="lgpu_destroy")
      import :: c_ptr, c_int
      type(c_ptr), value :: ctx
    end function lgpu_destroy

    integer(c_int) function lgpu_assemble(ctx, settings, base_grid, gauss_grid, fields) &
      bind(C, name="lgpu_assemble")
      import :: c_ptr, c_int, c_double, lgpu_settings
      type(c_ptr), value :: ctx
      type(lgpu_settings), intent(in) :: settings
      real(c_double), intent(in) :: base_grid(*), gauss_grid(*)
      type(c_ptr), intent(in) :: fields(*)
    end function lgpu_assemble

    integer(c_int) function lgpu_export_coo(ctx, which, nnz, rows, cols, vals) &
      bind(C, name="lgpu_export_coo")
      import :: c_ptr, c_int, c_int32_t, c_int64_t
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: which
      integer(c_int64_t), intent(inout) :: nnz
      type(c_ptr), value :: rows, cols, vals
    end function lgpu_export_coo

    integer(c_int) function lgpu_shift_invert(ctx, cfg, resid0, omega, vr, stats) &
      bind(C, name="lgpu_shift_invert")
      import :: c_ptr, c_int, c_double_complex, lgpu_arnoldi, lgpu_stats
      type(c_ptr), value :: ctx
      type(lgpu_arnoldi), intent(in) :: cfg
      complex(c_double_complex), intent(in) :: resid0(*)
      complex(c_double_complex), intent(out) :: omega(*)
      complex(c_double_complex), intent(out) :: vr(*)
      type(lgpu_stats), intent(out) :: stats
    end function lgpu_shift_invert

    integer(c_int) function lgpu_arnoldi_general(ctx, cfg, resid0, omega, vr, stats) &
      bind(C, name="lgpu_arnoldi_general")
      import :: c_ptr, c_int, c_double_complex, lgpu_arnoldi, lgpu_stats
      type(c_ptr), value :: ctx
      type(lgpu_arnoldi), intent(in) :: cfg
      complex(c_double_complex), intent(in) :: resid0(*)
      complex(c_double_complex), intent(out) :: omega(*)
      complex(c_double_complex), intent(out) :: vr(*)
      type(lgpu_stats), intent(out) :: stats
    end function lgpu_arnoldi_general

    integer(c_int) function lgpu_inverse_iteration(ctx, sigma_re, sigma_im, maxiter, tol, omega, vr, stats) &
      bind(C, name="lgpu_inverse_iteration")
      import :: c_ptr, c_int, c_double, c_double_complex, lgpu_stats
      type(c_ptr), value :: ctx
      real(c_double), value :: sigma_re, sigma_im, tol
      integer(c_int), value :: maxiter
      complex(c_double_complex), intent(out) :: omega(*)
      complex(c_double_complex), intent(out) :: vr(*)
      type(lgpu_stats), intent(out) :: stats
    end function lgpu_inverse_iteration

    integer(c_int) function lgpu_eigenfunctions(ctx, vr, nsel, idxs, out) bind(C, name="lgpu_eigenfunctions")
      import :: c_ptr, c_int, c_double_complex
      type(c_ptr), value :: ctx
      complex(c_double_complex), intent(in) :: vr(*)
      integer(c_int), value :: nsel
      integer(c_int), intent(in) :: idxs(*)
      complex(c_double_complex), intent(out) :: out(*)
    end function lgpu_eigenfunctions

    integer(c_int) function lgpu_residuals(ctx, nev, omega, vr, res) bind(C, name="lgpu_residuals")
      import :: c_ptr, c_int, c_double, c_double_complex
      type(c_ptr), value :: ctx
      integer(c_int), value :: nev
      complex(c_double_complex), intent(in) :: omega(*)
      complex(c_double_complex), intent(in) :: vr(*)
      real(c_double), intent(out) :: res(*)
    end function lgpu_residuals
  end interface

  !> one context per run: owns the device-resident A, B, factors and Krylov basis
  type(c_ptr), save :: gpu_ctx = c_null_ptr

  public :: build_matrices_gpu, solve_arpack_shift_invert_gpu, materialise_matrix_gpu
  public :: solve_arpack_general_gpu
  public :: inverse_iteration_gpu, residuals_gpu, base_eigenfunctions_gpu

contains

  !> Samples the 38 background / physics procedure pointers at grid%gaussian_grid
  !! (same accessors as src/dataIO/mod_output.f08:354-400) and assembles A and B on the GPU.
  subroutine build_matrices_gpu(settings, grid, background, physics)
    use mod_settings, only: settings_t
    use mod_grid, only: grid_t
    use mod_background, only: background_t
    use mod_physics, only: physics_t
    use mod_function_utils, only: from_function
    use mod_equilibrium_params, only: k2, k3
    use mod_global_variables, only: gaussian_nodes, gaussian_weights
    use mod_logging, only: logger, str
    type(settings_t), intent(in) :: settings
    type(grid_t), intent(in) :: grid
    type(background_t), intent(in) :: background
    type(physics_t), intent(in) :: physics

    real(dp), allocatable, target :: f(:, :)
    type(c_ptr) :: ptrs(LGPU_N_FIELDS)
    type(lgpu_settings) :: cs
    integer :: ng, i, rc

    if (.not. c_associated(gpu_ctx)) then
      rc = lgpu_create(gpu_ctx, 0_c_int32_t, int(logger%get_logging_level(), c_int32_t))
      if (rc /= 0) then
        call logger%error("legolas_b200: no CUDA device (lgpu_create = " // str(rc) // ")")
        return
      end if
    end if

    ng = settings%grid%get_gauss_gridpts()
    allocate(f(ng, LGPU_N_FIELDS))
    f(:, 1) = from_function(background%density%rho0, grid%gaussian_grid)
    f(:, 2) = from_function(background%density%drho0, grid%gaussian_grid)
    f(:, 3) = from_function(background%temperature%T0, grid%gaussian_grid)
    f(:, 4) = from_function(background%temperature%dT0, grid%gaussian_grid)
    f(:, 5) = from_function(background%temperature%ddT0, grid%gaussian_grid)
    f(:, 6) = from_function(background%magnetic%B01, grid%gaussian_grid)
    f(:, 7) = from_function(background%magnetic%B02, grid%gaussian_grid)
    f(:, 8) = from_function(background%magnetic%dB02, grid%gaussian_grid)
    f(:, 9) = from_function(background%magnetic%ddB02, grid%gaussian_grid)
    f(:, 10) = from_function(background%magnetic%B03, grid%gaussian_grid)
    f(:, 11) = from_function(background%magnetic%dB03, grid%gaussian_grid)
    f(:, 12) = from_function(background%magnetic%ddB03, grid%gaussian_grid)
    f(:, 13) = from_function(background%velocity%v01, grid%gaussian_grid)
    f(:, 14) = from_function(background%velocity%dv01, grid%gaussian_grid)
    f(:, 15) = from_function(background%velocity%ddv01, grid%gaussian_grid)
    f(:, 16) = from_function(background%velocity%v02, grid%gaussian_grid)
    f(:, 17) = from_function(background%velocity%dv02, grid%gaussian_grid)
    f(:, 18) = from_function(background%velocity%ddv02, grid%gaussian_grid)
    f(:, 19) = from_function(background%velocity%v03, grid%gaussian_grid)
    f(:, 20) = from_function(background%velocity%dv03, grid%gaussian_grid)
    f(:, 21) = from_function(background%velocity%ddv03, grid%gaussian_grid)
    f(:, 22) = from_function(physics%gravity%g0, grid%gaussian_grid)
    f(:, 23) = from_function(physics%resistivity%eta, grid%gaussian_grid)
    f(:, 24) = from_function(physics%resistivity%detadT, grid%gaussian_grid)
    f(:, 25) = from_function(physics%resistivity%detadr, grid%gaussian_grid)
    f(:, 26) = physics%heatloss%get_L0(grid%gaussian_grid)
    f(:, 27) = physics%heatloss%get_dLdT(grid%gaussian_grid)
    f(:, 28) = physics%heatloss%get_dLdrho(grid%gaussian_grid)
    f(:, 29) = from_function(physics%conduction%tcpara, grid%gaussian_grid)
    f(:, 30) = from_function(physics%conduction%dtcparadT, grid%gaussian_grid)
    f(:, 31) = from_function(physics%conduction%tcperp, grid%gaussian_grid)
    f(:, 32) = from_function(physics%conduction%dtcperpdrho, grid%gaussian_grid)
    f(:, 33) = from_function(physics%conduction%dtcperpdT, grid%gaussian_grid)
    f(:, 34) = from_function(physics%conduction%dtcperpdB2, grid%gaussian_grid)
    f(:, 35) = physics%conduction%get_tcprefactor(grid%gaussian_grid)
    f(:, 36) = physics%conduction%get_dtcprefactordr(grid%gaussian_grid)
    f(:, 37) = from_function(physics%hall%hallfactor, grid%gaussian_grid)
    f(:, 38) = from_function(physics%hall%inertiafactor, grid%gaussian_grid)
    do i = 1, LGPU_N_FIELDS
      ptrs(i) = c_loc(f(1, i))
    end do

    cs%gridpts = settings%grid%get_gridpts()
    select case(settings%get_physics_type())
    case("hd"); cs%physics_type = 1
    case("hd-1d"); cs%physics_type = 2
    case default; cs%physics_type = 0
    end select
    cs%geometry = merge(1, 0, settings%grid%get_geometry() == "cylindrical")
    cs%incompressible = merge(1, 0, settings%physics%is_incompressible)
    cs%flow = merge(1, 0, settings%physics%flow%is_enabled())
    cs%resistivity = merge(1, 0, settings%physics%resistivity%is_enabled())
    cs%cooling = merge(1, 0, settings%physics%cooling%is_enabled())
    cs%heating = merge(1, 0, settings%physics%heating%is_enabled())
    cs%conduction = merge(1, 0, settings%physics%conduction%is_enabled())
    cs%perpendicular_conduction = &
      merge(1, 0, settings%physics%conduction%has_perpendicular_conduction())
    cs%viscosity = merge(1, 0, settings%physics%viscosity%is_enabled())
    cs%viscous_heating = merge(1, 0, settings%physics%viscosity%has_viscous_heating())
    cs%hall = merge(1, 0, settings%physics%hall%is_enabled())
    cs%electron_inertia = merge(1, 0, settings%physics%hall%has_electron_inertia())
    cs%gravity = merge(1, 0, settings%physics%gravity%is_enabled())
    cs%boundary_type = merge(1, 0, settings%equilibrium%get_boundary_type() == "wall_weak")
    cs%coaxial = merge(1, 0, settings%grid%coaxial)
    cs%reserved = 0
    cs%k2 = k2
    cs%k3 = k3
    cs%gamma = settings%physics%get_gamma()
    cs%viscosity_value = settings%physics%viscosity%get_viscosity_value()
    cs%electron_fraction = settings%physics%hall%get_electron_fraction()
    cs%gauss_nodes = gaussian_nodes
    cs%gauss_weights = gaussian_weights

    rc = lgpu_assemble(gpu_ctx, cs, grid%base_grid, grid%gaussian_grid, ptrs)
    if (rc /= 0) call logger%error("legolas_b200: lgpu_assemble failed with " // str(rc))
    deallocate(f)
  end subroutine build_matrices_gpu


  !> Drop-in for solve_arpack_shift_invert: arpack_cfg has been built by new_arpack_config
  !! (validation, ncv / maxiter defaults, zlarnv start vector) exactly as before.
  subroutine solve_arpack_shift_invert_gpu(arpack_cfg, settings, omega, vr)
    use mod_arpack_type, only: arpack_t
    use mod_settings, only: settings_t
    use mod_logging, only: logger, str
    type(arpack_t), intent(inout) :: arpack_cfg
    type(settings_t), intent(in) :: settings
    complex(dp), intent(out) :: omega(:)
    complex(dp), intent(out) :: vr(:, :)

    type(lgpu_arnoldi) :: ca
    type(lgpu_stats) :: st
    character(len=2) :: which
    logical :: converged
    integer :: rc

    ca%nev = arpack_cfg%get_nev()
    ca%ncv = arpack_cfg%get_ncv()
    ca%maxiter = arpack_cfg%get_maxiter()
    which = arpack_cfg%get_which()
    ca%which(1) = which(1:1)
    ca%which(2) = which(2:2)
    ca%pad = 0
    ca%tol = arpack_cfg%get_tolerance()
    ca%sigma_re = real(settings%solvers%sigma)
    ca%sigma_im = aimag(settings%solvers%sigma)
    ca%refine_steps = 0
    ca%reserved = 0

    rc = lgpu_shift_invert(gpu_ctx, ca, arpack_cfg%residual, omega, vr, st)
    if (rc /= 0) then
      call logger%error("legolas_b200: lgpu_shift_invert failed with " // str(rc))
      return
    end if
    if (st%lu_info /= 0) call logger%warning("factorisation: singular pivot, info = " // str(st%lu_info))
    ! what znaupd / zneupd would have left behind, so the reference's parsers work unchanged
    arpack_cfg%info = st%info
    arpack_cfg%iparam(5) = st%nconv
    arpack_cfg%iparam(9) = st%n_op
    arpack_cfg%iparam(10) = st%n_bx
    arpack_cfg%iparam(11) = st%n_reorth
    call arpack_cfg%parse_znaupd_info(converged)
    arpack_cfg%info = 0
    call arpack_cfg%parse_zneupd_info()
    call arpack_cfg%parse_finished_stats()
  end subroutine solve_arpack_shift_invert_gpu


  !> Drop-in for solve_arpack_general (src/solvers/arnoldi/smod_arpack_general.f08:14-131):
  !! OP = B^-1 A, arpack_cfg built by new_arpack_config(mode=1, bmat="I") as before.
  subroutine solve_arpack_general_gpu(arpack_cfg, omega, vr)
    use mod_arpack_type, only: arpack_t
    use mod_logging, only: logger, str
    type(arpack_t), intent(inout) :: arpack_cfg
    complex(dp), intent(out) :: omega(:)
    complex(dp), intent(out) :: vr(:, :)

    type(lgpu_arnoldi) :: ca
    type(lgpu_stats) :: st
    character(len=2) :: which
    logical :: converged
    integer :: rc

    ca%nev = arpack_cfg%get_nev()
    ca%ncv = arpack_cfg%get_ncv()
    ca%maxiter = arpack_cfg%get_maxiter()
    which = arpack_cfg%get_which()
    ca%which(1) = which(1:1)
    ca%which(2) = which(2:2)
    ca%pad = 0
    ca%tol = arpack_cfg%get_tolerance()
    ca%sigma_re = 0.0d0
    ca%sigma_im = 0.0d0
    ca%refine_steps = 0
    ca%reserved = 0

    rc = lgpu_arnoldi_general(gpu_ctx, ca, arpack_cfg%residual, omega, vr, st)
    if (rc /= 0) then
      call logger%error("legolas_b200: lgpu_arnoldi_general failed with " // str(rc))
      return
    end if
    arpack_cfg%info = st%info
    arpack_cfg%iparam(5) = st%nconv
    arpack_cfg%iparam(9) = st%n_op
    arpack_cfg%iparam(10) = st%n_bx
    arpack_cfg%iparam(11) = st%n_reorth
    call arpack_cfg%parse_znaupd_info(converged)
    arpack_cfg%info = 0
    call arpack_cfg%parse_zneupd_info()
    call arpack_cfg%parse_finished_stats()
  end subroutine solve_arpack_general_gpu


  !> Drop-in for the body of inverse_iteration (src/solvers/smod_inverse_iteration.f08:16-205);
  !! the argument checks of the reference stay in front of it.
  subroutine inverse_iteration_gpu(settings, omega, vr)
    use mod_settings, only: settings_t
    use mod_logging, only: logger, str
    type(settings_t), intent(in) :: settings
    complex(dp), intent(out) :: omega(:)
    complex(dp), intent(out) :: vr(:, :)
    type(lgpu_stats) :: st
    integer :: rc

    rc = lgpu_inverse_iteration( &
      gpu_ctx, real(settings%solvers%sigma), aimag(settings%solvers%sigma), &
      settings%solvers%maxiter, settings%solvers%tolerance, omega, vr, st &
    )
    if (rc /= 0) then
      call logger%error("legolas_b200: lgpu_inverse_iteration failed with " // str(rc))
      return
    end if
    call logger%info("Iteration completed after " // str(st%n_op) // " iterations.")
    if (st%info /= 0) call logger%warning("Inverse iteration failed to converge! (maxiter reached)")
  end subroutine inverse_iteration_gpu


  !> Drop-in for the loop over base_efs(i)%assemble in eigenfunctions_t%assemble: quantities(:, :, p)
  !! is base_efs(p)%quantities for the p-th state-vector entry.
  subroutine base_eigenfunctions_gpu(idxs_to_assemble, right_eigenvectors, quantities)
    use mod_logging, only: logger, str
    integer, intent(in) :: idxs_to_assemble(:)
    complex(dp), intent(in) :: right_eigenvectors(:, :)
    complex(dp), intent(out) :: quantities(:, :, :)   ! (ef_gridpts, size(idxs), 8)
    integer :: rc

    rc = lgpu_eigenfunctions(gpu_ctx, right_eigenvectors, size(idxs_to_assemble), idxs_to_assemble, quantities)
    if (rc /= 0) call logger%error("legolas_b200: lgpu_eigenfunctions failed with " // str(rc))
  end subroutine base_eigenfunctions_gpu


  !> Drop-in for the loop of write_residual_data (src/dataIO/mod_output.f08:445-473).
  subroutine residuals_gpu(eigenvalues, eigenvectors, residuals)
    use mod_logging, only: logger, str
    complex(dp), intent(in) :: eigenvalues(:)
    complex(dp), intent(in) :: eigenvectors(:, :)
    real(dp), intent(out) :: residuals(:)
    integer :: rc

    rc = lgpu_residuals(gpu_ctx, size(eigenvalues), eigenvalues, eigenvectors, residuals)
    if (rc /= 0) call logger%error("legolas_b200: lgpu_residuals failed with " // str(rc))
  end subroutine residuals_gpu


  !> Lazily rebuilds a matrix_t from the device matrix (only needed for write_matrices,
  !! write_residuals or a non-GPU solver): triplets arrive in the reference's insertion order.
  subroutine materialise_matrix_gpu(which, matrix)
    use mod_matrix_structure, only: matrix_t
    integer, intent(in) :: which  ! 0 = A, 1 = B
    type(matrix_t), intent(inout) :: matrix
    integer(c_int64_t) :: nnz, k
    integer(c_int32_t), allocatable, target :: rows(:), cols(:)
    complex(dp), allocatable, target :: vals(:)
    integer :: rc

    nnz = 0
    rc = lgpu_export_coo(gpu_ctx, int(which, c_int32_t), nnz, c_null_ptr, c_null_ptr, c_null_ptr)
    allocate(rows(nnz), cols(nnz), vals(nnz))
    rc = lgpu_export_coo(gpu_ctx, int(which, c_int32_t), nnz, c_loc(rows), c_loc(cols), c_loc(vals))
    do k = 1, nnz
      call matrix%add_element(row=int(rows(k)), column=int(cols(k)), element=vals(k))
    end do
  end subroutine materialise_matrix_gpu

end module mod_gpu_bridge
