!> nhydro_mgb.f90 -- drop-in replacement of the reference's `module nhydro` (src/nhydro.f90) that forwards every
!> procedure to the B200 library libmgb.so through ISO_C_BINDING.  Same module name, same procedure names, same
!> argument lists, so the ocean model (or src/mg_testseamount.f90 etc.) links against this file instead of
!> src/nhydro.f90 + the solver sources (mg_relax, mg_intergrids, mg_define_matrix, mg_zr_zw, mg_compute_rhs,
!> mg_correct_uvw, mg_gather, mg_solvers).  (Not compiled in this repository: the build image has no Fortran
!> compiler; tests drive the same C entry points through ctypes with identically laid-out arrays.)
!>
!> Like src/nhydro.f90:3-8 it re-exports the reference's own support modules, which the stock drivers reach THROUGH
!> `use nhydro`: `myrank`/`nprocs` (mg_mpi, src/mg_mpi.f90:7-8), `rp`/`ip` and the namelist (mg_namelist), tic/toc
!> (mg_tictoc), `grid(:)` and the neighbour topology (mg_grids: the drivers read grid(1)%b, src/mg_testseamount.f90:201, and
!> mg_mpi_exchange_ijk needs grid(1)%neighb), `fill_halo_2D_bmask` (mg_mpi_exchange, src/mg_testseamount.f90:101) and
!> `write_netcdf` (mg_netcdf_out).  Those six files are compiled unchanged (INTEGRATION.md); they hold no solver arithmetic.
module nhydro

  use iso_c_binding
  use mpi
  use mg_mpi            ! myrank, nprocs, mg_mpi_init                      (src/mg_mpi.f90)
  use mg_namelist       ! rp, ip, read_nhnamelist, netcdf_output ...        (src/mg_namelist.f90)
  use mg_tictoc         ! tic, toc, print_tictoc                            (src/mg_tictoc.f90)
  use mg_grids          ! grid(:), define_grids, define_neighbours          (src/mg_grids.f90)
  use mg_mpi_exchange   ! fill_halo_2D_bmask for the drivers' rmask         (src/mg_mpi_exchange.f90)
  use mg_netcdf_out     ! write_netcdf (or a stub of it when NetCDF is absent and netcdf_output=.false.)

  implicit none

  interface
     integer(c_int) function mgb_init(nx, ny, nz, npx, npy, rank, namelist_path, device) bind(C, name='mgb_init')
       import :: c_int, c_char
       integer(c_int), value :: nx, ny, nz, npx, npy, rank, device
       character(kind=c_char), dimension(*) :: namelist_path
     end function mgb_init
     integer(c_int) function mgb_comm_unique_id(id) bind(C, name='mgb_comm_unique_id')
       import :: c_int, c_char
       character(kind=c_char), dimension(128) :: id
     end function mgb_comm_unique_id
     integer(c_int) function mgb_comm_init(id) bind(C, name='mgb_comm_init')
       import :: c_int, c_char
       character(kind=c_char), dimension(128) :: id
     end function mgb_comm_init
     integer(c_int) function mgb_matrices(dx, dy, zeta, h, rmask, hc, theta_b, theta_s) bind(C, name='mgb_matrices')
       import :: c_int, c_double
       real(c_double), dimension(*) :: dx, dy, zeta, h, rmask
       real(c_double), value :: hc, theta_b, theta_s
     end function mgb_matrices
     integer(c_int) function mgb_solve(rmask, u, v, w) bind(C, name='mgb_solve')
       import :: c_int, c_double
       real(c_double), dimension(*) :: rmask, u, v, w
     end function mgb_solve
     integer(c_int) function mgb_check_nondivergence(rmask, u, v, w) bind(C, name='mgb_check_nondivergence')
       import :: c_int, c_double
       real(c_double), dimension(*) :: rmask, u, v, w
     end function mgb_check_nondivergence
     integer(c_int) function mgb_write_fort100(path) bind(C, name='mgb_write_fort100')
       import :: c_int, c_char
       character(kind=c_char), dimension(*) :: path
     end function mgb_write_fort100
     integer(c_int) function mgb_clean() bind(C, name='mgb_clean')
       import :: c_int
     end function mgb_clean
     function mgb_last_error() bind(C, name='mgb_last_error') result(msg)
       import :: c_ptr
       type(c_ptr) :: msg
     end function mgb_last_error
  end interface

  interface
     integer(c_int) function mgb_get_field(lev, name, out) bind(C, name='mgb_get_field')
       import :: c_int, c_char, c_double
       integer(c_int), value :: lev
       character(kind=c_char), dimension(*) :: name
       real(c_double), dimension(*) :: out
     end function mgb_get_field
  end interface

contains

  subroutine mgb_check(rc, where)
    integer(c_int), intent(in) :: rc
    character(len=*), intent(in) :: where
    character(kind=c_char), pointer :: msg(:)
    if (rc /= 0) then                      ! the reference stops with a message (e.g. mg_grids.f90:529-530)
       call c_f_pointer(mgb_last_error(), msg, [256])
       write(*,*) 'Error in ', where, ': ', msg(1:index(transfer(msg, repeat(' ',256)), c_null_char)-1)
       stop -1
    endif
  end subroutine mgb_check

  !--------------------------------------------------------------
  subroutine nhydro_init(nx, ny, nz, npxg, npyg)          ! src/nhydro.f90:18
    integer(kind=ip), intent(in) :: nx, ny, nz, npxg, npyg
    character(kind=c_char), dimension(128) :: id
    integer :: ierr, ngpus_per_node
    call mg_mpi_init()                                      ! myrank, nprocs of module mg_mpi (src/mg_mpi.f90:12-36)
    call read_nhnamelist(vbrank=myrank)                     ! the drivers read netcdf_output etc. from the module (src/nhydro.f90:25)
    call define_grids(npxg, npyg, nx, ny, nz)               ! host-side grid(:) for the drivers: topology + grid(1)%b mirror
    call define_neighbours()                                ! (src/nhydro.f90:27-29); the solver state itself lives on the GPU
    ngpus_per_node = 8
    call mgb_check(mgb_init(nx, ny, nz, npxg, npyg, myrank, 'nh_namelist'//c_null_char, mod(myrank, ngpus_per_node)), 'nhydro_init')
    if (nprocs > 1) then                                   ! NCCL bootstrap over MPI: rank 0 creates the id, everybody joins
       if (myrank == 0) call mgb_check(mgb_comm_unique_id(id), 'mgb_comm_unique_id')
       call mpi_bcast(id, 128, mpi_character, 0, mpi_comm_world, ierr)
       call mgb_check(mgb_comm_init(id), 'mgb_comm_init')
    endif
  end subroutine nhydro_init

  !--------------------------------------------------------------
  subroutine nhydro_matrices(dx, dy, zeta, h, rmask, hc, theta_b, theta_s)   ! src/nhydro.f90:36
    real(kind=rp), dimension(:,:), intent(in) :: dx, dy, zeta, h               ! (0:ny+1,0:nx+1)
    real(kind=rp), dimension(:,:), pointer, intent(in) :: rmask
    real(kind=rp), intent(in) :: hc, theta_b, theta_s
    real(kind=rp), dimension(:,:), allocatable :: a, b, c, d, m                ! contiguous copies (assumed-shape in)
    allocate(a, source=dx); allocate(b, source=dy); allocate(c, source=zeta); allocate(d, source=h); allocate(m, source=rmask)
    call mgb_check(mgb_matrices(a, b, c, d, m, hc, theta_b, theta_s), 'nhydro_matrices')
  end subroutine nhydro_matrices

  !--------------------------------------------------------------
  subroutine nhydro_solve(nx, ny, nz, rmaska, ua, va, wa)                      ! src/nhydro.f90:53
    integer(kind=ip), intent(in) :: nx, ny, nz
    real(kind=rp), dimension(0:nx+1,0:ny+1),      target, intent(inout) :: rmaska
    real(kind=rp), dimension(1:nx+1,0:ny+1,1:nz), target, intent(inout) :: ua
    real(kind=rp), dimension(0:nx+1,1:ny+1,1:nz), target, intent(inout) :: va
    real(kind=rp), dimension(0:nx+1,0:ny+1,0:nz), target, intent(inout) :: wa
    call mgb_check(mgb_solve(rmaska, ua, va, wa), 'nhydro_solve')              ! explicit-shape = contiguous: passed as is
    if (myrank == 0) call mgb_check(mgb_write_fort100('fort.100'//c_null_char), 'fort.100')   ! mg_solvers.f90:59,72
  end subroutine nhydro_solve

  !--------------------------------------------------------------
  subroutine nhydro_check_nondivergence(nx, ny, nz, rmaska, ua, va, wa)        ! src/nhydro.f90:105
    integer(kind=ip), intent(in) :: nx, ny, nz
    real(kind=rp), dimension(0:nx+1,0:ny+1),      target, intent(inout) :: rmaska
    real(kind=rp), dimension(1:nx+1,0:ny+1,1:nz), target, intent(inout) :: ua
    real(kind=rp), dimension(0:nx+1,1:ny+1,1:nz), target, intent(inout) :: va
    real(kind=rp), dimension(0:nx+1,0:ny+1,0:nz), target, intent(inout) :: wa
    call mgb_check(mgb_check_nondivergence(rmaska, ua, va, wa), 'nhydro_check_nondivergence')
    ! the drivers dump grid(1)%b afterwards (src/mg_testseamount.f90:201): refresh the host mirror, same (k,j,i) layout
    call mgb_check(mgb_get_field(1, 'b'//c_null_char, grid(1)%b), 'mgb_get_field(b)')
  end subroutine nhydro_check_nondivergence

  !--------------------------------------------------------------
  subroutine nhydro_clean()                                                    ! src/nhydro.f90:137
    call mgb_check(mgb_clean(), 'nhydro_clean')
    call grids_dealloc()                                    ! src/nhydro.f90:139
  end subroutine nhydro_clean

end module nhydro
