!> nhydro_mgb.f90 -- drop-in replacement of the reference's `module nhydro` (src/nhydro.f90) that forwards every
!> procedure to the B200 library libmgb.so through ISO_C_BINDING.  Same module name, same procedure names, same
!> argument lists, so the ocean model (or src/mg_testseamount.f90 etc.) links against this file instead of
!> src/nhydro.f90 + the mg_*.f90 solver sources.  (Not compiled in this repository: the build image has no Fortran
!> compiler; tests drive the same C entry points through ctypes with identically laid-out arrays.)
module nhydro

  use iso_c_binding
  use mpi

  implicit none

  integer(kind=4), parameter :: rp = 8, ip = 4

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

  integer(kind=ip) :: myrank = 0, nprocs = 1

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
    call mpi_comm_rank(mpi_comm_world, myrank, ierr)       ! mg_mpi_init, src/mg_mpi.f90:12-36
    call mpi_comm_size(mpi_comm_world, nprocs, ierr)
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
  end subroutine nhydro_check_nondivergence

  !--------------------------------------------------------------
  subroutine nhydro_clean()                                                    ! src/nhydro.f90:137
    call mgb_check(mgb_clean(), 'nhydro_clean')
  end subroutine nhydro_clean

end module nhydro
