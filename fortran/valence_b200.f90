! ISO_C_BINDING interface to libvalence_b200.so (the C-ABI of include/valence_b200.h) for a Fortran host.
!
! Two ways to use the library from VALENCE's own Fortran:
!   (1) unchanged: the external procedures valence_api_initialize / valence_api_calculate_energy / valence_api_finalize
!       and init / getn / calcsurface / finalize (src/valence_api.F90:9,37,114, src/valence_api_nitrogen.F90:4-36) are
!       exported with gfortran's mangling -- link -lvalence_b200 instead of -lvalence, nothing else changes;
!   (2) keep VALENCE's control flow (input parser, minimize_energy, xm_output) and replace only the hot path: this
!       module.  In guess_energy (src/valence.F90:337-344) the calls  wfndet / schwarz_ints / vsvb_energy /
!       xm_equalize_scalar x2  become one vb_engine_energy; in first_order_opt (:666-764) the (ib,jb) loop and the two
!       xm_equalize(ham/ovl) become one vb_engine_first_order (column-major norbas x norbas, what rsg receives at :771).
!
! Parallel runs (the replacement of src/xm_module.F90:711-910): one process per GPU; after
!   ierr = vb_engine_attach_comm(eng, rank, nranks, c_null_char)
! the three calls above are collective -- the ranks form an NCCL communicator inside the library, the tile pass is
! sharded and ONE all-reduce per energy (one per orbital for ham) replaces the reference's MPI_Allreduce calls.  A host that
! already owns an ncclComm_t passes it with vb_engine_attach_nccl.
!
! This file cannot be compiled in the build image (no Fortran compiler there); it is kept in step with the header by
! tests/test_capi_symbols.py (every procedure bound here must be declared in include/valence_b200.h).
module valence_b200
  use iso_c_binding
  implicit none

  integer, parameter :: vb_cnt_schwarz_erep = 1, vb_cnt_schwarz_exch = 2, vb_cnt_value_erep = 3, vb_cnt_value_exch = 4, &
                        vb_cnt_int2e = 5, vb_cnt_shell_quartets = 6, vb_cnt_shortcut = 7, vb_cnt_entries = 8

  type, bind(c) :: vb_energy_result
     real(c_double) :: energy, enucrep, numerator, wfnorm, e1, e2     ! energy = numerator/wfnorm + enucrep (valence.F90:344)
     integer(c_long_long) :: counters(8)                              ! screening / quartet counts, reference task semantics
     integer(c_long_long) :: n_entries, n_groups, n_pairgroups, n_tiles, n_tiles_mine
     integer(c_long_long) :: n_ao_quartets, n_prim_quartets
     real(c_double) :: flops_model
     integer(c_long_long) :: ref_shell_quartets
     real(c_double) :: t_total_ms, t_host_setup_ms, t_1e_ms, t_density_ms, t_diag_ms, t_tiles_ms
     integer(c_int) :: launches, diag_launches, tile_launches
     real(c_double) :: min_pivot_ratio
     integer(c_long_long) :: h2d_bytes, d2h_bytes
     real(c_double) :: flops_transform
  end type vb_energy_result

  interface
     function vb_last_error() bind(c, name="vb_last_error") result(msg)
       import; type(c_ptr) :: msg
     end function
     integer(c_int) function vb_engine_create(path, device, eng) bind(c, name="vb_engine_create")
       import; character(kind=c_char), intent(in) :: path(*)
       integer(c_int), value :: device; type(c_ptr), intent(out) :: eng
     end function
     subroutine vb_engine_destroy(eng) bind(c, name="vb_engine_destroy")
       import; type(c_ptr), value :: eng
     end subroutine
     integer(c_int) function vb_engine_natom(eng) bind(c, name="vb_engine_natom")
       import; type(c_ptr), value :: eng
     end function
     integer(c_int) function vb_engine_nelec(eng) bind(c, name="vb_engine_nelec")
       import; type(c_ptr), value :: eng
     end function
     integer(c_int) function vb_engine_norbas(eng, iorb) bind(c, name="vb_engine_norbas")
       import; type(c_ptr), value :: eng; integer(c_int), value :: iorb
     end function
     integer(c_int) function vb_engine_attach_comm(eng, rank, nranks, key) bind(c, name="vb_engine_attach_comm")
       import; type(c_ptr), value :: eng; integer(c_int), value :: rank, nranks
       character(kind=c_char), intent(in) :: key(*)
     end function
     integer(c_int) function vb_engine_attach_nccl(eng, rank, nranks, nccl_comm) bind(c, name="vb_engine_attach_nccl")
       import; type(c_ptr), value :: eng, nccl_comm; integer(c_int), value :: rank, nranks
     end function
     integer(c_int) function vb_engine_set_coords(eng, x) bind(c, name="vb_engine_set_coords")
       import; type(c_ptr), value :: eng; real(c_double), intent(in) :: x(*)      ! 3*natom, atom-major, Angstrom
     end function
     integer(c_int) function vb_engine_energy(eng, res) bind(c, name="vb_engine_energy")
       import; type(c_ptr), value :: eng; type(vb_energy_result), intent(out) :: res
     end function
     integer(c_int) function vb_engine_first_order(eng, iorb, ham, ovl, cap, norbas, res) bind(c, name="vb_engine_first_order")
       import; type(c_ptr), value :: eng; integer(c_int), value :: iorb, cap
       real(c_double), intent(out) :: ham(*), ovl(*); integer(c_int), intent(out) :: norbas
       type(vb_energy_result), intent(out) :: res
     end function
     integer(c_int) function vb_engine_run(eng, print, enucrep, guess_energy, total_energy, converged, iterations) &
          bind(c, name="vb_engine_run")
       import; type(c_ptr), value :: eng; integer(c_int), value :: print
       real(c_double), intent(out) :: enucrep, guess_energy, total_energy
       integer(c_int), intent(out) :: converged, iterations
     end function
     ! a host that does its own reduction (MPI_Allreduce on a copy, or NCCL on the device pointer) instead of attach_comm:
     integer(c_int) function vb_engine_energy_partial(eng, rank, nranks, res) bind(c, name="vb_engine_energy_partial")
       import; type(c_ptr), value :: eng; integer(c_int), value :: rank, nranks; type(vb_energy_result), intent(out) :: res
     end function
     integer(c_int) function vb_engine_energy_finish(eng, res) bind(c, name="vb_engine_energy_finish")
       import; type(c_ptr), value :: eng; type(vb_energy_result), intent(inout) :: res
     end function
     function vb_engine_accum_device(eng) bind(c, name="vb_engine_accum_device") result(p)
       import; type(c_ptr), value :: eng; type(c_ptr) :: p
     end function
     integer(c_int) function vb_engine_accum_len(eng) bind(c, name="vb_engine_accum_len")
       import; type(c_ptr), value :: eng
     end function
     integer(c_int) function vb_engine_first_order_sharded(eng, iorb, rank, nranks, ham, ovl, cap, norbas, res) &
          bind(c, name="vb_engine_first_order_sharded")
       import; type(c_ptr), value :: eng; integer(c_int), value :: iorb, rank, nranks, cap
       real(c_double), intent(out) :: ham(*), ovl(*); integer(c_int), intent(out) :: norbas
       type(vb_energy_result), intent(out) :: res
     end function
  end interface

contains

  ! guess_energy (src/valence.F90:309-345) on the GPU: energy = numerator/wfnorm + enucrep
  subroutine vb_guess_energy(eng, coords_angstrom, energy, ierr)
    type(c_ptr), intent(in) :: eng
    real(c_double), intent(in) :: coords_angstrom(*)
    real(c_double), intent(out) :: energy
    integer, intent(out) :: ierr
    type(vb_energy_result) :: res
    ierr = vb_engine_set_coords(eng, coords_angstrom)
    if (ierr /= 0) return
    ierr = vb_engine_energy(eng, res)
    energy = res%energy
  end subroutine vb_guess_energy

  ! the ham / ovl build of first_order_opt (src/valence.F90:674-764) for orbital iorb; hdim = leading dimension of the
  ! caller's ham / ovl (hdim = max(xpmax, nspinc), src/valence.F90:225)
  subroutine vb_first_order_matrices(eng, iorb, hdim, ham, ovl, norbas, ierr)
    type(c_ptr), intent(in) :: eng
    integer, intent(in) :: iorb, hdim
    real(c_double), intent(out) :: ham(hdim, hdim), ovl(hdim, hdim)
    integer, intent(out) :: norbas, ierr
    real(c_double), allocatable :: h(:), s(:)
    type(vb_energy_result) :: res
    integer(c_int) :: n
    integer :: i, j
    n = vb_engine_norbas(eng, int(iorb, c_int))
    allocate(h(n*n), s(n*n))
    ierr = vb_engine_first_order(eng, int(iorb, c_int), h, s, n*n, n, res)
    norbas = n
    if (ierr == 0) then
       do j = 1, n
          do i = 1, n
             ham(i, j) = h((j-1)*n + i)
             ovl(i, j) = s((j-1)*n + i)
          end do
       end do
    end if
    deallocate(h, s)
  end subroutine vb_first_order_matrices

end module valence_b200
