! b200_module.f90 -- Fortran binding of the batched C ABI of libgimic_b200.so (include/gimic_b200.h).
!
! This is the file a GIMIC maintainer adds to src/fgimic/ (and to src/fgimic/CMakeLists.txt) to put the CUDA hot path
! behind the existing Fortran driver; see INTEGRATION.md for the three call sites that change
! (jfield.f90:114-129, integral.f90:107-165 / 245-304 / 462-499, gimic.F90:141-160).
! It cannot be compiled in the development container of this repository (no Fortran compiler there); it is exercised
! indirectly: tests/ drive exactly these entry points through ctypes with the same argument conventions.
module b200_module
    use iso_c_binding
    implicit none

    integer(c_int), parameter :: B200_ALPHA = 0, B200_BETA = 1, B200_TOTAL = 2, B200_SPINDENS = 3

    type, bind(c) :: b200_opts
        integer(c_int) :: uhf, giao, diamag, paramag, screening
        real(c_double) :: screening_thrs
        integer(c_int) :: device, spherical
    end type

    type, bind(c) :: b200_grid          ! gimic_b200_grid: what gridpoint()/get_weight() need of grid_t
        real(c_double) :: origin(3)
        real(c_double) :: basv(9)       ! basv(c + 3*(v-1)) = grid%basv(c, v)
        integer(c_int) :: npts(3)
        type(c_ptr) :: pts(3)           ! c_loc(grid%gdata(d)%pts)
        type(c_ptr) :: wgt(3)           ! c_loc(grid%gdata(d)%wgt)
        real(c_double) :: radius
    end type

    type(c_ptr), save :: b200_handle = c_null_ptr

    type, bind(c) :: b200_run_opts
        integer(c_int) :: flags, device, ndevices
        type(c_ptr) :: devices, workdir, title, report_path      ! c_null_ptr = default
    end type
    interface
        subroutine gimic_b200_default_opts(opts) bind(c)
            import; type(b200_opts) :: opts
        end subroutine
        integer(c_int) function gimic_b200_create(h, mol, xdens, opts) bind(c)
            import; type(c_ptr) :: h; character(kind=c_char) :: mol(*), xdens(*); type(b200_opts) :: opts
        end function
        integer(c_int) function gimic_b200_destroy(h) bind(c)
            import; type(c_ptr), value :: h
        end function
        ! tens(9,n): tens(m + 3*(b-1), i) = dJ_m/dB_b at r(:,i)  -- the layout of jfield_t%tens
        integer(c_int) function gimic_b200_calc_jtensors(h, n, r, spincase, tens, flags) bind(c)
            import; type(c_ptr), value :: h; integer(c_long), value :: n
            real(c_double) :: r(3,*), tens(9,*); integer(c_int), value :: spincase, flags
        end function
        integer(c_int) function gimic_b200_calc_jtensors_grid(h, g, lo, hi, spincase, tens, flags) bind(c)
            import; type(c_ptr), value :: h; type(b200_grid) :: g; integer(c_long), value :: lo, hi
            real(c_double) :: tens(9,*); integer(c_int), value :: spincase, flags
        end function
        ! out7 = (J, J+, J-, |J|, |J|+, |J|-, sum w*ACID) over plane rows jlo+1..jhi; what: 1 current, 2 modulus, 4 acid
        integer(c_int) function gimic_b200_integrate(h, g, b, spincase, what, jlo, jhi, out7) bind(c)
            import; type(c_ptr), value :: h; type(b200_grid) :: g; real(c_double) :: b(3), out7(7)
            integer(c_int), value :: spincase, what, jlo, jhi
        end function
        ! the same sums for ngrids grids in one tensor pass (current-profile scans): b(3,ngrids), out7(7,ngrids)
        integer(c_int) function gimic_b200_integrate_batch(h, ngrids, grids, b, spincase, what, out7) bind(c)
            import; type(c_ptr), value :: h; integer(c_int), value :: ngrids, spincase, what
            type(b200_grid) :: grids(*); real(c_double) :: b(3,*), out7(7,*)
        end function
        ! get_property (jfield.f90:584-929): part(5, nseg, natoms+1) partial sums per point block; seg_end = cumulative nelpts
        integer(c_int) function gimic_b200_property(h, n, r, w, tens, natoms, coords, nseg, seg_end, part, flags) bind(c)
            import; type(c_ptr), value :: h; integer(c_long), value :: n; integer(c_int), value :: natoms, nseg, flags
            real(c_double) :: r(3,*), w(*), tens(9,*), coords(3,*), part(5,nseg,*); integer(c_long) :: seg_end(*)
        end function
        ! per-point integrands for one nucleus (centre3) or the magnetizability (centre3 = c_null_ptr): out4(4, n)
        integer(c_int) function gimic_b200_property_integrand(h, n, r, tens, centre3, out4, flags) bind(c)
            import; type(c_ptr), value :: h, centre3; integer(c_long), value :: n; integer(c_int), value :: flags
            real(c_double) :: r(3,*), tens(9,*), out4(4,*)
        end function
        ! n values with the edit descriptor Ew.d into a character buffer (vtkplot.f90 writers); returns the bytes written
        integer(c_long) function gimic_b200_format_e(n, v, w, d, per_line, first_count, prefix, out, cap) bind(c)
            import; integer(c_long), value :: n, cap; integer(c_int), value :: w, d, per_line, first_count
            real(c_double) :: v(*); character(kind=c_char) :: prefix(*), out(*)
        end function
        integer(c_long) function gimic_b200_format_f(n, v, w, d, per_line, first_count, prefix, out, cap) bind(c)
            import; integer(c_long), value :: n, cap; integer(c_int), value :: w, d, per_line, first_count
            real(c_double) :: v(*); character(kind=c_char) :: prefix(*), out(*)
        end function
        ! geometry of a MOL file without a context (dry run, gimic.F90:142-204): returns natoms; symbols2 = 2 characters per atom
        integer(c_int) function gimic_b200_mol_geometry(mol, max_atoms, xyz, symbols2) bind(c)
            import; integer(c_int), value :: max_atoms
            character(kind=c_char) :: mol(*), symbols2(*); real(c_double) :: xyz(3,*)
        end function
        ! signed |J| of given J vectors (jmod2_vtkplot, jfield.f90:446-489)
        integer(c_int) function gimic_b200_jmod_from_jvec(h, n, r, jvec, b, jmod, flags) bind(c)
            import; type(c_ptr), value :: h; integer(c_long), value :: n; integer(c_int), value :: flags
            real(c_double) :: r(3,*), jvec(3,*), b(3), jmod(*)
        end function
        ! ---- include/gimic_b200_driver.h (libgimic_b200_driver.so): run modes and writers ------------------------------
        ! `gimic gimic.inp` as one call (src/gimic.in:116-159 + program gimic); flags: 1 dry run, 2 appended-binary .vti
        integer(c_int) function gimic_b200_run_input(inpfile, workdir, device, flags, report_path) bind(c)
            import; character(c_char) :: inpfile(*), workdir(*), report_path(*); integer(c_int), value :: device, flags
        end function
        ! general form: type(b200_run_opts) mirrors gimic_b200_run_opts (flags, device, ndevices, devices, workdir, title, report_path)
        integer(c_int) function gimic_b200_run(inpfile, opts) bind(c)
            import; character(c_char) :: inpfile(*); type(b200_run_opts) :: opts
        end function
        ! writers of vtkplot.f90:14-391 / jfield.f90:531-541 on the grid that gimic.inp describes; kind = 'vti_scalar',
        ! 'vti_vector', 'jmod_txt', 'vtu_vector', 'vtu_scalar' (NUL-terminated)
        integer(c_int) function gimic_b200_write_field(inpfile, workdir, kind, data, n, fname, flags) bind(c)
            import; character(c_char) :: inpfile(*), workdir(*), kind(*), fname(*); real(c_double) :: data(*)
            integer(c_long), value :: n; integer(c_int), value :: flags
        end function
        function gimic_b200_last_error() bind(c) result(msg)
            import; type(c_ptr) :: msg
        end function
    end interface

contains

    integer(c_int) function b200_spin_code(spincase) result(code)
        character(*), intent(in) :: spincase
        select case (trim(spincase))
            case ('alpha');    code = B200_ALPHA
            case ('beta');     code = B200_BETA
            case ('spindens'); code = B200_SPINDENS
            case default;      code = B200_TOTAL
        end select
    end function

    ! replaces new_basis + new_dens + read_dens in driver (gimic.F90:144-158)
    subroutine b200_init(molfile, densfile, is_uhf, use_giao, use_diamag, use_paramag, use_screening, screen_thrs, use_spherical)
        character(*), intent(in) :: molfile, densfile
        logical, intent(in) :: is_uhf, use_giao, use_diamag, use_paramag, use_screening, use_spherical
        real(c_double), intent(in) :: screen_thrs
        type(b200_opts) :: o
        call gimic_b200_default_opts(o)
        o%uhf = merge(1, 0, is_uhf); o%giao = merge(1, 0, use_giao); o%diamag = merge(1, 0, use_diamag)
        o%paramag = merge(1, 0, use_paramag); o%screening = merge(1, 0, use_screening); o%screening_thrs = screen_thrs
        o%spherical = merge(1, 0, use_spherical)
        if (gimic_b200_create(b200_handle, trim(molfile)//c_null_char, trim(densfile)//c_null_char, o) /= 0) then
            stop 'gimic_b200_create failed'
        end if
    end subroutine

    ! replaces the "!$omp parallel ... call ctensor(jt, coord, tens(:,n-lo+1), spincase)" loop of calc_jtensors
    subroutine b200_calc_jtensors(coords, spincase, tens)
        real(c_double), intent(in) :: coords(:,:)        ! (3, npts) = gridpoint() of the rank's lo..hi
        character(*), intent(in) :: spincase
        real(c_double), intent(out) :: tens(:,:)         ! (9, npts)
        if (gimic_b200_calc_jtensors(b200_handle, int(size(coords, 2), c_long), coords, b200_spin_code(spincase), tens, 0) /= 0) then
            stop 'gimic_b200_calc_jtensors failed'
        end if
    end subroutine

    subroutine b200_finalize()
        integer(c_int) :: ierr
        if (c_associated(b200_handle)) ierr = gimic_b200_destroy(b200_handle)
        b200_handle = c_null_ptr
    end subroutine
end module
