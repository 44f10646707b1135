! solvers_gpu.f90 -- replaces the reference's solvers.f90 (SUBROUTINE sprsBCGstabWR, solvers.f90:3-63):
! the same external procedure, forwarding to libec3d_gpu.so through ISO_C_BINDING (include/ec3d_gpu.h).
! Shipped as source: no Fortran compiler exists in the build image of this repository, so this file is
! not compiled or link-tested there (the C ABI it binds is tested through ctypes in tests/).
!
! Build (gfortran, Linux), in the reference's src/ directory:
!   gfortran -c -O2 -finit-local-zero -ffree-line-length-none ec3d_gpu_mod.f90 solvers_gpu.f90
!   and link EC3D with  -L<repo>/eddy_currents_3d_b200 -lec3d_gpu -lcudart -lnccl  instead of solvers.o
! (without this file the link-time drop-in works as well: the library exports sprsbcgstabwr_ itself).
SUBROUTINE sprsBCGstabWR (valA, irow, jcol, n, b, x, tolerance, itmax, iter)
  USE ec3d_gpu_mod
  IMPLICIT NONE
  INTEGER jcol(*), irow(n+1)
  REAL(8) valA(*)
  INTEGER iter, itmax, n
  REAL(8) tolerance, b(n), x(n)
  INTEGER(C_INT) :: rc
  rc = ec3d_bicgstabwr_csr(valA, irow, jcol, INT(n, C_INT), b, x, REAL(tolerance, C_DOUBLE), INT(itmax, C_INT), iter)
  IF (rc /= 0) THEN
     PRINT *, 'ec3d_gpu: error code ', rc
     STOP 'sprsBCGstabWR (GPU) failed'
  END IF
END SUBROUTINE sprsBCGstabWR
