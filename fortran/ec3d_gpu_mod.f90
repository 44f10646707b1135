! ec3d_gpu_mod.f90 -- ISO_C_BINDING interfaces of libec3d_gpu.so (include/ec3d_gpu.h).
! Level 1: the solver only (ec3d_bicgstabwr_csr).  Level 2: GPU-resident time stepping that replaces
! EC3D.f90:275-433 (ec3d_create / ec3d_step / ec3d_get_fields / ec3d_get_vtk_fields / ec3d_destroy).
! Shipped as source, not compiled in this repository's build image (no Fortran compiler there).
MODULE ec3d_gpu_mod
  USE, INTRINSIC :: ISO_C_BINDING
  IMPLICIT NONE

  ! mirrors `ec3d_config` field by field; arrays are passed as C_LOC of contiguous TARGET arrays
  TYPE, BIND(C) :: ec3d_config
     INTEGER(C_INT32_T) :: sdx, sdy, sdz
     REAL(C_DOUBLE)     :: delta(3)
     REAL(C_DOUBLE)     :: dt
     REAL(C_DOUBLE)     :: BND(2,3)          ! C BND[axis][side]: store BND(side,axis) = Fortran BND(axis,side)
     REAL(C_DOUBLE)     :: tolerance
     INTEGER(C_INT32_T) :: itmax
     INTEGER(C_INT32_T) :: nmat
     TYPE(C_PTR)        :: valPHYS           ! REAL(8) (5,nmat): column m = valPHYS(m,1:5)
     TYPE(C_PTR)        :: geoPHYS           ! INTEGER(1) (sdx,sdy,sdz)
     TYPE(C_PTR)        :: geoPHYS_C         ! INTEGER    (sdx,sdy,sdz)
     INTEGER(C_INT32_T) :: size_PHYS_C
     TYPE(C_PTR)        :: cond_nod_ptr      ! INTEGER (size_PHYS_C+1): 0, nCells0
     TYPE(C_PTR)        :: cond_nod          ! PHYS_C(1)%nod
     TYPE(C_PTR)        :: cond_valdom       ! PHYS_C(:)%valdom
     INTEGER(C_INT32_T) :: numfun
     TYPE(C_PTR)        :: fun_ex            ! CHARACTER(1) (numfun): 'X' / 'Y'
     TYPE(C_PTR)        :: fun_nod_ptr       ! INTEGER (numfun+1), 0-based offsets into fun_nods
     TYPE(C_PTR)        :: fun_nods          ! nods_Fx / nods_Fy of every function, concatenated
     TYPE(C_PTR)        :: fun_num_Vmech     ! INTEGER (3,numfun)
     TYPE(C_PTR)        :: fun_move          ! INTEGER (3,numfun)
     TYPE(C_PTR)        :: fun_vel_Vmech     ! REAL(8) (3,numfun)
     INTEGER(C_INT32_T) :: numMech
     INTEGER(C_INT32_T) :: nranks, rank
     TYPE(C_PTR)        :: nccl_id           ! 128 bytes from ec3d_nccl_unique_id of rank 0 (nranks > 1)
     INTEGER(C_INT32_T) :: device            ! CUDA device ordinal, -1 = current
  END TYPE ec3d_config

  INTERFACE
     INTEGER(C_INT) FUNCTION ec3d_bicgstabwr_csr(valA, irow, jcol, n, b, x, tolerance, itmax, iter) &
          BIND(C, NAME='ec3d_bicgstabwr_csr')
       IMPORT :: C_INT, C_DOUBLE
       REAL(C_DOUBLE), INTENT(IN)    :: valA(*), b(*)
       INTEGER(C_INT), INTENT(IN)    :: irow(*), jcol(*)
       INTEGER(C_INT), VALUE         :: n, itmax
       REAL(C_DOUBLE), VALUE         :: tolerance
       REAL(C_DOUBLE), INTENT(INOUT) :: x(*)
       INTEGER(C_INT), INTENT(OUT)   :: iter
     END FUNCTION ec3d_bicgstabwr_csr

     INTEGER(C_INT) FUNCTION ec3d_create(cfg, handle) BIND(C, NAME='ec3d_create')
       IMPORT :: C_INT, C_PTR, ec3d_config
       TYPE(ec3d_config), INTENT(IN) :: cfg
       TYPE(C_PTR), INTENT(OUT)      :: handle
     END FUNCTION ec3d_create

     INTEGER(C_INT) FUNCTION ec3d_step(handle, fun_vely, vmech_vely, iter) BIND(C, NAME='ec3d_step')
       IMPORT :: C_INT, C_PTR, C_DOUBLE
       TYPE(C_PTR), VALUE          :: handle
       REAL(C_DOUBLE), INTENT(IN)  :: fun_vely(*), vmech_vely(*)
       INTEGER(C_INT), INTENT(OUT) :: iter
     END FUNCTION ec3d_step

     INTEGER(C_INT) FUNCTION ec3d_get_fields(handle, Uaf, Jaf) BIND(C, NAME='ec3d_get_fields')
       IMPORT :: C_INT, C_PTR, C_DOUBLE
       TYPE(C_PTR), VALUE          :: handle
       REAL(C_DOUBLE), INTENT(OUT) :: Uaf(*), Jaf(*)
     END FUNCTION ec3d_get_fields

     INTEGER(C_INT) FUNCTION ec3d_get_vtk_fields(handle, fieldA, eddy, source, fieldB, big_endian) &
          BIND(C, NAME='ec3d_get_vtk_fields')
       IMPORT :: C_INT, C_PTR, C_FLOAT, C_INT32_T
       TYPE(C_PTR), VALUE         :: handle
       REAL(C_FLOAT), INTENT(OUT) :: fieldA(*), eddy(*), source(*), fieldB(*)
       INTEGER(C_INT32_T), VALUE  :: big_endian
     END FUNCTION ec3d_get_vtk_fields

     INTEGER(C_INT) FUNCTION ec3d_destroy(handle) BIND(C, NAME='ec3d_destroy')
       IMPORT :: C_INT, C_PTR
       TYPE(C_PTR), VALUE :: handle
     END FUNCTION ec3d_destroy
  END INTERFACE
END MODULE ec3d_gpu_mod
