/*
 * ec3d_oracle.h -- CPU oracle for the EC3D hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the reference's algorithm (JNSresearcher/eddy_currents_3d,
 * Fortran): sparse assembly, per-timestep RHS/history/motion and the BiCGSTAB-with-restart solve.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it, and only as the checker or the timed CPU baseline -- never as part of the product path.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or sample outputs, and no Fortran
 * compiler exists in this image, so this restatement cannot be diffed against a gfortran run.  It is
 * pinned only by (1) the structural counts the reference prints (EC3D.f90:113,968-971,993), derived
 * for the three shipped decks, (2) README cell counts, (3) analytic stencil identities.
 *
 * Conventions: all index VALUES are 1-based exactly as in the Fortran arrays (irow, jcol, node
 * lists, geoPHYS_C); the C arrays holding them are 0-based.  Build with
 *   gcc -O2 -ffp-contract=off   (mirrors gfortran -O2 on generic x86-64: no FMA, no reassociation).
 */
#ifndef EC3D_ORACLE_H
#define EC3D_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Grid + material description consumed by gen_sparse_matrix (EC3D.f90:465-1049). */
typedef struct {
    int32_t sdx, sdy, sdz;
    double delta[3];
    double dt;
    double BND[3][2];          /* BND[axis][side] == Fortran BND(axis+1, side+1) */
    int32_t nmat;              /* rows of valPHYS */
    const double *valPHYS;     /* [nmat][5], valPHYS(m,c) at [(m-1)*5 + (c-1)] */
    const int8_t *geoPHYS;     /* [sdx*sdy*sdz], x fastest, 1-based material ids */
    const int32_t *geoPHYS_C;  /* [sdx*sdy*sdz], 0 or 3*nC + m */
    int32_t size_PHYS_C;       /* number of conductor domains (0 => no cel_bnd lists) */
    int32_t nCells0;           /* number of U unknowns */
} orc_grid;

/* Outputs of gen_sparse_matrix.  Caller allocates upper bounds:
 *   irow: n+1, jcol/valA: orc_nnz_upper_bound(), each cel_bnd list: nCells0. */
typedef struct {
    int32_t *irow;
    int32_t *jcol;
    double *valA;
    int32_t *cel_bndX, *cel_bndY, *cel_bndZ;
    int32_t *cel_bndUx, *cel_bndUy, *cel_bndUz;
    /* filled on return */
    int64_t num_nzX, num_nzY, num_nzZ, num_nzU, num_nz;
    int32_t num_bndX, num_bndY, num_bndZ, num_bndUx, num_bndUy, num_bndUz;
    int32_t err_cell, err_col; /* set when the reference would STOP */
} orc_csr;

int64_t orc_nnz_upper_bound(const orc_grid *g);

/* EC3D.f90:465-1049.  Returns 0, or >0 where the reference STOPs:
 *  1 = column <= 0 (EC3D.f90:717-720 etc), 2 = duplicate U column (EC3D.f90:924-936),
 *  3 = nnz does not fit default INTEGER. */
int orc_gen_sparse_matrix(const orc_grid *g, orc_csr *out);

/* solvers.f90:3-50 (+ sprsAx :54-61).  Prints norm2(R) on iter > itmax like the reference.
 * Returns 0. */
int orc_sprsBCGstabWR(const double *valA, const int32_t *irow, const int32_t *jcol, int32_t n,
                      const double *b, double *x, double tolerance, int32_t itmax, int32_t *iter);

/* Bridge between the reference and the CUDA path: solvers.f90:3-50 with EXACTLY ROUNDED inner
 * products (double-double sums of the product terms the CUDA kernels form, see ec3d_oracle.c) and
 * every vector operation / SpMV as in the reference.  The CUDA path must equal THIS bit for bit; its
 * distance to orc_sprsBCGstabWR is what the reference's own summation error does to its iterates.
 * Needs the grid (even sdx) to know the cell pairs.  Returns 0, -2 for odd sdx. */
int orc_sprsBCGstabWR_exact_dots(const double *valA, const int32_t *irow, const int32_t *jcol, int32_t n,
                                 const double *b, double *x, double tolerance, int32_t itmax, int32_t *iter,
                                 int32_t sdx, int32_t sdy, int32_t sdz, const int32_t *geoPHYS_C);

/* solvers.f90:54-61 */
void orc_sprsAx(const double *valA, const int32_t *irow, const int32_t *jcol, int32_t n,
                const double *v, double *y);

/* Sensitivity experiments only: 0 = reference order (default), 1 = pairwise reductions. */
void orc_set_dot_mode(int mode);

/* gfortran NORM2 (scaled sum of squares) and sequential DOT_PRODUCT */
double orc_norm2(const double *x, int64_t n);
double orc_dot(const double *a, const double *b, int64_t n);

/* Source / motion description (m_vxc2data.f90:20-31 tfun_nod, :9-17 tFun). */
typedef struct {
    int32_t numfun;
    const char *ex;            /* [numfun] 'X','Y','Z' (anything else => STOP, EC3D.f90:337,363) */
    const int32_t *nod_ptr;    /* [numfun+1] offsets into nods */
    const int32_t *nods;       /* global 1-based unknown indices as in nods_Fx/Fy/Fz */
    const int32_t *num_Vmech;  /* [numfun][3] */
    const int32_t *move;       /* [numfun][3] */
    const double *vel_Vmech;   /* [numfun][3] */
    double *Distance;          /* [numfun][3] state */
    double *shift;             /* [numfun][3] state (filled by orc_motion_prepare) */
    int32_t *length;           /* [numfun][3] state */
    int32_t movestop[3];       /* state, initialised to 1 (EC3D.f90:238) */
    int32_t flag_move;         /* set by orc_motion_prepare */
} orc_sources;

/* Conductor description (m_vxc2data.f90:34-39 tPHYS). */
typedef struct {
    int32_t size_PHYS_C;
    const int32_t *nod_ptr;    /* [size_PHYS_C+1] */
    const int32_t *nod;        /* cell numbers 1..nC */
    const double *valdom;      /* [size_PHYS_C] = 2*C/dt */
} orc_conductors;

/* EC3D.f90:157-186 */
void orc_motion_prepare(orc_sources *s, const double delta[3], double dt);

/* EC3D.f90:275-367 (source scatter incl. motion_calc/new_m :1052-1114).
 * fun_vely[numfun] already multiplied by the mu0 literal (EC3D.f90:254); vmech_vely[numMech].
 * new_nodes (optional, may be NULL): [total nodes] moved cell numbers in (function,node) order. */
int orc_scatter_sources(const orc_grid *g, orc_sources *s, const orc_conductors *c,
                        const double *fun_vely, const double *vmech_vely, double *Jaf,
                        double *Jafbuf, int32_t *new_nodes);

/* EC3D.f90:370-404 */
void orc_rhs_pre(const orc_grid *g, const orc_conductors *c, const orc_csr *A, const double *Uaf,
                 double *Jaf);

/* EC3D.f90:412-433 */
void orc_rhs_post(const orc_grid *g, const orc_conductors *c, const orc_csr *A, double *Uaf,
                  double *Jaf);

/* utilites.f90:222-290 (writeVtk_field): the four per-point float32 vector fields of a field_N.vtk in
 * file order (x fastest, 3 components per point): Field_A = Uaf; Vector_field_eddy = s*Jaf on
 * conductor cells, 0 elsewhere (s = -0.0795774715459...d7, only written when size_PHYS_C != 0);
 * Vector_field_SOURCE = Jaf on non-conductor cells (all cells when size_PHYS_C == 0);
 * Vector_field_B = curl A by central differences with the indices clamped at the domain faces.
 * Values are computed in fp64 in the reference's order and rounded once to float.  Pointers may be NULL. */
void orc_vtk_fields(int32_t sdx, int32_t sdy, int32_t sdz, const double delta[3], const double *Uaf,
                    const double *Jaf, const int32_t *geoPHYS_C, int32_t size_PHYS_C, float *fieldA,
                    float *eddy, float *source, float *fieldB);

#ifdef __cplusplus
}
#endif
#endif
