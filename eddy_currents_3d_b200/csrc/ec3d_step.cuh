// ec3d_step.cuh -- K7/K8: per-timestep kernels around the solve (EC3D.f90:275-433): source motion
// and scatter, inertial sources / U-row right-hand side, history update.  Uaf, Jaf and the coil
// positions stay on the device.
#pragma once
#include "ec3d_common.cuh"
#include "ec3d_rows.cuh"
#include "ec3d_kernels.cuh"

struct MotionState {          // device-resident (tfun_nod Distance/length + movestop, EC3D.f90:238)
    double Distance[EC3D_MAX_FUN][3];
    double shift[EC3D_MAX_FUN][3];
    int length[EC3D_MAX_FUN][3];
    int movestop[3];
    int oob;                  // set when a moved cell falls outside the grid (never for valid input)
};

struct SourceDesc {           // device arrays describing the source functions
    const int *nod_ptr;       // [numfun+1]
    const int *nods;          // global 1-based unknown indices (nods_Fx / nods_Fy)
    const int *num_Vmech;     // [numfun][3]
    const int *comp;          // [numfun] 0 = X, 1 = Y, 2 = Z
    int numfun;
};

// new_m (EC3D.f90:1064-1114): shift cell m (1-based) by length[], clamp to [2, sd-2].
// The two ceiling() expressions are evaluated in default REAL (single precision) as in the
// reference.  clamp[a] is set when axis a was clamped.
__device__ __forceinline__ long long new_m_dev(int m, int sdx, int sdy, int sdz, const int len[3], bool clamp[3])
{
    const int kdz = sdx * sdy;
    const int L = (int)ceilf(__fdiv_rn((float)m, (float)kdz));
    int Lnew = L + len[2];
    clamp[2] = false;
    if (Lnew > sdz - 2) { clamp[2] = true; Lnew = sdz - 2; }
    else if (Lnew < 2) { clamp[2] = true; Lnew = 2; }
    const int nij = (L == 1) ? m : m - (L - 1) * kdz;
    const int j = (int)ceilf(__fdiv_rn((float)nij, (float)sdx));
    int jnew = j + len[1];
    clamp[1] = false;
    if (jnew > sdy - 2) { clamp[1] = true; jnew = sdy - 2; }
    else if (jnew < 2) { clamp[1] = true; jnew = 2; }
    const int i = nij - (j - 1) * sdx;
    int inew = i + len[0];
    clamp[0] = false;
    if (inew > sdx - 2) { clamp[0] = true; inew = sdx - 2; }
    else if (inew < 2) { clamp[0] = true; inew = 2; }
    return (long long)inew + (long long)sdx * (jnew - 1) + (long long)kdz * (Lnew - 1);
}

// motion_calc for every function in order (EC3D.f90:1052-1062), with the movestop flags carried
// from function to function exactly as the sequential node loop leaves them: after all nodes of a
// function, movestop(a) = 0 if its LAST node was clamped along a, else 1 (the re-enable test
// `Lnew < sd-2 .or. Lnew > 2` is always true for sd > 4, which ec3d_create requires for motion).
__global__ void k_motion(const SlabGeom G, const SourceDesc sd, MotionState *ms, const double *__restrict__ vmech,
                         const double dt, const double d0, const double d1, const double d2)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double delta[3] = {d0, d1, d2};
    for (int n = 0; n < sd.numfun; ++n) {
        for (int a = 0; a < 3; ++a) {
            const int nv = sd.num_Vmech[3 * n + a];
            if (nv == 0)
                ms->Distance[n][a] = DADD(ms->Distance[n][a], DMUL((double)ms->movestop[0], ms->shift[n][a]));
            else
                ms->Distance[n][a] = DADD(ms->Distance[n][a], DMUL(vmech[nv - 1], dt) / delta[a]);
            ms->length[n][a] = (int)llrint(copysign(floor(fabs(ms->Distance[n][a]) + 0.5), ms->Distance[n][a]));
        }
        const int nb = sd.nod_ptr[n], ne = sd.nod_ptr[n + 1];
        if (ne > nb) {
            const int m = sd.nods[ne - 1] - sd.comp[n] * (int)G.nC;
            bool cl[3];
            int len[3] = {ms->length[n][0], ms->length[n][1], ms->length[n][2]};
            new_m_dev(m, G.sdx, G.sdy, G.sdz, len, cl);
            for (int a = 0; a < 3; ++a) ms->movestop[a] = cl[a] ? 0 : 1;
        }
    }
}

// Moving sources: Jaf keeps only the A entries of conductor cells (EC3D.f90:277-295).
__global__ void k_clear_nonconductor(const SlabGeom G, const int *__restrict__ geo, double *__restrict__ Jaf,
                                     const int keep_conductor)
{
    const long long cells = (long long)G.nzl * G.kdz;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < cells + G.nUown;
         q += (long long)gridDim.x * blockDim.x) {
        if (q < cells) {
            const int g = geo[q + 2 * (long long)G.kdz];
            if (!(keep_conductor && g != 0)) {
                const long long l = q + G.kdz;
                Jaf[l] = 0.0; Jaf[G.segA + l] = 0.0; Jaf[2 * G.segA + l] = 0.0;
            }
        } else {
            Jaf[G.offU + G.nUlo + (q - cells)] = 0.0;
        }
    }
}

// Source scatter of ONE function (EC3D.f90:301-367): Jaf(target) = a.  Functions are launched in
// order on one stream, which preserves the reference's last-writer-wins order between functions;
// within a function every node writes the same value.
__global__ void k_scatter(const SlabGeom G, const SourceDesc sd, const MotionState *__restrict__ ms, const int fn,
                          const int moving, const double *__restrict__ fun_vely, double *__restrict__ Jaf,
                          int *__restrict__ new_nodes, int *__restrict__ oob)
{
    const int nb = sd.nod_ptr[fn], ne = sd.nod_ptr[fn + 1];
    const int q = nb + blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= ne) return;
    const int comp = sd.comp[fn];
    long long m = sd.nods[q] - (long long)comp * G.nC;       // cell number 1..nC
    if (moving) {
        bool cl[3];
        int len[3] = {ms->length[fn][0], ms->length[fn][1], ms->length[fn][2]};
        m = new_m_dev((int)m, G.sdx, G.sdy, G.sdz, len, cl);
    }
    new_nodes[q] = (int)m;
    if (m < 1 || m > G.nC) { *oob = 1; return; }
    const long long cell0 = m - 1;
    const int k = (int)(cell0 / G.kdz);
    if (k < G.k0 || k >= G.k1) return;                        // another rank owns the target plane
    Jaf[comp * G.segA + (cell0 - (long long)(G.k0 - 1) * G.kdz)] = fun_vely[fn];
}

// Inertial sources and the U-row right-hand side (EC3D.f90:370-404), one thread per conductor cell.
__global__ void __launch_bounds__(256)
k_rhs_pre(const SlabGeom G, const Coef cf, const int *__restrict__ geo, const int *__restrict__ cond_cells,
          const int ncond, const unsigned char *__restrict__ flags, const double valdom,
          const double *__restrict__ Uaf, double *__restrict__ Jaf)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncond) return;
    const int cell0 = cond_cells[t];
    const int k = cell0 / G.kdz, rem = cell0 - k * G.kdz, j = rem / G.sdx, i = rem - j * G.sdx;
    const GeoView gv = dense_view(G, geo);
    const long long cell_shift = (long long)(G.k0 - 1) * G.kdz;
    const long long lp = (long long)cell0 - cell_shift;
    const int f = flags[t];
#pragma unroll
    for (int comp = 0; comp < 3; ++comp) {
        const long long idx = comp * G.segA + lp;
        // Jaf = valdom*Uaf + Jaf (:381-383), then zero on cel_bndX/Y/Z (:400-402)
        Jaf[idx] = ((f >> comp) & 1) ? 0.0 : DADD(DMUL(valdom, Uaf[idx]), Jaf[idx]);
    }
    // U row: sum over the A columns of the row (:385-392), zero on cel_bndUx/Uy/Uz (:396-398)
    GatherVisitor v{Uaf, G.segA, G.offU, cell_shift, 0.0};
    cond_u_row(cf, gv, i, j, k, 1, v);
    const long long lu = u_local(G, i, j, k);
    Jaf[lu] = ((f >> 3) & 7) ? 0.0 : v.s;
}

// History update after the solve (EC3D.f90:412-433).
__global__ void __launch_bounds__(256)
k_rhs_post(const SlabGeom G, const int *__restrict__ cond_cells, const int ncond,
           const unsigned char *__restrict__ flags, const double valdom, double *__restrict__ Uaf,
           double *__restrict__ Jaf)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncond) return;
    const int cell0 = cond_cells[t];
    const long long lp = (long long)cell0 - (long long)(G.k0 - 1) * G.kdz;
    const int f = flags[t];
#pragma unroll
    for (int comp = 0; comp < 3; ++comp) {
        const long long idx = comp * G.segA + lp;
        if ((f >> comp) & 1) { Jaf[idx] = 0.0; Uaf[idx] = 0.0; }
        else Jaf[idx] = DSUB(DMUL(valdom, Uaf[idx]), Jaf[idx]);
    }
}

// Output post-processing on the device (writeVtk_field, utilites.f90:222-290): the
// per-point float32 triples of a field_N.vtk in file order, for the owned planes.  what: 0 Field_A
// (= Uaf), 1 Vector_field_eddy (s*Jaf on conductor cells, 0 elsewhere), 2 Vector_field_SOURCE (Jaf on
// non-conductor cells; all cells without a conductor), 3 Vector_field_B (curl A, central differences,
// indices clamped at the domain faces; needs fresh Uaf halo planes).  fp64 in the reference's order,
// one rounding to float; big_endian != 0 byte-swaps like convert="big_endian" (:192).
__device__ __forceinline__ float out_f(const double v, const int big_endian)
{
    const float f = __double2float_rn(v);
    if (!big_endian) return f;
    return __uint_as_float(__byte_perm(__float_as_uint(f), 0u, 0x0123));
}

__global__ void __launch_bounds__(256)
k_vtk_field(const SlabGeom G, const int *__restrict__ geo, const double *__restrict__ Uaf, const double *__restrict__ Jaf,
            const int what, const int has_conductor, const double d0, const double d1, const double d2,
            const int big_endian, float *__restrict__ out)
{
    const long long cells = (long long)G.nzl * G.kdz;
    const double sfac = -0.07957747154594766788444e7;             // utilites.f90:239
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < cells; q += (long long)gridDim.x * blockDim.x) {
        const int kl = (int)(q / G.kdz), rem = (int)(q - (long long)kl * G.kdz), j = rem / G.sdx, i = rem - j * G.sdx;
        const int k = G.k0 + kl;
        const long long l = q + G.kdz;                            // local A offset (one halo plane below)
        const int g = geo[q + 2 * (long long)G.kdz];              // geo planes start at k0-2
        double v0 = 0.0, v1 = 0.0, v2 = 0.0;
        if (what == 0) {
            v0 = Uaf[l]; v1 = Uaf[G.segA + l]; v2 = Uaf[2 * G.segA + l];
        } else if (what == 1) {
            if (has_conductor && g != 0) { v0 = DMUL(sfac, Jaf[l]); v1 = DMUL(sfac, Jaf[G.segA + l]); v2 = DMUL(sfac, Jaf[2 * G.segA + l]); }
        } else if (what == 2) {
            if (!has_conductor || g == 0) { v0 = Jaf[l]; v1 = Jaf[G.segA + l]; v2 = Jaf[2 * G.segA + l]; }
        } else {
            const long long im = (i == 0) ? l : l - 1, ip = (i == G.sdx - 1) ? l : l + 1;
            const long long jm = (j == 0) ? l : l - G.sdx, jp = (j == G.sdy - 1) ? l : l + G.sdx;
            const long long km = (k == 0) ? l : l - G.kdz, kp = (k == G.sdz - 1) ? l : l + G.kdz;
            const double *A0 = Uaf, *A1 = Uaf + G.segA, *A2 = Uaf + 2 * G.segA;
            v0 = DSUB(DMUL(0.5, DSUB(A2[jp], A2[jm])) / d1, DMUL(0.5, DSUB(A1[kp], A1[km])) / d2);
            v1 = DSUB(DMUL(0.5, DSUB(A0[kp], A0[km])) / d2, DMUL(0.5, DSUB(A2[ip], A2[im])) / d0);
            v2 = DSUB(DMUL(0.5, DSUB(A1[ip], A1[im])) / d0, DMUL(0.5, DSUB(A0[jp], A0[jm])) / d1);
        }
        out[3 * q] = out_f(v0, big_endian);
        out[3 * q + 1] = out_f(v1, big_endian);
        out[3 * q + 2] = out_f(v2, big_endian);
    }
}
