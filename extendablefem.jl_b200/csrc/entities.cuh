// entities.cuh -- assembly on boundary faces (entities = ON_BFACES) and the ItemIntegrator.
//
//   ON_BFACES   src/common_operators/bilinear_operator.jl:707-714 (the entity branch of build_assembler!: item geometries,
//               volumes, regions and the dofmap come from the boundary faces, the loop :820-951 is unchanged),
//               src/common_operators/linear_operator.jl:531-544; users: Example108:61 (Robin), Example330:105 (traction).
//   ItemIntegrator  src/common_operators/item_integrator.jl:191-249 (assembly_loop), :323-352 (evaluate).
//
// Boundary faces are O(boundary) work, so these kernels favour simplicity: one thread per local entry, Identity operators
// only (the restriction of an H1 function to a face; gradients on faces are outside the hot path).  Their contributions
// are added to the CSC values by an owner-computes scatter (one thread owns one matrix column / vector row, walks the
// faces adjacent to its dof in ascending order and finds the row by binary search): deterministic, no atomics.
#pragma once
#include "common.cuh"
#include "kernels_generic.cuh"
#include "solver.cuh"

namespace extfem {

// x at quadrature point q of item `f` of a (tdim)-simplex mesh embedded in SDIM space
template <int SDIM>
__device__ __forceinline__ void item_x(const OpDev &op, int tdim, long long f, int q, double *x)
{
    const int *cn = op.cellnodes + f * (tdim + 1);
    double l0 = 1.0;
    for (int r = 0; r < tdim; ++r) l0 -= op.qx[q * tdim + r];
#pragma unroll
    for (int d = 0; d < SDIM; ++d) x[d] = l0 * op.coords[(size_t)cn[0] * SDIM + d];
    for (int r = 0; r < tdim; ++r) {
        const double l = op.qx[q * tdim + r];
#pragma unroll
        for (int d = 0; d < SDIM; ++d) x[d] += l * op.coords[(size_t)cn[r + 1] * SDIM + d];
    }
}

// |F| of boundary faces: 1 for the vertices of a 1D grid, edge length in 2D, triangle area in 3D (xgrid[BFaceVolumes])
__global__ void face_volumes_kernel(int sdim, long long nfaces, const double *__restrict__ coords, const int *__restrict__ facenodes,
                                    double *__restrict__ vol)
{
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nfaces) return;
    if (sdim == 1) { vol[f] = 1.0; return; }
    const int *fn = facenodes + f * sdim;
    double e[2][3] = {{0, 0, 0}, {0, 0, 0}};
    for (int r = 0; r < sdim - 1; ++r)
        for (int d = 0; d < sdim; ++d) e[r][d] = coords[(size_t)fn[r + 1] * sdim + d] - coords[(size_t)fn[0] * sdim + d];
    if (sdim == 2) { vol[f] = sqrt(e[0][0] * e[0][0] + e[0][1] * e[0][1]); return; }
    const double cx = e[0][1] * e[1][2] - e[0][2] * e[1][1], cy = e[0][2] * e[1][0] - e[0][0] * e[1][2], cz = e[0][0] * e[1][1] - e[0][1] * e[1][0];
    vol[f] = 0.5 * sqrt(cx * cx + cy * cy + cz * cz);
}

__device__ __forceinline__ bool item_visited(const OpDev &op, long long f)
{
    if (op.nregions == 0) return true;
    const int reg = op.cellregions[f];
    bool vis = false;
    for (int k = 0; k < op.nregions; ++k) vis |= (op.regions[k] == reg);
    return vis;
}

// input_args at point q of face f (Identity operators only)
__device__ __forceinline__ void face_eval_args(const OpDev &op, long long f, int q, double *u)
{
    for (int d = 0; d < op.nin; ++d) u[d] = 0.0;
    for (int id = 0; id < op.nargs; ++id) {
        const ArgDev &a = op.args[id];
        const int *dofs = a.celldofs + f * a.nd;
        for (int j = 0; j < a.nd; ++j) {
            const int c = j / a.nscalar, k = j - c * a.nscalar;
            u[a.opoff + c] += op.sol[a.soloff + dofs[j]] * __ldg(a.refvals + q * a.nscalar + k);
        }
    }
}

// local matrices of a BilinearOperator on boundary faces: loc[f][NC][NR]   (bilinear_operator.jl:876-920 with face items)
template <int SDIM>
__global__ void __launch_bounds__(256) face_bilinear_kernel(const __grid_constant__ OpDev op, int tdim, double *__restrict__ loc)
{
    const int NRC = op.NR * op.NC;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= op.ncells * NRC) return;
    const long long f = e / NRC;
    const int r = (int)(e - f * NRC), j = r / op.NR, k = r - j * op.NR;
    double acc = 0.0;
    if (item_visited(op, f)) {
        for (int id = 0; id < op.nansatz; ++id) {
            const ArgDev &aa = op.ansatz[id];
            if (j < aa.locoff || j >= aa.locoff + aa.nd) continue;
            const int jj = j - aa.locoff, ca = jj / aa.nscalar, ka = jj - ca * aa.nscalar;
            for (int idt = 0; idt < op.ntest; ++idt) {
                const ArgDev &ta = op.test[idt];
                if (k < ta.locoff || k >= ta.locoff + ta.nd) continue;
                if (!op.coupling[id * op.ntest + idt]) continue;
                const int kk = k - ta.locoff, ct = kk / ta.nscalar, kt = kk - ct * ta.nscalar;
                double a = 0.0;
                for (int q = 0; q < op.nq; ++q) {
                    double in[MAXOP], res[MAXOP], ain[MAXOP];
                    for (int d = 0; d < MAXOP; ++d) in[d] = 0.0;
                    in[aa.opoff + ca] = __ldg(aa.refvals + q * aa.nscalar + ka);
                    if (op.nargs > 0) face_eval_args(op, f, q, ain);
                    bl_apply(op.kernel_id, SDIM, in, ain, op.params, res, op.nout);
                    a += (res[ta.opoff + ct] * (op.factor * op.qw[q])) * __ldg(ta.refvals + q * ta.nscalar + kt);
                }
                acc += a * op.cellvolumes[f];
            }
        }
    }
    loc[e] = acc;
}

// local vectors of a LinearOperator on boundary faces: bloc[f][NR]   (linear_operator.jl:619-637 with face items)
template <int SDIM>
__global__ void __launch_bounds__(256) face_linear_kernel(const __grid_constant__ OpDev op, int tdim, double *__restrict__ bloc)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= op.ncells * op.NR) return;
    const long long f = e / op.NR;
    const int k = (int)(e - f * op.NR);
    double acc = 0.0;
    if (item_visited(op, f)) {
        for (int idt = 0; idt < op.ntest; ++idt) {
            const ArgDev &ta = op.test[idt];
            if (k < ta.locoff || k >= ta.locoff + ta.nd) continue;
            const int kk = k - ta.locoff, ct = kk / ta.nscalar, kt = kk - ct * ta.nscalar;
            for (int q = 0; q < op.nq; ++q) {
                double r[MAXOP], x[3] = {0.0, 0.0, 0.0};
                if (op.nargs > 0) face_eval_args(op, f, q, r);   // standard kernel: result = input_args
                else {
                    item_x<SDIM>(op, tdim, f, q, x);
                    const double *tab = op.tabulated ? op.tabulated + ((size_t)f * op.nq + q) * op.nout : nullptr;
                    lin_apply(op.kernel_id, x, op.params, r, op.nout, tab);
                }
                acc += (r[ta.opoff + ct] * (op.factor * op.qw[q] * op.cellvolumes[f])) * __ldg(ta.refvals + q * ta.nscalar + kt);
            }
        }
    }
    bloc[e] = acc;
}

// ---- owner-computes scatter of face-local contributions into the cell pattern ---------------------------------------
struct FaceScatterArgs {
    long long ncols;                 // dofs of the column space
    long long colbase;               // global index of the block's first column
    const long long *adjptr;         // dof -> faces adjacency of the column space's BFaceDofs
    const int *adjface;
    const unsigned char *adjloc;
    int collocoff;                   // operator-local offset of the column block
    int nrb;                         // row blocks the operator touches
    const int *rowfacedofs[MAXARGS]; // [nbfaces][nd], 0-based block-local
    int rownd[MAXARGS];
    long long rowoff[MAXARGS];       // global offset of the row block
    int rowlocoff[MAXARGS];          // operator-local offset of the row block
    int NRop, NCop;
    const double *loc;
    const long long *colptr;
    const int *rowval;
    double *nzval;
    int *error;
};

__global__ void __launch_bounds__(256) face_scatter_matrix_kernel(const __grid_constant__ FaceScatterArgs A)
{
    const long long kc = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (kc >= A.ncols) return;
    const long long p0 = A.adjptr[kc], p1 = A.adjptr[kc + 1];
    for (long long p = p0; p < p1; ++p) {
        const long long f = A.adjface[p];
        const int kl = A.collocoff + A.adjloc[p];
        const double *src = A.loc + (f * A.NCop + kl) * A.NRop;
        for (int r = 0; r < A.nrb; ++r)
            for (int t = 0; t < A.rownd[r]; ++t) {
                const int row = (int)(A.rowoff[r] + A.rowfacedofs[r][f * A.rownd[r] + t]);
                const long long pos = find_in_column(A.colptr, A.rowval, A.colbase + kc, row);
                if (pos < 0) { atomicExch(A.error, 1); continue; }
                A.nzval[pos] += src[A.rowlocoff[r] + t];
            }
    }
}

struct FaceScatterVecArgs {
    long long nrows;                 // dofs of the row space
    long long rowbase;
    const long long *adjptr;
    const int *adjface;
    const unsigned char *adjloc;
    int rowlocoff, NRop;
    const double *bloc;
    double *b;
};

__global__ void __launch_bounds__(256) face_scatter_vector_kernel(const __grid_constant__ FaceScatterVecArgs A)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.nrows) return;
    const long long p0 = A.adjptr[i], p1 = A.adjptr[i + 1];
    if (p1 == p0) return;
    double s = A.b[A.rowbase + i];
    for (long long p = p0; p < p1; ++p) s += A.bloc[(long long)A.adjface[p] * A.NRop + A.rowlocoff + A.adjloc[p]];
    A.b[A.rowbase + i] = s;
}

// ---- ItemIntegrator ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ii_apply(int id, const double *in, const double *x, const double *p, const double *tab, double *r,
                                         int nin, int resultdim)
{
    switch (id) {
    case EXTFEM_II_STANDARD: for (int d = 0; d < resultdim; ++d) r[d] = d < nin ? in[d] : 0.0; break;
    case EXTFEM_II_L2NORM: for (int d = 0; d < resultdim; ++d) r[d] = d < nin ? in[d] * in[d] : 0.0; break;
    case EXTFEM_II_L2DIFF_TABULATED: for (int d = 0; d < resultdim; ++d) { const double e = tab[d] - in[d]; r[d] = e * e; } break;
    case EXTFEM_II_L2ERR_SINCOS301: { const double e = sin(1.7 * x[0]) * cos(3.9 * x[1]) - in[0]; r[0] = e * e; } break;
    case EXTFEM_II_L2ERR_EXP108: { const double e = exp(x[0]) - in[0]; r[0] = e * e; } break;
    }
    (void)p;
}

// one thread per cell: out[cell][resultdim] = sum_q kernel(input_args_q, x_q) * factor * w_q * |T|   (item_integrator.jl:213-246)
template <int DIM>
__global__ void __launch_bounds__(128) item_integrate_kernel(const __grid_constant__ OpDev op, int resultdim, double *__restrict__ out)
{
    const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= op.ncells) return;
    double acc[MAXOP];
    for (int d = 0; d < resultdim; ++d) acc[d] = 0.0;
    CellGeo<DIM> G;
    load_geo<DIM>(op, cell, G);
    if (G.visited) {
        for (int q = 0; q < op.nq; ++q) {
            double u[MAXOP], r[MAXOP], x[DIM];
            eval_args<DIM>(op, cell, q, G, u);
            eval_x<DIM>(G, op.qx + q * DIM, x);
            const double *tab = op.tabulated ? op.tabulated + ((size_t)cell * op.nq + q) * resultdim : nullptr;
            ii_apply(op.kernel_id, u, x, op.params, tab, r, op.nin, resultdim);
            const double sc = op.factor * op.qw[q] * G.vol;
            for (int d = 0; d < resultdim; ++d) acc[d] += r[d] * sc;
        }
    }
    for (int d = 0; d < resultdim; ++d) out[cell * resultdim + d] = acc[d];
}

// deterministic sum over the items of component d: partial[b][d], then total[d]
__global__ void __launch_bounds__(256) ii_reduce_partial_kernel(long long n, int resultdim, const double *__restrict__ v, double *__restrict__ partial)
{
    __shared__ double sh[256];
    for (int d = 0; d < resultdim; ++d) {
        double s = 0.0;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s += v[i * resultdim + d];
        sh[threadIdx.x] = s;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
        if (threadIdx.x == 0) partial[(size_t)blockIdx.x * resultdim + d] = sh[0];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) ii_reduce_final_kernel(int nb, int resultdim, const double *__restrict__ partial, double *__restrict__ total)
{
    __shared__ double sh[256];
    for (int d = 0; d < resultdim; ++d) {
        double s = 0.0;
        for (int i = threadIdx.x; i < nb; i += 256) s += partial[(size_t)i * resultdim + d];
        sh[threadIdx.x] = s;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
        if (threadIdx.x == 0) total[d] = sh[0];
        __syncthreads();
    }
}

__global__ void apply_values_kernel(long long n, const long long *__restrict__ dofs, const double *__restrict__ values, double *__restrict__ sol,
                                    long long nsol, int *err)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long d = dofs[i] - 1;
    if (d < 0 || d >= nsol) { atomicExch(err, 1); return; }
    sol[d] = values ? values[i] : 0.0;
}

} // namespace extfem
