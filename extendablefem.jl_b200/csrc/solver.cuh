// solver.cuh -- device-resident use of the assembled system: SpMV, residual b - A*x
// (compute_nonlinear_residual!, src/solvers.jl:38-43), boundary penalties
// (apply_penalties!, homogeneousdata_operator.jl:186-201) and a Jacobi-preconditioned CG that
// exercises the matrix without gathering it to the host (north_star).
#pragma once
#include <cub/cub.cuh>

#include "common.cuh"

namespace extfem {

struct SolverWork {
    // row-major (CSR) view of the CSC matrix: rowptr/colidx plus perm (CSR slot -> CSC slot)
    void *rowptr = nullptr, *colidx = nullptr, *perm = nullptr;
    bool csr_ready = false;
    // CG vectors
    void *vec[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    void *partial = nullptr, *scalars = nullptr;
    long long n_alloc = 0;
    ~SolverWork()
    {
        for (void *p : {rowptr, colidx, perm, partial, scalars}) if (p) cudaFree(p);
        for (void *p : vec) if (p) cudaFree(p);
    }
};

__global__ void csr_keys_kernel(const long long *__restrict__ colptr, long long ncols, const int *__restrict__ rowval,
                                unsigned long long *keys, unsigned *vals)
{
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncols) return;
    for (long long p = colptr[c]; p < colptr[c + 1]; ++p) {
        keys[p] = ((unsigned long long)(unsigned)rowval[p] << 32) | (unsigned long long)c;
        vals[p] = (unsigned)p;
    }
}

__global__ void csr_finish_kernel(const unsigned long long *__restrict__ keys, long long nnz, long long nrows,
                                  long long *rowptr, int *colidx)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nnz) colidx[i] = (int)(keys[i] & 0xffffffffull);
    if (i <= nrows) {
        unsigned long long target = (unsigned long long)i << 32;
        long long lo = 0, hi = nnz;
        while (lo < hi) {
            long long mid = (lo + hi) >> 1;
            if (keys[mid] < target) lo = mid + 1; else hi = mid;
        }
        rowptr[i] = lo;
    }
}

static inline int build_csr(cudaStream_t st, long long nrows, long long ncols, long long nnz, const long long *colptr,
                            const int *rowval, SolverWork &W, long long *launches)
{
    if (W.csr_ready) return 0;
    if (nnz >= (1ll << 32)) return -10;
    unsigned long long *k1, *k2;
    unsigned *v1;
    void *tmp = nullptr;
    if (cudaMalloc(&k1, nnz * 8) || cudaMalloc(&k2, nnz * 8) || cudaMalloc(&v1, nnz * 4)) return -1;
    if (cudaMalloc(&W.perm, nnz * 4) || cudaMalloc(&W.colidx, nnz * 4) || cudaMalloc(&W.rowptr, (nrows + 1) * 8)) return -1;
    csr_keys_kernel<<<(unsigned)((ncols + 255) / 256), 256, 0, st>>>(colptr, ncols, rowval, k1, v1);
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, k1, k2, v1, (unsigned *)W.perm, nnz, 0, 64, st);
    if (cudaMalloc(&tmp, tb)) return -1;
    cub::DeviceRadixSort::SortPairs(tmp, tb, k1, k2, v1, (unsigned *)W.perm, nnz, 0, 64, st);
    long long n = std::max(nnz, nrows + 1);
    csr_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(k2, nnz, nrows, (long long *)W.rowptr, (int *)W.colidx);
    if (launches) *launches += 3;
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(k1); cudaFree(k2); cudaFree(v1); cudaFree(tmp);
    if (e != cudaSuccess) return -2;
    W.csr_ready = true;
    return 0;
}

// y = A*x, 8 lanes per row.  perm == nullptr: (ptr, idx, val) is already row-major
// (symmetric matrices: the CSC arrays of A are the CSR arrays of A^T = A).
constexpr int SPMV_LANES = 8;
__global__ void __launch_bounds__(256)
spmv_kernel(long long nrows, const long long *__restrict__ ptr, const int *__restrict__ idx, const unsigned *__restrict__ perm,
            const double *__restrict__ val, const double *__restrict__ x, double *__restrict__ y)
{
    long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / SPMV_LANES;
    int lane = threadIdx.x & (SPMV_LANES - 1);
    double s = 0.0;
    if (row < nrows) {
        long long p0 = ptr[row], p1 = ptr[row + 1];
        if (perm)
            for (long long p = p0 + lane; p < p1; p += SPMV_LANES) s += val[perm[p]] * x[idx[p]];
        else
            for (long long p = p0 + lane; p < p1; p += SPMV_LANES) s += val[p] * x[idx[p]];
    }
#pragma unroll
    for (int o = SPMV_LANES / 2; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, SPMV_LANES);
    if (row < nrows && lane == 0) y[row] = s;
}

static inline int csc_spmv(cudaStream_t st, long long nrows, long long ncols, long long nnz, const long long *colptr,
                           const int *rowval, const double *nzval, const double *x, double *y, SolverWork &W, long long *launches)
{
    if (int rc = build_csr(st, nrows, ncols, nnz, colptr, rowval, W, launches)) return rc;
    long long nthreads = nrows * SPMV_LANES;
    spmv_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, st>>>(nrows, (const long long *)W.rowptr, (const int *)W.colidx,
                                                                    (const unsigned *)W.perm, nzval, x, y);
    if (launches) *launches += 1;
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

__global__ void residual_kernel(long long n, const double *__restrict__ b, double *__restrict__ y)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = b[i] - y[i];
}

__device__ __forceinline__ long long find_in_column(const long long *colptr, const int *rowval, long long col, int row)
{
    long long lo = colptr[col], hi = colptr[col + 1] - 1;
    while (lo <= hi) {
        long long mid = (lo + hi) >> 1;
        int r = rowval[mid];
        if (r == row) return mid;
        if (r < row) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}

__global__ void penalties_kernel(long long n, const long long *__restrict__ dofs, const double *__restrict__ values, double penalty,
                                 const long long *__restrict__ colptr, const int *__restrict__ rowval, double *__restrict__ nzval,
                                 double *__restrict__ b, long long nrows, const double *__restrict__ owned, int *err)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long d = dofs[i] - 1;
    if (d < 0 || d >= nrows) { atomicExch(err, 1); return; }
    long long p = find_in_column(colptr, rowval, d, (int)d);
    if (p < 0) { atomicExch(err, 1); return; }
    nzval[p] = (owned && owned[d] == 0.0) ? 0.0 : penalty;
    b[d] = penalty * (values ? values[i] : 0.0);
}

// ---- Jacobi-preconditioned CG ------------------------------------------------------------------
__global__ void diag_inv_kernel(long long n, const long long *__restrict__ colptr, const int *__restrict__ rowval,
                                const double *__restrict__ nzval, double *__restrict__ dinv)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long p = find_in_column(colptr, rowval, i, (int)i);
    double d = p >= 0 ? nzval[p] : 1.0;
    dinv[i] = d != 0.0 ? 1.0 / d : 1.0;
}

constexpr int RED_BLOCKS = 592; // 4 per SM
// partial[b] = sum over the block's slice of a[i]*b[i] (fixed assignment -> deterministic)
__global__ void __launch_bounds__(256) dot_partial_kernel(long long n, const double *__restrict__ a, const double *__restrict__ b,
                                                          double *__restrict__ partial)
{
    __shared__ double sh[256];
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s += a[i] * b[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
// scalars: [0]=rz  [1]=pq  [2]=rz_new  [3]=alpha  [4]=beta  [5]=rr  [6]=bnorm2
__global__ void __launch_bounds__(256) dot_final_kernel(const double *__restrict__ partial, int nb, double *scalars, int slot, int mode)
{
    __shared__ double sh[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < nb; i += 256) s += partial[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) {
        scalars[slot] = sh[0];
        if (mode == 1) scalars[3] = scalars[0] / sh[0];                                   // alpha = rz / pq
        if (mode == 2) { scalars[4] = sh[0] / scalars[0]; scalars[0] = sh[0]; }           // beta = rz_new / rz ; rz = rz_new
    }
}
__global__ void cg_init_kernel(long long n, const double *__restrict__ b, const double *__restrict__ Ax, const double *__restrict__ dinv,
                               double *__restrict__ r, double *__restrict__ z, double *__restrict__ p)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double ri = b[i] - Ax[i];
    r[i] = ri;
    double zi = dinv[i] * ri;
    z[i] = zi;
    p[i] = zi;
}
__global__ void cg_update_xr_kernel(long long n, const double *__restrict__ scalars, const double *__restrict__ p,
                                    const double *__restrict__ q, const double *__restrict__ dinv, double *__restrict__ x,
                                    double *__restrict__ r, double *__restrict__ z)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double alpha = scalars[3];
    x[i] += alpha * p[i];
    double ri = r[i] - alpha * q[i];
    r[i] = ri;
    z[i] = dinv[i] * ri;
}
__global__ void cg_update_p_kernel(long long n, const double *__restrict__ scalars, const double *__restrict__ z, double *__restrict__ p)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    p[i] = z[i] + scalars[4] * p[i];
}

// CG for symmetric positive definite systems: the CSC arrays are used directly as CSR.
static inline int jacobi_cg(cudaStream_t st, long long n, long long nnz, const long long *colptr, const int *rowval,
                            const double *nzval, const double *b, double *x, double rtol, int maxit, int *iters, double *relres,
                            SolverWork &W, long long *launches)
{
    (void)nnz;
    if (W.n_alloc < n) {
        W.n_alloc = 0;
        for (auto &v : W.vec) { if (v) { cudaFree(v); v = nullptr; } if (cudaMalloc(&v, n * 8)) { v = nullptr; return -1; } }
        if (!W.partial && cudaMalloc(&W.partial, RED_BLOCKS * 8)) return -1;
        if (!W.scalars && cudaMalloc(&W.scalars, 8 * 8)) return -1;
        W.n_alloc = n;
    }
    double *r = (double *)W.vec[0], *z = (double *)W.vec[1], *p = (double *)W.vec[2], *q = (double *)W.vec[3], *dinv = (double *)W.vec[4];
    double *partial = (double *)W.partial, *sc = (double *)W.scalars;
    unsigned gb = (unsigned)((n + 255) / 256), gs = (unsigned)((n * SPMV_LANES + 255) / 256);
    long long nl = 0;
    auto dot = [&](const double *a, const double *c, int slot, int mode) {
        dot_partial_kernel<<<RED_BLOCKS, 256, 0, st>>>(n, a, c, partial);
        dot_final_kernel<<<1, 256, 0, st>>>(partial, RED_BLOCKS, sc, slot, mode);
        nl += 2;
    };
    diag_inv_kernel<<<gb, 256, 0, st>>>(n, colptr, rowval, nzval, dinv);
    spmv_kernel<<<gs, 256, 0, st>>>(n, colptr, rowval, nullptr, nzval, x, q);
    cg_init_kernel<<<gb, 256, 0, st>>>(n, b, q, dinv, r, z, p);
    nl += 3;
    dot(b, b, 6, 0);
    dot(r, z, 0, 0);
    dot(r, r, 5, 0);
    double h[8];
    if (cudaMemcpyAsync(h, sc, 64, cudaMemcpyDeviceToHost, st) || cudaStreamSynchronize(st)) return -2;
    // the residual is measured against the INITIAL residual |b - A x0| (== |b| for x0 = 0): with penalised Dirichlet rows
    // (1e30 * value) in b, |b| would hide the interior residual; callers pass x0 with the boundary values set
    double bnorm = std::sqrt(h[5]);
    if (bnorm == 0.0) bnorm = 1.0;
    double res = std::sqrt(h[5]) / bnorm;
    int it = 0;
    while (res > rtol && it < maxit) {
        spmv_kernel<<<gs, 256, 0, st>>>(n, colptr, rowval, nullptr, nzval, p, q);
        dot(p, q, 1, 1);
        cg_update_xr_kernel<<<gb, 256, 0, st>>>(n, sc, p, q, dinv, x, r, z);
        dot(r, z, 2, 2);
        dot(r, r, 5, 0);
        cg_update_p_kernel<<<gb, 256, 0, st>>>(n, sc, z, p);
        nl += 3;
        ++it;
        if (cudaMemcpyAsync(h, sc, 64, cudaMemcpyDeviceToHost, st) || cudaStreamSynchronize(st)) return -2;
        if (!(h[1] > 0.0) || !std::isfinite(h[5])) { *iters = it; *relres = std::sqrt(h[5]) / bnorm; if (launches) *launches += nl; return -4; }
        res = std::sqrt(h[5]) / bnorm;
    }
    *iters = it;
    *relres = res;
    if (launches) *launches += nl;
    return 0;
}

} // namespace extfem
