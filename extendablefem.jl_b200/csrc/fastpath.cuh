// fastpath.cuh -- fused assembly kernels for the headline configurations.
//
// Scalar H1 Lagrange spaces on affine simplices with the standard kernel:
//     Laplace  BilinearOperator([grad(u)])   A_loc[i,j] = sum_{d<=e} G_de * S^{de}[i][j]
//     mass     BilinearOperator([id(u)])     A_loc[i,j] = |T| f  * S^{0}[i][j]
// where G = factor |T| J^-1 J^-T is the per-cell geometry factor (phase 0, coalesced) and
// S^{de}[i][j] = sum_q w_q d_d phi_i(q) d_e phi_j(q) (+ transposed term) are reference tables
// computed on the host WITH THE SAME quadrature rule and reference basis as the generic path,
// i.e. the reference's sum over quadrature points (bilinear_operator.jl:876-916) reassociated.
//
// The matrix is produced by an OWNER-COMPUTES gather without atomics and without an
// intermediate cell-local buffer: one thread owns one CSC column, computes the local column of
// every adjacent cell on the fly and accumulates it into shared memory laid out as the CSC
// segment of its CTA's contiguous column chunk, which is then written with unit-stride stores
// (each nzval byte is written exactly once).  Inside a chunk, columns are assigned to lanes
// sorted by a signature of their adjacency so that warps are (nearly) divergence-free, and the
// per-(column, cell) records are stored warp-transposed (ELL per warp) so every record load is a
// fully coalesced 128-byte access per word.
#pragma once
#include "common.cuh"

namespace extfem {

template <int DIM>
__global__ void cell_volumes_kernel(long long ncells, const double *__restrict__ coords, const int *__restrict__ cellnodes,
                                    double *__restrict__ vol)
{
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    const int *cn = cellnodes + c * (DIM + 1);
    double A[DIM][DIM];
    const double *p0 = coords + (size_t)cn[0] * DIM;
    for (int r = 0; r < DIM; ++r) {
        const double *pr = coords + (size_t)cn[r + 1] * DIM;
        for (int d = 0; d < DIM; ++d) A[d][r] = pr[d] - p0[d];
    }
    double det;
    if (DIM == 1) det = A[0][0];
    else if (DIM == 2) det = A[0][0] * A[1 % DIM][1 % DIM] - A[0][1 % DIM] * A[1 % DIM][0];
    else {
        constexpr int I1 = 1 % DIM, I2 = 2 % DIM;
        det = A[0][0] * (A[I1][I1] * A[I2][I2] - A[I1][I2] * A[I2][I1]) + A[0][I1] * (A[I1][I2] * A[I2][0] - A[I1][0] * A[I2][I2]) +
              A[0][I2] * (A[I1][0] * A[I2][I1] - A[I1][I1] * A[I2][0]);
    }
    const double fact = DIM == 1 ? 1.0 : (DIM == 2 ? 2.0 : 6.0);
    vol[c] = fabs(det) / fact;
}

static inline int launch_cell_volumes(cudaStream_t st, int dim, long long ncells, const double *coords, const int *cellnodes, double *vol)
{
    unsigned g = (unsigned)((ncells + 255) / 256);
    if (dim == 1) cell_volumes_kernel<1><<<g, 256, 0, st>>>(ncells, coords, cellnodes, vol);
    else if (dim == 2) cell_volumes_kernel<2><<<g, 256, 0, st>>>(ncells, coords, cellnodes, vol);
    else cell_volumes_kernel<3><<<g, 256, 0, st>>>(ncells, coords, cellnodes, vol);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// ------------------------------------------------------------------------------------------------
constexpr int FP_THREADS = 448;         // columns (threads) per chunk: 14 warps
constexpr int FP_WARPS = FP_THREADS / 32;
constexpr int FP_MAXNNZ = 11008;        // CSC entries per chunk held in shared memory
constexpr int FP_SMEM_DOUBLES = FP_MAXNNZ + FP_MAXNNZ / 16 + 8;
enum { FP_FORM_LAPLACE = 0, FP_FORM_MASS = 1 };

__host__ __device__ constexpr int fp_ng(int dim, int form) { return form == FP_FORM_MASS ? 1 : dim * (dim + 1) / 2; }
// number of 32-bit words of one (column, cell) record: cell id, then 1 + NS bytes (kl, pos[NS])
__host__ __device__ constexpr int fp_rw(int ns) { return 1 + (1 + ns + 3) / 4; }

// reference tables S[g][t][kl] of the current launch (uploaded per operator)
__constant__ double c_fp_S[6 * 10 * 10];

// phase 0: per-cell geometry factor  G = factor * |T| * J^-1 J^-T  (upper triangle) or factor*|T|
template <int DIM, int FORM>
__global__ void __launch_bounds__(256)
fp_geo_kernel(long long ncells, const double *__restrict__ coords, const int *__restrict__ cellnodes,
              const int *__restrict__ regions, const double *__restrict__ vol, double factor, int nregions,
              const int *__restrict__ visit /* device copy of regions list */, double *__restrict__ geo)
{
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    constexpr int NG = fp_ng(DIM, FORM);
    double f = factor * vol[c];
    if (nregions > 0) {
        int reg = regions[c], vis = 0;
        for (int k = 0; k < nregions; ++k) vis |= (visit[k] == reg);
        if (!vis) f = 0.0;
    }
    double *g = geo + c * NG;
    if (FORM == FP_FORM_MASS) { g[0] = f; return; }
    const int *cn = cellnodes + c * (DIM + 1);
    double A[DIM][DIM], B[DIM][DIM]; // B = A^-1 (rows = gradients of lambda_1..DIM)
    const double *p0 = coords + (size_t)cn[0] * DIM;
#pragma unroll
    for (int r = 0; r < DIM; ++r) {
        const double *pr = coords + (size_t)cn[r + 1] * DIM;
#pragma unroll
        for (int d = 0; d < DIM; ++d) A[d][r] = pr[d] - p0[d];
    }
    if (DIM == 1) {
        B[0][0] = 1.0 / A[0][0];
    } else if (DIM == 2) {
        constexpr int I1 = 1 % DIM;
        double id = 1.0 / (A[0][0] * A[I1][I1] - A[0][I1] * A[I1][0]);
        B[0][0] = A[I1][I1] * id; B[0][I1] = -A[0][I1] * id; B[I1][0] = -A[I1][0] * id; B[I1][I1] = A[0][0] * id;
    } else {
        constexpr int I1 = 1 % DIM, I2 = 2 % DIM;
        double c00 = A[I1][I1] * A[I2][I2] - A[I1][I2] * A[I2][I1];
        double c01 = A[I1][I2] * A[I2][0] - A[I1][0] * A[I2][I2];
        double c02 = A[I1][0] * A[I2][I1] - A[I1][I1] * A[I2][0];
        double id = 1.0 / (A[0][0] * c00 + A[0][I1] * c01 + A[0][I2] * c02);
        B[0][0] = c00 * id; B[I1][0] = c01 * id; B[I2][0] = c02 * id;
        B[0][I1] = (A[0][I2] * A[I2][I1] - A[0][I1] * A[I2][I2]) * id;
        B[I1][I1] = (A[0][0] * A[I2][I2] - A[0][I2] * A[I2][0]) * id;
        B[I2][I1] = (A[0][I1] * A[I2][0] - A[0][0] * A[I2][I1]) * id;
        B[0][I2] = (A[0][I1] * A[I1][I2] - A[0][I2] * A[I1][I1]) * id;
        B[I1][I2] = (A[0][I2] * A[I1][0] - A[0][0] * A[I1][I2]) * id;
        B[I2][I2] = (A[0][0] * A[I1][I1] - A[0][I1] * A[I1][0]) * id;
    }
    int o = 0;
#pragma unroll
    for (int d = 0; d < DIM; ++d)
#pragma unroll
        for (int e = d; e < DIM; ++e) {
            double s = 0.0;
#pragma unroll
            for (int x = 0; x < DIM; ++x) s += B[d][x] * B[e][x];
            g[o++] = f * s;
        }
}

struct FastPlanDev {
    int nchunks;
    const int *chunkptr;          // [nchunks+1] column ranges
    const int *slotcol;           // [nchunks*FP_THREADS] column of sorted slot (-1: idle)
    const int *warpniter;         // [nchunks*FP_WARPS]
    const long long *warpoff;     // [nchunks*FP_WARPS] offset (in u32 words) of the warp's record block
    const unsigned *rec;          // warp-transposed records
};

struct FastArgs {
    FastPlanDev plan;
    const long long *colptr;
    double *nzval;
    const double *geo;
    int overwrite;
};

__device__ __forceinline__ int fp_pad(int i) { return i + (i >> 4); }

// local column KL of the cell matrix, accumulated at the byte positions of the record
template <int NS, int NG, int KL>
__device__ __forceinline__ void fp_accumulate(const double (&G)[NG], double *__restrict__ a, int aoff, const unsigned (&w)[fp_rw(NS)])
{
#pragma unroll
    for (int t = 0; t < NS; ++t) {
        double v = 0.0;
#pragma unroll
        for (int g = 0; g < NG; ++g) v += G[g] * c_fp_S[(g * NS + t) * NS + KL];
        const int byte = 1 + t;                                   // byte 0 of word 1 is kl
        const int pos = (w[1 + byte / 4] >> (8 * (byte % 4))) & 0xff;
        a[fp_pad(aoff + pos)] += v;
    }
}

template <int NS, int NG, int KL>
struct FpSwitch {
    __device__ __forceinline__ static void run(int kl, const double (&G)[NG], double *a, int aoff, const unsigned (&w)[fp_rw(NS)])
    {
        if (kl == KL) fp_accumulate<NS, NG, KL>(G, a, aoff, w);
        else FpSwitch<NS, NG, KL + 1>::run(kl, G, a, aoff, w);
    }
};
template <int NS, int NG>
struct FpSwitch<NS, NG, NS> {
    __device__ __forceinline__ static void run(int, const double (&)[NG], double *, int, const unsigned (&)[fp_rw(NS)]) {}
};

template <int NS, int NG>
__global__ void __launch_bounds__(FP_THREADS, 2)
fp_gather_kernel(const __grid_constant__ FastArgs A)
{
    extern __shared__ double acc[];
    constexpr int RW = fp_rw(NS);
    const int chunk = blockIdx.x;
    const int k0 = A.plan.chunkptr[chunk], k1 = A.plan.chunkptr[chunk + 1];
    const long long base = A.colptr[k0];
    const int n = (int)(A.colptr[k1] - base);
    if (A.overwrite)
        for (int i = threadIdx.x; i < n; i += FP_THREADS) acc[fp_pad(i)] = 0.0;
    else
        for (int i = threadIdx.x; i < n; i += FP_THREADS) acc[fp_pad(i)] = A.nzval[base + i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k = A.plan.slotcol[(size_t)chunk * FP_THREADS + threadIdx.x];
    const int niter = A.plan.warpniter[chunk * FP_WARPS + warp];
    const unsigned *rec = A.plan.rec + A.plan.warpoff[chunk * FP_WARPS + warp] + lane;
    const int aoff = k >= 0 ? (int)(A.colptr[k] - base) : 0;
    unsigned w[RW], wn[RW];
    if (niter > 0) {
#pragma unroll
        for (int j = 0; j < RW; ++j) wn[j] = __ldg(rec + j * 32);
    }
    for (int r = 0; r < niter; ++r) {
#pragma unroll
        for (int j = 0; j < RW; ++j) w[j] = wn[j];
        if (r + 1 < niter) {
#pragma unroll
            for (int j = 0; j < RW; ++j) wn[j] = __ldg(rec + ((size_t)(r + 1) * RW + j) * 32);
        }
        const int cell = (int)w[0];
        if (cell < 0) continue;
        double G[NG];
        const double *gp = A.geo + (size_t)cell * NG;
        if (NG % 2 == 0) {
#pragma unroll
            for (int g = 0; g < NG; g += 2) {
                double2 t2 = __ldg(reinterpret_cast<const double2 *>(gp + g));
                G[g] = t2.x; G[(g + 1) % NG] = t2.y;
            }
        } else {
#pragma unroll
            for (int g = 0; g < NG; ++g) G[g] = __ldg(gp + g);
        }
        const int kl = w[1] & 0xff;
        FpSwitch<NS, NG, 0>::run(kl, G, acc, aoff, w);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += FP_THREADS) A.nzval[base + i] = acc[fp_pad(i)];
}

// ---- plan construction (setup, once per pattern) -----------------------------------------------
__global__ void fp_signature_kernel(long long ncols, const long long *__restrict__ adjptr, const unsigned char *__restrict__ adjloc,
                                    unsigned long long *__restrict__ sig)
{
    long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= ncols) return;
    long long p0 = adjptr[k], p1 = adjptr[k + 1];
    unsigned long long h = 1469598103934665603ull;
    for (long long p = p0; p < p1; ++p) { h ^= adjloc[p]; h *= 1099511628211ull; }
    sig[k] = ((unsigned long long)(p1 - p0) << 48) | (h & 0xffffffffffffull);
}

// one CTA per chunk: sort the chunk's columns by signature, emit slot->column and per-warp iteration counts
__global__ void __launch_bounds__(512)
fp_sort_kernel(const int *__restrict__ chunkptr, const unsigned long long *__restrict__ sig, const long long *__restrict__ adjptr,
               int *__restrict__ slotcol, int *__restrict__ warpniter)
{
    __shared__ unsigned long long key[512];
    __shared__ int val[512];
    const int chunk = blockIdx.x, k0 = chunkptr[chunk], k1 = chunkptr[chunk + 1];
    const int nc = k1 - k0, t = threadIdx.x;
    key[t] = t < nc ? sig[k0 + t] : ~0ull;
    val[t] = t < nc ? k0 + t : -1;
    __syncthreads();
    for (int k = 2; k <= 512; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            int ixj = t ^ j;
            if (ixj > t) {
                bool asc = ((t & k) == 0);
                unsigned long long a = key[t], b = key[ixj];
                int va = val[t], vb = val[ixj];
                // ties broken by column index to keep the order deterministic
                bool gt = (a > b) || (a == b && va > vb);
                if (gt == asc) { key[t] = b; key[ixj] = a; val[t] = vb; val[ixj] = va; }
            }
            __syncthreads();
        }
    if (t < FP_THREADS) slotcol[(size_t)chunk * FP_THREADS + t] = val[t];
    __syncthreads();
    if (t < FP_WARPS) {
        int m = 0;
        for (int l = 0; l < 32; ++l) {
            int c = val[t * 32 + l];
            if (c >= 0) m = max(m, (int)(adjptr[c + 1] - adjptr[c]));
        }
        warpniter[chunk * FP_WARPS + t] = m;
    }
}

template <typename PosT>
__global__ void fp_fill_kernel(long long nslots, int ns, int rw, int posstride, const int *__restrict__ slotcol,
                               const int *__restrict__ warpniter, const long long *__restrict__ warpoff,
                               const long long *__restrict__ adjptr, const int *__restrict__ adjcell,
                               const unsigned char *__restrict__ adjloc, const PosT *__restrict__ posmap, unsigned *__restrict__ rec)
{
    long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslots) return;
    long long wg = s >> 5;
    int lane = (int)(s & 31);
    int niter = warpniter[wg];
    unsigned *out = rec + warpoff[wg] + lane;
    int k = slotcol[s];
    long long p0 = 0, p1 = 0;
    if (k >= 0) { p0 = adjptr[k]; p1 = adjptr[k + 1]; }
    for (int r = 0; r < niter; ++r) {
        unsigned w[8] = {0xffffffffu, 0, 0, 0, 0, 0, 0, 0};
        long long p = p0 + r;
        if (p < p1) {
            w[0] = (unsigned)adjcell[p];
            unsigned bytes[28];
            bytes[0] = adjloc[p];
            for (int t = 0; t < ns; ++t) bytes[1 + t] = (unsigned)posmap[p * posstride + t] & 0xff;
            for (int b = 0; b < 1 + ns; ++b) w[1 + b / 4] |= bytes[b] << (8 * (b % 4));
        }
        for (int j = 0; j < rw; ++j) out[((size_t)r * rw + j) * 32] = w[j];
    }
}

} // namespace extfem
