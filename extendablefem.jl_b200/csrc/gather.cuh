// gather.cuh -- atomic-free reduction of cell-local contributions into the global CSC values.
//
// Replaces ExtendableSparse's rawupdateindex!(A, +, v, i, j) / flush!
// (bilinear_operator.jl:921-930,993) by an OWNER-COMPUTES gather: one thread owns one matrix
// column (one ansatz dof), walks the cells adjacent to that dof in ascending cell order -- the
// reference's accumulation order -- and adds the cell-local column into shared-memory
// accumulators laid out exactly like the CSC segment of its CTA's column chunk; the chunk is
// then written with unit-stride stores.  Positions inside the column come from a precomputed
// byte map (posmap), so the hot path does integer loads only: no search, no atomics.
#pragma once
#include "common.cuh"

namespace extfem {

constexpr int MAXBLOCKS = 8;
constexpr int GATHER_THREADS = 128;
constexpr int GATHER_MAXNNZ = 6112; // doubles of a column chunk in shared memory (48 KB per CTA with the 32 dummy targets of the warp kernel)

template <typename PosT>
struct GatherArgs {
    const int *chunkptr;            // [nchunks+1] global column index ranges (never straddle column blocks)
    const long long *colptr;        // [ncols+1] 0-based
    double *nzval;
    int ncb;
    long long coloff[MAXBLOCKS + 1];
    const long long *adjptr[MAXBLOCKS];
    const int *adjcell[MAXBLOCKS];
    const unsigned char *adjloc[MAXBLOCKS];
    const PosT *posmap[MAXBLOCKS];
    int collocoff[MAXBLOCKS];       // operator-local offset of this column block, -1: untouched by the operator
    int colnd[MAXBLOCKS];
    int NRpat;                      // posmap row width (all row blocks of the pattern)
    int NRop, NCop;                 // operator-local matrix: loc[cell][NCop][NRop]
    int nrows_g;                    // rows gathered per pair
    int rowmap[MAXLOC];             // gathered row t -> pattern-local row (posmap index)
    int rowsrc[MAXLOC];             // gathered row t -> operator-local index on the "other" side
    const double *loc;
    int overwrite;                  // 1: write sums; 0: add to existing values
    int transposed;                 // 0: A[row,col] += loc[col][row]; 1: transposed copy
    double scale;                   // transposed_copy factor (+1 / -1)
};

template <typename PosT>
__global__ void __launch_bounds__(GATHER_THREADS)
gather_columns_kernel(const __grid_constant__ GatherArgs<PosT> g)
{
    __shared__ double acc[GATHER_MAXNNZ];
    const int k0 = g.chunkptr[blockIdx.x], k1 = g.chunkptr[blockIdx.x + 1];
    const long long base = g.colptr[k0];
    const int n = (int)(g.colptr[k1] - base);
    int cb = 0;
    while (cb + 1 < g.ncb && k0 >= g.coloff[cb + 1]) ++cb;
    const int cloc = g.collocoff[cb];
    if (cloc < 0) { // column block untouched by this operator
        if (g.overwrite)
            for (int i = threadIdx.x; i < n; i += blockDim.x) g.nzval[base + i] = 0.0;
        return;
    }
    if (g.overwrite)
        for (int i = threadIdx.x; i < n; i += blockDim.x) acc[i] = 0.0;
    else
        for (int i = threadIdx.x; i < n; i += blockDim.x) acc[i] = g.nzval[base + i];
    __syncthreads();
    const int k = k0 + threadIdx.x;
    if (k < k1) {
        double *a = acc + (g.colptr[k] - base);
        const long long kk = k - g.coloff[cb];
        const long long p0 = g.adjptr[cb][kk], p1 = g.adjptr[cb][kk + 1];
        const PosT SENT = (PosT)~(PosT)0;
        for (long long p = p0; p < p1; ++p) {
            const long long cell = g.adjcell[cb][p];
            const int kl = cloc + g.adjloc[cb][p];
            const PosT *pm = g.posmap[cb] + p * g.NRpat;
            if (!g.transposed) {
                const double *src = g.loc + (cell * g.NCop + kl) * g.NRop;
                for (int t = 0; t < g.nrows_g; ++t) {
                    PosT pos = pm[g.rowmap[t]];
                    if (pos != SENT) a[pos] += src[g.rowsrc[t]];
                }
            } else {
                // column = test dof (operator-local row kl); rows = ansatz dofs (operator-local cols)
                const double *src = g.loc + (cell * g.NCop) * g.NRop + kl;
                for (int t = 0; t < g.nrows_g; ++t) {
                    PosT pos = pm[g.rowmap[t]];
                    if (pos != SENT) a[pos] += g.scale * src[(size_t)g.rowsrc[t] * g.NRop];
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) g.nzval[base + i] = acc[i];
}

// The same reduction with the lanes of a warp on the ROWS of a cell-local column: the column (NRop contiguous doubles) and its
// position bytes are read with one coalesced load each instead of 32 lanes walking 32 different cells.  A warp owns a run of at
// most 32 consecutive columns of its CTA's chunk and streams through their (column, adjacent cell) pairs -- contiguous in the
// adjacency arrays -- eight at a time, so the loads of a batch are in flight together; the column of a pair comes from a ballot
// over the per-lane column ends (no search, no memory).  Pairs of one column are still added in ascending cell order: every sum
// has the same order of additions as the thread-per-column kernel (bit-identical).  Columns of at most 16 rows are handled two
// pairs at a time by the half warps.
constexpr int GATHER_WARP_THREADS = 256;   // 8 warps: runs of at most 16 columns

template <typename PosT, bool TR, int H>
__global__ void __launch_bounds__(GATHER_WARP_THREADS, 3)
gather_columns_warp_kernel(const __grid_constant__ GatherArgs<PosT> g)
{
    __shared__ double acc[GATHER_MAXNNZ + 32];   // + one dummy target per lane
    const int k0 = g.chunkptr[blockIdx.x], k1 = g.chunkptr[blockIdx.x + 1];
    const long long base = g.colptr[k0];
    const int n = (int)(g.colptr[k1] - base);
    int cb = 0;
    while (cb + 1 < g.ncb && k0 >= g.coloff[cb + 1]) ++cb;
    const int cloc = g.collocoff[cb];
    if (cloc < 0) { // column block untouched by this operator
        if (g.overwrite)
            for (int i = threadIdx.x; i < n; i += blockDim.x) g.nzval[base + i] = 0.0;
        return;
    }
    if (g.overwrite)
        for (int i = threadIdx.x; i < n; i += blockDim.x) acc[i] = 0.0;
    else
        for (int i = threadIdx.x; i < n; i += blockDim.x) acc[i] = g.nzval[base + i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const PosT SENT = (PosT)~(PosT)0;
    const long long *adjptr = g.adjptr[cb] - g.coloff[cb];      // indexed by the global column
    const int *adjcell = g.adjcell[cb];
    const unsigned char *adjloc = g.adjloc[cb];
    const PosT *posmap = g.posmap[cb];
    const int rstride = TR ? g.NRop : 1, cstride = TR ? 1 : g.NRop, CS = g.NCop * g.NRop;
    const double scale = TR ? g.scale : 1.0;
    constexpr int W = 32 / H;                                    // H pairs per warp step, W lanes per pair
    const int sub = lane / W, tl = lane - sub * W;
    constexpr int U = 8;
    // this warp's run of columns (the chunk has at most GATHER_THREADS = 32 nw columns)
    const int cpw = (k1 - k0 + nw - 1) / nw;
    const int ka = min(k0 + warp * cpw, k1), kb = min(ka + cpw, k1);
    // pair indices relative to the start of the run: 32-bit arithmetic in the loop, clamped loads instead of branches, and a
    // per-lane dummy accumulator for lanes without a target, so that the loop body is straight-line code
    const long long P0 = adjptr[ka];
    const int nrun = (int)(adjptr[kb] - P0);
    const int pend = lane < kb - ka ? (int)(adjptr[ka + lane + 1] - P0) : 0x7fffffff;
    const int cbase = lane < kb - ka ? (int)(g.colptr[ka + lane] - base) : 0;
    const int *adjc = adjcell + P0;
    const unsigned char *adjl = adjloc + P0;
    const int NRpat = g.NRpat, dummy = GATHER_MAXNNZ + lane;
    for (int t0 = 0; t0 < g.nrows_g && nrun > 0; t0 += W) {
        const bool on = t0 + tl < g.nrows_g;
        const int rm = on ? g.rowmap[t0 + tl] : 0;
        const double *lsrc = g.loc + (on ? g.rowsrc[t0 + tl] : 0) * rstride + cloc * cstride;
        const PosT *pm = posmap + P0 * NRpat + rm;
        // (cell, local column) of the pairs one batch ahead of the value loads
        int ncell[U], nkl[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int pp = min(u * H + sub, nrun - 1);
            ncell[u] = adjc[pp]; nkl[u] = adjl[pp];
        }
        for (int p = 0; p < nrun; p += U * H) {
            int ad[U];
            double v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int pu = p + u * H, pp = pu + sub;
                // column of the pair = number of columns of the run that end at or before it
                int col = __popc(__ballot_sync(0xffffffffu, pend <= pu));
                if (H == 2) { const int col1 = __popc(__ballot_sync(0xffffffffu, pend <= pu + 1)); col = sub ? col1 : col; }
                const int ab = __shfl_sync(0xffffffffu, cbase, min(col, 31));
                const PosT pos = pm[(long long)(min(pp, nrun - 1) * NRpat)];
                v[u] = lsrc[(long long)ncell[u] * CS + nkl[u] * cstride];
                ad[u] = (on && pp < nrun && pos != SENT) ? ab + (int)pos : dummy;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int pp = min(p + (U + u) * H + sub, nrun - 1);
                ncell[u] = adjc[pp]; nkl[u] = adjl[pp];
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const double w = TR ? scale * v[u] : v[u];
                if (H == 2) {
                    if (sub == 0) acc[ad[u]] += w;
                    __syncwarp();
                    if (sub == 1) acc[ad[u]] += w;
                } else acc[ad[u]] += w;
                __syncwarp();              // another lane may own the same row in the next cell
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) g.nzval[base + i] = acc[i];
}

struct GatherVecArgs {
    int nrb;
    long long rowoff[MAXBLOCKS + 1];
    const long long *adjptr[MAXBLOCKS];
    const int *adjcell[MAXBLOCKS];
    const unsigned char *adjloc[MAXBLOCKS];
    int rowlocoff[MAXBLOCKS]; // operator-local row offset of the block, -1: untouched
    int NRop;
    const double *bloc;       // [ncells][NRop]
    double *b;
    int overwrite;
    long long nrows;
};

// b[dof] += sum over adjacent cells (ascending) of the cell-local vector entry
// (linear_operator.jl:629-636, nonlinear_operator.jl:407-414 without the scattered +=)
__global__ void gather_rows_kernel(const __grid_constant__ GatherVecArgs g)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.nrows) return;
    int rb = 0;
    while (rb + 1 < g.nrb && i >= g.rowoff[rb + 1]) ++rb;
    int lo = g.rowlocoff[rb];
    if (lo < 0) { if (g.overwrite) g.b[i] = 0.0; return; }
    long long kk = i - g.rowoff[rb];
    double s = g.overwrite ? 0.0 : g.b[i];
    for (long long p = g.adjptr[rb][kk]; p < g.adjptr[rb][kk + 1]; ++p)
        s += g.bloc[(long long)g.adjcell[rb][p] * g.NRop + lo + g.adjloc[rb][p]];
    g.b[i] = s;
}

} // namespace extfem
