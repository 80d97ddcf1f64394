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
constexpr int GATHER_MAXNNZ = 6144; // doubles of shared memory per CTA (48 KB)

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
