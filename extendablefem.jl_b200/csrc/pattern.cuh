// pattern.cuh -- integer-only construction of the dof->cell adjacency, the structural CSC
// pattern and the byte position map (cell-local entry -> position inside its CSC column).
//
// Replaces what ExtendableSparse does implicitly on first assembly (linked-list insertion in
// rawupdateindex! + merge in flush!, call sites bilinear_operator.jl:926,993): the pattern is
// computed once, bit-exactly (sorted unique integers), and reused by every assembly.
#pragma once
#include <cub/cub.cuh>

#include "common.cuh"
#include "gather.cuh"

namespace extfem {

constexpr int PAT_CAP = 2048;  // candidate rows per column handled by one warp
constexpr int PAT_WARPS = 4;

// ---- adjacency (transpose of the cell dof map), sorted by (dof, cell) -----------------------
__global__ void adj_keys_kernel(const int *__restrict__ celldofs, long long n, int nd, unsigned long long *keys,
                                unsigned char *vals)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long cell = i / nd;
    keys[i] = ((unsigned long long)(unsigned)celldofs[i] << 32) | (unsigned long long)cell;
    vals[i] = (unsigned char)(i - cell * nd);
}

__global__ void adj_ptr_kernel(const unsigned long long *__restrict__ keys, long long n, long long ndofs,
                               long long *adjptr)
{
    long long d = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (d > ndofs) return;
    unsigned long long target = (unsigned long long)d << 32;
    long long lo = 0, hi = n;
    while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        if (keys[mid] < target) lo = mid + 1; else hi = mid;
    }
    adjptr[d] = lo;
}

__global__ void adj_cells_kernel(const unsigned long long *__restrict__ keys, long long n, int *adjcell)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) adjcell[i] = (int)(keys[i] & 0xffffffffull);
}

// ---- warp-level bitonic sort in shared memory ----------------------------------------------
__device__ __forceinline__ void warp_bitonic_sort(int *s, int n, int lane)
{
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < n; i += 32) {
                int ixj = i ^ j;
                if (ixj > i) {
                    int a = s[i], b = s[ixj];
                    bool asc = ((i & k) == 0);
                    if ((a > b) == asc) { s[i] = b; s[ixj] = a; }
                }
            }
            __syncwarp();
        }
    }
}

struct PatArgs {
    long long ncolsb;                 // columns of this column block
    long long colbase;                // global index of the block's first column
    const long long *adjptr;          // adjacency of the column space
    const int *adjcell;
    int nrb;
    unsigned char coupled[MAXBLOCKS]; // row block r present in this column block
    const int *celldofs[MAXBLOCKS];   // row spaces' cell dof maps (0-based)
    int nd[MAXBLOCKS];
    long long rowoff[MAXBLOCKS];
    int rowlocoff[MAXBLOCKS];         // offset of the row block inside the posmap row
    int NRpat;
    int *error;
};

// gathers the candidate rows of column k into shared memory; returns count (or -1 on overflow)
__device__ __forceinline__ int pat_load_candidates(const PatArgs &P, long long k, int *s, int lane)
{
    long long p0 = P.adjptr[k], p1 = P.adjptr[k + 1];
    int ncoupled = 0;
    for (int r = 0; r < P.nrb; ++r) if (P.coupled[r]) ncoupled += P.nd[r];
    long long total = (p1 - p0) * ncoupled;
    if (total > PAT_CAP) return -1;
    int off = 0;
    for (int r = 0; r < P.nrb; ++r) {
        if (!P.coupled[r]) continue;
        int nd = P.nd[r];
        int cnt = (int)(p1 - p0) * nd;
        for (int i = lane; i < cnt; i += 32) {
            int pi = i / nd, t = i - pi * nd;
            long long cell = P.adjcell[p0 + pi];
            s[off + i] = (int)(P.rowoff[r] + P.celldofs[r][cell * nd + t]);
        }
        off += cnt;
    }
    int n = 1;
    while (n < off) n <<= 1;
    if (n < 32) n = 32;
    for (int i = off + lane; i < n; i += 32) s[i] = 0x7fffffff;
    __syncwarp();
    return n;
}

// in-place unique of a sorted array by one warp; returns number of unique entries (< INT_MAX)
__device__ __forceinline__ int warp_unique(int *s, int n, int lane)
{
    int count = 0;
    for (int i0 = 0; i0 < n; i0 += 32) {
        int i = i0 + lane;
        int v = s[i];
        int prev = (i == 0) ? -1 : s[i - 1];
        bool keep = (v != prev) && (v != 0x7fffffff);
        unsigned m = __ballot_sync(0xffffffffu, keep);
        __syncwarp();
        if (keep) s[count + __popc(m & ((1u << lane) - 1u))] = v;
        count += __popc(m);
        __syncwarp();
    }
    return count;
}

__global__ void __launch_bounds__(PAT_WARPS * 32) pattern_count_kernel(const __grid_constant__ PatArgs P, long long *collen)
{
    __shared__ int buf[PAT_WARPS][PAT_CAP];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long k = (long long)blockIdx.x * PAT_WARPS + warp;
    if (k >= P.ncolsb) return;
    int *s = buf[warp];
    int n = pat_load_candidates(P, k, s, lane);
    if (n < 0) { if (lane == 0) atomicExch(P.error, 1); return; }
    warp_bitonic_sort(s, n, lane);
    // count uniques without compaction
    int count = 0;
    for (int i0 = 0; i0 < n; i0 += 32) {
        int i = i0 + lane;
        int v = s[i];
        int prev = (i == 0) ? -1 : s[i - 1];
        count += __popc(__ballot_sync(0xffffffffu, (v != prev) && (v != 0x7fffffff)));
    }
    if (lane == 0) collen[P.colbase + k] = count;
}

template <typename PosT>
__global__ void __launch_bounds__(PAT_WARPS * 32)
pattern_fill_kernel(const __grid_constant__ PatArgs P, const long long *__restrict__ colptr, int *__restrict__ rowval,
                    PosT *__restrict__ posmap)
{
    __shared__ int buf[PAT_WARPS][PAT_CAP];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long k = (long long)blockIdx.x * PAT_WARPS + warp;
    if (k >= P.ncolsb) return;
    int *s = buf[warp];
    int n = pat_load_candidates(P, k, s, lane);
    if (n < 0) return;
    warp_bitonic_sort(s, n, lane);
    int m = warp_unique(s, n, lane);
    long long base = colptr[P.colbase + k];
    for (int i = lane; i < m; i += 32) rowval[base + i] = s[i];
    // position map: for every (pair, pattern-local row) the rank of that row in the column
    long long p0 = P.adjptr[k], p1 = P.adjptr[k + 1];
    const PosT SENT = (PosT)~(PosT)0;
    for (int r = 0; r < P.nrb; ++r) {
        int nd = P.nd[r];
        int cnt = (int)(p1 - p0) * nd;
        for (int i = lane; i < cnt; i += 32) {
            int pi = i / nd, t = i - pi * nd;
            PosT pos = SENT;
            if (P.coupled[r]) {
                long long cell = P.adjcell[p0 + pi];
                int row = (int)(P.rowoff[r] + P.celldofs[r][cell * nd + t]);
                int lo = 0, hi = m - 1;
                while (lo < hi) {
                    int mid = (lo + hi) >> 1;
                    if (s[mid] < row) lo = mid + 1; else hi = mid;
                }
                pos = (PosT)lo;
            }
            posmap[(p0 + pi) * P.NRpat + P.rowlocoff[r] + t] = pos;
        }
    }
}

__global__ void fill_i64_kernel(long long *p, long long n, long long v)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void convert_index_kernel(const void *__restrict__ in, int index_bytes, long long n, long long nmax, int *__restrict__ out,
                                     int *__restrict__ err)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long v = index_bytes == 8 ? reinterpret_cast<const long long *>(in)[i] : reinterpret_cast<const int *>(in)[i];
    if (v < 1 || v > nmax) { atomicExch(err, 1); v = 1; }
    out[i] = (int)(v - 1);
}

__global__ void csc_export_kernel(const long long *__restrict__ colptr, long long ncols1, const int *__restrict__ rowval,
                                  long long nnz, long long *__restrict__ colptr_out, long long *__restrict__ rowval_out)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (colptr_out && i < ncols1) colptr_out[i] = colptr[i] + 1;
    if (rowval_out && i < nnz) rowval_out[i] = (long long)rowval[i] + 1;
}

// ---- lower triangle of a square CSC matrix (rows sorted per column: the entries with row >= column are a SUFFIX of the column) --
// lstart[j] = index of the first entry of column j with row >= j; lcount[j] = entries from there to the end of the column
__global__ void lower_start_kernel(long long ncols, const long long *__restrict__ colptr, const int *__restrict__ rowval,
                                   long long *__restrict__ lstart, long long *__restrict__ lcount)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncols) return;
    long long lo = colptr[j], hi = colptr[j + 1];
    const long long end = hi;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (rowval[mid] < j) lo = mid + 1; else hi = mid;
    }
    lstart[j] = lo; lcount[j] = end - lo;
}

// LPC lanes per column of [c0, c1): out[lcolptr[j] + k] = src[lstart[j] + k]  (values: T = double; row indices: int -> 1-based
// int64).  The lower part of a P2 column holds 10 to 33 entries: 8 lanes per column keep the lanes busy and let one warp run
// four independent (start, offset) -> copy chains; consecutive columns are adjacent in `out`.
template <typename TI, typename TO, int LPC>
__global__ void __launch_bounds__(256) lower_pack_kernel(long long c0, long long c1, const long long *__restrict__ lstart,
                                                         const long long *__restrict__ lcolptr, const TI *__restrict__ src,
                                                         TO *__restrict__ out, TO add)
{
    const long long j = c0 + ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LPC;
    const int lane = threadIdx.x % LPC;
    if (j >= c1) return;
    const long long s = lstart[j], o = lcolptr[j], n = lcolptr[j + 1] - o;
    for (long long k = lane; k < n; k += LPC) out[o + k] = (TO)src[s + k] + add;
}

__global__ void add_one_kernel(long long n, const long long *__restrict__ in, long long *__restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] + 1;
}

} // namespace extfem
