// fastpath.cuh -- fused owner-computes assembly kernels for the headline configurations
// (DESIGN.md section 4).
//
// Scalar H1 Lagrange spaces on affine simplices with the standard kernel:
//     Laplace  BilinearOperator([grad(u)])   A_loc = factor |T| sum_q w_q grad phi_i . grad phi_j
//     mass     BilinearOperator([id(u)])     A_loc = factor |T| sum_q w_q phi_i phi_j
// i.e. the reference's sum over quadrature points (bilinear_operator.jl:876-916) reassociated into
// (per-cell geometry factor) x (reference-element table).
//
// The matrix is produced by an OWNER-COMPUTES gather without atomics and without a cell-local
// buffer: one thread owns one CSC column, computes the local column of every adjacent cell on the
// fly and accumulates it into shared memory laid out as the CSC segment of its CTA's contiguous
// column chunk, which is then written with unit-stride stores (every nzval byte is written exactly
// once).  Columns are sorted by a signature of their adjacency (number of cells + local indices) inside
// windows of FP_WINDOW consecutive columns and cut into chunks of FP_T columns, so that the lanes of a
// warp own structurally identical columns and run (nearly) divergence-free; chunks are grouped into
// shared-memory size classes so that short (edge-dof) and long (vertex-dof) columns both reach high
// occupancy, and where shared memory limits occupancy the local rows are split by entity class
// (vertex rows | edge rows: disjoint global rows) over two warp groups that share the accumulators.
// The per-(column, cell) records are stored lane-contiguous per warp round so that one vector load per
// lane fetches a record and a warp reads one contiguous block.
//
// Two evaluators produce the local column:
//   EvalTable<NS,NG>   v[t] = sum_g G[g] S[kl][t][g], reference tables S computed on the host WITH THE
//                      SAME quadrature rule and basis as the generic path (any form / dimension);
//   EvalBary<DIM,ORD>  closed form for the Laplace form of P1/P2 in terms of the barycentric Gram
//                      matrix D_ab = factor |T| grad(lambda_a).grad(lambda_b): every P2 entry is a
//                      combination of <= 4 D-values with small integer coefficients.  The host checks
//                      it against the tables before it is used (extfem.cu: verify_bary_against_tables).
#pragma once
#include "common.cuh"

namespace extfem {

// layout of the per-cell scratch records (geometry, RHS point values)
struct GeoLayout {
    int soa;            // 0: array of structures [cell][NG]; 1: [g][Npad] in period-P transposed cell order
    int permuted;       // the cell kernels read mesh arrays that were copied into the transposed order
    int P;              // period (1: identity)
    long long N, Npad;  // N = ceil(ncells / P), Npad = N * P
};

__host__ __device__ __forceinline__ long long geo_perm(const GeoLayout &Lg, long long c)
{
    return Lg.P == 1 ? c : (c % Lg.P) * Lg.N + c / Lg.P;
}

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
constexpr int FP_T = 128;               // columns per chunk (threads per warp group)
constexpr int FP_W = FP_T / 32;
constexpr int FP_WINDOW = 2048;         // columns per sorting window (multiple of FP_T)
constexpr int FP_NCLASS = 3;            // shared-memory size classes of chunks (doubles per chunk)
constexpr int FP_CAP0 = 3584, FP_CAP1 = 5632, FP_CAP2 = 8448;
__host__ __device__ constexpr int fp_cap(int cls) { return cls == 0 ? FP_CAP0 : (cls == 1 ? FP_CAP1 : FP_CAP2); }
constexpr int FP_KLBITS = 4, FP_CELLBITS = 32 - FP_KLBITS; // record word 0 = cell | kl << 28
enum { FP_FORM_LAPLACE = 0, FP_FORM_MASS = 1 };
enum { FP_GEO_METRIC = 0, FP_GEO_VOLUME = 1, FP_GEO_BARY = 2 };

__host__ __device__ constexpr int fp_ng(int dim, int geo) { return geo == FP_GEO_VOLUME ? 1 : dim * (dim + 1) / 2; }
// 32-bit words of one (column, cell) record: word 0 = cell | kl << 28, then NS position bytes
__host__ __device__ constexpr int fp_rw(int ns) { return 1 + (ns + 3) / 4; }

// reference tables S[kl][t][g] of the current launch (uploaded per operator)
__constant__ double c_fp_S[6 * 10 * 10];

// ---- local edge tables (ExtendableGrids local_celledgenodes order; grids.py TRI_EDGES / TET_EDGES) ------
template <int DIM> __host__ __device__ constexpr int fp_edge_a(int e)
{
    if (DIM == 1) return 0;
    if (DIM == 2) return e;                       // (0,1) (1,2) (2,0)
    return e < 3 ? 0 : (e < 5 ? 1 : 2);           // (0,1) (0,2) (0,3) (1,2) (1,3) (2,3)
}
template <int DIM> __host__ __device__ constexpr int fp_edge_b(int e)
{
    if (DIM == 1) return 1;
    if (DIM == 2) return (e + 1) % 3;
    return e < 3 ? e + 1 : (e < 5 ? e - 1 : 3);
}
// index of the off-diagonal D_ab (a < b) in the per-cell geometry record: pairs in lexicographic order
template <int DIM> __host__ __device__ constexpr int fp_pair_index(int a, int b)
{
    // lexicographic rank of (a,b), a<b, among pairs of {0..DIM}
    return a * (2 * DIM + 1 - a) / 2 + (b - a - 1);
}

// Closed-form local Laplace entry (t, kl) from the full symmetric barycentric Gram matrix M (already scaled:
// P1: M = D; P2: M = D/5 in 3D, D/3 in 2D -- see fp_bary_scale), valid for DIM = 2, 3.
//   P1:  A[t][kl] = D[t][kl]
//   P2:  vertex-vertex  3 M_ii | -M_ij
//        vertex i - edge (a,b)   c(i==a) M_ib + c(i==b) M_ia,  c = 3 | -1 (3D),  4 | 0 (2D)
//        edge (a,b) - edge (c,d) w(a==c) M_bd + w(a==d) M_bc + w(b==c) M_ad + w(b==d) M_ac,  w = 8 | 4
// (exact integrals of the barycentric polynomials: int l_i = 1/(d+1), int l_i l_j = (1+delta_ij)/((d+1)(d+2)))
__host__ __device__ constexpr double fp_bary_scale(int dim, int order) { return order == 1 ? 1.0 : (dim == 3 ? 0.2 : 1.0 / 3.0); }

template <int DIM, int ORDER>
__host__ __device__ __forceinline__ double fp_bary_acc(int t, int kl, const double (&M)[DIM + 1][DIM + 1], double r)
{
    // returns r + A_loc[t][kl] as a chain of fused multiply-adds
    constexpr int NV = DIM + 1;
    if (ORDER == 1) return r + M[t][kl];
    const bool tv = t < NV, kv = kl < NV;
    if (tv && kv) return t == kl ? fma(3.0, M[t][t], r) : r - M[t][kl];
    if (tv != kv) {
        const int i = tv ? t : kl, e = (tv ? kl : t) - NV;
        const int a = fp_edge_a<DIM>(e), b = fp_edge_b<DIM>(e);
        if (DIM == 3) {
            r = (i == a) ? fma(3.0, M[i][b], r) : r - M[i][b];
            r = (i == b) ? fma(3.0, M[i][a], r) : r - M[i][a];
        } else {
            if (i == a) r = fma(4.0, M[i][b], r);
            if (i == b) r = fma(4.0, M[i][a], r);
        }
        return r;
    }
    const int a = fp_edge_a<DIM>(t - NV), b = fp_edge_b<DIM>(t - NV), c = fp_edge_a<DIM>(kl - NV), d = fp_edge_b<DIM>(kl - NV);
    r = fma(a == c ? 8.0 : 4.0, M[b][d], r);
    r = fma(a == d ? 8.0 : 4.0, M[b][c], r);
    r = fma(b == c ? 8.0 : 4.0, M[a][d], r);
    r = fma(b == d ? 8.0 : 4.0, M[a][c], r);
    return r;
}

// full symmetric M from the NG stored off-diagonals (rows of the Gram matrix of barycentric gradients sum to 0)
template <int DIM>
__host__ __device__ __forceinline__ void fp_bary_expand(const double (&G)[DIM * (DIM + 1) / 2], double (&M)[DIM + 1][DIM + 1])
{
#pragma unroll
    for (int a = 0; a <= DIM; ++a)
#pragma unroll
        for (int b = a + 1; b <= DIM; ++b) M[a][b] = M[b][a] = G[fp_pair_index<DIM>(a, b)];
#pragma unroll
    for (int a = 0; a <= DIM; ++a) {
        double s = 0.0;
#pragma unroll
        for (int b = 0; b <= DIM; ++b)
            if (b != a) s -= M[a][b];
        M[a][a] = s;
    }
}

// ---- phase 0: per-cell geometry record -------------------------------------------------------------
//   FP_GEO_METRIC  f |T| J^-1 J^-T (upper triangle)      table Laplace
//   FP_GEO_VOLUME  f |T|                                 mass
//   FP_GEO_BARY    s f |T| grad(l_a).grad(l_b), a<b      closed-form Laplace (s = fp_bary_scale)
template <int DIM, int GEO, bool SOA>
__global__ void __launch_bounds__(256)
fp_geo_kernel(long long ncells, const double *__restrict__ coords, const int *__restrict__ cellnodes,
              const int *__restrict__ regions, const double *__restrict__ vol, double factor, int nregions,
              const int *__restrict__ visit /* device copy of regions list */, double *__restrict__ geo, const GeoLayout Lg)
{
    // SOA with Lg.permuted: cellnodes / regions / vol are copies in the transposed cell order (fastplan: tp_permute_*), the
    // thread index IS the transposed index: every load and store of the kernel is coalesced
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= (SOA && Lg.permuted ? Lg.Npad : ncells)) return;
    if (SOA && Lg.permuted && cellnodes[c * (DIM + 1)] < 0) return;   // padding slot of the transposed order
    constexpr int NG = fp_ng(DIM, GEO);
    double f = factor * vol[c];
    if (nregions > 0) {
        int reg = regions[c], vis = 0;
        for (int k = 0; k < nregions; ++k) vis |= (visit[k] == reg);
        if (!vis) f = 0.0;
    }
    double *g = SOA ? geo + (Lg.permuted ? c : geo_perm(Lg, c)) : geo + c * NG;
    if (GEO == FP_GEO_VOLUME) { g[0] = f; return; }
    int cn[DIM + 1];
    if (DIM == 3) {
        int4 q = __ldg(reinterpret_cast<const int4 *>(cellnodes) + c);
        cn[0] = q.x; cn[1 % (DIM + 1)] = q.y; cn[2 % (DIM + 1)] = q.z; cn[3 % (DIM + 1)] = q.w;
    } else {
#pragma unroll
        for (int r = 0; r <= DIM; ++r) cn[r] = cellnodes[c * (DIM + 1) + r];
    }
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
    double out[NG];
    if (GEO == FP_GEO_METRIC) {
        int o = 0;
#pragma unroll
        for (int d = 0; d < DIM; ++d)
#pragma unroll
            for (int e = d; e < DIM; ++e) {
                double s = 0.0;
#pragma unroll
                for (int x = 0; x < DIM; ++x) s += B[d][x] * B[e][x];
                out[o++] = f * s;
            }
    } else {
        // gradients of the barycentric coordinates: l_r = row r-1 of B (r >= 1), l_0 = -sum
        double L[DIM + 1][DIM];
#pragma unroll
        for (int x = 0; x < DIM; ++x) {
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < DIM; ++r) { L[r + 1][x] = B[r][x]; s -= B[r][x]; }
            L[0][x] = s;
        }
#pragma unroll
        for (int a = 0; a <= DIM; ++a)
#pragma unroll
            for (int b = a + 1; b <= DIM; ++b) {
                double s = 0.0;
#pragma unroll
                for (int x = 0; x < DIM; ++x) s += L[a][x] * L[b][x];
                out[fp_pair_index<DIM>(a, b)] = f * s;   // f already contains fp_bary_scale (host)
            }
    }
    if (SOA) {
#pragma unroll
        for (int o = 0; o < NG; ++o) g[(size_t)o * Lg.Npad] = out[o];
    } else if (NG % 2 == 0) {
#pragma unroll
        for (int o = 0; o < NG; o += 2) reinterpret_cast<double2 *>(g)[o / 2] = make_double2(out[o], out[(o + 1) % NG]);
    } else {
#pragma unroll
        for (int o = 0; o < NG; ++o) g[o] = out[o];
    }
}

// ---- plan ------------------------------------------------------------------------------------------
struct FastPlanDev {
    const int *chunklist;         // [nchunks] chunk ids ordered by shared-memory class
    const int *slotcol;           // [nchunks*FP_T] column (block-local) of the slot, sorted by signature; -1: idle
    const int *slotoff;           // [nchunks*FP_T] offset of the slot's column segment in the chunk's accumulators
    const int *chunktot;          // [nchunks] accumulator doubles of the chunk
    const int *warpniter;         // [nchunks*FP_W]
    const long long *warpoff;     // [nchunks*FP_W] offset (in u32 words) of the warp's record block
    const unsigned *rec;          // records: [round][lane][RW] per warp
};

struct FastArgs {
    FastPlanDev plan;
    const long long *colptr;      // of the column block
    double *nzval;
    const double *geo;
    long long Npad;               // structure-of-arrays geometry: plane stride
    int overwrite;
    int chunk0;                   // first entry of chunklist of this launch (class)
};

// ---- evaluators: cur[t] += A_loc[t][KL] for local rows t in [T0, T1) ---------------------------------
template <int NS_, int NG_, int NV_>
struct EvalTable {
    static constexpr int NS = NS_, NG = NG_, NV = NV_;
    static constexpr bool BARY = false;          // no plan-time specialisation (jit.cuh)
    static constexpr int DIM_ = NV_ - 1, ORDER_ = 0;
    // rows that EVERY adjacent cell of a column contributes to can be accumulated in registers (fastplan: template kernel):
    // slot of local row t for a column with local index kl, -1: none.  Any element: the column dof itself.
    static constexpr int NCOMMON = 1;
    __host__ __device__ static constexpr int common_slot(int kl, int t) { return t == kl ? 0 : -1; }
    // geometry values (bit g) that local column kl reads: all of them for the table evaluator
    __host__ __device__ static constexpr unsigned plane_mask(int) { return (1u << NG) - 1u; }
    template <int KL, int T0, int T1>
    __device__ __forceinline__ static void column(const double (&G)[NG], double (&cur)[NS])
    {
#pragma unroll
        for (int t = T0; t < T1; ++t) {
            double s = cur[t];
#pragma unroll
            for (int g = 0; g < NG; ++g) s = fma(G[g], c_fp_S[(KL * NS + t) * NG + g], s);
            cur[t] = s;
        }
    }
};

template <int DIM, int ORDER>
struct EvalBary {
    static constexpr int NV = DIM + 1, NS = ORDER == 1 ? DIM + 1 : (DIM + 1) * (DIM + 2) / 2, NG = DIM * (DIM + 1) / 2;
    static constexpr bool BARY = true;           // closed form: can be specialised per template at plan time (jit.cuh)
    static constexpr int DIM_ = DIM, ORDER_ = ORDER;
    // geometry values (bit g = pair index of D_ab) that local column kl reads: the closed form only touches rows of the Gram
    // matrix that belong to the column dof's vertices (the vertex itself, or the two end points of its edge)
    __host__ __device__ static constexpr unsigned plane_mask(int kl)
    {
        unsigned m = 0;
        const int v0 = kl < NV ? kl : fp_edge_a<DIM>(kl - NV), v1 = kl < NV ? kl : fp_edge_b<DIM>(kl - NV);
        for (int a = 0; a <= DIM; ++a)
            for (int b = a + 1; b <= DIM; ++b)
                if (a == v0 || b == v0 || a == v1 || b == v1) m |= 1u << fp_pair_index<DIM>(a, b);
        return m;
    }
    // P2: every cell around an edge dof also holds the edge's two end points: slots 1 / 2 (local edge vertices a / b)
    static constexpr int NCOMMON = ORDER == 2 ? 3 : 1;
    __host__ __device__ static constexpr int common_slot(int kl, int t)
    {
        if (t == kl) return 0;
        if (ORDER == 2 && kl >= NV) {
            if (t == fp_edge_a<DIM>(kl - NV)) return 1;
            if (t == fp_edge_b<DIM>(kl - NV)) return 2;
        }
        return -1;
    }
    template <int KL, int T0, int T1>
    __device__ __forceinline__ static void column(const double (&G)[NG], double (&cur)[NS])
    {
        double M[DIM + 1][DIM + 1];
        fp_bary_expand<DIM>(G, M);
#pragma unroll
        for (int t = T0; t < T1; ++t) cur[t] = fp_bary_acc<DIM, ORDER>(t, KL, M, cur[t]);
    }
};

// Column KL of the local matrix, rows [T0, T1), accumulated at the byte positions of the record.  The positions
// of one (column, cell) pair are distinct rows of the column, so all loads are issued before all stores: the
// read-modify-write chains run concurrently instead of being serialised by possible aliasing.
template <class EV, int KL, int T0, int T1>
__device__ __forceinline__ void fp_column_rmw(const double (&G)[EV::NG], double *__restrict__ a, const unsigned (&w)[fp_rw(EV::NS)])
{
    if (KL < EV::NS) {
        constexpr int K = KL < EV::NS ? KL : 0;
        double cur[EV::NS];
        int pos[EV::NS];
#pragma unroll
        for (int t = T0; t < T1; ++t) {
            pos[t] = (w[1 + t / 4] >> (8 * (t % 4))) & 0xff;
            cur[t] = a[pos[t]];
        }
        EV::template column<K, T0, T1>(G, cur);
#pragma unroll
        for (int t = T0; t < T1; ++t) a[pos[t]] = cur[t];
    }
}

template <class EV, int T0, int T1>
__device__ __forceinline__ void fp_dispatch(int kl, const double (&G)[EV::NG], double *__restrict__ a, const unsigned (&w)[fp_rw(EV::NS)])
{
    static_assert(EV::NS <= 10, "fp_dispatch handles up to 10 local dofs");
    switch (kl) {
    case 0: fp_column_rmw<EV, 0, T0, T1>(G, a, w); break;
    case 1: fp_column_rmw<EV, 1, T0, T1>(G, a, w); break;
    case 2: fp_column_rmw<EV, 2, T0, T1>(G, a, w); break;
    case 3: fp_column_rmw<EV, 3, T0, T1>(G, a, w); break;
    case 4: fp_column_rmw<EV, 4, T0, T1>(G, a, w); break;
    case 5: fp_column_rmw<EV, 5, T0, T1>(G, a, w); break;
    case 6: fp_column_rmw<EV, 6, T0, T1>(G, a, w); break;
    case 7: fp_column_rmw<EV, 7, T0, T1>(G, a, w); break;
    case 8: fp_column_rmw<EV, 8, T0, T1>(G, a, w); break;
    case 9: fp_column_rmw<EV, 9, T0, T1>(G, a, w); break;
    }
}

template <int RW>
__device__ __forceinline__ void fp_load_rec(const unsigned *p, unsigned (&w)[RW])
{
    if (RW == 4) {
        uint4 q = __ldcs(reinterpret_cast<const uint4 *>(p));
        w[0] = q.x; w[1 % RW] = q.y; w[2 % RW] = q.z; w[3 % RW] = q.w;
    } else if (RW == 2) {
        uint2 q = __ldcs(reinterpret_cast<const uint2 *>(p));
        w[0] = q.x; w[1 % RW] = q.y;
    } else {
#pragma unroll
        for (int j = 0; j < RW; ++j) w[j] = __ldcs(p + j);
    }
}

template <int NG, bool SOA>
__device__ __forceinline__ void fp_load_geo(const double *__restrict__ geo, long long Npad, int cell, double (&G)[NG])
{
    if (SOA) { // `cell` is the transposed index (stored in the record)
#pragma unroll
        for (int g = 0; g < NG; ++g) G[g] = __ldg(geo + (size_t)g * Npad + cell);
        return;
    }
    const double *gp = geo + (size_t)cell * NG;
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
}

// rounds of one warp over local rows [T0, T1): software pipeline with word 0 (cell | kl) two rounds ahead -- which
// also pulls the record's line into L1 --, full record and geometry one round ahead, double-buffered by parity
template <class EV, int T0, int T1, bool SOA>
__device__ __forceinline__ void fp_rounds(const unsigned *__restrict__ rec, int niter, const double *__restrict__ geo, long long Npad,
                                          double *__restrict__ a)
{
    constexpr int NG = EV::NG, RW = fp_rw(EV::NS);
    constexpr unsigned CELLMASK = (1u << FP_CELLBITS) - 1u, NONE = 0xffffffffu;
    unsigned w[2][RW];
    double G[2][NG];
    unsigned c1 = NONE; // word 0 of round r+1
    if (niter > 0) {
        fp_load_rec<RW>(rec, w[0]);
        if (w[0][0] != NONE) fp_load_geo<NG, SOA>(geo, Npad, w[0][0] & CELLMASK, G[0]);
    }
    if (niter > 1) c1 = __ldcs(rec + (size_t)32 * RW);
#define FP_ROUND(CUR, NXT)                                                                                  \
    {                                                                                                       \
        unsigned c2 = NONE;                                                                                 \
        if (r + 1 < niter) {                                                                                \
            if (c1 != NONE) fp_load_geo<NG, SOA>(geo, Npad, c1 & CELLMASK, G[NXT]);                                    \
            fp_load_rec<RW>(rec + (size_t)(r + 1) * 32 * RW, w[NXT]);                                       \
            if (r + 2 < niter) c2 = __ldcs(rec + (size_t)(r + 2) * 32 * RW);                                \
        }                                                                                                   \
        if (w[CUR][0] != NONE) fp_dispatch<EV, T0, T1>((int)(w[CUR][0] >> FP_CELLBITS), G[CUR], a, w[CUR]); \
        c1 = c2;                                                                                            \
    }
    int r = 0;
    for (; r + 1 < niter; r += 2) {
        FP_ROUND(0, 1)
        ++r;
        FP_ROUND(1, 0)
        --r;
    }
    if (r < niter) FP_ROUND(0, 1)
#undef FP_ROUND
}

// NGRP == 2: warp group 0 accumulates the vertex rows [0, NV), group 1 the remaining rows [NV, NS) of the same
// columns; the two groups touch disjoint global rows, hence disjoint accumulators, and need no synchronisation.
template <class EV, int NGRP, int MINB, bool SOA>
__global__ void __launch_bounds__(FP_T * NGRP, MINB)
fp_gather_kernel(const __grid_constant__ FastArgs A)
{
    extern __shared__ double acc[];
    __shared__ long long s_cp[FP_T];
    __shared__ int s_len[FP_T], s_off[FP_T];
    constexpr int NS = EV::NS, RW = fp_rw(NS), NT = FP_T * NGRP;
    const int chunk = A.plan.chunklist[A.chunk0 + blockIdx.x];
    const int n = A.plan.chunktot[chunk];
    const int slot = threadIdx.x % FP_T, grp = threadIdx.x / FP_T;
    const int lane = threadIdx.x & 31, warp = slot >> 5;
    const int k = A.plan.slotcol[(size_t)chunk * FP_T + slot];
    const int off = A.plan.slotoff[(size_t)chunk * FP_T + slot];
    if (grp == 0) {
        long long c0 = 0;
        int len = 0;
        if (k >= 0) { c0 = A.colptr[k]; len = (int)(A.colptr[k + 1] - c0); }
        s_cp[slot] = c0; s_len[slot] = len; s_off[slot] = off;
    }
    if (A.overwrite) {
        for (int i = threadIdx.x; i < n; i += NT) acc[i] = 0.0;
        __syncthreads();
    } else {
        __syncthreads();
        for (int s = threadIdx.x >> 5; s < FP_T; s += NT / 32) {
            const long long c0 = s_cp[s];
            const int len = s_len[s], o = s_off[s];
            for (int i = lane; i < len; i += 32) acc[o + i] = A.nzval[c0 + i];
        }
        __syncthreads();
    }
    const int niter = A.plan.warpniter[chunk * FP_W + warp];
    const unsigned *rec = A.plan.rec + A.plan.warpoff[chunk * FP_W + warp] + lane * RW;
    double *a = acc + off;
    if (NGRP == 1) fp_rounds<EV, 0, NS, SOA>(rec, niter, A.geo, A.Npad, a);
    else if (grp == 0) fp_rounds<EV, 0, EV::NV, SOA>(rec, niter, A.geo, A.Npad, a);
    else fp_rounds<EV, EV::NV, NS, SOA>(rec, niter, A.geo, A.Npad, a);
    __syncthreads();
    for (int s = threadIdx.x >> 5; s < FP_T; s += NT / 32) {
        const long long c0 = s_cp[s];
        const int len = s_len[s], o = s_off[s];
        for (int i = lane; i < len; i += 32) __stcs(A.nzval + c0 + i, acc[o + i]);
    }
}

// ---- plan construction (setup, once per pattern) -----------------------------------------------
// sort key: window | number of adjacent cells | hash of the local indices; the radix sort is stable, so equal
// signatures stay in column order
// (columns = the list `cols` of block-local column ids, ascending; the window is taken over the list index)
__global__ void fp_key_kernel(long long ncols, const int *__restrict__ cols, const long long *__restrict__ adjptr,
                              const unsigned char *__restrict__ adjloc, unsigned long long *__restrict__ key, int *__restrict__ col)
{
    long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= ncols) return;
    const int kc = cols[k];
    long long p0 = adjptr[kc], p1 = adjptr[kc + 1];
    unsigned long long h = 1469598103934665603ull;
    for (long long p = p0; p < p1; ++p) { h ^= adjloc[p]; h *= 1099511628211ull; }
    unsigned long long cnt = (unsigned long long)min((long long)255, p1 - p0);
    key[k] = ((unsigned long long)(k / FP_WINDOW) << 42) | (cnt << 34) | ((h ^ (h >> 34)) & ((1ull << 34) - 1));
    col[k] = kc;
}

// one CTA per chunk of FP_T sorted slots: slot -> column, accumulator offsets, chunk total, rounds per warp
__global__ void __launch_bounds__(FP_T)
fp_chunk_kernel(long long ncols, const int *__restrict__ order, const long long *__restrict__ colptr, const long long *__restrict__ adjptr,
                int *__restrict__ slotcol, int *__restrict__ slotoff, int *__restrict__ chunktot, int *__restrict__ warpniter)
{
    __shared__ int sc[FP_T];
    const int t = threadIdx.x;
    const long long slot = (long long)blockIdx.x * FP_T + t;
    const int c = slot < ncols ? order[slot] : -1;
    int len = 0, cnt = 0;
    if (c >= 0) { len = (int)(colptr[c + 1] - colptr[c]); cnt = (int)(adjptr[c + 1] - adjptr[c]); }
    sc[t] = len;
    __syncthreads();
    for (int o = 1; o < FP_T; o <<= 1) { // inclusive Hillis-Steele scan
        int v = t >= o ? sc[t - o] : 0;
        __syncthreads();
        sc[t] += v;
        __syncthreads();
    }
    slotcol[slot] = c;
    slotoff[slot] = sc[t] - len;
    if (t == FP_T - 1) chunktot[blockIdx.x] = sc[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt = max(cnt, __shfl_xor_sync(0xffffffffu, cnt, o));
    if ((t & 31) == 0) warpniter[blockIdx.x * FP_W + (t >> 5)] = cnt;
}

template <typename PosT>
__global__ void fp_fill_kernel(long long nslots, int ns, int rw, int posstride, const int *__restrict__ slotcol,
                               const int *__restrict__ warpniter, const long long *__restrict__ warpoff,
                               const long long *__restrict__ adjptr, const int *__restrict__ adjcell,
                               const unsigned char *__restrict__ adjloc, const PosT *__restrict__ posmap, unsigned *__restrict__ rec,
                               const GeoLayout Lg)
{
    long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslots) return;
    long long wg = s >> 5;
    int lane = (int)(s & 31);
    int niter = warpniter[wg];
    unsigned *out = rec + warpoff[wg] + lane * rw;
    int k = slotcol[s];
    long long p0 = 0, p1 = 0;
    if (k >= 0) { p0 = adjptr[k]; p1 = adjptr[k + 1]; }
    for (int r = 0; r < niter; ++r) {
        unsigned w[4] = {0xffffffffu, 0, 0, 0};
        long long p = p0 + r;
        if (p < p1) {
            w[0] = (unsigned)(Lg.soa ? geo_perm(Lg, adjcell[p]) : (long long)adjcell[p]) | ((unsigned)adjloc[p] << FP_CELLBITS);
            for (int t = 0; t < ns; ++t) w[1 + t / 4] |= ((unsigned)posmap[p * posstride + t] & 0xff) << (8 * (t % 4));
        }
        for (int j = 0; j < rw; ++j) out[(size_t)r * 32 * rw + j] = w[j];
    }
}

} // namespace extfem
