// dist.cuh -- multi-GPU plumbing of the sharded system.
//
// One process per GPU; cells are partitioned into contiguous ranges and every rank assembles ITS cells into a
// local system over its local dofs (the single-GPU kernels, unchanged).  The global system is the sum of the
// local ones, A = sum_g R_g^T A_g R_g, b = sum_g R_g^T b_g; dofs on a partition interface exist on several
// ranks.  The only data that crosses NVLink are INTERFACE-ROW contributions:
//   * of the right-hand side once per assembly              (extfem_dist_sum_rhs),
//   * of the matrix diagonal once per solve (Jacobi)        (inside extfem_dist_cg),
//   * of y = A x once per SpMV                              (extfem_dist_spmv / extfem_dist_cg),
// packed by a gather kernel, exchanged with grouped ncclSend/ncclRecv between neighbouring ranks on the
// context's stream, and added by a scatter kernel; plus one ncclAllReduce of a scalar per dot product.
// NCCL is resolved at run time (dlopen of libnccl.so.2: the copy already loaded by torch if there is one).
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"
#include "solver.cuh"

namespace extfem {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string load()
    {
        if (handle) return "";
        handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!handle) handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!handle) return std::string("cannot load libnccl.so.2: ") + dlerror();
#define EXTFEM_NCCL_SYM(name)                                                       \
        *(void **)(&name) = dlsym(handle, "nccl" #name);                            \
        if (!name) return std::string("libnccl.so.2 lacks nccl" #name);
        EXTFEM_NCCL_SYM(GetUniqueId) EXTFEM_NCCL_SYM(CommInitRank) EXTFEM_NCCL_SYM(CommDestroy) EXTFEM_NCCL_SYM(GroupStart)
        EXTFEM_NCCL_SYM(GroupEnd) EXTFEM_NCCL_SYM(Send) EXTFEM_NCCL_SYM(Recv) EXTFEM_NCCL_SYM(AllReduce) EXTFEM_NCCL_SYM(GetErrorString)
#undef EXTFEM_NCCL_SYM
        return "";
    }
};

static NcclApi g_nccl;

struct DistState {
    bool ready = false;
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr;
};

// interface rows of one pattern shared with the neighbouring ranks
struct IfacePlan {
    bool ready = false;
    std::vector<int> ranks;           // neighbour ranks
    std::vector<long long> ptr;       // [nneigh + 1] into rows / buffers
    void *rows = nullptr;             // int, 0-based local rows, per neighbour in the order both sides agree on
    void *sendbuf = nullptr, *recvbuf = nullptr;
    void *weight = nullptr;           // double [nrows]: 1 for rows this rank owns, 0 otherwise
    ~IfacePlan()
    {
        for (void *p : {rows, sendbuf, recvbuf, weight}) if (p) cudaFree(p);
    }
};

__global__ void iface_pack_kernel(long long n, const int *__restrict__ rows, const double *__restrict__ v, double *__restrict__ buf)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) buf[i] = v[rows[i]];
}

// rows of one neighbour are distinct, so one launch per neighbour adds without atomics (a row shared with several
// neighbours appears once in each neighbour's list)
__global__ void iface_add_kernel(long long n, const int *__restrict__ rows, const double *__restrict__ buf, double *__restrict__ v)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[rows[i]] += buf[i];
}

__global__ void __launch_bounds__(256) dotw_partial_kernel(long long n, const double *__restrict__ a, const double *__restrict__ b,
                                                           const double *__restrict__ w, double *__restrict__ partial)
{
    __shared__ double sh[256];
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s += w[i] * a[i] * b[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

// after the all-reduce of scalars[slot]: mode 1: alpha = rz / pq; mode 2: beta = rz_new / rz, rz = rz_new
__global__ void cg_scalar_kernel(double *scalars, int slot, int mode)
{
    const double v = scalars[slot];
    if (mode == 1) scalars[3] = scalars[0] / v;
    if (mode == 2) { scalars[4] = v / scalars[0]; scalars[0] = v; }
}

__global__ void diag_kernel(long long n, const long long *__restrict__ colptr, const int *__restrict__ rowval,
                            const double *__restrict__ nzval, double *__restrict__ diag)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long p = find_in_column(colptr, rowval, i, (int)i);
    diag[i] = p >= 0 ? nzval[p] : 0.0;
}

__global__ void invert_kernel(long long n, double *__restrict__ d)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = d[i] != 0.0 ? 1.0 / d[i] : 1.0;
}

// v_i <- sum over the ranks sharing row i of their v_i: pack, grouped send/recv with every neighbour, add
static inline int iface_exchange_add(cudaStream_t st, DistState &D, IfacePlan &I, double *v, long long *launches, std::string *err)
{
    const long long ntot = I.ptr.empty() ? 0 : I.ptr.back();
    if (ntot == 0 || D.world == 1) return 0;
    iface_pack_kernel<<<(unsigned)((ntot + 255) / 256), 256, 0, st>>>(ntot, (const int *)I.rows, v, (double *)I.sendbuf);
    if (launches) ++*launches;
    ncclResult_t r = g_nccl.GroupStart();
    for (size_t k = 0; r == ncclSuccess && k < I.ranks.size(); ++k) {
        const long long o = I.ptr[k], c = I.ptr[k + 1] - I.ptr[k];
        if (c == 0) continue;
        r = g_nccl.Send((const double *)I.sendbuf + o, (size_t)c, ncclDouble, I.ranks[k], D.comm, st);
        if (r == ncclSuccess) r = g_nccl.Recv((double *)I.recvbuf + o, (size_t)c, ncclDouble, I.ranks[k], D.comm, st);
    }
    ncclResult_t r2 = g_nccl.GroupEnd();
    if (r == ncclSuccess) r = r2;
    if (r != ncclSuccess) { if (err) *err = std::string("NCCL send/recv: ") + g_nccl.GetErrorString(r); return -1; }
    for (size_t k = 0; k < I.ranks.size(); ++k) {
        const long long o = I.ptr[k], c = I.ptr[k + 1] - I.ptr[k];
        if (c == 0) continue;
        iface_add_kernel<<<(unsigned)((c + 255) / 256), 256, 0, st>>>(c, (const int *)I.rows + o, (const double *)I.recvbuf + o, v);
        if (launches) ++*launches;
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// Jacobi-preconditioned CG on the additive sharded system.  All vectors are CONSISTENT (interface rows hold the same
// value on every sharing rank); dot products weight every row by its ownership and are all-reduced.
static inline int dist_jacobi_cg(cudaStream_t st, DistState &D, IfacePlan &I, long long n, const long long *colptr, const int *rowval,
                                 const double *nzval, const double *b, double *x, double rtol, int maxit, int *iters, double *relres,
                                 SolverWork &W, long long *launches, std::string *err)
{
    if (W.n_alloc < n) {
        W.n_alloc = 0;
        for (auto &v : W.vec) { if (v) { cudaFree(v); v = nullptr; } if (cudaMalloc(&v, n * 8)) { v = nullptr; return -1; } }
        if (!W.partial && cudaMalloc(&W.partial, RED_BLOCKS * 8)) return -1;
        if (!W.scalars && cudaMalloc(&W.scalars, 8 * 8)) return -1;
        W.n_alloc = n;
    }
    double *r = (double *)W.vec[0], *z = (double *)W.vec[1], *p = (double *)W.vec[2], *q = (double *)W.vec[3], *dinv = (double *)W.vec[4];
    double *partial = (double *)W.partial, *sc = (double *)W.scalars;
    const double *wgt = (const double *)I.weight;
    unsigned gb = (unsigned)((n + 255) / 256), gs = (unsigned)((n * SPMV_LANES + 255) / 256);
    long long nl = 0;
    int fail = 0;
    auto dot = [&](const double *a, const double *c, int slot, int mode) {
        dotw_partial_kernel<<<RED_BLOCKS, 256, 0, st>>>(n, a, c, wgt, partial);
        dot_final_kernel<<<1, 256, 0, st>>>(partial, RED_BLOCKS, sc, slot, 0);
        if (D.world > 1 && g_nccl.AllReduce(sc + slot, sc + slot, 1, ncclDouble, ncclSum, D.comm, st) != ncclSuccess) fail = 1;
        if (mode) cg_scalar_kernel<<<1, 1, 0, st>>>(sc, slot, mode);
        nl += 3;
    };
    auto spmv = [&](const double *in, double *out) -> int {
        spmv_kernel<<<gs, 256, 0, st>>>(n, colptr, rowval, nullptr, nzval, in, out);
        ++nl;
        return iface_exchange_add(st, D, I, out, &nl, err);
    };
    diag_kernel<<<gb, 256, 0, st>>>(n, colptr, rowval, nzval, dinv);
    if (iface_exchange_add(st, D, I, dinv, &nl, err)) return -5;
    invert_kernel<<<gb, 256, 0, st>>>(n, dinv);
    if (spmv(x, q)) return -5;
    cg_init_kernel<<<gb, 256, 0, st>>>(n, b, q, dinv, r, z, p);
    nl += 3;
    dot(b, b, 6, 0);
    dot(r, z, 0, 0);
    dot(r, r, 5, 0);
    double h[8];
    if (fail || cudaMemcpyAsync(h, sc, 64, cudaMemcpyDeviceToHost, st) || cudaStreamSynchronize(st)) return -2;
    // the residual is measured against the INITIAL residual |b - A x0| (== |b| for x0 = 0): with penalised Dirichlet rows
    // (1e30 * value) in b, |b| would hide the interior residual; callers pass x0 with the boundary values set
    double bnorm = std::sqrt(h[5]);
    if (bnorm == 0.0) bnorm = 1.0;
    double res = std::sqrt(h[5]) / bnorm;
    int it = 0;
    while (res > rtol && it < maxit) {
        if (spmv(p, q)) return -5;
        dot(p, q, 1, 1);
        cg_update_xr_kernel<<<gb, 256, 0, st>>>(n, sc, p, q, dinv, x, r, z);
        dot(r, z, 2, 2);
        dot(r, r, 5, 0);
        cg_update_p_kernel<<<gb, 256, 0, st>>>(n, sc, z, p);
        nl += 2;
        ++it;
        if (fail || cudaMemcpyAsync(h, sc, 64, cudaMemcpyDeviceToHost, st) || cudaStreamSynchronize(st)) return -2;
        if (!(h[1] > 0.0) || !std::isfinite(h[5])) { *iters = it; *relres = std::sqrt(h[5]) / bnorm; if (launches) *launches += nl; return -4; }
        res = std::sqrt(h[5]) / bnorm;
    }
    *iters = it;
    *relres = res;
    if (launches) *launches += nl;
    return 0;
}


// ---- owned-row form ------------------------------------------------------------------------------------------------------
// Every dof has one owner rank; a rank's local mesh carries one layer of zero-volume ghost cells, so the CSC column of an
// interface dof has the same rows on every rank that holds it (host/dist.py: OwnedShard).  After the local assembly the
// non-owners SEND their column segments / vector rows of shared dofs to the owner, which adds them: the owner then holds the
// complete column (= row, the systems handled by the CG are symmetric).  Vectors are refreshed on ghost rows by a halo
// exchange (owner -> copies) before every SpMV.  Only interface-row data crosses NVLink.
struct OwnedExchange {
    std::vector<long long> send_ptr, recv_ptr;   // [nneigh + 1] rows per neighbour
    void *send_rows = nullptr, *recv_rows = nullptr;   // int, 0-based local rows
    long long nsend = 0, nrecv = 0;
    ~OwnedExchange() { for (void *p : {send_rows, recv_rows}) if (p) cudaFree(p); }
};

struct OwnedPlanDev {
    bool ready = false;
    std::vector<int> ranks;
    OwnedExchange red, halo;
    // matrix reduction: value offsets of the listed columns inside the packed buffers
    void *send_seg = nullptr, *recv_seg = nullptr;     // long long [n + 1] prefix sums of the column lengths
    std::vector<long long> send_vptr, recv_vptr;       // [nneigh + 1] value offsets per neighbour
    void *sendbuf = nullptr, *recvbuf = nullptr;       // doubles: max(matrix values, vector rows)
    size_t bufdoubles = 0;
    void *weight = nullptr;                            // double [nrows]: 1 owned, 0 ghost
    ~OwnedPlanDev() { for (void *p : {send_seg, recv_seg, sendbuf, recvbuf, weight}) if (p) cudaFree(p); }
};

__global__ void seg_pack_kernel(long long ncolsl, const int *__restrict__ cols, const long long *__restrict__ seg,
                                const long long *__restrict__ colptr, const double *__restrict__ nzval, double *__restrict__ buf)
{
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= ncolsl) return;
    const long long c0 = colptr[cols[w]], n = seg[w + 1] - seg[w], o = seg[w];
    for (long long i = lane; i < n; i += 32) buf[o + i] = nzval[c0 + i];
}

__global__ void seg_add_kernel(long long ncolsl, const int *__restrict__ cols, const long long *__restrict__ seg,
                               const long long *__restrict__ colptr, const double *__restrict__ buf, double *__restrict__ nzval)
{
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= ncolsl) return;
    const long long c0 = colptr[cols[w]], n = seg[w + 1] - seg[w], o = seg[w];
    for (long long i = lane; i < n; i += 32) nzval[c0 + i] += buf[o + i];
}

__global__ void iface_assign_kernel(long long n, const int *__restrict__ rows, const double *__restrict__ buf, double *__restrict__ v)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[rows[i]] = buf[i];
}

// grouped send / recv of per-neighbour slices (counts in doubles)
static inline int owned_sendrecv(cudaStream_t st, DistState &D, const std::vector<int> &ranks, const std::vector<long long> &sp,
                                 const std::vector<long long> &rp, const double *sendbuf, double *recvbuf, std::string *err)
{
    ncclResult_t r = g_nccl.GroupStart();
    for (size_t k = 0; r == ncclSuccess && k < ranks.size(); ++k) {
        const long long cs = sp[k + 1] - sp[k], cr = rp[k + 1] - rp[k];
        if (cs > 0) r = g_nccl.Send(sendbuf + sp[k], (size_t)cs, ncclDouble, ranks[k], D.comm, st);
        if (r == ncclSuccess && cr > 0) r = g_nccl.Recv(recvbuf + rp[k], (size_t)cr, ncclDouble, ranks[k], D.comm, st);
    }
    ncclResult_t r2 = g_nccl.GroupEnd();
    if (r == ncclSuccess) r = r2;
    if (r != ncclSuccess) { if (err) *err = std::string("NCCL send/recv: ") + g_nccl.GetErrorString(r); return -1; }
    return 0;
}

// vector rows: non-owners -> owner (add)
static inline int owned_reduce_vector(cudaStream_t st, DistState &D, OwnedPlanDev &O, double *v, long long *launches, std::string *err)
{
    if (D.world == 1 || (O.red.nsend == 0 && O.red.nrecv == 0)) return 0;
    if (O.red.nsend) iface_pack_kernel<<<(unsigned)((O.red.nsend + 255) / 256), 256, 0, st>>>(O.red.nsend, (const int *)O.red.send_rows, v, (double *)O.sendbuf);
    if (owned_sendrecv(st, D, O.ranks, O.red.send_ptr, O.red.recv_ptr, (const double *)O.sendbuf, (double *)O.recvbuf, err)) return -1;
    if (O.red.nrecv) {
        // a row can be listed for several neighbours: one launch per neighbour keeps the adds race-free
        for (size_t k = 0; k < O.ranks.size(); ++k) {
            const long long o = O.red.recv_ptr[k], c = O.red.recv_ptr[k + 1] - o;
            if (c > 0) iface_add_kernel<<<(unsigned)((c + 255) / 256), 256, 0, st>>>(c, (const int *)O.red.recv_rows + o, (const double *)O.recvbuf + o, v);
        }
    }
    if (launches) *launches += 2;
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// matrix columns: non-owners -> owner (add), whole column segments
static inline int owned_reduce_matrix(cudaStream_t st, DistState &D, OwnedPlanDev &O, const long long *colptr, double *nzval,
                                      long long *launches, std::string *err)
{
    if (D.world == 1 || (O.red.nsend == 0 && O.red.nrecv == 0)) return 0;
    if (O.red.nsend)
        seg_pack_kernel<<<(unsigned)((O.red.nsend * 32 + 255) / 256), 256, 0, st>>>(O.red.nsend, (const int *)O.red.send_rows, (const long long *)O.send_seg,
                                                                                   colptr, nzval, (double *)O.sendbuf);
    if (owned_sendrecv(st, D, O.ranks, O.send_vptr, O.recv_vptr, (const double *)O.sendbuf, (double *)O.recvbuf, err)) return -1;
    for (size_t k = 0; k < O.ranks.size(); ++k) {
        const long long o = O.red.recv_ptr[k], c = O.red.recv_ptr[k + 1] - o;
        if (c > 0)
            seg_add_kernel<<<(unsigned)((c * 32 + 255) / 256), 256, 0, st>>>(c, (const int *)O.red.recv_rows + o, (const long long *)O.recv_seg + o, colptr,
                                                                           (const double *)O.recvbuf, nzval);
    }
    if (launches) *launches += 2;
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// ghost rows <- owner's value
static inline int owned_halo(cudaStream_t st, DistState &D, OwnedPlanDev &O, double *v, long long *launches, std::string *err)
{
    if (D.world == 1 || (O.halo.nsend == 0 && O.halo.nrecv == 0)) return 0;
    if (O.halo.nsend) iface_pack_kernel<<<(unsigned)((O.halo.nsend + 255) / 256), 256, 0, st>>>(O.halo.nsend, (const int *)O.halo.send_rows, v, (double *)O.sendbuf);
    if (owned_sendrecv(st, D, O.ranks, O.halo.send_ptr, O.halo.recv_ptr, (const double *)O.sendbuf, (double *)O.recvbuf, err)) return -1;
    if (O.halo.nrecv) iface_assign_kernel<<<(unsigned)((O.halo.nrecv + 255) / 256), 256, 0, st>>>(O.halo.nrecv, (const int *)O.halo.recv_rows, (const double *)O.recvbuf, v);
    if (launches) *launches += 2;
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

__global__ void mask_rows_kernel(long long n, const double *__restrict__ w, double *__restrict__ v)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && w[i] == 0.0) v[i] = 0.0;
}

// Jacobi-preconditioned CG on the owned-row system (symmetric positive definite): vectors are meaningful on owned rows,
// ghost rows of the search direction are refreshed by the halo exchange before every SpMV, dot products run over owned rows.
static inline int owned_jacobi_cg(cudaStream_t st, DistState &D, OwnedPlanDev &O, long long n, const long long *colptr, const int *rowval,
                                  const double *nzval, const double *b, double *x, double rtol, int maxit, int *iters, double *relres,
                                  SolverWork &W, long long *launches, std::string *err)
{
    if (W.n_alloc < n) {
        W.n_alloc = 0;
        for (auto &v : W.vec) { if (v) { cudaFree(v); v = nullptr; } if (cudaMalloc(&v, n * 8)) { v = nullptr; return -1; } }
        if (!W.partial && cudaMalloc(&W.partial, RED_BLOCKS * 8)) return -1;
        if (!W.scalars && cudaMalloc(&W.scalars, 8 * 8)) return -1;
        W.n_alloc = n;
    }
    double *r = (double *)W.vec[0], *z = (double *)W.vec[1], *p = (double *)W.vec[2], *q = (double *)W.vec[3], *dinv = (double *)W.vec[4];
    double *partial = (double *)W.partial, *sc = (double *)W.scalars;
    const double *wgt = (const double *)O.weight;
    unsigned gb = (unsigned)((n + 255) / 256), gs = (unsigned)((n * SPMV_LANES + 255) / 256);
    long long nl = 0;
    int fail = 0;
    auto dot = [&](const double *a, const double *c, int slot, int mode) {
        dotw_partial_kernel<<<RED_BLOCKS, 256, 0, st>>>(n, a, c, wgt, partial);
        dot_final_kernel<<<1, 256, 0, st>>>(partial, RED_BLOCKS, sc, slot, 0);
        if (D.world > 1 && g_nccl.AllReduce(sc + slot, sc + slot, 1, ncclDouble, ncclSum, D.comm, st) != ncclSuccess) fail = 1;
        if (mode) cg_scalar_kernel<<<1, 1, 0, st>>>(sc, slot, mode);
        nl += 3;
    };
    auto spmv = [&](double *in, double *out) -> int {
        if (owned_halo(st, D, O, in, &nl, err)) return -5;
        spmv_kernel<<<gs, 256, 0, st>>>(n, colptr, rowval, nullptr, nzval, in, out);
        ++nl;
        return 0;
    };
    diag_kernel<<<gb, 256, 0, st>>>(n, colptr, rowval, nzval, dinv);
    invert_kernel<<<gb, 256, 0, st>>>(n, dinv);
    if (spmv(x, q)) return -5;
    cg_init_kernel<<<gb, 256, 0, st>>>(n, b, q, dinv, r, z, p);
    nl += 3;
    dot(b, b, 6, 0);
    dot(r, z, 0, 0);
    dot(r, r, 5, 0);
    double h[8];
    if (fail || cudaMemcpyAsync(h, sc, 64, cudaMemcpyDeviceToHost, st) || cudaStreamSynchronize(st)) return -2;
    // the residual is measured against the INITIAL residual |b - A x0| (== |b| for x0 = 0): with penalised Dirichlet rows
    // (1e30 * value) in b, |b| would hide the interior residual; callers pass x0 with the boundary values set
    double bnorm = std::sqrt(h[5]);
    if (bnorm == 0.0) bnorm = 1.0;
    double res = std::sqrt(h[5]) / bnorm;
    int it = 0;
    while (res > rtol && it < maxit) {
        if (spmv(p, q)) return -5;
        dot(p, q, 1, 1);
        cg_update_xr_kernel<<<gb, 256, 0, st>>>(n, sc, p, q, dinv, x, r, z);
        dot(r, z, 2, 2);
        dot(r, r, 5, 0);
        cg_update_p_kernel<<<gb, 256, 0, st>>>(n, sc, z, p);
        nl += 2;
        ++it;
        if (fail || cudaMemcpyAsync(h, sc, 64, cudaMemcpyDeviceToHost, st) || cudaStreamSynchronize(st)) return -2;
        if (!(h[1] > 0.0) || !std::isfinite(h[5])) { *iters = it; *relres = std::sqrt(h[5]) / bnorm; if (launches) *launches += nl; return -4; }
        res = std::sqrt(h[5]) / bnorm;
    }
    if (owned_halo(st, D, O, x, &nl, err)) return -5;   // the solution is returned consistent on every local dof
    *iters = it;
    *relres = res;
    if (launches) *launches += nl;
    return 0;
}

} // namespace extfem
