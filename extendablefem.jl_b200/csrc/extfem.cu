// extfem.cu -- C-ABI implementation of libextfem_cuda.so (see include/extfem_cuda.h).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <set>
#include <tuple>

#include "common.cuh"
#include "fe_tables.h"
#include "gather.cuh"
#include "kernels_generic.cuh"
#include "pattern.cuh"
#include "fastpath.cuh"
#include "fastplan.cuh"
#include "jit.cuh"
#include "solver.cuh"
#include "dist.cuh"
#include "entities.cuh"

namespace extfem {

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct Mesh {
    int dim = 0;                     // space dimension
    int tdim = 0;                    // topological dimension of the items (== dim for cells, dim - 1 for boundary faces)
    long long ncells = 0, nnodes = 0;
    DevBuf coords, cellnodes, regions, vol;
    long long vol_version = 0;       // bumped when the cell volumes change
    Mesh *parent = nullptr;          // boundary-face meshes borrow the coordinates of their grid
    std::unique_ptr<Mesh> bf;        // xgrid[BFaceNodes / BFaceRegions / BFaceVolumes] (extfem_mesh_set_bfaces)
    bool bf_vol_given = false;
    double *coordsp() const { return parent ? parent->coords.as<double>() : coords.as<double>(); }
};

struct Space {
    int id = -1;                     // index in Ctx::spaces
    int tabkey = -1;                 // table cache key of a host-supplied basis (2 * id + face), -1: built-in
    int mesh = -1, fetype = 0, order = 0, ncomp = 0, nscalar = 0, nd = 0;
    long long ndofs = 0;
    DevBuf celldofs, adjptr, adjcell, adjloc;
    // host-supplied polynomial reference basis (EXTFEM_FE_TABULATED, extfem_space_set_tables)
    bool poly_ready = false;
    std::vector<double> poly, bpoly; // [nscalar][nmono(order, tdim)], boundary-face restriction [nscalar_bf][nmono(order, tdim-1)]
    int nscalar_bf = 0;
    std::unique_ptr<Space> bf;       // FES[BFaceDofs] (extfem_space_set_bfacedofs): item dof map over the boundary faces
};

struct FastPlan {
    bool ready = false, usable = false;
    int nchunks = 0;
    int cls_start[FP_NCLASS + 1] = {};  // chunks are ordered by shared-memory size class
    std::vector<int> cls_maxtot = std::vector<int>(FP_NCLASS, 0);
    DevBuf chunklist, slotcol, slotoff, chunktot, warpniter, warpoff, rec;
};

// template plan of one diagonal block (fastplan.cuh)
struct TemplatePlan {
    bool ready = false;
    GeoLayout Lg{0, 0, 1, 0, 0};
    int nwarps = 0, nctas = 0, ngroups = 0, ntemplates = 0, pool_bytes = 0;
    long long nrounds = 0, nleft = 0, ncols = 0;
    DevBuf wdesc, slotcol, slotpb, slotptr, tmpl, leftcols, dump;
    DevBuf cn_p, reg_p, vol_p;       // mesh arrays in the transposed cell order (Lg.permuted)
    long long vol_version = -1;
    std::vector<JitTemplate> jtmpl;  // host copy of the templates (plan-time specialisation, jit.cuh)
    std::unique_ptr<JitModule> jit;
    DevBuf walk_live[2], walk_next;  // persistent walk kernel: launch-order warps of the short / long class, queue heads
    int walk_nlive[2] = {0, 0}, walk_doubles[2] = {0, 0}, walk_nl = 1, walk_ns = 0, walk_ctas = 1;
    DevBuf jit_live[2];              // launch-order warps of the short- / long-column class
    int jit_nlive[2] = {0, 0};
    DevBuf walk;                     // walk records of the templates (walkplan.h), [nrounds][TW_RW] words, same round indices
    bool walk_ok = false;
};

struct Pattern {
    std::vector<std::unique_ptr<FastPlan>> fastplans; // per column block (diagonal blocks only)
    std::vector<std::unique_ptr<TemplatePlan>> tplans;
    std::vector<long long> hcolptr;                   // host copy of colptr (plan construction)
    std::vector<int> rowspaces, colspaces;
    std::vector<long long> rowoff, coloff;     // size n+1
    std::vector<int> rowlocoff;                // per row block: offset in the posmap row
    std::vector<unsigned char> coupling;       // [c*nrow + r]
    int NRpat = 0;
    long long nrows = 0, ncols = 0, nnz = 0;
    int poswidth = 1;                          // bytes per posmap entry
    int maxcollen = 0;
    DevBuf colptr, rowval, nzval, b;
    DevBuf lstart, lcolptr, lpack;             // lower-triangle export (extfem_values_get_lower): suffix starts, packed colptr, staging
    long long nnz_lower = -1;
    std::vector<long long> lchunk_col, lchunk_off;   // column chunks of the pipelined export and the packed offsets at their ends
    std::vector<std::unique_ptr<DevBuf>> posmap; // per column block
    DevBuf chunkptr;
    int nchunks = 0;
    bool square = false;
    SolverWork cg;
    IfacePlan iface;                           // partition interfaces (multi-GPU, additive form)
    OwnedPlanDev owned;                        // exchange lists of the owned-row form
};

struct TableKey {
    int dim, order, quadorder, space;   // space >= 0: host-supplied basis of that space (2 * id + face)
    bool operator<(const TableKey &o) const { return std::tie(dim, order, quadorder, space) < std::tie(o.dim, o.order, o.quadorder, o.space); }
};
struct DevTables {
    int nq = 0;
    DevBuf vals, grads;
};
struct DevQuad {
    int nq = 0;
    DevBuf w, x;
    std::vector<double> hw, hx;
};

struct Ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    DistState dist;
    std::string err;
    std::vector<std::unique_ptr<Mesh>> meshes;
    std::vector<std::unique_ptr<Space>> spaces;
    std::vector<std::unique_ptr<Pattern>> patterns;
    std::map<TableKey, std::unique_ptr<DevTables>> tables;
    std::map<std::pair<int, int>, std::unique_ptr<DevQuad>> quads; // (dim, order)
    DevBuf custom_qw, custom_qx;
    DevBuf loc, bloc, sol, params_scratch, tab, geo, visit, fq;
    bool fast_enabled = true;   // option "fastpath"
    int nl_cpw = 8, nl_warps = 4; // options "nonlinear_cells_per_warp", "nonlinear_warps_per_cta" of the tensor-core contraction kernel
    bool nl_sparse_jac = true;  // option "nonlinear_sparse_jacobian": the tensor-core nonlinear path moves only the structural non-zeros of J
    bool nl_point_cache = true; // option "nonlinear_point_cache": nl_point_kernel caches the physical basis values of its point in shared memory
    bool nl_rowwise = true;     // option "nonlinear_rowwise": Neo-Hooke point kernel with the Jacobian produced row by row (no spills)
    bool gather_warp = true;    // option "gather_warp": generic matrix reduction with one warp (1) / one thread (0) per column
    int nl_version = 4;         // option "nonlinear_kernel": 1 entry-wise local kernel, 2 staged per block, 3 warp per cell,
                                // 4 warp per cell with the contractions on FP64 tensor cores (falls back to 3 when not applicable)
    DevBuf nl2buf;
    DevBuf nlpt;                // w J and (J u - F) per (cell, quadrature point) of the tensor-core nonlinear path
    bool tmpl_enabled = true;   // option "fastpath_templates": 0 keeps every column on the record kernel
    int tmpl_ahead = 1024;      // option "template_prefetch_ctas": CTAs ahead whose start-up data is prefetched into L2
    int sm_count = 148;
    int tmpl_walk_roles = 0;       // option "template_walk_roles": 10 nl + ns forces the long / short warps of a persistent CTA (0: modelled)
    int tmpl_persistent = 0;       // option "template_persistent": N > 0 runs the walk kernel as persistent CTAs (at most N per SM) fed from
                                   // two queues; measured slower than one CTA per 4 groups in the release build (profiles/r02_tuning_log.md)
    int tmpl_pool = TP_POOL_BYTES; // option "template_pool_bytes": shared-memory pool of one template CTA
    bool jit_enabled = false;   // option "template_jit": plan-time specialisation of the templates (jit.cuh); off by default
    long long jit_mincols = 200000; // option "template_jit_min_cols": smallest column block that is worth the compile time
    bool tmpl_planemask = true; // option "template_plane_mask": rounds load only the geometry values their local column reads
    bool tmpl_const = true;     // option "template_constant_memory": template rounds in constant memory when they fit
    bool tmpl_walk = true;      // option "template_walk": 3D P2 columns run as walk programs (walkplan.h) when every template verifies
    bool tmpl_permute_mesh = true; // option "template_permute_mesh": cell kernels read mesh copies in the transposed order
    int tmpl_mincols = 24;      // option "template_min_cols": smallest group of columns that gets a template
    int tmpl_classmask = 3;     // option "template_class_mask": tuning aid (time the short- / long-column warps alone)
    int tmpl_flat = 0;          // option "template_flat_writeout": the walk kernel writes columns of up to this many entries out over
                                // the flat (column, position) index (0: none)
    bool rhs_local = true;      // option "rhs_local": fast right-hand side through cell-local vectors (one value per (dof, cell) pair)
    bool rhs_fast_trig = true;  // option "rhs_fast_trig": sin / cos of the registered right-hand sides through tp_sin / tp_cos (fastplan.cuh)
    int rhs_gather_ctas = 8;    // option "rhs_gather_ctas": resident CTAs per SM the cell-local gather is compiled for (5, 6, 8): it is
                                // latency-bound, full occupancy (32 registers) wins over the few spilled bytes it costs
    int rhs_groups = 1;         // option "rhs_groups": column groups a warp of the cell-local gather serves at once (1, 2, 4)
    int rhs_ahead = -1;         // option "rhs_prefetch_warps": prefetch distance of its descriptors in launch-order warps (-1: derived)
    bool bary_enabled = true;   // option "fastpath_closed_form": 0 keeps the table evaluator
    long long launches = 0;
    cudaStream_t stream2 = nullptr;   // exchange stream: the interface reduction of the matrix runs beside the rhs assembly
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool aux_pending = false;         // work on stream2 that later calls on the main stream must wait for
    cudaEvent_t cev[8] = {};          // chunk events of the pipelined lower-triangle export (created on first use)
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t uev[16] = {};
    double last_ms[3] = {0, 0, 0};
    bool timing_pending = false;
};

static std::string g_last_error;

void set_error(Ctx *ctx, int code, const std::string &msg)
{
    std::string m = "extfem error " + std::to_string(code) + ": " + msg;
    if (ctx) ctx->err = m;
    g_last_error = m;
}

static int fail(Ctx *ctx, int code, const std::string &msg)
{
    set_error(ctx, code, msg);
    return code;
}

// Kernel attributes and __constant__ banks are per DEVICE, not per context: remember per device what was set / who
// owns the constant-memory copies, so that several contexts (also on different devices) can live in one process.
constexpr int EXTFEM_MAXDEV = 64;
static std::set<const void *> g_smem_attr_done[EXTFEM_MAXDEV];
static const void *g_const_tmpl_owner[EXTFEM_MAXDEV] = {};
static const void *g_planemask_owner[EXTFEM_MAXDEV] = {};

// Contexts on one device share the __constant__ tables (templates, plane masks, reference tables).  Every launch that
// reads them records an event; an upload by ANOTHER context first makes its stream wait for that event, so a context can
// never overwrite tables a kernel of the other context is still reading (contexts used alternately without synchronize).
static const void *g_const_last_ctx[EXTFEM_MAXDEV] = {};
static cudaEvent_t g_const_ev[EXTFEM_MAXDEV] = {};
static int const_acquire(Ctx *ctx);
static int const_release(Ctx *ctx);

static int smem_attr(Ctx *ctx, const void *kernel, int bytes)
{
    auto &done = g_smem_attr_done[ctx->device % EXTFEM_MAXDEV];
    if (done.count(kernel)) return 0;
    EXTFEM_CUDA_CHECK(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    done.insert(kernel);
    return 0;
}

static int ensure(Ctx *ctx, DevBuf &b, size_t bytes)
{
    if (b.bytes >= bytes && b.p) return 0;
    if (b.p) { cudaFree(b.p); b.p = nullptr; b.bytes = 0; }
    if (bytes == 0) bytes = 16;
    EXTFEM_CUDA_CHECK(ctx, cudaMalloc(&b.p, bytes));
    b.bytes = bytes;
    return 0;
}

static inline unsigned nblocks(long long n, int bs) { return (unsigned)((n + bs - 1) / bs); }

static int const_acquire(Ctx *ctx)
{
    const int d = ctx->device % EXTFEM_MAXDEV;
    if (g_const_last_ctx[d] && g_const_last_ctx[d] != ctx && g_const_ev[d])
        EXTFEM_CUDA_CHECK(ctx, cudaStreamWaitEvent(ctx->stream, g_const_ev[d], 0));
    return 0;
}

static int const_release(Ctx *ctx)
{
    const int d = ctx->device % EXTFEM_MAXDEV;
    if (!g_const_ev[d]) EXTFEM_CUDA_CHECK(ctx, cudaEventCreateWithFlags(&g_const_ev[d], cudaEventDisableTiming));
    EXTFEM_CUDA_CHECK(ctx, cudaEventRecord(g_const_ev[d], ctx->stream));
    g_const_last_ctx[d] = ctx;
    return 0;
}

#define LAUNCHED(ctx) (++(ctx)->launches)

// ---------------------------------------------------------------------------------------------
static int upload(Ctx *ctx, DevBuf &dst, const void *src, size_t bytes)
{
    if (int rc = ensure(ctx, dst, bytes)) return rc;
    if (bytes) EXTFEM_CUDA_CHECK(ctx, cudaMemcpyAsync(dst.p, src, bytes, cudaMemcpyDefault, ctx->stream));
    return 0;
}

static int upload_indices(Ctx *ctx, DevBuf &dst, const void *src, int index_bytes, long long n, long long nmax, const char *what)
{
    // src may be host or device; stage raw bytes on the device, then convert to 0-based int32 (range-checked: 1..nmax)
    DevBuf raw, err;
    if (int rc = upload(ctx, raw, src, (size_t)n * index_bytes)) return rc;
    if (int rc = ensure(ctx, dst, (size_t)n * sizeof(int))) return rc;
    if (int rc = ensure(ctx, err, 4)) return rc;
    EXTFEM_CUDA_CHECK(ctx, cudaMemsetAsync(err.p, 0, 4, ctx->stream));
    convert_index_kernel<<<nblocks(n, 256), 256, 0, ctx->stream>>>(raw.p, index_bytes, n, nmax, dst.as<int>(), err.as<int>());
    LAUNCHED(ctx);
    int herr = 0;
    EXTFEM_CUDA_CHECK(ctx, cudaMemcpyAsync(&herr, err.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    EXTFEM_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    if (herr) return fail(ctx, EXTFEM_ERR_BAD_ARGUMENT, std::string(what) + ": index outside 1.." + std::to_string(nmax) + " (indices are 1-based)");
    return 0;
}

static int build_adjacency(Ctx *ctx, Space &S, long long ncells)
{
    long long n = ncells * S.nd;
    DevBuf keys, keys2, vals, vals2, tmp;
    if (int rc = ensure(ctx, keys, n * 8)) return rc;
    if (int rc = ensure(ctx, keys2, n * 8)) return rc;
    if (int rc = ensure(ctx, vals, n)) return rc;
    if (int rc = ensure(ctx, vals2, n)) return rc;
    adj_keys_kernel<<<nblocks(n, 256), 256, 0, ctx->stream>>>(S.celldofs.as<int>(), n, S.nd,
                                                              keys.as<unsigned long long>(), vals.as<unsigned char>());
    LAUNCHED(ctx);
    int cellbits = 1, dofbits = 1;
    while ((1ll << cellbits) < ncells) ++cellbits;
    while ((1ll << dofbits) < S.ndofs + 1) ++dofbits;
    (void)cellbits;
    size_t tmpbytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmpbytes, keys.as<unsigned long long>(), keys2.as<unsigned long long>(),
                                    vals.as<unsigned char>(), vals2.as<unsigned char>(), n, 0, 32 + dofbits, ctx->stream);
    if (int rc = ensure(ctx, tmp, tmpbytes)) return rc;
    EXTFEM_CUDA_CHECK(ctx, cub::DeviceRadixSort::SortPairs(tmp.p, tmpbytes, keys.as<unsigned long long>(),
                                                           keys2.as<unsigned long long>(), vals.as<unsigned char>(),
                                                           vals2.as<unsigned char>(), n, 0, 32 + dofbits, ctx->stream));
    LAUNCHED(ctx);
    if (int rc = ensure(ctx, S.adjptr, (S.ndofs + 1) * 8)) return rc;
    if (int rc = ensure(ctx, S.adjcell, n * 4)) return rc;
    if (int rc = ensure(ctx, S.adjloc, n)) return rc;
    adj_ptr_kernel<<<nblocks(S.ndofs + 1, 256), 256, 0, ctx->stream>>>(keys2.as<unsigned long long>(), n, S.ndofs,
                                                                       S.adjptr.as<long long>());
    LAUNCHED(ctx);
    adj_cells_kernel<<<nblocks(n, 256), 256, 0, ctx->stream>>>(keys2.as<unsigned long long>(), n, S.adjcell.as<int>());
    LAUNCHED(ctx);
    EXTFEM_CUDA_CHECK(ctx, cudaMemcpyAsync(S.adjloc.p, vals2.p, n, cudaMemcpyDeviceToDevice, ctx->stream));
    EXTFEM_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

static int get_quad(Ctx *ctx, int dim, int order, DevQuad **out)
{
    auto key = std::make_pair(dim, order);
    auto it = ctx->quads.find(key);
    if (it == ctx->quads.end()) {
        QuadRule Q = quadrature_rule(dim, order);
        auto dq = std::make_unique<DevQuad>();
        dq->nq = Q.nq;
        dq->hw = Q.w;
        dq->hx = Q.x;
        if (int rc = upload(ctx, dq->w, Q.w.data(), Q.w.size() * 8)) return rc;
        if (int rc = upload(ctx, dq->x, Q.x.data(), Q.x.size() * 8)) return rc;
        EXTFEM_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
        it = ctx->quads.emplace(key, std::move(dq)).first;
    }
    *out = it->second.get();
    return 0;
}

// scalar reference basis of space S (built-in H1P1/H1P2 or host-supplied polynomials) at the points of Q
static void space_ref_basis(const Space &S, int tdim, const QuadRule &Q, std::vector<double> &v, std::vector<double> &g)
{
    if (S.fetype == EXTFEM_FE_TABULATED) poly_basis(S.order, tdim, S.nscalar, S.poly, Q, v, g);
    else ref_basis(S.order, tdim, Q, v, g);
}

static int get_tables(Ctx *ctx, const Space &S, int dim, int quadkey, const QuadRule &Q, DevTables **out)
{
    TableKey key{dim, S.order, quadkey, S.tabkey};
    auto it = ctx->tables.find(key);
    if (it == ctx->tables.end() || quadkey < 0) {
        std::vector<double> v, g;
        space_ref_basis(S, dim, Q, v, g);
        auto dt = std::make_unique<DevTables>();
        dt->nq = Q.nq;
        if (int rc = upload(ctx, dt->vals, v.data(), v.size() * 8)) return rc;
        if (int rc = upload(ctx, dt->grads, g.data(), g.size() * 8)) return rc;
        EXTFEM_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
        if (it != ctx->tables.end()) ctx->tables.erase(it);
        it = ctx->tables.emplace(key, std::move(dt)).first;
    }
    *out = it->second.get();
    return 0;
}

static int oplen_of(int op, int ncomp, int dim)
{
    switch (op) {
    case EXTFEM_OP_ID: return ncomp;
    case EXTFEM_OP_GRAD: return ncomp * dim;
    case EXTFEM_OP_DIV: return 1;
    case EXTFEM_OP_SYMGRAD_VOIGT: return dim == 1 ? 1 : (dim == 2 ? 3 : 6);
    }
    return -1;
}

struct Prepared {
    OpDev op;
    Mesh *mesh = nullptr;
    std::vector<int> testblocks, colblocks; // unique blocks (ascending)
    std::vector<int> testlocoff, collocoff;  // operator-local offsets of these blocks
    int quadorder = 0;
    int entities = EXTFEM_ON_CELLS;
    QuadRule Q;                              // host copy of the operator's quadrature rule
};

enum OpKind { KIND_BILINEAR = 0, KIND_LINEAR = 1, KIND_NONLINEAR = 2, KIND_INTEGRATE = 3 };

static bool kernel_known(int kind, int id)
{
    if (kind == KIND_BILINEAR) return id >= EXTFEM_BLK_STANDARD && id <= EXTFEM_BLK_ROBIN108;
    if (kind == KIND_LINEAR) return id >= EXTFEM_LIN_CONSTANT_ONE && id <= EXTFEM_LIN_STEP105;
    if (kind == KIND_INTEGRATE) return id >= EXTFEM_II_STANDARD && id <= EXTFEM_II_L2ERR_EXP108;
    return id >= EXTFEM_NL_NSE2D && id <= EXTFEM_NL_POROUS106;
}

static int prepare(Ctx *ctx, Pattern &P, const extfem_opdesc *d, int kind, const double *sol, Prepared &R)
{
    if (!d) return fail(ctx, EXTFEM_ERR_BAD_ARGUMENT, "opdesc is NULL");
    if (kind == KIND_LINEAR && d->nargs > 0) {
        if (d->kernel_id != EXTFEM_BLK_STANDARD)
            return fail(ctx, EXTFEM_ERR_UNREGISTERED_KERNEL, "LinearOperator with args supports only the standard kernel");
    } else if (!kernel_known(kind, d->kernel_id))
        return fail(ctx, EXTFEM_ERR_UNREGISTERED_KERNEL,
                    "kernel id " + std::to_string(d->kernel_id) + " is not in the registry of this operator type");
    if (d->ntest < (kind == KIND_INTEGRATE ? 0 : 1) || d->ntest > MAXARGS || d->nansatz < 0 || d->nansatz > MAXARGS || d->nargs < 0 ||
        d->nargs > MAXARGS)
        return fail(ctx, EXTFEM_ERR_BAD_ARGUMENT, "number of operator arguments out of range");
    if (kind == KIND_INTEGRATE && d->nargs < 1) return fail(ctx, EXTFEM_ERR_BAD_ARGUMENT, "ItemIntegrator needs args");
    const bool onbf = d->entities == EXTFEM_ON_BFACES;
    if (d->entities != EXTFEM_ON_CELLS && !onbf) return fail(ctx, EXTFEM_ERR_UNSUPPORTED_ELEMENT, "entities must be ON_CELLS or ON_BFACES");
    if (onbf && (kind == KIND_NONLINEAR || kind == KIND_INTEGRATE))
        return fail(ctx, EXTFEM_ERR_UNSUPPORTED_ELEMENT, "ON_BFACES is built for BilinearOperator and LinearOperator only");
    if (kind == KIND_BILINEAR && d->nansatz < 1) return fail(ctx, EXTFEM_ERR_BAD_ARGUMENT, "BilinearOperator needs ansatz arguments");
    if (kind == KIND_NONLINEAR && d->nargs < 1) return fail(ctx, EXTFEM_ERR_BAD_ARGUMENT, "NonlinearOperator needs args");
    if (d->nparams > MAXPARAMS) return fail(ctx, EXTFEM_ERR_CAPACITY, "too many kernel parameters");
    if (d->nregions > MAXREGIONS) return fail(ctx, EXTFEM_ERR_CAPACITY, "too many regions");
    if ((d->nargs > 0) && !sol) return fail(ctx, EXTFEM_ERR_BAD_ARGUMENT, "operator with args needs a solution vector");

    OpDev &op = R.op;
    memset(&op, 0, sizeof(op));
    Space &S0 = *ctx->spaces[P.rowspaces[0]];
    Mesh &Mc = *ctx->meshes[S0.mesh];
    if (onbf && !Mc.bf) return fail(ctx, EXTFEM_ERR_BAD_ARGUMENT, "ON_BFACES needs extfem_mesh_set_bfaces on the grid");
    Mesh &M = onbf ? *Mc.bf : Mc;    // the item mesh: cells or boundary faces
    R.mesh = &M;
    R.entities = d->entities;
    op.dim = M.dim;
    op.ncells = M.ncells;
    op.coords = M.coordsp();
    op.cellnodes = M.cellnodes.as<int>();
    op.cellregions = M.regions.as<int>();
    op.cellvolumes = M.vol.as<double>();
    op.ntest = d->ntest; op.nansatz = d->nansatz; op.nargs = d->nargs;
    op.kernel_id = d->kernel_id;
    op.nparams = d->nparams;
    for (int i = 0; i < d->nparams; ++i) op.params[i] = d->params[i];
    op.factor = d->factor; op.time = d->time; op.offdiag = d->symgrad_offdiag;
    op.nregions = d->nregions;
    for (int i = 0; i < d->nregions; ++i) op.regions[i] = d->regions[i];
    op.lump = d->lump;
    for (int i = 0; i < MAXARGS * MAXARGS; ++i) op.coupling[i] = 1;
    if (d->coupling)
        for (int i = 0; i < d->nansatz * d->ntest; ++i) op.coupling[i] = d->coupling[i];

    // --- resolve spaces / polynomial orders
    // item spaces: the cell dof map, or FES[BFaceDofs] for ON_BFACES (nullptr when missing)
    auto item = [&](Space *S) -> Space * { return S && onbf ? S->bf.get() : S; };
    auto rowspace = [&](int b) -> Space * { return item((b >= 0 && b < (int)P.rowspaces.size()) ? ctx->spaces[P.rowspaces[b]].get() : nullptr); };
    auto colspace = [&](int b) -> Space * { return item((b >= 0 && b < (int)P.colspaces.size()) ? ctx->spaces[P.colspaces[b]].get() : nullptr); };
    const char *nospace = onbf ? "block index out of range or space without BFaceDofs (extfem_space_set_bfacedofs)" : "block index out of range";
    int poly_t = 0, poly_a = 0, poly_g = 0;
    auto poly = [&](Space *S, int o) { return S->order - (o == EXTFEM_OP_ID ? 0 : 1); };
    for (int i = 0; i < d->ntest; ++i) {
        Space *S = rowspace(d->test_block[i]);
        if (!S) return fail(ctx, EXTFEM_ERR_BAD_ARGUMENT, std::string("test: ") + nospace);
        poly_t = std::max(poly_t, poly(S, d->test_op[i]));
    }
    for (int i = 0; i < d->nansatz; ++i) {
        Space *S = colspace(d->ansatz_block[i]);
        if (!S) return fail(ctx, EXTFEM_ERR_BAD_ARGUMENT, std::string("ansatz: ") + nospace);
        poly_a = std::max(poly_a, poly(S, d->ansatz_op[i]));
    }
    for (int i = 0; i < d->nargs; ++i) {
        Space *S = colspace(d->args_block[i]);
        if (!S) return fail(ctx, EXTFEM_ERR_BAD_ARGUMENT, std::string("args: ") + nospace);
        poly_g = std::max(poly_g, poly(S, d->args_op[i]));
    }
    // quadrature order (bilinear_operator.jl:728-734, linear_operator.jl:534-538 / :299-305,
    // nonlinear_operator.jl:187-193)
    int qo;
    if (d->quadorder >= 0) qo = d->quadorder + d->bonus_quadorder;
    else if (kind == KIND_BILINEAR) qo = poly_a + poly_t + d->bonus_quadorder;
    else if (kind == KIND_LINEAR) qo = poly_t + (d->nargs > 0 ? poly_g : 0) + d->bonus_quadorder;
    else if (kind == KIND_INTEGRATE) qo = poly_g + d->bonus_quadorder;                    // item_integrator.jl:147-151
    else qo = poly_g + poly_t + d->bonus_quadorder;
    R.quadorder = qo;
    QuadRule Q;
    int quadkey = qo;
    if (d->nq_custom > 0) {
        if (!d->qweights || !d->qpoints) return fail(ctx, EXTFEM_ERR_BAD_ARGUMENT, "custom quadrature needs weights and points");
        Q.dim = M.tdim; Q.nq = d->nq_custom;
        Q.w.assign(d->qweights, d->qweights + Q.nq);
        Q.x.assign(d->qpoints, d->qpoints + (size_t)Q.nq * M.tdim);
        if (int rc = upload(ctx, ctx->custom_qw, Q.w.data(), Q.w.size() * 8)) return rc;
        if (int rc = upload(ctx, ctx->custom_qx, Q.x.data(), Q.x.size() * 8)) return rc;
        op.qw = ctx->custom_qw.as<double>(); op.qx = ctx->custom_qx.as<double>();
        quadkey = -1;
    } else {
        DevQuad *dq;
        if (int rc = get_quad(ctx, M.tdim, qo, &dq)) return rc;
        Q.dim = M.tdim; Q.nq = dq->nq; Q.w = dq->hw; Q.x = dq->hx;
        op.qw = dq->w.as<double>(); op.qx = dq->x.as<double>();
    }
    op.nq = Q.nq;
    R.Q = Q;

    // --- unique blocks and operator-local layout
    auto uniq = [](std::vector<int> v) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); return v; };
    std::vector<int> tb, cb;
    for (int i = 0; i < d->ntest; ++i) tb.push_back(d->test_block[i]);
    if (kind == KIND_NONLINEAR || kind == KIND_INTEGRATE) for (int i = 0; i < d->nargs; ++i) cb.push_back(d->args_block[i]);
    else for (int i = 0; i < d->nansatz; ++i) cb.push_back(d->ansatz_block[i]);
    R.testblocks = uniq(tb);
    R.colblocks = uniq(cb);
    int NR = 0, NC = 0;
    for (int b : R.testblocks) { R.testlocoff.push_back(NR); NR += rowspace(b)->nd; }
    for (int b : R.colblocks) { R.collocoff.push_back(NC); NC += colspace(b)->nd; }
    if (NR > MAXLOC || NC > MAXLOC) return fail(ctx, EXTFEM_ERR_CAPACITY, "operator-local matrix larger than MAXLOC");
    op.NR = NR; op.NC = NC;
    auto locoff_of = [&](const std::vector<int> &blocks, const std::vector<int> &offs, int b) {
        for (size_t i = 0; i < blocks.size(); ++i) if (blocks[i] == b) return offs[i];
        return -1;
    };
    auto fill_arg = [&](ArgDev &a, Space *S, int o, int block, int &opoff, int locoff, long long soloff) -> int {
        if (S->fetype == EXTFEM_FE_TABULATED && !S->poly_ready)
            return fail(ctx, EXTFEM_ERR_UNSUPPORTED_ELEMENT, onbf ? "host-tabulated space without a boundary-face basis (extfem_space_set_tables: bface_coeffs)"
                                                                  : "host-tabulated space without a basis (extfem_space_set_tables)");
        if (o < EXTFEM_OP_ID || o > EXTFEM_OP_SYMGRAD_VOIGT)
            return fail(ctx, EXTFEM_ERR_UNSUPPORTED_ELEMENT, "unsupported function operator");
        if (onbf && o != EXTFEM_OP_ID)
            return fail(ctx, EXTFEM_ERR_UNSUPPORTED_ELEMENT, "ON_BFACES supports Identity operators only");
        a.ncomp = S->ncomp; a.nscalar = S->nscalar; a.op = o; a.nd = S->nd;
        a.oplen = oplen_of(o, S->ncomp, M.dim);
        a.opoff = opoff; opoff += a.oplen;
        a.locoff = locoff; a.block = block;
        a.celldofs = S->celldofs.as<int>();
        a.soloff = soloff;
        DevTables *T;
        if (int rc = get_tables(ctx, *S, M.tdim, quadkey, Q, &T)) return rc;
        a.refvals = T->vals.as<double>(); a.refgrads = T->grads.as<double>();
        return 0;
    };
    int off_t = 0, off_a = 0, off_g = 0;
    for (int i = 0; i < d->ntest; ++i)
        if (int rc = fill_arg(op.test[i], rowspace(d->test_block[i]), d->test_op[i], d->test_block[i], off_t,
                              locoff_of(R.testblocks, R.testlocoff, d->test_block[i]), 0)) return rc;
    for (int i = 0; i < d->nansatz; ++i)
        if (int rc = fill_arg(op.ansatz[i], colspace(d->ansatz_block[i]), d->ansatz_op[i], d->ansatz_block[i], off_a,
                              locoff_of(R.colblocks, R.collocoff, d->ansatz_block[i]), 0)) return rc;
    for (int i = 0; i < d->nargs; ++i)
        if (int rc = fill_arg(op.args[i], colspace(d->args_block[i]), d->args_op[i], d->args_block[i], off_g,
                              (kind == KIND_NONLINEAR || kind == KIND_INTEGRATE) ? locoff_of(R.colblocks, R.collocoff, d->args_block[i]) : 0,
                              P.coloff[d->args_block[i]])) return rc;
    op.nout = off_t;
    op.nin = (kind == KIND_BILINEAR && d->nargs == 0) ? off_a : off_g;
    if (off_t > MAXOP || off_a > MAXOP || off_g > MAXOP) return fail(ctx, EXTFEM_ERR_CAPACITY, "operator vector longer than MAXOP");
    // sanity of kernel shapes
    if (kind == KIND_BILINEAR && d->kernel_id == EXTFEM_BLK_STANDARD && off_t != off_a)
        return fail(ctx, EXTFEM_ERR_BAD_ARGUMENT, "standard kernel needs equal operator lengths of test and ansatz");
    if (kind == KIND_NONLINEAR && (d->kernel_id == EXTFEM_NL_NSE2D || d->kernel_id == EXTFEM_NL_LINNSE7) &&
        (off_g != 7 || off_t != 7 || M.dim != 2))
        return fail(ctx, EXTFEM_ERR_BAD_ARGUMENT, "nse2d kernels need [id(u),grad(u),id(p)] in 2D");
    if (kind == KIND_NONLINEAR && d->kernel_id == EXTFEM_NL_NEOHOOKE3D && (off_g != 9 || off_t != 9 || M.dim != 3))
        return fail(ctx, EXTFEM_ERR_BAD_ARGUMENT, "neohooke3d needs [grad(u)] in 3D");
    if (kind == KIND_NONLINEAR && d->kernel_id == EXTFEM_NL_STVENANT230 &&
        (off_g != 4 || off_t != 4 || M.dim != 2 || d->nparams < 4 || d->nparams != 1 + 3 * (int)d->params[0]))
        return fail(ctx, EXTFEM_ERR_BAD_ARGUMENT, "stvenant230 needs [grad(u)] in 2D and params R, lambda[R], mu[R], epsT[R]");
    if (kind == KIND_NONLINEAR && (d->kernel_id == EXTFEM_NL_RCD || d->kernel_id == EXTFEM_NL_NLPOISSON105) &&
        (off_g != 1 + M.dim || off_t != 1 + M.dim || (d->kernel_id == EXTFEM_NL_NLPOISSON105 && d->nparams < 1)))
        return fail(ctx, EXTFEM_ERR_BAD_ARGUMENT, "rcd / nlpoisson105 need [id(u), grad(u)] of a scalar unknown");
    if (kind == KIND_NONLINEAR && d->kernel_id == EXTFEM_NL_POROUS106 && (off_g != 1 + M.dim || off_t != M.dim || d->nparams < 1))
        return fail(ctx, EXTFEM_ERR_BAD_ARGUMENT, "porous106 needs test [grad(u)], args [id(u), grad(u)] of a scalar unknown and params m");
    if (kind == KIND_BILINEAR && d->kernel_id == EXTFEM_BLK_ROBIN108 && (d->nparams < 1 || off_t != off_a))
        return fail(ctx, EXTFEM_ERR_BAD_ARGUMENT, "robin108 needs params g and equal operator lengths");
    if (sol) {
        if (int rc = upload(ctx, ctx->sol, sol, (size_t)P.ncols * 8)) return rc;
        op.sol = ctx->sol.as<double>();
    }
    if (kind == KIND_LINEAR && d->kernel_id == EXTFEM_LIN_TABULATED && d->nargs == 0) {
        if (!d->tabulated) return fail(ctx, EXTFEM_ERR_BAD_ARGUMENT, "tabulated kernel needs values");
        if (int rc = upload(ctx, ctx->tab, d->tabulated, (size_t)M.ncells * op.nq * op.nout * 8)) return rc;
        op.tabulated = ctx->tab.as<double>();
    }
    if (kind == KIND_INTEGRATE) op.tabulated = nullptr;   // uploaded by extfem_integrate (needs resultdim)
    return 0;
}

template <typename F>
static int dispatch_dim(Ctx *ctx, int dim, F &&f)
{
    switch (dim) {
    case 1: f(std::integral_constant<int, 1>{}); return 0;
    case 2: f(std::integral_constant<int, 2>{}); return 0;
    case 3: f(std::integral_constant<int, 3>{}); return 0;
    }
    return fail(ctx, EXTFEM_ERR_BAD_ARGUMENT, "dim must be 1, 2 or 3");
}

static int cells_per_block(size_t per_cell_bytes)
{
    size_t budget = 40 * 1024;
    int cpb = (int)(budget / per_cell_bytes);
    return std::max(1, std::min(16, cpb));
}

// --- gather launches ---------------------------------------------------------------------------
template <typename PosT>
static int launch_gather_cols(Ctx *ctx, Pattern &P, const Prepared &R, int overwrite, int transposed, double scale)
{
    GatherArgs<PosT> g;
    memset(&g, 0, sizeof(g));
    g.chunkptr = P.chunkptr.as<int>();
    g.colptr = P.colptr.as<long long>();
    g.nzval = P.nzval.as<double>();
    g.ncb = (int)P.colspaces.size();
    for (int c = 0; c <= g.ncb; ++c) g.coloff[c] = P.coloff[c];
    g.NRpat = P.NRpat; g.NRop = R.op.NR; g.NCop = R.op.NC;
    g.loc = ctx->loc.as<double>();
    g.overwrite = overwrite; g.transposed = transposed; g.scale = scale;
    const std::vector<int> &colside = transposed ? R.testblocks : R.colblocks;
    const std::vector<int> &colsideoff = transposed ? R.testlocoff : R.collocoff;
    const std::vector<int> &rowside = transposed ? R.colblocks : R.testblocks;
    const std::vector<int> &rowsideoff = transposed ? R.collocoff : R.testlocoff;
    for (int c = 0; c < g.ncb; ++c) {
        Space &S = *ctx->spaces[P.colspaces[c]];
        g.adjptr[c] = S.adjptr.as<long long>(); g.adjcell[c] = S.adjcell.as<int>(); g.adjloc[c] = S.adjloc.as<unsigned char>();
        g.posmap[c] = P.posmap[c]->as<PosT>();
        g.collocoff[c] = -1; g.colnd[c] = S.nd;
        for (size_t i = 0; i < colside.size(); ++i) if (colside[i] == c) g.collocoff[c] = colsideoff[i];
    }
    int t = 0;
    for (size_t i = 0; i < rowside.size(); ++i) {
        int b = rowside[i];
        if (b >= (int)P.rowspaces.size()) return fail(ctx, EXTFEM_ERR_BAD_ARGUMENT, "transposed copy needs a square block system");
        int nd = ctx->spaces[P.rowspaces[b]]->nd;
        for (int j = 0; j < nd; ++j, ++t) { g.rowmap[t] = P.rowlocoff[b] + j; g.rowsrc[t] = rowsideoff[i] + j; }
    }
    g.nrows_g = t;
    if (g.nrows_g >= 8 && ctx->gather_warp) {
        const bool two = g.nrows_g <= 16;
        if (transposed) { if (two) gather_columns_warp_kernel<PosT, true, 2><<<P.nchunks, GATHER_WARP_THREADS, 0, ctx->stream>>>(g);
                          else gather_columns_warp_kernel<PosT, true, 1><<<P.nchunks, GATHER_WARP_THREADS, 0, ctx->stream>>>(g); }
        else { if (two) gather_columns_warp_kernel<PosT, false, 2><<<P.nchunks, GATHER_WARP_THREADS, 0, ctx->stream>>>(g);
               else gather_columns_warp_kernel<PosT, false, 1><<<P.nchunks, GATHER_WARP_THREADS, 0, ctx->stream>>>(g); }
    }
    else gather_columns_kernel<PosT><<<P.nchunks, GATHER_THREADS, 0, ctx->stream>>>(g);
    LAUNCHED(ctx);
    EXTFEM_CUDA_CHECK(ctx, cudaGetLastError());
    return 0;
}

static int gather_matrix(Ctx *ctx, Pattern &P, const Prepared &R, int accumulate, int transposed_copy)
{
    int rc = P.poswidth == 1 ? launch_gather_cols<unsigned char>(ctx, P, R, !accumulate, 0, 1.0)
                             : launch_gather_cols<unsigned short>(ctx, P, R, !accumulate, 0, 1.0);
    if (rc) return rc;
    if (transposed_copy != 0) {
        if (!P.square) return fail(ctx, EXTFEM_ERR_BAD_ARGUMENT, "transposed_copy needs a square block system");
        rc = P.poswidth == 1 ? launch_gather_cols<unsigned char>(ctx, P, R, 0, 1, (double)transposed_copy)
                             : launch_gather_cols<unsigned short>(ctx, P, R, 0, 1, (double)transposed_copy);
    }
    return rc;
}

static int gather_vector(Ctx *ctx, Pattern &P, const Prepared &R, int accumulate)
{
    GatherVecArgs g;
    memset(&g, 0, sizeof(g));
    g.nrb = (int)P.rowspaces.size();
    for (int r = 0; r <= g.nrb; ++r) g.rowoff[r] = P.rowoff[r];
    for (int r = 0; r < g.nrb; ++r) {
        Space &S = *ctx->spaces[P.rowspaces[r]];
        g.adjptr[r] = S.adjptr.as<long long>(); g.adjcell[r] = S.adjcell.as<int>(); g.adjloc[r] = S.adjloc.as<unsigned char>();
        g.rowlocoff[r] = -1;
        for (size_t i = 0; i < R.testblocks.size(); ++i) if (R.testblocks[i] == r) g.rowlocoff[r] = R.testlocoff[i];
    }
    g.NRop = R.op.NR; g.bloc = ctx->bloc.as<double>(); g.b = P.b.as<double>(); g.overwrite = !accumulate; g.nrows = P.nrows;
    gather_rows_kernel<<<nblocks(P.nrows, 256), 256, 0, ctx->stream>>>(g);
    LAUNCHED(ctx);
    EXTFEM_CUDA_CHECK(ctx, cudaGetLastError());
    return 0;
}

// Results that stay device-resident do not make the call wait for the GPU (the host runs ahead and prepares the next
// operator); copies into caller memory do.  extfem_last_timings / extfem_synchronize wait.
static int collect_timing(Ctx *ctx);
static int finish_timing(Ctx *ctx, bool wait)
{
    EXTFEM_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
    ctx->timing_pending = true;
    if (!wait) return 0;
    return collect_timing(ctx);
}

static int collect_timing(Ctx *ctx)
{
    if (!ctx->timing_pending) return 0;
    ctx->timing_pending = false;
    EXTFEM_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    float a = 0, b = 0, c = 0;
    cudaEventElapsedTime(&a, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&b, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&c, ctx->ev[0], ctx->ev[3]);
    ctx->last_ms[0] = a; ctx->last_ms[1] = b; ctx->last_ms[2] = c;
    return 0;
}

// ---- template plan (fastplan.cuh) -------------------------------------------------------------------
template <typename K, typename V>
static int radix_sort_pairs(Ctx *ctx, const K *kin, K *kout, const V *vin, V *vout, long long n, int endbit = (int)sizeof(K) * 8)
{
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, kin, kout, vin, vout, n, 0, endbit, ctx->stream);
    DevBuf tmp;
    if (int rc = ensure(ctx, tmp, tb)) return rc;
    EXTFEM_CUDA_CHECK(ctx, cub::DeviceRadixSort::SortPairs(tmp.p, tb, kin, kout, vin, vout, n, 0, endbit, ctx->stream));
    LAUNCHED(ctx);
    EXTFEM_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream)); // tmp is freed on return
    return 0;
}

static int build_template_plan(Ctx *ctx, Pattern &P, int b, int ns)
{
    if (!P.tplans[b]) P.tplans[b] = std::make_unique<TemplatePlan>();
    TemplatePlan &T = *P.tplans[b];
    if (T.ready) return 0;
    T.ready = true;
    Space &S = *ctx->spaces[P.colspaces[b]];
    Mesh &M = *ctx->meshes[S.mesh];
    const long long ncols = S.ndofs;
    T.ncols = ncols;
    T.Lg = GeoLayout{0, 0, 1, M.ncells, M.ncells};
    cudaStream_t st = ctx->stream;
    auto all_left = [&]() -> int { // no templates: every column goes to the record kernel, geometry stays [cell][NG]
        T.nwarps = T.nctas = T.ntemplates = 0;
        T.Lg = GeoLayout{0, 0, 1, M.ncells, M.ncells};
        if (int rc = ensure(ctx, T.leftcols, (size_t)ncols * 4)) return rc;
        tp_iota_kernel<<<nblocks(ncols, 256), 256, 0, st>>>(ncols, T.leftcols.as<int>());
        LAUNCHED(ctx);
        T.nleft = ncols;
        EXTFEM_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
        return 0;
    };
    if (!ctx->tmpl_enabled || ncols < 64 || ns > 10 || P.poswidth != 1 || M.ncells >= (1ll << 31) - 4096) return all_left();
    const long long *colptr = P.colptr.as<long long>() + P.coloff[b];
    const unsigned char *posmap = P.posmap[b]->as<unsigned char>() + P.rowlocoff[b];
    const int posstride = P.NRpat;
    const long long *adjptr = S.adjptr.as<long long>();
    const int *adjcell = S.adjcell.as<int>();
    const unsigned char *adjloc = S.adjloc.as<unsigned char>();
    const unsigned gb = nblocks(ncols, 256);

    DevBuf hash, base, col, key, skey, order, flag, gid1, gstart, ok, hist;
    if (int rc = ensure(ctx, hash, ncols * 8)) return rc;
    if (int rc = ensure(ctx, base, ncols * 4)) return rc;
    if (int rc = ensure(ctx, col, ncols * 4)) return rc;
    if (int rc = ensure(ctx, key, ncols * 8)) return rc;
    if (int rc = ensure(ctx, skey, ncols * 8)) return rc;
    if (int rc = ensure(ctx, order, ncols * 4)) return rc;
    tp_sig_kernel<<<gb, 256, 0, st>>>(ncols, ns, posstride, colptr, adjptr, adjcell, adjloc, posmap, hash.as<unsigned long long>(),
                                      base.as<int>(), col.as<int>());
    LAUNCHED(ctx);
    // period P of the transposed geometry order: most frequent base-cell stride inside a signature class
    if (int rc = radix_sort_pairs(ctx, hash.as<unsigned long long>(), skey.as<unsigned long long>(), col.as<int>(), order.as<int>(), ncols))
        return rc;
    if (int rc = ensure(ctx, hist, 65 * 8)) return rc;
    EXTFEM_CUDA_CHECK(ctx, cudaMemsetAsync(hist.p, 0, 65 * 8, st));
    tp_stride_hist_kernel<<<gb, 256, 0, st>>>(ncols, skey.as<unsigned long long>(), order.as<int>(), base.as<int>(),
                                              hist.as<unsigned long long>());
    LAUNCHED(ctx);
    unsigned long long hh[65];
    EXTFEM_CUDA_CHECK(ctx, cudaMemcpyAsync(hh, hist.p, sizeof(hh), cudaMemcpyDeviceToHost, st));
    EXTFEM_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    int Pp = 1;
    for (int d = 1; d <= 64; ++d) if (hh[d] > hh[Pp]) Pp = d;
    if (hh[Pp] == 0) Pp = 1;
    GeoLayout Lg;
    Lg.soa = 1; Lg.permuted = 0; Lg.P = Pp; Lg.N = (M.ncells + Pp - 1) / Pp; Lg.Npad = Lg.N * Pp;
    // final signature includes the residue of the base cell modulo P
    tp_key2_kernel<<<gb, 256, 0, st>>>(ncols, Pp, hash.as<unsigned long long>(), base.as<int>(), key.as<unsigned long long>(), col.as<int>());
    LAUNCHED(ctx);
    if (int rc = radix_sort_pairs(ctx, key.as<unsigned long long>(), skey.as<unsigned long long>(), col.as<int>(), order.as<int>(), ncols))
        return rc;
    if (int rc = ensure(ctx, flag, ncols * 4)) return rc;
    if (int rc = ensure(ctx, gid1, ncols * 4)) return rc;
    tp_flag_kernel<<<gb, 256, 0, st>>>(ncols, skey.as<unsigned long long>(), flag.as<int>());
    LAUNCHED(ctx);
    {
        size_t tb = 0;
        cub::DeviceScan::InclusiveSum(nullptr, tb, flag.as<int>(), gid1.as<int>(), ncols, st);
        DevBuf tmp;
        if (int rc = ensure(ctx, tmp, tb)) return rc;
        EXTFEM_CUDA_CHECK(ctx, cub::DeviceScan::InclusiveSum(tmp.p, tb, flag.as<int>(), gid1.as<int>(), ncols, st));
        LAUNCHED(ctx);
        EXTFEM_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    }
    int ngroups = 0;
    EXTFEM_CUDA_CHECK(ctx, cudaMemcpy(&ngroups, gid1.as<int>() + (ncols - 1), 4, cudaMemcpyDeviceToHost));
    T.ngroups = ngroups;
    if (int rc = ensure(ctx, gstart, ((size_t)ngroups + 1) * 4)) return rc;
    tp_gstart_kernel<<<gb, 256, 0, st>>>(ncols, flag.as<int>(), gid1.as<int>(), gstart.as<int>());
    LAUNCHED(ctx);
    if (int rc = ensure(ctx, ok, ncols)) return rc;
    tp_verify_kernel<<<gb, 256, 0, st>>>(ncols, ns, posstride, Pp, order.as<int>(), gid1.as<int>(), gstart.as<int>(), colptr, adjptr,
                                         adjcell, adjloc, posmap, ok.as<unsigned char>());
    LAUNCHED(ctx);
    DevBuf gnw, gw0, gnr, gr0;
    if (int rc = ensure(ctx, gnw, ((size_t)ngroups + 1) * 4)) return rc;
    if (int rc = ensure(ctx, gw0, ((size_t)ngroups + 1) * 4)) return rc;
    if (int rc = ensure(ctx, gnr, ((size_t)ngroups + 1) * 8)) return rc;
    if (int rc = ensure(ctx, gr0, ((size_t)ngroups + 1) * 8)) return rc;
    EXTFEM_CUDA_CHECK(ctx, cudaMemsetAsync(gnw.p, 0, ((size_t)ngroups + 1) * 4, st));
    EXTFEM_CUDA_CHECK(ctx, cudaMemsetAsync(gnr.p, 0, ((size_t)ngroups + 1) * 8, st));
    tp_group_kernel<<<nblocks(ngroups, 256), 256, 0, st>>>(ngroups, ctx->tmpl_mincols, gstart.as<int>(), order.as<int>(), adjptr,
                                                           gnw.as<int>(), gnr.as<long long>());
    LAUNCHED(ctx);
    {
        size_t tb1 = 0, tb2 = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb1, gnw.as<int>(), gw0.as<int>(), ngroups + 1, st);
        cub::DeviceScan::ExclusiveSum(nullptr, tb2, gnr.as<long long>(), gr0.as<long long>(), ngroups + 1, st);
        DevBuf tmp;
        if (int rc = ensure(ctx, tmp, std::max(tb1, tb2))) return rc;
        EXTFEM_CUDA_CHECK(ctx, cub::DeviceScan::ExclusiveSum(tmp.p, tb1, gnw.as<int>(), gw0.as<int>(), ngroups + 1, st));
        EXTFEM_CUDA_CHECK(ctx, cub::DeviceScan::ExclusiveSum(tmp.p, tb2, gnr.as<long long>(), gr0.as<long long>(), ngroups + 1, st));
        LAUNCHED(ctx); LAUNCHED(ctx);
        EXTFEM_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    }
    int nwarps = 0;
    long long nrounds = 0;
    EXTFEM_CUDA_CHECK(ctx, cudaMemcpy(&nwarps, gw0.as<int>() + ngroups, 4, cudaMemcpyDeviceToHost));
    EXTFEM_CUDA_CHECK(ctx, cudaMemcpy(&nrounds, gr0.as<long long>() + ngroups, 8, cudaMemcpyDeviceToHost));
    if (nwarps == 0 || nrounds >= (1ll << 31) - 1) return all_left();

    // slots, warp descriptors (template order), leftover columns
    DevBuf wd0, wkey, wkey2, widx, widx2, left, leftsel, nsel, slotcol0, slotpb0;
    if (int rc = ensure(ctx, T.tmpl, (size_t)nrounds * TP_TW * 4)) return rc;
    const size_t nslots0 = (size_t)nwarps * 32 * TP_K;   // template order
    if (int rc = ensure(ctx, slotcol0, nslots0 * 4)) return rc;
    if (int rc = ensure(ctx, slotpb0, nslots0 * 4)) return rc;
    if (int rc = ensure(ctx, wd0, (size_t)nwarps * 16)) return rc;
    if (int rc = ensure(ctx, wkey, (size_t)nwarps * 4)) return rc;
    if (int rc = ensure(ctx, wkey2, (size_t)nwarps * 4)) return rc;
    if (int rc = ensure(ctx, widx, (size_t)nwarps * 4)) return rc;
    if (int rc = ensure(ctx, widx2, (size_t)nwarps * 4)) return rc;
    if (int rc = ensure(ctx, left, ncols)) return rc;
    if (int rc = ensure(ctx, leftsel, ncols * 4)) return rc;
    if (int rc = ensure(ctx, nsel, 8)) return rc;
    EXTFEM_CUDA_CHECK(ctx, cudaMemsetAsync(slotcol0.p, 0xff, nslots0 * 4, st));
    EXTFEM_CUDA_CHECK(ctx, cudaMemsetAsync(slotpb0.p, 0, nslots0 * 4, st));
    tp_tmpl_kernel<<<nblocks(ngroups, 128), 128, 0, st>>>(ngroups, ns, posstride, M.dim, S.order, Lg, gstart.as<int>(), order.as<int>(), gnr.as<long long>(),
                                                          gr0.as<long long>(), adjptr, adjcell, adjloc, posmap, T.tmpl.as<unsigned>());
    LAUNCHED(ctx);
    tp_slot_kernel<<<gb, 256, 0, st>>>(ncols, Lg, order.as<int>(), gid1.as<int>(), gstart.as<int>(), gnw.as<int>(), gw0.as<int>(),
                                       gnr.as<long long>(), gr0.as<long long>(), ok.as<unsigned char>(), base.as<int>(), colptr,
                                       slotcol0.as<int>(), slotpb0.as<int>(), wd0.as<int4>(), wkey.as<unsigned>(), widx.as<int>(),
                                       left.as<unsigned char>());
    LAUNCHED(ctx);
    {
        size_t tb = 0;
        cub::DeviceSelect::Flagged(nullptr, tb, order.as<int>(), left.as<unsigned char>(), leftsel.as<int>(), nsel.as<long long>(), ncols, st);
        DevBuf tmp;
        if (int rc = ensure(ctx, tmp, tb)) return rc;
        EXTFEM_CUDA_CHECK(ctx, cub::DeviceSelect::Flagged(tmp.p, tb, order.as<int>(), left.as<unsigned char>(), leftsel.as<int>(),
                                                          nsel.as<long long>(), ncols, st));
        LAUNCHED(ctx);
        EXTFEM_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    }
    long long nleft = 0;
    EXTFEM_CUDA_CHECK(ctx, cudaMemcpy(&nleft, nsel.p, 8, cudaMemcpyDeviceToHost));
    if (ncols - nleft < ncols / 2) return all_left(); // templates cover less than half of the columns: not worth the layout
    T.nleft = nleft;
    if (int rc = ensure(ctx, T.leftcols, (size_t)std::max(nleft, 1ll) * 4)) return rc;
    if (nleft > 0) {
        size_t tb = 0;
        cub::DeviceRadixSort::SortKeys(nullptr, tb, leftsel.as<int>(), T.leftcols.as<int>(), nleft, 0, 32, st);
        DevBuf tmp;
        if (int rc = ensure(ctx, tmp, tb)) return rc;
        EXTFEM_CUDA_CHECK(ctx, cub::DeviceRadixSort::SortKeys(tmp.p, tb, leftsel.as<int>(), T.leftcols.as<int>(), nleft, 0, 32, st));
        LAUNCHED(ctx);
        EXTFEM_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    }
    // launch order: warps sorted by the first adjacent cell of their first column, so that all templates sweep the
    // mesh together (geometry is read from DRAM once, output streams advance sequentially)
    if (int rc = radix_sort_pairs(ctx, wkey.as<unsigned>(), wkey2.as<unsigned>(), widx.as<int>(), widx2.as<int>(), nwarps)) return rc;
    std::vector<int4> hw((size_t)nwarps);
    std::vector<int> hidx((size_t)nwarps);
    EXTFEM_CUDA_CHECK(ctx, cudaMemcpy(hw.data(), wd0.p, (size_t)nwarps * 16, cudaMemcpyDeviceToHost));
    EXTFEM_CUDA_CHECK(ctx, cudaMemcpy(hidx.data(), widx2.p, (size_t)nwarps * 4, cudaMemcpyDeviceToHost));
    auto wneed = [&](const int4 &d) { return (tp_warp_smem(tp_desc_L(d.y), tp_desc_m(d.y)) + 1) & ~1; }; // doubles, even
    auto wcost = [&](const int4 &d) { return tp_desc_m(d.y) * tp_desc_ng(d.y); };                      // rounds of the warp
    int pool = ctx->tmpl_pool / 8;
    for (int w = 0; w < nwarps; ++w) pool = std::max(pool, wneed(hw[w]));
    // inside windows of the launch order, warps of similar cost (template rounds) go to the same CTA: a CTA's shared
    // memory and registers are held until its longest warp finishes
    for (int i0 = 0; i0 < nwarps; i0 += TP_WINDOW) {
        const int i1 = std::min(nwarps, i0 + TP_WINDOW);
        std::stable_sort(hidx.begin() + i0, hidx.begin() + i1, [&](int x, int y) { return wcost(hw[x]) > wcost(hw[y]); });
    }
    // CTAs are padded to TP_MAXW descriptors (rounds == 0: the warp exits), so that warp q = blockIdx * TP_MAXW + warp
    std::vector<int4> launch;
    std::vector<int> wpos((size_t)nwarps);
    launch.reserve((size_t)nwarps + nwarps / 2);
    int cur_w = 0, cur_s = 0, cur_m = -1;
    auto close_cta = [&]() { while (launch.size() % TP_MAXW) launch.push_back(make_int4(0, 0, 0, 0)); cur_w = 0; cur_s = 0; };
    for (int i = 0; i < nwarps; ++i) {
        const int4 d = hw[hidx[i]];
        const int mm = wcost(d);
        const int need = wneed(d);
        // new CTA: full, out of shared memory, or this warp is much cheaper than the CTA's first (longest) one
        if (cur_w > 0 && (cur_w == TP_MAXW || cur_s + need > pool || 2 * mm < cur_m)) close_cta();
        if (cur_w == 0) cur_m = mm;
        wpos[hidx[i]] = (int)launch.size();
        launch.push_back(make_int4(d.x, d.y, cur_s, 0));
        cur_s += need; ++cur_w;
    }
    close_cta();
    T.nctas = (int)(launch.size() / TP_MAXW);
    const long long nslots_l = (long long)launch.size() * 32 * TP_K;
    if (int rc = ensure(ctx, T.slotcol, (size_t)nslots_l * 4)) return rc;
    if (int rc = ensure(ctx, T.slotpb, (size_t)nslots_l * 4)) return rc;
    if (int rc = ensure(ctx, T.slotptr, (size_t)nslots_l * 8)) return rc;
    if (int rc = ensure(ctx, T.dump, 512 * 8)) return rc;
    {
        DevBuf dwpos;
        if (int rc = upload(ctx, dwpos, wpos.data(), wpos.size() * 4)) return rc;
        tp_slot_init_kernel<<<nblocks(nslots_l, 256), 256, 0, st>>>(nslots_l, T.dump.as<double>(), T.slotcol.as<int>(), T.slotpb.as<int>(),
                                                                  T.slotptr.as<double *>());
        LAUNCHED(ctx);
        tp_slot_permute_kernel<<<nblocks((long long)nslots0, 256), 256, 0, st>>>(
            (long long)nslots0, dwpos.as<int>(), slotcol0.as<int>(), slotpb0.as<int>(), colptr, P.nzval.as<double>(),
            T.dump.as<double>(), T.slotcol.as<int>(), T.slotpb.as<int>(), T.slotptr.as<double *>());
        LAUNCHED(ctx);
        EXTFEM_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    }
    T.nwarps = nwarps;
    T.nrounds = nrounds;
    T.pool_bytes = pool * 8;
    T.Lg = Lg;
    if (Lg.P > 1 && ctx->tmpl_permute_mesh) {
        const int nv = M.dim + 1;
        if (int rc = ensure(ctx, T.cn_p, (size_t)Lg.Npad * nv * 4)) return rc;
        if (int rc = ensure(ctx, T.reg_p, (size_t)Lg.Npad * 4)) return rc;
        if (int rc = ensure(ctx, T.vol_p, (size_t)Lg.Npad * 8)) return rc;
        const unsigned gp = nblocks(Lg.Npad, 256);
        if (nv == 2) tp_permute_mesh_kernel<2><<<gp, 256, 0, st>>>(Lg, M.ncells, M.cellnodes.as<int>(), M.regions.as<int>(), T.cn_p.as<int>(), T.reg_p.as<int>());
        else if (nv == 3) tp_permute_mesh_kernel<3><<<gp, 256, 0, st>>>(Lg, M.ncells, M.cellnodes.as<int>(), M.regions.as<int>(), T.cn_p.as<int>(), T.reg_p.as<int>());
        else tp_permute_mesh_kernel<4><<<gp, 256, 0, st>>>(Lg, M.ncells, M.cellnodes.as<int>(), M.regions.as<int>(), T.cn_p.as<int>(), T.reg_p.as<int>());
        LAUNCHED(ctx);
        T.Lg.permuted = 1;
        T.vol_version = -1;
    }
    {
        std::vector<int> hnw((size_t)ngroups);
        EXTFEM_CUDA_CHECK(ctx, cudaMemcpy(hnw.data(), gnw.p, (size_t)ngroups * 4, cudaMemcpyDeviceToHost));
        T.ntemplates = 0;
        for (int g = 0; g < ngroups; ++g) T.ntemplates += hnw[g] > 0;
    }
    {   // host copy of the templates: one entry per distinct first round
        std::vector<unsigned> hwords((size_t)nrounds * TP_TW);
        EXTFEM_CUDA_CHECK(ctx, cudaMemcpy(hwords.data(), T.tmpl.p, hwords.size() * 4, cudaMemcpyDeviceToHost));
        std::map<int, std::pair<int, int>> seen;
        for (int w = 0; w < nwarps; ++w) seen[hw[w].x] = std::make_pair(tp_desc_m(hw[w].y), tp_desc_L(hw[w].y));
        T.jtmpl.clear();
        for (auto &kv : seen) {
            JitTemplate jt;
            jt.r0 = kv.first; jt.m = kv.second.first; jt.L = kv.second.second;
            jt.words.assign(hwords.begin() + (size_t)jt.r0 * TP_TW, hwords.begin() + (size_t)(jt.r0 + jt.m) * TP_TW);
            T.jtmpl.push_back(std::move(jt));
        }
        // walk programs of 3D P2 columns (walkplan.h): all templates must plan and verify, else the plan keeps the
        // read-modify-write kernel
        T.walk_ok = false;
        if (ctx->tmpl_walk && M.dim == 3 && S.order == 2 && S.fetype == EXTFEM_FE_H1P2 && ns == 10 && nrounds <= TP_CONST_ROUNDS) {
            std::vector<unsigned> wrec((size_t)nrounds * TW_RW, 0u), one;
            bool ok = true;
            for (const JitTemplate &jt : T.jtmpl) {
                WalkTemplateIn W;
                W.m = jt.m; W.L = jt.L;
                for (int i = 0; i < jt.m && ok; ++i) {
                    const unsigned *w = &jt.words[(size_t)i * TP_TW];
                    W.celloff.push_back((int)w[0]);
                    W.kl.push_back((int)(w[1] & 0xff));
                    W.orient.push_back((int)((w[1] >> 20) & 1u));
                    std::array<int, 10> pp;
                    for (int t = 0; t < 10; ++t) { pp[t] = (int)(w[2 + t] / (TP_LD * 8)); ok = ok && pp[t] < jt.L; }
                    W.pos.push_back(pp);
                }
                ok = ok && tw_plan_template(W, Lg.Npad, one);
                if (!ok) break;
                std::copy(one.begin(), one.end(), wrec.begin() + (size_t)jt.r0 * TW_RW);
            }
            if (ok) {
                if (int rc = upload(ctx, T.walk, wrec.data(), wrec.size() * 4)) return rc;
                EXTFEM_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
                T.walk_ok = true;
            }
        }
        if (T.walk_ok) {
            // persistent walk kernel: the column groups are split by length into a short and a long class; a CTA has nl warps with
            // an accumulator area for the longest group (they serve the long queue, then the short one) and ns warps with a short
            // area.  Split and (nl, ns) maximise the modelled throughput: work = rounds + a per-group constant, every warp slot the
            // shared memory admits is busy until the queues are empty.
            std::map<int, std::pair<long long, int>> byL;        // L -> (work, groups)
            for (size_t i = 0; i < launch.size(); ++i)
                if (tp_desc_m(launch[i].y) > 0) { auto &e = byL[tp_desc_L(launch[i].y)]; e.first += tp_desc_m(launch[i].y) + 2; ++e.second; }
            auto area = [](int L) { return (L * TP_LD + 32 + 1) & ~1; };   // accumulators + column pointers (records are in the constant bank)
            const int Lmax = byL.empty() ? 1 : byL.rbegin()->first;
            double best = 1e300;
            int bsplit = Lmax, bnl = 1, bns = 0, bctas = 1;
            for (auto &cand : byL) {
                const int Ls = cand.first;
                double wS = 0, wL = 0;
                for (auto &kv : byL) (kv.first <= Ls ? wS : wL) += (double)kv.second.first;
                const double aS = area(Ls) * 8.0, aL = area(Lmax) * 8.0;
                for (int nl = 1; nl <= 4; ++nl)
                    for (int ns = 0; ns + nl <= 8; ++ns) {
                        if (Ls == Lmax && ns > 0) continue;      // one class only
                        if (ctx->tmpl_walk_roles > 0 && (nl != ctx->tmpl_walk_roles / 10 || ns != ctx->tmpl_walk_roles % 10)) continue;
                        const int w = nl + ns;
                        const int ctas = std::min({(int)((227.0 * 1024) / (nl * aL + ns * aS + 1024)), 24 / w, ctx->tmpl_persistent});
                        if (ctas < 1) continue;
                        const double nL = nl * ctas, nS = ns * ctas;
                        double t = wL / nL;
                        if (!(nS > 0 && wS / nS <= t)) t += (wS - nS * t) / (nL + nS);
                        if (t < best) { best = t; bsplit = Ls; bnl = nl; bns = ns; bctas = ctas; }
                    }
            }
            std::vector<int> q[2];
            for (size_t i = 0; i < launch.size(); ++i)
                if (tp_desc_m(launch[i].y) > 0) q[tp_desc_L(launch[i].y) > bsplit ? 1 : 0].push_back((int)i);
            T.walk_doubles[0] = area(bsplit); T.walk_doubles[1] = area(Lmax);
            T.walk_nl = bnl; T.walk_ns = bns; T.walk_ctas = bctas;
            for (int c = 0; c < 2; ++c) {
                T.walk_nlive[c] = (int)q[c].size();
                if (int rc = upload(ctx, T.walk_live[c], q[c].empty() ? (const int *)&c : q[c].data(), std::max<size_t>(q[c].size(), 1) * 4)) return rc;
            }
            if (int rc = ensure(ctx, T.walk_next, 8)) return rc;
            if (getenv("EXTFEM_VERBOSE")) {
                fprintf(stderr, "extfem: walk queues short %d (L <= %d, area %d doubles) long %d (area %d doubles); CTA = %d long + %d short warps, %d per SM; "
                        "column lengths (groups):", T.walk_nlive[0], bsplit, T.walk_doubles[0], T.walk_nlive[1], T.walk_doubles[1], bnl, bns, bctas);
                for (auto &kv : byL) fprintf(stderr, " %d:%d", kv.first, kv.second.second);
                fprintf(stderr, "\n");
            }
        }
        std::vector<int> live[2];
        for (size_t i = 0; i < launch.size(); ++i)
            if (tp_desc_m(launch[i].y) > 0) live[tp_desc_L(launch[i].y) > JIT_SPLIT_L ? 1 : 0].push_back((int)i);
        // CTAs of one template: inside windows of 256 warps of the sweep (the same mesh neighbourhood, so that the interleaved
        // output columns still merge in L2) the warps are grouped by template -- the 4 warps of a CTA then run the same
        // straight-line code and fetch it once
        for (int c = 0; c < 2; ++c)
            for (size_t i0 = 0; i0 < live[c].size(); i0 += 256) {
                const size_t i1 = std::min(live[c].size(), i0 + 256);
                std::stable_sort(live[c].begin() + i0, live[c].begin() + i1, [&](int x, int y) { return launch[x].x < launch[y].x; });
            }
        for (int c = 0; c < 2; ++c) {
            T.jit_nlive[c] = (int)live[c].size();
            if (!live[c].empty()) if (int rc = upload(ctx, T.jit_live[c], live[c].data(), live[c].size() * 4)) return rc;
        }
    }
    if (int rc = upload(ctx, T.wdesc, launch.data(), launch.size() * 16)) return rc;
    tp_affine_kernel<<<nblocks((long long)launch.size(), 8), 256, 0, st>>>((int)launch.size(), T.slotcol.as<int>(), T.slotptr.as<double *>(),
                                                                         T.wdesc.as<int4>());
    LAUNCHED(ctx);
    EXTFEM_CUDA_CHECK(ctx, cudaGetLastError());
    EXTFEM_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    return 0;
}

// ---- fast path (fastpath.cuh) --------------------------------------------------------------------
// Plan of one diagonal block: columns sorted by adjacency signature inside windows, cut into chunks of FP_T
// slots, chunks ordered by shared-memory class; lane-contiguous records per warp round.
static int build_fast_plan(Ctx *ctx, Pattern &P, int b, int ns)
{
    if (!P.fastplans[b]) P.fastplans[b] = std::make_unique<FastPlan>();
    FastPlan &F = *P.fastplans[b];
    if (F.ready) return 0;
    F.ready = true;
    Space &S = *ctx->spaces[P.colspaces[b]];
    Mesh &M = *ctx->meshes[S.mesh];
    if (M.ncells + 64 >= (1ll << FP_CELLBITS) || ns > (1 << FP_KLBITS)) return 0; // record word 0 cannot hold cell | kl
    if (int rc = build_template_plan(ctx, P, b, ns)) return rc;
    TemplatePlan &T = *P.tplans[b];
    const long long ncols = T.nleft;            // the record kernel owns the columns without a template
    if (ncols == 0) { F.nchunks = 0; F.usable = true; return 0; }
    const long long *colptr = P.colptr.as<long long>() + P.coloff[b];
    F.nchunks = (int)((ncols + FP_T - 1) / FP_T);
    const long long nslots = (long long)F.nchunks * FP_T;
    {
        DevBuf key, key2, col, order, tmp;
        if (int rc = ensure(ctx, key, ncols * 8)) return rc;
        if (int rc = ensure(ctx, key2, ncols * 8)) return rc;
        if (int rc = ensure(ctx, col, ncols * 4)) return rc;
        if (int rc = ensure(ctx, order, ncols * 4)) return rc;
        fp_key_kernel<<<nblocks(ncols, 256), 256, 0, ctx->stream>>>(ncols, T.leftcols.as<int>(), S.adjptr.as<long long>(),
                                                                   S.adjloc.as<unsigned char>(), key.as<unsigned long long>(), col.as<int>());
        LAUNCHED(ctx);
        size_t tb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb, key.as<unsigned long long>(), key2.as<unsigned long long>(), col.as<int>(),
                                        order.as<int>(), ncols, 0, 64, ctx->stream);
        if (int rc = ensure(ctx, tmp, tb)) return rc;
        EXTFEM_CUDA_CHECK(ctx, cub::DeviceRadixSort::SortPairs(tmp.p, tb, key.as<unsigned long long>(), key2.as<unsigned long long>(),
                                                               col.as<int>(), order.as<int>(), ncols, 0, 64, ctx->stream));
        LAUNCHED(ctx);
        if (int rc = ensure(ctx, F.slotcol, (size_t)nslots * 4)) return rc;
        if (int rc = ensure(ctx, F.slotoff, (size_t)nslots * 4)) return rc;
        if (int rc = ensure(ctx, F.chunktot, (size_t)F.nchunks * 4)) return rc;
        if (int rc = ensure(ctx, F.warpniter, (size_t)F.nchunks * FP_W * 4)) return rc;
        fp_chunk_kernel<<<F.nchunks, FP_T, 0, ctx->stream>>>(ncols, order.as<int>(), colptr, S.adjptr.as<long long>(), F.slotcol.as<int>(),
                                                            F.slotoff.as<int>(), F.chunktot.as<int>(), F.warpniter.as<int>());
        LAUNCHED(ctx);
        EXTFEM_CUDA_CHECK(ctx, cudaGetLastError());
        EXTFEM_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    std::vector<int> hn((size_t)F.nchunks * FP_W), htot((size_t)F.nchunks);
    EXTFEM_CUDA_CHECK(ctx, cudaMemcpyAsync(hn.data(), F.warpniter.p, hn.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    EXTFEM_CUDA_CHECK(ctx, cudaMemcpyAsync(htot.data(), F.chunktot.p, htot.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    EXTFEM_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<int> lists[FP_NCLASS];
    for (int c = 0; c < F.nchunks; ++c) {
        if (htot[c] > fp_cap(FP_NCLASS - 1)) return 0; // chunk exceeds the largest shared-memory class: generic path
        int cls = 0;
        while (htot[c] > fp_cap(cls)) ++cls;
        lists[cls].push_back(c);
        F.cls_maxtot[cls] = std::max(F.cls_maxtot[cls], htot[c]);
    }
    std::vector<int> all;
    for (int c = 0; c < FP_NCLASS; ++c) {
        F.cls_start[c] = (int)all.size();
        all.insert(all.end(), lists[c].begin(), lists[c].end());
    }
    F.cls_start[FP_NCLASS] = (int)all.size();
    if (int rc = upload(ctx, F.chunklist, all.data(), all.size() * 4)) return rc;
    const int rw = fp_rw(ns);
    std::vector<long long> hoff(hn.size());
    long long tot = 0;
    for (size_t i = 0; i < hn.size(); ++i) { hoff[i] = tot; tot += (long long)hn[i] * rw * 32; }
    if (int rc = upload(ctx, F.warpoff, hoff.data(), hoff.size() * 8)) return rc;
    if (int rc = ensure(ctx, F.rec, (size_t)tot * 4 + 16)) return rc;
    fp_fill_kernel<unsigned char><<<nblocks(nslots, 256), 256, 0, ctx->stream>>>(
        nslots, ns, rw, P.NRpat, F.slotcol.as<int>(), F.warpniter.as<int>(), F.warpoff.as<long long>(), S.adjptr.as<long long>(),
        S.adjcell.as<int>(), S.adjloc.as<unsigned char>(), P.posmap[b]->as<unsigned char>() + P.rowlocoff[b], F.rec.as<unsigned>(), T.Lg);
    LAUNCHED(ctx);
    EXTFEM_CUDA_CHECK(ctx, cudaGetLastError());
    EXTFEM_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    F.usable = true;
    return 0;
}

// transposed-order copy of the cell volumes, refreshed when the mesh's volumes changed
static int refresh_permuted_volumes(Ctx *ctx, Mesh &M, TemplatePlan &T)
{
    if (!T.Lg.permuted || T.vol_version == M.vol_version) return 0;
    tp_permute_vol_kernel<<<nblocks(T.Lg.Npad, 256), 256, 0, ctx->stream>>>(T.Lg, M.ncells, M.vol.as<double>(), T.vol_p.as<double>());
    LAUNCHED(ctx);
    EXTFEM_CUDA_CHECK(ctx, cudaGetLastError());
    T.vol_version = M.vol_version;
    return 0;
}

template <class EV, bool FIRST, bool CT>
static int launch_template(Ctx *ctx, Pattern &P, TemplatePlan &T, int b, int accumulate)
{
    TPArgs A;
    A.wdesc = T.wdesc.as<int4>(); A.slotpb = T.slotpb.as<int>(); A.slotptr = T.slotptr.as<double *>();
    A.tmpl = T.tmpl.as<unsigned>(); A.geo = ctx->geo.as<double>(); A.Npad = T.Lg.Npad; A.overwrite = !accumulate;
    A.nwarps = T.nctas * TP_MAXW; A.ahead = ctx->tmpl_ahead * TP_MAXW; A.classmask = ctx->tmpl_classmask; A.flat = 0;
    (void)P; (void)b;
    auto k = tp_gather_kernel<EV, FIRST, CT>;
    if (int rc = smem_attr(ctx, (const void *)k, 227 * 1024)) return rc;
    k<<<T.nctas, TP_MAXW * 32, T.pool_bytes, ctx->stream>>>(A);
    LAUNCHED(ctx);
    EXTFEM_CUDA_CHECK(ctx, cudaGetLastError());
    return 0;
}

template <int DIM, int GEO, class EV, bool SOA>
static int launch_fast_layout(Ctx *ctx, Pattern &P, FastPlan &F, TemplatePlan &T, int b, const Prepared &R, const extfem_opdesc *d,
                              double geoscale, int accumulate)
{
    constexpr int NG = fp_ng(DIM, GEO);
    static_assert(NG == EV::NG, "geometry record / evaluator mismatch");
    Mesh &M = *R.mesh;
    if (int rc = ensure(ctx, ctx->geo, (size_t)(SOA ? T.Lg.Npad : M.ncells) * NG * 8)) return rc;
    if (d->nregions > 0) if (int rc = upload(ctx, ctx->visit, d->regions, (size_t)d->nregions * 4)) return rc;
    if (int rc = refresh_permuted_volumes(ctx, M, T)) return rc;
    const bool perm = SOA && T.Lg.permuted;
    fp_geo_kernel<DIM, GEO, SOA><<<nblocks(perm ? T.Lg.Npad : M.ncells, 256), 256, 0, ctx->stream>>>(
        M.ncells, M.coords.as<double>(), perm ? T.cn_p.as<int>() : M.cellnodes.as<int>(), perm ? T.reg_p.as<int>() : M.regions.as<int>(),
        perm ? T.vol_p.as<double>() : M.vol.as<double>(), d->factor * geoscale, d->nregions, ctx->visit.as<int>(), ctx->geo.as<double>(), T.Lg);
    LAUNCHED(ctx);
    EXTFEM_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    bool jit_done = false;
    if (SOA && T.nwarps > 0 && EV::BARY && ctx->jit_enabled && T.ncols >= ctx->jit_mincols && !T.jtmpl.empty()) {
        // plan-time specialisation of the templates (jit.cuh); any failure keeps the static kernel
        if (!T.jit) T.jit = std::make_unique<JitModule>();
        JitModule &J = *T.jit;
        if (!J.tried) {
            J.tried = true;
            const auto t0 = std::chrono::steady_clock::now();
            JitGen gen{EV::DIM_, EV::ORDER_, EV::NV, EV::NS};
            bool present[2];
            const std::string src = gen.source(T.jtmpl, JIT_SPLIT_L, present);
            J.ok = jit_compile(src, present, J);
            J.compile_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (!J.ok && getenv("EXTFEM_JIT_VERBOSE")) fprintf(stderr, "extfem: template specialisation failed: %s\n", J.log.c_str());
        }
        if (J.ok) {
            JitArgs A;
            A.wdesc = T.wdesc.as<int4>(); A.slotpb = T.slotpb.as<int>(); A.slotptr = T.slotptr.as<double *>();
            A.geo = ctx->geo.as<double>(); A.Npad = T.Lg.Npad; A.overwrite = !accumulate; A.nlive = 0; A.live = nullptr;
            void *args[] = {&A};
            for (int c = 1; c >= 0; --c) {   // long columns first
                if (!J.present[c] || T.jit_nlive[c] == 0) continue;
                A.nlive = T.jit_nlive[c]; A.live = T.jit_live[c].as<int>();
                EXTFEM_CUDA_CHECK(ctx, cudaLaunchKernel((const void *)J.kernel[c], dim3(nblocks(A.nlive, TP_MAXW)), dim3(TP_MAXW * 32), args,
                                                        (size_t)TP_MAXW * (32 * TP_LD + 32) * 8, ctx->stream));
                LAUNCHED(ctx);
            }
            jit_done = true;
        }
    }
    bool walk_done = false;
    if (SOA && T.nwarps > 0 && !jit_done && std::is_same<EV, EvalBary<3, 2>>::value && T.walk_ok && ctx->tmpl_walk) {
        // walk programs (walkplan.h): records live in the constant bank of the templates
        const bool first = !accumulate && P.rowspaces.size() == 1;
        const void *owner = (const char *)&T + 1;
        if (g_const_tmpl_owner[ctx->device % EXTFEM_MAXDEV] != owner) {
            EXTFEM_CUDA_CHECK(ctx, cudaMemcpyToSymbolAsync(c_tp_tmpl, T.walk.p, (size_t)T.nrounds * TW_RW * 4, 0, cudaMemcpyDeviceToDevice, ctx->stream));
            g_const_tmpl_owner[ctx->device % EXTFEM_MAXDEV] = owner;
        }
        TPArgs A;
        A.wdesc = T.wdesc.as<int4>(); A.slotpb = T.slotpb.as<int>(); A.slotptr = T.slotptr.as<double *>();
        A.tmpl = T.walk.as<unsigned>(); A.geo = ctx->geo.as<double>(); A.Npad = T.Lg.Npad; A.overwrite = !accumulate;
        A.nwarps = T.nctas * TP_MAXW; A.ahead = ctx->tmpl_ahead * TP_MAXW; A.classmask = ctx->tmpl_classmask; A.flat = ctx->tmpl_flat;
        if (ctx->tmpl_persistent && ctx->tmpl_classmask == 3 && T.walk_nlive[0] + T.walk_nlive[1] > 0) {
            // persistent CTAs: tmpl_persistent CTAs per SM (limited by the accumulator areas), the warps take groups from two queues
            auto k = first ? tw_gather_persistent_kernel<true> : tw_gather_persistent_kernel<false>;
            if (int rc = smem_attr(ctx, (const void *)k, 227 * 1024)) return rc;
            TWQueues Q;
            for (int c = 0; c < 2; ++c) { Q.live[c] = T.walk_live[c].as<int>(); Q.n[c] = T.walk_nlive[c]; Q.doubles[c] = T.walk_doubles[c]; }
            Q.next = T.walk_next.as<int>();
            EXTFEM_CUDA_CHECK(ctx, cudaMemsetAsync(Q.next, 0, 8, ctx->stream));
            Q.nlong = T.walk_nl;
            const int w = T.walk_nl + T.walk_ns;
            const size_t smem = (size_t)(T.walk_nl * T.walk_doubles[1] + T.walk_ns * T.walk_doubles[0]) * 8;
            const int grid = std::min((T.walk_nlive[0] + T.walk_nlive[1] + w - 1) / w, ctx->sm_count * T.walk_ctas);
            k<<<grid, w * 32, smem, ctx->stream>>>(A, Q);
        } else {
            auto k = first ? tw_gather_kernel<true> : tw_gather_kernel<false>;
            if (int rc = smem_attr(ctx, (const void *)k, 227 * 1024)) return rc;
            k<<<T.nctas, TP_MAXW * 32, T.pool_bytes, ctx->stream>>>(A);
        }
        LAUNCHED(ctx);
        EXTFEM_CUDA_CHECK(ctx, cudaGetLastError());
        walk_done = true;
    }
    if (SOA && T.nwarps > 0 && !jit_done && !walk_done) {
        // first-touch stores need: overwrite, and column segments that hold rows of this block only
        const bool first = !accumulate && P.rowspaces.size() == 1;
        // templates in constant memory when they fit (re-uploaded when another plan used the bank in between)
        const bool ct = ctx->tmpl_const && T.nrounds <= TP_CONST_ROUNDS;
        if (ct && g_const_tmpl_owner[ctx->device % EXTFEM_MAXDEV] != &T) {
            EXTFEM_CUDA_CHECK(ctx, cudaMemcpyToSymbolAsync(c_tp_tmpl, T.tmpl.p, (size_t)T.nrounds * TP_TW * 4, 0, cudaMemcpyDeviceToDevice, ctx->stream));
            g_const_tmpl_owner[ctx->device % EXTFEM_MAXDEV] = &T;
        }
        {
            static const char ev_tag = 0;   // one per evaluator instantiation
            const void *owner = ctx->tmpl_planemask ? (const void *)&ev_tag : (const void *)ctx;
            if (g_planemask_owner[ctx->device % EXTFEM_MAXDEV] != owner) {
                unsigned pm[16];
                for (int kl = 0; kl < 16; ++kl) pm[kl] = ctx->tmpl_planemask ? EV::plane_mask(kl < EV::NS ? kl : 0) : (1u << EV::NG) - 1u;
                EXTFEM_CUDA_CHECK(ctx, cudaMemcpyToSymbolAsync(c_tp_planemask, pm, sizeof(pm), 0, cudaMemcpyHostToDevice, ctx->stream));
                EXTFEM_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));   // pm lives on the stack
                g_planemask_owner[ctx->device % EXTFEM_MAXDEV] = owner;
            }
        }
        int rc;
        if (ct) rc = first ? launch_template<EV, true, true>(ctx, P, T, b, accumulate) : launch_template<EV, false, true>(ctx, P, T, b, accumulate);
        else rc = first ? launch_template<EV, true, false>(ctx, P, T, b, accumulate) : launch_template<EV, false, false>(ctx, P, T, b, accumulate);
        if (rc) return rc;
    }
    if (F.nchunks > 0) {
        FastArgs A;
        A.plan.chunklist = F.chunklist.as<int>(); A.plan.slotcol = F.slotcol.as<int>(); A.plan.slotoff = F.slotoff.as<int>();
        A.plan.chunktot = F.chunktot.as<int>();
        A.plan.warpniter = F.warpniter.as<int>(); A.plan.warpoff = F.warpoff.as<long long>(); A.plan.rec = F.rec.as<unsigned>();
        A.colptr = P.colptr.as<long long>() + P.coloff[b];
        A.nzval = P.nzval.as<double>(); A.geo = ctx->geo.as<double>(); A.Npad = T.Lg.Npad; A.overwrite = !accumulate;
        // One launch per shared-memory class.  The small class runs one warp group per chunk; the larger classes are
        // limited by shared memory, so their local rows are split over two warp groups (vertex rows | other rows) when
        // the element has both (EV::NS > EV::NV).
        constexpr int NGRP_BIG = EV::NS > EV::NV ? 2 : 1;
        auto k0 = fp_gather_kernel<EV, 1, 6, SOA>;
        auto k1 = fp_gather_kernel<EV, NGRP_BIG, (NGRP_BIG == 2 ? 3 : 4), SOA>;
        auto k2 = fp_gather_kernel<EV, NGRP_BIG, 3, SOA>;
        if (int rc = smem_attr(ctx, (const void *)k0, fp_cap(FP_NCLASS - 1) * 8)) return rc;
        if (int rc = smem_attr(ctx, (const void *)k1, fp_cap(FP_NCLASS - 1) * 8)) return rc;
        if (int rc = smem_attr(ctx, (const void *)k2, fp_cap(FP_NCLASS - 1) * 8)) return rc;
        for (int c = FP_NCLASS - 1; c >= 0; --c) { // longest-running class first
            int n = F.cls_start[c + 1] - F.cls_start[c];
            if (n == 0) continue;
            A.chunk0 = F.cls_start[c];
            const size_t smem = (size_t)F.cls_maxtot[c] * 8; // what the class actually needs
            if (c == 0) k0<<<n, FP_T, smem, ctx->stream>>>(A);
            else if (c == 1) k1<<<n, FP_T * NGRP_BIG, smem, ctx->stream>>>(A);
            else k2<<<n, FP_T * NGRP_BIG, smem, ctx->stream>>>(A);
            LAUNCHED(ctx);
        }
    }
    EXTFEM_CUDA_CHECK(ctx, cudaGetLastError());
    EXTFEM_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    return 0;
}

template <int DIM, int GEO, class EV>
static int launch_fast(Ctx *ctx, Pattern &P, FastPlan &F, int b, const Prepared &R, const extfem_opdesc *d, double geoscale,
                       int accumulate)
{
    TemplatePlan &T = *P.tplans[b];
    return T.Lg.soa ? launch_fast_layout<DIM, GEO, EV, true>(ctx, P, F, T, b, R, d, geoscale, accumulate)
                    : launch_fast_layout<DIM, GEO, EV, false>(ctx, P, F, T, b, R, d, geoscale, accumulate);
}

// reference tables S[kl][t][g] of the table evaluator, from the SAME quadrature rule / basis as the generic path
static std::vector<double> fast_tables(const QuadRule &Q, int order, int dim, int form)
{
    std::vector<double> v, g;
    ref_basis(order, dim, Q, v, g);
    const int ns = nscalar_of(order, dim), ng = fp_ng(dim, form == FP_FORM_MASS ? FP_GEO_VOLUME : FP_GEO_METRIC);
    std::vector<double> T((size_t)ng * ns * ns, 0.0);
    for (int q = 0; q < Q.nq; ++q)
        for (int t = 0; t < ns; ++t)
            for (int kl = 0; kl < ns; ++kl) {
                double *out = &T[((size_t)kl * ns + t) * ng];
                if (form == FP_FORM_MASS) { out[0] += Q.w[q] * v[(size_t)q * ns + t] * v[(size_t)q * ns + kl]; continue; }
                int gi = 0;
                for (int dd = 0; dd < dim; ++dd)
                    for (int ee = dd; ee < dim; ++ee, ++gi) {
                        double a = g[((size_t)q * ns + t) * dim + dd] * g[((size_t)q * ns + kl) * dim + ee];
                        if (ee != dd) a += g[((size_t)q * ns + t) * dim + ee] * g[((size_t)q * ns + kl) * dim + dd];
                        out[gi] += Q.w[q] * a;
                    }
            }
    return T;
}

// The closed-form evaluator hard-codes the P1/P2 basis and exact integration.  Before it is used, check it on the
// host against the tables of the operator's actual quadrature rule / basis on a generic (non-degenerate) simplex;
// any mismatch (different basis convention, inexact rule) keeps the table evaluator.
template <int DIM, int ORDER>
static bool verify_bary_against_tables(const std::vector<double> &T)
{
    constexpr int NV = DIM + 1, NS = ORDER == 1 ? NV : NV * (NV + 1) / 2, NG = DIM * (DIM + 1) / 2;
    // gradients of barycentric coordinates for J^-1 = B (rows), a fixed generic matrix
    double B[3][3] = {{1.3, -0.2, 0.1}, {0.4, 0.9, -0.3}, {-0.5, 0.25, 1.1}};
    double L[NV][DIM];
    for (int x = 0; x < DIM; ++x) {
        double s = 0;
        for (int r = 0; r < DIM; ++r) { L[r + 1][x] = B[r][x]; s -= B[r][x]; }
        L[0][x] = s;
    }
    double G[NG], Dg[NG], Mfull[NV][NV];
    int o = 0;
    for (int dd = 0; dd < DIM; ++dd)
        for (int ee = dd; ee < DIM; ++ee) {
            double s = 0;
            for (int x = 0; x < DIM; ++x) s += B[dd][x] * B[ee][x];
            G[o++] = s;
        }
    for (int a = 0; a < NV; ++a)
        for (int c = a + 1; c < NV; ++c) {
            double s = 0;
            for (int x = 0; x < DIM; ++x) s += L[a][x] * L[c][x];
            Dg[fp_pair_index<DIM>(a, c)] = s * fp_bary_scale(DIM, ORDER);
        }
    fp_bary_expand<DIM>(Dg, Mfull);
    double scale = 0, err = 0;
    for (int kl = 0; kl < NS; ++kl)
        for (int t = 0; t < NS; ++t) {
            double ref = 0;
            for (int g = 0; g < NG; ++g) ref += G[g] * T[((size_t)kl * NS + t) * NG + g];
            double got = fp_bary_acc<DIM, ORDER>(t, kl, Mfull, 0.0);
            scale = std::max(scale, std::fabs(ref));
            err = std::max(err, std::fabs(ref - got));
        }
    return err <= 1e-13 * scale;
}

// fast paths for the headline configurations; *fast == false -> generic path
static int try_fast_bilinear(Ctx *ctx, Pattern &P, const Prepared &R, const extfem_opdesc *d, int accumulate, bool *fast)
{
    *fast = false;
    if (!ctx->fast_enabled) return 0;
    const OpDev &op = R.op;
    if (d->kernel_id != EXTFEM_BLK_STANDARD || d->ntest != 1 || d->nansatz != 1 || d->nargs != 0 || d->lump != 0 ||
        d->transposed_copy != 0 || d->test_block[0] != d->ansatz_block[0] || d->entities != EXTFEM_ON_CELLS)
        return 0;
    const int b = d->test_block[0];
    if (b >= (int)P.colspaces.size() || b >= (int)P.rowspaces.size() || P.rowspaces[b] != P.colspaces[b]) return 0;
    if (!P.coupling[(size_t)b * P.rowspaces.size() + b] || P.poswidth != 1) return 0;
    Space &S = *ctx->spaces[P.colspaces[b]];
    if (S.ncomp != 1 || S.nscalar > 10 || S.fetype == EXTFEM_FE_TABULATED) return 0;
    int form;
    if (d->test_op[0] == EXTFEM_OP_GRAD && d->ansatz_op[0] == EXTFEM_OP_GRAD) form = FP_FORM_LAPLACE;
    else if (d->test_op[0] == EXTFEM_OP_ID && d->ansatz_op[0] == EXTFEM_OP_ID) form = FP_FORM_MASS;
    else return 0;
    const int dim = op.dim, ns = S.nscalar;
    if (int rc = build_fast_plan(ctx, P, b, ns)) return rc;
    FastPlan &F = *P.fastplans[b];
    if (!F.usable) return 0;
    // other column blocks of the pattern are not touched by this operator: zero them when overwriting
    if (!accumulate && P.colspaces.size() > 1) {
        for (size_t c = 0; c < P.colspaces.size(); ++c) {
            if ((int)c == b) continue;
            long long n0 = P.hcolptr[P.coloff[c]], n1 = P.hcolptr[P.coloff[c + 1]];
            EXTFEM_CUDA_CHECK(ctx, cudaMemsetAsync(P.nzval.as<double>() + n0, 0, (size_t)(n1 - n0) * 8, ctx->stream));
        }
    }
    const QuadRule &Q = R.Q;   // host copy kept by prepare(): no device round trip
    std::vector<double> T = fast_tables(Q, S.order, dim, form);
    int rc = -1;
    if (int rca = const_acquire(ctx)) return rca;
    // closed-form barycentric evaluators (Laplace, P1/P2, 2D/3D), verified against the tables
    if (form == FP_FORM_LAPLACE && ctx->bary_enabled) {
#define FP_BARY(D, O) \
        if (rc == -1 && dim == D && S.order == O && verify_bary_against_tables<D, O>(T)) \
            rc = launch_fast<D, FP_GEO_BARY, EvalBary<D, O>>(ctx, P, F, b, R, d, fp_bary_scale(D, O), accumulate);
        FP_BARY(3, 2) FP_BARY(3, 1) FP_BARY(2, 2) FP_BARY(2, 1)
#undef FP_BARY
    }
    if (rc == -1) {
        EXTFEM_CUDA_CHECK(ctx, cudaMemcpyToSymbolAsync(c_fp_S, T.data(), T.size() * 8, 0, cudaMemcpyHostToDevice, ctx->stream));
#define FP_TAB(D, N, FRM, GEO) \
        if (rc == -1 && dim == D && ns == N && form == FRM) \
            rc = launch_fast<D, GEO, EvalTable<N, fp_ng(D, GEO), D + 1>>(ctx, P, F, b, R, d, 1.0, accumulate);
        FP_TAB(1, 2, FP_FORM_LAPLACE, FP_GEO_METRIC) FP_TAB(1, 3, FP_FORM_LAPLACE, FP_GEO_METRIC)
        FP_TAB(2, 3, FP_FORM_LAPLACE, FP_GEO_METRIC) FP_TAB(2, 6, FP_FORM_LAPLACE, FP_GEO_METRIC)
        FP_TAB(3, 4, FP_FORM_LAPLACE, FP_GEO_METRIC) FP_TAB(3, 10, FP_FORM_LAPLACE, FP_GEO_METRIC)
        FP_TAB(1, 2, FP_FORM_MASS, FP_GEO_VOLUME) FP_TAB(1, 3, FP_FORM_MASS, FP_GEO_VOLUME)
        FP_TAB(2, 3, FP_FORM_MASS, FP_GEO_VOLUME) FP_TAB(2, 6, FP_FORM_MASS, FP_GEO_VOLUME)
        FP_TAB(3, 4, FP_FORM_MASS, FP_GEO_VOLUME) FP_TAB(3, 10, FP_FORM_MASS, FP_GEO_VOLUME)
#undef FP_TAB
    }
    if (rc == -1) return 0; // no instantiation: generic path
    if (rc) return rc;
    if (int rcr = const_release(ctx)) return rcr;
    *fast = true;
    return 0;
}

// host tables of the warp-per-cell nonlinear local kernel (kernels_generic.cuh: local_nonlinear_kernel3)
struct NL3Dof { int n = 0, space = 0, scalar = 0; int out[CVC_MAX], src[CVC_MAX]; double scale[CVC_MAX]; };

static bool nl3_dofops(const ArgDev *args, int nargs, int dim, double offdiag, int loc, const std::vector<const double *> &spaces, NL3Dof &D)
{
    int n = 0;
    for (int i = 0; i < nargs; ++i) {
        const ArgDev &a = args[i];
        if (loc < a.locoff || loc >= a.locoff + a.nd) continue;
        const int jj = loc - a.locoff, c = jj / a.nscalar, k = jj % a.nscalar;
        int sp = 0;
        while (sp < (int)spaces.size() && spaces[sp] != a.refvals) ++sp;
        if (n > 0 && (D.space != sp || D.scalar != k)) return false;   // arguments on one block must share the space
        D.space = sp; D.scalar = k;
        auto add = [&](int out, int src, double scale) { if (n < CVC_MAX) { D.out[n] = a.opoff + out; D.src[n] = src; D.scale[n] = scale; } ++n; };
        switch (a.op) {
        case EXTFEM_OP_ID: add(c, 0, 1.0); break;
        case EXTFEM_OP_GRAD: for (int d = 0; d < dim; ++d) add(c * dim + d, 1 + d, 1.0); break;
        case EXTFEM_OP_DIV: add(0, 1 + (c < dim ? c : 0), 1.0); break;
        default: // SYMGRAD_VOIGT, mirrors eval_cv
            if (dim == 1) add(0, 1, 1.0);
            else if (dim == 2) { add(c, 1 + c, 1.0); add(2, 1 + ((1 - c) & 1), offdiag); }
            else {
                const int o1 = (c == 0) ? 4 : 3, o2 = (c == 2) ? 4 : 5, g1 = (c == 0) ? 2 : ((c == 1) ? 2 : 1), g2 = (c == 0) ? 1 : 0;
                add(c, 1 + c, 1.0); add(o1, 1 + g1, offdiag); add(o2, 1 + g2, offdiag);
            }
        }
    }
    D.n = n;
    return n <= CVC_MAX;
}

static int build_nl3_tables(Ctx *ctx, const OpDev &op, NL3Tables &T, bool *ok, bool want_dense = false)
{
    *ok = false;
    memset(&T, 0, sizeof(T));
    if (op.NR > 255 || op.NC > 255 || op.nin > 255 || op.nout > 255 || op.nq > 255) return 0;
    std::vector<const double *> spaces, sgrads;
    std::vector<int> sns;
    auto note = [&](const ArgDev &a) {
        for (auto p : spaces) if (p == a.refvals) return;
        spaces.push_back(a.refvals); sgrads.push_back(a.refgrads); sns.push_back(a.nscalar);
    };
    for (int i = 0; i < op.nargs; ++i) note(op.args[i]);
    for (int i = 0; i < op.ntest; ++i) note(op.test[i]);
    if ((int)spaces.size() > NL2_MAXSP_) return 0;
    std::vector<NL3Dof> cols(op.NC), rows(op.NR);
    for (int j = 0; j < op.NC; ++j) if (!nl3_dofops(op.args, op.nargs, op.dim, op.offdiag, j, spaces, cols[j])) return 0;
    for (int k = 0; k < op.NR; ++k) if (!nl3_dofops(op.test, op.ntest, op.dim, op.offdiag, k, spaces, rows[k])) return 0;
    std::vector<int> phi_off;
    int off = 0;
    for (size_t s = 0; s < spaces.size(); ++s) { phi_off.push_back(off); off += op.nq * sns[s] * (1 + op.dim); }
    bool same = op.NC == op.NR;
    for (int j = 0; same && j < op.NC; ++j) {
        same = cols[j].n == rows[j].n && cols[j].space == rows[j].space && cols[j].scalar == rows[j].scalar;
        for (int x = 0; same && x < cols[j].n; ++x)
            same = cols[j].out[x] == rows[j].out[x] && cols[j].src[x] == rows[j].src[x] && cols[j].scale[x] == rows[j].scale[x];
    }
    std::vector<unsigned char> host;
    auto reserve = [&](size_t bytes) { size_t o = (host.size() + 15) / 16 * 16; host.resize(o + bytes, 0); return (int)o; };
    auto pack = [&](const std::vector<NL3Dof> &D, int N, int &E, int &o_n, int &o_out, int &o_idx, int &o_sc) {
        E = 1;
        for (int i = 0; i < N; ++i) E = std::max(E, D[i].n);
        o_n = reserve(N);
        o_out = reserve((size_t)E * N);
        o_idx = reserve((size_t)op.nq * E * N * 4);
        o_sc = reserve((size_t)op.nq * E * N * 8);
        for (int i = 0; i < N; ++i) {
            host[o_n + i] = (unsigned char)D[i].n;
            for (int x = 0; x < E; ++x) {
                host[o_out + (size_t)x * N + i] = x < D[i].n ? (unsigned char)D[i].out[x] : 0;
                for (int q = 0; q < op.nq; ++q) {
                    const size_t it = ((size_t)q * E + x) * N + i;
                    reinterpret_cast<int *>(host.data() + o_idx)[it] =
                        x < D[i].n ? phi_off[D[i].space] + (q * sns[D[i].space] + D[i].scalar) * (1 + op.dim) + D[i].src[x] : -1;
                    reinterpret_cast<double *>(host.data() + o_sc)[it] = x < D[i].n ? D[i].scale[x] : 0.0;
                }
            }
        }
    };
    T.NC = op.NC; T.NR = op.NR; T.same = same ? 1 : 0;
    pack(cols, op.NC, T.EC, T.o_cn, T.o_cout, T.o_bgidx, T.o_bgsc);
    {   // dense operator matrix of the tensor-core kernel (kernels_generic.cuh: local_nonlinear_kernel4)
        T.NT4 = (std::max(op.NC, op.NR) + 7) / 8;
        T.Np4 = T.NT4 <= 2 ? 20 : 36;
        T.dense_ok = same && op.nin == op.nout && op.nin <= 9 && T.NT4 <= 4 && T.EC * op.NC <= 128 &&
                     nl4_warp_doubles(op.nq, op.nin, op.nin * op.nout, T.Np4, off) * 8 * 4 <= 120 * 1024;
        // structural non-zeros of the kernel's Jacobian (probed on the device with the operator's parameters)
        T.nnzJ = op.nin * op.nout;
        T.o_jslot = reserve((size_t)op.nin * op.nout);
        for (int e = 0; e < op.nin * op.nout; ++e) host[T.o_jslot + e] = (unsigned char)e;
        if (want_dense && T.dense_ok && ctx->nl_sparse_jac) {
            DevBuf dm;
            std::vector<unsigned char> hm((size_t)op.nin * op.nout);
            if (int rc = ensure(ctx, dm, hm.size())) return rc;
            EXTFEM_CUDA_CHECK(ctx, cudaMemsetAsync(dm.p, 0, hm.size(), ctx->stream));
            nl_mask_kernel<<<1, 32, 0, ctx->stream>>>(op, op.dim, dm.as<unsigned char>());
            LAUNCHED(ctx);
            EXTFEM_CUDA_CHECK(ctx, cudaMemcpyAsync(hm.data(), dm.p, hm.size(), cudaMemcpyDeviceToHost, ctx->stream));
            EXTFEM_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
            int n = 0;
            for (size_t e = 0; e < hm.size(); ++e) host[T.o_jslot + e] = hm[e] ? (unsigned char)n++ : (unsigned char)255;
            T.nnzJ = std::max(n, 1);
        }
        T.o_pt = reserve((size_t)T.EC * op.NC * 4);
        for (int x = 0; x < T.EC; ++x)
            for (int j = 0; j < op.NC; ++j) {
                unsigned char *e = host.data() + T.o_pt + ((size_t)x * op.NC + j) * 4;
                const bool live = x < cols[j].n;
                e[0] = live ? (unsigned char)cols[j].space : 0; e[1] = live ? (unsigned char)cols[j].scalar : 0;
                e[2] = live ? (unsigned char)cols[j].src[x] : 0; e[3] = 0;
            }
        T.o_ddst = reserve((size_t)T.EC * op.NC * 2);
        for (int x = 0; x < T.EC; ++x)
            for (int j = 0; j < op.NC; ++j)
                reinterpret_cast<unsigned short *>(host.data() + T.o_ddst)[(size_t)x * op.NC + j] =
                    (unsigned short)((T.dense_ok && x < cols[j].n ? cols[j].out[x] : 0) * T.Np4 + j);
    }
    if (same) { T.ER = T.EC; T.o_rn = T.o_cn; T.o_rout = T.o_cout; T.o_btidx = T.o_bgidx; T.o_btsc = T.o_bgsc; }
    else pack(rows, op.NR, T.ER, T.o_rn, T.o_rout, T.o_btidx, T.o_btsc);
    if ((size_t)op.nq * op.nin * op.nout >= 65536 || (size_t)op.nq * T.EC * op.NC >= 65536 || (size_t)op.nq * T.ER * op.NR >= 65536 ||
        (size_t)op.nq * op.NC * op.nout >= 65536)
        return 0;   // 16-bit offsets
    const std::vector<NL3Dof> &rws = same ? cols : rows;
    const bool skip_sparse = want_dense && T.dense_ok;   // the tensor-core kernel needs only the B tables and their dense places
    if (!skip_sparse) {
    T.o_gjj = reserve((size_t)op.nq * op.NC * op.nout * T.EC * 2);
    T.o_gjb = reserve((size_t)op.nq * op.NC * op.nout * T.EC * 2);
    {
        size_t i = 0;
        for (int q = 0; q < op.nq; ++q)
            for (int j = 0; j < op.NC; ++j)
                for (int t = 0; t < op.nout; ++t, ++i)
                    for (int x = 0; x < T.EC; ++x) {
                        const bool live = x < cols[j].n;
                        reinterpret_cast<unsigned short *>(host.data() + T.o_gjj)[i * T.EC + x] =
                            (unsigned short)(q * op.nin * op.nout + t * op.nin + (live ? cols[j].out[x] : 0));
                        reinterpret_cast<unsigned short *>(host.data() + T.o_gjb)[i * T.EC + x] = (unsigned short)((q * T.EC + x) * op.NC + j);
                    }
    }
    T.o_enb = reserve((size_t)op.NR * op.NC * T.ER * 2);
    T.o_eng = reserve((size_t)op.NR * op.NC * T.ER * 2);
    for (int j = 0; j < op.NC; ++j)
        for (int k = 0; k < op.NR; ++k)
            for (int x = 0; x < T.ER; ++x) {
                const size_t i = ((size_t)j * op.NR + k) * T.ER + x;
                const bool live = x < rws[k].n;
                reinterpret_cast<unsigned short *>(host.data() + T.o_enb)[i] = (unsigned short)(x * op.NR + k);
                reinterpret_cast<unsigned short *>(host.data() + T.o_eng)[i] = (unsigned short)(j * op.nout + (live ? rws[k].out[x] : 0));
            }
    }
    {
        std::vector<unsigned short> uptr(op.nin + 1, 0), ulist;
        for (int o = 0; o < op.nin; ++o) {
            for (int j = 0; j < op.NC; ++j)
                for (int x = 0; x < cols[j].n; ++x)
                    if (cols[j].out[x] == o) ulist.push_back((unsigned short)(j | (x << 8)));
            uptr[o + 1] = (unsigned short)ulist.size();
        }
        T.o_uptr = reserve(uptr.size() * 2);
        memcpy(host.data() + T.o_uptr, uptr.data(), uptr.size() * 2);
        T.o_ulist = reserve(std::max<size_t>(ulist.size(), 1) * 2);
        if (!ulist.empty()) memcpy(host.data() + T.o_ulist, ulist.data(), ulist.size() * 2);
    }
    reserve(0);
    host.resize((host.size() + 15) / 16 * 16, 0);
    T.tab_bytes = (int)host.size();
    if (int rc = upload(ctx, ctx->nl2buf, host.data(), host.size())) return rc;
    EXTFEM_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    T.tab = ctx->nl2buf.as<unsigned char>();
    T.nspaces = (int)spaces.size();
    for (int s = 0; s < T.nspaces; ++s) { T.ns[s] = sns[s]; T.refvals[s] = spaces[s]; T.refgrads[s] = sgrads[s]; T.phi_off[s] = phi_off[s]; }
    T.phi_off[T.nspaces] = off;
    for (int i = 0; i < op.nargs; ++i) {
        bool seen = false;
        for (int b = 0; b < T.ncolblocks; ++b) seen |= T.blk_locoff[b] == op.args[i].locoff;
        if (seen) continue;
        const int b = T.ncolblocks++;
        T.blk_locoff[b] = op.args[i].locoff; T.blk_nd[b] = op.args[i].nd; T.blk_celldofs[b] = op.args[i].celldofs;
        T.blk_soloff[b] = op.args[i].soloff;
    }
    *ok = true;
    return 0;
}

// rules (sizes of the engine's own rules up to order 4 or 5) the cell-local form of the fast right-hand side is compiled for
static bool rhs_local_rule(int dim, int nq)
{
    if (dim == 1) return nq >= 1 && nq <= 4;
    if (dim == 2) return nq == 1 || nq == 3 || nq == 4 || nq == 9;
    return nq == 1 || nq == 4 || nq == 8;
}

// fast right-hand side: LinearOperator(f, [id(u)]) on a scalar P1/P2 space (linear_operator.jl:584-640).  Per cell the
// point values factor*w_q*|T|*f(x_q) (structure-of-arrays, geometry order), per dof an owner-computes sum over the
// adjacent cells driven by the template plan of the block.
static int try_fast_linear(Ctx *ctx, Pattern &P, const Prepared &R, const extfem_opdesc *d, int accumulate, bool *fast)
{
    *fast = false;
    if (!ctx->fast_enabled) return 0;
    const OpDev &op = R.op;
    if (d->ntest != 1 || d->nargs != 0 || d->test_op[0] != EXTFEM_OP_ID || op.nout != 1 || op.nq > TP_NQMAX || d->entities != EXTFEM_ON_CELLS) return 0;
    const int b = d->test_block[0];
    if (!P.square || b >= (int)P.colspaces.size() || P.poswidth != 1 || !P.coupling[(size_t)b * P.rowspaces.size() + b]) return 0;
    Space &S = *ctx->spaces[P.rowspaces[b]];
    if (S.ncomp != 1 || S.nscalar > 10 || S.fetype == EXTFEM_FE_TABULATED) return 0;
    Mesh &M = *R.mesh;
    const int dim = op.dim, ns = S.nscalar;
    if (int rc = build_template_plan(ctx, P, b, ns)) return rc;
    TemplatePlan &T = *P.tplans[b];
    // reference basis values at the operator's quadrature points
    const QuadRule &Q = R.Q;   // host copy kept by prepare(): no device round trip
    std::vector<double> v, g, phi((size_t)10 * TP_NQMAX, 0.0);
    ref_basis(S.order, dim, Q, v, g);
    for (int q = 0; q < op.nq; ++q)
        for (int kl = 0; kl < ns; ++kl) phi[(size_t)kl * TP_NQMAX + q] = v[(size_t)q * ns + kl];
    if (int rca = const_acquire(ctx)) return rca;
    EXTFEM_CUDA_CHECK(ctx, cudaMemcpyToSymbolAsync(c_tp_phi, phi.data(), phi.size() * 8, 0, cudaMemcpyHostToDevice, ctx->stream));
    // cell-local form (one plane per local dof) for the rules with a compile-time kernel, else point values (one plane per point)
    const bool local = ctx->rhs_local && rhs_local_rule(dim, op.nq) && ns == (S.order == 1 ? dim + 1 : (dim + 1) * (dim + 2) / 2);
    if (int rc = ensure(ctx, ctx->fq, (size_t)T.Lg.Npad * (local ? ns : op.nq) * 8)) return rc;
    double *bblk = P.b.as<double>() + P.rowoff[b];
    if (!accumulate && P.rowspaces.size() > 1) {
        for (size_t r = 0; r < P.rowspaces.size(); ++r) {
            if ((int)r == b) continue;
            EXTFEM_CUDA_CHECK(ctx, cudaMemsetAsync(P.b.as<double>() + P.rowoff[r], 0, (size_t)(P.rowoff[r + 1] - P.rowoff[r]) * 8, ctx->stream));
        }
    }
    RhsCellArgs C;
    memset(&C, 0, sizeof(C));
    if (int rc = refresh_permuted_volumes(ctx, M, T)) return rc;
    const bool perm = T.Lg.permuted != 0;
    C.ncells = M.ncells; C.coords = M.coords.as<double>(); C.cellnodes = perm ? T.cn_p.as<int>() : M.cellnodes.as<int>();
    C.regions = perm ? T.reg_p.as<int>() : M.regions.as<int>();
    C.vol = perm ? T.vol_p.as<double>() : M.vol.as<double>(); C.qw = op.qw; C.qx = op.qx; C.tabulated = op.tabulated; C.nq = op.nq; C.kernel_id = op.kernel_id;
    C.nregions = op.nregions;
    for (int i = 0; i < op.nregions; ++i) C.visit[i] = op.regions[i];
    for (int i = 0; i < op.nparams; ++i) C.params[i] = op.params[i];
    C.factor = op.factor; C.Lg = T.Lg; C.fq = ctx->fq.as<double>();
    const unsigned gcell = nblocks(perm ? T.Lg.Npad : M.ncells, 256);
    if (local) {
        RhsCellLocalArgs CL;
        memset(&CL, 0, sizeof(CL));
        CL.C = C; CL.ns = ns;
        for (int q = 0; q < op.nq; ++q) {
            CL.qw[q] = Q.w[q];
            for (int r = 0; r < dim; ++r) CL.qx[q * dim + r] = Q.x[(size_t)q * dim + r];
        }
        bool launched = false;
#define RHS_LOCAL(D, N, S) \
        if (!launched && dim == D && op.nq == N && ns == S) { \
            if (op.kernel_id == EXTFEM_LIN_SINCOS301 && ctx->rhs_fast_trig) tp_rhs_cell_local_kernel<D, N, S, TP_KID_SINCOS301_FT><<<gcell, 256, 0, ctx->stream>>>(CL); \
            else tp_rhs_cell_local_kernel<D, N, S, -1><<<gcell, 256, 0, ctx->stream>>>(CL); \
            launched = true; \
        }
#define RHS_LOCAL2(D, N, S1, S2) RHS_LOCAL(D, N, S1) RHS_LOCAL(D, N, S2)
        RHS_LOCAL2(1, 1, 2, 3) RHS_LOCAL2(1, 2, 2, 3) RHS_LOCAL2(1, 3, 2, 3) RHS_LOCAL2(1, 4, 2, 3)
        RHS_LOCAL2(2, 1, 3, 6) RHS_LOCAL2(2, 3, 3, 6) RHS_LOCAL2(2, 4, 3, 6) RHS_LOCAL2(2, 9, 3, 6)
        RHS_LOCAL2(3, 1, 4, 10) RHS_LOCAL2(3, 4, 4, 10) RHS_LOCAL2(3, 8, 4, 10)
#undef RHS_LOCAL2
#undef RHS_LOCAL
        if (!launched) return fail(ctx, EXTFEM_ERR_CAPACITY, "fast right-hand side: no cell kernel for this rule");
    }
    else if (dim == 1) tp_rhs_cell_kernel<1><<<gcell, 256, 0, ctx->stream>>>(C);
    else if (dim == 2) tp_rhs_cell_kernel<2><<<gcell, 256, 0, ctx->stream>>>(C);
    else tp_rhs_cell_kernel<3><<<gcell, 256, 0, ctx->stream>>>(C);
    LAUNCHED(ctx);
    EXTFEM_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    if (T.nwarps > 0) {
        TPRhsArgs A;
        A.nwarps = T.nctas * TP_MAXW; A.wdesc = T.wdesc.as<int4>(); A.slotcol = T.slotcol.as<int>(); A.slotpb = T.slotpb.as<int>();
        A.tmpl = T.tmpl.as<unsigned>(); A.fq = ctx->fq.as<double>(); A.Npad = T.Lg.Npad; A.nq = op.nq; A.b = bblk; A.overwrite = !accumulate;
        const unsigned gr = nblocks((long long)T.nctas * TP_MAXW * TP_K, 8);
        // descriptors are prefetched about one and a half waves of resident warps ahead
        const int G = ctx->rhs_groups >= 4 ? 4 : ctx->rhs_groups >= 2 ? 2 : 1;
        A.ahead = ctx->rhs_ahead >= 0 ? ctx->rhs_ahead : ctx->sm_count * 48 * G;
        if (local && ctx->rhs_groups >= 4) tp_rhs_local_kernel<4, 2><<<nblocks((long long)A.nwarps, 32), 256, 0, ctx->stream>>>(A);
        else if (local && ctx->rhs_groups >= 2) tp_rhs_local_kernel<2, 4><<<nblocks((long long)A.nwarps, 16), 256, 0, ctx->stream>>>(A);
        else if (local && ctx->rhs_gather_ctas >= 8) tp_rhs_local_kernel<1, 8><<<gr, 256, 0, ctx->stream>>>(A);
        else if (local && ctx->rhs_gather_ctas >= 6) tp_rhs_local_kernel<1, 6><<<gr, 256, 0, ctx->stream>>>(A);
        else if (local) tp_rhs_local_kernel<1, 5><<<gr, 256, 0, ctx->stream>>>(A);
        else if (op.nq == 1) tp_rhs_kernel<1><<<gr, 256, 0, ctx->stream>>>(A);
        else if (op.nq == 3) tp_rhs_kernel<3><<<gr, 256, 0, ctx->stream>>>(A);
        else if (op.nq == 4) tp_rhs_kernel<4><<<gr, 256, 0, ctx->stream>>>(A);
        else if (op.nq == 6) tp_rhs_kernel<6><<<gr, 256, 0, ctx->stream>>>(A);
        else tp_rhs_kernel<0><<<gr, 256, 0, ctx->stream>>>(A);
        LAUNCHED(ctx);
    }
    if (T.nleft > 0) {
        RhsLeftArgs A;
        A.nleft = T.nleft; A.leftcols = T.leftcols.as<int>(); A.adjptr = S.adjptr.as<long long>(); A.adjcell = S.adjcell.as<int>();
        A.adjloc = S.adjloc.as<unsigned char>(); A.fq = ctx->fq.as<double>(); A.Lg = T.Lg; A.nq = op.nq; A.b = bblk; A.overwrite = !accumulate;
        A.local = local ? 1 : 0;
        tp_rhs_left_kernel<<<nblocks(T.nleft, 256), 256, 0, ctx->stream>>>(A);
        LAUNCHED(ctx);
    }
    EXTFEM_CUDA_CHECK(ctx, cudaGetLastError());
    EXTFEM_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    if (int rcr = const_release(ctx)) return rcr;
    *fast = true;
    return 0;
}

// ---- ON_BFACES (entities.cuh): face-local kernels + owner-computes scatter into the cell pattern ---------------------
static int scatter_faces_matrix(Ctx *ctx, Pattern &P, const Prepared &R)
{
    DevBuf err;
    if (int rc = ensure(ctx, err, 4)) return rc;
    EXTFEM_CUDA_CHECK(ctx, cudaMemsetAsync(err.p, 0, 4, ctx->stream));
    for (size_t ic = 0; ic < R.colblocks.size(); ++ic) {
        const int cb = R.colblocks[ic];
        Space &Sc = *ctx->spaces[P.colspaces[cb]]->bf;
        FaceScatterArgs A;
        memset(&A, 0, sizeof(A));
        A.ncols = Sc.ndofs; A.colbase = P.coloff[cb];
        A.adjptr = Sc.adjptr.as<long long>(); A.adjface = Sc.adjcell.as<int>(); A.adjloc = Sc.adjloc.as<unsigned char>();
        A.collocoff = R.collocoff[ic];
        A.nrb = 0;
        for (size_t ir = 0; ir < R.testblocks.size(); ++ir) {
            const int rb = R.testblocks[ir];
            if (!P.coupling[(size_t)cb * P.rowspaces.size() + rb]) continue;   // block not in the pattern
            Space &Sr = *ctx->spaces[P.rowspaces[rb]]->bf;
            A.rowfacedofs[A.nrb] = Sr.celldofs.as<int>(); A.rownd[A.nrb] = Sr.nd; A.rowoff[A.nrb] = P.rowoff[rb];
            A.rowlocoff[A.nrb] = R.testlocoff[ir];
            ++A.nrb;
        }
        A.NRop = R.op.NR; A.NCop = R.op.NC; A.loc = ctx->loc.as<double>();
        A.colptr = P.colptr.as<long long>(); A.rowval = P.rowval.as<int>(); A.nzval = P.nzval.as<double>(); A.error = err.as<int>();
        face_scatter_matrix_kernel<<<nblocks(A.ncols, 256), 256, 0, ctx->stream>>>(A);
        LAUNCHED(ctx);
    }
    int herr = 0;
    EXTFEM_CUDA_CHECK(ctx, cudaMemcpyAsync(&herr, err.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    EXTFEM_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    if (herr) return fail(ctx, EXTFEM_ERR_BAD_ARGUMENT, "ON_BFACES: a boundary-face entry is not part of the matrix pattern (BFaceDofs inconsistent with CellDofs?)");
    return 0;
}

static int scatter_faces_vector(Ctx *ctx, Pattern &P, const Prepared &R)
{
    for (size_t ir = 0; ir < R.testblocks.size(); ++ir) {
        const int rb = R.testblocks[ir];
        Space &Sr = *ctx->spaces[P.rowspaces[rb]]->bf;
        FaceScatterVecArgs A;
        A.nrows = Sr.ndofs; A.rowbase = P.rowoff[rb];
        A.adjptr = Sr.adjptr.as<long long>(); A.adjface = Sr.adjcell.as<int>(); A.adjloc = Sr.adjloc.as<unsigned char>();
        A.rowlocoff = R.testlocoff[ir]; A.NRop = R.op.NR; A.bloc = ctx->bloc.as<double>(); A.b = P.b.as<double>();
        face_scatter_vector_kernel<<<nblocks(A.nrows, 256), 256, 0, ctx->stream>>>(A);
        LAUNCHED(ctx);
    }
    EXTFEM_CUDA_CHECK(ctx, cudaGetLastError());
    return 0;
}

static int assemble_bfaces(Ctx *ctx, Pattern &P, const Prepared &R, const extfem_opdesc *d, int kind, int accumulate)
{
    const OpDev &op = R.op;
    if (d->transposed_copy != 0 || d->lump != 0)
        return fail(ctx, EXTFEM_ERR_UNSUPPORTED_ELEMENT, "ON_BFACES: transposed_copy / lump are not built");
    if (kind == KIND_LINEAR && d->nargs > 0 && d->kernel_id != EXTFEM_BLK_STANDARD)
        return fail(ctx, EXTFEM_ERR_UNREGISTERED_KERNEL, "LinearOperator with args supports only the standard kernel");
    const int tdim = R.mesh->tdim;
    EXTFEM_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    if (kind == KIND_BILINEAR) {
        if (int rc = ensure(ctx, ctx->loc, (size_t)op.ncells * op.NR * op.NC * 8)) return rc;
        if (!accumulate) EXTFEM_CUDA_CHECK(ctx, cudaMemsetAsync(P.nzval.p, 0, (size_t)P.nnz * 8, ctx->stream));
        const long long n = op.ncells * op.NR * op.NC;
        int rcd = dispatch_dim(ctx, op.dim, [&](auto dimc) {
            constexpr int DIM = decltype(dimc)::value;
            face_bilinear_kernel<DIM><<<nblocks(n, 256), 256, 0, ctx->stream>>>(op, tdim, ctx->loc.as<double>());
        });
        if (rcd) return rcd;
        LAUNCHED(ctx);
        EXTFEM_CUDA_CHECK(ctx, cudaGetLastError());
        EXTFEM_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
        if (int rc = scatter_faces_matrix(ctx, P, R)) return rc;
    } else {
        if (int rc = ensure(ctx, ctx->bloc, (size_t)op.ncells * op.NR * 8)) return rc;
        if (!accumulate) EXTFEM_CUDA_CHECK(ctx, cudaMemsetAsync(P.b.p, 0, (size_t)P.nrows * 8, ctx->stream));
        const long long n = op.ncells * op.NR;
        int rcd = dispatch_dim(ctx, op.dim, [&](auto dimc) {
            constexpr int DIM = decltype(dimc)::value;
            face_linear_kernel<DIM><<<nblocks(n, 256), 256, 0, ctx->stream>>>(op, tdim, ctx->bloc.as<double>());
        });
        if (rcd) return rcd;
        LAUNCHED(ctx);
        EXTFEM_CUDA_CHECK(ctx, cudaGetLastError());
        EXTFEM_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
        if (int rc = scatter_faces_vector(ctx, P, R)) return rc;
    }
    EXTFEM_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    return 0;
}

} // namespace extfem

using namespace extfem;

// work queued on the exchange stream is ordered before anything a later call puts on the main stream (every entry point except
// extfem_assemble_linear, which touches neither the matrix values nor the exchange buffers)
static inline void join_aux(Ctx *C)
{
    if (C->aux_pending) { cudaStreamWaitEvent(C->stream, C->ev_join, 0); C->aux_pending = false; }
}

#define CTX_GUARD_NOJOIN(ctx)                                                           \
    if (!(ctx)) { set_error(nullptr, EXTFEM_ERR_BAD_ARGUMENT, "ctx is NULL"); return EXTFEM_ERR_BAD_ARGUMENT; } \
    Ctx *C = reinterpret_cast<Ctx *>(ctx);                                              \
    cudaSetDevice(C->device);
#define CTX_GUARD(ctx) CTX_GUARD_NOJOIN(ctx) join_aux(C);

#define GET_PATTERN(id)                                                                 \
    if ((id) < 0 || (id) >= (int)C->patterns.size() || !C->patterns[id])                \
        return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "invalid pattern handle");             \
    Pattern &P = *C->patterns[id];

extern "C" {

int extfem_ctx_create(int device, extfem_ctx **out)
{
    if (!out) return fail(nullptr, EXTFEM_ERR_BAD_ARGUMENT, "out is NULL");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, EXTFEM_ERR_CUDA, std::string("no CUDA device available (there is no CPU fallback): ") +
                                                  cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, EXTFEM_ERR_BAD_ARGUMENT, "device index out of range");
    Ctx *C = new Ctx();
    C->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&C->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete C;
        return fail(nullptr, EXTFEM_ERR_CUDA, "cannot initialise device");
    }
    if (cudaDeviceGetAttribute(&C->sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || C->sm_count <= 0) C->sm_count = 148;
    for (auto &ev : C->ev) cudaEventCreate(&ev);
    for (auto &ev : C->uev) cudaEventCreate(&ev);
    cudaStreamCreateWithFlags(&C->stream2, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&C->ev_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&C->ev_join, cudaEventDisableTiming);
    if (const char *e = getenv("EXTFEM_DISABLE_FASTPATH")) C->fast_enabled = !(e[0] == '1');
    *out = reinterpret_cast<extfem_ctx *>(C);
    if (const char *e = getenv("EXTFEM_OPTIONS")) {   // tuning knobs for every context of the process: "key=value,key=value"
        std::string all(e);
        size_t i = 0;
        while (i < all.size()) {
            size_t j = all.find(',', i);
            if (j == std::string::npos) j = all.size();
            const std::string kv = all.substr(i, j - i);
            const size_t q = kv.find('=');
            if (q != std::string::npos) extfem_set_option(*out, kv.substr(0, q).c_str(), atoi(kv.c_str() + q + 1));
            i = j + 1;
        }
    }
    return EXTFEM_OK;
}

int extfem_ctx_destroy(extfem_ctx *ctx)
{
    if (!ctx) return EXTFEM_OK;
    Ctx *C = reinterpret_cast<Ctx *>(ctx);
    cudaSetDevice(C->device);
    cudaStreamSynchronize(C->stream);
    for (auto &ev : C->ev) if (ev) cudaEventDestroy(ev);
    for (auto &ev : C->uev) if (ev) cudaEventDestroy(ev);
    for (auto &ev : C->cev) if (ev) cudaEventDestroy(ev);
    if (C->stream2) { cudaStreamSynchronize(C->stream2); cudaStreamDestroy(C->stream2); }
    if (C->ev_fork) cudaEventDestroy(C->ev_fork);
    if (C->ev_join) cudaEventDestroy(C->ev_join);
    C->patterns.clear(); C->spaces.clear(); C->meshes.clear();
    g_const_tmpl_owner[C->device % EXTFEM_MAXDEV] = nullptr;   // plans of this context may have owned the constant banks
    g_planemask_owner[C->device % EXTFEM_MAXDEV] = nullptr;
    if (g_const_last_ctx[C->device % EXTFEM_MAXDEV] == C) g_const_last_ctx[C->device % EXTFEM_MAXDEV] = nullptr;
    if (C->dist.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(C->dist.comm);
    cudaStream_t s = C->stream;
    delete C;
    cudaStreamDestroy(s);
    return EXTFEM_OK;
}

const char *extfem_last_error(extfem_ctx *ctx)
{
    if (ctx) return reinterpret_cast<Ctx *>(ctx)->err.c_str();
    return g_last_error.c_str();
}

int extfem_kernel_id(const char *name)
{
    static const struct { const char *n; int id; } tab[] = {
        {"standard", EXTFEM_BLK_STANDARD}, {"dcr", EXTFEM_BLK_DCR}, {"stokes", EXTFEM_BLK_STOKES},
        {"linnse7", EXTFEM_BLK_LINNSE7}, {"hooke_grad", EXTFEM_BLK_HOOKE_GRAD}, {"hooke_voigt", EXTFEM_BLK_HOOKE_VOIGT},
        {"convect_args", EXTFEM_BLK_CONVECT_ARGS}, {"robin108", EXTFEM_BLK_ROBIN108},
        {"constant_one", EXTFEM_LIN_CONSTANT_ONE}, {"constant_params", EXTFEM_LIN_CONSTANT_PARAMS}, {"xy", EXTFEM_LIN_XY},
        {"sincos301", EXTFEM_LIN_SINCOS301}, {"tabulated", EXTFEM_LIN_TABULATED}, {"exp2x", EXTFEM_LIN_EXP2X}, {"step105", EXTFEM_LIN_STEP105},
        {"nse2d", EXTFEM_NL_NSE2D}, {"nl_linnse7", EXTFEM_NL_LINNSE7}, {"neohooke3d", EXTFEM_NL_NEOHOOKE3D}, {"rcd", EXTFEM_NL_RCD},
        {"nlpoisson105", EXTFEM_NL_NLPOISSON105}, {"stvenant230", EXTFEM_NL_STVENANT230}, {"porous106", EXTFEM_NL_POROUS106},
        {"ii_standard", EXTFEM_II_STANDARD}, {"l2norm", EXTFEM_II_L2NORM}, {"l2diff_tabulated", EXTFEM_II_L2DIFF_TABULATED},
        {"l2err_sincos301", EXTFEM_II_L2ERR_SINCOS301}, {"l2err_exp108", EXTFEM_II_L2ERR_EXP108}};
    if (name)
        for (auto &t : tab) if (!strcmp(t.n, name)) return t.id;
    set_error(nullptr, EXTFEM_ERR_UNREGISTERED_KERNEL, std::string("kernel '") + (name ? name : "(null)") + "' is not registered");
    return EXTFEM_ERR_UNREGISTERED_KERNEL;
}

int extfem_synchronize(extfem_ctx *ctx)
{
    CTX_GUARD(ctx);
    EXTFEM_CUDA_CHECK(C, cudaStreamSynchronize(C->stream));
    return EXTFEM_OK;
}

int extfem_set_option(extfem_ctx *ctx, const char *key, int value)
{
    CTX_GUARD(ctx);
    if (key && !strcmp(key, "fastpath")) { C->fast_enabled = value != 0; return EXTFEM_OK; }
    if (key && !strcmp(key, "fastpath_closed_form")) { C->bary_enabled = value != 0; return EXTFEM_OK; }
    if (key && !strcmp(key, "nonlinear_v2")) { C->nl_version = value != 0 ? 3 : 1; return EXTFEM_OK; }
    if (key && !strcmp(key, "nonlinear_cells_per_warp")) { C->nl_cpw = std::min(std::max(value, 1), 1024); return EXTFEM_OK; }
    if (key && !strcmp(key, "nonlinear_warps_per_cta")) { C->nl_warps = std::min(std::max(value, 1), 4); return EXTFEM_OK; }
    if (key && !strcmp(key, "nonlinear_sparse_jacobian")) { C->nl_sparse_jac = value != 0; return EXTFEM_OK; }
    if (key && !strcmp(key, "nonlinear_point_cache")) { C->nl_point_cache = value != 0; return EXTFEM_OK; }
    if (key && !strcmp(key, "nonlinear_rowwise")) { C->nl_rowwise = value != 0; return EXTFEM_OK; }
    if (key && !strcmp(key, "gather_warp")) { C->gather_warp = value != 0; return EXTFEM_OK; }
    if (key && !strcmp(key, "nonlinear_kernel")) { C->nl_version = std::min(std::max(value, 1), 4); return EXTFEM_OK; }
    if (key && !strcmp(key, "fastpath_templates")) { C->tmpl_enabled = value != 0; return EXTFEM_OK; }
    if (key && !strcmp(key, "template_prefetch_ctas")) { C->tmpl_ahead = value < 0 ? 0 : value; return EXTFEM_OK; }
    if (key && !strcmp(key, "template_walk_roles")) { C->tmpl_walk_roles = std::max(value, 0); return EXTFEM_OK; }
    if (key && !strcmp(key, "template_persistent")) { C->tmpl_persistent = std::min(std::max(value, 0), 16); return EXTFEM_OK; }
    if (key && !strcmp(key, "template_pool_bytes")) { C->tmpl_pool = std::min(std::max(value, 4096), 200 * 1024); return EXTFEM_OK; }
    if (key && !strcmp(key, "template_jit")) { C->jit_enabled = value != 0; return EXTFEM_OK; }
    if (key && !strcmp(key, "template_jit_min_cols")) { C->jit_mincols = value < 0 ? 0 : value; return EXTFEM_OK; }
    if (key && !strcmp(key, "template_plane_mask")) { C->tmpl_planemask = value != 0; return EXTFEM_OK; }
    if (key && !strcmp(key, "template_constant_memory")) { C->tmpl_const = value != 0; return EXTFEM_OK; }
    if (key && !strcmp(key, "template_walk")) { C->tmpl_walk = value != 0; return EXTFEM_OK; }
    if (key && !strcmp(key, "template_permute_mesh")) { C->tmpl_permute_mesh = value != 0; return EXTFEM_OK; }
    if (key && !strcmp(key, "template_min_cols")) { C->tmpl_mincols = value < 1 ? 1 : value; return EXTFEM_OK; }
    if (key && !strcmp(key, "template_class_mask")) { C->tmpl_classmask = value & 3; return EXTFEM_OK; }
    if (key && !strcmp(key, "template_flat_writeout")) { C->tmpl_flat = std::max(value, 0); return EXTFEM_OK; }
    if (key && !strcmp(key, "rhs_local")) { C->rhs_local = value != 0; return EXTFEM_OK; }
    if (key && !strcmp(key, "rhs_prefetch_warps")) { C->rhs_ahead = value; return EXTFEM_OK; }
    if (key && !strcmp(key, "rhs_fast_trig")) { C->rhs_fast_trig = value != 0; return EXTFEM_OK; }
    if (key && !strcmp(key, "rhs_gather_ctas")) { C->rhs_gather_ctas = value; return EXTFEM_OK; }
    if (key && !strcmp(key, "rhs_groups")) { C->rhs_groups = std::min(std::max(value, 1), 4); return EXTFEM_OK; }
    return fail(C, EXTFEM_ERR_BAD_ARGUMENT, std::string("unknown option ") + (key ? key : "(null)"));
}

int64_t extfem_launch_count(extfem_ctx *ctx) { return ctx ? reinterpret_cast<Ctx *>(ctx)->launches : 0; }

int extfem_last_timings(extfem_ctx *ctx, double *ms3)
{
    CTX_GUARD(ctx);
    if (int rc = collect_timing(C)) return rc;
    for (int i = 0; i < 3; ++i) ms3[i] = C->last_ms[i];
    return EXTFEM_OK;
}

int extfem_event_record(extfem_ctx *ctx, int slot)
{
    CTX_GUARD(ctx);
    if (slot < 0 || slot >= 16) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "event slot out of range");
    EXTFEM_CUDA_CHECK(C, cudaEventRecord(C->uev[slot], C->stream));
    return EXTFEM_OK;
}

int extfem_event_elapsed_ms(extfem_ctx *ctx, int a, int b, double *ms)
{
    CTX_GUARD(ctx);
    if (a < 0 || a >= 16 || b < 0 || b >= 16 || !ms) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "event slot out of range");
    EXTFEM_CUDA_CHECK(C, cudaEventSynchronize(C->uev[b]));
    float f = 0;
    EXTFEM_CUDA_CHECK(C, cudaEventElapsedTime(&f, C->uev[a], C->uev[b]));
    *ms = f;
    return EXTFEM_OK;
}

int extfem_mesh_set(extfem_ctx *ctx, int dim, int64_t ncells, int64_t nnodes, const double *coords, const void *cellnodes,
                    int index_bytes, const int32_t *cellregions, const double *cellvolumes, int *mesh_out)
{
    CTX_GUARD(ctx);
    if (dim < 1 || dim > 3 || ncells <= 0 || nnodes <= 0 || !coords || !cellnodes || !mesh_out || (index_bytes != 4 && index_bytes != 8))
        return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "extfem_mesh_set: bad argument");
    if (ncells >= (1ll << 31) || nnodes >= (1ll << 31)) return fail(C, EXTFEM_ERR_CAPACITY, "mesh too large for 32-bit device indices");
    auto M = std::make_unique<Mesh>();
    M->dim = dim; M->tdim = dim; M->ncells = ncells; M->nnodes = nnodes;
    if (int rc = upload(C, M->coords, coords, (size_t)nnodes * dim * 8)) return rc;
    if (int rc = upload_indices(C, M->cellnodes, cellnodes, index_bytes, ncells * (dim + 1), nnodes, "cellnodes")) return rc;
    if (cellregions) { if (int rc = upload(C, M->regions, cellregions, (size_t)ncells * 4)) return rc; }
    else {
        std::vector<int> ones((size_t)ncells, 1);
        if (int rc = upload(C, M->regions, ones.data(), (size_t)ncells * 4)) return rc;
        EXTFEM_CUDA_CHECK(C, cudaStreamSynchronize(C->stream));
    }
    if (int rc = ensure(C, M->vol, (size_t)ncells * 8)) return rc;
    if (cellvolumes) { EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(M->vol.p, cellvolumes, (size_t)ncells * 8, cudaMemcpyDefault, C->stream)); }
    else {
        if (int rc = launch_cell_volumes(C->stream, dim, ncells, M->coords.as<double>(), M->cellnodes.as<int>(), M->vol.as<double>()))
            return fail(C, EXTFEM_ERR_CUDA, "cell volume kernel failed");
        LAUNCHED(C);
    }
    EXTFEM_CUDA_CHECK(C, cudaStreamSynchronize(C->stream));
    C->meshes.push_back(std::move(M));
    *mesh_out = (int)C->meshes.size() - 1;
    return EXTFEM_OK;
}

int extfem_mesh_update_coords(extfem_ctx *ctx, int mesh, const double *coords, const double *cellvolumes)
{
    CTX_GUARD(ctx);
    if (mesh < 0 || mesh >= (int)C->meshes.size() || !coords) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "invalid mesh handle");
    Mesh &M = *C->meshes[mesh];
    EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(M.coords.p, coords, (size_t)M.nnodes * M.dim * 8, cudaMemcpyDefault, C->stream));
    if (cellvolumes) { EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(M.vol.p, cellvolumes, (size_t)M.ncells * 8, cudaMemcpyDefault, C->stream)); }
    else {
        launch_cell_volumes(C->stream, M.dim, M.ncells, M.coords.as<double>(), M.cellnodes.as<int>(), M.vol.as<double>());
        LAUNCHED(C);
    }
    ++M.vol_version;
    if (M.bf && !M.bf_vol_given) {
        face_volumes_kernel<<<nblocks(M.bf->ncells, 256), 256, 0, C->stream>>>(M.dim, M.bf->ncells, M.coords.as<double>(), M.bf->cellnodes.as<int>(),
                                                                              M.bf->vol.as<double>());
        LAUNCHED(C);
    }
    return EXTFEM_OK;
}

int extfem_mesh_set_bfaces(extfem_ctx *ctx, int mesh, int64_t nbfaces, const void *bfacenodes, int index_bytes, const int32_t *bfaceregions,
                           const double *bfacevolumes)
{
    CTX_GUARD(ctx);
    if (mesh < 0 || mesh >= (int)C->meshes.size() || nbfaces <= 0 || !bfacenodes || (index_bytes != 4 && index_bytes != 8))
        return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "extfem_mesh_set_bfaces: bad argument");
    Mesh &M = *C->meshes[mesh];
    auto B = std::make_unique<Mesh>();
    B->dim = M.dim; B->tdim = M.dim - 1; B->ncells = nbfaces; B->nnodes = M.nnodes; B->parent = &M;
    if (int rc = upload_indices(C, B->cellnodes, bfacenodes, index_bytes, nbfaces * M.dim, M.nnodes, "bfacenodes")) return rc;
    if (bfaceregions) { if (int rc = upload(C, B->regions, bfaceregions, (size_t)nbfaces * 4)) return rc; }
    else {
        std::vector<int> ones((size_t)nbfaces, 1);
        if (int rc = upload(C, B->regions, ones.data(), (size_t)nbfaces * 4)) return rc;
        EXTFEM_CUDA_CHECK(C, cudaStreamSynchronize(C->stream));
    }
    if (int rc = ensure(C, B->vol, (size_t)nbfaces * 8)) return rc;
    M.bf_vol_given = bfacevolumes != nullptr;
    if (bfacevolumes) { EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(B->vol.p, bfacevolumes, (size_t)nbfaces * 8, cudaMemcpyDefault, C->stream)); }
    else {
        face_volumes_kernel<<<nblocks(nbfaces, 256), 256, 0, C->stream>>>(M.dim, nbfaces, M.coords.as<double>(), B->cellnodes.as<int>(), B->vol.as<double>());
        LAUNCHED(C);
    }
    EXTFEM_CUDA_CHECK(C, cudaStreamSynchronize(C->stream));
    M.bf = std::move(B);
    // spaces created before the boundary faces keep their BFaceDofs handle only if the face count is unchanged
    for (auto &S : C->spaces)
        if (S && S->mesh == mesh && S->bf && S->bf->celldofs.bytes != (size_t)nbfaces * S->bf->nd * 4) S->bf.reset();
    return EXTFEM_OK;
}

int extfem_space_set(extfem_ctx *ctx, int mesh, int fetype, int ncomp, const void *celldofs, int index_bytes, int ndofs4cell,
                     int64_t ndofs, int *space_out)
{
    CTX_GUARD(ctx);
    if (mesh < 0 || mesh >= (int)C->meshes.size() || !celldofs || !space_out || ncomp < 1 || (index_bytes != 4 && index_bytes != 8))
        return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "extfem_space_set: bad argument");
    Mesh &M = *C->meshes[mesh];
    int order = fetype == EXTFEM_FE_H1P1 ? 1 : (fetype == EXTFEM_FE_H1P2 ? 2 : (fetype == EXTFEM_FE_TABULATED ? 0 : -1));
    if (order < 0)
        return fail(C, EXTFEM_ERR_UNSUPPORTED_ELEMENT, "fetype " + std::to_string(fetype) +
                                                        " is not supported (H1P1, H1P2 / H1Pk order<=2, or EXTFEM_FE_TABULATED + extfem_space_set_tables)");
    int ns = fetype == EXTFEM_FE_TABULATED ? ndofs4cell / ncomp : nscalar_of(order, M.dim);
    if (ndofs4cell != ns * ncomp || ns < 1)
        return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "ndofs4cell does not match the element (expected " + std::to_string(ns * ncomp) + ")");
    if (ndofs4cell > 255) return fail(C, EXTFEM_ERR_CAPACITY, "ndofs4cell > 255");
    if (ndofs >= (1ll << 31)) return fail(C, EXTFEM_ERR_CAPACITY, "too many dofs for 32-bit device indices");
    auto S = std::make_unique<Space>();
    S->mesh = mesh; S->fetype = fetype; S->order = order; S->ncomp = ncomp; S->nscalar = ns; S->nd = ndofs4cell; S->ndofs = ndofs;
    if (int rc = upload_indices(C, S->celldofs, celldofs, index_bytes, M.ncells * ndofs4cell, ndofs, "celldofs")) return rc;
    if (int rc = build_adjacency(C, *S, M.ncells)) return rc;
    S->id = (int)C->spaces.size();
    C->spaces.push_back(std::move(S));
    *space_out = (int)C->spaces.size() - 1;
    return EXTFEM_OK;
}

int extfem_space_set_tables(extfem_ctx *ctx, int space, int order, int nscalar, const double *coeffs, int nscalar_bface, const double *bface_coeffs)
{
    CTX_GUARD(ctx);
    if (space < 0 || space >= (int)C->spaces.size() || !coeffs) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "extfem_space_set_tables: bad argument");
    Space &S = *C->spaces[space];
    if (S.fetype != EXTFEM_FE_TABULATED) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "extfem_space_set_tables: the space is not EXTFEM_FE_TABULATED");
    if (order < 0 || order > 8) return fail(C, EXTFEM_ERR_CAPACITY, "extfem_space_set_tables: polynomial order must be 0..8");
    if (nscalar != S.nscalar) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "extfem_space_set_tables: nscalar does not match ndofs4cell / ncomp");
    Mesh &M = *C->meshes[S.mesh];
    S.order = order;
    S.poly.assign(coeffs, coeffs + (size_t)nscalar * nmonomials(order, M.dim));
    S.tabkey = 2 * S.id;
    S.poly_ready = true;
    S.nscalar_bf = 0; S.bpoly.clear();
    if (bface_coeffs && nscalar_bface > 0) {
        S.nscalar_bf = nscalar_bface;
        S.bpoly.assign(bface_coeffs, bface_coeffs + (size_t)nscalar_bface * nmonomials(order, M.dim - 1));
    }
    if (S.bf) {   // BFaceDofs attached before the basis
        S.bf->order = order; S.bf->poly = S.bpoly; S.bf->poly_ready = !S.bpoly.empty() && S.bf->nscalar == S.nscalar_bf;
    }
    // cached tables of an earlier basis of this space are stale
    for (auto it = C->tables.begin(); it != C->tables.end();)
        if (it->first.space == 2 * S.id || it->first.space == 2 * S.id + 1) it = C->tables.erase(it); else ++it;
    return EXTFEM_OK;
}

int extfem_space_set_bfacedofs(extfem_ctx *ctx, int space, const void *bfacedofs, int index_bytes, int ndofs4bface)
{
    CTX_GUARD(ctx);
    if (space < 0 || space >= (int)C->spaces.size() || !bfacedofs || (index_bytes != 4 && index_bytes != 8) || ndofs4bface < 1)
        return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "extfem_space_set_bfacedofs: bad argument");
    Space &S = *C->spaces[space];
    Mesh &M = *C->meshes[S.mesh];
    if (!M.bf) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "extfem_space_set_bfacedofs: call extfem_mesh_set_bfaces on the grid first");
    const bool tab = S.fetype == EXTFEM_FE_TABULATED;
    const int ns = tab ? ndofs4bface / S.ncomp : nscalar_of(S.order, M.dim - 1);
    if (ndofs4bface != ns * S.ncomp || ndofs4bface > 255)
        return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "ndofs4bface does not match the element (expected " + std::to_string(ns * S.ncomp) + ")");
    auto B = std::make_unique<Space>();
    B->id = S.id; B->mesh = S.mesh; B->fetype = S.fetype; B->order = S.order; B->ncomp = S.ncomp; B->nscalar = ns; B->nd = ndofs4bface; B->ndofs = S.ndofs;
    B->tabkey = tab ? 2 * S.id + 1 : -1;
    if (tab) { B->poly = S.bpoly; B->poly_ready = S.poly_ready && !S.bpoly.empty() && S.nscalar_bf == ns; }
    if (int rc = upload_indices(C, B->celldofs, bfacedofs, index_bytes, M.bf->ncells * ndofs4bface, S.ndofs, "bfacedofs")) return rc;
    if (int rc = build_adjacency(C, *B, M.bf->ncells)) return rc;
    S.bf = std::move(B);
    return EXTFEM_OK;
}

int extfem_pattern_build(extfem_ctx *ctx, int nrow, const int *rowspaces, int ncol, const int *colspaces,
                         const uint8_t *block_coupling, int *pattern_out)
{
    CTX_GUARD(ctx);
    if (nrow < 1 || ncol < 1 || nrow > MAXBLOCKS || ncol > MAXBLOCKS || !rowspaces || !colspaces || !pattern_out)
        return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "extfem_pattern_build: bad argument");
    auto Pp = std::make_unique<Pattern>();
    Pattern &P = *Pp;
    int mesh = -1;
    P.rowoff.push_back(0); P.coloff.push_back(0);
    for (int r = 0; r < nrow; ++r) {
        int s = rowspaces[r];
        if (s < 0 || s >= (int)C->spaces.size()) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "invalid row space handle");
        if (mesh < 0) mesh = C->spaces[s]->mesh;
        if (C->spaces[s]->mesh != mesh) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "all spaces of a pattern must live on one mesh");
        P.rowspaces.push_back(s);
        P.rowlocoff.push_back(P.NRpat);
        P.NRpat += C->spaces[s]->nd;
        P.rowoff.push_back(P.rowoff.back() + C->spaces[s]->ndofs);
    }
    for (int c = 0; c < ncol; ++c) {
        int s = colspaces[c];
        if (s < 0 || s >= (int)C->spaces.size()) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "invalid column space handle");
        if (C->spaces[s]->mesh != mesh) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "all spaces of a pattern must live on one mesh");
        P.colspaces.push_back(s);
        P.coloff.push_back(P.coloff.back() + C->spaces[s]->ndofs);
    }
    P.nrows = P.rowoff.back(); P.ncols = P.coloff.back();
    if (P.nrows >= (1ll << 31) - 1) return fail(C, EXTFEM_ERR_CAPACITY, "too many rows for 32-bit row indices");
    P.square = (P.rowspaces == P.colspaces);
    P.coupling.assign((size_t)ncol * nrow, 1);
    if (block_coupling) for (int i = 0; i < ncol * nrow; ++i) P.coupling[i] = block_coupling[i] ? 1 : 0;
    Mesh &M = *C->meshes[mesh];
    (void)M;

    DevBuf collen, err, tmp;
    if (int rc = ensure(C, collen, (P.ncols + 1) * 8)) return rc;
    if (int rc = ensure(C, err, 4)) return rc;
    EXTFEM_CUDA_CHECK(C, cudaMemsetAsync(collen.p, 0, (P.ncols + 1) * 8, C->stream));
    EXTFEM_CUDA_CHECK(C, cudaMemsetAsync(err.p, 0, 4, C->stream));
    std::vector<PatArgs> pargs(ncol);
    for (int c = 0; c < ncol; ++c) {
        Space &Sc = *C->spaces[P.colspaces[c]];
        PatArgs &A = pargs[c];
        memset(&A, 0, sizeof(A));
        A.ncolsb = Sc.ndofs; A.colbase = P.coloff[c];
        A.adjptr = Sc.adjptr.as<long long>(); A.adjcell = Sc.adjcell.as<int>();
        A.nrb = nrow; A.NRpat = P.NRpat; A.error = err.as<int>();
        for (int r = 0; r < nrow; ++r) {
            Space &Sr = *C->spaces[P.rowspaces[r]];
            A.coupled[r] = P.coupling[(size_t)c * nrow + r];
            A.celldofs[r] = Sr.celldofs.as<int>(); A.nd[r] = Sr.nd; A.rowoff[r] = P.rowoff[r]; A.rowlocoff[r] = P.rowlocoff[r];
        }
        pattern_count_kernel<<<nblocks(Sc.ndofs, PAT_WARPS), PAT_WARPS * 32, 0, C->stream>>>(A, collen.as<long long>());
        LAUNCHED(C);
    }
    EXTFEM_CUDA_CHECK(C, cudaGetLastError());
    int herr = 0;
    EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(&herr, err.p, 4, cudaMemcpyDeviceToHost, C->stream));
    EXTFEM_CUDA_CHECK(C, cudaStreamSynchronize(C->stream));
    if (herr) return fail(C, EXTFEM_ERR_CAPACITY, "a matrix column has more than " + std::to_string(PAT_CAP) + " candidate rows");
    // max column length -> width of the position map
    {
        DevBuf dmax;
        if (int rc = ensure(C, dmax, 8)) return rc;
        size_t tb = 0;
        cub::DeviceReduce::Max(nullptr, tb, collen.as<long long>(), dmax.as<long long>(), (long long)P.ncols, C->stream);
        if (int rc = ensure(C, tmp, tb)) return rc;
        EXTFEM_CUDA_CHECK(C, cub::DeviceReduce::Max(tmp.p, tb, collen.as<long long>(), dmax.as<long long>(), (long long)P.ncols, C->stream));
        LAUNCHED(C);
        long long hmax = 0;
        EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(&hmax, dmax.p, 8, cudaMemcpyDeviceToHost, C->stream));
        EXTFEM_CUDA_CHECK(C, cudaStreamSynchronize(C->stream));
        P.maxcollen = (int)hmax;
        P.poswidth = hmax <= 254 ? 1 : 2;
        if (hmax > 65534 || hmax > GATHER_MAXNNZ) return fail(C, EXTFEM_ERR_CAPACITY, "matrix column too long");
    }
    if (int rc = ensure(C, P.colptr, (P.ncols + 1) * 8)) return rc;
    {
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, collen.as<long long>(), P.colptr.as<long long>(), (long long)(P.ncols + 1), C->stream);
        DevBuf t2;
        if (int rc = ensure(C, t2, tb)) return rc;
        EXTFEM_CUDA_CHECK(C, cub::DeviceScan::ExclusiveSum(t2.p, tb, collen.as<long long>(), P.colptr.as<long long>(),
                                                           (long long)(P.ncols + 1), C->stream));
        LAUNCHED(C);
        EXTFEM_CUDA_CHECK(C, cudaStreamSynchronize(C->stream));
    }
    std::vector<long long> hcolptr((size_t)P.ncols + 1);
    EXTFEM_CUDA_CHECK(C, cudaMemcpy(hcolptr.data(), P.colptr.p, (P.ncols + 1) * 8, cudaMemcpyDeviceToHost));
    P.nnz = hcolptr.back();
    if (int rc = ensure(C, P.rowval, (size_t)P.nnz * 4)) return rc;
    if (int rc = ensure(C, P.nzval, (size_t)P.nnz * 8)) return rc;
    if (int rc = ensure(C, P.b, (size_t)P.nrows * 8)) return rc;
    EXTFEM_CUDA_CHECK(C, cudaMemsetAsync(P.nzval.p, 0, (size_t)P.nnz * 8, C->stream));
    EXTFEM_CUDA_CHECK(C, cudaMemsetAsync(P.b.p, 0, (size_t)P.nrows * 8, C->stream));
    for (int c = 0; c < ncol; ++c) {
        Space &Sc = *C->spaces[P.colspaces[c]];
        long long npairs = M.ncells * Sc.nd;
        P.posmap.push_back(std::make_unique<DevBuf>());
        if (int rc = ensure(C, *P.posmap.back(), (size_t)npairs * P.NRpat * P.poswidth)) return rc;
        if (P.poswidth == 1)
            pattern_fill_kernel<unsigned char><<<nblocks(Sc.ndofs, PAT_WARPS), PAT_WARPS * 32, 0, C->stream>>>(
                pargs[c], P.colptr.as<long long>(), P.rowval.as<int>(), P.posmap.back()->as<unsigned char>());
        else
            pattern_fill_kernel<unsigned short><<<nblocks(Sc.ndofs, PAT_WARPS), PAT_WARPS * 32, 0, C->stream>>>(
                pargs[c], P.colptr.as<long long>(), P.rowval.as<int>(), P.posmap.back()->as<unsigned short>());
        LAUNCHED(C);
    }
    EXTFEM_CUDA_CHECK(C, cudaGetLastError());
    // column chunks of the gather kernel: <= GATHER_THREADS columns, <= GATHER_MAXNNZ entries, one column block
    std::vector<int> chunks;
    chunks.push_back(0);
    for (int c = 0; c < ncol; ++c) {
        long long k = P.coloff[c], kend = P.coloff[c + 1];
        while (k < kend) {
            long long k2 = k;
            long long base = hcolptr[k];
            while (k2 < kend && k2 - k < GATHER_THREADS && hcolptr[k2 + 1] - base <= GATHER_MAXNNZ) ++k2;
            if (k2 == k) return fail(C, EXTFEM_ERR_CAPACITY, "matrix column too long for the gather kernel");
            chunks.push_back((int)k2);
            k = k2;
        }
    }
    P.nchunks = (int)chunks.size() - 1;
    P.hcolptr = std::move(hcolptr);
    P.fastplans.resize(ncol);
    P.tplans.resize(ncol);
    if (int rc = upload(C, P.chunkptr, chunks.data(), chunks.size() * 4)) return rc;
    EXTFEM_CUDA_CHECK(C, cudaStreamSynchronize(C->stream));
    C->patterns.push_back(std::move(Pp));
    *pattern_out = (int)C->patterns.size() - 1;
    return EXTFEM_OK;
}

int extfem_pattern_dims(extfem_ctx *ctx, int pattern, int64_t *nrows, int64_t *ncols, int64_t *nnz)
{
    CTX_GUARD(ctx);
    GET_PATTERN(pattern);
    if (nrows) *nrows = P.nrows;
    if (ncols) *ncols = P.ncols;
    if (nnz) *nnz = P.nnz;
    return EXTFEM_OK;
}

int extfem_pattern_get(extfem_ctx *ctx, int pattern, int64_t *colptr, int64_t *rowval)
{
    CTX_GUARD(ctx);
    GET_PATTERN(pattern);
    DevBuf c64, r64;
    if (colptr) if (int rc = ensure(C, c64, (P.ncols + 1) * 8)) return rc;
    if (rowval) if (int rc = ensure(C, r64, (size_t)P.nnz * 8)) return rc;
    long long n = std::max(P.ncols + 1, P.nnz);
    csc_export_kernel<<<nblocks(n, 256), 256, 0, C->stream>>>(P.colptr.as<long long>(), P.ncols + 1, P.rowval.as<int>(), P.nnz,
                                                              colptr ? c64.as<long long>() : nullptr, rowval ? r64.as<long long>() : nullptr);
    LAUNCHED(C);
    if (colptr) EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(colptr, c64.p, (P.ncols + 1) * 8, cudaMemcpyDefault, C->stream));
    if (rowval) EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(rowval, r64.p, (size_t)P.nnz * 8, cudaMemcpyDefault, C->stream));
    EXTFEM_CUDA_CHECK(C, cudaStreamSynchronize(C->stream));
    return EXTFEM_OK;
}

int extfem_assemble_bilinear(extfem_ctx *ctx, int pattern, const extfem_opdesc *d, const double *sol, int accumulate, double *nzval_out)
{
    CTX_GUARD(ctx);
    GET_PATTERN(pattern);
    Prepared R;
    if (int rc = prepare(C, P, d, KIND_BILINEAR, d && d->nargs > 0 ? sol : nullptr, R)) return rc;
    if (R.entities == EXTFEM_ON_BFACES) {
        if (int rc = assemble_bfaces(C, P, R, d, KIND_BILINEAR, accumulate)) return rc;
        if (nzval_out) EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(nzval_out, P.nzval.p, (size_t)P.nnz * 8, cudaMemcpyDefault, C->stream));
        return finish_timing(C, nzval_out != nullptr);
    }
    EXTFEM_CUDA_CHECK(C, cudaEventRecord(C->ev[0], C->stream));
    bool fast = false;
    if (int rc = try_fast_bilinear(C, P, R, d, accumulate, &fast)) return rc;
    if (!fast) {
        const OpDev &op = R.op;
        if (int rc = ensure(C, C->loc, (size_t)op.ncells * op.NR * op.NC * 8)) return rc;
        int rcd = dispatch_dim(C, op.dim, [&](auto dimc) {
            constexpr int DIM = decltype(dimc)::value;
            size_t per_cell = sizeof(CellGeo<DIM>) + (op.nargs > 0 ? (size_t)op.nq * op.nin * 8 : 0);
            int cpb = cells_per_block(per_cell);
            local_bilinear_kernel<DIM><<<nblocks(op.ncells, cpb), 256, cpb * per_cell, C->stream>>>(op, C->loc.as<double>(), cpb);
        });
        if (rcd) return rcd;
        LAUNCHED(C);
        EXTFEM_CUDA_CHECK(C, cudaGetLastError());
        EXTFEM_CUDA_CHECK(C, cudaEventRecord(C->ev[1], C->stream));
        if (int rc = gather_matrix(C, P, R, accumulate, d->transposed_copy)) return rc;
        EXTFEM_CUDA_CHECK(C, cudaEventRecord(C->ev[2], C->stream));
    }
    if (nzval_out) EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(nzval_out, P.nzval.p, (size_t)P.nnz * 8, cudaMemcpyDefault, C->stream));
    return finish_timing(C, nzval_out != nullptr);
}

int extfem_assemble_linear(extfem_ctx *ctx, int pattern, const extfem_opdesc *d, const double *sol, int accumulate, double *b_out)
{
    CTX_GUARD_NOJOIN(ctx);
    GET_PATTERN(pattern);
    Prepared R;
    if (int rc = prepare(C, P, d, KIND_LINEAR, d && d->nargs > 0 ? sol : nullptr, R)) return rc;
    const OpDev &op = R.op;
    if (R.entities == EXTFEM_ON_BFACES) {
        if (int rc = assemble_bfaces(C, P, R, d, KIND_LINEAR, accumulate)) return rc;
        if (b_out) EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(b_out, P.b.p, (size_t)P.nrows * 8, cudaMemcpyDefault, C->stream));
        return finish_timing(C, b_out != nullptr);
    }
    EXTFEM_CUDA_CHECK(C, cudaEventRecord(C->ev[0], C->stream));
    bool fast = false;
    if (int rc = try_fast_linear(C, P, R, d, accumulate, &fast)) return rc;
    if (!fast) {
        if (int rc = ensure(C, C->bloc, (size_t)op.ncells * op.NR * 8)) return rc;
        int rcd = dispatch_dim(C, op.dim, [&](auto dimc) {
            constexpr int DIM = decltype(dimc)::value;
            size_t per_cell = sizeof(CellGeo<DIM>) + (size_t)op.nq * op.nout * 8;
            int cpb = cells_per_block(per_cell);
            local_linear_kernel<DIM><<<nblocks(op.ncells, cpb), 256, cpb * per_cell, C->stream>>>(op, C->bloc.as<double>(), cpb);
        });
        if (rcd) return rcd;
        LAUNCHED(C);
        EXTFEM_CUDA_CHECK(C, cudaGetLastError());
        EXTFEM_CUDA_CHECK(C, cudaEventRecord(C->ev[1], C->stream));
        if (int rc = gather_vector(C, P, R, accumulate)) return rc;
        EXTFEM_CUDA_CHECK(C, cudaEventRecord(C->ev[2], C->stream));
    }
    if (b_out) EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(b_out, P.b.p, (size_t)P.nrows * 8, cudaMemcpyDefault, C->stream));
    return finish_timing(C, b_out != nullptr);
}

/* 1: the templates of the block run as plan-time specialised kernels (jit.cuh); 0: not (disabled, below the size
 * threshold, not a closed-form operator, or not assembled yet); -1: specialisation was tried and failed (static kernel) */
int extfem_plan_jit_status(extfem_ctx *ctx, int pattern, int block)
{
    CTX_GUARD(ctx);
    GET_PATTERN(pattern);
    if (block < 0 || block >= (int)P.tplans.size()) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "extfem_plan_jit_status: bad block");
    if (!P.tplans[block] || !P.tplans[block]->jit || !P.tplans[block]->jit->tried) return 0;
    return P.tplans[block]->jit->ok ? 1 : -1;
}

/* statistics of the fast-path plans of column block `block` (built on first use):
 * [0] period P of the geometry order, [1] templates, [2] template warps, [3] columns on the record kernel,
 * [4] CTAs of the template kernel, [5] shared-memory pool bytes, [6] template rounds, [7] columns of the block */
int extfem_plan_stats(extfem_ctx *ctx, int pattern, int block, int64_t *stats8)
{
    CTX_GUARD(ctx);
    GET_PATTERN(pattern);
    if (block < 0 || block >= (int)P.tplans.size() || !stats8) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "extfem_plan_stats: bad argument");
    for (int i = 0; i < 8; ++i) stats8[i] = 0;
    if (!P.tplans[block] || !P.tplans[block]->ready) return EXTFEM_OK;
    TemplatePlan &T = *P.tplans[block];
    stats8[0] = T.Lg.P; stats8[1] = T.ntemplates; stats8[2] = T.nwarps; stats8[3] = T.nleft; stats8[4] = T.nctas;
    stats8[5] = T.pool_bytes; stats8[6] = T.nrounds; stats8[7] = T.ncols;
    return EXTFEM_OK;
}

int extfem_assemble_nonlinear(extfem_ctx *ctx, int pattern, const extfem_opdesc *d, const double *sol, int accumulate,
                              double *nzval_out, double *b_out)
{
    CTX_GUARD(ctx);
    GET_PATTERN(pattern);
    Prepared R;
    if (int rc = prepare(C, P, d, KIND_NONLINEAR, sol, R)) return rc;
    const OpDev &op = R.op;
    if (int rc = ensure(C, C->loc, (size_t)op.ncells * op.NR * op.NC * 8)) return rc;
    if (int rc = ensure(C, C->bloc, (size_t)op.ncells * op.NR * 8)) return rc;
    EXTFEM_CUDA_CHECK(C, cudaEventRecord(C->ev[0], C->stream));
    // combined operator vectors of v2 hold <= CVC_MAX entries per dof (all arguments on the dof's block together)
    auto cv_entries = [&](const ArgDev *a, int n) {
        int worst = 0;
        for (int i = 0; i < n; ++i) {
            int sum = 0;
            for (int j = 0; j < n; ++j)
                if (a[j].locoff == a[i].locoff) sum += a[j].op == EXTFEM_OP_ID ? 1 : (a[j].op == EXTFEM_OP_GRAD ? op.dim : (a[j].op == EXTFEM_OP_DIV ? 1 : op.dim));
            worst = std::max(worst, sum);
        }
        return worst;
    };
    const bool v2 = C->nl_version >= 2 && cv_entries(op.args, op.nargs) <= CVC_MAX && cv_entries(op.test, op.ntest) <= CVC_MAX && op.nin + op.nout <= 255;
    NL3Tables T3;
    bool v3 = false;
    // the tensor-core path evaluates the kernel per (cell, point) with its vectors in registers: instantiated input lengths
    const bool nl4_points_ok = op.nin == op.nout && (op.nin == 2 || op.nin == 3 || op.nin == 4 || op.nin == 7 || op.nin == 9);
    if (C->nl_version >= 4 && nl4_points_ok)
        if (int rc = ensure(C, C->nlpt, (size_t)op.ncells * op.nq * (op.nin * op.nout + op.nout) * 8)) return rc;
    if (C->nl_version >= 3) if (int rc = build_nl3_tables(C, op, T3, &v3, C->nl_version >= 4 && nl4_points_ok)) return rc;
    int rcd = dispatch_dim(C, op.dim, [&](auto dimc) {
        constexpr int DIM = decltype(dimc)::value;
        if (v3 && C->nl_version >= 4 && T3.dense_ok && nl4_points_ok) {
            const size_t wd = nl4_warp_doubles(op.nq, op.nin, T3.nnzJ, T3.Np4, T3.phi_off[T3.nspaces]);
            const int nw = C->nl_warps, cpw = C->nl_cpw;
            const size_t smem = (size_t)nw * wd * 8 + T3.tab_bytes;
            if (smem <= 220 * 1024) {
                const long long ntot = op.ncells * op.nq;
                double *wJ = C->nlpt.as<double>(), *rqg = wJ + (size_t)ntot * T3.nnzJ;
                // (1) kernel value and Jacobian per (cell, point), one thread each
                const unsigned gp = nblocks(ntot, 128);
                // physical basis values of a thread cached in shared memory when they fit (kernels_generic.cuh: PVC)
                const int npv = T3.phi_off[T3.nspaces] / op.nq;
                const bool pvc = C->nl_point_cache && npv < 65536 && nl_point_smem(T3.tab_bytes, op.nq, op.NC, T3.EC, npv, true) <= 72 * 1024;
                const size_t psm = nl_point_smem(T3.tab_bytes, op.nq, op.NC, T3.EC, npv, pvc);
                switch (op.nin) {
#define EXTFEM_NLPT2(D, N, R, P) { smem_attr(C, (const void *)nl_point_kernel<D, N, R, P>, 200 * 1024); \
                                   nl_point_kernel<D, N, R, P><<<gp, 128, psm, C->stream>>>(op, T3, wJ, rqg); }
#define EXTFEM_NLPT(N) case N: if (pvc) EXTFEM_NLPT2(DIM, N, false, true) else EXTFEM_NLPT2(DIM, N, false, false) break;
                EXTFEM_NLPT(2) EXTFEM_NLPT(3) EXTFEM_NLPT(4) EXTFEM_NLPT(7)
                case 9:
                    if (DIM == 3 && op.kernel_id == EXTFEM_NL_NEOHOOKE3D && C->nl_rowwise) {   // Jacobian row by row: no spills
                        if (pvc) EXTFEM_NLPT2(3, 9, true, true) else EXTFEM_NLPT2(3, 9, true, false)
                    } else {
                        if (pvc) EXTFEM_NLPT2(DIM, 9, false, true) else EXTFEM_NLPT2(DIM, 9, false, false)
                    }
                    break;
#undef EXTFEM_NLPT
#undef EXTFEM_NLPT2
                }
                ++C->launches;
                // (2) contractions on the FP64 tensor cores, one warp per cell: k-steps / rank-1 remainder by the kernel's vector length
                const unsigned gb = nblocks(op.ncells, nw * cpw);
                const int shape = op.nin <= 4 ? 0 : (op.nin <= 8 ? 1 : 2);
#define EXTFEM_NL4(N, KS, R1) smem_attr(C, (const void *)local_nonlinear_kernel4<DIM, N, KS, R1>, 227 * 1024); \
                      local_nonlinear_kernel4<DIM, N, KS, R1><<<gb, nw * 32, smem, C->stream>>>(op, T3, wJ, rqg, C->loc.as<double>(), C->bloc.as<double>(), cpw);
#define EXTFEM_NL4S(N) case N: if (shape == 0) { EXTFEM_NL4(N, 1, false) } else if (shape == 1) { EXTFEM_NL4(N, 2, false) } else { EXTFEM_NL4(N, 2, true) } break;
                switch (T3.NT4) { EXTFEM_NL4S(1) EXTFEM_NL4S(2) EXTFEM_NL4S(3) EXTFEM_NL4S(4) }
#undef EXTFEM_NL4S
#undef EXTFEM_NL4
                return;
            }
        }
        if (v3 && !(C->nl_version >= 4 && T3.dense_ok && nl4_points_ok)) {   // (with the dense tables the sparse ones of v3 were not built)
            const size_t wd = nl3_warp_doubles(op.nq, op.nin, op.nout, op.NC, op.NR, T3.EC, T3.ER, T3.same, T3.phi_off[T3.nspaces]);
            int nw = 8;
            while (nw > 1 && (size_t)nw * wd * 8 + T3.tab_bytes > 100 * 1024) nw >>= 1;
            const size_t smem = (size_t)nw * wd * 8 + T3.tab_bytes;
            if (smem <= 200 * 1024) {
                smem_attr(C, (const void *)local_nonlinear_kernel3<DIM>, 200 * 1024);
                const int cpw = 4;   // cells per warp: amortises the table load of a block
                local_nonlinear_kernel3<DIM><<<nblocks(op.ncells, nw * cpw), nw * 32, smem, C->stream>>>(op, T3, C->loc.as<double>(),
                                                                                                      C->bloc.as<double>(), cpw);
                return;
            }
        }
        if (v2) {
            const size_t per_cell = nl2_cell_bytes((int)sizeof(CellGeo<DIM>), op.nq, op.nin, op.nout, op.NR, op.NC);
            const int cpb = (int)std::max<size_t>(1, std::min<size_t>(8, (72 * 1024) / per_cell));
            smem_attr(C, (const void *)local_nonlinear_kernel2<DIM>, 200 * 1024);
            local_nonlinear_kernel2<DIM><<<nblocks(op.ncells, cpb), 256, cpb * per_cell, C->stream>>>(op, C->loc.as<double>(),
                                                                                                    C->bloc.as<double>(), cpb);
            return;
        }
        size_t per_cell = sizeof(CellGeo<DIM>) + (size_t)op.nq * (op.nout * op.nin + op.nout) * 8;
        int cpb = cells_per_block(per_cell);
        local_nonlinear_kernel<DIM><<<nblocks(op.ncells, cpb), 256, cpb * per_cell, C->stream>>>(op, C->loc.as<double>(),
                                                                                               C->bloc.as<double>(), cpb);
    });
    if (rcd) return rcd;
    LAUNCHED(C);
    EXTFEM_CUDA_CHECK(C, cudaGetLastError());
    EXTFEM_CUDA_CHECK(C, cudaEventRecord(C->ev[1], C->stream));
    if (int rc = gather_matrix(C, P, R, accumulate, 0)) return rc;
    if (int rc = gather_vector(C, P, R, accumulate)) return rc;
    EXTFEM_CUDA_CHECK(C, cudaEventRecord(C->ev[2], C->stream));
    if (nzval_out) EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(nzval_out, P.nzval.p, (size_t)P.nnz * 8, cudaMemcpyDefault, C->stream));
    if (b_out) EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(b_out, P.b.p, (size_t)P.nrows * 8, cudaMemcpyDefault, C->stream));
    return finish_timing(C, nzval_out != nullptr || b_out != nullptr);
}

int extfem_quadrature_points(extfem_ctx *ctx, int pattern, const extfem_opdesc *d, int is_linear, int *nq_out, double *xq)
{
    CTX_GUARD(ctx);
    GET_PATTERN(pattern);
    Prepared R;
    extfem_opdesc dd = *d;
    if (is_linear && dd.kernel_id == EXTFEM_LIN_TABULATED) dd.kernel_id = EXTFEM_LIN_CONSTANT_ONE; // values not needed here
    if (int rc = prepare(C, P, &dd, is_linear ? KIND_LINEAR : KIND_BILINEAR, nullptr, R)) return rc;
    if (nq_out) *nq_out = R.op.nq;
    if (!xq) return EXTFEM_OK;
    const OpDev &op = R.op;
    DevBuf dx;
    long long n = op.ncells * op.nq;
    if (int rc = ensure(C, dx, (size_t)n * op.dim * 8)) return rc;
    int rcd = dispatch_dim(C, op.dim, [&](auto dimc) {
        constexpr int DIM = decltype(dimc)::value;
        quadpoints_kernel<DIM><<<nblocks(n, 256), 256, 0, C->stream>>>(op, dx.as<double>());
    });
    if (rcd) return rcd;
    LAUNCHED(C);
    EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(xq, dx.p, (size_t)n * op.dim * 8, cudaMemcpyDefault, C->stream));
    EXTFEM_CUDA_CHECK(C, cudaStreamSynchronize(C->stream));
    return EXTFEM_OK;
}

int extfem_integrate(extfem_ctx *ctx, int pattern, const extfem_opdesc *d, const double *sol, int resultdim, int piecewise, double *out)
{
    CTX_GUARD(ctx);
    GET_PATTERN(pattern);
    if (!out || !sol) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "extfem_integrate: sol / out is NULL");
    Prepared R;
    if (int rc = prepare(C, P, d, KIND_INTEGRATE, sol, R)) return rc;
    OpDev &op = R.op;
    if (resultdim <= 0) resultdim = op.nin;     // :resultdim == 0: length of the arguments (item_integrator.jl:188-192)
    if (resultdim > MAXOP) return fail(C, EXTFEM_ERR_CAPACITY, "resultdim larger than MAXOP");
    if ((d->kernel_id == EXTFEM_II_L2ERR_SINCOS301 || d->kernel_id == EXTFEM_II_L2ERR_EXP108) && (resultdim != 1 || op.nin < 1))
        return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "exact-error kernels integrate one scalar argument");
    if (d->kernel_id == EXTFEM_II_L2DIFF_TABULATED) {
        if (!d->tabulated || resultdim > op.nin) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "l2diff_tabulated needs reference values [ncells][nq][resultdim]");
        if (int rc = upload(C, C->tab, d->tabulated, (size_t)op.ncells * op.nq * resultdim * 8)) return rc;
        op.tabulated = C->tab.as<double>();
    }
    DevBuf &vals = C->loc;
    if (int rc = ensure(C, vals, (size_t)op.ncells * resultdim * 8)) return rc;
    EXTFEM_CUDA_CHECK(C, cudaEventRecord(C->ev[0], C->stream));
    int rcd = dispatch_dim(C, op.dim, [&](auto dimc) {
        constexpr int DIM = decltype(dimc)::value;
        item_integrate_kernel<DIM><<<nblocks(op.ncells, 128), 128, 0, C->stream>>>(op, resultdim, vals.as<double>());
    });
    if (rcd) return rcd;
    LAUNCHED(C);
    EXTFEM_CUDA_CHECK(C, cudaGetLastError());
    EXTFEM_CUDA_CHECK(C, cudaEventRecord(C->ev[1], C->stream));
    if (piecewise) {
        EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(out, vals.p, (size_t)op.ncells * resultdim * 8, cudaMemcpyDefault, C->stream));
    } else {
        DevBuf part;
        if (int rc = ensure(C, part, (size_t)(RED_BLOCKS + 1) * resultdim * 8)) return rc;
        ii_reduce_partial_kernel<<<RED_BLOCKS, 256, 0, C->stream>>>(op.ncells, resultdim, vals.as<double>(), part.as<double>());
        ii_reduce_final_kernel<<<1, 256, 0, C->stream>>>(RED_BLOCKS, resultdim, part.as<double>(), part.as<double>() + (size_t)RED_BLOCKS * resultdim);
        LAUNCHED(C); LAUNCHED(C);
        EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(out, part.as<double>() + (size_t)RED_BLOCKS * resultdim, (size_t)resultdim * 8, cudaMemcpyDefault, C->stream));
        EXTFEM_CUDA_CHECK(C, cudaStreamSynchronize(C->stream));   // part is freed on return
    }
    EXTFEM_CUDA_CHECK(C, cudaEventRecord(C->ev[2], C->stream));
    return finish_timing(C, true);
}

int extfem_values_zero(extfem_ctx *ctx, int pattern, int zero_matrix, int zero_rhs)
{
    CTX_GUARD(ctx);
    GET_PATTERN(pattern);
    if (zero_matrix) EXTFEM_CUDA_CHECK(C, cudaMemsetAsync(P.nzval.p, 0, (size_t)P.nnz * 8, C->stream));
    if (zero_rhs) EXTFEM_CUDA_CHECK(C, cudaMemsetAsync(P.b.p, 0, (size_t)P.nrows * 8, C->stream));
    return EXTFEM_OK;
}

int extfem_apply_values(extfem_ctx *ctx, int64_t ndofs, const int64_t *dofs, const double *values, double *sol, int64_t nsol)
{
    CTX_GUARD(ctx);
    if (ndofs == 0) return EXTFEM_OK;
    if (!dofs || !sol || nsol <= 0) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "extfem_apply_values: bad argument");
    cudaPointerAttributes at;
    const bool dev = cudaPointerGetAttributes(&at, sol) == cudaSuccess && (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged);
    cudaGetLastError();
    DevBuf dd, dv, ds, err;
    if (int rc = upload(C, dd, dofs, (size_t)ndofs * 8)) return rc;
    if (values) if (int rc = upload(C, dv, values, (size_t)ndofs * 8)) return rc;
    double *target = sol;
    if (!dev) { if (int rc = upload(C, ds, sol, (size_t)nsol * 8)) return rc; target = ds.as<double>(); }
    if (int rc = ensure(C, err, 4)) return rc;
    EXTFEM_CUDA_CHECK(C, cudaMemsetAsync(err.p, 0, 4, C->stream));
    apply_values_kernel<<<nblocks(ndofs, 256), 256, 0, C->stream>>>(ndofs, dd.as<long long>(), values ? dv.as<double>() : nullptr, target, nsol, err.as<int>());
    LAUNCHED(C);
    if (!dev) EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(sol, target, (size_t)nsol * 8, cudaMemcpyDeviceToHost, C->stream));
    int herr = 0;
    EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(&herr, err.p, 4, cudaMemcpyDeviceToHost, C->stream));
    EXTFEM_CUDA_CHECK(C, cudaStreamSynchronize(C->stream));
    if (herr) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "extfem_apply_values: dof out of range");
    return EXTFEM_OK;
}

int extfem_values_get(extfem_ctx *ctx, int pattern, double *nzval, double *b)
{
    CTX_GUARD(ctx);
    GET_PATTERN(pattern);
    if (nzval) EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(nzval, P.nzval.p, (size_t)P.nnz * 8, cudaMemcpyDefault, C->stream));
    if (b) EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(b, P.b.p, (size_t)P.nrows * 8, cudaMemcpyDefault, C->stream));
    EXTFEM_CUDA_CHECK(C, cudaStreamSynchronize(C->stream));
    return EXTFEM_OK;
}

// suffix starts and packed column pointers of the lower triangle, once per pattern
static int ensure_lower(Ctx *C, Pattern &P)
{
    if (P.nnz_lower >= 0) return EXTFEM_OK;
    if (!P.square || P.nrows != P.ncols) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "the lower triangle needs a square pattern");
    DevBuf cnt, tmp;
    if (int rc = ensure(C, P.lstart, (size_t)P.ncols * 8)) return rc;
    if (int rc = ensure(C, P.lcolptr, (size_t)(P.ncols + 1) * 8)) return rc;
    if (int rc = ensure(C, cnt, (size_t)(P.ncols + 1) * 8)) return rc;
    EXTFEM_CUDA_CHECK(C, cudaMemsetAsync(cnt.p, 0, (size_t)(P.ncols + 1) * 8, C->stream));
    lower_start_kernel<<<nblocks(P.ncols, 256), 256, 0, C->stream>>>(P.ncols, P.colptr.as<long long>(), P.rowval.as<int>(), P.lstart.as<long long>(),
                                                                   cnt.as<long long>());
    LAUNCHED(C);
    size_t tb = 0;
    EXTFEM_CUDA_CHECK(C, cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt.as<long long>(), P.lcolptr.as<long long>(), (int)(P.ncols + 1), C->stream));
    if (int rc = ensure(C, tmp, tb)) return rc;
    EXTFEM_CUDA_CHECK(C, cub::DeviceScan::ExclusiveSum(tmp.p, tb, cnt.as<long long>(), P.lcolptr.as<long long>(), (int)(P.ncols + 1), C->stream));
    LAUNCHED(C);
    long long n = 0;
    EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(&n, P.lcolptr.as<long long>() + P.ncols, 8, cudaMemcpyDeviceToHost, C->stream));
    EXTFEM_CUDA_CHECK(C, cudaStreamSynchronize(C->stream));
    P.nnz_lower = n;
    // column chunks of the pipelined export (extfem_values_get_lower): packed offsets at the chunk ends
    const int nch = n > (1ll << 22) ? 8 : 1;
    P.lchunk_col.assign(nch + 1, 0); P.lchunk_off.assign(nch + 1, 0);
    for (int k = 0; k <= nch; ++k) {
        P.lchunk_col[k] = P.ncols * k / nch;
        EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(&P.lchunk_off[k], P.lcolptr.as<long long>() + P.lchunk_col[k], 8, cudaMemcpyDeviceToHost, C->stream));
    }
    EXTFEM_CUDA_CHECK(C, cudaStreamSynchronize(C->stream));
    return EXTFEM_OK;
}

int extfem_pattern_get_lower(extfem_ctx *ctx, int pattern, int64_t *nnz_lower, int64_t *colptr, int64_t *rowval)
{
    CTX_GUARD(ctx);
    GET_PATTERN(pattern);
    if (int rc = ensure_lower(C, P)) return rc;
    if (nnz_lower) *nnz_lower = P.nnz_lower;
    DevBuf c64, r64;
    if (colptr) {
        if (int rc = ensure(C, c64, (size_t)(P.ncols + 1) * 8)) return rc;
        add_one_kernel<<<nblocks(P.ncols + 1, 256), 256, 0, C->stream>>>(P.ncols + 1, P.lcolptr.as<long long>(), c64.as<long long>());
        LAUNCHED(C);
        EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(colptr, c64.p, (size_t)(P.ncols + 1) * 8, cudaMemcpyDefault, C->stream));
    }
    if (rowval) {
        if (int rc = ensure(C, r64, (size_t)std::max<long long>(P.nnz_lower, 1) * 8)) return rc;
        lower_pack_kernel<int, long long, 8><<<nblocks(P.ncols * 8, 256), 256, 0, C->stream>>>(0, P.ncols, P.lstart.as<long long>(), P.lcolptr.as<long long>(),
                                                                                             P.rowval.as<int>(), r64.as<long long>(), 1LL);
        LAUNCHED(C);
        EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(rowval, r64.p, (size_t)P.nnz_lower * 8, cudaMemcpyDefault, C->stream));
    }
    EXTFEM_CUDA_CHECK(C, cudaStreamSynchronize(C->stream));
    return EXTFEM_OK;
}

int extfem_values_get_lower(extfem_ctx *ctx, int pattern, double *nzval_lower, double *b)
{
    CTX_GUARD(ctx);
    GET_PATTERN(pattern);
    if (int rc = ensure_lower(C, P)) return rc;
    // pack kernels of column chunks on the main stream, the copy of a chunk on the exchange stream as soon as it is packed: the
    // packing hides behind the PCIe transfer (one chunk for small systems)
    const int nch = nzval_lower ? (int)P.lchunk_col.size() - 1 : 0;
    const bool piped = nch > 1 && C->stream2;
    if (piped) {
        for (int k = 0; k < nch; ++k)
            if (!C->cev[k]) EXTFEM_CUDA_CHECK(C, cudaEventCreateWithFlags(&C->cev[k], cudaEventDisableTiming));
        EXTFEM_CUDA_CHECK(C, cudaEventRecord(C->ev_fork, C->stream));
        EXTFEM_CUDA_CHECK(C, cudaStreamWaitEvent(C->stream2, C->ev_fork, 0));
    }
    cudaStream_t cs = piped ? C->stream2 : C->stream;
    if (b && piped) EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(b, P.b.p, (size_t)P.nrows * 8, cudaMemcpyDefault, cs));
    if (nzval_lower) {
        if (int rc = ensure(C, P.lpack, (size_t)std::max<long long>(P.nnz_lower, 1) * 8)) return rc;
        for (int k = 0; k < nch; ++k) {
            const long long c0 = P.lchunk_col[k], c1 = P.lchunk_col[k + 1], o0 = P.lchunk_off[k], o1 = P.lchunk_off[k + 1];
            if (c1 > c0) {
                lower_pack_kernel<double, double, 8><<<nblocks((c1 - c0) * 8, 256), 256, 0, C->stream>>>(
                    c0, c1, P.lstart.as<long long>(), P.lcolptr.as<long long>(), P.nzval.as<double>(), P.lpack.as<double>(), 0.0);
                LAUNCHED(C);
            }
            if (piped) {
                EXTFEM_CUDA_CHECK(C, cudaEventRecord(C->cev[k], C->stream));
                EXTFEM_CUDA_CHECK(C, cudaStreamWaitEvent(cs, C->cev[k], 0));
            }
            if (o1 > o0) EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(nzval_lower + o0, P.lpack.as<double>() + o0, (size_t)(o1 - o0) * 8, cudaMemcpyDefault, cs));
        }
    }
    if (b && !piped) EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(b, P.b.p, (size_t)P.nrows * 8, cudaMemcpyDefault, C->stream));
    if (piped) EXTFEM_CUDA_CHECK(C, cudaStreamSynchronize(C->stream2));
    EXTFEM_CUDA_CHECK(C, cudaStreamSynchronize(C->stream));
    return EXTFEM_OK;
}

int extfem_values_set(extfem_ctx *ctx, int pattern, const double *nzval, const double *b)
{
    CTX_GUARD(ctx);
    GET_PATTERN(pattern);
    if (nzval) EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(P.nzval.p, nzval, (size_t)P.nnz * 8, cudaMemcpyDefault, C->stream));
    if (b) EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(P.b.p, b, (size_t)P.nrows * 8, cudaMemcpyDefault, C->stream));
    EXTFEM_CUDA_CHECK(C, cudaStreamSynchronize(C->stream));
    return EXTFEM_OK;
}

int extfem_device_ptrs(extfem_ctx *ctx, int pattern, void **colptr, void **rowval, void **nzval, void **b)
{
    CTX_GUARD(ctx);
    GET_PATTERN(pattern);
    if (colptr) *colptr = P.colptr.p;
    if (rowval) *rowval = P.rowval.p;
    if (nzval) *nzval = P.nzval.p;
    if (b) *b = P.b.p;
    return EXTFEM_OK;
}

int extfem_apply_penalties(extfem_ctx *ctx, int pattern, int64_t ndofs, const int64_t *dofs, const double *values, double penalty)
{
    CTX_GUARD(ctx);
    GET_PATTERN(pattern);
    if (ndofs == 0) return EXTFEM_OK;
    if (!dofs) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "dofs is NULL");
    DevBuf dd, dv;
    if (int rc = upload(C, dd, dofs, (size_t)ndofs * 8)) return rc;
    if (values) if (int rc = upload(C, dv, values, (size_t)ndofs * 8)) return rc;
    DevBuf err;
    if (int rc = ensure(C, err, 4)) return rc;
    EXTFEM_CUDA_CHECK(C, cudaMemsetAsync(err.p, 0, 4, C->stream));
    // sharded system: the diagonal lives on the owning rank only (additive form), b stays consistent
    const double *owned = nullptr;
    if (C->dist.ready && C->dist.world > 1) owned = P.owned.ready ? (const double *)P.owned.weight : (P.iface.ready ? (const double *)P.iface.weight : nullptr);
    penalties_kernel<<<nblocks(ndofs, 256), 256, 0, C->stream>>>(ndofs, dd.as<long long>(), values ? dv.as<double>() : nullptr, penalty,
                                                                P.colptr.as<long long>(), P.rowval.as<int>(), P.nzval.as<double>(),
                                                                P.b.as<double>(), P.nrows, owned, err.as<int>());
    LAUNCHED(C);
    int herr = 0;
    EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(&herr, err.p, 4, cudaMemcpyDeviceToHost, C->stream));
    EXTFEM_CUDA_CHECK(C, cudaStreamSynchronize(C->stream));
    if (herr) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "penalty dof out of range or without diagonal entry");
    return EXTFEM_OK;
}

int extfem_spmv(extfem_ctx *ctx, int pattern, const double *x, double *y)
{
    CTX_GUARD(ctx);
    GET_PATTERN(pattern);
    if (!x || !y) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "x or y is NULL");
    DevBuf dx, dy;
    if (int rc = upload(C, dx, x, (size_t)P.ncols * 8)) return rc;
    if (int rc = ensure(C, dy, (size_t)P.nrows * 8)) return rc;
    if (int rc = csc_spmv(C->stream, P.nrows, P.ncols, P.nnz, P.colptr.as<long long>(), P.rowval.as<int>(), P.nzval.as<double>(),
                          dx.as<double>(), dy.as<double>(), P.cg, &C->launches))
        return fail(C, EXTFEM_ERR_CUDA, "spmv failed");
    EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(y, dy.p, (size_t)P.nrows * 8, cudaMemcpyDefault, C->stream));
    EXTFEM_CUDA_CHECK(C, cudaStreamSynchronize(C->stream));
    return EXTFEM_OK;
}

int extfem_residual(extfem_ctx *ctx, int pattern, const double *sol, double *res_out)
{
    CTX_GUARD(ctx);
    GET_PATTERN(pattern);
    if (!sol || !res_out) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "sol or res_out is NULL");
    DevBuf dx, dy;
    if (int rc = upload(C, dx, sol, (size_t)P.ncols * 8)) return rc;
    if (int rc = ensure(C, dy, (size_t)P.nrows * 8)) return rc;
    if (int rc = csc_spmv(C->stream, P.nrows, P.ncols, P.nnz, P.colptr.as<long long>(), P.rowval.as<int>(), P.nzval.as<double>(),
                          dx.as<double>(), dy.as<double>(), P.cg, &C->launches))
        return fail(C, EXTFEM_ERR_CUDA, "spmv failed");
    residual_kernel<<<nblocks(P.nrows, 256), 256, 0, C->stream>>>(P.nrows, P.b.as<double>(), dy.as<double>());
    LAUNCHED(C);
    EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(res_out, dy.p, (size_t)P.nrows * 8, cudaMemcpyDefault, C->stream));
    EXTFEM_CUDA_CHECK(C, cudaStreamSynchronize(C->stream));
    return EXTFEM_OK;
}

int extfem_cg(extfem_ctx *ctx, int pattern, const double *b, double *x, double rtol, int maxit, int *iters, double *relres)
{
    CTX_GUARD(ctx);
    GET_PATTERN(pattern);
    if (!P.square || P.nrows != P.ncols) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "CG needs a square system");
    if (!x) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "x is NULL");
    DevBuf db, dx;
    const double *bptr = P.b.as<double>();
    if (b) { if (int rc = upload(C, db, b, (size_t)P.nrows * 8)) return rc; bptr = db.as<double>(); }
    if (int rc = upload(C, dx, x, (size_t)P.nrows * 8)) return rc;
    int it = 0; double rr = 0;
    if (int rc = jacobi_cg(C->stream, P.nrows, P.nnz, P.colptr.as<long long>(), P.rowval.as<int>(), P.nzval.as<double>(), bptr,
                           dx.as<double>(), rtol, maxit, &it, &rr, P.cg, &C->launches))
        return fail(C, EXTFEM_ERR_CUDA, "CG failed (CUDA error or breakdown), code " + std::to_string(rc));
    EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(x, dx.p, (size_t)P.nrows * 8, cudaMemcpyDefault, C->stream));
    EXTFEM_CUDA_CHECK(C, cudaStreamSynchronize(C->stream));
    if (iters) *iters = it;
    if (relres) *relres = rr;
    return EXTFEM_OK;
}

/* ---- multi-GPU: one process per GPU, NCCL communicator inside the context (dist.cuh) ---------------------- */
int extfem_dist_unique_id(char *id128)
{
    if (!id128) return fail(nullptr, EXTFEM_ERR_BAD_ARGUMENT, "id128 is NULL");
    std::string e = g_nccl.load();
    if (!e.empty()) return fail(nullptr, EXTFEM_ERR_NCCL, e);
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) return fail(nullptr, EXTFEM_ERR_NCCL, std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r));
    memcpy(id128, &id, 128);
    return EXTFEM_OK;
}

int extfem_dist_init(extfem_ctx *ctx, int rank, int world, const char *id128)
{
    CTX_GUARD(ctx);
    if (world < 1 || rank < 0 || rank >= world) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "extfem_dist_init: bad rank / world");
    C->dist.rank = rank; C->dist.world = world;
    if (world > 1) {
        if (!id128) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "extfem_dist_init: id128 is NULL");
        std::string e = g_nccl.load();
        if (!e.empty()) return fail(C, EXTFEM_ERR_NCCL, e);
        ncclUniqueId id;
        memcpy(&id, id128, 128);
        ncclResult_t r = g_nccl.CommInitRank(&C->dist.comm, world, id, rank);
        if (r != ncclSuccess) return fail(C, EXTFEM_ERR_NCCL, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r));
    }
    C->dist.ready = true;
    return EXTFEM_OK;
}

int extfem_dist_set_interfaces(extfem_ctx *ctx, int pattern, int nneigh, const int32_t *neigh_ranks, const int64_t *ptr,
                               const int64_t *rows, const uint8_t *owned)
{
    CTX_GUARD(ctx);
    GET_PATTERN(pattern);
    if (!C->dist.ready) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "extfem_dist_init has not been called");
    if (nneigh < 0 || (nneigh > 0 && (!neigh_ranks || !ptr || !rows))) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "extfem_dist_set_interfaces: bad argument");
    IfacePlan &I = P.iface;
    I.ranks.assign(neigh_ranks, neigh_ranks + nneigh);
    I.ptr.assign(1, 0);
    for (int k = 0; k < nneigh; ++k) {
        if (ptr[k + 1] < ptr[k] || neigh_ranks[k] < 0 || neigh_ranks[k] >= C->dist.world || neigh_ranks[k] == C->dist.rank)
            return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "extfem_dist_set_interfaces: bad neighbour list");
        I.ptr.push_back(ptr[k + 1] - ptr[0]);
    }
    const long long ntot = I.ptr.back();
    std::vector<int> r0((size_t)std::max(ntot, 1ll));
    for (long long i = 0; i < ntot; ++i) {
        long long r = rows[ptr[0] + i] - 1;
        if (r < 0 || r >= P.nrows) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "extfem_dist_set_interfaces: row out of range");
        r0[i] = (int)r;
    }
    std::vector<double> w((size_t)P.nrows, 1.0);
    if (owned) for (long long i = 0; i < P.nrows; ++i) w[i] = owned[i] ? 1.0 : 0.0;
    for (void **p : {&I.rows, &I.sendbuf, &I.recvbuf, &I.weight}) if (*p) { cudaFree(*p); *p = nullptr; }
    EXTFEM_CUDA_CHECK(C, cudaMalloc(&I.rows, r0.size() * 4));
    EXTFEM_CUDA_CHECK(C, cudaMalloc(&I.sendbuf, r0.size() * 8));
    EXTFEM_CUDA_CHECK(C, cudaMalloc(&I.recvbuf, r0.size() * 8));
    EXTFEM_CUDA_CHECK(C, cudaMalloc(&I.weight, w.size() * 8));
    EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(I.rows, r0.data(), r0.size() * 4, cudaMemcpyHostToDevice, C->stream));
    EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(I.weight, w.data(), w.size() * 8, cudaMemcpyHostToDevice, C->stream));
    EXTFEM_CUDA_CHECK(C, cudaStreamSynchronize(C->stream));
    I.ready = true;
    return EXTFEM_OK;
}

#define GET_IFACE()                                                                                    \
    if (!C->dist.ready || !P.iface.ready)                                                              \
        return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "extfem_dist_init / extfem_dist_set_interfaces have not been called");

int extfem_dist_sum_rhs(extfem_ctx *ctx, int pattern)
{
    CTX_GUARD(ctx);
    GET_PATTERN(pattern);
    GET_IFACE();
    std::string e;
    if (iface_exchange_add(C->stream, C->dist, P.iface, P.b.as<double>(), &C->launches, &e)) return fail(C, EXTFEM_ERR_NCCL, e.empty() ? "interface exchange failed" : e);
    return EXTFEM_OK;
}

int extfem_dist_spmv(extfem_ctx *ctx, int pattern, const double *x, double *y)
{
    CTX_GUARD(ctx);
    GET_PATTERN(pattern);
    GET_IFACE();
    if (!x || !y) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "x or y is NULL");
    DevBuf dx, dy;
    if (int rc = upload(C, dx, x, (size_t)P.ncols * 8)) return rc;
    if (int rc = ensure(C, dy, (size_t)P.nrows * 8)) return rc;
    if (int rc = csc_spmv(C->stream, P.nrows, P.ncols, P.nnz, P.colptr.as<long long>(), P.rowval.as<int>(), P.nzval.as<double>(),
                          dx.as<double>(), dy.as<double>(), P.cg, &C->launches))
        return fail(C, EXTFEM_ERR_CUDA, "spmv failed");
    std::string e;
    if (iface_exchange_add(C->stream, C->dist, P.iface, dy.as<double>(), &C->launches, &e)) return fail(C, EXTFEM_ERR_NCCL, e.empty() ? "interface exchange failed" : e);
    EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(y, dy.p, (size_t)P.nrows * 8, cudaMemcpyDefault, C->stream));
    EXTFEM_CUDA_CHECK(C, cudaStreamSynchronize(C->stream));
    return EXTFEM_OK;
}

int extfem_dist_cg(extfem_ctx *ctx, int pattern, const double *b, double *x, double rtol, int maxit, int *iters, double *relres)
{
    CTX_GUARD(ctx);
    GET_PATTERN(pattern);
    GET_IFACE();
    if (!P.square || P.nrows != P.ncols) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "CG needs a square system");
    if (!x) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "x is NULL");
    DevBuf db, dx;
    const double *bptr = P.b.as<double>();
    if (b) { if (int rc = upload(C, db, b, (size_t)P.nrows * 8)) return rc; bptr = db.as<double>(); }
    if (int rc = upload(C, dx, x, (size_t)P.nrows * 8)) return rc;
    int it = 0; double rr = 0;
    std::string e;
    if (int rc = dist_jacobi_cg(C->stream, C->dist, P.iface, P.nrows, P.colptr.as<long long>(), P.rowval.as<int>(), P.nzval.as<double>(), bptr,
                                dx.as<double>(), rtol, maxit, &it, &rr, P.cg, &C->launches, &e))
        return fail(C, rc == -5 ? EXTFEM_ERR_NCCL : EXTFEM_ERR_CUDA, "distributed CG failed, code " + std::to_string(rc) + " " + e);
    EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(x, dx.p, (size_t)P.nrows * 8, cudaMemcpyDefault, C->stream));
    EXTFEM_CUDA_CHECK(C, cudaStreamSynchronize(C->stream));
    if (iters) *iters = it;
    if (relres) *relres = rr;
    return EXTFEM_OK;
}

/* ---- owned-row form (dist.cuh) ------------------------------------------------------------------------------------- */
static int upload_rows(Ctx *C, long long nrows, const int64_t *ptr, const int64_t *rows, int nneigh, OwnedExchange &X, long long *count,
                       void **dev, const char *what)
{
    std::vector<long long> &pp = (dev == &X.send_rows) ? X.send_ptr : X.recv_ptr;
    pp.assign(1, 0);
    for (int k = 0; k < nneigh; ++k) {
        if (ptr[k + 1] < ptr[k]) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, std::string("extfem_dist_set_owned: bad pointer array of ") + what);
        pp.push_back(ptr[k + 1] - ptr[0]);
    }
    const long long n = pp.back();
    std::vector<int> r0((size_t)std::max(n, 1ll), 0);
    for (long long i = 0; i < n; ++i) {
        const long long r = rows[ptr[0] + i] - 1;
        if (r < 0 || r >= nrows) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, std::string("extfem_dist_set_owned: row out of range in ") + what);
        r0[i] = (int)r;
    }
    if (*dev) { cudaFree(*dev); *dev = nullptr; }
    EXTFEM_CUDA_CHECK(C, cudaMalloc(dev, r0.size() * 4));
    EXTFEM_CUDA_CHECK(C, cudaMemcpy(*dev, r0.data(), r0.size() * 4, cudaMemcpyHostToDevice));
    *count = n;
    return 0;
}

int extfem_dist_set_owned(extfem_ctx *ctx, int pattern, int nneigh, const int32_t *neigh_ranks, const int64_t *red_send_ptr,
                          const int64_t *red_send_rows, const int64_t *red_recv_ptr, const int64_t *red_recv_rows,
                          const int64_t *halo_send_ptr, const int64_t *halo_send_rows, const int64_t *halo_recv_ptr,
                          const int64_t *halo_recv_rows, const uint8_t *owned)
{
    CTX_GUARD(ctx);
    GET_PATTERN(pattern);
    if (!C->dist.ready) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "extfem_dist_init has not been called");
    if (!P.square) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "the owned-row form needs a square system");
    if (nneigh < 0 || !owned || (nneigh > 0 && (!neigh_ranks || !red_send_ptr || !red_recv_ptr || !halo_send_ptr || !halo_recv_ptr)))
        return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "extfem_dist_set_owned: bad argument");
    OwnedPlanDev &O = P.owned;
    O.ready = false;
    O.ranks.assign(neigh_ranks, neigh_ranks + nneigh);
    for (int k = 0; k < nneigh; ++k)
        if (neigh_ranks[k] < 0 || neigh_ranks[k] >= C->dist.world || neigh_ranks[k] == C->dist.rank)
            return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "extfem_dist_set_owned: bad neighbour list");
    static const int64_t zero2[2] = {0, 0};
    auto P0 = [&](const int64_t *p) { return nneigh > 0 ? p : zero2; };
    if (int rc = upload_rows(C, P.nrows, P0(red_send_ptr), red_send_rows, nneigh, O.red, &O.red.nsend, &O.red.send_rows, "red_send")) return rc;
    if (int rc = upload_rows(C, P.nrows, P0(red_recv_ptr), red_recv_rows, nneigh, O.red, &O.red.nrecv, &O.red.recv_rows, "red_recv")) return rc;
    if (int rc = upload_rows(C, P.nrows, P0(halo_send_ptr), halo_send_rows, nneigh, O.halo, &O.halo.nsend, &O.halo.send_rows, "halo_send")) return rc;
    if (int rc = upload_rows(C, P.nrows, P0(halo_recv_ptr), halo_recv_rows, nneigh, O.halo, &O.halo.nrecv, &O.halo.recv_rows, "halo_recv")) return rc;
    // column segments of the matrix reduction
    auto segs = [&](const std::vector<long long> &ptr, const int64_t *rows0, long long n, std::vector<long long> &seg, std::vector<long long> &vptr) {
        seg.assign((size_t)n + 1, 0);
        for (long long i = 0; i < n; ++i) { const long long c = rows0[i] - 1; seg[i + 1] = seg[i] + (P.hcolptr[c + 1] - P.hcolptr[c]); }
        vptr.assign(ptr.size(), 0);
        for (size_t k = 0; k < ptr.size(); ++k) vptr[k] = seg[ptr[k]];
    };
    std::vector<long long> sseg, rseg;
    segs(O.red.send_ptr, red_send_rows ? red_send_rows + P0(red_send_ptr)[0] : nullptr, O.red.nsend, sseg, O.send_vptr);
    segs(O.red.recv_ptr, red_recv_rows ? red_recv_rows + P0(red_recv_ptr)[0] : nullptr, O.red.nrecv, rseg, O.recv_vptr);
    // the receive-side segment table is indexed per listed column (global offsets into recvbuf)
    for (void **pp : {&O.send_seg, &O.recv_seg, &O.sendbuf, &O.recvbuf, &O.weight}) if (*pp) { cudaFree(*pp); *pp = nullptr; }
    EXTFEM_CUDA_CHECK(C, cudaMalloc(&O.send_seg, sseg.size() * 8));
    EXTFEM_CUDA_CHECK(C, cudaMalloc(&O.recv_seg, rseg.size() * 8));
    EXTFEM_CUDA_CHECK(C, cudaMemcpy(O.send_seg, sseg.data(), sseg.size() * 8, cudaMemcpyHostToDevice));
    EXTFEM_CUDA_CHECK(C, cudaMemcpy(O.recv_seg, rseg.data(), rseg.size() * 8, cudaMemcpyHostToDevice));
    O.bufdoubles = (size_t)std::max<long long>({sseg.back(), rseg.back(), O.halo.nsend, O.halo.nrecv, O.red.nsend, O.red.nrecv, 1});
    EXTFEM_CUDA_CHECK(C, cudaMalloc(&O.sendbuf, O.bufdoubles * 8));
    EXTFEM_CUDA_CHECK(C, cudaMalloc(&O.recvbuf, O.bufdoubles * 8));
    std::vector<double> w((size_t)P.nrows);
    for (long long i = 0; i < P.nrows; ++i) w[i] = owned[i] ? 1.0 : 0.0;
    EXTFEM_CUDA_CHECK(C, cudaMalloc(&O.weight, w.size() * 8));
    EXTFEM_CUDA_CHECK(C, cudaMemcpy(O.weight, w.data(), w.size() * 8, cudaMemcpyHostToDevice));
    // both sides of every exchange must agree on the column lengths (same rows in the same order): exchange and compare
    if (C->dist.world > 1 && (O.red.nsend > 0 || O.red.nrecv > 0)) {
        std::vector<double> mylen((size_t)std::max(O.red.nsend, 1ll)), got((size_t)std::max(O.red.nrecv, 1ll));
        for (long long i = 0; i < O.red.nsend; ++i) mylen[i] = (double)(sseg[i + 1] - sseg[i]);
        EXTFEM_CUDA_CHECK(C, cudaMemcpy(O.sendbuf, mylen.data(), (size_t)O.red.nsend * 8, cudaMemcpyHostToDevice));
        std::string e;
        if (owned_sendrecv(C->stream, C->dist, O.ranks, O.red.send_ptr, O.red.recv_ptr, (const double *)O.sendbuf, (double *)O.recvbuf, &e))
            return fail(C, EXTFEM_ERR_NCCL, e);
        EXTFEM_CUDA_CHECK(C, cudaStreamSynchronize(C->stream));
        EXTFEM_CUDA_CHECK(C, cudaMemcpy(got.data(), O.recvbuf, (size_t)O.red.nrecv * 8, cudaMemcpyDeviceToHost));
        for (long long i = 0; i < O.red.nrecv; ++i)
            if ((long long)got[i] != rseg[i + 1] - rseg[i])
                return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "extfem_dist_set_owned: a shared column has different lengths on the two ranks "
                                                        "(the local meshes need one layer of ghost cells, host/dist.py: OwnedShard)");
    }
    O.ready = true;
    return EXTFEM_OK;
}

#define GET_OWNED()                                                                                    \
    if (!C->dist.ready || !P.owned.ready)                                                              \
        return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "extfem_dist_init / extfem_dist_set_owned have not been called");

int extfem_dist_reduce_system(extfem_ctx *ctx, int pattern, int matrix, int rhs)
{
    CTX_GUARD(ctx);
    GET_PATTERN(pattern);
    GET_OWNED();
    std::string e;
    if (matrix) {
        // on the exchange stream, behind everything queued so far: a following extfem_assemble_linear overlaps with it
        EXTFEM_CUDA_CHECK(C, cudaEventRecord(C->ev_fork, C->stream));
        EXTFEM_CUDA_CHECK(C, cudaStreamWaitEvent(C->stream2, C->ev_fork, 0));
        if (owned_reduce_matrix(C->stream2, C->dist, P.owned, P.colptr.as<long long>(), P.nzval.as<double>(), &C->launches, &e))
            return fail(C, EXTFEM_ERR_NCCL, e.empty() ? "matrix reduction failed" : e);
        EXTFEM_CUDA_CHECK(C, cudaEventRecord(C->ev_join, C->stream2));
        C->aux_pending = true;
    }
    if (rhs) {
        join_aux(C);   // the exchange buffers are shared
        if (owned_reduce_vector(C->stream, C->dist, P.owned, P.b.as<double>(), &C->launches, &e))
            return fail(C, EXTFEM_ERR_NCCL, e.empty() ? "rhs reduction failed" : e);
    }
    return EXTFEM_OK;
}

int extfem_dist_spmv_owned(extfem_ctx *ctx, int pattern, const double *x, double *y)
{
    CTX_GUARD(ctx);
    GET_PATTERN(pattern);
    GET_OWNED();
    if (!x || !y) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "x or y is NULL");
    DevBuf dx, dy;
    if (int rc = upload(C, dx, x, (size_t)P.ncols * 8)) return rc;
    if (int rc = ensure(C, dy, (size_t)P.nrows * 8)) return rc;
    std::string e;
    if (owned_halo(C->stream, C->dist, P.owned, dx.as<double>(), &C->launches, &e)) return fail(C, EXTFEM_ERR_NCCL, e.empty() ? "halo exchange failed" : e);
    spmv_kernel<<<nblocks(P.nrows * SPMV_LANES, 256), 256, 0, C->stream>>>(P.nrows, P.colptr.as<long long>(), P.rowval.as<int>(), nullptr,
                                                                           P.nzval.as<double>(), dx.as<double>(), dy.as<double>());
    mask_rows_kernel<<<nblocks(P.nrows, 256), 256, 0, C->stream>>>(P.nrows, (const double *)P.owned.weight, dy.as<double>());
    LAUNCHED(C); LAUNCHED(C);
    EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(y, dy.p, (size_t)P.nrows * 8, cudaMemcpyDefault, C->stream));
    EXTFEM_CUDA_CHECK(C, cudaStreamSynchronize(C->stream));
    return EXTFEM_OK;
}

int extfem_dist_cg_owned(extfem_ctx *ctx, int pattern, const double *b, double *x, double rtol, int maxit, int *iters, double *relres)
{
    CTX_GUARD(ctx);
    GET_PATTERN(pattern);
    GET_OWNED();
    if (!x) return fail(C, EXTFEM_ERR_BAD_ARGUMENT, "x is NULL");
    DevBuf db, dx;
    const double *bptr = P.b.as<double>();
    if (b) { if (int rc = upload(C, db, b, (size_t)P.nrows * 8)) return rc; bptr = db.as<double>(); }
    if (int rc = upload(C, dx, x, (size_t)P.nrows * 8)) return rc;
    int it = 0; double rr = 0;
    std::string e;
    if (int rc = owned_jacobi_cg(C->stream, C->dist, P.owned, P.nrows, P.colptr.as<long long>(), P.rowval.as<int>(), P.nzval.as<double>(), bptr,
                                 dx.as<double>(), rtol, maxit, &it, &rr, P.cg, &C->launches, &e))
        return fail(C, rc == -5 ? EXTFEM_ERR_NCCL : EXTFEM_ERR_CUDA, "owned-row CG failed, code " + std::to_string(rc) + " " + e);
    EXTFEM_CUDA_CHECK(C, cudaMemcpyAsync(x, dx.p, (size_t)P.nrows * 8, cudaMemcpyDefault, C->stream));
    EXTFEM_CUDA_CHECK(C, cudaStreamSynchronize(C->stream));
    if (iters) *iters = it;
    if (relres) *relres = rr;
    return EXTFEM_OK;
}

} // extern "C"

