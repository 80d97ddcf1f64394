/*
 * extfem_cuda.h -- C-ABI of libextfem_cuda.so, the B200 (sm_100a) assembly engine behind
 * ExtendableFEM.jl's operator API.
 *
 * The reference has no FFI: its hot path sits behind Julia multiple dispatch.  The seam this
 * library replaces is the closure `O.assembler` that `build_assembler!` creates and
 * `assemble!` invokes with raw arrays:
 *
 *   BilinearOperator   src/common_operators/bilinear_operator.jl:659 (build_assembler!),
 *                      :820-951 / :451-596 (assembly_loop), :955-1003 (assembler), :1013/:1050 (assemble!)
 *   LinearOperator     src/common_operators/linear_operator.jl:482 / :241 (build_assembler!),
 *                      :584-640 / :359-438 (assembly_loop), :697/:774 (assemble!)
 *   NonlinearOperator  src/common_operators/nonlinear_operator.jl:129 (build_assembler!),
 *                      :283-436 (assembly_loop), :479/:510 (assemble!)
 *   assemble_system!   src/solvers.jl:124-195 (zeroing, flush!), residual src/solvers.jl:38-43
 *
 * Conventions
 *  - every function returns EXTFEM_OK (0) or a negative EXTFEM_ERR_* code; the message is
 *    available from extfem_last_error(ctx).  No C++ exception crosses the boundary.
 *  - index arrays handed in are 1-based (Julia), 4 or 8 bytes wide (`index_bytes`), laid out
 *    exactly like the Julia arrays: coords[dim, nnodes], cellnodes[dim+1, ncells],
 *    celldofs[ndofs4cell, ncells] column-major (== C row-major [nitems][k]).
 *  - host arrays are borrowed for the duration of the call only.  Every data pointer may be a
 *    host pointer OR a device pointer (unified addressing; copies use cudaMemcpyDefault), so a
 *    caller that already has device-resident data (CUDA.jl, torch) pays no PCIe transfer.
 *  - the matrix is CSC, Int64, 1-based, rows sorted per column -- the layout of
 *    `A.entries.cscmatrix` (src/solver_config.jl:190, src/solvers.jl:134).  The CSR arrays
 *    north_star mentions are the same arrays for the transposed matrix.
 *  - the sparsity pattern is STRUCTURAL (dofs sharing a cell, filtered by the block coupling);
 *    the reference's value-dependent pattern (entries exactly 0.0 are never inserted,
 *    bilinear_operator.jl:925) is always a subset of it.
 *  - a context is not re-entrant; calls on one context are serialised by the caller.  Several contexts may live in one
 *    process (one per device, or several on one device): contexts on the SAME device share the device's __constant__
 *    tables; the library orders a context's upload behind the other context's last kernel that read them (event wait), so
 *    they may be used alternately without extfem_synchronize, but not from two host threads at once.
 *  - there is no CPU fallback: without a CUDA device extfem_ctx_create fails.
 */
#ifndef EXTFEM_CUDA_H
#define EXTFEM_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct extfem_ctx extfem_ctx;

/* ---- error codes ----------------------------------------------------------------------- */
enum {
    EXTFEM_OK = 0,
    EXTFEM_ERR_UNREGISTERED_KERNEL = -1, /* Julia closure has no registry entry (north_star)   */
    EXTFEM_ERR_UNSUPPORTED_ELEMENT = -2, /* fetype / operator / entity outside the hot path     */
    EXTFEM_ERR_BAD_ARGUMENT = -3,
    EXTFEM_ERR_CUDA = -4,
    EXTFEM_ERR_NCCL = -5,
    EXTFEM_ERR_CAPACITY = -6             /* an internal fixed capacity was exceeded             */
};

/* ---- finite elements / function operators ----------------------------------------------- */
enum { EXTFEM_FE_H1P1 = 1, EXTFEM_FE_H1P2 = 2, EXTFEM_FE_TABULATED = 100 };
/* assembly entities (:entities kwarg, bilinear_operator.jl:61, 707-714): cells or boundary faces */
enum { EXTFEM_ON_CELLS = 0, EXTFEM_ON_BFACES = 1 };
/* function operators (ExtendableFEMBase: Identity, Gradient, Divergence, SymmetricGradient) */
enum { EXTFEM_OP_ID = 0, EXTFEM_OP_GRAD = 1, EXTFEM_OP_DIV = 2, EXTFEM_OP_SYMGRAD_VOIGT = 3 };

/* ---- kernel registry; extfem_kernel_id("name") resolves names ------ */
enum { /* BilinearOperator kernels: result = C(x, args) * input */
    EXTFEM_BLK_STANDARD = 1,     /* "standard"      ExtendableFEMBase.standard_kernel (bilinear_operator.jl:248) */
    EXTFEM_BLK_DCR = 2,          /* "dcr"           Example220:66-72; params alpha, nu, beta[dim]                */
    EXTFEM_BLK_STOKES = 3,       /* "stokes"        docs/src/bilinearoperator.md:37-45; params mu                */
    EXTFEM_BLK_LINNSE7 = 4,      /* "linnse7"       test/test_nonlinear_operator.jl:18-28; params mu, alpha      */
    EXTFEM_BLK_HOOKE_GRAD = 5,   /* "hooke_grad"    isotropic Hooke on grad(u); params mu, lambda                */
    EXTFEM_BLK_HOOKE_VOIGT = 6,  /* "hooke_voigt"   Example312:55 sigma = C*epsV(u); params C row-major          */
    EXTFEM_BLK_CONVECT_ARGS = 7, /* "convect_args"  (args . grad) u, kernel with args (bilinear_operator.jl:536)  */
    EXTFEM_BLK_ROBIN108 = 8      /* "robin108"      Example108:48-51 result = params[0] - input (used ON_BFACES)     */
};
enum { /* LinearOperator kernels f(x) */
    EXTFEM_LIN_CONSTANT_ONE = 1,    /* "constant_one"    constant_one_kernel (linear_operator.jl:159)  */
    EXTFEM_LIN_CONSTANT_PARAMS = 2, /* "constant_params" result .= params (Example330:44, Example312:58) */
    EXTFEM_LIN_XY = 3,              /* "xy"              README.md:37-40 / Example201:32-35              */
    EXTFEM_LIN_SINCOS301 = 4,       /* "sincos301"       Example301:33-35; params mu                     */
    EXTFEM_LIN_TABULATED = 5,       /* "tabulated"       any Julia closure, evaluated by the host at the
                                                         quadrature points (extfem_quadrature_points)    */
    EXTFEM_LIN_EXP2X = 6,           /* "exp2x"           Example108:31-34  f = exp(2 x1)                 */
    EXTFEM_LIN_STEP105 = 7          /* "step105"         Example105:35-38  f = x1 < 0.5 ? -1 : 1         */
};
enum { /* NonlinearOperator kernels with analytic Jacobians */
    EXTFEM_NL_NSE2D = 1,       /* "nse2d"       Example250:59-74; params mu          */
    EXTFEM_NL_LINNSE7 = 2,     /* "nl_linnse7"  test/test_nonlinear_operator.jl:18-28 */
    EXTFEM_NL_NEOHOOKE3D = 3,  /* "neohooke3d"  Example330:49-57 (DW); params mu, lambda */
    EXTFEM_NL_RCD = 4,         /* "rcd"         Example108:40-45                        */
    EXTFEM_NL_NLPOISSON105 = 5,/* "nlpoisson105" Example105:45-50 [exp(u)-exp(-u), eps grad u]; params eps */
    EXTFEM_NL_STVENANT230 = 6, /* "stvenant230" Example230:39-72 (2D, [grad(u)]); params R, lambda[R], mu[R], epsT[R]
                                                 indexed by the cell region                  */
    EXTFEM_NL_POROUS106 = 7    /* "porous106"   Example106:47-52: test [grad(u)], args [id(u), grad(u)], m u^(m-1) grad u; params m */
};
enum { /* ItemIntegrator kernels (item_integrator.jl:26-28, 71-81 and the examples' exact_error! closures) */
    EXTFEM_II_STANDARD = 1,         /* "ii_standard"      result = input (ItemIntegrator(oa_args), item_integrator.jl:78-81) */
    EXTFEM_II_L2NORM = 2,           /* "l2norm"           l2norm_kernel: result = input.^2 (item_integrator.jl:26-28)        */
    EXTFEM_II_L2DIFF_TABULATED = 3, /* "l2diff_tabulated" (ref - input).^2 with ref evaluated by the host at the quadrature
                                                          points (any exact_error! closure)                               */
    EXTFEM_II_L2ERR_SINCOS301 = 4,  /* "l2err_sincos301"  Example301:62-67 (sin(1.7x)cos(3.9y) - u)^2                      */
    EXTFEM_II_L2ERR_EXP108 = 5      /* "l2err_exp108"     Example108:86-90 (exp(x) - u)^2                                  */
};

#define EXTFEM_MAXARGS 4 /* (unknown, operator) pairs per role */

/* One operator, flattened from the Julia kwargs tables (bilinear_operator.jl:56-73,
 * linear_operator.jl:34-47, nonlinear_operator.jl:31-48).  `*_block` index the row (test) /
 * column (ansatz, args) blocks of the pattern, like `A[j,j]` in bilinear_operator.jl:1030. */
typedef struct {
    int32_t ntest, test_block[EXTFEM_MAXARGS], test_op[EXTFEM_MAXARGS];
    int32_t nansatz, ansatz_block[EXTFEM_MAXARGS], ansatz_op[EXTFEM_MAXARGS];
    int32_t nargs, args_block[EXTFEM_MAXARGS], args_op[EXTFEM_MAXARGS];
    int32_t kernel_id;
    int32_t nparams;
    const double *params;      /* qpinfo.params                                        */
    double factor;             /* :factor                                              */
    double time;               /* qpinfo.time                                          */
    double symgrad_offdiag;    /* offdiagval of SymmetricGradient operators            */
    int32_t quadorder;         /* :quadorder, -1 == "auto"                             */
    int32_t bonus_quadorder;   /* :bonus_quadorder                                     */
    int32_t nregions;          /* :regions ([] == all)                                 */
    const int32_t *regions;
    int32_t transposed_copy;   /* :transposed_copy (0, 1, -1)                          */
    int32_t lump;              /* :lump (0, 1, 2)                                      */
    const uint8_t *coupling;   /* [nansatz][ntest] couples_with (NULL == all couple)   */
    /* optional host-supplied quadrature rule (from ExtendableFEMBase's QuadratureRule) */
    int32_t nq_custom;
    const double *qweights;    /* [nq]                                                 */
    const double *qpoints;     /* [nq][dim]                                            */
    const double *tabulated;   /* EXTFEM_LIN_TABULATED / EXTFEM_II_L2DIFF_TABULATED: [nitems][nq][oplen] */
    int32_t entities;          /* :entities, EXTFEM_ON_CELLS (0) or EXTFEM_ON_BFACES   */
} extfem_opdesc;

/* ---- context ---------------------------------------------------------------------------- */
int extfem_ctx_create(int device, extfem_ctx **out);
int extfem_ctx_destroy(extfem_ctx *ctx);
const char *extfem_last_error(extfem_ctx *ctx);      /* ctx may be NULL: last global error    */
int extfem_kernel_id(const char *name);              /* <0: EXTFEM_ERR_UNREGISTERED_KERNEL     */
int extfem_synchronize(extfem_ctx *ctx);
/* engine options: "fastpath" (default 1): 0 forces the generic two-phase path for every operator */
int extfem_set_option(extfem_ctx *ctx, const char *key, int value);
/* further options: "fastpath_closed_form" (1), "fastpath_templates" (1): 0 keeps every column on the per-pair
 * record kernel, "template_min_cols" (24): smallest group of structurally identical columns that gets a template,
 * "template_pool_bytes" (37376), "template_prefetch_ctas" (1024), "template_constant_memory" (1),
 * "template_permute_mesh" (1): tuning knobs of the template kernel,
 * "nonlinear_kernel" (3): local kernel of NonlinearOperator, 1 entry-wise | 2 staged per block | 3 warp per cell,
 * "template_jit" (0) / "template_jit_min_cols" (200000): experimental plan-time specialisation of the templates through
 * NVRTC (jit.cuh), "template_plane_mask" (1) */
/* number of kernel launches issued by this context since creation (bench "gpu_launches") */
int64_t extfem_launch_count(extfem_ctx *ctx);
/* device time in ms of the phases of the last assemble call (CUDA events):
 * [0] local (cell) kernel, [1] gather kernel(s), [2] total incl. copies.  */
int extfem_last_timings(extfem_ctx *ctx, double *ms3);

/* user-level device timers (the reference wraps assemble! in TimerOutputs sections,
 * src/solvers.jl:140): record event `slot` (0..15) on the context's stream / elapsed ms. */
int extfem_event_record(extfem_ctx *ctx, int slot);
int extfem_event_elapsed_ms(extfem_ctx *ctx, int slot_start, int slot_stop, double *ms);

/* ---- grid:  xgrid[Coordinates], xgrid[CellNodes], xgrid[CellRegions], xgrid[CellVolumes]
 *      (bilinear_operator.jl:693-695) ------------------------------------------------------- */
int extfem_mesh_set(extfem_ctx *ctx, int dim, int64_t ncells, int64_t nnodes, const double *coords,
                    const void *cellnodes, int index_bytes, const int32_t *cellregions /* NULL: all 1 */,
                    const double *cellvolumes /* NULL: computed */, int *mesh_out);
/* boundary faces of the grid: xgrid[BFaceNodes], xgrid[BFaceRegions], xgrid[BFaceVolumes] (read by the ON_BFACES branch of
 * build_assembler!, bilinear_operator.jl:693-714).  bfacenodes[dim, nbfaces] column-major, 1-based. */
int extfem_mesh_set_bfaces(extfem_ctx *ctx, int mesh, int64_t nbfaces, const void *bfacenodes, int index_bytes,
                           const int32_t *bfaceregions /* NULL: all 1 */, const double *bfacevolumes /* NULL: computed */);
/* new node coordinates for an existing mesh (moving meshes; refreshes the geometry cache) */
int extfem_mesh_update_coords(extfem_ctx *ctx, int mesh, const double *coords, const double *cellvolumes);

/* ---- FESpace: FES[CellDofs] via get_dofmap (src/helper_functions.jl:561-567) ------------- */
int extfem_space_set(extfem_ctx *ctx, int mesh, int fetype, int ncomp, const void *celldofs, int index_bytes,
                     int ndofs4cell, int64_t ndofs, int *space_out);
/* EXTFEM_FE_TABULATED (any affine H1-conforming scalar-Lagrange-type element, e.g. H1Pk order 3 of README.md:50 /
 * Example201:66): the host supplies the scalar reference basis as POLYNOMIALS in the reference coordinates, so the engine can
 * evaluate values and gradients at the points of whatever quadrature rule an operator asks for.  coeffs[nscalar][nmono]
 * (row-major), monomials x^i y^j z^k with i+j+k <= order enumerated by `for k: for j: for i` (i fastest; lower dimensions drop
 * the outer loops): nmono = binomial(order + dim, dim).  extfem_space_set with fetype == EXTFEM_FE_TABULATED creates the space
 * (nscalar = ndofs4cell / ncomp); it cannot be used before its basis is set.  bface_coeffs (may be NULL) is the basis of the
 * restriction to a boundary face in the face's own reference coordinates, [nscalar_bface][binomial(order + dim-1, dim-1)].
 * Cell-wise orientation-dependent elements keep ONE reference basis here: the host orders `celldofs` per cell accordingly. */
int extfem_space_set_tables(extfem_ctx *ctx, int space, int order, int nscalar, const double *coeffs,
                            int nscalar_bface, const double *bface_coeffs);
/* FES[BFaceDofs] (read through get_dofmap for ON_BFACES, src/helper_functions.jl:561-567): bfacedofs[ndofs4bface, nbfaces],
 * 1-based, same global numbering as celldofs; needs extfem_mesh_set_bfaces on the space's mesh.                        */
int extfem_space_set_bfacedofs(extfem_ctx *ctx, int space, const void *bfacedofs, int index_bytes, int ndofs4bface);

/* ---- FEMatrix pattern (replaces ExtendableSparse rawupdateindex!/flush!) ------------------
 * Block system with row blocks `rowspaces` and column blocks `colspaces`; block (r,c) is
 * present when block_coupling[c*nrow + r] != 0 (NULL: all).                                   */
int extfem_pattern_build(extfem_ctx *ctx, int nrowspaces, const int *rowspaces, int ncolspaces, const int *colspaces,
                         const uint8_t *block_coupling, int *pattern_out);
int extfem_pattern_dims(extfem_ctx *ctx, int pattern, int64_t *nrows, int64_t *ncols, int64_t *nnz);
int extfem_pattern_get(extfem_ctx *ctx, int pattern, int64_t *colptr /*[ncols+1]*/, int64_t *rowval /*[nnz]*/);

/* ---- assembly.  accumulate == 0 first zeroes the device values (fill!(nzval,0),
 *      src/solvers.jl:130-135); out pointers may be NULL (values stay device-resident).  A call that copies results into
 *      caller memory returns when the copy is complete; a call whose results stay device-resident returns as soon as its
 *      kernels are enqueued on the context's stream (later extfem_* calls are ordered behind them; extfem_synchronize,
 *      extfem_last_timings and every call that returns data wait).                           */
int extfem_assemble_bilinear(extfem_ctx *ctx, int pattern, const extfem_opdesc *op, const double *sol /* args */,
                             int accumulate, double *nzval_out);
int extfem_assemble_linear(extfem_ctx *ctx, int pattern, const extfem_opdesc *op, const double *sol /* args */,
                           int accumulate, double *b_out);
int extfem_assemble_nonlinear(extfem_ctx *ctx, int pattern, const extfem_opdesc *op, const double *sol,
                              int accumulate, double *nzval_out, double *b_out);
/* statistics of the scatter-map plans of diagonal block `block` (built on first fast-path use), stats8 =
 * { period P of the geometry order, templates, template warps, columns on the record kernel, CTAs of the template
 *   kernel, shared-memory pool bytes, template rounds, columns of the block }                                    */
int extfem_plan_stats(extfem_ctx *ctx, int pattern, int block, int64_t *stats8);
/* 1: the block's templates run as kernels specialised at plan time (NVRTC, sm_100a; options "template_jit",
 * "template_jit_min_cols"), 0: they do not, -1: specialisation failed and the static template kernel is used */
int extfem_plan_jit_status(extfem_ctx *ctx, int pattern, int block);
/* x at the quadrature points the operator will use: xq[ncells][nq][dim]; *nq_out = nq.
 * xq may be NULL to query nq only.  (Host evaluates its closure there -> EXTFEM_LIN_TABULATED.) */
int extfem_quadrature_points(extfem_ctx *ctx, int pattern, const extfem_opdesc *op, int is_linear, int *nq_out,
                             double *xq);

/* ---- ItemIntegrator (src/common_operators/item_integrator.jl:191-249, evaluate :323-352): integrates
 *      kernel(input_args(sol), qpinfo) over every cell.  `op` uses args_* (blocks index the pattern's column blocks), kernel_id
 *      (EXTFEM_II_*), params, factor, quadorder ("auto" = max polynomial order of the arguments, :147-151), regions, tabulated.
 *      piecewise != 0: out[resultdim, ncells] column-major (the matrix `evaluate` returns); else out[resultdim] (sum over cells). */
int extfem_integrate(extfem_ctx *ctx, int pattern, const extfem_opdesc *op, const double *sol, int resultdim, int piecewise,
                     double *out);

/* ---- device-resident system --------------------------------------------------------------- */
/* fill!(A.entries.cscmatrix.nzval, 0) / fill!(b.entries, 0) at the start of assemble_system! (src/solvers.jl:130-135) */
int extfem_values_zero(extfem_ctx *ctx, int pattern, int zero_matrix, int zero_rhs);
int extfem_values_get(extfem_ctx *ctx, int pattern, double *nzval /*NULL ok*/, double *b /*NULL ok*/);
int extfem_values_set(extfem_ctx *ctx, int pattern, const double *nzval /*NULL ok*/, const double *b /*NULL ok*/);
/* Symmetric forms: only the lower triangle crosses PCIe (half the bytes of extfem_values_get).  Rows are sorted per column, so
 * the entries with row >= column are a suffix of every column; they are packed on the device in column order.  The Julia side
 * wraps the result as Symmetric(SparseMatrixCSC(n, n, colptr, rowval, nzval), :L) -- what CHOLMOD / a CG need of an SPD matrix.
 * pattern_get_lower: nnz_lower, colptr [ncols+1] and rowval [nnz_lower] (Int64, 1-based; any of them may be NULL).
 * values_get_lower packs column chunks on the context's stream and copies every chunk on its exchange stream as soon as it is
 * packed (pinned host arrays make the copies asynchronous); it returns when both arrays are complete.                          */
int extfem_pattern_get_lower(extfem_ctx *ctx, int pattern, int64_t *nnz_lower, int64_t *colptr, int64_t *rowval);
int extfem_values_get_lower(extfem_ctx *ctx, int pattern, double *nzval_lower /*[nnz_lower], NULL ok*/, double *b /*NULL ok*/);
/* raw device pointers (colptr int64 0-based, rowval int32 0-based, nzval, b) for zero-copy users */
int extfem_device_ptrs(extfem_ctx *ctx, int pattern, void **colptr, void **rowval, void **nzval, void **b);
/* apply_penalties! (homogeneousdata_operator.jl:186-201): A[d,d] = penalty, b[d] = penalty*value[d].  On a sharded system
 * (after extfem_dist_set_interfaces) every rank passes ALL its local boundary dofs, like the reference does per partition: the
 * diagonal is written on the owning rank only (0 elsewhere, the additive sum is `penalty`) and b[d] = penalty*value[d] on every
 * sharing rank (b stays consistent; call it after extfem_dist_sum_rhs).                                                */
int extfem_apply_penalties(extfem_ctx *ctx, int pattern, int64_t ndofs, const int64_t *dofs /*1-based*/,
                           const double *values /*NULL: 0*/, double penalty);
/* the assemble_sol leg of apply_penalties! (homogeneousdata_operator.jl:198-201, interpolateboundarydata_operator.jl:160-167):
 * sol[dofs] = values (NULL: 0) on a caller vector (host or device pointer)                                              */
int extfem_apply_values(extfem_ctx *ctx, int64_t ndofs, const int64_t *dofs /*1-based*/, const double *values, double *sol,
                        int64_t nsol);
/* compute_nonlinear_residual! (src/solvers.jl:38-43): res = b - A*sol */
int extfem_residual(extfem_ctx *ctx, int pattern, const double *sol, double *res_out);
/* y = A*x on the device-resident matrix */
int extfem_spmv(extfem_ctx *ctx, int pattern, const double *x, double *y);
/* Jacobi-preconditioned CG on the device-resident system (square patterns); b == NULL uses the
 * device-resident right-hand side.  x is in/out (initial guess); the iteration stops when |b - A x| <= rtol * |b - A x0| (pass x0
 * with the Dirichlet values set when rows are penalised).  The matrix must be SYMMETRIC positive definite: the CSC
 * arrays are traversed as CSR (A^T x); use extfem_spmv / extfem_residual (which build a true row-major view) for
 * non-symmetric systems such as Newton matrices of convection terms.                           */
int extfem_cg(extfem_ctx *ctx, int pattern, const double *b, double *x, double rtol, int maxit, int *iters,
              double *relres);

/* ---- multi-GPU: one process per GPU, cells partitioned into contiguous ranges.  Every rank assembles
 *      its own cells into a local system over its local dofs with the calls above; the global system is the SUM of the
 *      local ones (the reference's analogue: thread-private partitions merged by flush!, bilinear_operator.jl:969-993).
 *      Only interface-row contributions cross NVLink (grouped ncclSend/ncclRecv between neighbouring ranks on the
 *      context's stream).  With world == 1 every call below degenerates to its single-GPU meaning.                   */
int extfem_dist_unique_id(char *id128);                 /* rank 0: ncclGetUniqueId; broadcast the 128 bytes to all ranks */
int extfem_dist_init(extfem_ctx *ctx, int rank, int world, const char *id128);
/* rows of `pattern` shared with each neighbouring rank: rows[ptr[k] .. ptr[k+1]) (1-based, local) are shared with
 * neigh_ranks[k], listed in an order both sides agree on (ascending global dof id); owned[nrows] != 0 marks the rows
 * this rank owns (every global row is owned by exactly one rank; NULL: all).                                       */
int extfem_dist_set_interfaces(extfem_ctx *ctx, int pattern, int nneigh, const int32_t *neigh_ranks, const int64_t *ptr,
                               const int64_t *rows, const uint8_t *owned);
/* device-resident rhs: b_i <- sum over the ranks sharing row i of their b_i (afterwards b is consistent) */
int extfem_dist_sum_rhs(extfem_ctx *ctx, int pattern);
/* y = A x for the sharded matrix: local SpMV + interface sum; x and y are consistent local vectors */
int extfem_dist_spmv(extfem_ctx *ctx, int pattern, const double *x, double *y);
/* Jacobi-preconditioned CG on the sharded system; b == NULL uses the (consistent) device-resident rhs */
int extfem_dist_cg(extfem_ctx *ctx, int pattern, const double *b, double *x, double rtol, int maxit, int *iters,
                   double *relres);

/* ---- owned-row form (north_star: "owned-row assembly ... NCCL exchanges only interface-row contributions"): every dof has ONE
 *      owner rank; the local mesh of a rank carries one layer of zero-volume ghost cells (host/dist.py: OwnedShard), so the
 *      column of a shared dof has the same rows, in the same order, on every rank that holds it.  After the local assemble calls,
 *      extfem_dist_reduce_system sends the non-owners' column segments (matrix) and rows (rhs) of shared dofs to the owner, which
 *      adds them: the owner holds the complete, merged rows -- what flush! produces from the partitions in the reference
 *      (bilinear_operator.jl:969-993).  Rows are per neighbour in an order both sides agree on (ascending global dof id):
 *        red_send   my NON-owned rows that receive contributions of my own cells   -> their owner
 *        red_recv   my owned rows the neighbour contributes to
 *        halo_send  my owned rows of which the neighbour holds a copy (for SpMV)   / halo_recv  my copies of the neighbour's rows */
int extfem_dist_set_owned(extfem_ctx *ctx, int pattern, int nneigh, const int32_t *neigh_ranks, const int64_t *red_send_ptr,
                          const int64_t *red_send_rows, const int64_t *red_recv_ptr, const int64_t *red_recv_rows,
                          const int64_t *halo_send_ptr, const int64_t *halo_send_rows, const int64_t *halo_recv_ptr,
                          const int64_t *halo_recv_rows, const uint8_t *owned /*[nrows]*/);
/* interface contributions of the device-resident matrix (column segments) and / or rhs to their owners.  The matrix reduction
 * runs on the context's exchange stream: an extfem_assemble_linear issued next overlaps with it; every other call (and
 * extfem_synchronize) is ordered behind it.                                                                            */
int extfem_dist_reduce_system(extfem_ctx *ctx, int pattern, int matrix, int rhs);
/* y = A x on owned rows (0 elsewhere) after a halo exchange of x; symmetric matrices (the CSC arrays are traversed as CSR) */
int extfem_dist_spmv_owned(extfem_ctx *ctx, int pattern, const double *x, double *y);
/* Jacobi-preconditioned CG on the owned-row system (SPD); b == NULL: device-resident rhs; x comes back consistent on all local dofs */
int extfem_dist_cg_owned(extfem_ctx *ctx, int pattern, const double *b, double *x, double rtol, int maxit, int *iters,
                         double *relres);

#ifdef __cplusplus
}
#endif
#endif /* EXTFEM_CUDA_H */
