/*
 * oracle/assembly_ref.c -- CPU restatement of ExtendableFEM.jl's cell assembly loops.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under extendablefem.jl_b200/ may include, link or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / CPU baseline.
 *
 * PARITY STATUS: "partially pinned".  The reference (Julia) cannot run in the build
 * container and its arithmetic lives in three un-vendored packages (ExtendableFEMBase,
 * ExtendableGrids, ExtendableSparse).  This file restates the loop nests
 * that ARE in /root/reference, in the reference's accumulation order:
 *
 *   ora_assemble_bilinear   <- src/common_operators/bilinear_operator.jl:820-951 (no args)
 *                              and :451-596 (with args)
 *   ora_assemble_linear     <- src/common_operators/linear_operator.jl:584-640 (no args)
 *                              and :359-438 (with args)
 *   ora_assemble_nonlinear  <- src/common_operators/nonlinear_operator.jl:283-436
 *   ora_csc_insert_sorted   <- sequential rawupdateindex!(A, +, v, i, j) accumulation
 *                              (bilinear_operator.jl:926) followed by flush! (:993)
 *
 * The pieces the reference delegates to ExtendableFEMBase (update_basis!, eval_trafo!,
 * quadrature tables) are restated from their published definitions for affine simplices
 * and H1 Lagrange elements; tables are passed in by oracle/fetables.py.  The restatement
 * is pinned against the reference's own golden value for Example201
 * (examples/Example201_PoissonProblem.jl:80) in tests/test_oracle_golden.py.
 *
 * Local Jacobians of nonlinear kernels: the reference uses ForwardDiff
 * (nonlinear_operator.jl:358-365), exact up to rounding.  Here they are obtained by
 * complex-step differentiation of the same kernel source (also exact up to rounding).
 */
#include <complex.h>
#undef I /* keep the identifier free; use _Complex_I */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORA_MAXOP 32   /* max total operator length (input/result vector)      */
#define ORA_MAXARGS 4  /* max number of (unknown, operator) pairs per role      */

enum { ORA_OP_ID = 0, ORA_OP_GRAD = 1, ORA_OP_DIV = 2, ORA_OP_SYMGRAD_VOIGT = 3 };

typedef struct {
    int dim;
    int64_t ncells, nnodes;
    const double *coords;      /* [nnodes][dim]               */
    const int32_t *cellnodes;  /* [ncells][dim+1], 1-based    */
    const int32_t *cellregions;/* [ncells]                    */
    const double *cellvolumes; /* [ncells]                    */
    int tdim;                  /* topological dimension of the items: dim (cells) or dim-1 (boundary faces:
                                  xgrid[BFaceNodes/BFaceRegions/BFaceVolumes], bilinear_operator.jl:693-714) */
} ora_mesh;

/* one (FESpace, FunctionOperator) pair == one FEEvaluator of the reference */
typedef struct {
    int ncomp, nscalar, op;
    double offdiag;            /* SymmetricGradient off-diagonal factor */
    const int32_t *celldofs;   /* [ncells][ncomp*nscalar], 1-based, block-local */
    int64_t offset;            /* block offset in the global system (FE.offset) */
    const double *refvals;     /* [nq][nscalar]       */
    const double *refgrads;    /* [nq][nscalar][dim]  */
} ora_arg;

typedef struct {
    double x[3];
    double time, volume;
    int region;
    int64_t item;
    const double *params;
    int nparams;
} ora_qpinfo;

typedef struct {               /* COO sink in the reference's insertion order */
    int64_t *I, *J;
    double *V;
    int64_t n, cap;
} ora_coo;

/* Test aid: when set, every TERM of every sum (basis evaluations, solution coefficients, kernel results, Jacobian entries) enters
 * by its absolute value, so an assembly returns, entry by entry, the sum of |terms| behind that entry: the scale of a
 * componentwise backward-error comparison |fl(sum t_i) - sum t_i| <= c * eps * sum |t_i| between two summation orders
 * (tests/util.py: check_values_entrywise). */
static int g_abs_accumulate = 0;
void ora_set_abs_accumulate(int on) { g_abs_accumulate = on; }
#define ORA_ACC(v) (g_abs_accumulate ? fabs(v) : (v))
static void ora_abs_vec(double *v, int n) { if (g_abs_accumulate) for (int i = 0; i < n; ++i) v[i] = fabs(v[i]); }

static int arg_ndofs(const ora_arg *a) { return a->ncomp * a->nscalar; }
static int arg_oplen(const ora_arg *a, int dim)
{
    switch (a->op) {
    case ORA_OP_ID: return a->ncomp;
    case ORA_OP_GRAD: return a->ncomp * dim;
    case ORA_OP_DIV: return 1;
    case ORA_OP_SYMGRAD_VOIGT: return dim == 2 ? 3 : (dim == 3 ? 6 : 1);
    }
    return 0;
}

/* ---- geometry: L2GTransformer for affine simplices ------------------------------ */
typedef struct { double x0[3]; double A[3][3]; double Ainv[3][3]; double det; } ora_trafo;

static void update_trafo(ora_trafo *T, const ora_mesh *m, int64_t cell)
{
    int dim = m->dim, tdim = m->tdim;
    const int32_t *cn = m->cellnodes + cell * (tdim + 1);
    const double *p0 = m->coords + (int64_t)(cn[0] - 1) * dim;
    for (int d = 0; d < dim; ++d) T->x0[d] = p0[d];
    for (int r = 0; r < tdim; ++r) {
        const double *pr = m->coords + (int64_t)(cn[r + 1] - 1) * dim;
        for (int d = 0; d < dim; ++d) T->A[d][r] = pr[d] - p0[d];
    }
    if (tdim < dim) return; /* boundary faces: only the affine map is needed (Identity operators) */
    if (dim == 1) {
        T->det = T->A[0][0];
        T->Ainv[0][0] = 1.0 / T->A[0][0];
    } else if (dim == 2) {
        double a = T->A[0][0], b = T->A[0][1], c = T->A[1][0], d = T->A[1][1];
        T->det = a * d - b * c;
        T->Ainv[0][0] = d / T->det;  T->Ainv[0][1] = -b / T->det;
        T->Ainv[1][0] = -c / T->det; T->Ainv[1][1] = a / T->det;
    } else {
        double (*A)[3] = T->A;
        double c00 = A[1][1] * A[2][2] - A[1][2] * A[2][1];
        double c01 = A[1][2] * A[2][0] - A[1][0] * A[2][2];
        double c02 = A[1][0] * A[2][1] - A[1][1] * A[2][0];
        T->det = A[0][0] * c00 + A[0][1] * c01 + A[0][2] * c02;
        double id = 1.0 / T->det;
        T->Ainv[0][0] = c00 * id;
        T->Ainv[1][0] = c01 * id;
        T->Ainv[2][0] = c02 * id;
        T->Ainv[0][1] = (A[0][2] * A[2][1] - A[0][1] * A[2][2]) * id;
        T->Ainv[1][1] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) * id;
        T->Ainv[2][1] = (A[0][1] * A[2][0] - A[0][0] * A[2][1]) * id;
        T->Ainv[0][2] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) * id;
        T->Ainv[1][2] = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) * id;
        T->Ainv[2][2] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) * id;
    }
}

static void eval_trafo(double *x, const ora_trafo *T, const double *xref, int dim, int tdim)
{
    for (int d = 0; d < dim; ++d) {
        double s = T->x0[d];
        for (int r = 0; r < tdim; ++r) s += T->A[d][r] * xref[r];
        x[d] = s;
    }
}

/* ---- update_basis!: cvals[d, j, qp] for one evaluator on one cell ----------------- */
/* layout: cvals[(qp*ndofs + j)*oplen + d]  (Julia cvals[d,j,qp], column-major)        */
static void update_basis(double *cvals, const ora_arg *a, const ora_trafo *T, int dim, int nq)
{   /* on boundary faces (tdim < dim) only Identity operators are evaluated (refgrads unused) */
    int ndofs = arg_ndofs(a), oplen = arg_oplen(a, dim), ns = a->nscalar;
    memset(cvals, 0, sizeof(double) * (size_t)nq * ndofs * oplen);
    for (int qp = 0; qp < nq; ++qp) {
        for (int k = 0; k < ns; ++k) {
            double g[3] = {0, 0, 0};
            if (a->op != ORA_OP_ID) {
                const double *rg = a->refgrads + ((size_t)qp * ns + k) * dim;
                for (int d = 0; d < dim; ++d) {
                    double s = 0;
                    for (int r = 0; r < dim; ++r) s += T->Ainv[r][d] * rg[r];
                    g[d] = s;
                }
            }
            for (int c = 0; c < a->ncomp; ++c) {
                double *cv = cvals + ((size_t)qp * ndofs + (c * ns + k)) * oplen;
                switch (a->op) {
                case ORA_OP_ID: cv[c] = a->refvals[(size_t)qp * ns + k]; break;
                case ORA_OP_GRAD: for (int d = 0; d < dim; ++d) cv[c * dim + d] = g[d]; break;
                case ORA_OP_DIV: cv[0] = g[c]; break;
                case ORA_OP_SYMGRAD_VOIGT:
                    if (dim == 2) {
                        cv[c] = g[c];
                        cv[2] = a->offdiag * g[1 - c];
                    } else {
                        cv[c] = g[c];
                        /* Voigt order 23, 13, 12 */
                        if (c == 0) { cv[4] = a->offdiag * g[2]; cv[5] = a->offdiag * g[1]; }
                        if (c == 1) { cv[3] = a->offdiag * g[2]; cv[5] = a->offdiag * g[0]; }
                        if (c == 2) { cv[3] = a->offdiag * g[1]; cv[4] = a->offdiag * g[0]; }
                    }
                    break;
                }
            }
        }
    }
    ora_abs_vec(cvals, nq * ndofs * oplen);
}

/* =====================================================================================
 * Kernel registry (restated from the reference's examples / docs)
 * ===================================================================================== */
enum {
    ORA_BLK_STANDARD = 1,  /* ExtendableFEMBase.standard_kernel: result .= input (bilinear_operator.jl:248) */
    ORA_BLK_DCR = 2,       /* examples/Example220_ReactionConvectionDiffusion.jl:66-72 ; params alpha,nu,beta[dim] */
    ORA_BLK_STOKES = 3,    /* docs/src/bilinearoperator.md:37-45 ; params mu                      */
    ORA_BLK_LINNSE7 = 4,   /* test/test_nonlinear_operator.jl:18-28 ; params mu, alpha            */
    ORA_BLK_HOOKE_GRAD = 5,/* isotropic Hooke on grad(u): mu(G+G^T)+lambda tr(G) I ; params mu,lambda */
    ORA_BLK_HOOKE_VOIGT = 6,/* examples/Example312_PeriodicElasticity3D.jl:55 sigma = C*eps ; params C row-major */
    ORA_BLK_CONVECT_ARGS = 7,/* with-args kernel: (beta=u_args . grad)u ; test kernel for :451-596 */
    ORA_BLK_ROBIN108 = 8     /* examples/Example108_RobinBoundaryCondition.jl:48-51 ; params g     */
};
enum {
    ORA_LIN_CONSTANT_ONE = 1, /* ExtendableFEMBase.constant_one_kernel (linear_operator.jl:159) */
    ORA_LIN_CONSTANT_PARAMS = 2, /* result .= params (Example330 apply_force!, Example312 linear_kernel!) */
    ORA_LIN_XY = 3,           /* README.md:37-40, Example201:32-35  f = x*y                      */
    ORA_LIN_SINCOS301 = 4,    /* examples/Example301_PoissonProblem.jl:33-35                     */
    ORA_LIN_TABULATED = 5,    /* values supplied per (cell, qp, component)                       */
    ORA_LIN_EXP2X = 6,        /* examples/Example108_RobinBoundaryCondition.jl:31-34             */
    ORA_LIN_STEP105 = 7       /* examples/Example105_NonlinearPoissonEquation.jl:35-38           */
};
enum {
    ORA_NL_NSE2D = 1,      /* examples/Example250_NSELidDrivenCavity.jl:59-74 ; params mu        */
    ORA_NL_LINNSE7 = 2,    /* test/test_nonlinear_operator.jl:18-28 ; params mu, alpha           */
    ORA_NL_NEOHOOKE3D = 3, /* examples/Example330_HyperElasticity.jl:49-57 (DW) ; params mu,lambda */
    ORA_NL_RCD = 4,        /* examples/Example108_RobinBoundaryCondition.jl:40-45 (any dim: u*du/dx1+u ; grad) */
    ORA_NL_NLPOISSON105 = 5, /* examples/Example105_NonlinearPoissonEquation.jl:45-50 ; params eps */
    ORA_NL_STVENANT230 = 6, /* examples/Example230_NonlinearElasticity.jl:39-72 ; params R, lambda[R], mu[R], epsT[R] */
    ORA_NL_POROUS106 = 7   /* examples/Example106_NonlinearDiffusion.jl:47-52 ; params m */
};
enum {
    ORA_II_STANDARD = 1,   /* ExtendableFEMBase.standard_kernel (item_integrator.jl:78-81)              */
    ORA_II_L2NORM = 2,     /* l2norm_kernel (item_integrator.jl:26-28)                                  */
    ORA_II_L2DIFF_TABULATED = 3, /* (ref - input)^2, ref supplied per (cell, qp, component)              */
    ORA_II_L2ERR_SINCOS301 = 4,  /* examples/Example301_PoissonProblem.jl:37-40,62-67                    */
    ORA_II_L2ERR_EXP108 = 5      /* examples/Example108_RobinBoundaryCondition.jl:35-38,86-90           */
};

static int bl_kernel(int id, int dim, double *r, const double *in, const double *args, const ora_qpinfo *qp, int oplen)
{
    const double *p = qp->params;
    switch (id) {
    case ORA_BLK_STANDARD:
        for (int d = 0; d < oplen; ++d) r[d] = in[d];
        return 0;
    case ORA_BLK_DCR: {
        double s = p[0] * in[0];
        for (int d = 0; d < dim; ++d) s += p[2 + d] * in[1 + d];
        r[0] = s;
        for (int d = 0; d < dim; ++d) r[1 + d] = p[1] * in[1 + d];
        return 0;
    }
    case ORA_BLK_STOKES: { /* input = [grad u (dim*dim), p] */
        int n = dim * dim;
        double div = 0;
        for (int d = 0; d < n; ++d) r[d] = p[0] * in[d];
        for (int c = 0; c < dim; ++c) { r[c * dim + c] -= in[n]; div += in[c * dim + c]; }
        r[n] = -div;
        return 0;
    }
    case ORA_BLK_LINNSE7: { /* u(2), grad u(4), p(1) */
        const double *u = in, *g = in + 2, *pr = in + 6;
        double mu = p[0], al = p[1];
        r[0] = g[0] + al * u[0];
        r[1] = g[2] + al * u[1];
        r[2] = mu * g[0] - pr[0];
        r[3] = mu * g[1];
        r[4] = mu * g[2];
        r[5] = mu * g[3] - pr[0];
        r[6] = -(g[0] + g[3]);
        return 0;
    }
    case ORA_BLK_HOOKE_GRAD: {
        double mu = p[0], la = p[1], tr = 0;
        for (int c = 0; c < dim; ++c) tr += in[c * dim + c];
        for (int c = 0; c < dim; ++c)
            for (int d = 0; d < dim; ++d)
                r[c * dim + d] = mu * (in[c * dim + d] + in[d * dim + c]) + (c == d ? la * tr : 0.0);
        return 0;
    }
    case ORA_BLK_HOOKE_VOIGT:
        for (int i = 0; i < oplen; ++i) {
            double s = 0;
            for (int j = 0; j < oplen; ++j) s += p[i * oplen + j] * in[j];
            r[i] = s;
        }
        return 0;
    case ORA_BLK_CONVECT_ARGS: { /* ansatz: grad u (ncomp*dim) ; args: id(beta) (dim) ; result id (ncomp) */
        int nc = oplen;
        for (int c = 0; c < nc; ++c) {
            double s = 0;
            for (int d = 0; d < dim; ++d) s += args[d] * in[c * dim + d];
            r[c] = s;
        }
        return 0;
    }
    case ORA_BLK_ROBIN108: /* result[1] = 2 - input[1] with g = params[0] */
        for (int d = 0; d < oplen; ++d) r[d] = p[0] - in[d];
        return 0;
    }
    return -1;
}

static int lin_kernel(int id, double *r, const ora_qpinfo *qp, int oplen, const double *tab)
{
    const double *p = qp->params;
    switch (id) {
    case ORA_LIN_CONSTANT_ONE: for (int d = 0; d < oplen; ++d) r[d] = 1.0; return 0;
    case ORA_LIN_CONSTANT_PARAMS: for (int d = 0; d < oplen; ++d) r[d] = p[d]; return 0;
    case ORA_LIN_XY: r[0] = qp->x[0] * qp->x[1]; return 0;
    case ORA_LIN_SINCOS301:
        r[0] = p[0] * (1.7 * 1.7 + 3.9 * 3.9) * sin(1.7 * qp->x[0]) * cos(3.9 * qp->x[1]);
        return 0;
    case ORA_LIN_TABULATED: for (int d = 0; d < oplen; ++d) r[d] = tab[d]; return 0;
    case ORA_LIN_EXP2X: r[0] = exp(2 * qp->x[0]); return 0;
    case ORA_LIN_STEP105: r[0] = qp->x[0] < 0.5 ? -1 : 1; return 0;
    }
    return -1;
}

/* nonlinear kernels, complex-typed so that the Jacobian is a complex-step derivative */
static int nl_kernel(int id, int dim, double complex *r, const double complex *in, const ora_qpinfo *qp)
{
    const double *p = qp->params;
    switch (id) {
    case ORA_NL_NSE2D: { /* u(2) | grad u (4): [d1u1,d2u1,d1u2,d2u2] | p */
        const double complex *u = in, *g = in + 2, pr = in[6];
        double mu = p[0];
        /* tmul!(v, grad_u, u): v_i = sum_j G[j,i] u_j with G column-major 2x2 view of g,
           i.e. G[j,i] = g[j + 2 i]  (src/helper_functions.jl:610-618, src/tensors.jl:120-122) */
        r[0] = g[0] * u[0] + g[1] * u[1];
        r[1] = g[2] * u[0] + g[3] * u[1];
        r[2] = mu * g[0] - pr;
        r[3] = mu * g[1];
        r[4] = mu * g[2];
        r[5] = mu * g[3] - pr;
        r[6] = -(g[0] + g[3]);
        return 0;
    }
    case ORA_NL_LINNSE7: {
        const double complex *u = in, *g = in + 2, pr = in[6];
        double mu = p[0], al = p[1];
        r[0] = g[0] + al * u[0];
        r[1] = g[2] + al * u[1];
        r[2] = mu * g[0] - pr;
        r[3] = mu * g[1];
        r[4] = mu * g[2];
        r[5] = mu * g[3] - pr;
        r[6] = -(g[0] + g[3]);
        return 0;
    }
    case ORA_NL_NEOHOOKE3D: {
        /* DW of W = mu/2 (F:F - 3 - 2 log detF) + lambda/2 (log detF)^2, F = I + grad u,
           F[1..9] indexed as in Example330:50-54.  DW = mu F + (lambda log detF - mu) d(detF)/dF / detF */
        double mu = p[0], la = p[1];
        double complex F[9];
        for (int i = 0; i < 9; ++i) F[i] = in[i];
        F[0] += 1; F[4] += 1; F[8] += 1;
        double complex det = -(F[2] * (F[4] * F[6] - F[3] * F[7]) + F[1] * (-F[5] * F[6] + F[3] * F[8]) +
                               F[0] * (F[5] * F[7] - F[4] * F[8]));
        double complex dd[9];
        dd[0] = -(F[5] * F[7] - F[4] * F[8]);
        dd[1] = -(-F[5] * F[6] + F[3] * F[8]);
        dd[2] = -(F[4] * F[6] - F[3] * F[7]);
        dd[3] = -(-F[2] * F[7] + F[1] * F[8]);
        dd[4] = -(F[2] * F[6] - F[0] * F[8]);
        dd[5] = -(-F[1] * F[6] + F[0] * F[7]);
        dd[6] = -(F[2] * F[4] - F[1] * F[5]);
        dd[7] = -(-F[2] * F[3] + F[0] * F[5]);
        dd[8] = -(F[1] * F[3] - F[0] * F[4]);
        double complex c = (la * clog(det) - mu) / det;
        for (int i = 0; i < 9; ++i) r[i] = mu * F[i] + c * dd[i];
        return 0;
    }
    case ORA_NL_RCD: { /* input u, grad u (dim) ; result[0] = u*d1u + u ; result[1..] = grad u */
        r[0] = in[0] * in[1] + in[0];
        for (int d = 0; d < dim; ++d) r[1 + d] = in[1 + d];
        return 0;
    }
    case ORA_NL_NLPOISSON105: { /* u, grad u ; result[1] = exp(u) - exp(-u) ; result[2] = eps * grad u */
        r[0] = cexp(in[0]) - cexp(-in[0]);
        for (int d = 0; d < dim; ++d) r[1 + d] = p[0] * in[1 + d];
        return 0;
    }
    case ORA_NL_POROUS106: { /* input u, grad u ; result = m u^(m-1) grad u (dim components) */
        double complex um1 = (p[0] == 2.0) ? in[0] : cpow(in[0], p[0] - 1.0);
        for (int d = 0; d < dim; ++d) r[d] = p[0] * um1 * in[1 + d];
        return 0;
    }
    case ORA_NL_STVENANT230: { /* input = grad(u) as a vector, Voigt strain, isotropic stress, per-region material */
        int R = (int)p[0], reg = qp->region >= 1 && qp->region <= R ? qp->region - 1 : 0;
        double la = p[1 + reg], mu = p[1 + R + reg], eT = p[1 + 2 * R + reg];
        double complex e1 = in[0], e2 = in[3], e3 = in[1] + in[2];
        e1 += 0.5 * (in[0] * in[0] + in[2] * in[2]);
        e2 += 0.5 * (in[1] * in[1] + in[3] * in[3]);
        e3 += in[0] * in[1] + in[2] * in[3];
        e1 -= eT; e2 -= eT;
        double complex a = la * (e1 + e2) + 2 * mu * e1, b = la * (e1 + e2) + 2 * mu * e2, c = 2 * mu * e3;
        r[0] = a; r[1] = c; r[2] = c; r[3] = b;
        return 0;
    }
    }
    return -1;
}

/* scalar energy of Example330:49-57, exported so tests can check DW against dW/dF */
double ora_neohooke_energy(const double *gradu, double mu, double la)
{
    double F[9];
    for (int i = 0; i < 9; ++i) F[i] = gradu[i];
    F[0] += 1; F[4] += 1; F[8] += 1;
    double det = -(F[2] * (F[4] * F[6] - F[3] * F[7]) + F[1] * (-F[5] * F[6] + F[3] * F[8]) +
                   F[0] * (F[5] * F[7] - F[4] * F[8]));
    double ff = 0;
    for (int i = 0; i < 9; ++i) ff += F[i] * F[i];
    return mu / 2 * (ff - 3 - 2 * log(det)) + la / 2 * log(det) * log(det);
}

int ora_nl_value_and_jacobian(int id, int dim, int nin, int nout, const double *in, const double *params,
                              int nparams, double *value, double *jac /* [nout][nin] row-major */)
{
    ora_qpinfo qp;
    memset(&qp, 0, sizeof qp);
    qp.params = params; qp.nparams = nparams;
    double complex zin[ORA_MAXOP], zout[ORA_MAXOP];
    for (int i = 0; i < nin; ++i) zin[i] = in[i];
    if (nl_kernel(id, dim, zout, zin, &qp)) return -1;
    for (int k = 0; k < nout; ++k) value[k] = creal(zout[k]);
    const double h = 1e-40;
    for (int j = 0; j < nin; ++j) {
        zin[j] = in[j] + h * _Complex_I;
        nl_kernel(id, dim, zout, zin, &qp);
        for (int k = 0; k < nout; ++k) jac[k * nin + j] = cimag(zout[k]) / h;
        zin[j] = in[j];
    }
    return 0;
}

/* ---- COO sink ------------------------------------------------------------------- */
static int coo_push(ora_coo *s, int64_t i, int64_t j, double v)
{
    if (s->n >= s->cap) return -1;
    s->I[s->n] = i; s->J[s->n] = j; s->V[s->n] = v; s->n++;
    return 0;
}

/* optional direct CSC insertion (rawupdateindex! into an existing pattern) for timing */
typedef struct { const int64_t *colptr; const int64_t *rowval; double *nzval; } ora_csc;
static int csc_add(ora_csc *A, int64_t i, int64_t j, double v)
{
    int64_t lo = A->colptr[j - 1] - 1, hi = A->colptr[j] - 2;
    while (lo <= hi) {
        int64_t mid = (lo + hi) >> 1;
        int64_t r = A->rowval[mid];
        if (r == i) { A->nzval[mid] += v; return 0; }
        if (r < i) lo = mid + 1; else hi = mid - 1;
    }
    return -2;
}


typedef struct { ora_coo *coo; ora_csc *csc; } ora_sink;
static int sink_add(ora_sink *s, int64_t i, int64_t j, double v)
{
    v = ORA_ACC(v);
    if (s->csc) return csc_add(s->csc, i, j, v);
    return coo_push(s->coo, i, j, v);
}

static int region_visited(const int32_t *regions, int nregions, int r)
{
    if (nregions == 0) return 1;
    for (int k = 0; k < nregions; ++k) if (regions[k] == r) return 1;
    return 0;
}

/* =====================================================================================
 * BilinearOperator assembly_loop  (bilinear_operator.jl:820-951 / :451-596)
 * ===================================================================================== */
int ora_assemble_bilinear(const ora_mesh *m, int ntest, const ora_arg *test, int nansatz, const ora_arg *ansatz,
                          int nargs, const ora_arg *args, const double *sol /* global entries, may be NULL */,
                          const int64_t *args_sol_offsets,
                          const uint8_t *coupling /* [nansatz][ntest] */, int nq, const double *qw, const double *qx,
                          int kernel_id, const double *params, int nparams, double factor, double time,
                          const int32_t *regions, int nregions, int transposed_copy, int lump, double entry_tol,
                          ora_coo *coo, ora_csc *csc)
{
    int dim = m->dim;
    ora_sink sink = {coo, csc};
    int oplen_t[ORA_MAXARGS], oplen_a[ORA_MAXARGS], oplen_g[ORA_MAXARGS];
    int off_t[ORA_MAXARGS + 1] = {0}, off_a[ORA_MAXARGS + 1] = {0}, off_g[ORA_MAXARGS + 1] = {0};
    int nd_t[ORA_MAXARGS], nd_a[ORA_MAXARGS], nd_g[ORA_MAXARGS];
    double *cv_t[ORA_MAXARGS], *cv_a[ORA_MAXARGS], *cv_g[ORA_MAXARGS];
    double *Aloc[ORA_MAXARGS][ORA_MAXARGS];
    for (int j = 0; j < ntest; ++j) {
        oplen_t[j] = arg_oplen(&test[j], dim); off_t[j + 1] = off_t[j] + oplen_t[j]; nd_t[j] = arg_ndofs(&test[j]);
        cv_t[j] = malloc(sizeof(double) * (size_t)nq * nd_t[j] * oplen_t[j]);
    }
    for (int j = 0; j < nansatz; ++j) {
        oplen_a[j] = arg_oplen(&ansatz[j], dim); off_a[j + 1] = off_a[j] + oplen_a[j]; nd_a[j] = arg_ndofs(&ansatz[j]);
        cv_a[j] = malloc(sizeof(double) * (size_t)nq * nd_a[j] * oplen_a[j]);
    }
    for (int j = 0; j < nargs; ++j) {
        oplen_g[j] = arg_oplen(&args[j], dim); off_g[j + 1] = off_g[j] + oplen_g[j]; nd_g[j] = arg_ndofs(&args[j]);
        cv_g[j] = malloc(sizeof(double) * (size_t)nq * nd_g[j] * oplen_g[j]);
    }
    for (int j = 0; j < ntest; ++j)
        for (int k = 0; k < nansatz; ++k) Aloc[j][k] = calloc((size_t)nd_t[j] * nd_a[k], sizeof(double));
    double input_ansatz[ORA_MAXOP], input_args[ORA_MAXOP], result[ORA_MAXOP];
    ora_qpinfo qp;
    memset(&qp, 0, sizeof qp);
    qp.params = params; qp.nparams = nparams; qp.time = time;
    ora_trafo T;
    int rc = 0;
    for (int64_t item = 0; item < m->ncells && !rc; ++item) {
        int reg = m->cellregions[item];
        if (reg > 0) { if (!region_visited(regions, nregions, reg)) continue; }
        else if (nregions > 0 && nargs == 0) continue;
        qp.region = reg; qp.item = item + 1; qp.volume = m->cellvolumes[item];
        update_trafo(&T, m, item);
        for (int j = 0; j < ntest; ++j) update_basis(cv_t[j], &test[j], &T, dim, nq);
        for (int j = 0; j < nansatz; ++j) update_basis(cv_a[j], &ansatz[j], &T, dim, nq);
        for (int j = 0; j < nargs; ++j) update_basis(cv_g[j], &args[j], &T, dim, nq);
        for (int q = 0; q < nq; ++q) {
            if (nargs > 0) {
                for (int d = 0; d < off_g[nargs]; ++d) input_args[d] = 0;
                for (int id = 0; id < nargs; ++id)
                    for (int j = 0; j < nd_g[id]; ++j) {
                        int64_t dof = args[id].celldofs[item * nd_g[id] + j] - 1 + args_sol_offsets[id];
                        for (int d = 0; d < oplen_g[id]; ++d)
                            input_args[d + off_g[id]] += ORA_ACC(sol[dof]) * cv_g[id][((size_t)q * nd_g[id] + j) * oplen_g[id] + d];
                    }
            }
            eval_trafo(qp.x, &T, qx + (size_t)q * m->tdim, dim, m->tdim);
            for (int id = 0; id < nansatz; ++id) {
                for (int j = 0; j < nd_a[id]; ++j) {
                    for (int d = 0; d < off_a[nansatz]; ++d) input_ansatz[d] = 0;
                    for (int d = 0; d < oplen_a[id]; ++d)
                        input_ansatz[d + off_a[id]] = cv_a[id][((size_t)q * nd_a[id] + j) * oplen_a[id] + d];
                    if (bl_kernel(kernel_id, dim, result, input_ansatz, input_args, &qp, off_t[ntest])) { rc = -3; goto done; }
                    ora_abs_vec(result, off_t[ntest]);
                    for (int d = 0; d < off_t[ntest]; ++d) result[d] *= factor * qw[q];
                    if (lump == 1) {
                        for (int d = 0; d < oplen_t[id]; ++d)
                            Aloc[id][id][j * nd_a[id] + j] += result[d + off_t[id]] * cv_t[id][((size_t)q * nd_t[id] + j) * oplen_t[id] + d];
                    } else if (lump == 2) {
                        for (int k = 0; k < nd_t[id]; ++k)
                            for (int d = 0; d < oplen_t[id]; ++d)
                                Aloc[id][id][j * nd_a[id] + j] += result[d + off_t[id]] * cv_t[id][((size_t)q * nd_t[id] + k) * oplen_t[id] + d];
                    } else {
                        for (int idt = 0; idt < ntest; ++idt) {
                            if (!coupling[id * ntest + idt]) continue;
                            for (int k = 0; k < nd_t[idt]; ++k)
                                for (int d = 0; d < oplen_t[idt]; ++d)
                                    Aloc[idt][id][k * nd_a[id] + j] += result[d + off_t[idt]] * cv_t[idt][((size_t)q * nd_t[idt] + k) * oplen_t[idt] + d];
                        }
                    }
                }
            }
        }
        /* add local matrices to the global matrix (:919-930) */
        for (int id = 0; id < nansatz; ++id)
            for (int idt = 0; idt < ntest; ++idt) {
                if (!coupling[id * ntest + idt] && nargs == 0) continue;
                double *L = Aloc[idt][id];
                for (int e = 0; e < nd_t[idt] * nd_a[id]; ++e) L[e] *= m->cellvolumes[item];
                for (int j = 0; j < nd_t[idt]; ++j) {
                    int64_t dof_j = test[idt].celldofs[item * nd_t[idt] + j] + test[idt].offset;
                    for (int k = 0; k < nd_a[id]; ++k) {
                        int64_t dof_k = ansatz[id].celldofs[item * nd_a[id] + k] + ansatz[id].offset;
                        if (fabs(L[j * nd_a[id] + k]) > entry_tol)
                            if (sink_add(&sink, dof_j, dof_k, L[j * nd_a[id] + k])) { rc = -2; goto done; }
                    }
                }
            }
        if (transposed_copy != 0)
            for (int id = 0; id < nansatz; ++id)
                for (int idt = 0; idt < ntest; ++idt) {
                    if (!coupling[id * ntest + idt] && nargs == 0) continue;
                    double *L = Aloc[idt][id];
                    for (int e = 0; e < nd_t[idt] * nd_a[id]; ++e) L[e] *= transposed_copy;
                    for (int j = 0; j < nd_t[idt]; ++j) {
                        int64_t dof_j = test[idt].celldofs[item * nd_t[idt] + j] + test[idt].offset;
                        for (int k = 0; k < nd_a[id]; ++k) {
                            int64_t dof_k = ansatz[id].celldofs[item * nd_a[id] + k] + ansatz[id].offset;
                            if (fabs(L[j * nd_a[id] + k]) > entry_tol)
                                if (sink_add(&sink, dof_k, dof_j, L[j * nd_a[id] + k])) { rc = -2; goto done; }
                        }
                    }
                }
        for (int id = 0; id < nansatz; ++id)
            for (int idt = 0; idt < ntest; ++idt) memset(Aloc[idt][id], 0, sizeof(double) * (size_t)nd_t[idt] * nd_a[id]);
    }
done:
    for (int j = 0; j < ntest; ++j) free(cv_t[j]);
    for (int j = 0; j < nansatz; ++j) free(cv_a[j]);
    for (int j = 0; j < nargs; ++j) free(cv_g[j]);
    for (int j = 0; j < ntest; ++j) for (int k = 0; k < nansatz; ++k) free(Aloc[j][k]);
    return rc;
}

/* =====================================================================================
 * LinearOperator assembly_loop  (linear_operator.jl:584-640 / :359-438)
 * b[dof] is accumulated directly, qp by qp, pre-scaled by factor*w*|T| (:626,:633)
 * With nargs > 0 the kernel is evaluated on input_args (standard kernel: result = input_args).
 * ===================================================================================== */
int ora_assemble_linear(const ora_mesh *m, int ntest, const ora_arg *test, int nargs, const ora_arg *args,
                        const double *sol, const int64_t *args_sol_offsets, int nq, const double *qw, const double *qx,
                        int kernel_id, const double *params, int nparams, double factor, double time,
                        const int32_t *regions, int nregions, const double *tabulated /* [ncells][nq][oplen] */,
                        double *b)
{
    int dim = m->dim;
    int oplen_t[ORA_MAXARGS], oplen_g[ORA_MAXARGS], off_t[ORA_MAXARGS + 1] = {0}, off_g[ORA_MAXARGS + 1] = {0};
    int nd_t[ORA_MAXARGS], nd_g[ORA_MAXARGS];
    double *cv_t[ORA_MAXARGS], *cv_g[ORA_MAXARGS];
    for (int j = 0; j < ntest; ++j) {
        oplen_t[j] = arg_oplen(&test[j], dim); off_t[j + 1] = off_t[j] + oplen_t[j]; nd_t[j] = arg_ndofs(&test[j]);
        cv_t[j] = malloc(sizeof(double) * (size_t)nq * nd_t[j] * oplen_t[j]);
    }
    for (int j = 0; j < nargs; ++j) {
        oplen_g[j] = arg_oplen(&args[j], dim); off_g[j + 1] = off_g[j] + oplen_g[j]; nd_g[j] = arg_ndofs(&args[j]);
        cv_g[j] = malloc(sizeof(double) * (size_t)nq * nd_g[j] * oplen_g[j]);
    }
    double input_args[ORA_MAXOP], result[ORA_MAXOP];
    ora_qpinfo qp;
    memset(&qp, 0, sizeof qp);
    qp.params = params; qp.nparams = nparams; qp.time = time;
    ora_trafo T;
    int rc = 0, oplen = off_t[ntest];
    for (int64_t item = 0; item < m->ncells && !rc; ++item) {
        int reg = m->cellregions[item];
        if (reg > 0) { if (!region_visited(regions, nregions, reg)) continue; }
        else if (nregions > 0 && nargs == 0) continue;
        qp.region = reg; qp.item = item + 1; qp.volume = m->cellvolumes[item];
        update_trafo(&T, m, item);
        for (int j = 0; j < ntest; ++j) update_basis(cv_t[j], &test[j], &T, dim, nq);
        for (int j = 0; j < nargs; ++j) update_basis(cv_g[j], &args[j], &T, dim, nq);
        for (int q = 0; q < nq; ++q) {
            if (nargs > 0) {
                for (int d = 0; d < off_g[nargs]; ++d) input_args[d] = 0;
                for (int id = 0; id < nargs; ++id)
                    for (int j = 0; j < nd_g[id]; ++j) {
                        int64_t dof = args[id].celldofs[item * nd_g[id] + j] - 1 + args_sol_offsets[id];
                        for (int d = 0; d < oplen_g[id]; ++d)
                            input_args[d + off_g[id]] += ORA_ACC(sol[dof]) * cv_g[id][((size_t)q * nd_g[id] + j) * oplen_g[id] + d];
                    }
            }
            eval_trafo(qp.x, &T, qx + (size_t)q * m->tdim, dim, m->tdim);
            if (nargs > 0) {
                if (kernel_id != ORA_BLK_STANDARD) { rc = -3; break; }
                for (int d = 0; d < oplen; ++d) result[d] = input_args[d];
            } else if (lin_kernel(kernel_id, result, &qp, oplen,
                                  tabulated ? tabulated + ((size_t)item * nq + q) * oplen : NULL)) { rc = -3; break; }
            ora_abs_vec(result, oplen);
            for (int d = 0; d < oplen; ++d) result[d] *= factor * qw[q] * m->cellvolumes[item];
            for (int idt = 0; idt < ntest; ++idt)
                for (int k = 0; k < nd_t[idt]; ++k) {
                    int64_t dof = test[idt].celldofs[item * nd_t[idt] + k] - 1 + test[idt].offset;
                    for (int d = 0; d < oplen_t[idt]; ++d)
                        b[dof] += ORA_ACC(result[d + off_t[idt]] * cv_t[idt][((size_t)q * nd_t[idt] + k) * oplen_t[idt] + d]);
                }
        }
    }
    for (int j = 0; j < ntest; ++j) free(cv_t[j]);
    for (int j = 0; j < nargs; ++j) free(cv_g[j]);
    return rc;
}

/* =====================================================================================
 * NonlinearOperator assembly_loop  (nonlinear_operator.jl:283-436), dense local Jacobian
 * path (:385-389; the sparse path :377-383 visits the same products minus exact zeros).
 * ===================================================================================== */
int ora_assemble_nonlinear(const ora_mesh *m, int ntest, const ora_arg *test, int nargs, const ora_arg *args,
                           const double *sol, const int64_t *args_sol_offsets, int nq, const double *qw, const double *qx,
                           int kernel_id, const double *params, int nparams, double factor, double time,
                           const int32_t *regions, int nregions, double entry_tol, ora_coo *coo, ora_csc *csc, double *b)
{
    int dim = m->dim;
    ora_sink sink = {coo, csc};
    int oplen_t[ORA_MAXARGS], oplen_g[ORA_MAXARGS], off_t[ORA_MAXARGS + 1] = {0}, off_g[ORA_MAXARGS + 1] = {0};
    int nd_t[ORA_MAXARGS], nd_g[ORA_MAXARGS];
    double *cv_t[ORA_MAXARGS], *cv_g[ORA_MAXARGS];
    double *Aloc[ORA_MAXARGS][ORA_MAXARGS];
    for (int j = 0; j < ntest; ++j) {
        oplen_t[j] = arg_oplen(&test[j], dim); off_t[j + 1] = off_t[j] + oplen_t[j]; nd_t[j] = arg_ndofs(&test[j]);
        cv_t[j] = malloc(sizeof(double) * (size_t)nq * nd_t[j] * oplen_t[j]);
    }
    for (int j = 0; j < nargs; ++j) {
        oplen_g[j] = arg_oplen(&args[j], dim); off_g[j + 1] = off_g[j] + oplen_g[j]; nd_g[j] = arg_ndofs(&args[j]);
        cv_g[j] = malloc(sizeof(double) * (size_t)nq * nd_g[j] * oplen_g[j]);
    }
    for (int j = 0; j < ntest; ++j)
        for (int k = 0; k < nargs; ++k) Aloc[j][k] = calloc((size_t)nd_t[j] * nd_g[k], sizeof(double));
    int nin = off_g[nargs], nout = off_t[ntest];
    double input_args[ORA_MAXOP], value[ORA_MAXOP], tempV[ORA_MAXOP], jac[ORA_MAXOP * ORA_MAXOP];
    ora_qpinfo qp;
    memset(&qp, 0, sizeof qp);
    qp.params = params; qp.nparams = nparams; qp.time = time;
    ora_trafo T;
    int rc = 0;
    for (int64_t item = 0; item < m->ncells && !rc; ++item) {
        int reg = m->cellregions[item];
        if (reg > 0 && !region_visited(regions, nregions, reg)) continue;
        qp.region = reg; qp.item = item + 1; qp.volume = m->cellvolumes[item];
        double vol = m->cellvolumes[item];
        update_trafo(&T, m, item);
        for (int j = 0; j < ntest; ++j) update_basis(cv_t[j], &test[j], &T, dim, nq);
        for (int j = 0; j < nargs; ++j) update_basis(cv_g[j], &args[j], &T, dim, nq);
        for (int q = 0; q < nq; ++q) {
            for (int d = 0; d < nin; ++d) input_args[d] = 0;
            for (int id = 0; id < nargs; ++id)
                for (int j = 0; j < nd_g[id]; ++j) {
                    int64_t dof = args[id].celldofs[item * nd_g[id] + j] - 1 + args_sol_offsets[id];
                    for (int d = 0; d < oplen_g[id]; ++d)
                        input_args[d + off_g[id]] += ORA_ACC(sol[dof]) * cv_g[id][((size_t)q * nd_g[id] + j) * oplen_g[id] + d];
                }
            eval_trafo(qp.x, &T, qx + (size_t)q * m->tdim, dim, m->tdim);
            {   /* value_and_jacobian! (:358-365) */
                double complex zin[ORA_MAXOP], zout[ORA_MAXOP];
                const double h = 1e-40;
                for (int i = 0; i < nin; ++i) zin[i] = input_args[i];
                if (nl_kernel(kernel_id, dim, zout, zin, &qp)) { rc = -3; goto done; }
                for (int k = 0; k < nout; ++k) value[k] = creal(zout[k]);
                for (int j = 0; j < nin; ++j) {
                    zin[j] = input_args[j] + h * _Complex_I;
                    nl_kernel(kernel_id, dim, zout, zin, &qp);
                    for (int k = 0; k < nout; ++k) jac[k * nin + j] = cimag(zout[k]) / h;
                    zin[j] = input_args[j];
                }
                ora_abs_vec(value, nout);
                ora_abs_vec(jac, nin * nout);
            }
            /* update matrix (:372-401) */
            for (int id = 0; id < nargs; ++id)
                for (int j = 0; j < nd_g[id]; ++j) {
                    for (int k = 0; k < nout; ++k) tempV[k] = 0;
                    for (int d = 0; d < oplen_g[id]; ++d)
                        for (int k = 0; k < nout; ++k)
                            tempV[k] += jac[k * nin + d + off_g[id]] * cv_g[id][((size_t)q * nd_g[id] + j) * oplen_g[id] + d];
                    for (int idt = 0; idt < ntest; ++idt)
                        for (int k = 0; k < nd_t[idt]; ++k)
                            for (int d = 0; d < oplen_t[idt]; ++d)
                                Aloc[idt][id][k * nd_g[id] + j] +=
                                    tempV[d + off_t[idt]] * cv_t[idt][((size_t)q * nd_t[idt] + k) * oplen_t[idt] + d] * qw[q];
                }
            /* update rhs (:404-414): (jac*u - value) * factor*w*|T| */
            for (int k = 0; k < nout; ++k) {
                double s = 0;
                for (int d = 0; d < nin; ++d) s += jac[k * nin + d] * input_args[d];
                tempV[k] = (g_abs_accumulate ? s + value[k] : s - value[k]) * (factor * qw[q] * vol);
            }
            for (int idt = 0; idt < ntest; ++idt)
                for (int j = 0; j < nd_t[idt]; ++j) {
                    int64_t dof = test[idt].celldofs[item * nd_t[idt] + j] - 1 + test[idt].offset;
                    for (int d = 0; d < oplen_t[idt]; ++d)
                        b[dof] += ORA_ACC(tempV[d + off_t[idt]] * cv_t[idt][((size_t)q * nd_t[idt] + j) * oplen_t[idt] + d]);
                }
        }
        for (int id = 0; id < nargs; ++id)
            for (int idt = 0; idt < ntest; ++idt) {
                double *L = Aloc[idt][id];
                for (int e = 0; e < nd_t[idt] * nd_g[id]; ++e) L[e] *= factor * vol;
                for (int j = 0; j < nd_t[idt]; ++j) {
                    int64_t dof_j = test[idt].celldofs[item * nd_t[idt] + j] + test[idt].offset;
                    for (int k = 0; k < nd_g[id]; ++k) {
                        int64_t dof_k = args[id].celldofs[item * nd_g[id] + k] + args[id].offset;
                        if (fabs(L[j * nd_g[id] + k]) > entry_tol)
                            if (sink_add(&sink, dof_j, dof_k, L[j * nd_g[id] + k])) { rc = -2; goto done; }
                    }
                }
                memset(L, 0, sizeof(double) * (size_t)nd_t[idt] * nd_g[id]);
            }
    }
done:
    for (int j = 0; j < ntest; ++j) free(cv_t[j]);
    for (int j = 0; j < nargs; ++j) free(cv_g[j]);
    for (int j = 0; j < ntest; ++j) for (int k = 0; k < nargs; ++k) free(Aloc[j][k]);
    return rc;
}

/* =====================================================================================
 * ItemIntegrator assembly_loop  (item_integrator.jl:191-249): piecewise b[1:resultdim, item] += kernel(...) * factor*w*|T|
 * ===================================================================================== */
int ora_integrate(const ora_mesh *m, int nargs, const ora_arg *args, const double *sol, const int64_t *args_sol_offsets,
                  int nq, const double *qw, const double *qx, int kernel_id, const double *params, int nparams, double factor,
                  double time, const int32_t *regions, int nregions, const double *tabulated /* [ncells][nq][resultdim] */,
                  int resultdim, double *b /* [ncells][resultdim] */)
{
    int dim = m->dim;
    int oplen_g[ORA_MAXARGS], off_g[ORA_MAXARGS + 1] = {0}, nd_g[ORA_MAXARGS];
    double *cv_g[ORA_MAXARGS];
    for (int j = 0; j < nargs; ++j) {
        oplen_g[j] = arg_oplen(&args[j], dim); off_g[j + 1] = off_g[j] + oplen_g[j]; nd_g[j] = arg_ndofs(&args[j]);
        cv_g[j] = malloc(sizeof(double) * (size_t)nq * nd_g[j] * oplen_g[j]);
    }
    int nin = off_g[nargs], rc = 0;
    double input_args[ORA_MAXOP], result[ORA_MAXOP];
    ora_qpinfo qp;
    memset(&qp, 0, sizeof qp);
    qp.params = params; qp.nparams = nparams; qp.time = time;
    ora_trafo T;
    for (int64_t item = 0; item < m->ncells && !rc; ++item) {
        int reg = m->cellregions[item];
        if (reg > 0 && !region_visited(regions, nregions, reg)) continue;
        qp.region = reg; qp.item = item + 1; qp.volume = m->cellvolumes[item];
        update_trafo(&T, m, item);
        for (int j = 0; j < nargs; ++j) update_basis(cv_g[j], &args[j], &T, dim, nq);
        for (int q = 0; q < nq; ++q) {
            for (int d = 0; d < nin; ++d) input_args[d] = 0;
            for (int id = 0; id < nargs; ++id)
                for (int j = 0; j < nd_g[id]; ++j) {
                    int64_t dof = args[id].celldofs[item * nd_g[id] + j] - 1 + args_sol_offsets[id];
                    for (int d = 0; d < oplen_g[id]; ++d)
                        input_args[d + off_g[id]] += ORA_ACC(sol[dof]) * cv_g[id][((size_t)q * nd_g[id] + j) * oplen_g[id] + d];
                }
            eval_trafo(qp.x, &T, qx + (size_t)q * m->tdim, dim, m->tdim);
            switch (kernel_id) {
            case ORA_II_STANDARD: for (int d = 0; d < resultdim; ++d) result[d] = d < nin ? input_args[d] : 0; break;
            case ORA_II_L2NORM: for (int d = 0; d < resultdim; ++d) result[d] = d < nin ? input_args[d] * input_args[d] : 0; break;
            case ORA_II_L2DIFF_TABULATED:
                for (int d = 0; d < resultdim; ++d) {
                    double e = tabulated[((size_t)item * nq + q) * resultdim + d] - input_args[d];
                    result[d] = e * e;
                }
                break;
            case ORA_II_L2ERR_SINCOS301: { double e = sin(1.7 * qp.x[0]) * cos(3.9 * qp.x[1]) - input_args[0]; result[0] = e * e; } break;
            case ORA_II_L2ERR_EXP108: { double e = exp(qp.x[0]) - input_args[0]; result[0] = e * e; } break;
            default: rc = -3;
            }
            for (int d = 0; d < resultdim; ++d) b[item * resultdim + d] += result[d] * (factor * qw[q] * m->cellvolumes[item]);
        }
    }
    for (int j = 0; j < nargs; ++j) free(cv_g[j]);
    return rc;
}

/* Sequential accumulation of COO triplets that were stably sorted by (col,row):
 * equals the reference's repeated rawupdateindex!(A,+,v,i,j) followed by flush!.   */
int64_t ora_coo_segment_sum(int64_t n, const int64_t *I, const int64_t *J, const double *V, int64_t *outI,
                            int64_t *outJ, double *outV)
{
    int64_t k = -1;
    for (int64_t e = 0; e < n; ++e) {
        if (k >= 0 && outI[k] == I[e] && outJ[k] == J[e]) outV[k] += V[e];
        else { ++k; outI[k] = I[e]; outJ[k] = J[e]; outV[k] = V[e]; }
    }
    return k + 1;
}

/* y = b - A*x  for CSC A  (compute_nonlinear_residual!, src/solvers.jl:38-43) */
void ora_residual(int64_t ncols, const int64_t *colptr, const int64_t *rowval, const double *nzval, const double *x,
                  const double *b, int64_t nrows, double *res)
{
    for (int64_t i = 0; i < nrows; ++i) res[i] = b[i];
    for (int64_t j = 0; j < ncols; ++j)
        for (int64_t p = colptr[j] - 1; p < colptr[j + 1] - 1; ++p) res[rowval[p] - 1] -= nzval[p] * x[j];
}
