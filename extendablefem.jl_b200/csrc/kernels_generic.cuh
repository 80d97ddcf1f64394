// kernels_generic.cuh -- generic (any registered kernel / element / block layout) local
// assembly kernels.  One CTA processes CPB consecutive cells cooperatively:
//
//   phase A  per (cell, qp)   : affine geometry, x_q, solution evaluation, kernel Jacobian /
//                                source value  -> shared memory                     (staging)
//   phase B  per local entry  : quadrature contraction with on-the-fly basis evaluation,
//                                written to the cell-local buffer with unit-stride stores
//
// The cell-local buffer is then reduced into the CSC values by the column-gather kernel
// (gather.cuh) without atomics.  These kernels restate the loop nests
//   bilinear_operator.jl:876-916 / :512-561, linear_operator.jl:619-637 / :408-435,
//   nonlinear_operator.jl:343-415
// with the quadrature index innermost per entry (summation order over qp is preserved).
#pragma once
#include "common.cuh"

namespace extfem {

// sparse operator evaluation of one basis function at one quadrature point: cvals[:, j, qp]
struct CV {
    int idx[3];
    double v[3];
    int len;
};

template <int DIM>
__device__ __forceinline__ CV eval_cv(const ArgDev &a, int j, int q, const double *__restrict__ Ainv, double offdiag)
{
    CV cv;
    int c = j / a.nscalar, k = j - c * a.nscalar;
    if (a.op == EXTFEM_OP_ID) {
        cv.len = 1;
        cv.idx[0] = c;
        cv.v[0] = __ldg(a.refvals + q * a.nscalar + k);
        return cv;
    }
    double g[DIM];
    const double *rg = a.refgrads + ((size_t)q * a.nscalar + k) * DIM;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
        double s = 0;
#pragma unroll
        for (int r = 0; r < DIM; ++r) s += Ainv[r * DIM + d] * __ldg(rg + r);
        g[d] = s;
    }
    if (a.op == EXTFEM_OP_GRAD) {
        cv.len = DIM;
#pragma unroll
        for (int d = 0; d < DIM; ++d) { cv.idx[d] = c * DIM + d; cv.v[d] = g[d]; }
    } else if (a.op == EXTFEM_OP_DIV) {
        cv.len = 1;
        cv.idx[0] = 0;
        cv.v[0] = g[c < DIM ? c : 0];
    } else { // SYMGRAD_VOIGT
        if (DIM == 1) { cv.len = 1; cv.idx[0] = 0; cv.v[0] = g[0]; }
        else if (DIM == 2) {
            cv.len = 2;
            cv.idx[0] = c; cv.v[0] = g[c];
            cv.idx[1] = 2; cv.v[1] = offdiag * g[(1 - c) & 1];
        } else {
            cv.len = 3;
            cv.idx[0] = c; cv.v[0] = g[c];
            // Voigt order 23, 13, 12
            int o1 = (c == 0) ? 4 : 3, o2 = (c == 2) ? 4 : 5;
            int g1 = (c == 0) ? 2 : ((c == 1) ? 2 : 1), g2 = (c == 0) ? 1 : 0;
            cv.idx[1] = o1; cv.v[1] = offdiag * g[g1 < DIM ? g1 : 0];
            cv.idx[2] = o2; cv.v[2] = offdiag * g[g2 < DIM ? g2 : 0];
        }
    }
    return cv;
}

// ---- registry: bilinear kernels (linear in `in`) ------------------------------------------
__device__ __forceinline__ void bl_apply(int id, int dim, const double *in, const double *ain, const double *p,
                                         double *r, int nout)
{
    switch (id) {
    case EXTFEM_BLK_STANDARD:
        for (int d = 0; d < nout; ++d) r[d] = in[d];
        break;
    case EXTFEM_BLK_DCR: {
        double s = p[0] * in[0];
        for (int d = 0; d < dim; ++d) s += p[2 + d] * in[1 + d];
        r[0] = s;
        for (int d = 0; d < dim; ++d) r[1 + d] = p[1] * in[1 + d];
    } break;
    case EXTFEM_BLK_STOKES: {
        int n = dim * dim;
        double div = 0;
        for (int d = 0; d < n; ++d) r[d] = p[0] * in[d];
        for (int c = 0; c < dim; ++c) { r[c * dim + c] -= in[n]; div += in[c * dim + c]; }
        r[n] = -div;
    } break;
    case EXTFEM_BLK_LINNSE7: {
        const double *u = in, *g = in + 2;
        double pr = in[6], mu = p[0], al = p[1];
        r[0] = g[0] + al * u[0];
        r[1] = g[2] + al * u[1];
        r[2] = mu * g[0] - pr;
        r[3] = mu * g[1];
        r[4] = mu * g[2];
        r[5] = mu * g[3] - pr;
        r[6] = -(g[0] + g[3]);
    } break;
    case EXTFEM_BLK_HOOKE_GRAD: {
        double mu = p[0], la = p[1], tr = 0;
        for (int c = 0; c < dim; ++c) tr += in[c * dim + c];
        for (int c = 0; c < dim; ++c)
            for (int d = 0; d < dim; ++d)
                r[c * dim + d] = mu * (in[c * dim + d] + in[d * dim + c]) + (c == d ? la * tr : 0.0);
    } break;
    case EXTFEM_BLK_HOOKE_VOIGT:
        for (int i = 0; i < nout; ++i) {
            double s = 0;
            for (int j = 0; j < nout; ++j) s += p[i * nout + j] * in[j];
            r[i] = s;
        }
        break;
    case EXTFEM_BLK_CONVECT_ARGS:
        for (int c = 0; c < nout; ++c) {
            double s = 0;
            for (int d = 0; d < dim; ++d) s += ain[d] * in[c * dim + d];
            r[c] = s;
        }
        break;
    case EXTFEM_BLK_ROBIN108:   // Example108:48-51: evaluated per basis function like every bilinear kernel (:891)
        for (int d = 0; d < nout; ++d) r[d] = p[0] - in[d];
        break;
    }
}

__device__ __forceinline__ void lin_apply(int id, const double *x, const double *p, double *r, int nout,
                                          const double *tab)
{
    switch (id) {
    case EXTFEM_LIN_CONSTANT_ONE: for (int d = 0; d < nout; ++d) r[d] = 1.0; break;
    case EXTFEM_LIN_CONSTANT_PARAMS: for (int d = 0; d < nout; ++d) r[d] = p[d]; break;
    case EXTFEM_LIN_XY: r[0] = x[0] * x[1]; break;
    case EXTFEM_LIN_SINCOS301: r[0] = p[0] * (1.7 * 1.7 + 3.9 * 3.9) * sin(1.7 * x[0]) * cos(3.9 * x[1]); break;
    case EXTFEM_LIN_TABULATED: for (int d = 0; d < nout; ++d) r[d] = tab[d]; break;
    case EXTFEM_LIN_EXP2X: r[0] = exp(2.0 * x[0]); break;
    case EXTFEM_LIN_STEP105: r[0] = x[0] < 0.5 ? -1.0 : 1.0; break;
    }
}

// ---- registry: nonlinear kernels with analytic Jacobians ----------------------------------
// value[nout], jac[nout][nin] (row-major, written densely)
__device__ __forceinline__ void nl_apply(int id, int dim, const double *in, const double *p, double *val, double *J,
                                         int nin, int nout, int region = 1)
{
#pragma unroll
    for (int i = 0; i < nin * nout; ++i) J[i] = 0.0;
    switch (id) {
    case EXTFEM_NL_NSE2D: {
        const double *u = in, *g = in + 2;
        double pr = in[6], mu = p[0];
        val[0] = g[0] * u[0] + g[1] * u[1];
        val[1] = g[2] * u[0] + g[3] * u[1];
        val[2] = mu * g[0] - pr;
        val[3] = mu * g[1];
        val[4] = mu * g[2];
        val[5] = mu * g[3] - pr;
        val[6] = -(g[0] + g[3]);
        J[0 * 7 + 0] = g[0]; J[0 * 7 + 1] = g[1]; J[0 * 7 + 2] = u[0]; J[0 * 7 + 3] = u[1];
        J[1 * 7 + 0] = g[2]; J[1 * 7 + 1] = g[3]; J[1 * 7 + 4] = u[0]; J[1 * 7 + 5] = u[1];
        J[2 * 7 + 2] = mu; J[2 * 7 + 6] = -1.0;
        J[3 * 7 + 3] = mu;
        J[4 * 7 + 4] = mu;
        J[5 * 7 + 5] = mu; J[5 * 7 + 6] = -1.0;
        J[6 * 7 + 2] = -1.0; J[6 * 7 + 5] = -1.0;
    } break;
    case EXTFEM_NL_LINNSE7: {
        const double *u = in, *g = in + 2;
        double pr = in[6], mu = p[0], al = p[1];
        val[0] = g[0] + al * u[0];
        val[1] = g[2] + al * u[1];
        val[2] = mu * g[0] - pr;
        val[3] = mu * g[1];
        val[4] = mu * g[2];
        val[5] = mu * g[3] - pr;
        val[6] = -(g[0] + g[3]);
        J[0 * 7 + 2] = 1.0; J[0 * 7 + 0] = al;
        J[1 * 7 + 4] = 1.0; J[1 * 7 + 1] = al;
        J[2 * 7 + 2] = mu; J[2 * 7 + 6] = -1.0;
        J[3 * 7 + 3] = mu;
        J[4 * 7 + 4] = mu;
        J[5 * 7 + 5] = mu; J[5 * 7 + 6] = -1.0;
        J[6 * 7 + 2] = -1.0; J[6 * 7 + 5] = -1.0;
    } break;
    case EXTFEM_NL_NEOHOOKE3D: {
        // DW = mu F + c * dd,  c = (lambda log(det) - mu)/det,  dd = d det / dF
        // D2W = mu I + c * d(dd)/dF + dd (x) dd * (lambda + mu - lambda log(det)) / det^2
        double mu = p[0], la = p[1];
        double F[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) F[i] = in[i];
        F[0] += 1.0; F[4] += 1.0; F[8] += 1.0;
        double dd[9];
        dd[0] = F[4] * F[8] - F[5] * F[7];
        dd[1] = F[5] * F[6] - F[3] * F[8];
        dd[2] = F[3] * F[7] - F[4] * F[6];
        dd[3] = F[2] * F[7] - F[1] * F[8];
        dd[4] = F[0] * F[8] - F[2] * F[6];
        dd[5] = F[1] * F[6] - F[0] * F[7];
        dd[6] = F[1] * F[5] - F[2] * F[4];
        dd[7] = F[2] * F[3] - F[0] * F[5];
        dd[8] = F[0] * F[4] - F[1] * F[3];
        double det = F[0] * dd[0] + F[1] * dd[1] + F[2] * dd[2];
        double ld = log(det);
        double c = (la * ld - mu) / det;
        double e = (la + mu - la * ld) / (det * det);
#pragma unroll
        for (int i = 0; i < 9; ++i) val[i] = mu * F[i] + c * dd[i];
#pragma unroll
        for (int i = 0; i < 9; ++i) {
#pragma unroll
            for (int j = 0; j < 9; ++j) J[i * 9 + j] = e * dd[i] * dd[j];
            J[i * 9 + i] += mu;
        }
        // c * d(dd_i)/dF_j : dd is the cofactor matrix, its derivative is +-F entries
#define EXTFEM_D2(i, j, s, k) J[(i) * 9 + (j)] += (s) * c * F[k];
        EXTFEM_D2(0, 4, 1, 8) EXTFEM_D2(0, 8, 1, 4) EXTFEM_D2(0, 5, -1, 7) EXTFEM_D2(0, 7, -1, 5)
        EXTFEM_D2(1, 5, 1, 6) EXTFEM_D2(1, 6, 1, 5) EXTFEM_D2(1, 3, -1, 8) EXTFEM_D2(1, 8, -1, 3)
        EXTFEM_D2(2, 3, 1, 7) EXTFEM_D2(2, 7, 1, 3) EXTFEM_D2(2, 4, -1, 6) EXTFEM_D2(2, 6, -1, 4)
        EXTFEM_D2(3, 2, 1, 7) EXTFEM_D2(3, 7, 1, 2) EXTFEM_D2(3, 1, -1, 8) EXTFEM_D2(3, 8, -1, 1)
        EXTFEM_D2(4, 0, 1, 8) EXTFEM_D2(4, 8, 1, 0) EXTFEM_D2(4, 2, -1, 6) EXTFEM_D2(4, 6, -1, 2)
        EXTFEM_D2(5, 1, 1, 6) EXTFEM_D2(5, 6, 1, 1) EXTFEM_D2(5, 0, -1, 7) EXTFEM_D2(5, 7, -1, 0)
        EXTFEM_D2(6, 1, 1, 5) EXTFEM_D2(6, 5, 1, 1) EXTFEM_D2(6, 2, -1, 4) EXTFEM_D2(6, 4, -1, 2)
        EXTFEM_D2(7, 2, 1, 3) EXTFEM_D2(7, 3, 1, 2) EXTFEM_D2(7, 0, -1, 5) EXTFEM_D2(7, 5, -1, 0)
        EXTFEM_D2(8, 0, 1, 4) EXTFEM_D2(8, 4, 1, 0) EXTFEM_D2(8, 1, -1, 3) EXTFEM_D2(8, 3, -1, 1)
#undef EXTFEM_D2
    } break;
    case EXTFEM_NL_RCD: {
        val[0] = in[0] * in[1] + in[0];
        J[0 * nin + 0] = in[1] + 1.0;
        J[0 * nin + 1] = in[0];
#pragma unroll
        for (int d = 0; d < dim; ++d) { val[1 + d] = in[1 + d]; J[(1 + d) * nin + 1 + d] = 1.0; }
    } break;
    case EXTFEM_NL_NLPOISSON105: {
        const double ep = exp(in[0]), em = exp(-in[0]);
        val[0] = ep - em;
        J[0] = ep + em;
#pragma unroll
        for (int d = 0; d < dim; ++d) { val[1 + d] = p[0] * in[1 + d]; J[(1 + d) * nin + 1 + d] = p[0]; }
    } break;
    case EXTFEM_NL_POROUS106: {
        // porous medium flux m u^(m-1) grad u (result has dim components, input u and grad u)
        const double m = p[0], um1 = pow(in[0], m - 1.0), um2 = (m == 2.0) ? 1.0 : pow(in[0], m - 2.0);
#pragma unroll
        for (int d = 0; d < dim; ++d) {
            val[d] = m * um1 * in[1 + d];
            J[d * nin + 0] = m * (m - 1.0) * um2 * in[1 + d];
            J[d * nin + 1 + d] = m * um1;
        }
    } break;
    case EXTFEM_NL_STVENANT230: {
        // Green-Lagrange strain (Voigt) minus thermal strain, isotropic Hooke; material by cell region
        const int R = (int)p[0], m = region >= 1 && region <= R ? region - 1 : 0;
        const double la = p[1 + m], mu = p[1 + R + m], eT = p[1 + 2 * R + m];
        const double g1 = in[0], g2 = in[1], g3 = in[2], g4 = in[3];
        const double e1 = g1 + 0.5 * (g1 * g1 + g3 * g3) - eT, e2 = g4 + 0.5 * (g2 * g2 + g4 * g4) - eT;
        const double e3 = g2 + g3 + g1 * g2 + g3 * g4;
        const double a = la * (e1 + e2) + 2.0 * mu * e1, b = la * (e1 + e2) + 2.0 * mu * e2, c = 2.0 * mu * e3;
        val[0] = a; val[1] = c; val[2] = c; val[3] = b;
        const double d1[4] = {1.0 + g1, 0.0, g3, 0.0}, d2[4] = {0.0, g2, 0.0, 1.0 + g4}, d3[4] = {g2, 1.0 + g1, 1.0 + g4, g3};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            J[0 * 4 + j] = (la + 2.0 * mu) * d1[j] + la * d2[j];
            J[3 * 4 + j] = la * d1[j] + (la + 2.0 * mu) * d2[j];
            J[1 * 4 + j] = J[2 * 4 + j] = 2.0 * mu * d3[j];
        }
    } break;
    }
}

// ---- shared-memory cell slot ------------------------------------------------------------------
template <int DIM>
struct CellGeo {
    double Ainv[DIM * DIM];
    double A[DIM * DIM];
    double x0[DIM];
    double vol;
    int visited;
    int pad;
};

template <int DIM>
__device__ __forceinline__ void load_geo(const OpDev &op, long long cell, CellGeo<DIM> &G)
{
    const int *cn = op.cellnodes + cell * (DIM + 1);
    const double *p0 = op.coords + (size_t)cn[0] * DIM;
    double A[DIM][DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) G.x0[d] = p0[d];
#pragma unroll
    for (int r = 0; r < DIM; ++r) {
        const double *pr = op.coords + (size_t)cn[r + 1] * DIM;
#pragma unroll
        for (int d = 0; d < DIM; ++d) A[d][r] = pr[d] - p0[d];
    }
#pragma unroll
    for (int d = 0; d < DIM; ++d)
#pragma unroll
        for (int r = 0; r < DIM; ++r) G.A[d * DIM + r] = A[d][r];
    if (DIM == 1) {
        G.Ainv[0] = 1.0 / A[0][0];
    } else if (DIM == 2) {
        double det = A[0][0] * A[1 % DIM][1 % DIM] - A[0][1 % DIM] * A[1 % DIM][0];
        double id = 1.0 / det;
        G.Ainv[0 * DIM + 0] = A[1 % DIM][1 % DIM] * id;
        G.Ainv[0 * DIM + 1 % DIM] = -A[0][1 % DIM] * id;
        G.Ainv[(1 % DIM) * DIM + 0] = -A[1 % DIM][0] * id;
        G.Ainv[(1 % DIM) * DIM + 1 % DIM] = A[0][0] * id;
    } else {
        constexpr int I1 = 1 % DIM, I2 = 2 % DIM;
        double c00 = A[I1][I1] * A[I2][I2] - A[I1][I2] * A[I2][I1];
        double c01 = A[I1][I2] * A[I2][0] - A[I1][0] * A[I2][I2];
        double c02 = A[I1][0] * A[I2][I1] - A[I1][I1] * A[I2][0];
        double det = A[0][0] * c00 + A[0][I1] * c01 + A[0][I2] * c02;
        double id = 1.0 / det;
        G.Ainv[0 * DIM + 0] = c00 * id;
        G.Ainv[I1 * DIM + 0] = c01 * id;
        G.Ainv[I2 * DIM + 0] = c02 * id;
        G.Ainv[0 * DIM + I1] = (A[0][I2] * A[I2][I1] - A[0][I1] * A[I2][I2]) * id;
        G.Ainv[I1 * DIM + I1] = (A[0][0] * A[I2][I2] - A[0][I2] * A[I2][0]) * id;
        G.Ainv[I2 * DIM + I1] = (A[0][I1] * A[I2][0] - A[0][0] * A[I2][I1]) * id;
        G.Ainv[0 * DIM + I2] = (A[0][I1] * A[I1][I2] - A[0][I2] * A[I1][I1]) * id;
        G.Ainv[I1 * DIM + I2] = (A[0][I2] * A[I1][0] - A[0][0] * A[I1][I2]) * id;
        G.Ainv[I2 * DIM + I2] = (A[0][0] * A[I1][I1] - A[0][I1] * A[I1][0]) * id;
    }
    G.vol = op.cellvolumes[cell];
    int reg = op.cellregions[cell];
    int vis = 1;
    if (op.nregions > 0) {
        vis = 0;
        for (int k = 0; k < op.nregions; ++k) vis |= (op.regions[k] == reg);
    }
    G.visited = vis;
}

template <int DIM>
__device__ __forceinline__ void eval_x(const CellGeo<DIM> &G, const double *xref, double *x)
{
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
        double s = G.x0[d];
#pragma unroll
        for (int r = 0; r < DIM; ++r) s += G.A[d * DIM + r] * xref[r];
        x[d] = s;
    }
}

// input_args[q] = sum_j sol[dof_j] * cvals[:, j, q]   (bilinear_operator.jl:513-521)
template <int DIM>
__device__ __forceinline__ void eval_args(const OpDev &op, long long cell, int q, const CellGeo<DIM> &G, double *u)
{
    for (int d = 0; d < op.nin; ++d) u[d] = 0.0;
    for (int id = 0; id < op.nargs; ++id) {
        const ArgDev &a = op.args[id];
        const int *dofs = a.celldofs + cell * a.nd;
        for (int j = 0; j < a.nd; ++j) {
            double s = op.sol[a.soloff + dofs[j]];
            CV cv = eval_cv<DIM>(a, j, q, G.Ainv, op.offdiag);
            for (int t = 0; t < cv.len; ++t) u[a.opoff + cv.idx[t]] += s * cv.v[t];
        }
    }
}

__device__ __forceinline__ const ArgDev *find_arg(const ArgDev *args, int n, int loc, int &jj, int from)
{
    // arguments whose block covers local index `loc`; iterate with `from`
    for (int i = from; i < n; ++i)
        if (loc >= args[i].locoff && loc < args[i].locoff + args[i].nd) { jj = loc - args[i].locoff; return &args[i]; }
    return nullptr;
}

// =============================== BilinearOperator ==============================================
template <int DIM>
__global__ void __launch_bounds__(256)
local_bilinear_kernel(const __grid_constant__ OpDev op, double *__restrict__ loc, int CPB)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CellGeo<DIM> *geo = reinterpret_cast<CellGeo<DIM> *>(smem_raw);
    double *uq = reinterpret_cast<double *>(geo + CPB); // [CPB][nq][nin_args] when nargs > 0
    const int nin_args = op.nargs > 0 ? op.nin : 0;
    long long c0 = (long long)blockIdx.x * CPB;
    int ncell = (int)min((long long)CPB, op.ncells - c0);
    for (int s = threadIdx.x; s < ncell; s += blockDim.x) load_geo<DIM>(op, c0 + s, geo[s]);
    __syncthreads();
    if (op.nargs > 0) {
        for (int t = threadIdx.x; t < ncell * op.nq; t += blockDim.x) {
            int s = t / op.nq, q = t - s * op.nq;
            double u[MAXOP];
            eval_args<DIM>(op, c0 + s, q, geo[s], u);
            for (int d = 0; d < nin_args; ++d) uq[(size_t)t * nin_args + d] = u[d];
        }
        __syncthreads();
    }
    const int NRC = op.NR * op.NC;
    // ansatz-vector length when kernel has args: the kernel input is the ansatz evaluation
    int nin_a = 0;
    for (int i = 0; i < op.nansatz; ++i) nin_a = max(nin_a, op.ansatz[i].opoff + op.ansatz[i].oplen);
    for (int e = threadIdx.x; e < ncell * NRC; e += blockDim.x) {
        int s = e / NRC, r = e - s * NRC;
        int j = r / op.NR, k = r - j * op.NR;
        const CellGeo<DIM> &G = geo[s];
        double acc = 0.0;
        if (G.visited) {
            for (int id = 0; id < op.nansatz; ++id) {
                const ArgDev &aa = op.ansatz[id];
                if (j < aa.locoff || j >= aa.locoff + aa.nd) continue;
                int jj = j - aa.locoff;
                for (int idt = 0; idt < op.ntest; ++idt) {
                    const ArgDev &ta = op.test[idt];
                    if (k < ta.locoff || k >= ta.locoff + ta.nd) continue;
                    int kk = k - ta.locoff;
                    if (op.lump) { if (idt != id || kk != jj) continue; }
                    else if (!op.coupling[id * op.ntest + idt]) continue;
                    double a = 0.0;
                    for (int q = 0; q < op.nq; ++q) {
                        CV ca = eval_cv<DIM>(aa, jj, q, G.Ainv, op.offdiag);
                        double fw = op.factor * op.qw[q];
                        if (op.kernel_id == EXTFEM_BLK_STANDARD) {
                            // result = input: only matching slots of the operator vectors contribute
                            if (op.lump == 2) {
                                for (int k2 = 0; k2 < ta.nd; ++k2) {
                                    CV ct = eval_cv<DIM>(ta, k2, q, G.Ainv, op.offdiag);
                                    for (int t = 0; t < ct.len; ++t)
                                        for (int t2 = 0; t2 < ca.len; ++t2)
                                            if (ct.idx[t] + ta.opoff == ca.idx[t2] + aa.opoff) a += (ca.v[t2] * fw) * ct.v[t];
                                }
                            } else {
                                CV ct = eval_cv<DIM>(ta, kk, q, G.Ainv, op.offdiag);
                                for (int t = 0; t < ct.len; ++t)
                                    for (int t2 = 0; t2 < ca.len; ++t2)
                                        if (ct.idx[t] + ta.opoff == ca.idx[t2] + aa.opoff) a += (ca.v[t2] * fw) * ct.v[t];
                            }
                        } else {
                            double in[MAXOP], res[MAXOP];
                            for (int d = 0; d < MAXOP; ++d) in[d] = 0.0;
                            for (int t2 = 0; t2 < ca.len; ++t2) in[aa.opoff + ca.idx[t2]] = ca.v[t2];
                            const double *ain = op.nargs > 0 ? uq + ((size_t)s * op.nq + q) * nin_args : nullptr;
                            bl_apply(op.kernel_id, DIM, in, ain, op.params, res, op.nout);
                            if (op.lump == 2) {
                                for (int k2 = 0; k2 < ta.nd; ++k2) {
                                    CV ct = eval_cv<DIM>(ta, k2, q, G.Ainv, op.offdiag);
                                    for (int t = 0; t < ct.len; ++t) a += (res[ta.opoff + ct.idx[t]] * fw) * ct.v[t];
                                }
                            } else {
                                CV ct = eval_cv<DIM>(ta, kk, q, G.Ainv, op.offdiag);
                                for (int t = 0; t < ct.len; ++t) a += (res[ta.opoff + ct.idx[t]] * fw) * ct.v[t];
                            }
                        }
                    }
                    acc += a * G.vol;
                }
            }
        }
        loc[(size_t)(c0 + s) * NRC + r] = acc;
    }
    (void)nin_a;
}

// =============================== LinearOperator ================================================
template <int DIM>
__global__ void __launch_bounds__(256)
local_linear_kernel(const __grid_constant__ OpDev op, double *__restrict__ bloc, int CPB)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CellGeo<DIM> *geo = reinterpret_cast<CellGeo<DIM> *>(smem_raw);
    double *fq = reinterpret_cast<double *>(geo + CPB); // [CPB][nq][nout]
    long long c0 = (long long)blockIdx.x * CPB;
    int ncell = (int)min((long long)CPB, op.ncells - c0);
    for (int s = threadIdx.x; s < ncell; s += blockDim.x) load_geo<DIM>(op, c0 + s, geo[s]);
    __syncthreads();
    for (int t = threadIdx.x; t < ncell * op.nq; t += blockDim.x) {
        int s = t / op.nq, q = t - s * op.nq;
        const CellGeo<DIM> &G = geo[s];
        double r[MAXOP];
        if (op.nargs > 0) {
            eval_args<DIM>(op, c0 + s, q, G, r); // standard kernel: result = input_args
        } else {
            double x[DIM];
            eval_x<DIM>(G, op.qx + q * DIM, x);
            const double *tab = op.tabulated ? op.tabulated + ((size_t)(c0 + s) * op.nq + q) * op.nout : nullptr;
            lin_apply(op.kernel_id, x, op.params, r, op.nout, tab);
        }
        double sc = op.factor * op.qw[q] * G.vol;
        for (int d = 0; d < op.nout; ++d) fq[(size_t)t * op.nout + d] = r[d] * sc;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < ncell * op.NR; e += blockDim.x) {
        int s = e / op.NR, k = e - s * op.NR;
        const CellGeo<DIM> &G = geo[s];
        double acc = 0.0;
        if (G.visited) {
            for (int idt = 0; idt < op.ntest; ++idt) {
                const ArgDev &ta = op.test[idt];
                if (k < ta.locoff || k >= ta.locoff + ta.nd) continue;
                int kk = k - ta.locoff;
                for (int q = 0; q < op.nq; ++q) {
                    CV ct = eval_cv<DIM>(ta, kk, q, G.Ainv, op.offdiag);
                    const double *f = fq + ((size_t)s * op.nq + q) * op.nout;
                    for (int t = 0; t < ct.len; ++t) acc += f[ta.opoff + ct.idx[t]] * ct.v[t];
                }
            }
        }
        bloc[(size_t)(c0 + s) * op.NR + k] = acc;
    }
}

// =============================== NonlinearOperator =============================================
template <int DIM>
__global__ void __launch_bounds__(256)
local_nonlinear_kernel(const __grid_constant__ OpDev op, double *__restrict__ loc, double *__restrict__ bloc, int CPB)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CellGeo<DIM> *geo = reinterpret_cast<CellGeo<DIM> *>(smem_raw);
    const int nin = op.nin, nout = op.nout, JS = nin * nout;
    double *jq = reinterpret_cast<double *>(geo + CPB); // [CPB][nq][nout*nin]
    double *rq = jq + (size_t)CPB * op.nq * JS;         // [CPB][nq][nout]   (jac*u - value)*factor*w*vol
    long long c0 = (long long)blockIdx.x * CPB;
    int ncell = (int)min((long long)CPB, op.ncells - c0);
    for (int s = threadIdx.x; s < ncell; s += blockDim.x) load_geo<DIM>(op, c0 + s, geo[s]);
    __syncthreads();
    for (int t = threadIdx.x; t < ncell * op.nq; t += blockDim.x) {
        int s = t / op.nq, q = t - s * op.nq;
        const CellGeo<DIM> &G = geo[s];
        double u[MAXOP], val[MAXOP];
        eval_args<DIM>(op, c0 + s, q, G, u);
        double *J = jq + (size_t)t * JS;
        nl_apply(op.kernel_id, DIM, u, op.params, val, J, nin, nout, op.cellregions[c0 + s]);
        double sc = op.factor * op.qw[q] * G.vol;
        for (int k = 0; k < nout; ++k) {
            double sum = 0.0;
            for (int d = 0; d < nin; ++d) sum += J[k * nin + d] * u[d];
            rq[(size_t)t * nout + k] = (sum - val[k]) * sc;
        }
    }
    __syncthreads();
    const int NRC = op.NR * op.NC;
    for (int e = threadIdx.x; e < ncell * NRC; e += blockDim.x) {
        int s = e / NRC, r = e - s * NRC;
        int j = r / op.NR, k = r - j * op.NR;
        const CellGeo<DIM> &G = geo[s];
        double acc = 0.0;
        if (G.visited) {
            for (int id = 0; id < op.nargs; ++id) {
                const ArgDev &ga = op.args[id];
                if (j < ga.locoff || j >= ga.locoff + ga.nd) continue;
                int jj = j - ga.locoff;
                for (int idt = 0; idt < op.ntest; ++idt) {
                    const ArgDev &ta = op.test[idt];
                    if (k < ta.locoff || k >= ta.locoff + ta.nd) continue;
                    int kk = k - ta.locoff;
                    double a = 0.0;
                    for (int q = 0; q < op.nq; ++q) {
                        CV cg = eval_cv<DIM>(ga, jj, q, G.Ainv, op.offdiag);
                        CV ct = eval_cv<DIM>(ta, kk, q, G.Ainv, op.offdiag);
                        const double *J = jq + ((size_t)s * op.nq + q) * JS;
                        double w = op.qw[q];
                        for (int t = 0; t < ct.len; ++t) {
                            const double *Jrow = J + (ta.opoff + ct.idx[t]) * nin + ga.opoff;
                            double tv = 0.0;
                            for (int t2 = 0; t2 < cg.len; ++t2) tv += Jrow[cg.idx[t2]] * cg.v[t2];
                            a += tv * ct.v[t] * w;
                        }
                    }
                    acc += a * (op.factor * G.vol);
                }
            }
        }
        loc[(size_t)(c0 + s) * NRC + r] = acc;
    }
    for (int e = threadIdx.x; e < ncell * op.NR; e += blockDim.x) {
        int s = e / op.NR, k = e - s * op.NR;
        const CellGeo<DIM> &G = geo[s];
        double acc = 0.0;
        if (G.visited) {
            for (int idt = 0; idt < op.ntest; ++idt) {
                const ArgDev &ta = op.test[idt];
                if (k < ta.locoff || k >= ta.locoff + ta.nd) continue;
                int kk = k - ta.locoff;
                for (int q = 0; q < op.nq; ++q) {
                    CV ct = eval_cv<DIM>(ta, kk, q, G.Ainv, op.offdiag);
                    const double *f = rq + ((size_t)s * op.nq + q) * nout;
                    for (int t = 0; t < ct.len; ++t) acc += f[ta.opoff + ct.idx[t]] * ct.v[t];
                }
            }
        }
        bloc[(size_t)(c0 + s) * op.NR + k] = acc;
    }
}

// ---- NonlinearOperator, restructured (v2) -------------------------------------------------------------
// Same sums as local_nonlinear_kernel, organised as small dense contractions staged in shared memory:
//   cvG/cvT   combined operator vectors of every local column / row dof at every quadrature point (all arguments that
//             live on the dof's block: e.g. id(u) and grad(u)), evaluated ONCE per cell instead of once per local entry
//   u, J, F   per quadrature point: input_args from cvG, kernel value and analytic Jacobian
//   GJ        w_q * J_q * cvG   (nout values per column dof and point)
//   A_loc     sum_q cvT . GJ    (<= 6 multiply-adds per point and entry), scaled by factor*|T| once per cell
//             (nonlinear_operator.jl:396,419); rhs sum_q cvT . (J u - F) pre-scaled per point (:406)
constexpr int CVC_MAX = 6;
constexpr int NL2_MAXSP_ = 4;
struct CVC {
    double v[CVC_MAX];
    unsigned char idx[CVC_MAX];
    unsigned char len, pad;
};

template <int DIM>
__device__ __forceinline__ void combined_cv(const ArgDev *args, int nargs, int loc, int q, const double *Ainv, double offdiag, CVC &out)
{
    int n = 0;
    for (int i = 0; i < nargs; ++i) {
        const ArgDev &a = args[i];
        if (loc < a.locoff || loc >= a.locoff + a.nd) continue;
        CV cv = eval_cv<DIM>(a, loc - a.locoff, q, Ainv, offdiag);
        for (int t = 0; t < cv.len && n < CVC_MAX; ++t, ++n) { out.v[n] = cv.v[t]; out.idx[n] = (unsigned char)(a.opoff + cv.idx[t]); }
    }
    out.len = (unsigned char)n;
}

__host__ __device__ inline size_t nl2_cell_bytes(int sizeof_geo, int nq, int nin, int nout, int NR, int NC)
{
    size_t b = (size_t)sizeof_geo + (size_t)nq * nin * nout * 8 + (size_t)nq * nout * 8 + (size_t)nq * (NR + NC) * sizeof(CVC) +
               (size_t)nq * NC * nout * 8;
    return (b + 15) & ~(size_t)15;
}

template <int DIM>
__global__ void __launch_bounds__(256)
local_nonlinear_kernel2(const __grid_constant__ OpDev op, double *__restrict__ loc, double *__restrict__ bloc, int CPB)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nin = op.nin, nout = op.nout, JS = nin * nout, nq = op.nq, NR = op.NR, NC = op.NC;
    const size_t cb = nl2_cell_bytes((int)sizeof(CellGeo<DIM>), nq, nin, nout, NR, NC);
    auto geo_of = [&](int s) { return reinterpret_cast<CellGeo<DIM> *>(smem_raw + s * cb); };
    auto jq_of = [&](int s) { return reinterpret_cast<double *>(smem_raw + s * cb + sizeof(CellGeo<DIM>)); };   // [nq][nout][nin]
    auto rq_of = [&](int s) { return jq_of(s) + (size_t)nq * JS; };                                              // [nq][nout]
    auto gj_of = [&](int s) { return rq_of(s) + (size_t)nq * nout; };                                            // [nq][NC][nout]
    auto cvg_of = [&](int s) { return reinterpret_cast<CVC *>(gj_of(s) + (size_t)nq * NC * nout); };             // [nq][NC]
    auto cvt_of = [&](int s) { return cvg_of(s) + (size_t)nq * NC; };                                            // [nq][NR]
    const long long c0 = (long long)blockIdx.x * CPB;
    const int ncell = (int)min((long long)CPB, op.ncells - c0);
    for (int s = threadIdx.x; s < ncell; s += blockDim.x) load_geo<DIM>(op, c0 + s, *geo_of(s));
    __syncthreads();
    // combined operator vectors
    for (int e = threadIdx.x; e < ncell * nq * (NC + NR); e += blockDim.x) {
        const int s = e / (nq * (NC + NR)), r = e - s * nq * (NC + NR);
        const int q = r / (NC + NR), i = r - q * (NC + NR);
        const CellGeo<DIM> &G = *geo_of(s);
        if (i < NC) combined_cv<DIM>(op.args, op.nargs, i, q, G.Ainv, op.offdiag, cvg_of(s)[q * NC + i]);
        else combined_cv<DIM>(op.test, op.ntest, i - NC, q, G.Ainv, op.offdiag, cvt_of(s)[q * NR + (i - NC)]);
    }
    __syncthreads();
    // input_args, kernel value and Jacobian per quadrature point
    for (int t = threadIdx.x; t < ncell * nq; t += blockDim.x) {
        const int s = t / nq, q = t - s * nq;
        const CellGeo<DIM> &G = *geo_of(s);
        double u[MAXOP], val[MAXOP];
        for (int d = 0; d < nin; ++d) u[d] = 0.0;
        const CVC *cg = cvg_of(s) + q * NC;
        for (int id = 0; id < op.nargs; ++id) {       // every column block once (arguments on one block share their dofs)
            const ArgDev &a = op.args[id];
            bool seen = false;
            for (int i2 = 0; i2 < id; ++i2) seen |= (op.args[i2].locoff == a.locoff);
            if (seen) continue;
            const int *dofs = a.celldofs + (c0 + s) * a.nd;
            for (int j = 0; j < a.nd; ++j) {
                const double sv = op.sol[a.soloff + dofs[j]];
                const CVC &c = cg[a.locoff + j];
                for (int x = 0; x < c.len; ++x) u[c.idx[x]] += sv * c.v[x];
            }
        }
        double *J = jq_of(s) + (size_t)q * JS;
        nl_apply(op.kernel_id, DIM, u, op.params, val, J, nin, nout, op.cellregions[c0 + s]);
        const double sc = op.factor * op.qw[q] * G.vol;
        for (int k = 0; k < nout; ++k) {
            double sum = 0.0;
            for (int d = 0; d < nin; ++d) sum += J[k * nin + d] * u[d];
            rq_of(s)[q * nout + k] = (sum - val[k]) * sc;
        }
    }
    __syncthreads();
    // GJ = w_q J_q cvG
    for (int e = threadIdx.x; e < ncell * nq * NC * nout; e += blockDim.x) {
        const int s = e / (nq * NC * nout), r = e - s * nq * NC * nout;
        const int q = r / (NC * nout), r2 = r - q * NC * nout;
        const int j = r2 / nout, t = r2 - j * nout;
        const CVC &c = cvg_of(s)[q * NC + j];
        const double *Jrow = jq_of(s) + (size_t)q * JS + t * nin;
        double a = 0.0;
        for (int x = 0; x < c.len; ++x) a += Jrow[c.idx[x]] * c.v[x];
        gj_of(s)[r] = a * op.qw[q];
    }
    __syncthreads();
    const int NRC = NR * NC;
    for (int e = threadIdx.x; e < ncell * NRC; e += blockDim.x) {
        const int s = e / NRC, r = e - s * NRC;
        const int j = r / NR, k = r - j * NR;
        const CellGeo<DIM> &G = *geo_of(s);
        double acc = 0.0;
        if (G.visited) {
            const double *gj = gj_of(s) + (size_t)j * nout;
            const CVC *ct = cvt_of(s) + k;
            for (int q = 0; q < nq; ++q) {
                const CVC &c = ct[q * NR];
                const double *g = gj + (size_t)q * NC * nout;
                for (int x = 0; x < c.len; ++x) acc += c.v[x] * g[c.idx[x]];
            }
            acc *= op.factor * G.vol;
        }
        loc[(size_t)(c0 + s) * NRC + r] = acc;
    }
    for (int e = threadIdx.x; e < ncell * NR; e += blockDim.x) {
        const int s = e / NR, k = e - s * NR;
        const CellGeo<DIM> &G = *geo_of(s);
        double acc = 0.0;
        if (G.visited) {
            for (int q = 0; q < nq; ++q) {
                const CVC &c = cvt_of(s)[q * NR + k];
                const double *f = rq_of(s) + q * nout;
                for (int x = 0; x < c.len; ++x) acc += f[c.idx[x]] * c.v[x];
            }
        }
        bloc[(size_t)(c0 + s) * NR + k] = acc;
    }
}

// ---- NonlinearOperator, warp per cell (v3) -------------------------------------------------------------
// One warp owns one cell at a time; phases are separated by __syncwarp only (no block barriers), every phase runs over a
// flattened item index decoded through small cell-independent lookup tables built on the host (no integer division):
//   PHI  value / physical gradient of every scalar basis function at every point        items (point, function)
//   B    operator vectors of the local dofs  B[q][x][j] = scale * PHI[...]               items (q, x, j)  -> gather table
//   u,J  input_args, kernel value and Jacobian (w_q-scaled) per point                    lanes = points
//   GJ   (w_q J_q) B_q[:, j]                                                             items (q, j, t)  -> packed table
//   A    sum_q B_q^T GJ, scaled by factor*|T| once per cell; rhs from (J u - F)          items (j, k)     -> packed table
// Rows and columns share one B table when test and args describe the same operators (the usual Newton setting).
struct NL3Tables {
    const unsigned char *tab;      // device table buffer (offsets below are bytes into it; 16-byte aligned pieces)
    int tab_bytes;
    int NC, NR, EC, ER, same;      // same != 0: rows use the column tables
    int o_cn, o_cout, o_rn, o_rout;            // u8 [N], u8 [E][N]
    int o_bgidx, o_bgsc, o_btidx, o_btsc;      // i32 / f64 [nq][E][N]: PHI index (-1: zero) and scale of B
    int o_gjj, o_gjb;                          // u16 [nq*NC*nout][EC]: offsets into J (q*JS + t*nin + out_x) and into B
    int o_enb, o_eng;                          // u16 [NR*NC][ER]: offsets of BT[0][x][k] and of GJ[0][j][out_x(k)]
    int o_uptr, o_ulist;                       // input_args: per result index o the (j | x << 8) pairs with out == o
    int o_ucol;                                // i32 [NC]: offset of the column dof in the cell's solution gather (block, dof)
    // dense form of the operator matrix for the tensor-core kernel (local_nonlinear_kernel4): B[(q, component)][dof], row
    // stride Np4 (== 4 mod 16: conflict-free fragment loads), Kp4 = nq * nin rounded up to a multiple of 4
    int dense_ok, NT4, Np4;                    // tensor-core path: number of 8-wide dof tiles and the row stride of B_q
    int o_ddst;                                // u16 [EC][NC]: offset of the sparse entry inside the dense B_q
    int nnzJ, o_jslot;                         // structurally non-zero Jacobian entries; u8 [nout*nin] compact slot or 255
    int o_pt;                                  // u8x4 [EC][NC]: (space, scalar basis function, source 0 value | 1+d gradient, -) of a B entry
    int nspaces, ns[NL2_MAXSP_];
    const double *refvals[NL2_MAXSP_], *refgrads[NL2_MAXSP_];
    int phi_off[NL2_MAXSP_ + 1];
    int ncolblocks, blk_locoff[MAXARGS], blk_nd[MAXARGS];
    const int *blk_celldofs[MAXARGS];
    long long blk_soloff[MAXARGS];
};

__host__ __device__ inline size_t nl3_warp_doubles(int nq, int nin, int nout, int NC, int NR, int EC, int ER, int same, int phid)
{
    size_t d = (size_t)phid + (size_t)nq * EC * NC + (same ? 0 : (size_t)nq * ER * NR) + (size_t)nq * nin * nout + (size_t)nq * nout +
               (size_t)nq * nin + (size_t)nq * NC * nout + (size_t)NC;
    return (d + 1) & ~(size_t)1;
}

template <int DIM>
__global__ void __launch_bounds__(256)
local_nonlinear_kernel3(const __grid_constant__ OpDev op, const __grid_constant__ NL3Tables T, double *__restrict__ loc,
                        double *__restrict__ bloc, int cells_per_warp)
{
    extern __shared__ __align__(16) double smem_d[];
    const int nin = op.nin, nout = op.nout, JS = nin * nout, nq = op.nq, NR = T.NR, NC = T.NC, NRC = NR * NC, EC = T.EC, ER = T.ER;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int phid = T.phi_off[T.nspaces];
    const size_t wd = nl3_warp_doubles(nq, nin, nout, NC, NR, EC, ER, T.same, phid);
    // block-shared tables behind the per-warp areas
    unsigned char *tb = reinterpret_cast<unsigned char *>(smem_d + (size_t)nwarp * wd);
    for (int i = threadIdx.x; i < T.tab_bytes / 16; i += blockDim.x) reinterpret_cast<uint4 *>(tb)[i] = __ldg(reinterpret_cast<const uint4 *>(T.tab) + i);
    __syncthreads();
    const unsigned char *r_n = tb + T.o_rn, *r_out = tb + T.o_rout;
    const int *bgidx = reinterpret_cast<const int *>(tb + T.o_bgidx), *btidx = reinterpret_cast<const int *>(tb + T.o_btidx);
    const double *bgsc = reinterpret_cast<const double *>(tb + T.o_bgsc), *btsc = reinterpret_cast<const double *>(tb + T.o_btsc);
    const unsigned short *gjj = reinterpret_cast<const unsigned short *>(tb + T.o_gjj), *gjb = reinterpret_cast<const unsigned short *>(tb + T.o_gjb);
    const unsigned short *enb = reinterpret_cast<const unsigned short *>(tb + T.o_enb), *eng = reinterpret_cast<const unsigned short *>(tb + T.o_eng);
    const unsigned short *uptr = reinterpret_cast<const unsigned short *>(tb + T.o_uptr), *ulist = reinterpret_cast<const unsigned short *>(tb + T.o_ulist);
    double *PHI = smem_d + (size_t)warp * wd;
    double *BG = PHI + phid;                                       // [nq][EC][NC]
    double *BT = T.same ? BG : BG + (size_t)nq * EC * NC;          // [nq][ER][NR]
    double *Jq = (T.same ? BG : BT) + (T.same ? (size_t)nq * EC * NC : (size_t)nq * ER * NR);   // [nq][nout][nin]
    double *rq = Jq + (size_t)nq * JS;                             // [nq][nout]
    double *uq = rq + (size_t)nq * nout;                           // [nq][nin]
    double *GJ = uq + (size_t)nq * nin;                            // [nq][NC][nout]
    double *solc = GJ + (size_t)nq * NC * nout;                    // [NC] solution values of the cell's column dofs
    const long long cbase = ((long long)blockIdx.x * nwarp + warp) * cells_per_warp;
    for (int ci = 0; ci < cells_per_warp; ++ci) {
        const long long cell = cbase + ci;
        if (cell >= op.ncells) break;
        // geometry: every lane computes it redundantly (registers, no staging)
        CellGeo<DIM> G;
        load_geo<DIM>(op, cell, G);
        __syncwarp();
        // PHI
        for (int sp = 0; sp < T.nspaces; ++sp) {
            double *ph = PHI + T.phi_off[sp];
            for (int e = lane; e < nq * T.ns[sp]; e += 32) {
                double *o = ph + (size_t)e * (1 + DIM);
                o[0] = __ldg(T.refvals[sp] + e);
                const double *rg = T.refgrads[sp] + (size_t)e * DIM;
#pragma unroll
                for (int d = 0; d < DIM; ++d) {
                    double g = 0.0;
#pragma unroll
                    for (int r = 0; r < DIM; ++r) g += G.Ainv[r * DIM + d] * __ldg(rg + r);
                    o[1 + d] = g;
                }
            }
        }
        __syncwarp();
        // B tables
        for (int i = lane; i < nq * EC * NC; i += 32) { const int p = bgidx[i]; BG[i] = p >= 0 ? bgsc[i] * PHI[p] : 0.0; }
        if (!T.same)
            for (int i = lane; i < nq * ER * NR; i += 32) { const int p = btidx[i]; BT[i] = p >= 0 ? btsc[i] * PHI[p] : 0.0; }
        __syncwarp();
        // solution values of the cell's column dofs
        for (int b = 0; b < T.ncolblocks; ++b) {
            const int *dofs = T.blk_celldofs[b] + cell * T.blk_nd[b];
            for (int jl = lane; jl < T.blk_nd[b]; jl += 32) solc[T.blk_locoff[b] + jl] = op.sol[T.blk_soloff[b] + dofs[jl]];
        }
        __syncwarp();
        // input_args: items (point, result index) sum over the (dof, entry) pairs that feed the index
        for (int i = lane; i < nq * nin; i += 32) {
            const int q = i / nin, o = i - q * nin;
            const double *bg = BG + (size_t)q * EC * NC;
            double a = 0.0;
            for (int p = uptr[o]; p < uptr[o + 1]; ++p) {
                const int j = ulist[p] & 0xff, x = ulist[p] >> 8;
                a += solc[j] * bg[x * NC + j];
            }
            uq[i] = a;
        }
        __syncwarp();
        // kernel value and Jacobian (lanes = points)
        for (int q = lane; q < nq; q += 32) {
            double ur[MAXOP], val[MAXOP];
            for (int d = 0; d < nin; ++d) ur[d] = uq[q * nin + d];
            double *J = Jq + (size_t)q * JS;
            nl_apply(op.kernel_id, DIM, ur, op.params, val, J, nin, nout, op.cellregions[cell]);
            const double w = op.qw[q], sc = op.factor * w * G.vol;
            for (int k = 0; k < nout; ++k) {
                double sum = 0.0;
                for (int d = 0; d < nin; ++d) { sum += J[k * nin + d] * ur[d]; J[k * nin + d] *= w; }
                rq[q * nout + k] = (sum - val[k]) * sc;
            }
        }
        __syncwarp();
        // GJ[q][j][t]: EC multiply-adds per item through direct offset tables (padding entries hit zeros of B)
        for (int i = lane; i < nq * NC * nout; i += 32) {
            const unsigned short *oj = gjj + (size_t)i * EC, *ob = gjb + (size_t)i * EC;
            double a = 0.0;
            for (int x = 0; x < EC; ++x) a += Jq[oj[x]] * BG[ob[x]];
            GJ[i] = a;
        }
        __syncwarp();
        // local matrix (lanes = consecutive entries = consecutive rows of one column) and vector
        const double fv = G.visited ? op.factor * G.vol : 0.0;
        double *out = loc + (size_t)cell * NRC;
        const int sb = ER * NR, sg = NC * nout;
        for (int e = lane; e < NRC; e += 32) {
            const unsigned short *ob = enb + (size_t)e * ER, *og = eng + (size_t)e * ER;
            double acc = 0.0;
            for (int x = 0; x < ER; ++x) {
                const double *b = BT + ob[x];
                const double *g = GJ + og[x];
                for (int q = 0; q < nq; ++q) acc += b[(size_t)q * sb] * g[(size_t)q * sg];
            }
            out[e] = acc * fv;
        }
        double *bout = bloc + (size_t)cell * NR;
        for (int k = lane; k < NR; k += 32) {
            const int n = r_n[k];
            double acc = 0.0;
            for (int x = 0; x < n; ++x) {
                const double *b = BT + (size_t)x * NR + k;
                const double *f = rq + r_out[x * NR + k];
                for (int q = 0; q < nq; ++q) acc += f[q * nout] * b[(size_t)q * ER * NR];
            }
            bout[k] = G.visited ? acc : 0.0;
        }
        __syncwarp();
    }
}

// ---- NonlinearOperator on FP64 tensor cores (v4) -------------------------------------------------------------------------
// Two kernels.
//  (1) nl_point_kernel<DIM, NIO>: one THREAD per (cell, quadrature point): input_args from the solution, kernel value and analytic
//      Jacobian in registers (nonlinear_operator.jl:343-369), stores w_q J_q and (J u - F) factor w_q |T| as structure-of-arrays
//      [entry][cell * nq + q] (coalesced writes; the reader below fetches the nq values of one cell as whole sectors).  Every lane
//      works -- in the warp-per-cell kernels this phase ran on nq of 32 lanes through shared-memory read-modify-write chains and
//      cost more than all contractions together.
//  (2) local_nonlinear_kernel4<DIM>: one warp per cell, the two contractions of nonlinear_operator.jl:372-401 as dense GEMMs on
//      mma.sync.m8n8k4.f64 (DMMA):
//          GJ_q = (w_q J_q) B_q          [nout x nin] x [nin x NC]   per quadrature point
//          A    = sum_q B_q^T GJ_q       [NR x (nq nout)] x [(nq nout) x NC]
//      where B[(q, component)][dof] is the operator matrix of the cell (values / physical gradients of the basis functions routed to
//      the components of the kernel's input vector): the sparse B tables written into a zero-initialised dense array whose zero
//      pattern is cell-independent.  Used when test and args describe the same operators (T.same: the usual Newton setting).
__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Structural non-zeros of a registered kernel's Jacobian for the operator's parameters: 32 probes at random states (all regions
// the material tables distinguish); an entry that is zero in every probe is never stored or multiplied.
__global__ void nl_mask_kernel(const __grid_constant__ OpDev op, int dim, unsigned char *__restrict__ mask)
{
    double in[16], val[16], J[256];
    unsigned s = threadIdx.x * 2654435761u + 12345u;
    for (int i = 0; i < op.nin; ++i) {
        s = s * 1664525u + 1013904223u;
        in[i] = ((s >> 8) & 0xffffu) / 65536.0 * 1.5 + 0.25;
    }
    nl_apply(op.kernel_id, dim, in, op.params, val, J, op.nin, op.nout, 1 + (threadIdx.x & 7));
    for (int e = 0; e < op.nin * op.nout; ++e)
        if (J[e] != 0.0) mask[e] = 1;
}

// nl_apply with compile-time vector lengths and the kernel id folded per case: after inlining every index is static, the value and
// Jacobian arrays live in registers
template <int DIM, int NIO>
__device__ __forceinline__ void nl_apply_fixed(int id, const double (&u)[NIO], const double *p, double (&val)[NIO], double (&J)[NIO * NIO], int region)
{
    switch (id) {
    case EXTFEM_NL_NSE2D: if constexpr (NIO == 7) nl_apply(EXTFEM_NL_NSE2D, DIM, u, p, val, J, NIO, NIO, region); break;
    case EXTFEM_NL_LINNSE7: if constexpr (NIO == 7) nl_apply(EXTFEM_NL_LINNSE7, DIM, u, p, val, J, NIO, NIO, region); break;
    case EXTFEM_NL_NEOHOOKE3D: if constexpr (NIO == 9) nl_apply(EXTFEM_NL_NEOHOOKE3D, DIM, u, p, val, J, NIO, NIO, region); break;
    case EXTFEM_NL_RCD: if constexpr (NIO == 1 + DIM) nl_apply(EXTFEM_NL_RCD, DIM, u, p, val, J, NIO, NIO, region); break;
    case EXTFEM_NL_NLPOISSON105: if constexpr (NIO == 1 + DIM) nl_apply(EXTFEM_NL_NLPOISSON105, DIM, u, p, val, J, NIO, NIO, region); break;
    case EXTFEM_NL_STVENANT230: if constexpr (NIO == 4) nl_apply(EXTFEM_NL_STVENANT230, DIM, u, p, val, J, NIO, NIO, region); break;
    }
}

// Neo-Hooke (EXTFEM_NL_NEOHOOKE3D in nl_apply) with the Jacobian produced ROW BY ROW.  In the array form all 81 entries of D2W are
// computed before the first one is stored (the store loop is a chain of conditional blocks the compiler does not sink them
// into): nl_point_kernel<3, 9> needed 168 registers plus 392 bytes of spills, ~480 local-memory loads per point
// (profiles/r02_ncu_nl_point_config4.txt).  Same formulas and operation order as the array form.
struct NeoState {
    double F[9], dd[9], mu, c, e;
};

__device__ __forceinline__ void neo_prepare(const double (&in)[9], const double *p, NeoState &S, double (&val)[9])
{
    const double mu = p[0], la = p[1];
#pragma unroll
    for (int i = 0; i < 9; ++i) S.F[i] = in[i];
    S.F[0] += 1.0; S.F[4] += 1.0; S.F[8] += 1.0;
    const double (&F)[9] = S.F;
    S.dd[0] = F[4] * F[8] - F[5] * F[7];
    S.dd[1] = F[5] * F[6] - F[3] * F[8];
    S.dd[2] = F[3] * F[7] - F[4] * F[6];
    S.dd[3] = F[2] * F[7] - F[1] * F[8];
    S.dd[4] = F[0] * F[8] - F[2] * F[6];
    S.dd[5] = F[1] * F[6] - F[0] * F[7];
    S.dd[6] = F[1] * F[5] - F[2] * F[4];
    S.dd[7] = F[2] * F[3] - F[0] * F[5];
    S.dd[8] = F[0] * F[4] - F[1] * F[3];
    const double det = F[0] * S.dd[0] + F[1] * S.dd[1] + F[2] * S.dd[2];
    const double ld = log(det);
    S.mu = mu;
    S.c = (la * ld - mu) / det;
    S.e = (la + mu - la * ld) / (det * det);
#pragma unroll
    for (int i = 0; i < 9; ++i) val[i] = mu * F[i] + S.c * S.dd[i];
}

// d(dd_i)/dF_j of the cofactor matrix dd: sign * (k + 1) for +-F[k], 0 for none (the EXTFEM_D2 list of nl_apply as a table)
__host__ __device__ constexpr int neo_d2(int i, int j)
{
    constexpr int T[9][8] = {{4, 9, 8, 5, 5, -8, 7, -6}, {5, 7, 6, 6, 3, -9, 8, -4}, {3, 8, 7, 4, 4, -7, 6, -5},
                             {2, 8, 7, 3, 1, -9, 8, -2}, {0, 9, 8, 1, 2, -7, 6, -3}, {1, 7, 6, 2, 0, -8, 7, -1},
                             {1, 6, 5, 2, 2, -5, 4, -3}, {2, 4, 3, 3, 0, -6, 5, -1}, {0, 5, 4, 1, 1, -4, 3, -2}};
    for (int q = 0; q < 4; ++q)
        if (T[i][2 * q] == j) return T[i][2 * q + 1];
    return 0;
}

template <int I, int J_>
__device__ __forceinline__ double neo_entry(const NeoState &S)
{
    constexpr int d = neo_d2(I, J_);
    double v = S.e * S.dd[I] * S.dd[J_];
    if (I == J_) v += S.mu;
    if (d > 0) v += S.c * S.F[d > 0 ? d - 1 : 0];
    if (d < 0) v += -S.c * S.F[d < 0 ? -d - 1 : 0];
    return v;
}

template <int I>
__device__ __forceinline__ void neo_row(const NeoState &S, double (&row)[9])
{
    row[0] = neo_entry<I, 0>(S); row[1] = neo_entry<I, 1>(S); row[2] = neo_entry<I, 2>(S);
    row[3] = neo_entry<I, 3>(S); row[4] = neo_entry<I, 4>(S); row[5] = neo_entry<I, 5>(S);
    row[6] = neo_entry<I, 6>(S); row[7] = neo_entry<I, 7>(S); row[8] = neo_entry<I, 8>(S);
}

// shared memory of nl_point_kernel: tables | solution coefficients of the block's cells | (PVC) u16 index of every B entry into
// the thread's physical basis values | (PVC) the values [entry][thread]
__host__ __device__ inline size_t nl_point_smem(int tab_bytes, int nq, int NC, int EC, int npv, bool pvc)
{
    size_t b = (size_t)tab_bytes + (size_t)(128 / nq + 2) * NC * 8;
    if (pvc) b += ((size_t)EC * NC * 2 + 15) / 16 * 16 + (size_t)npv * 128 * 8;
    return b;
}

// ROW: the row-wise flavour of a registered kernel (Neo-Hooke, NIO = 9).
// PVC: every thread first evaluates the physical basis values of its (cell, point) -- value and DIM physical derivatives of every
// scalar basis function of every space, the PHI layout of the contraction kernel -- into its column of a shared array; the
// input_args loop then costs one shared load per (dof, entry) pair.  Without it every pair fetches its reference gradient, picks
// a column of A^-1 (which the compiler turns into a dynamically indexed local-memory array) and transforms it: three times the
// work for vector-valued spaces and ~65 instructions per pair (8000 per point at config 4).
template <int DIM, int NIO, bool ROW, bool PVC>
__global__ void __launch_bounds__(128, ROW ? 4 : 3)
nl_point_kernel(const __grid_constant__ OpDev op, const __grid_constant__ NL3Tables T, double *__restrict__ wJ, double *__restrict__ rqg)
{
    extern __shared__ __align__(16) unsigned char tb[];
    for (int i = threadIdx.x; i < T.tab_bytes / 16; i += blockDim.x) reinterpret_cast<uint4 *>(tb)[i] = __ldg(reinterpret_cast<const uint4 *>(T.tab) + i);
    const unsigned short *uptr = reinterpret_cast<const unsigned short *>(tb + T.o_uptr), *ulist = reinterpret_cast<const unsigned short *>(tb + T.o_ulist);
    const uchar4 *pt = reinterpret_cast<const uchar4 *>(tb + T.o_pt);
    const double *bgsc = reinterpret_cast<const double *>(tb + T.o_bgsc);
    const unsigned char *jslot = tb + T.o_jslot;
    double *solc = reinterpret_cast<double *>(tb + T.tab_bytes);      // [cells of the block][NC] solution coefficients
    const int nq = op.nq, NC = T.NC;
    const long long ntot = op.ncells * nq;
    const long long t0 = (long long)blockIdx.x * blockDim.x, t = t0 + threadIdx.x;
    const long long c0 = t0 / nq, c1 = min((t0 + blockDim.x - 1) / nq, op.ncells - 1);
    // the coefficients of the block's cells, one independent load per (cell, dof)
    for (int i = threadIdx.x; i < (int)(c1 - c0 + 1) * NC; i += blockDim.x) {
        const int c = i / NC, j = i - c * NC;
        int b = 0;
        while (b + 1 < T.ncolblocks && j >= T.blk_locoff[b] + T.blk_nd[b]) ++b;
        solc[i] = op.sol[T.blk_soloff[b] + T.blk_celldofs[b][(c0 + c) * T.blk_nd[b] + (j - T.blk_locoff[b])]];
    }
    unsigned short *pidx = reinterpret_cast<unsigned short *>(solc + (size_t)(128 / nq + 2) * NC);
    double *pv = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(pidx) + ((size_t)T.EC * NC * 2 + 15) / 16 * 16) + threadIdx.x;
    if constexpr (PVC) {
        __syncthreads();    // the tables (pt) are complete
        // B entry (x, j) -> index of its physical basis value: (values before this space) + k (1 + DIM) + source
        for (int i = threadIdx.x; i < T.EC * NC; i += blockDim.x) {
            const uchar4 e = pt[i];
            pidx[i] = (unsigned short)(T.phi_off[e.x] / nq + e.y * (1 + DIM) + e.z);
        }
    }
    __syncthreads();
    if (t >= ntot) return;
    const long long cell = t / nq;
    const int q = (int)(t - cell * nq);
    const double *sc_ = solc + (size_t)(cell - c0) * NC;
    CellGeo<DIM> G;
    load_geo<DIM>(op, cell, G);
    if constexpr (PVC) {
        for (int s = 0; s < T.nspaces; ++s) {
            const int ns = T.ns[s];
            const double *rv = T.refvals[s] + (size_t)q * ns, *rg = T.refgrads[s] + (size_t)q * ns * DIM;
            double *dst = pv + (size_t)(T.phi_off[s] / nq) * 128;
            for (int k = 0; k < ns; ++k, dst += (1 + DIM) * 128) {
                dst[0] = __ldg(rv + k);
                double r_[DIM];
#pragma unroll
                for (int r = 0; r < DIM; ++r) r_[r] = __ldg(rg + k * DIM + r);
#pragma unroll
                for (int d = 0; d < DIM; ++d) {
                    double v = 0.0;
#pragma unroll
                    for (int r = 0; r < DIM; ++r) v = fma(G.Ainv[r * DIM + d], r_[r], v);
                    dst[(1 + d) * 128] = v;
                }
            }
        }
    }
    // input_args: for every component o the (dof, entry) pairs that feed it (static register index)
    double u[NIO];
#pragma unroll
    for (int o = 0; o < NIO; ++o) {
        double a = 0.0;
        for (int p = uptr[o]; p < uptr[o + 1]; ++p) {
            const int j = ulist[p] & 0xff, x = ulist[p] >> 8;
            if constexpr (PVC) {
                const int i = x * NC + j;
                a = fma(sc_[j] * bgsc[i], pv[(size_t)pidx[i] * 128], a);
                continue;
            }
            const uchar4 e = pt[x * NC + j];
            double v;
            if (e.z == 0) v = __ldg(T.refvals[e.x] + q * T.ns[e.x] + e.y);
            else {
                const double *rg = T.refgrads[e.x] + ((size_t)q * T.ns[e.x] + e.y) * DIM;
                double r_[DIM];
#pragma unroll
                for (int r = 0; r < DIM; ++r) r_[r] = __ldg(rg + r);
                v = 0.0;
#pragma unroll
                for (int d = 0; d < DIM; ++d)          // select the column of A^-1 without a dynamic register index
                    if (e.z - 1 == d) {
#pragma unroll
                        for (int r = 0; r < DIM; ++r) v = fma(G.Ainv[r * DIM + d], r_[r], v);
                    }
            }
            a = fma(sc_[j] * bgsc[x * NC + j], v, a);
        }
        u[o] = a;
    }
    const double w = op.qw[q], sc = op.factor * w * G.vol;
    if constexpr (ROW && NIO == 9) {
        double val[9];
        NeoState S;
        neo_prepare(u, op.params, S, val);
#define EXTFEM_NEO_ROW(K)                                                                    \
        {                                                                                    \
            double row[9];                                                                   \
            neo_row<K>(S, row);                                                              \
            double sum = 0.0;                                                                \
            _Pragma("unroll") for (int d = 0; d < 9; ++d) {                                  \
                sum = fma(row[d], u[d], sum);                                                \
                const int sl = jslot[K * 9 + d];                                             \
                if (sl != 255) wJ[(size_t)sl * ntot + t] = row[d] * w;                       \
            }                                                                                \
            rqg[(size_t)K * ntot + t] = (sum - val[K]) * sc;                                 \
        }
        EXTFEM_NEO_ROW(0) EXTFEM_NEO_ROW(1) EXTFEM_NEO_ROW(2) EXTFEM_NEO_ROW(3) EXTFEM_NEO_ROW(4)
        EXTFEM_NEO_ROW(5) EXTFEM_NEO_ROW(6) EXTFEM_NEO_ROW(7) EXTFEM_NEO_ROW(8)
#undef EXTFEM_NEO_ROW
        return;
    }
    double val[NIO], J[NIO * NIO];
    nl_apply_fixed<DIM, NIO>(op.kernel_id, u, op.params, val, J, op.cellregions[cell]);
#pragma unroll
    for (int k = 0; k < NIO; ++k) {
        double sum = 0.0;
#pragma unroll
        for (int d = 0; d < NIO; ++d) {
            sum = fma(J[k * NIO + d], u[d], sum);
            const int sl = jslot[k * NIO + d];        // structural zeros of the Jacobian are not stored
            if (sl != 255) wJ[(size_t)sl * ntot + t] = J[k * NIO + d] * w;
        }
        rqg[(size_t)k * ntot + t] = (sum - val[k]) * sc;
    }
}

// doubles of shared memory per warp: PHI | B_q [rows][Np] | GJ_q [rows][Np] | w J [nq][nnzJ] (compact) | residual terms [nq][nin]
__host__ __device__ inline size_t nl4_warp_doubles(int nq, int nin, int nnzJ, int Np, int phid)
{
    const int rows = (nin + 3) / 4 * 4;
    size_t d = (size_t)(phid + (phid & 1)) + 2 * (size_t)rows * Np + (size_t)nq * nnzJ + (size_t)nq * nin + (size_t)nin;
    return (d + 1) & ~(size_t)1;
}

// NT = number of 8-wide dof tiles (NR = NC <= 8 NT); the NT x NT accumulator tiles of the cell matrix stay in registers over the
// quadrature loop, B_q and GJ_q exist for one point at a time.  KS = k-steps of 4 kernel components done on the tensor cores;
// R1: one more component (nin = 4 KS + 1, e.g. the 9 gradient components of a 3D displacement) handled as a rank-1 update on
// the FP64 FMA pipe instead of a padded k-step (B200: DMMA and DFMA share one peak, padding is paid in full).
// The right-hand side rides along as column NC of GJ_q when the last dof tile has a free column.
template <int DIM, int NT, int KS, bool R1>
__global__ void __launch_bounds__(128, 4)
local_nonlinear_kernel4(const __grid_constant__ OpDev op, const __grid_constant__ NL3Tables T, const double *__restrict__ wJ,
                        const double *__restrict__ rqg, double *__restrict__ loc, double *__restrict__ bloc, int cells_per_warp)
{
    constexpr int Np = NT <= 2 ? 20 : 36;      // row stride = 4 mod 16 doubles: the fragment loads below are bank-conflict free
    constexpr int KD = KS * 4;                 // components on the tensor cores
    constexpr int MT = (KD + 7) / 8;           // 8-row tiles of GJ_q
    constexpr int rows = (KD + (R1 ? 1 : 0) + 3) / 4 * 4;
    extern __shared__ __align__(16) double smem_d[];
    const int nin = op.nin, nq = op.nq, NR = T.NR, NC = T.NC, NRC = NR * NC, ECNC = T.EC * NC;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int phid = T.phi_off[T.nspaces];
    const int JS = T.nnzJ;                     // compact Jacobian entries per point
    const size_t wd = nl4_warp_doubles(nq, nin, JS, Np, phid);
    unsigned char *tb = reinterpret_cast<unsigned char *>(smem_d + (size_t)nwarp * wd);
    for (int i = threadIdx.x; i < T.tab_bytes / 16; i += blockDim.x) reinterpret_cast<uint4 *>(tb)[i] = __ldg(reinterpret_cast<const uint4 *>(T.tab) + i);
    const int *bgidx = reinterpret_cast<const int *>(tb + T.o_bgidx);
    const double *bgsc = reinterpret_cast<const double *>(tb + T.o_bgsc);
    const unsigned short *ddst = reinterpret_cast<const unsigned short *>(tb + T.o_ddst);
    const unsigned char *jslot = tb + T.o_jslot;
    double *PHI = smem_d + (size_t)warp * wd;
    double *Bq = PHI + phid + (phid & 1);      // [rows][Np] operator matrix of one quadrature point (16-byte aligned)
    double *Gq = Bq + rows * Np;               // [rows][Np] (w J_q) B_q
    double *Jq = Gq + rows * Np;               // [nq][nnzJ] w J of the cell's points, structural non-zeros only
    double *rq = Jq + (size_t)nq * JS;         // [nq][nin] (J u - F) factor w |T|
    double *j8 = rq + (size_t)nq * nin;        // [nin] last Jacobian row of the current point, dense (R1)
    for (int i = lane; i < 2 * rows * Np; i += 32) Bq[i] = 0.0;   // structural zeros and padding stay zero for every cell and point
    __syncthreads();
    const int gid = lane >> 2, tig = lane & 3;
    // compact slots of this lane's Jacobian fragment (rows gid [+8], columns 4 ks + tig, and the rank-1 column): -1 = structural zero
    int sa[MT][KS], sr[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        const int o = mt * 8 + gid;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const int i = ks * 4 + tig;
            const int sl = (o < nin && i < nin) ? jslot[o * nin + i] : 255;
            sa[mt][ks] = sl == 255 ? -1 : sl;
        }
        const int sl = (R1 && o < nin) ? jslot[o * nin + KD] : 255;
        sr[mt] = sl == 255 ? -1 : sl;
    }
    const int s8 = (R1 && lane < nin) ? jslot[KD * nin + lane] : 255;
    // this lane's entries of the operator matrix (at most 4 x 32 sparse entries: NC <= 32 dofs with <= 4 operator entries each):
    // dense place, scale and the PHI index as a linear function of the point -- nothing of it depends on the cell
    // (kept in registers when the accumulator tiles leave room: NT <= 3; the 16 tiles of NT = 4 need them all)
    constexpr bool HOIST = NT <= 3;
    constexpr int NIT = HOIST ? 4 : 1;
    int sdst[NIT], sp0[NIT], sdp[NIT];
    double ssc[NIT];
#pragma unroll
    for (int it = 0; it < NIT && HOIST; ++it) {
        const int i = it * 32 + lane;
        const bool live = i < ECNC && bgidx[i] >= 0;
        sdst[it] = live ? ddst[i] : -1;
        sp0[it] = live ? bgidx[i] : 0;
        sdp[it] = (live && nq > 1) ? bgidx[ECNC + i] - bgidx[i] : 0;
        ssc[it] = live ? bgsc[i] : 0.0;
    }
    // the lane whose accumulator fragment holds column NC carries the right-hand side
    const bool rhs_in_gemm = (NC & 7) != 0;
    const int rt = NC >> 3, rslot = NC & 1;
    const bool rmine = rhs_in_gemm && ((NC & 7) >> 1) == tig;
    const long long ntot = op.ncells * nq;
    const long long cbase = ((long long)blockIdx.x * nwarp + warp) * cells_per_warp;
    for (int ci = 0; ci < cells_per_warp; ++ci) {
        const long long cell = cbase + ci;
        if (cell >= op.ncells) break;
        CellGeo<DIM> G;
        load_geo<DIM>(op, cell, G);
        __syncwarp();
        {   // w J and the residual terms of the cell's points (nl_point_kernel wrote [entry][cell nq + q]); (entry, q) advance
            // with the lane stride without a division
            const double *wJc = wJ + cell * nq, *rqc = rqg + cell * nq;
            const int de = 32 / nq, dq = 32 - de * nq;
            int e = lane / nq, q = lane - e * nq;
            for (int i = lane; i < JS * nq; i += 32) {
                Jq[q * JS + e] = __ldg(wJc + (size_t)e * ntot + q);
                e += de; q += dq;
                if (q >= nq) { q -= nq; ++e; }
            }
            e = lane / nq; q = lane - e * nq;
            for (int i = lane; i < nin * nq; i += 32) {
                rq[q * nin + e] = __ldg(rqc + (size_t)e * ntot + q);
                e += de; q += dq;
                if (q >= nq) { q -= nq; ++e; }
            }
        }
        for (int sp = 0; sp < T.nspaces; ++sp) {
            double *ph = PHI + T.phi_off[sp];
            for (int e = lane; e < nq * T.ns[sp]; e += 32) {
                double *o = ph + (size_t)e * (1 + DIM);
                o[0] = __ldg(T.refvals[sp] + e);
                const double *rg = T.refgrads[sp] + (size_t)e * DIM;
#pragma unroll
                for (int d = 0; d < DIM; ++d) {
                    double g = 0.0;
#pragma unroll
                    for (int r = 0; r < DIM; ++r) g += G.Ainv[r * DIM + d] * __ldg(rg + r);
                    o[1 + d] = g;
                }
            }
        }
        double acc[NT][NT][2];
#pragma unroll
        for (int a = 0; a < NT; ++a)
#pragma unroll
            for (int b = 0; b < NT; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
        double racc = 0.0;
        __syncwarp();
        for (int q = 0; q < nq; ++q) {
            // this lane's fragment of w J_q (component rows gid [+8], input columns 4 ks + tig) and its residual terms
            const double *J = Jq + q * JS, *r = rq + q * nin;
            double ja[MT][KS], jr[MT], rr[MT];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                const int o = mt * 8 + gid;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) ja[mt][ks] = sa[mt][ks] >= 0 ? J[sa[mt][ks]] : 0.0;
                jr[mt] = (R1 && sr[mt] >= 0) ? J[sr[mt]] : 0.0;
                rr[mt] = (rmine && o < nin) ? r[o] : 0.0;
            }
            if (R1 && lane < nin) j8[lane] = s8 != 255 ? J[s8] : 0.0;
            // operator matrix of the point: the sparse entries of the B tables into their dense places
            if (HOIST) {
#pragma unroll
                for (int it = 0; it < NIT; ++it)
                    if (sdst[it] >= 0) Bq[sdst[it]] = ssc[it] * PHI[sp0[it] + q * sdp[it]];
            } else {
                for (int i = lane; i < ECNC; i += 32) { const int p = bgidx[q * ECNC + i]; if (p >= 0) Bq[ddst[i]] = bgsc[i] * PHI[p]; }
            }
            __syncwarp();
            // GJ_q = (w J_q) B_q: tiles of 8 components x 8 dofs, k-steps of 4 input components
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                double g[NT][2];
#pragma unroll
                for (int t = 0; t < NT; ++t) g[t][0] = g[t][1] = 0.0;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    const double *pb = Bq + (ks * 4 + tig) * Np + gid;
#pragma unroll
                    for (int t = 0; t < NT; ++t) dmma_m8n8k4(g[t][0], g[t][1], ja[mt][ks], pb[t * 8]);
                }
                if (R1) {
#pragma unroll
                    for (int t = 0; t < NT; ++t) {
                        const double2 b8 = *reinterpret_cast<const double2 *>(Bq + KD * Np + t * 8 + 2 * tig);
                        g[t][0] = fma(jr[mt], b8.x, g[t][0]);
                        g[t][1] = fma(jr[mt], b8.y, g[t][1]);
                    }
                }
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    if (t == rt && rmine) { if (rslot) g[t][1] = rr[mt]; else g[t][0] = rr[mt]; }
                    if (mt * 8 + gid < rows) *reinterpret_cast<double2 *>(Gq + (mt * 8 + gid) * Np + t * 8 + 2 * tig) = make_double2(g[t][0], g[t][1]);
                }
            }
            if (R1 && lane < NT * 8) {          // last component row on the FMA pipe
                double g = 0.0;
                for (int i = 0; i < nin; ++i) g = fma(j8[i], Bq[i * Np + lane], g);
                if (rhs_in_gemm && lane == NC) g = r[KD];
                Gq[KD * Np + lane] = g;
            }
            __syncwarp();
            // A += B_q^T GJ_q: every 8 x 8 tile of the cell matrix, k-steps of 4 components
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const double *pa = Bq + (ks * 4 + tig) * Np + gid, *pg = Gq + (ks * 4 + tig) * Np + gid;
                double av[NT], bv[NT];
#pragma unroll
                for (int t = 0; t < NT; ++t) { av[t] = pa[t * 8]; bv[t] = pg[t * 8]; }
#pragma unroll
                for (int a = 0; a < NT; ++a)
#pragma unroll
                    for (int b = 0; b < NT; ++b) dmma_m8n8k4(acc[a][b][0], acc[a][b][1], av[a], bv[b]);
            }
            if (R1) {
                double av[NT];
                double2 bv[NT];
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    av[t] = Bq[KD * Np + t * 8 + gid];
                    bv[t] = *reinterpret_cast<const double2 *>(Gq + KD * Np + t * 8 + 2 * tig);
                }
#pragma unroll
                for (int a = 0; a < NT; ++a)
#pragma unroll
                    for (int b = 0; b < NT; ++b) {
                        acc[a][b][0] = fma(av[a], bv[b].x, acc[a][b][0]);
                        acc[a][b][1] = fma(av[a], bv[b].y, acc[a][b][1]);
                    }
            }
            if (!rhs_in_gemm && lane < NR)     // no free column: rhs_k += sum_component (J u - F) B
                for (int o = 0; o < nin; ++o) racc = fma(r[o], Bq[o * Np + lane], racc);
            __syncwarp();
        }
        const double fv = G.visited ? op.factor * G.vol : 0.0;
        double *out = loc + (size_t)cell * NRC;
        double *bout = bloc + (size_t)cell * NR;
#pragma unroll
        for (int a = 0; a < NT; ++a) {
            const int k = a * 8 + gid;
#pragma unroll
            for (int b = 0; b < NT; ++b) {
                const int j = b * 8 + 2 * tig;
                if (k < NR && j < NC) out[(size_t)j * NR + k] = acc[a][b][0] * fv;
                if (k < NR && j + 1 < NC) out[(size_t)(j + 1) * NR + k] = acc[a][b][1] * fv;
                if (b == rt && rmine && k < NR) bout[k] = G.visited ? (rslot ? acc[a][b][1] : acc[a][b][0]) : 0.0;
            }
        }
        if (!rhs_in_gemm && lane < NR) bout[lane] = G.visited ? racc : 0.0;
    }
}

// x at quadrature points (extfem_quadrature_points)
template <int DIM>
__global__ void quadpoints_kernel(const __grid_constant__ OpDev op, double *__restrict__ xq)
{
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= op.ncells * op.nq) return;
    long long cell = t / op.nq;
    int q = (int)(t - cell * op.nq);
    CellGeo<DIM> G;
    load_geo<DIM>(op, cell, G);
    double x[DIM];
    eval_x<DIM>(G, op.qx + q * DIM, x);
    for (int d = 0; d < DIM; ++d) xq[t * DIM + d] = x[d];
}

} // namespace extfem
