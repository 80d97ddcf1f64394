// fe_tables.h -- host-side quadrature rules and H1 Lagrange reference bases of the engine.
//
// These stand in for ExtendableFEMBase's QuadratureRule{Tv,EG}(order) and FEEvaluator
// reference tables (call sites: src/common_operators/bilinear_operator.jl:738,744-747),
// which are not part of the reference tree.  The Julia glue may override them with the
// tables of ExtendableFEMBase itself (extfem_opdesc::qweights/qpoints,
// extfem_space_set_tables), so the drop-in does not depend on these conventions.
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

namespace extfem {

// Golub-Welsch for the Jacobi weight (1-x)^alpha on [-1,1], mapped to [0,1].
// Returns nodes s_i in (0,1) ascending and weights normalised to sum 1.
inline void gauss_jacobi01(int n, double alpha, std::vector<double> &x, std::vector<double> &w)
{
    std::vector<double> a(n), b(n, 0.0);
    for (int k = 0; k < n; ++k) {
        double s = 2.0 * k + alpha;
        a[k] = (k == 0) ? -alpha / (alpha + 2.0) : -alpha * alpha / (s * (s + 2.0));
        if (k >= 1) {
            double kk = k;
            double num = 4.0 * kk * (kk + alpha) * kk * (kk + alpha);
            double den = s * s * (s + 1.0) * (s - 1.0);
            b[k] = std::sqrt(num / den);
        }
    }
    // cyclic Jacobi eigenvalue iteration on the symmetric tridiagonal matrix (n is tiny)
    std::vector<double> M(n * n, 0.0), V(n * n, 0.0);
    for (int i = 0; i < n; ++i) {
        M[i * n + i] = a[i];
        V[i * n + i] = 1.0;
        if (i + 1 < n) M[i * n + i + 1] = M[(i + 1) * n + i] = b[i + 1];
    }
    for (int sweep = 0; sweep < 100; ++sweep) {
        double off = 0;
        for (int i = 0; i < n; ++i)
            for (int j = i + 1; j < n; ++j) off += M[i * n + j] * M[i * n + j];
        if (off < 1e-300) break;
        for (int p = 0; p < n; ++p)
            for (int q = p + 1; q < n; ++q) {
                if (std::fabs(M[p * n + q]) < 1e-300) continue;
                double theta = (M[q * n + q] - M[p * n + p]) / (2.0 * M[p * n + q]);
                double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; ++k) {
                    double mkp = M[k * n + p], mkq = M[k * n + q];
                    M[k * n + p] = c * mkp - s * mkq;
                    M[k * n + q] = s * mkp + c * mkq;
                }
                for (int k = 0; k < n; ++k) {
                    double mpk = M[p * n + k], mqk = M[q * n + k];
                    M[p * n + k] = c * mpk - s * mqk;
                    M[q * n + k] = s * mpk + c * mqk;
                }
                for (int k = 0; k < n; ++k) {
                    double vkp = V[k * n + p], vkq = V[k * n + q];
                    V[k * n + p] = c * vkp - s * vkq;
                    V[k * n + q] = s * vkp + c * vkq;
                }
            }
    }
    std::vector<int> idx(n);
    for (int i = 0; i < n; ++i) idx[i] = i;
    std::sort(idx.begin(), idx.end(), [&](int i, int j) { return M[i * n + i] < M[j * n + j]; });
    x.resize(n);
    w.resize(n);
    double sum = 0;
    for (int i = 0; i < n; ++i) {
        x[i] = 0.5 * M[idx[i] * n + idx[i]] + 0.5;
        w[i] = V[0 * n + idx[i]] * V[0 * n + idx[i]];
        sum += w[i];
    }
    for (int i = 0; i < n; ++i) w[i] /= sum;
}

struct QuadRule {
    int dim = 0, nq = 0;
    std::vector<double> w;   // [nq], sums to 1
    std::vector<double> x;   // [nq][dim]
};

// Same conventions as documented in oracle/fetables.py (independent implementation).
inline QuadRule quadrature_rule(int dim, int order)
{
    QuadRule Q;
    Q.dim = dim;
    auto push = [&](std::initializer_list<double> p, double wt) {
        for (double v : p) Q.x.push_back(v);
        Q.w.push_back(wt);
    };
    if (dim == 0) {
        Q.w.push_back(1.0);   // a vertex (boundary "face" of a 1D grid): one point of weight 1, no coordinates
    } else if (order <= 1) {
        if (dim == 1) push({0.5}, 1.0);
        if (dim == 2) push({1.0 / 3, 1.0 / 3}, 1.0);
        if (dim == 3) push({0.25, 0.25, 0.25}, 1.0);
    } else if (order == 2) {
        if (dim == 1) { push({0.0}, 1.0 / 6); push({1.0}, 1.0 / 6); push({0.5}, 2.0 / 3); }
        if (dim == 2) { push({0.5, 0.5}, 1.0 / 3); push({0.0, 0.5}, 1.0 / 3); push({0.5, 0.0}, 1.0 / 3); }
        if (dim == 3) {
            const double a = 0.1381966011250105, b = 0.5854101966249685;
            push({a, a, a}, 0.25); push({b, a, a}, 0.25); push({a, b, a}, 0.25); push({a, a, b}, 0.25);
        }
    } else {
        int n = order / 2 + 1;
        std::vector<double> r, a, s, b, t, c;
        gauss_jacobi01(n, 0.0, r, a);
        gauss_jacobi01(n, 1.0, s, b);
        gauss_jacobi01(n, 2.0, t, c);
        if (dim == 1)
            for (int i = 0; i < n; ++i) push({r[i]}, a[i]);
        if (dim == 2)
            for (int j = 0; j < n; ++j)
                for (int i = 0; i < n; ++i) push({s[j], r[i] * (1 - s[j])}, a[i] * b[j]);
        if (dim == 3)
            for (int k = 0; k < n; ++k)
                for (int j = 0; j < n; ++j)
                    for (int i = 0; i < n; ++i)
                        push({t[k], s[j] * (1 - t[k]), r[i] * (1 - s[j]) * (1 - t[k])}, a[i] * b[j] * c[k]);
    }
    Q.nq = (int)Q.w.size();
    double sum = 0;
    for (double v : Q.w) sum += v;
    for (double &v : Q.w) v /= sum;
    return Q;
}

inline int nscalar_of(int order, int dim)
{
    if (dim == 0) return 1;
    if (order == 1) return dim + 1;
    if (order == 2) return dim == 1 ? 3 : (dim == 2 ? 6 : 10);
    return -1;
}

// Scalar reference basis: vals[nq][ns], grads[nq][ns][dim]
inline void ref_basis(int order, int dim, const QuadRule &Q, std::vector<double> &vals, std::vector<double> &grads)
{
    static const int tri_e[3][2] = {{0, 1}, {1, 2}, {2, 0}};
    static const int tet_e[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
    static const int seg_e[1][2] = {{0, 1}};
    int ns = nscalar_of(order, dim), nq = Q.nq;
    vals.assign((size_t)nq * ns, 0.0);
    grads.assign((size_t)nq * ns * dim, 0.0);
    if (dim == 0) { for (int q = 0; q < nq; ++q) vals[q] = 1.0; return; }
    for (int q = 0; q < nq; ++q) {
        double lam[4], dlam[4][3];
        double s = 0;
        for (int d = 0; d < dim; ++d) s += Q.x[q * dim + d];
        lam[0] = 1.0 - s;
        for (int d = 0; d < dim; ++d) { lam[d + 1] = Q.x[q * dim + d]; }
        for (int i = 0; i <= dim; ++i)
            for (int d = 0; d < dim; ++d) dlam[i][d] = (i == 0) ? -1.0 : (i - 1 == d ? 1.0 : 0.0);
        double *v = &vals[(size_t)q * ns];
        double *g = &grads[(size_t)q * ns * dim];
        if (order == 1) {
            for (int i = 0; i <= dim; ++i) {
                v[i] = lam[i];
                for (int d = 0; d < dim; ++d) g[i * dim + d] = dlam[i][d];
            }
        } else {
            for (int i = 0; i <= dim; ++i) {
                v[i] = lam[i] * (2 * lam[i] - 1);
                for (int d = 0; d < dim; ++d) g[i * dim + d] = (4 * lam[i] - 1) * dlam[i][d];
            }
            int ne = ns - dim - 1;
            const int(*E)[2] = dim == 1 ? seg_e : (dim == 2 ? tri_e : tet_e);
            for (int e = 0; e < ne; ++e) {
                int a = E[e][0], b = E[e][1], i = dim + 1 + e;
                v[i] = 4 * lam[a] * lam[b];
                for (int d = 0; d < dim; ++d) g[i * dim + d] = 4 * (lam[a] * dlam[b][d] + lam[b] * dlam[a][d]);
            }
        }
    }
}

// ---- host-supplied polynomial reference bases (EXTFEM_FE_TABULATED, extfem_space_set_tables) ----------------------
// monomials x^i y^j z^k, i+j+k <= order, enumerated `for k: for j: for i` (i fastest)
inline int nmonomials(int order, int dim)
{
    int n = 1;
    for (int d = 1; d <= dim; ++d) n = n * (order + d) / d;
    return n;
}

inline void poly_basis(int order, int dim, int ns, const std::vector<double> &coeffs, const QuadRule &Q, std::vector<double> &vals,
                       std::vector<double> &grads)
{
    const int nq = Q.nq, nm = nmonomials(order, dim);
    vals.assign((size_t)nq * ns, 0.0);
    grads.assign((size_t)nq * ns * std::max(dim, 1), 0.0);
    if (dim == 0) { for (int q = 0; q < nq; ++q) for (int s = 0; s < ns; ++s) vals[(size_t)q * ns + s] = coeffs[s]; return; }
    std::vector<double> mv(nm), mg((size_t)nm * dim);
    for (int q = 0; q < nq; ++q) {
        double x[3] = {0, 0, 0};
        for (int d = 0; d < dim; ++d) x[d] = Q.x[(size_t)q * dim + d];
        int m = 0;
        const int K = dim >= 3 ? order : 0;
        for (int k = 0; k <= K; ++k) {
            const int J = dim >= 2 ? order - k : 0;
            for (int j = 0; j <= J; ++j)
                for (int i = 0; i <= order - k - j; ++i, ++m) {
                    const int e[3] = {i, j, k};
                    double v = 1.0;
                    for (int d = 0; d < dim; ++d) v *= std::pow(x[d], e[d]);
                    mv[m] = v;
                    for (int d = 0; d < dim; ++d) {
                        double g = e[d];
                        if (e[d] > 0)
                            for (int d2 = 0; d2 < dim; ++d2) g *= std::pow(x[d2], d2 == d ? e[d2] - 1 : e[d2]);
                        mg[(size_t)m * dim + d] = e[d] > 0 ? g : 0.0;
                    }
                }
        }
        for (int s = 0; s < ns; ++s) {
            double v = 0.0, g[3] = {0, 0, 0};
            for (int mm = 0; mm < nm; ++mm) {
                const double c = coeffs[(size_t)s * nm + mm];
                v += c * mv[mm];
                for (int d = 0; d < dim; ++d) g[d] += c * mg[(size_t)mm * dim + d];
            }
            vals[(size_t)q * ns + s] = v;
            for (int d = 0; d < dim; ++d) grads[((size_t)q * ns + s) * dim + d] = g[d];
        }
    }
}

} // namespace extfem
