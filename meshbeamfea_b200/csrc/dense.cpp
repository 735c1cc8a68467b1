// Dense SPD inverse on the host for the coarsest level of the multilevel preconditioner (SURVEY 8(f) rank 4;
// not on the solver path yet).  The coarsest Galerkin matrix has a few hundred 6x6 block rows; its explicit
// inverse turns the coarse solve of every CG iteration into one small dense matrix-vector product on the
// device.  Cholesky A = L L^T (right-looking, panels of 48 columns, trailing update on all cores), then
// A^-1 = U U^T with U = L^-T.  Row-major; every inner product runs over contiguous rows.
#include <algorithm>
#include <cmath>
#include <thread>
#include <vector>

#include "mbfea_internal.hpp"

namespace mbfea {

// fn(i) for i in [0, n), index i on thread i % threads: the loops below have triangular cost profiles
template <class Fn>
static void par_cyclic(int64_t n, int64_t min_per_thread, Fn fn)
{
    const unsigned hw = std::thread::hardware_concurrency();
    const int64_t nt = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(hw ? hw : 1, 32), n / std::max<int64_t>(1, min_per_thread)));
    if (nt <= 1) { for (int64_t i = 0; i < n; ++i) fn(i); return; }
    std::vector<std::thread> th;
    for (int64_t t = 0; t < nt; ++t) th.emplace_back([=] { for (int64_t i = t; i < n; i += nt) fn(i); });
    for (auto &x : th) x.join();
}

// returns 0, or 1 + the index of the first non-positive pivot
int dense_spd_inverse(int64_t n, const double *A, double *Ainv)
{
    if (n <= 0) return 0;
    std::vector<double> Lm(A, A + n * n);   // lower triangle becomes L
    double *L = Lm.data();
    constexpr int64_t kPanel = 48;
    for (int64_t j0 = 0; j0 < n; j0 += kPanel) {
        const int64_t j1 = std::min(n, j0 + kPanel);
        // diagonal block: unblocked Cholesky of the kPanel x kPanel corner (its trailing updates applied already)
        for (int64_t j = j0; j < j1; ++j) {
            double d = L[j * n + j];
            for (int64_t k = j0; k < j; ++k) d -= L[j * n + k] * L[j * n + k];
            if (!(d > 0.0) || !std::isfinite(d)) return (int)(1 + j);
            const double piv = std::sqrt(d);
            L[j * n + j] = piv;
            for (int64_t i = j + 1; i < j1; ++i) {
                double s = L[i * n + j];
                for (int64_t k = j0; k < j; ++k) s -= L[i * n + k] * L[j * n + k];
                L[i * n + j] = s / piv;
            }
        }
        // panel below it: row i of L21 solves  L21[i] L11^T = A21[i]  -- rows are independent
        par_cyclic(n - j1, 32, [&](int64_t r) {
            const int64_t i = j1 + r;
            for (int64_t j = j0; j < j1; ++j) {
                double s = L[i * n + j];
                for (int64_t k = j0; k < j; ++k) s -= L[i * n + k] * L[j * n + k];
                L[i * n + j] = s / L[j * n + j];
            }
        });
        // trailing update: A22 -= L21 L21^T (lower triangle only)
        par_cyclic(n - j1, 16, [&](int64_t r) {
            const int64_t i = j1 + r;
            const double *li = L + i * n + j0;
            for (int64_t c = j1; c <= i; ++c) {
                const double *lc = L + c * n + j0;
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;   // four chains: the panel width is a multiple of 4
                for (int64_t k = 0; k + 3 < j1 - j0; k += 4) {
                    s0 += li[k] * lc[k]; s1 += li[k + 1] * lc[k + 1]; s2 += li[k + 2] * lc[k + 2]; s3 += li[k + 3] * lc[k + 3];
                }
                for (int64_t k = (j1 - j0) & ~(int64_t)3; k < j1 - j0; ++k) s0 += li[k] * lc[k];
                L[i * n + c] -= (s0 + s1) + (s2 + s3);
            }
        });
    }
    // U = L^-T (upper triangular, row-major): row c of U is column c of L^-1, from L y = e_c (y = 0 above c).
    // Every inner product below runs over contiguous row segments.
    std::vector<double> Um((size_t)(n * n), 0.0);
    double *U = Um.data();
    par_cyclic(n, 8, [&](int64_t c) {
        double *y = U + c * n;
        for (int64_t i = c; i < n; ++i) {
            const double *li = L + i * n;
            double s0 = (i == c) ? 1.0 : 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int64_t k = c;
            for (; k + 3 < i; k += 4) { s0 -= li[k] * y[k]; s1 -= li[k + 1] * y[k + 1]; s2 -= li[k + 2] * y[k + 2]; s3 -= li[k + 3] * y[k + 3]; }
            for (; k < i; ++k) s0 -= li[k] * y[k];
            y[i] = ((s0 + s1) + (s2 + s3)) / li[i];
        }
    });
    // A^-1 = L^-T L^-1 = U U^T:  (i, j <= i) = sum over k >= i of U[i][k] U[j][k]
    par_cyclic(n, 8, [&](int64_t i) {
        for (int64_t j = 0; j <= i; ++j) {
            const double *ui = U + i * n, *uj = U + j * n;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int64_t k = i;
            for (; k + 3 < n; k += 4) { s0 += ui[k] * uj[k]; s1 += ui[k + 1] * uj[k + 1]; s2 += ui[k + 2] * uj[k + 2]; s3 += ui[k + 3] * uj[k + 3]; }
            for (; k < n; ++k) s0 += ui[k] * uj[k];
            const double sum = (s0 + s1) + (s2 + s3);
            Ainv[i * n + j] = sum; Ainv[j * n + i] = sum;
        }
    });
    return 0;
}

}  // namespace mbfea
