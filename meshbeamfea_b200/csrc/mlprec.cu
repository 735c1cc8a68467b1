// Multilevel (rigid-body aggregate) preconditioner, device half -- SURVEY 8(f) rank 4.
//
//   z~ = r~ + P0~ ( D1^-1 + P1 ( D2^-1 + ... P_{L-1} A_L^-1 P_{L-1}^T ... ) P1^T ) P0~^T r~
//
// additive over the levels (no coarse SpMV in the apply), for CG on the block-Jacobi-scaled system K~:
//  * level 0 -> 1: the TENTATIVE prolongator in scaled variables, one 6x6 block per fine row,
//    T_v = L_v^T B(d_v)  (L_v: Cholesky factor of the row's diagonal block, B(d) = [[I, -[d]x], [0, I]] the
//    rigid-body modes of the aggregate about its centroid).  Stored in FP32 (144 B per row): the same rounded
//    values serve restriction and prolongation, so the preconditioner stays exactly symmetric.
//  * level l -> l+1 (l >= 1): general 6x6-block sparse P_l, tentative or smoothed P = (I - w D^-1 A) T;
//    Galerkin matrices A_{l+1} = P^T (A P) by two sparse block products over host-built patterns.
//  * every sum runs in a fixed order (member lists, transposed lists and pattern rows ascend): no atomics,
//    results are bit-reproducible run to run.
// All kernels are HBM/L2-bound gathers of 6x6 blocks; tensor cores do not apply.
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.cuh"
#include "multilevel_math.hpp"

namespace mbfea { namespace dev {

namespace {
constexpr int kT = 256;
inline unsigned blocks_for(int64_t n, int per = kT) { return (unsigned)((n + per - 1) / per); }
}  // namespace

// ---------------------------------------------------------------------------
// setup
// ---------------------------------------------------------------------------
// T_v = L_v^T B(d_v) in FP32, row-major 6x6 (one thread per (row, entry))
__global__ void __launch_bounds__(kT) k_ml_make_t0(int64_t nb, const double *__restrict__ lfac,
                                                   const double *__restrict__ off, float *__restrict__ t0)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t v = t / 36;
    if (v >= nb) return;
    const int e = (int)(t - v * 36), i = e / 6, j = e - 6 * i;
    const double *L = lfac + v * 36, *d = off + 3 * v;
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) s = fma(__ldg(L + 6 * k + i), rigid_entry(d, k, j), s);   // (L^T)[i][k] B[k][j]
    t0[t] = (float)s;
}

// D^-1 = S^T S per row (S = L^-1 lower triangular, full 6x6 storage), one thread per (row, entry)
__global__ void __launch_bounds__(kT) k_ml_diag_inv(int64_t n, const double *__restrict__ sinv, double *__restrict__ dinv)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t v = t / 36;
    if (v >= n) return;
    const int e = (int)(t - v * 36), i = e / 6, j = e - 6 * i;
    const double *S = sinv + v * 36;
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) s = fma(__ldg(S + 6 * k + i), __ldg(S + 6 * k + j), s);
    dinv[t] = s;
}

// P = T - w D^-1 A T on the pattern (prow, pcol): one thread per P block (i, I).
//   T(i, agg i) = B(off_i);   (A T)(i, I) = sum over blocks (i, j) with agg j == I of A_ij B(off_j)   (ascending j)
// w == 0: tentative prolongator (P has one block per row).
__global__ void __launch_bounds__(128) k_ml_make_p(int64_t n, const int32_t *__restrict__ prow,
                                                   const int32_t *__restrict__ pcol, const int32_t *__restrict__ prow_of,
                                                   const int32_t *__restrict__ arow, const int32_t *__restrict__ acol,
                                                   const double *__restrict__ avals, const int32_t *__restrict__ agg,
                                                   const double *__restrict__ off, const double *__restrict__ dinv,
                                                   double omega, double *__restrict__ pval)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= (int64_t)__ldg(prow + n)) return;
    const int32_t i = __ldg(prow_of + b), I = __ldg(pcol + b);
    double M[36];
#pragma unroll
    for (int e = 0; e < 36; ++e) M[e] = 0.0;
    if (omega != 0.0) {
        for (int32_t k = __ldg(arow + i); k < __ldg(arow + i + 1); ++k) {
            const int32_t j = __ldg(acol + k);
            if (__ldg(agg + j) != I) continue;
            const double *A = avals + (int64_t)k * 36, *d = off + 3 * (int64_t)j;
            const double dx = __ldg(d), dy = __ldg(d + 1), dz = __ldg(d + 2);
            // A B(d): columns 0..2 unchanged; column 3+c gets A[:, 0..2] (-[d]x)[:, c] + A[:, 3+c]
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                const double a0 = __ldg(A + 6 * r), a1 = __ldg(A + 6 * r + 1), a2 = __ldg(A + 6 * r + 2);
                M[6 * r + 0] += a0; M[6 * r + 1] += a1; M[6 * r + 2] += a2;
                // -[d]x = [[0, dz, -dy], [-dz, 0, dx], [dy, -dx, 0]]
                M[6 * r + 3] += __ldg(A + 6 * r + 3) + (-a1 * dz + a2 * dy);
                M[6 * r + 4] += __ldg(A + 6 * r + 4) + (a0 * dz - a2 * dx);
                M[6 * r + 5] += __ldg(A + 6 * r + 5) + (-a0 * dy + a1 * dx);
            }
        }
    }
    const bool own = __ldg(agg + i) == I;
    const double *d = off + 3 * (int64_t)i;
    const double *Di = dinv + (int64_t)i * 36;
    double *out = pval + b * 36;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            double s = 0.0;
            if (omega != 0.0) {
#pragma unroll
                for (int k = 0; k < 6; ++k) s = fma(__ldg(Di + 6 * r + k), M[6 * k + c], s);
                s *= -omega;
            }
            if (own) s += rigid_entry(d, r, c);
            out[6 * r + c] = s;
        }
    }
}

// C(i, J) = sum over k in row i of A, with (k, J) in P:  A_ik P_kJ   (one thread per block of C = A P; k ascending)
__global__ void __launch_bounds__(128) k_ml_ap(int64_t n, const int32_t *__restrict__ crow, const int32_t *__restrict__ ccol,
                                               const int32_t *__restrict__ crow_of, const int32_t *__restrict__ arow,
                                               const int32_t *__restrict__ acol, const double *__restrict__ avals,
                                               const int32_t *__restrict__ prow, const int32_t *__restrict__ pcol,
                                               const double *__restrict__ pval, double *__restrict__ cval)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= (int64_t)__ldg(crow + n)) return;
    const int32_t i = __ldg(crow_of + b), J = __ldg(ccol + b);
    double acc[36];
#pragma unroll
    for (int e = 0; e < 36; ++e) acc[e] = 0.0;
    for (int32_t ka = __ldg(arow + i); ka < __ldg(arow + i + 1); ++ka) {
        const int32_t k = __ldg(acol + ka);
        int32_t lo = __ldg(prow + k), hi = __ldg(prow + k + 1);
        while (lo < hi) {   // lower bound of J in P's row k
            const int32_t mid = (lo + hi) >> 1;
            if (__ldg(pcol + mid) < J) lo = mid + 1; else hi = mid;
        }
        if (lo >= __ldg(prow + k + 1) || __ldg(pcol + lo) != J) continue;
        const double *A = avals + (int64_t)ka * 36, *P = pval + (int64_t)lo * 36;
        double p[36];
#pragma unroll
        for (int e = 0; e < 36; ++e) p[e] = __ldg(P + e);
#pragma unroll
        for (int r = 0; r < 6; ++r) {
#pragma unroll
            for (int m = 0; m < 6; ++m) {
                const double a = __ldg(A + 6 * r + m);
#pragma unroll
                for (int c = 0; c < 6; ++c) acc[6 * r + c] = fma(a, p[6 * m + c], acc[6 * r + c]);
            }
        }
    }
    double *out = cval + b * 36;
#pragma unroll
    for (int e = 0; e < 36; ++e) out[e] = acc[e];
}

// G(I, J) = sum over the P blocks (i, I) of column I (ascending i), with (i, J) in C = A P:  P_iI^T C_iJ
__global__ void __launch_bounds__(128) k_ml_rap(int64_t nc, const int32_t *__restrict__ grow, const int32_t *__restrict__ gcol,
                                                const int32_t *__restrict__ grow_of, const int32_t *__restrict__ rrow,
                                                const int32_t *__restrict__ rblk, const int32_t *__restrict__ rfine,
                                                const double *__restrict__ pval, const int32_t *__restrict__ crow,
                                                const int32_t *__restrict__ ccol, const double *__restrict__ cval,
                                                double *__restrict__ gval)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= (int64_t)__ldg(grow + nc)) return;
    const int32_t I = __ldg(grow_of + b), J = __ldg(gcol + b);
    double acc[36];
#pragma unroll
    for (int e = 0; e < 36; ++e) acc[e] = 0.0;
    for (int32_t e = __ldg(rrow + I); e < __ldg(rrow + I + 1); ++e) {
        const int32_t i = __ldg(rfine + e);
        int32_t lo = __ldg(crow + i), hi = __ldg(crow + i + 1);
        const int32_t end = hi;
        while (lo < hi) {
            const int32_t mid = (lo + hi) >> 1;
            if (__ldg(ccol + mid) < J) lo = mid + 1; else hi = mid;
        }
        if (lo >= end || __ldg(ccol + lo) != J) continue;
        const double *P = pval + (int64_t)__ldg(rblk + e) * 36, *C = cval + (int64_t)lo * 36;
        double cc[36];
#pragma unroll
        for (int x = 0; x < 36; ++x) cc[x] = __ldg(C + x);
#pragma unroll
        for (int m = 0; m < 6; ++m) {
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                const double p = __ldg(P + 6 * m + r);   // (P^T)[r][m]
#pragma unroll
                for (int c = 0; c < 6; ++c) acc[6 * r + c] = fma(p, cc[6 * m + c], acc[6 * r + c]);
            }
        }
    }
    double *out = gval + b * 36;
#pragma unroll
    for (int e = 0; e < 36; ++e) out[e] = acc[e];
}

// row index of every block of a CSR pattern (setup helper): row_of[k] = i for k in [rowptr[i], rowptr[i+1])
__global__ void __launch_bounds__(kT) k_ml_row_of(int64_t n, const int32_t *__restrict__ rowptr, int32_t *__restrict__ row_of)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int32_t k = __ldg(rowptr + i); k < __ldg(rowptr + i + 1); ++k) row_of[k] = (int32_t)i;
}

// coarsest level: 6x6-block rows -> dense n x n (row-major), zero elsewhere (dense pre-zeroed)
__global__ void __launch_bounds__(kT) k_ml_blocks_to_dense(int64_t nrows, const int32_t *__restrict__ rowptr,
                                                           const int32_t *__restrict__ colidx,
                                                           const int32_t *__restrict__ row_of,
                                                           const double *__restrict__ vals, double *__restrict__ dense)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t k = t / 36;
    if (k >= (int64_t)__ldg(rowptr + nrows)) return;
    const int e = (int)(t - k * 36), r = e / 6, c = e - 6 * r;
    const int64_t n = 6 * nrows;
    dense[(6 * (int64_t)__ldg(row_of + k) + r) * n + 6 * (int64_t)__ldg(colidx + k) + c] = vals[t];
}

// In-place Gauss-Jordan inverse of a dense SPD matrix (no pivoting needed), ONE CTA of 1024 threads, the matrix
// (<= a few MB) stays L2-resident.  *flag |= 2 on a non-positive pivot.  Step k:
//   p = 1 / a_kk;  row k *= p (a_kk := p);  for i != k: f = a_ik, a_ik := 0, row i -= f * row k.
__global__ void __launch_bounds__(1024) k_ml_dense_inverse(int n, double *__restrict__ a, int *flag)
{
    __shared__ double piv;
    extern __shared__ double colk[];   // n doubles: column k before the step
    for (int k = 0; k < n; ++k) {
        if (threadIdx.x == 0) {
            const double d = a[(size_t)k * n + k];
            if (!(d > 0.0) || !isfinite(d)) { atomicOr(flag, 2); piv = 1.0; } else piv = 1.0 / d;
        }
        for (int i = threadIdx.x; i < n; i += blockDim.x) colk[i] = a[(size_t)i * n + k];
        __syncthreads();
        const double p = piv;
        // row k
        for (int j = threadIdx.x; j < n; j += blockDim.x) a[(size_t)k * n + j] = (j == k) ? p : a[(size_t)k * n + j] * p;
        __syncthreads();
        // every other row: a_ij -= f_i * a_kj (j != k), a_ik = -f_i * p
        for (int64_t t = threadIdx.x; t < (int64_t)n * n; t += blockDim.x) {
            const int i = (int)(t / n), j = (int)(t - (int64_t)i * n);
            if (i == k) continue;
            const double f = colk[i];
            a[t] = (j == k) ? -f * p : fma(-f, a[(size_t)k * n + j], a[t]);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// apply
// ---------------------------------------------------------------------------
// fine restriction: rc[I] = sum over the members v of aggregate I (ascending) of T_v^T r~_v.  G lanes per
// aggregate (8 when no aggregate has more than 8 members, else 32); lanes take members round-robin, then a
// shuffle tree in a fixed pattern within the group.
template <int G>
__global__ void __launch_bounds__(kT) k_ml_restrict0(int64_t n_agg, const int32_t *__restrict__ agg_ptr,
                                                     const int32_t *__restrict__ agg_rows, const float *__restrict__ t0,
                                                     const double *__restrict__ r, double *__restrict__ rc)
{
    const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int lane = threadIdx.x % G;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    if (gid < n_agg) {
        for (int32_t q = __ldg(agg_ptr + gid) + lane; q < __ldg(agg_ptr + gid + 1); q += G) {
            const int64_t v = __ldg(agg_rows + q);
            const float4 *T = reinterpret_cast<const float4 *>(t0 + v * 36);
            float tt[36];
#pragma unroll
            for (int x = 0; x < 9; ++x) { const float4 f = __ldg(T + x); tt[4 * x] = f.x; tt[4 * x + 1] = f.y; tt[4 * x + 2] = f.z; tt[4 * x + 3] = f.w; }
            const double2 r0 = *reinterpret_cast<const double2 *>(r + v * 6), r1 = *reinterpret_cast<const double2 *>(r + v * 6 + 2),
                          r2 = *reinterpret_cast<const double2 *>(r + v * 6 + 4);
            const double rv[6] = {r0.x, r0.y, r1.x, r1.y, r2.x, r2.y};
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int j = 0; j < 6; ++j) acc[j] = fma((double)tt[6 * i + j], rv[i], acc[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < 6; ++j)
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) acc[j] += __shfl_down_sync(0xffffffffu, acc[j], o, G);
    if (gid < n_agg && lane == 0) {
        double2 *o = reinterpret_cast<double2 *>(rc + gid * 6);
        o[0] = make_double2(acc[0], acc[1]); o[1] = make_double2(acc[2], acc[3]); o[2] = make_double2(acc[4], acc[5]);
    }
}

// fine prolongation + dot: z~_v = r~_v + T_v z1[agg v]; part[blockIdx] = sum of r~ . z~ over the CTA's entries.
// Thread = (row, scalar i): reads row i of T_v (24 B, consecutive threads -> consecutive addresses).
__global__ void __launch_bounds__(kT, 4) k_ml_finish0(int64_t nb, const int32_t *__restrict__ agg, const float *__restrict__ t0,
                                                      const double *__restrict__ rt, const double *__restrict__ z1,
                                                      double *__restrict__ zt, double *__restrict__ part)
{
    __shared__ double red[kT / 32];
    double dot = 0.0;
    const int64_t n = 6 * nb;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = t / 6;
        const double *zc = z1 + (int64_t)__ldg(agg + v) * 6;
        const float2 *T = reinterpret_cast<const float2 *>(t0 + t * 6);
        const float2 a = __ldg(T), b = __ldg(T + 1), c = __ldg(T + 2);
        const double ri = rt[t];
        double s = ri;
        s = fma((double)a.x, zc[0], s); s = fma((double)a.y, zc[1], s);
        s = fma((double)b.x, zc[2], s); s = fma((double)b.y, zc[3], s);
        s = fma((double)c.x, zc[4], s); s = fma((double)c.y, zc[5], s);
        zt[t] = s;
        dot = fma(ri, s, dot);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_down_sync(0xffffffffu, dot, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kT / 32; ++w) s += red[w];
        part[blockIdx.x] = s;
    }
}

// coarse restriction: rc[I] = sum over the P blocks (i, I) of column I (ascending) of P_iI^T r_i; one warp per coarse row
__global__ void __launch_bounds__(kT) k_ml_restrict(int64_t nc, const int32_t *__restrict__ rrow,
                                                    const int32_t *__restrict__ rblk, const int32_t *__restrict__ rfine,
                                                    const double *__restrict__ pval, const double *__restrict__ r,
                                                    double *__restrict__ rc)
{
    const int lane = threadIdx.x & 31;
    const int64_t I = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (I >= nc) return;   // whole warps leave together
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (int32_t e = __ldg(rrow + I) + lane; e < __ldg(rrow + I + 1); e += 32) {
        const double *P = pval + (int64_t)__ldg(rblk + e) * 36;
        const double *rv = r + (int64_t)__ldg(rfine + e) * 6;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const double ri = rv[i];
#pragma unroll
            for (int j = 0; j < 6; ++j) acc[j] = fma(__ldg(P + 6 * i + j), ri, acc[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < 6; ++j)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_down_sync(0xffffffffu, acc[j], o);
    if (lane == 0)
#pragma unroll
        for (int j = 0; j < 6; ++j) rc[I * 6 + j] = acc[j];
}

// coarse level on the way down: z_i = D_i^-1 r_i + sum over row i of P of P_iJ zc_J; thread = (row, scalar)
__global__ void __launch_bounds__(kT) k_ml_smooth_prolong(int64_t n, const double *__restrict__ dinv,
                                                          const int32_t *__restrict__ prow, const int32_t *__restrict__ pcol,
                                                          const double *__restrict__ pval, const double *__restrict__ r,
                                                          const double *__restrict__ zc, double *__restrict__ z)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t v = t / 6;
    if (v >= n) return;
    const int i = (int)(t - v * 6);
    const double *D = dinv + v * 36 + 6 * i, *rv = r + v * 6;
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 6; ++j) s = fma(__ldg(D + j), rv[j], s);
    for (int32_t b = __ldg(prow + v); b < __ldg(prow + v + 1); ++b) {
        const double *P = pval + (int64_t)b * 36 + 6 * i, *c = zc + (int64_t)__ldg(pcol + b) * 6;
#pragma unroll
        for (int j = 0; j < 6; ++j) s = fma(__ldg(P + j), c[j], s);
    }
    z[t] = s;
}

// ---- launchers ----------------------------------------------------------------------------------------------
void launch_ml_make_t0(cudaStream_t st, int64_t nb, const double *lfac, const double *off, float *t0)
{
    if (nb > 0) k_ml_make_t0<<<blocks_for(nb * 36), kT, 0, st>>>(nb, lfac, off, t0);
}
void launch_ml_diag_inv(cudaStream_t st, int64_t n, const double *sinv, double *dinv)
{
    if (n > 0) k_ml_diag_inv<<<blocks_for(n * 36), kT, 0, st>>>(n, sinv, dinv);
}
void launch_ml_row_of(cudaStream_t st, int64_t n, const int32_t *rowptr, int32_t *row_of)
{
    if (n > 0) k_ml_row_of<<<blocks_for(n), kT, 0, st>>>(n, rowptr, row_of);
}
void launch_ml_make_p(cudaStream_t st, int64_t n, int64_t nnzp, const int32_t *prow, const int32_t *pcol, const int32_t *prow_of,
                      const int32_t *arow, const int32_t *acol, const double *avals, const int32_t *agg, const double *off,
                      const double *dinv, double omega, double *pval)
{
    if (nnzp > 0) k_ml_make_p<<<blocks_for(nnzp, 128), 128, 0, st>>>(n, prow, pcol, prow_of, arow, acol, avals, agg, off, dinv, omega, pval);
}
void launch_ml_ap(cudaStream_t st, int64_t n, int64_t nnzc, const int32_t *crow, const int32_t *ccol, const int32_t *crow_of,
                  const int32_t *arow, const int32_t *acol, const double *avals, const int32_t *prow, const int32_t *pcol,
                  const double *pval, double *cval)
{
    if (nnzc > 0) k_ml_ap<<<blocks_for(nnzc, 128), 128, 0, st>>>(n, crow, ccol, crow_of, arow, acol, avals, prow, pcol, pval, cval);
}
void launch_ml_rap(cudaStream_t st, int64_t nc, int64_t nnzg, const int32_t *grow, const int32_t *gcol, const int32_t *grow_of,
                   const int32_t *rrow, const int32_t *rblk, const int32_t *rfine, const double *pval, const int32_t *crow,
                   const int32_t *ccol, const double *cval, double *gval)
{
    if (nnzg > 0) k_ml_rap<<<blocks_for(nnzg, 128), 128, 0, st>>>(nc, grow, gcol, grow_of, rrow, rblk, rfine, pval, crow, ccol, cval, gval);
}
void launch_ml_blocks_to_dense(cudaStream_t st, int64_t nrows, int64_t nnz, const int32_t *rowptr, const int32_t *colidx,
                               const int32_t *row_of, const double *vals, double *dense)
{
    if (nnz > 0) k_ml_blocks_to_dense<<<blocks_for(nnz * 36), kT, 0, st>>>(nrows, rowptr, colidx, row_of, vals, dense);
}
int ml_dense_inverse_prepare(int n)
{
    return (int)cudaFuncSetAttribute(k_ml_dense_inverse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(n * sizeof(double)));
}
void launch_ml_dense_inverse(cudaStream_t st, int n, double *a, int *flag)
{
    if (n > 0) k_ml_dense_inverse<<<1, 1024, (size_t)n * sizeof(double), st>>>(n, a, flag);
}
void launch_ml_restrict0(cudaStream_t st, int64_t n_agg, int group, const int32_t *agg_ptr, const int32_t *agg_rows,
                         const float *t0, const double *r, double *rc)
{
    if (n_agg <= 0) return;
    if (group == 8) k_ml_restrict0<8><<<blocks_for(n_agg * 8), kT, 0, st>>>(n_agg, agg_ptr, agg_rows, t0, r, rc);
    else k_ml_restrict0<32><<<blocks_for(n_agg * 32), kT, 0, st>>>(n_agg, agg_ptr, agg_rows, t0, r, rc);
}
void launch_ml_finish0(cudaStream_t st, int grid, int64_t nb, const int32_t *agg, const float *t0, const double *rt,
                       const double *z1, double *zt, double *part)
{
    if (nb > 0 && grid > 0) k_ml_finish0<<<(unsigned)grid, kT, 0, st>>>(nb, agg, t0, rt, z1, zt, part);
}
void launch_ml_restrict(cudaStream_t st, int64_t nc, const int32_t *rrow, const int32_t *rblk, const int32_t *rfine,
                        const double *pval, const double *r, double *rc)
{
    if (nc > 0) k_ml_restrict<<<blocks_for(nc * 32), kT, 0, st>>>(nc, rrow, rblk, rfine, pval, r, rc);
}
void launch_ml_smooth_prolong(cudaStream_t st, int64_t n, const double *dinv, const int32_t *prow, const int32_t *pcol,
                              const double *pval, const double *r, const double *zc, double *z)
{
    if (n > 0) k_ml_smooth_prolong<<<blocks_for(n * 6), kT, 0, st>>>(n, dinv, prow, pcol, pval, r, zc, z);
}

}}  // namespace mbfea::dev
