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
__global__ void __launch_bounds__(kT) k_ml_diag_inv(int64_t n, const double *__restrict__ sinv, double *__restrict__ dinv,
                                                    float *__restrict__ dinv32)
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
    dinv32[t] = (float)s;
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

// ---------------------------------------------------------------------------
// Galerkin matrix of level 1 with the tentative fine prolongator:  coarse block k = sum over its contributor list
// (ascending fine blocks q) of B(d_a)^T A_q B(d_b), a = frow[q], b = fcol[q]  (A = the assembled, unscaled K).
// EIGHT lanes per coarse block: lanes take contributors round-robin (each reads its own 288-byte block: whole
// sectors), keep the 6x6 partial sum in registers, then a fixed shuffle tree over the 8 lanes; lane 0 stores.
// Diagonal coarse blocks have ~30 contributors, off-diagonal ones ~4: eight lanes keep both busy.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void rigid_sandwich(const double *__restrict__ A, const double *da, const double *db, double *acc)
{
    double M[36];
    const double bx = db[0], by = db[1], bz = db[2];
#pragma unroll
    for (int r = 0; r < 6; ++r) {   // M = A B(db)
        const double2 a01 = __ldg(reinterpret_cast<const double2 *>(A + 6 * r)), a23 = __ldg(reinterpret_cast<const double2 *>(A + 6 * r + 2)),
                      a45 = __ldg(reinterpret_cast<const double2 *>(A + 6 * r + 4));
        M[6 * r + 0] = a01.x; M[6 * r + 1] = a01.y; M[6 * r + 2] = a23.x;
        M[6 * r + 3] = a23.y + (-a01.y * bz + a23.x * by);
        M[6 * r + 4] = a45.x + (a01.x * bz - a23.x * bx);
        M[6 * r + 5] = a45.y + (-a01.x * by + a01.y * bx);
    }
    const double ax = da[0], ay = da[1], az = da[2];
#pragma unroll
    for (int c = 0; c < 6; ++c) {   // acc += B(da)^T M: rows 0..2 unchanged, row 3+r += ([da]x M[0..2, :])_r
        const double m0 = M[c], m1 = M[6 + c], m2 = M[12 + c];
        acc[c] += m0; acc[6 + c] += m1; acc[12 + c] += m2;
        acc[18 + c] += M[18 + c] + (ay * m2 - az * m1);
        acc[24 + c] += M[24 + c] + (az * m0 - ax * m2);
        acc[30 + c] += M[30 + c] + (ax * m1 - ay * m0);
    }
}

__global__ void __launch_bounds__(128) k_ml_galerkin0(int64_t n_coarse_blocks, const int32_t *__restrict__ gal_ptr,
                                                      const int32_t *__restrict__ gal_src, const int32_t *__restrict__ frow,
                                                      const int32_t *__restrict__ fcol, const double *__restrict__ fvals,
                                                      const double *__restrict__ off, double *__restrict__ cvals)
{
    const int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int lane = threadIdx.x & 7;
    double acc[36];
#pragma unroll
    for (int e = 0; e < 36; ++e) acc[e] = 0.0;
    if (k < n_coarse_blocks) {
        for (int32_t s = __ldg(gal_ptr + k) + lane; s < __ldg(gal_ptr + k + 1); s += 8) {
            const int32_t q = __ldg(gal_src + s);
            const double *da = off + 3 * (int64_t)__ldg(frow + q), *db = off + 3 * (int64_t)__ldg(fcol + q);
            const double dav[3] = {__ldg(da), __ldg(da + 1), __ldg(da + 2)}, dbv[3] = {__ldg(db), __ldg(db + 1), __ldg(db + 2)};
            rigid_sandwich(fvals + (int64_t)q * 36, dav, dbv, acc);
        }
    }
#pragma unroll
    for (int e = 0; e < 36; ++e)
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) acc[e] += __shfl_down_sync(0xffffffffu, acc[e], o, 8);
    if (k < n_coarse_blocks && lane == 0) {
        double2 *out = reinterpret_cast<double2 *>(cvals + k * 36);
#pragma unroll
        for (int e = 0; e < 18; ++e) out[e] = make_double2(acc[2 * e], acc[2 * e + 1]);
    }
}

// ---------------------------------------------------------------------------
// Explicit inverse of the coarsest matrix (dense n x n, n = 6 m, SPD): block Gauss-Jordan with 24-wide pivot blocks, three
// launches per pivot block (no pivot search: SPD).  Step over the pivot columns [k0, k0 + bw), bw <= 24:
//   invert: Pinv = (A_kk)^-1 (one warp, a row per lane in registers)
//   rows  : the column block is saved; block row k := Pinv A_k,: (A_kk := Pinv)                        -- n threads
//   update: A_ij := [j outside the pivot columns] A_ij - sum_c C_ic R_cj for i outside the pivot rows, C the saved column
//           block, R the new block row -- a rank-24 update in 64 x 64 tiles (4 x 4 outputs per thread)  -- (n / 64)^2 CTAs
// The matrix stays L2-resident; wider pivots cut the passes over it 4x against 6-wide ones (n = 2058: 15 ms -> 3 ms).
// *flag |= 2 on a non-positive pivot.
// ---------------------------------------------------------------------------
constexpr int kGjW = 24;

// step 1 (one warp): Pinv = inverse of the pivot block [k0, k0 + bw)^2 (padded with the identity to 24 x 24), by
// Gauss-Jordan with a row per lane in registers, -> pinv[24 * 24]
__global__ void __launch_bounds__(32) k_ml_gj_invert(int n, int k0, int bw, const double *__restrict__ a, double *__restrict__ pinv, int *flag)
{
    const int i = threadIdx.x;
    double row[kGjW];
#pragma unroll
    for (int j = 0; j < kGjW; ++j) row[j] = (i < bw && j < bw) ? a[(size_t)(k0 + i) * n + k0 + j] : (i == j ? 1.0 : 0.0);
    bool bad = false;
#pragma unroll
    for (int p = 0; p < kGjW; ++p) {
        double d = __shfl_sync(0xffffffffu, row[p], p);
        if (!(d > 0.0) || !isfinite(d)) { bad = true; d = 1.0; }
        const double inv = 1.0 / d;
        const double f = row[p];   // this row's entry in the pivot column
#pragma unroll
        for (int j = 0; j < kGjW; ++j) {
            const double pj = __shfl_sync(0xffffffffu, row[j], p) * inv;   // scaled pivot row, entry j (j != p)
            if (i == p) row[j] = (j == p) ? inv : pj;
            else row[j] = (j == p) ? -f * inv : fma(-f, pj, row[j]);
        }
    }
    if (i < kGjW)
#pragma unroll
        for (int j = 0; j < kGjW; ++j) pinv[i * kGjW + j] = row[j];
    if (bad && i == 0) atomicOr(flag, 2);
}

// step 2: thread t < n as a COLUMN: new block row k := Pinv (old pivot rows), Pinv itself in the pivot columns;
//         thread t < n as a ROW outside the pivot rows: its pivot-column entries saved to colk[t][0 .. 24)
// (the two halves touch disjoint entries: pivot rows are written, other rows are read)
__global__ void __launch_bounds__(128) k_ml_gj_rows(int n, int k0, int bw, double *__restrict__ a, const double *__restrict__ pinv,
                                                    double *__restrict__ colk)
{
    __shared__ double P[kGjW * kGjW];
    for (int t = threadIdx.x; t < kGjW * kGjW; t += blockDim.x) P[t] = pinv[t];
    __syncthreads();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    if (t < k0 || t >= k0 + bw) {
        const double *src = a + (size_t)t * n + k0;
        double *dst = colk + (size_t)t * kGjW;
#pragma unroll
        for (int c = 0; c < kGjW; ++c) dst[c] = c < bw ? src[c] : 0.0;
    }
    double v[kGjW];
#pragma unroll
    for (int c = 0; c < kGjW; ++c) v[c] = c < bw ? a[(size_t)(k0 + c) * n + t] : 0.0;
    const bool in = t >= k0 && t < k0 + bw;
    for (int r = 0; r < bw; ++r) {
        double s2 = 0.0;
        if (in) s2 = P[r * kGjW + (t - k0)];
        else
#pragma unroll
            for (int c = 0; c < kGjW; ++c) s2 = fma(P[r * kGjW + c], v[c], s2);
        a[(size_t)(k0 + r) * n + t] = s2;
    }
}

// step 3: rank-24 update of every row outside the pivot rows, 64 x 64 tiles, 4 x 4 outputs per thread (rows tr + x,
// columns tc + 16 y: consecutive lanes -> consecutive columns, conflict-free shared-memory reads and coalesced stores)
__global__ void __launch_bounds__(256) k_ml_gj_update(int n, int k0, int bw, double *__restrict__ a, const double *__restrict__ colk)
{
    __shared__ double Cs[64][kGjW + 1];
    __shared__ double Rs[kGjW][64];
    const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
    for (int t = threadIdx.x; t < 64 * kGjW; t += 256) {
        const int r = t / kGjW, c = t - r * kGjW;
        const int i = i0 + r;
        Cs[r][c] = (i < n && (i < k0 || i >= k0 + bw)) ? colk[(size_t)i * kGjW + c] : 0.0;
    }
    for (int t = threadIdx.x; t < kGjW * 64; t += 256) {
        const int c = t / 64, j = t - c * 64;
        Rs[c][j] = (c < bw && j0 + j < n) ? a[(size_t)(k0 + c) * n + j0 + j] : 0.0;
    }
    __syncthreads();
    const int tr = (threadIdx.x / 16) * 4, tc = threadIdx.x % 16;
    double acc[4][4];
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = 0.0;
#pragma unroll 8
    for (int c = 0; c < kGjW; ++c) {
        double cv[4], rv[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) { cv[x] = Cs[tr + x][c]; rv[x] = Rs[c][tc + 16 * x]; }
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) acc[x][y] = fma(cv[x], rv[y], acc[x][y]);
    }
#pragma unroll
    for (int x = 0; x < 4; ++x) {
        const int i = i0 + tr + x;
        if (i >= n || (i >= k0 && i < k0 + bw)) continue;   // the pivot rows were finished by k_ml_gj_rows
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            const int j = j0 + tc + 16 * y;
            if (j >= n) continue;
            double *q = a + (size_t)i * n + j;
            *q = ((j >= k0 && j < k0 + bw) ? 0.0 : *q) - acc[x][y];
        }
    }
}

// symmetrised FP32 copy of the inverse for the apply (rounding the same double for (i, j) and (j, i) keeps it symmetric)
__global__ void __launch_bounds__(256) k_ml_dense_to_f32(int n, const double *__restrict__ a, float *__restrict__ out)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n * n) return;
    const int i = (int)(t / n), j = (int)(t - (int64_t)i * n);
    out[t] = (float)(0.5 * (a[t] + a[(size_t)j * n + i]));
}

// coarsest level: z = Ainv r, Ainv dense n x n FP32 (L2-resident); one warp per row
__global__ void __launch_bounds__(kT) k_ml_dense_matvec(int n, const float *__restrict__ ainv, const double *__restrict__ r,
                                                        double *__restrict__ z)
{
    const int lane = threadIdx.x & 31;
    const int row = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (row >= n) return;
    const float *A = ainv + (size_t)row * n;
    double s = 0.0;
    for (int j = lane; j < n; j += 32) s = fma((double)__ldg(A + j), r[j], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) z[row] = s;
}

// FP32 copies for the apply: the same rounded values serve restriction and prolongation
__global__ void __launch_bounds__(kT) k_ml_to_f32(int64_t n, const double *__restrict__ a, float *__restrict__ out)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) out[t] = (float)a[t];
}

// 16-bit transfer blocks (multilevel_math.hpp): 36 values / largest magnitude as halfs, + that magnitude; thread per block
template <class S>
__global__ void __launch_bounds__(kT) k_ml_pack16(int64_t nblk, const S *__restrict__ src, unsigned char *__restrict__ dst)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    float v[36];
    float m = 0.0f;
#pragma unroll
    for (int x = 0; x < 36; ++x) { v[x] = (float)src[b * 36 + x]; m = fmaxf(m, fabsf(v[x])); }
    const float inv = m > 0.0f ? 1.0f / m : 0.0f;
    uint32_t w[20];
#pragma unroll
    for (int x = 0; x < 18; ++x) {
        const __half2 h = __floats2half2_rn(v[2 * x] * inv, v[2 * x + 1] * inv);
        w[x] = *reinterpret_cast<const uint32_t *>(&h);
    }
    w[18] = __float_as_uint(m); w[19] = 0u;
    uint4 *o = reinterpret_cast<uint4 *>(dst + b * kTBlock16Bytes);
#pragma unroll
    for (int x = 0; x < 5; ++x) o[x] = make_uint4(w[4 * x], w[4 * x + 1], w[4 * x + 2], w[4 * x + 3]);
}

// slices of the hierarchy image (several ranks): dst[i] = src[i] - delta, and the offsets of the ghost rows
__global__ void __launch_bounds__(kT) k_ml_shift_i32(int64_t n, const int32_t *__restrict__ src, int32_t delta, int32_t *__restrict__ dst)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i] - delta;
}
__global__ void __launch_bounds__(kT) k_ml_gather3(int64_t n, const double *__restrict__ src, const int64_t *__restrict__ idx,
                                                   double *__restrict__ dst)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < 3 * n) dst[t] = src[3 * idx[t / 3] + t % 3];
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

// ---------------------------------------------------------------------------
// apply
// ---------------------------------------------------------------------------
// fine restriction: rc[I] = sum over the members v of aggregate I (ascending) of T_v^T r~_v.  G lanes per
// aggregate (8 when no aggregate has more than 8 members, else 32); lanes take members round-robin, then a
// shuffle tree in a fixed pattern within the group.
template <int G, bool H16>
__global__ void __launch_bounds__(kT) k_ml_restrict0(int64_t n_agg, const int32_t *__restrict__ agg_ptr,
                                                     const int32_t *__restrict__ agg_rows, const void *__restrict__ t0,
                                                     const double *__restrict__ r, double *__restrict__ rc)
{
    const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int lane = threadIdx.x % G;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    if (gid < n_agg) {
        for (int32_t q = __ldg(agg_ptr + gid) + lane; q < __ldg(agg_ptr + gid + 1); q += G) {
            const int64_t v = __ldg(agg_rows + q);
            float tt[36];
            double sc;
            tblock_load<H16>(t0, v, tt, sc);
            const double2 r0 = *reinterpret_cast<const double2 *>(r + v * 6), r1 = *reinterpret_cast<const double2 *>(r + v * 6 + 2),
                          r2 = *reinterpret_cast<const double2 *>(r + v * 6 + 4);
            const double rv[6] = {sc * r0.x, sc * r0.y, sc * r1.x, sc * r1.y, sc * r2.x, sc * r2.y};
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

// coarse restriction, general P (FP32 blocks): rc[I] = sum over the P blocks (i, I) of column I (ascending) of
// P_iI^T r_i; one warp per coarse row, lanes over the entries, fixed shuffle tree
// [fine_lo, fine_hi): only the fine rows of this window contribute (several ranks: each rank sums over its own rows of the
// finer level and the partial results are all-reduced)
template <bool H16>
__global__ void __launch_bounds__(kT) k_ml_restrict(int64_t nc, const int32_t *__restrict__ rrow,
                                                    const int32_t *__restrict__ rblk, const int32_t *__restrict__ rfine,
                                                    const void *__restrict__ pval, const double *__restrict__ r,
                                                    double *__restrict__ rc, int32_t fine_lo, int32_t fine_hi)
{
    const int lane = threadIdx.x & 31;
    const int64_t I = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (I >= nc) return;   // whole warps leave together
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (int32_t e = __ldg(rrow + I) + lane; e < __ldg(rrow + I + 1); e += 32) {
        const int32_t fi = __ldg(rfine + e);
        if (fi < fine_lo || fi >= fine_hi) continue;
        const double *rv = r + (int64_t)fi * 6;
        float pp[36];
        double sc;
        tblock_load<H16>(pval, (int64_t)__ldg(rblk + e), pp, sc);
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const double ri = sc * rv[i];
#pragma unroll
            for (int j = 0; j < 6; ++j) acc[j] = fma((double)pp[6 * i + j], ri, acc[j]);
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

// the same for coarse levels whose rows are long (smoothed prolongators between small levels: hundreds of P blocks per
// coarse row, a few hundred rows): one CTA of 128 threads per coarse row, one entry per thread and pass, fixed-order
// CTA reduction -- the warp-per-row kernel above is latency-bound there (7 dependent passes per warp, 216 warps)
template <bool H16>
__global__ void __launch_bounds__(128) k_ml_restrict_wide(int64_t nc, const int32_t *__restrict__ rrow,
                                                          const int32_t *__restrict__ rblk, const int32_t *__restrict__ rfine,
                                                          const void *__restrict__ pval, const double *__restrict__ r,
                                                          double *__restrict__ rc, int32_t fine_lo, int32_t fine_hi)
{
    __shared__ double red[4][6];
    const int64_t I = blockIdx.x;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (int32_t e = __ldg(rrow + I) + threadIdx.x; e < __ldg(rrow + I + 1); e += 128) {
        const int32_t fi = __ldg(rfine + e);
        if (fi < fine_lo || fi >= fine_hi) continue;
        const double *rv = r + (int64_t)fi * 6;
        float pp[36];
        double sc;
        tblock_load<H16>(pval, (int64_t)__ldg(rblk + e), pp, sc);
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const double ri = sc * rv[i];
#pragma unroll
            for (int j = 0; j < 6; ++j) acc[j] = fma((double)pp[6 * i + j], ri, acc[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < 6; ++j)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_down_sync(0xffffffffu, acc[j], o);
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int j = 0; j < 6; ++j) red[threadIdx.x >> 5][j] = acc[j];
    __syncthreads();
    if (threadIdx.x < 6) rc[I * 6 + threadIdx.x] = (red[0][threadIdx.x] + red[1][threadIdx.x]) + (red[2][threadIdx.x] + red[3][threadIdx.x]);
}

// coarse restriction, TENTATIVE P (one rigid block per row, generated from the offsets: 24 bytes per member instead
// of a 6x6 block): rc[I] = sum over the members i of aggregate I (ascending) of B(off_i)^T r_i; 8 lanes per aggregate
__global__ void __launch_bounds__(kT) k_ml_restrict_tent(int64_t nc, const int32_t *__restrict__ rrow,
                                                         const int32_t *__restrict__ rfine, const double *__restrict__ off,
                                                         const double *__restrict__ r, double *__restrict__ rc,
                                                         int32_t fine_lo, int32_t fine_hi)
{
    const int lane = threadIdx.x & 7;
    const int64_t I = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    if (I < nc) {
        for (int32_t e = __ldg(rrow + I) + lane; e < __ldg(rrow + I + 1); e += 8) {
            const int64_t i = __ldg(rfine + e);
            if (i < fine_lo || i >= fine_hi) continue;
            double w[6], t[6];
#pragma unroll
            for (int x = 0; x < 6; ++x) w[x] = r[i * 6 + x];
            const double d[3] = {__ldg(off + 3 * i), __ldg(off + 3 * i + 1), __ldg(off + 3 * i + 2)};
            rigid_apply_t(d, w, t);
#pragma unroll
            for (int x = 0; x < 6; ++x) acc[x] += t[x];
        }
    }
#pragma unroll
    for (int j = 0; j < 6; ++j)
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) acc[j] += __shfl_down_sync(0xffffffffu, acc[j], o, 8);
    if (I < nc && lane == 0)
#pragma unroll
        for (int j = 0; j < 6; ++j) rc[I * 6 + j] = acc[j];
}

// coarse level on the way down: z_i = D_i^-1 r_i + sum over row i of P of P_iJ zc_J; thread = (row, scalar); FP32 D^-1, P
template <bool H16>
__global__ void __launch_bounds__(kT) k_ml_smooth_prolong(int64_t n, const float *__restrict__ dinv,
                                                          const int32_t *__restrict__ prow, const int32_t *__restrict__ pcol,
                                                          const void *__restrict__ pval, const double *__restrict__ r,
                                                          const double *__restrict__ zc, double *__restrict__ z, int64_t row0)
{
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t0 / 6 >= n) return;
    const int64_t t = t0 + 6 * row0, v = t / 6;   // rows [row0, row0 + n)
    const int i = (int)(t - v * 6);
    const float2 *D = reinterpret_cast<const float2 *>(dinv + v * 36 + 6 * i);
    const double *rv = r + v * 6;
    const float2 d0 = __ldg(D), d1 = __ldg(D + 1), d2 = __ldg(D + 2);
    double s = (double)d0.x * rv[0];
    s = fma((double)d0.y, rv[1], s); s = fma((double)d1.x, rv[2], s); s = fma((double)d1.y, rv[3], s);
    s = fma((double)d2.x, rv[4], s); s = fma((double)d2.y, rv[5], s);
    for (int32_t b = __ldg(prow + v); b < __ldg(prow + v + 1); ++b) {
        const double *c = zc + (int64_t)__ldg(pcol + b) * 6;
        float pr[6];
        double sc;
        tblock_load_row<H16>(pval, b, i, pr, sc);
        double a = (double)pr[0] * c[0];
        a = fma((double)pr[1], c[1], a); a = fma((double)pr[2], c[2], a);
        a = fma((double)pr[3], c[3], a); a = fma((double)pr[4], c[4], a); a = fma((double)pr[5], c[5], a);
        s = fma(sc, a, s);
    }
    z[t] = s;
}

// the same with one WARP per row, lanes over the row's P blocks (rows of smoothed prolongators between small levels hold
// tens of blocks: the thread-per-scalar kernel above walks them one dependent gather after the other)
template <bool H16>
__global__ void __launch_bounds__(kT) k_ml_smooth_prolong_warp(int64_t n, const float *__restrict__ dinv,
                                                               const int32_t *__restrict__ prow, const int32_t *__restrict__ pcol,
                                                               const void *__restrict__ pval, const double *__restrict__ r,
                                                               const double *__restrict__ zc, double *__restrict__ z, int64_t row0)
{
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n) return;   // whole warps leave together
    const int64_t v = w + row0;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (int32_t b = __ldg(prow + v) + lane; b < __ldg(prow + v + 1); b += 32) {
        const double *c = zc + (int64_t)__ldg(pcol + b) * 6;
        float pp[36];
        double sc;
        tblock_load<H16>(pval, b, pp, sc);
        const double cc[6] = {sc * c[0], sc * c[1], sc * c[2], sc * c[3], sc * c[4], sc * c[5]};
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = 0; j < 6; ++j) acc[i] = fma((double)pp[6 * i + j], cc[j], acc[i]);
    }
    if (lane < 6) {   // D^-1 r: lane i adds row i
        const float2 *D = reinterpret_cast<const float2 *>(dinv + v * 36 + 6 * lane);
        const double *rv = r + v * 6;
        const float2 d0 = __ldg(D), d1 = __ldg(D + 1), d2 = __ldg(D + 2);
        double s = (double)d0.x * rv[0];
        s = fma((double)d0.y, rv[1], s); s = fma((double)d1.x, rv[2], s); s = fma((double)d1.y, rv[3], s);
        s = fma((double)d2.x, rv[4], s); s = fma((double)d2.y, rv[5], s);
#pragma unroll
        for (int i = 0; i < 6; ++i) if (i == lane) acc[i] += s;
    }
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_down_sync(0xffffffffu, acc[i], o);
    if (lane == 0)
#pragma unroll
        for (int i = 0; i < 6; ++i) z[v * 6 + i] = acc[i];
}

// the same for a TENTATIVE P: z_i = D_i^-1 r_i + B(off_i) zc[agg i]
__global__ void __launch_bounds__(kT) k_ml_smooth_prolong_tent(int64_t n, const float *__restrict__ dinv,
                                                               const int32_t *__restrict__ agg, const double *__restrict__ off,
                                                               const double *__restrict__ r, const double *__restrict__ zc,
                                                               double *__restrict__ z, int64_t row0)
{
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t0 / 6 >= n) return;
    const int64_t t = t0 + 6 * row0, v = t / 6;   // rows [row0, row0 + n)
    const int i = (int)(t - v * 6);
    const float2 *D = reinterpret_cast<const float2 *>(dinv + v * 36 + 6 * i);
    const double *rv = r + v * 6;
    const float2 d0 = __ldg(D), d1 = __ldg(D + 1), d2 = __ldg(D + 2);
    double s = (double)d0.x * rv[0];
    s = fma((double)d0.y, rv[1], s); s = fma((double)d1.x, rv[2], s); s = fma((double)d1.y, rv[3], s);
    s = fma((double)d2.x, rv[4], s); s = fma((double)d2.y, rv[5], s);
    const double *c = zc + (int64_t)__ldg(agg + v) * 6;
    const double d[3] = {__ldg(off + 3 * v), __ldg(off + 3 * v + 1), __ldg(off + 3 * v + 2)};
    double cc[6], tt[6];
#pragma unroll
    for (int x = 0; x < 6; ++x) cc[x] = c[x];
    rigid_apply(d, cc, tt);
    z[t] = s + tt[i];
}

// ---- launchers ----------------------------------------------------------------------------------------------
void launch_ml_make_t0(cudaStream_t st, int64_t nb, const double *lfac, const double *off, float *t0)
{
    if (nb > 0) k_ml_make_t0<<<blocks_for(nb * 36), kT, 0, st>>>(nb, lfac, off, t0);
}
void launch_ml_diag_inv(cudaStream_t st, int64_t n, const double *sinv, double *dinv, float *dinv32)
{
    if (n > 0) k_ml_diag_inv<<<blocks_for(n * 36), kT, 0, st>>>(n, sinv, dinv, dinv32);
}
void launch_ml_shift_i32(cudaStream_t st, int64_t n, const int32_t *src, int32_t delta, int32_t *dst)
{
    if (n > 0) k_ml_shift_i32<<<blocks_for(n), kT, 0, st>>>(n, src, delta, dst);
}
void launch_ml_gather3(cudaStream_t st, int64_t n, const double *src, const int64_t *idx, double *dst)
{
    if (n > 0) k_ml_gather3<<<blocks_for(3 * n), kT, 0, st>>>(n, src, idx, dst);
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
void launch_ml_galerkin0(cudaStream_t st, int64_t n_coarse_blocks, const int32_t *gal_ptr, const int32_t *gal_src,
                         const int32_t *frow, const int32_t *fcol, const double *fvals, const double *off, double *cvals)
{
    if (n_coarse_blocks > 0)
        k_ml_galerkin0<<<blocks_for(n_coarse_blocks * 8, 128), 128, 0, st>>>(n_coarse_blocks, gal_ptr, gal_src, frow, fcol, fvals, off, cvals);
}
// a: n x n (n = 6 m) overwritten by its inverse; colk: (n + 24) * 24 doubles of scratch (the last 576: the pivot inverse);
// 3 launches per 24 pivot columns
void launch_ml_dense_inverse(cudaStream_t st, int n, double *a, double *colk, int *flag)
{
    if (n <= 0) return;
    const dim3 grid((unsigned)((n + 63) / 64), (unsigned)((n + 63) / 64));
    double *pinv = colk + (size_t)n * kGjW;
    for (int k0 = 0; k0 < n; k0 += kGjW) {
        const int bw = n - k0 < kGjW ? n - k0 : kGjW;
        k_ml_gj_invert<<<1, 32, 0, st>>>(n, k0, bw, a, pinv, flag);
        k_ml_gj_rows<<<blocks_for(n, 128), 128, 0, st>>>(n, k0, bw, a, pinv, colk);
        k_ml_gj_update<<<grid, 256, 0, st>>>(n, k0, bw, a, colk);
    }
}
void launch_ml_dense_to_f32(cudaStream_t st, int n, const double *a, float *out)
{
    if (n > 0) k_ml_dense_to_f32<<<blocks_for((int64_t)n * n, 256), 256, 0, st>>>(n, a, out);
}
void launch_ml_dense_matvec(cudaStream_t st, int n, const float *ainv, const double *r, double *z)
{
    if (n > 0) k_ml_dense_matvec<<<blocks_for((int64_t)n * 32), kT, 0, st>>>(n, ainv, r, z);
}
void launch_ml_to_f32(cudaStream_t st, int64_t n, const double *a, float *out)
{
    if (n > 0) k_ml_to_f32<<<blocks_for(n), kT, 0, st>>>(n, a, out);
}
void launch_ml_pack16(cudaStream_t st, int64_t nblk, const float *src, void *dst)
{
    if (nblk > 0) k_ml_pack16<float><<<blocks_for(nblk), kT, 0, st>>>(nblk, src, static_cast<unsigned char *>(dst));
}
void launch_ml_pack16(cudaStream_t st, int64_t nblk, const double *src, void *dst)
{
    if (nblk > 0) k_ml_pack16<double><<<blocks_for(nblk), kT, 0, st>>>(nblk, src, static_cast<unsigned char *>(dst));
}
void launch_ml_restrict0(cudaStream_t st, int64_t n_agg, int group, const int32_t *agg_ptr, const int32_t *agg_rows,
                         const void *t0, bool half, const double *r, double *rc)
{
    if (n_agg <= 0) return;
    if (group == 8) {
        if (half) k_ml_restrict0<8, true><<<blocks_for(n_agg * 8), kT, 0, st>>>(n_agg, agg_ptr, agg_rows, t0, r, rc);
        else k_ml_restrict0<8, false><<<blocks_for(n_agg * 8), kT, 0, st>>>(n_agg, agg_ptr, agg_rows, t0, r, rc);
    } else {
        if (half) k_ml_restrict0<32, true><<<blocks_for(n_agg * 32), kT, 0, st>>>(n_agg, agg_ptr, agg_rows, t0, r, rc);
        else k_ml_restrict0<32, false><<<blocks_for(n_agg * 32), kT, 0, st>>>(n_agg, agg_ptr, agg_rows, t0, r, rc);
    }
}
void launch_ml_restrict(cudaStream_t st, int64_t nc, int64_t nnzp, const int32_t *rrow, const int32_t *rblk, const int32_t *rfine,
                        const void *pval, bool half, const double *r, double *rc, int64_t fine_lo, int64_t fine_hi)
{
    if (nc <= 0) return;
    const int32_t lo = (int32_t)fine_lo, hi = (int32_t)fine_hi;
    if (nnzp > 48 * nc && nc <= 65535) {   // long rows on a small level
        if (half) k_ml_restrict_wide<true><<<(unsigned)nc, 128, 0, st>>>(nc, rrow, rblk, rfine, pval, r, rc, lo, hi);
        else k_ml_restrict_wide<false><<<(unsigned)nc, 128, 0, st>>>(nc, rrow, rblk, rfine, pval, r, rc, lo, hi);
    } else {
        if (half) k_ml_restrict<true><<<blocks_for(nc * 32), kT, 0, st>>>(nc, rrow, rblk, rfine, pval, r, rc, lo, hi);
        else k_ml_restrict<false><<<blocks_for(nc * 32), kT, 0, st>>>(nc, rrow, rblk, rfine, pval, r, rc, lo, hi);
    }
}
void launch_ml_restrict_tent(cudaStream_t st, int64_t nc, const int32_t *rrow, const int32_t *rfine, const double *off,
                             const double *r, double *rc, int64_t fine_lo, int64_t fine_hi)
{
    if (nc > 0) k_ml_restrict_tent<<<blocks_for(nc * 8), kT, 0, st>>>(nc, rrow, rfine, off, r, rc, (int32_t)fine_lo, (int32_t)fine_hi);
}
void launch_ml_smooth_prolong(cudaStream_t st, int64_t row0, int64_t n, int64_t n_level, int64_t nnzp, const float *dinv,
                              const int32_t *prow, const int32_t *pcol, const void *pval, bool half, const double *r, const double *zc,
                              double *z)
{
    if (n <= 0) return;
    if (nnzp > 6 * n_level) {   // rows of many blocks: a warp per row
        if (half) k_ml_smooth_prolong_warp<true><<<blocks_for(n * 32), kT, 0, st>>>(n, dinv, prow, pcol, pval, r, zc, z, row0);
        else k_ml_smooth_prolong_warp<false><<<blocks_for(n * 32), kT, 0, st>>>(n, dinv, prow, pcol, pval, r, zc, z, row0);
    } else {
        if (half) k_ml_smooth_prolong<true><<<blocks_for(n * 6), kT, 0, st>>>(n, dinv, prow, pcol, pval, r, zc, z, row0);
        else k_ml_smooth_prolong<false><<<blocks_for(n * 6), kT, 0, st>>>(n, dinv, prow, pcol, pval, r, zc, z, row0);
    }
}
void launch_ml_smooth_prolong_tent(cudaStream_t st, int64_t row0, int64_t n, const float *dinv, const int32_t *agg, const double *off,
                                   const double *r, const double *zc, double *z)
{
    if (n > 0) k_ml_smooth_prolong_tent<<<blocks_for(n * 6), kT, 0, st>>>(n, dinv, agg, off, r, zc, z, row0);
}

}}  // namespace mbfea::dev
