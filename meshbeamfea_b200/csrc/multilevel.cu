// Device half of the multilevel (rigid-body aggregate) preconditioner, SURVEY 8(f) rank 4 -- DRAFT.
// These kernels consume the arrays of csrc/hierarchy.cpp (HierLevel) and csrc/dense.cpp.  They compile into
// libmbfea.so but NOTHING calls them yet: they were written at the end of round 1 without GPU time left to
// validate them, so mbfea_solve does not use them and no parity claim covers them.  Wiring them into the PCG
// driver (z~ = r~ + L^T B z_1 replaces z~ = r~; k_cg_pupdate takes z~ where it takes r~) and testing them
// against tests/tools/proto_hierarchy_pcg.py is the first task of the next round.
//
// Notation.  Level l rows carry 6 unknowns.  agg[v] = aggregate (row of level l+1) of row v, off[3v..] = d_v =
// position of v minus the centroid of its aggregate, and the 6x6 prolongation block is
//     B(d) = [[I, -[d]x], [0, I]]   (u = t + w x d, theta = w),
// so  B(d) (t, w) = (t + w x d, w)   and   B(d)^T (f, m) = (f, m + d x f).
// The fine level lives in the block-Jacobi-scaled variables of the solver (y = L^T x): its transfer operators
// are L_v^T B(d_v), with L_v the row's Cholesky factor (lfac, full 6x6 storage, zeros above the diagonal).
// All sums run in a fixed order (member lists and contributor lists ascend): results are reproducible.
#if !defined(MBFEA_HOST_EMULATION)   // tests/multilevel_emul.cpp compiles the kernel bodies for the CPU
#include <cuda_runtime.h>
#endif

#include <cstdint>

#if !defined(MBFEA_HOST_EMULATION)
#include "kernels.cuh"
#endif
#include "multilevel_math.hpp"

namespace mbfea { namespace dev {

namespace {

constexpr int kMlThreads = 256;

}  // namespace

// ---------------------------------------------------------------------------
// Galerkin assembly of one coarse level: coarse block k = sum over its contributor list (ascending finer
// blocks q = gal_src[..]) of B(d_a)^T A_q B(d_b), a = frow[q], b = fcol[q].  One thread per (coarse block,
// entry): 36 consecutive threads share a block.  fvals: finer level's blocks, 36 doubles each, row-major.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kMlThreads) k_ml_galerkin(int64_t n_coarse_blocks, const int32_t *__restrict__ gal_ptr,
                                                            const int32_t *__restrict__ gal_src,
                                                            const int32_t *__restrict__ frow,
                                                            const int32_t *__restrict__ fcol,
                                                            const double *__restrict__ fvals,
                                                            const double *__restrict__ off, double *__restrict__ cvals)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t k = t / 36;
    if (k >= n_coarse_blocks) return;
    const int e = (int)(t - k * 36), i = e / 6, m = e - 6 * i;
    double acc = 0.0;
    for (int32_t s = __ldg(gal_ptr + k); s < __ldg(gal_ptr + k + 1); ++s) {
        const int32_t q = __ldg(gal_src + s);
        const double *A = fvals + (int64_t)q * 36;
        const double *da = off + 3 * (int64_t)__ldg(frow + q), *db = off + 3 * (int64_t)__ldg(fcol + q);
        double Ab[36];
#pragma unroll
        for (int x = 0; x < 36; ++x) Ab[x] = __ldg(A + x);
        const double v = galerkin_entry(da, db, Ab, i, m);
        acc += v;
    }
    cvals[k * 36 + e] = acc;
}

// ---------------------------------------------------------------------------
// Restriction: rc[I] = sum over the members v of aggregate I (ascending) of B(d_v)^T w_v, one warp per aggregate.
// SCALED (fine level): w_v = L_v r~_v, the unscaled residual.  Lanes take members round-robin, then a
// shuffle tree in a fixed pattern.
// ---------------------------------------------------------------------------
template <bool SCALED>
__global__ void __launch_bounds__(kMlThreads) k_ml_restrict(int64_t n_agg, const int32_t *__restrict__ agg_ptr,
                                                            const int32_t *__restrict__ agg_rows,
                                                            const double *__restrict__ off,
                                                            const double *__restrict__ lfac,
                                                            const double *__restrict__ r, double *__restrict__ rc)
{
    const int lane = threadIdx.x & 31;
    const int64_t I = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (I >= n_agg) return;   // whole warps leave together
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (int32_t q = __ldg(agg_ptr + I) + lane; q < __ldg(agg_ptr + I + 1); q += 32) {
        const int64_t v = __ldg(agg_rows + q);
        double w[6], t[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) w[i] = r[v * 6 + i];
        if (SCALED) {
            const double *L = lfac + v * 36;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                double s = 0.0;
#pragma unroll
                for (int j = 0; j <= i; ++j) s = fma(__ldg(L + 6 * i + j), w[j], s);
                t[i] = s;
            }
#pragma unroll
            for (int i = 0; i < 6; ++i) w[i] = t[i];
        }
        rigid_apply_t(off + 3 * v, w, t);
#pragma unroll
        for (int i = 0; i < 6; ++i) acc[i] += t[i];
    }
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_down_sync(0xffffffffu, acc[i], o);
    if (lane == 0)
#pragma unroll
        for (int i = 0; i < 6; ++i) rc[I * 6 + i] = acc[i];
}

// z = D^-1 r on an intermediate level, D^-1 = S^T S with S = L^-1 from k_block_chol (one thread per row)
__global__ void __launch_bounds__(kMlThreads) k_ml_smooth(int64_t n, const double *__restrict__ sinv,
                                                          const double *__restrict__ r, double *__restrict__ z)
{
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const double *S = sinv + v * 36;
    double t[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j <= i; ++j) s = fma(__ldg(S + 6 * i + j), r[v * 6 + j], s);
        t[i] = s;
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double s = 0.0;
#pragma unroll
        for (int j = i; j < 6; ++j) s = fma(__ldg(S + 6 * j + i), t[j], s);   // (S^T t)_i
        z[v * 6 + i] = s;
    }
}

// coarsest level: z = Ainv r, Ainv dense n x n row-major (the host's explicit inverse); one warp per row
__global__ void __launch_bounds__(kMlThreads) k_ml_dense_matvec(int64_t n, const double *__restrict__ Ainv,
                                                                const double *__restrict__ r, double *__restrict__ z)
{
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    double s = 0.0;
    for (int64_t j = lane; j < n; j += 32) s = fma(__ldg(Ainv + row * n + j), r[j], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) z[row] = s;
}

// prolongation on the way down: z[v] += B(d_v) zc[agg v]   (one thread per row of an intermediate level)
__global__ void __launch_bounds__(kMlThreads) k_ml_prolong(int64_t n, const int32_t *__restrict__ agg,
                                                           const double *__restrict__ off,
                                                           const double *__restrict__ zc, double *__restrict__ z)
{
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    double c[6], t[6];
    const int64_t I = __ldg(agg + v);
#pragma unroll
    for (int i = 0; i < 6; ++i) c[i] = zc[I * 6 + i];
    rigid_apply(off + 3 * v, c, t);
#pragma unroll
    for (int i = 0; i < 6; ++i) z[v * 6 + i] += t[i];
}

// fine level, scaled variables: z~_v = r~_v + L_v^T B(d_v) z1[agg v];  part[blockIdx] = sum of r~ . z~ over the
// CTA's rows (grid-stride, one thread per row) -- the rho of the preconditioned CG
__global__ void __launch_bounds__(kMlThreads) k_ml_finish(int64_t nb, const int32_t *__restrict__ agg,
                                                          const double *__restrict__ off,
                                                          const double *__restrict__ lfac,
                                                          const double *__restrict__ rt, const double *__restrict__ z1,
                                                          double *__restrict__ zt, double *__restrict__ part)
{
    __shared__ double red[kMlThreads / 32];
    double dot = 0.0;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nb; v += (int64_t)gridDim.x * blockDim.x) {
        double c[6], t[6];
        const int64_t I = __ldg(agg + v);
#pragma unroll
        for (int i = 0; i < 6; ++i) c[i] = z1[I * 6 + i];
        rigid_apply(off + 3 * v, c, t);
        const double *L = lfac + v * 36;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            double s = rt[v * 6 + i];
            const double ri = s;
#pragma unroll
            for (int j = i; j < 6; ++j) s = fma(__ldg(L + 6 * j + i), t[j], s);   // (L^T t)_i
            zt[v * 6 + i] = s;
            dot = fma(ri, s, dot);
        }
    }
    // fixed-order CTA sum
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_down_sync(0xffffffffu, dot, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < kMlThreads / 32; ++w) s += red[w];
        part[blockIdx.x] = s;
    }
}

#if !defined(MBFEA_HOST_EMULATION)
// ---- launchers (no caller yet, see the header comment) -------------------------------------------------------
void launch_ml_galerkin(cudaStream_t st, int64_t n_coarse_blocks, const int32_t *gal_ptr, const int32_t *gal_src,
                        const int32_t *frow, const int32_t *fcol, const double *fvals, const double *off, double *cvals)
{
    if (n_coarse_blocks <= 0) return;
    const int64_t threads = n_coarse_blocks * 36;
    k_ml_galerkin<<<(unsigned)((threads + kMlThreads - 1) / kMlThreads), kMlThreads, 0, st>>>(
        n_coarse_blocks, gal_ptr, gal_src, frow, fcol, fvals, off, cvals);
}

void launch_ml_restrict(cudaStream_t st, int64_t n_agg, const int32_t *agg_ptr, const int32_t *agg_rows, const double *off,
                        const double *lfac, const double *r, double *rc)
{
    if (n_agg <= 0) return;
    const unsigned grid = (unsigned)((n_agg * 32 + kMlThreads - 1) / kMlThreads);
    if (lfac != nullptr) k_ml_restrict<true><<<grid, kMlThreads, 0, st>>>(n_agg, agg_ptr, agg_rows, off, lfac, r, rc);
    else k_ml_restrict<false><<<grid, kMlThreads, 0, st>>>(n_agg, agg_ptr, agg_rows, off, nullptr, r, rc);
}

void launch_ml_smooth(cudaStream_t st, int64_t n, const double *sinv, const double *r, double *z)
{
    if (n <= 0) return;
    k_ml_smooth<<<(unsigned)((n + kMlThreads - 1) / kMlThreads), kMlThreads, 0, st>>>(n, sinv, r, z);
}

void launch_ml_dense_matvec(cudaStream_t st, int64_t n, const double *Ainv, const double *r, double *z)
{
    if (n <= 0) return;
    k_ml_dense_matvec<<<(unsigned)((n * 32 + kMlThreads - 1) / kMlThreads), kMlThreads, 0, st>>>(n, Ainv, r, z);
}

void launch_ml_prolong(cudaStream_t st, int64_t n, const int32_t *agg, const double *off, const double *zc, double *z)
{
    if (n <= 0) return;
    k_ml_prolong<<<(unsigned)((n + kMlThreads - 1) / kMlThreads), kMlThreads, 0, st>>>(n, agg, off, zc, z);
}

void launch_ml_finish(cudaStream_t st, int grid, int64_t nb, const int32_t *agg, const double *off, const double *lfac,
                      const double *rt, const double *z1, double *zt, double *part)
{
    if (nb <= 0 || grid <= 0) return;
    k_ml_finish<<<(unsigned)grid, kMlThreads, 0, st>>>(nb, agg, off, lfac, rt, z1, zt, part);
}

#endif   // !MBFEA_HOST_EMULATION

}}  // namespace mbfea::dev
