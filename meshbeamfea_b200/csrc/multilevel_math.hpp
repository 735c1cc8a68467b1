// Rigid-body transfer blocks of the multilevel preconditioner (see multilevel.cu), usable from host code too so
// that the arithmetic is unit-tested on the CPU (tests/test_host_logic.py::test_multilevel_transfer_math).
#pragma once
#if defined(__CUDACC__)
#include <cuda_fp16.h>

#include <cstdint>

#include "kernels.cuh"
#define MBFEA_HD __host__ __device__ __forceinline__
#else
#define MBFEA_HD inline
#endif

namespace mbfea { namespace dev {

// B(d) = [[I, -[d]x], [0, I]]:  out = B(d) c  (c = translation t, rotation w:  u = t + w x d, theta = w)
MBFEA_HD void rigid_apply(const double *d, const double *c, double *out)
{
    const double tx = c[0], ty = c[1], tz = c[2], wx = c[3], wy = c[4], wz = c[5];
    out[0] = tx + (wy * d[2] - wz * d[1]);
    out[1] = ty + (wz * d[0] - wx * d[2]);
    out[2] = tz + (wx * d[1] - wy * d[0]);
    out[3] = wx; out[4] = wy; out[5] = wz;
}
// out = B(d)^T r  (r = force f, moment m:  f, m + d x f)
MBFEA_HD void rigid_apply_t(const double *d, const double *r, double *out)
{
    const double fx = r[0], fy = r[1], fz = r[2];
    out[0] = fx; out[1] = fy; out[2] = fz;
    out[3] = r[3] + (d[1] * fz - d[2] * fy);
    out[4] = r[4] + (d[2] * fx - d[0] * fz);
    out[5] = r[5] + (d[0] * fy - d[1] * fx);
}
// entry (j, i) of B(d): identity plus -[d]x = [[0, dz, -dy], [-dz, 0, dx], [dy, -dx, 0]] in the upper right block
MBFEA_HD double rigid_entry(const double *d, int j, int i)
{
    if (j == i) return 1.0;
    if (j >= 3 || i < 3) return 0.0;
    const int c = i - 3;
    if (j == 0) return c == 1 ? d[2] : (c == 2 ? -d[1] : 0.0);
    if (j == 1) return c == 0 ? -d[2] : (c == 2 ? d[0] : 0.0);
    return c == 0 ? d[1] : (c == 1 ? -d[0] : 0.0);
}
// (B(da)^T A B(db))[i][m] for a 6x6 block A (row-major)
MBFEA_HD double galerkin_entry(const double *da, const double *db, const double *A, int i, int m)
{
    double v = 0.0;
    for (int j = 0; j < 6; ++j) {
        const double ba = rigid_entry(da, j, i);
        if (ba == 0.0) continue;
        double row = 0.0;
        for (int l = 0; l < 6; ++l) {
            const double bb = rigid_entry(db, l, m);
            if (bb != 0.0) row += A[6 * j + l] * bb;
        }
        v += ba * row;
    }
    return v;
}

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------
// Transfer blocks of the multilevel apply (fine T_v, coarse P_iJ).  The apply is bandwidth-bound and needs no more than
// ~3 digits of the transfer entries (the preconditioner only has to stay symmetric: restriction and prolongation read
// the SAME rounded values), so a block is stored as 36 16-bit floats normalised by the block's largest magnitude plus
// that magnitude as a float: 80 bytes, 16-byte aligned, instead of 144 (FP32).  The per-block scale makes the format
// independent of the units of the mesh (FP16 alone would overflow on L_v^T entries above 65504).
// H16 = false: plain FP32 blocks of 144 bytes -- the DEFAULT: the 16-bit format is an opt-in experiment (MBFEA_ML_HALF=1),
// fine on stiff 3-D lattices, not accurate enough for thin shells (see MlPrec::half16).
// ---------------------------------------------------------------------------

__device__ __forceinline__ float2 tb_half2(uint32_t u)
{
    return __half22float2(*reinterpret_cast<const __half2 *>(&u));
}
// whole block -> pp[36] (row-major), scale
template <bool H16>
__device__ __forceinline__ void tblock_load(const void *base, int64_t b, float (&pp)[36], double &scale)
{
    if (H16) {
        const uint4 *q = reinterpret_cast<const uint4 *>(static_cast<const unsigned char *>(base) + b * kTBlock16Bytes);
        uint4 w[5];
#pragma unroll
        for (int x = 0; x < 5; ++x) w[x] = __ldg(q + x);
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            const float2 a = tb_half2(w[x].x), c = tb_half2(w[x].y), d = tb_half2(w[x].z), e = tb_half2(w[x].w);
            pp[8 * x] = a.x; pp[8 * x + 1] = a.y; pp[8 * x + 2] = c.x; pp[8 * x + 3] = c.y;
            pp[8 * x + 4] = d.x; pp[8 * x + 5] = d.y; pp[8 * x + 6] = e.x; pp[8 * x + 7] = e.y;
        }
        const float2 a = tb_half2(w[4].x), c = tb_half2(w[4].y);
        pp[32] = a.x; pp[33] = a.y; pp[34] = c.x; pp[35] = c.y;
        scale = (double)__uint_as_float(w[4].z);
    } else {
        const float4 *q = reinterpret_cast<const float4 *>(static_cast<const float *>(base) + b * 36);
#pragma unroll
        for (int x = 0; x < 9; ++x) { const float4 f = __ldg(q + x); pp[4 * x] = f.x; pp[4 * x + 1] = f.y; pp[4 * x + 2] = f.z; pp[4 * x + 3] = f.w; }
        scale = 1.0;
    }
}
// row i of a block -> row[6], scale
template <bool H16>
__device__ __forceinline__ void tblock_load_row(const void *base, int64_t b, int i, float (&row)[6], double &scale)
{
    if (H16) {
        const unsigned char *p = static_cast<const unsigned char *>(base) + b * kTBlock16Bytes;
        const uint32_t *q = reinterpret_cast<const uint32_t *>(p + 12 * i);
        const float2 a = tb_half2(__ldg(q)), c = tb_half2(__ldg(q + 1)), d = tb_half2(__ldg(q + 2));
        row[0] = a.x; row[1] = a.y; row[2] = c.x; row[3] = c.y; row[4] = d.x; row[5] = d.y;
        scale = (double)__ldg(reinterpret_cast<const float *>(p + 72));
    } else {
        const float2 *q = reinterpret_cast<const float2 *>(static_cast<const float *>(base) + b * 36 + 6 * i);
        const float2 a = __ldg(q), c = __ldg(q + 1), d = __ldg(q + 2);
        row[0] = a.x; row[1] = a.y; row[2] = c.x; row[3] = c.y; row[4] = d.x; row[5] = d.y;
        scale = 1.0;
    }
}
#endif

}}  // namespace mbfea::dev
