// Rigid-body transfer blocks of the multilevel preconditioner (see multilevel.cu), usable from host code too so
// that the arithmetic is unit-tested on the CPU (tests/test_host_logic.py::test_multilevel_transfer_math).
#pragma once
#if defined(__CUDACC__)
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

}}  // namespace mbfea::dev
