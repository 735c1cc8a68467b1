// Register-resident Euler-Bernoulli frame element: local frame, the ten
// distinct stiffness coefficients and one ROW of one 6x6 sub-block of
// K_e = A^T K_l A.  Replaces threed-beam-fea's calcKelem / calcAelem and the
// two dense 12x12 products behind fea::solve [REF src/beamsimulator.cpp:65-66].
//
// The 12x12 A and K_l are never formed.  K_l's four 6x6 sub-blocks all have the
// sparsity
//        | d0  .   .   .   .   .  |
//        | .   d1  .   .   .   c15|
//        | .   .   d2  .   c24 .  |
//        | .   .   .   d3  .   .  |
//        | .   .   c42 .   d4  .  |
//        | .   c51 .   .   .   d5 |
// and a row of Q^T S Q (Q = blockdiag(R,R), rows of R = nx,ny,nz) needs 5
// products for w = S^T q and 18 for the row.  Every operation is an explicit
// IEEE round-to-nearest mul/add/div/sqrt in the oracle's order (no FMA
// contraction), so assembled K entries are bit-identical to oracle/mbfea_oracle.c
// up to the sign of zero.
#pragma once
#include <cstdint>

namespace mbfea { namespace dev {

struct __align__(16) ElemRec {   // 64 B: one aligned record per beam
    double n[3];                 // edge normal (local y) as given
    double props[4];             // EA, EIz, EIy, GJ (float-quirk values widened)
    int32_t a, b;                // end vertices
};
static_assert(sizeof(ElemRec) == 64, "ElemRec must be 64 bytes");

struct __align__(32) Node4 { double x, y, z, w; };

struct Frame { double nx[3], ny[3], nz[3]; double L; };

// K1 output: the local frame and the ten stiffness coefficients of one beam,
// computed once per beam and consumed by the assembly slices (160 B, 32-B aligned).
struct __align__(32) ElemGeom { double v[20]; };
static_assert(sizeof(ElemGeom) == 160, "ElemGeom must be 160 bytes");

struct Stiff {  // the ten coefficients of K_l
    double ea, gj, z12, z6, z1x4, z1x2, y12, y6, y1x4, y1x2;
};

#define MUL(a, b) __dmul_rn((a), (b))
#define ADD(a, b) __dadd_rn((a), (b))
#define SUB(a, b) __dadd_rn((a), -(b))
#define DIV(a, b) __ddiv_rn((a), (b))

__device__ __forceinline__ void make_frame(const Node4 &p0, const Node4 &p1, const double nr[3], Frame &f)
{
    const double d0 = SUB(p1.x, p0.x), d1 = SUB(p1.y, p0.y), d2 = SUB(p1.z, p0.z);
    const double L = __dsqrt_rn(ADD(ADD(MUL(d0, d0), MUL(d1, d1)), MUL(d2, d2)));
    f.L = L;
    f.nx[0] = DIV(d0, L); f.nx[1] = DIV(d1, L); f.nx[2] = DIV(d2, L);
    const double nl = __dsqrt_rn(ADD(ADD(MUL(nr[0], nr[0]), MUL(nr[1], nr[1])), MUL(nr[2], nr[2])));
    const double y0 = DIV(nr[0], nl), y1 = DIV(nr[1], nl), y2 = DIV(nr[2], nl);
    // z = x cross y, divided by its squared norm (upstream convention)
    double z0 = SUB(MUL(f.nx[1], y2), MUL(f.nx[2], y1));
    double z1 = SUB(MUL(f.nx[2], y0), MUL(f.nx[0], y2));
    double z2 = SUB(MUL(f.nx[0], y1), MUL(f.nx[1], y0));
    const double dlz = ADD(ADD(MUL(z0, z0), MUL(z1, z1)), MUL(z2, z2));
    z0 = DIV(z0, dlz); z1 = DIV(z1, dlz); z2 = DIV(z2, dlz);
    f.nz[0] = z0; f.nz[1] = z1; f.nz[2] = z2;
    // y = z cross x, divided by its squared norm
    double w0 = SUB(MUL(z1, f.nx[2]), MUL(z2, f.nx[1]));
    double w1 = SUB(MUL(z2, f.nx[0]), MUL(z0, f.nx[2]));
    double w2 = SUB(MUL(z0, f.nx[1]), MUL(z1, f.nx[0]));
    const double dly = ADD(ADD(MUL(w0, w0), MUL(w1, w1)), MUL(w2, w2));
    f.ny[0] = DIV(w0, dly); f.ny[1] = DIV(w1, dly); f.ny[2] = DIV(w2, dly);
}

__device__ __forceinline__ void make_stiff(const double pr[4], double L, Stiff &s)
{
    const double EA = pr[0], EIz = pr[1], EIy = pr[2], GJ = pr[3];
    const double LL = MUL(L, L), LLL = MUL(LL, L);
    s.ea = DIV(EA, L);
    s.gj = DIV(GJ, L);
    s.z12 = DIV(MUL(12.0, EIz), LLL);
    s.z6 = DIV(MUL(6.0, EIz), LL);
    const double z1 = DIV(EIz, L);
    s.z1x4 = MUL(4.0, z1); s.z1x2 = MUL(2.0, z1);
    s.y12 = DIV(MUL(12.0, EIy), LLL);
    s.y6 = DIV(MUL(6.0, EIy), LL);
    const double y1 = DIV(EIy, L);
    s.y1x4 = MUL(4.0, y1); s.y1x2 = MUL(2.0, y1);
}

__device__ __forceinline__ double pick3(const double v[3], int i)
{
    return i == 0 ? v[0] : (i == 1 ? v[1] : v[2]);
}

// Row r (0..5) of sub-block `which` (0:00 1:01 2:10 3:11) of K_e.
__device__ __forceinline__ void kelem_row(const Frame &f, const Stiff &s, int which, int r, double out[6])
{
    const bool diag = (which == 0) || (which == 3);
    const double d0 = diag ? s.ea : -s.ea;
    const double d1 = diag ? s.z12 : -s.z12;
    const double d2 = diag ? s.y12 : -s.y12;
    const double d3 = diag ? s.gj : -s.gj;
    const double d4 = diag ? s.y1x4 : s.y1x2;
    const double d5 = diag ? s.z1x4 : s.z1x2;
    // couplings: 00:(+z6,+z6,-y6,-y6) 01:(+z6,-z6,-y6,+y6) 10:(-z6,+z6,+y6,-y6) 11:(-z6,-z6,+y6,+y6)
    const double c15 = (which <= 1) ? s.z6 : -s.z6;
    const double c51 = (which == 0 || which == 2) ? s.z6 : -s.z6;
    const double c24 = (which <= 1) ? -s.y6 : s.y6;
    const double c42 = (which == 0 || which == 2) ? -s.y6 : s.y6;
    const int r3 = r < 3 ? r : r - 3;
    const double qx = pick3(f.nx, r3), qy = pick3(f.ny, r3), qz = pick3(f.nz, r3);
    double w0, w1, w2, w3, w4, w5;
    if (r < 3) {  // q = (qx,qy,qz,0,0,0)
        w0 = MUL(qx, d0); w1 = MUL(qy, d1); w2 = MUL(qz, d2);
        w3 = 0.0;         w4 = MUL(qz, c24); w5 = MUL(qy, c15);
    } else {      // q = (0,0,0,qx,qy,qz)
        w0 = 0.0;         w1 = MUL(qz, c51); w2 = MUL(qy, c42);
        w3 = MUL(qx, d3); w4 = MUL(qy, d4);  w5 = MUL(qz, d5);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        out[c]     = ADD(ADD(MUL(w0, f.nx[c]), MUL(w1, f.ny[c])), MUL(w2, f.nz[c]));
        out[3 + c] = ADD(ADD(MUL(w3, f.nx[c]), MUL(w4, f.ny[c])), MUL(w5, f.nz[c]));
    }
}

// f_e = K_l (A u_e): ul = local displacements (12), fe = 12 end forces, in the
// oracle's summation order (ascending column of K_l).
__device__ __forceinline__ void local_forces(const Stiff &s, const double ul[12], double fe[12])
{
    fe[0]  = ADD(MUL(s.ea, ul[0]), MUL(-s.ea, ul[6]));
    fe[1]  = ADD(ADD(ADD(MUL(s.z12, ul[1]), MUL(s.z6, ul[5])), MUL(-s.z12, ul[7])), MUL(s.z6, ul[11]));
    fe[2]  = ADD(ADD(ADD(MUL(s.y12, ul[2]), MUL(-s.y6, ul[4])), MUL(-s.y12, ul[8])), MUL(-s.y6, ul[10]));
    fe[3]  = ADD(MUL(s.gj, ul[3]), MUL(-s.gj, ul[9]));
    fe[4]  = ADD(ADD(ADD(MUL(-s.y6, ul[2]), MUL(s.y1x4, ul[4])), MUL(s.y6, ul[8])), MUL(s.y1x2, ul[10]));
    fe[5]  = ADD(ADD(ADD(MUL(s.z6, ul[1]), MUL(s.z1x4, ul[5])), MUL(-s.z6, ul[7])), MUL(s.z1x2, ul[11]));
    fe[6]  = ADD(MUL(-s.ea, ul[0]), MUL(s.ea, ul[6]));
    fe[7]  = ADD(ADD(ADD(MUL(-s.z12, ul[1]), MUL(-s.z6, ul[5])), MUL(s.z12, ul[7])), MUL(-s.z6, ul[11]));
    fe[8]  = ADD(ADD(ADD(MUL(-s.y12, ul[2]), MUL(s.y6, ul[4])), MUL(s.y12, ul[8])), MUL(s.y6, ul[10]));
    fe[9]  = ADD(MUL(-s.gj, ul[3]), MUL(s.gj, ul[9]));
    fe[10] = ADD(ADD(ADD(MUL(-s.y6, ul[2]), MUL(s.y1x2, ul[4])), MUL(s.y6, ul[8])), MUL(s.y1x4, ul[10]));
    fe[11] = ADD(ADD(ADD(MUL(s.z6, ul[1]), MUL(s.z1x2, ul[5])), MUL(-s.z6, ul[7])), MUL(s.z1x4, ul[11]));
}

__device__ __forceinline__ void load_elem(const ElemRec *__restrict__ rec, const Node4 *__restrict__ node4,
                                          int64_t e, Frame &f, Stiff &s, int32_t &a, int32_t &b)
{
    const double2 *q = reinterpret_cast<const double2 *>(rec + e);
    const double2 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2), q3 = __ldg(q + 3);
    const double nr[3] = {q0.x, q0.y, q1.x};
    const double pr[4] = {q1.y, q2.x, q2.y, q3.x};
    const long long ab = __double_as_longlong(q3.y);
    a = (int32_t)(ab & 0xffffffffLL); b = (int32_t)(ab >> 32);
    const double2 *pa = reinterpret_cast<const double2 *>(node4 + a);
    const double2 *pb = reinterpret_cast<const double2 *>(node4 + b);
    const double2 a0 = __ldg(pa), a1 = __ldg(pa + 1), b0 = __ldg(pb), b1 = __ldg(pb + 1);
    Node4 P0{a0.x, a0.y, a1.x, 0.0}, P1{b0.x, b0.y, b1.x, 0.0};
    make_frame(P0, P1, nr, f);
    make_stiff(pr, f.L, s);
}

__device__ __forceinline__ void store_geom(ElemGeom *__restrict__ geom, int64_t e, const Frame &f, const Stiff &s,
                                           int32_t a, int32_t b)
{
    double2 *o = reinterpret_cast<double2 *>(geom + e);
    const long long ab = (long long)(((unsigned long long)(uint32_t)b << 32) | (uint32_t)a);
    o[0] = make_double2(f.nx[0], f.nx[1]); o[1] = make_double2(f.nx[2], f.ny[0]);
    o[2] = make_double2(f.ny[1], f.ny[2]); o[3] = make_double2(f.nz[0], f.nz[1]);
    o[4] = make_double2(f.nz[2], s.ea);    o[5] = make_double2(s.gj, s.z12);
    o[6] = make_double2(s.z6, s.z1x4);     o[7] = make_double2(s.z1x2, s.y12);
    o[8] = make_double2(s.y6, s.y1x4);     o[9] = make_double2(s.y1x2, __longlong_as_double(ab));
}

__device__ __forceinline__ void load_geom(const ElemGeom *__restrict__ geom, int64_t e, Frame &f, Stiff &s)
{
    const double2 *q = reinterpret_cast<const double2 *>(geom + e);
    const double2 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2), q3 = __ldg(q + 3), q4 = __ldg(q + 4),
                  q5 = __ldg(q + 5), q6 = __ldg(q + 6), q7 = __ldg(q + 7), q8 = __ldg(q + 8), q9 = __ldg(q + 9);
    f.nx[0] = q0.x; f.nx[1] = q0.y; f.nx[2] = q1.x; f.ny[0] = q1.y; f.ny[1] = q2.x; f.ny[2] = q2.y;
    f.nz[0] = q3.x; f.nz[1] = q3.y; f.nz[2] = q4.x; f.L = 0.0;
    s.ea = q4.y; s.gj = q5.x; s.z12 = q5.y; s.z6 = q6.x; s.z1x4 = q6.y; s.z1x2 = q7.x; s.y12 = q7.y;
    s.y6 = q8.x; s.y1x4 = q8.y; s.y1x2 = q9.x;
}

}}  // namespace mbfea::dev
