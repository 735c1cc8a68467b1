/*
 * mbfea_oracle.c -- CPU ORACLE for the MeshBeamFEA simulation hot path.
 *
 * THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the
 * smoke check in __graft_entry__.py and bench.py's cpu_baseline / reference
 * arm may build, load or call it.  The product (libmbfea.so) never links it
 * and has no CPU fallback.
 *
 * What it restates
 * ----------------
 * The reference's BeamSimulator adapter [REF src/beamsimulator.cpp:40-210]
 * hands the whole numerical job to the third-party library `threed-beam-fea`
 * (https://github.com/IasonManolas/threed-beam-fea.git, GIT_TAG master --
 * floating, NOT pinned, NOT vendored: [REF CMakeLists.txt:15-23]).  That source
 * is absent from /root/reference and cannot be fetched (no network), and Eigen
 * is not installed, so the reference cannot be compiled here.  This file
 * therefore restates:
 *   (a) the parts the reference repo owns, verbatim in semantics:
 *       float section properties  [REF src/beamsimulator.cpp:173-188]
 *       6 BCs per fixed vertex     [REF src/beamsimulator.cpp:126-152]
 *       result layout              [REF src/beamsimulator.cpp:190-210]
 *   (b) the published algorithm of threed-beam-fea's fea::solve as evidenced
 *       by the reference's call sites [REF src/beamsimulator.cpp:42-66,89-91]:
 *       12x12 Euler-Bernoulli local stiffness, A = blockdiag(R,R,R,R),
 *       K_e = A^T K_l A, triplet assembly in element order, Lagrange or
 *       eliminated constraints, element forces f_e = K_l A u_e.
 *
 * PARITY PINNING: the reference holds no asserting tests.  This oracle is
 * pinned (tests/test_oracle_golden.py) against the only golden data the
 * reference carries: the README screenshot (Mz colour bar -200..100 for
 * Models/TwoLineSegments.ply, fixed {0}, force (2,1,100)) and the closed-form
 * answers of the BeamSimulatorTester cantilever cases
 * [REF src/beamsimulatortester.cpp:8-41,110-163].  Everything inside
 * fea::solve beyond those vectors is "parity unpinned".
 *
 * Build: see oracle/Makefile  (gcc -O2 -fopenmp, no fast-math).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------ */
/* (1) Section properties in IEEE float, promotion rules exactly as written
 *     in the reference  [REF src/beamsimulator.cpp:173-188].
 *     C has no std::pow(float,int) overload; C++11's promotes both arguments
 *     to double, so pow((double)h, 3.0) is the same libm call.             */
ORC_API void orc_beam_props(int64_t E, const float *dims_bh,
                            const float *mat_nuE, double *props4)
{
    for (int64_t e = 0; e < E; ++e) {
        const float b = dims_bh[2 * e + 0], h = dims_bh[2 * e + 1];
        const float nu = mat_nuE[2 * e + 0], Egpa = mat_nuE[2 * e + 1];
        const float crossArea = b * h;                                   /* :175 */
        const float I2 = (float)((double)b * pow((double)h, 3.0) / 12);  /* :176 */
        const float I3 = (float)((double)h * pow((double)b, 3.0) / 12);  /* :177 */
        const float polarInertia = I2 + I3;                              /* :178 */
        const float youngsPa = (float)((double)Egpa * pow(10.0, 9.0));   /* :180-181 */
        const float G = youngsPa / (2 * (1 + nu));                       /* :183 */
        const float EA = youngsPa * crossArea;                           /* :184 */
        const float EIy = youngsPa * I3;                                 /* :185 */
        const float EIz = youngsPa * I2;                                 /* :186 */
        const float GJ = G * polarInertia;                               /* :187 */
        /* order of fea::Props(EA, EIz, EIy, GJ, normal)  [REF :89-90]        */
        props4[4 * e + 0] = (double)EA;
        props4[4 * e + 1] = (double)EIz;
        props4[4 * e + 2] = (double)EIy;
        props4[4 * e + 3] = (double)GJ;
    }
}

/* ------------------------------------------------------------------------ */
/* (2) Element matrices.  Kl: local 12x12; A: 12x12 rotation; Ke = A^T Kl A.
 *     All row-major 12x12.  Returns L.                                      */
static void mat12_mul(const double *X, const double *Y, double *Z)
{
    for (int i = 0; i < 12; ++i)
        for (int j = 0; j < 12; ++j) {
            double s = 0.0;
            for (int k = 0; k < 12; ++k) s += X[i * 12 + k] * Y[k * 12 + j];
            Z[i * 12 + j] = s;
        }
}

static double elem_local(const double p0[3], const double p1[3],
                         const double nrm[3], const double pr[4],
                         double *Kl, double *A)
{
    const double EA = pr[0], EIz = pr[1], EIy = pr[2], GJ = pr[3];
    double d[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
    const double L = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    memset(Kl, 0, 144 * sizeof(double));
    memset(A, 0, 144 * sizeof(double));

    const double tmpea = EA / L, tmpgj = GJ / L;
    const double tmp12z = 12.0 * EIz / (L * L * L);
    const double tmp6z = 6.0 * EIz / (L * L);
    const double tmp1z = EIz / L;
    const double tmp12y = 12.0 * EIy / (L * L * L);
    const double tmp6y = 6.0 * EIy / (L * L);
    const double tmp1y = EIy / L;
#define KL(i, j) Kl[(i) * 12 + (j)]
    KL(0, 0) = tmpea;   KL(0, 6) = -tmpea;
    KL(1, 1) = tmp12z;  KL(1, 5) = tmp6z;   KL(1, 7) = -tmp12z;  KL(1, 11) = tmp6z;
    KL(2, 2) = tmp12y;  KL(2, 4) = -tmp6y;  KL(2, 8) = -tmp12y;  KL(2, 10) = -tmp6y;
    KL(3, 3) = tmpgj;   KL(3, 9) = -tmpgj;
    KL(4, 4) = 4.0 * tmp1y;  KL(4, 8) = tmp6y;   KL(4, 10) = 2.0 * tmp1y;
    KL(5, 5) = 4.0 * tmp1z;  KL(5, 7) = -tmp6z;  KL(5, 11) = 2.0 * tmp1z;
    KL(6, 6) = tmpea;
    KL(7, 7) = tmp12z;  KL(7, 11) = -tmp6z;
    KL(8, 8) = tmp12y;  KL(8, 10) = tmp6y;
    KL(9, 9) = tmpgj;
    KL(10, 10) = 4.0 * tmp1y;
    KL(11, 11) = 4.0 * tmp1z;
    for (int i = 0; i < 12; ++i)
        for (int j = i + 1; j < 12; ++j) KL(j, i) = KL(i, j);
#undef KL

    /* local frame: x = edge, y = given normal, z = x cross y.  The upstream
     * library normalises y, divides z = x cross y by its SQUARED norm, then
     * rebuilds y = z cross x (again divided by its squared norm).  For a unit
     * normal perpendicular to the edge all of these are no-ops up to an ulp.
     * Frame convention cross-evidenced by [REF src/mesh.hpp:163-179].        */
    double nx[3] = {d[0] / L, d[1] / L, d[2] / L};
    const double nl = sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
    double ny[3] = {nrm[0] / nl, nrm[1] / nl, nrm[2] / nl};
    double nz[3] = {nx[1] * ny[2] - nx[2] * ny[1], nx[2] * ny[0] - nx[0] * ny[2],
                    nx[0] * ny[1] - nx[1] * ny[0]};
    const double dlz = nz[0] * nz[0] + nz[1] * nz[1] + nz[2] * nz[2];
    nz[0] /= dlz; nz[1] /= dlz; nz[2] /= dlz;
    ny[0] = nz[1] * nx[2] - nz[2] * nx[1];
    ny[1] = nz[2] * nx[0] - nz[0] * nx[2];
    ny[2] = nz[0] * nx[1] - nz[1] * nx[0];
    const double dly = ny[0] * ny[0] + ny[1] * ny[1] + ny[2] * ny[2];
    ny[0] /= dly; ny[1] /= dly; ny[2] /= dly;
    for (int blk = 0; blk < 4; ++blk)
        for (int c = 0; c < 3; ++c) {
            A[(3 * blk + 0) * 12 + 3 * blk + c] = nx[c];
            A[(3 * blk + 1) * 12 + 3 * blk + c] = ny[c];
            A[(3 * blk + 2) * 12 + 3 * blk + c] = nz[c];
        }
    return L;
}

ORC_API double orc_kelem(const double *p0, const double *p1, const double *nrm,
                         const double *props4, double *Kl, double *A, double *Ke)
{
    double AT[144], T[144];
    const double L = elem_local(p0, p1, nrm, props4, Kl, A);
    for (int i = 0; i < 12; ++i)
        for (int j = 0; j < 12; ++j) AT[i * 12 + j] = A[j * 12 + i];
    mat12_mul(AT, Kl, T); /* (A^T Kl) A, left to right */
    mat12_mul(T, A, Ke);
    return L;
}

/* Ke for every element: out E x 144 (row-major 12x12 each). */
ORC_API void orc_kelem_all(int64_t V, int64_t E, const double *nodes_cm,
                           const int32_t *elems_cm, const double *normals_cm,
                           const double *props4, double *Ke_all)
{
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < E; ++e) {
        const int32_t a = elems_cm[e], b = elems_cm[E + e];
        double p0[3] = {nodes_cm[a], nodes_cm[V + a], nodes_cm[2 * V + a]};
        double p1[3] = {nodes_cm[b], nodes_cm[V + b], nodes_cm[2 * V + b]};
        double n[3] = {normals_cm[e], normals_cm[E + e], normals_cm[2 * E + e]};
        double Kl[144], A[144];
        orc_kelem(p0, p1, n, props4 + 4 * e, Kl, A, Ke_all + 144 * e);
    }
}

/* ------------------------------------------------------------------------ */
/* (3) Triplet assembly in element order into a scalar CSR of order 6V.
 *     Mirrors `Kelem.sparseView()` (exact zeros are not pushed) followed by
 *     SparseMatrix::setFromTriplets (duplicates summed in push order).
 *     Two calls: orc_assemble_count() sizes the output, orc_assemble_fill()
 *     writes it.  Serial, like the reference.                               */
typedef struct { int64_t row; int32_t col; double val; } trip_t;

static int64_t build_triplets(int64_t V, int64_t E, const double *nodes_cm,
                              const int32_t *elems_cm, const double *normals_cm,
                              const double *props4, trip_t *out /* may be NULL */,
                              int64_t *row_count /* 6V, may be NULL */)
{
    int64_t n = 0;
    for (int64_t e = 0; e < E; ++e) {
        const int32_t nd[2] = {elems_cm[e], elems_cm[E + e]};
        double p0[3] = {nodes_cm[nd[0]], nodes_cm[V + nd[0]], nodes_cm[2 * V + nd[0]]};
        double p1[3] = {nodes_cm[nd[1]], nodes_cm[V + nd[1]], nodes_cm[2 * V + nd[1]]};
        double nr[3] = {normals_cm[e], normals_cm[E + e], normals_cm[2 * E + e]};
        double Kl[144], A[144], Ke[144];
        orc_kelem(p0, p1, nr, props4 + 4 * e, Kl, A, Ke);
        for (int i = 0; i < 12; ++i)
            for (int j = 0; j < 12; ++j) {
                const double v = Ke[i * 12 + j];
                if (v == 0.0) continue; /* sparseView drops exact zeros */
                const int64_t r = 6 * (int64_t)nd[i / 6] + i % 6;
                const int64_t c = 6 * (int64_t)nd[j / 6] + j % 6;
                if (out) { out[n].row = r; out[n].col = (int32_t)c; out[n].val = v; }
                if (row_count) row_count[r]++;
                ++n;
            }
    }
    return n;
}

/* returns nnz of the summed matrix; rowptr (6V+1) is filled. */
ORC_API int64_t orc_assemble_count(int64_t V, int64_t E, const double *nodes_cm,
                                   const int32_t *elems_cm, const double *normals_cm,
                                   const double *props4, int64_t *rowptr)
{
    const int64_t N = 6 * V;
    int64_t *cnt = (int64_t *)calloc((size_t)N + 1, sizeof(int64_t));
    const int64_t nt = build_triplets(V, E, nodes_cm, elems_cm, normals_cm, props4, NULL, cnt);
    trip_t *t = (trip_t *)malloc((size_t)(nt ? nt : 1) * sizeof(trip_t));
    build_triplets(V, E, nodes_cm, elems_cm, normals_cm, props4, t, NULL);
    /* stable counting sort by row */
    int64_t *start = (int64_t *)malloc(((size_t)N + 1) * sizeof(int64_t));
    start[0] = 0;
    for (int64_t r = 0; r < N; ++r) start[r + 1] = start[r] + cnt[r];
    int32_t *cols = (int32_t *)malloc((size_t)(nt ? nt : 1) * sizeof(int32_t));
    int64_t *pos = (int64_t *)malloc(((size_t)N + 1) * sizeof(int64_t));
    memcpy(pos, start, ((size_t)N + 1) * sizeof(int64_t));
    for (int64_t k = 0; k < nt; ++k) cols[pos[t[k].row]++] = t[k].col;
    int64_t nnz = 0;
    rowptr[0] = 0;
    for (int64_t r = 0; r < N; ++r) {
        /* count unique columns of this row */
        int64_t b = start[r], e2 = start[r + 1];
        /* insertion sort (rows are short) */
        for (int64_t i = b + 1; i < e2; ++i) {
            int32_t c = cols[i]; int64_t j = i - 1;
            while (j >= b && cols[j] > c) { cols[j + 1] = cols[j]; --j; }
            cols[j + 1] = c;
        }
        for (int64_t i = b; i < e2; ++i)
            if (i == b || cols[i] != cols[i - 1]) ++nnz;
        rowptr[r + 1] = nnz;
    }
    free(cnt); free(t); free(start); free(cols); free(pos);
    return nnz;
}

ORC_API void orc_assemble_fill(int64_t V, int64_t E, const double *nodes_cm,
                               const int32_t *elems_cm, const double *normals_cm,
                               const double *props4, const int64_t *rowptr,
                               int32_t *colidx, double *vals)
{
    const int64_t N = 6 * V;
    int64_t *cnt = (int64_t *)calloc((size_t)N + 1, sizeof(int64_t));
    const int64_t nt = build_triplets(V, E, nodes_cm, elems_cm, normals_cm, props4, NULL, cnt);
    trip_t *t = (trip_t *)malloc((size_t)(nt ? nt : 1) * sizeof(trip_t));
    build_triplets(V, E, nodes_cm, elems_cm, normals_cm, props4, t, NULL);
    int64_t *start = (int64_t *)malloc(((size_t)N + 1) * sizeof(int64_t));
    start[0] = 0;
    for (int64_t r = 0; r < N; ++r) start[r + 1] = start[r] + cnt[r];
    trip_t *s = (trip_t *)malloc((size_t)(nt ? nt : 1) * sizeof(trip_t));
    int64_t *pos = (int64_t *)malloc(((size_t)N + 1) * sizeof(int64_t));
    memcpy(pos, start, ((size_t)N + 1) * sizeof(int64_t));
    for (int64_t k = 0; k < nt; ++k) s[pos[t[k].row]++] = t[k]; /* stable in push order */
    for (int64_t r = 0; r < N; ++r) {
        const int64_t b = start[r], e2 = start[r + 1];
        /* stable insertion sort by column keeps push order among duplicates */
        for (int64_t i = b + 1; i < e2; ++i) {
            trip_t x = s[i]; int64_t j = i - 1;
            while (j >= b && s[j].col > x.col) { s[j + 1] = s[j]; --j; }
            s[j + 1] = x;
        }
        int64_t o = rowptr[r] - 1;
        for (int64_t i = b; i < e2; ++i) {
            if (i == b || s[i].col != s[i - 1].col) { ++o; colidx[o] = s[i].col; vals[o] = s[i].val; }
            else vals[o] += s[i].val; /* duplicates summed in element order */
        }
    }
    free(cnt); free(t); free(start); free(s); free(pos);
}

/* ------------------------------------------------------------------------ */
/* (4) Constraint map: every fixed vertex loses all 6 DOFs
 *     [REF src/beamsimulator.cpp:139-150].  free_index[v] = -1 if fixed else
 *     rank among free vertices.  Returns #free vertices, or -1 when a fixed
 *     index is out of range (the reference throws runtime_error, :134-137). */
ORC_API int64_t orc_free_map(int64_t V, int64_t F, const int32_t *fixed, int32_t *free_index)
{
    for (int64_t v = 0; v < V; ++v) free_index[v] = 0;
    for (int64_t i = 0; i < F; ++i) {
        if (fixed[i] < 0 || (int64_t)fixed[i] > V - 1) return -1;
        free_index[fixed[i]] = -1;
    }
    int32_t k = 0;
    for (int64_t v = 0; v < V; ++v)
        if (free_index[v] == 0) free_index[v] = k++;
    return k;
}

/* ------------------------------------------------------------------------ */
/* (5) Element forces  f_e = Kl * (A * u_e), u_e = [u(n0,0..5), u(n1,0..5)];
 *     U is V x 6 ROW-major here (u[6v+d], the fea::Summary layout).
 *     Output in the SimulationResults layout [REF src/beamsimulator.cpp:198-207]:
 *     F6x2E[c*(2E) + 2e] = f_e[c], F6x2E[c*(2E) + 2e+1] = f_e[6+c].
 *     Optionally also the raw E x 12 table (may be NULL).                    */
ORC_API void orc_element_forces(int64_t V, int64_t E, const double *nodes_cm,
                                const int32_t *elems_cm, const double *normals_cm,
                                const double *props4, const double *U_rowmajor,
                                double *F6x2E, double *Fe_Ex12)
{
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < E; ++e) {
        const int32_t a = elems_cm[e], b = elems_cm[E + e];
        double p0[3] = {nodes_cm[a], nodes_cm[V + a], nodes_cm[2 * V + a]};
        double p1[3] = {nodes_cm[b], nodes_cm[V + b], nodes_cm[2 * V + b]};
        double n[3] = {normals_cm[e], normals_cm[E + e], normals_cm[2 * E + e]};
        double Kl[144], A[144], ue[12], ul[12], fe[12];
        elem_local(p0, p1, n, props4 + 4 * e, Kl, A);
        for (int d = 0; d < 6; ++d) { ue[d] = U_rowmajor[6 * (int64_t)a + d]; ue[6 + d] = U_rowmajor[6 * (int64_t)b + d]; }
        for (int i = 0; i < 12; ++i) { double s = 0; for (int k = 0; k < 12; ++k) s += A[i * 12 + k] * ue[k]; ul[i] = s; }
        for (int i = 0; i < 12; ++i) { double s = 0; for (int k = 0; k < 12; ++k) s += Kl[i * 12 + k] * ul[k]; fe[i] = s; }
        for (int c = 0; c < 6; ++c) {
            F6x2E[(int64_t)c * 2 * E + 2 * e] = fe[c];
            F6x2E[(int64_t)c * 2 * E + 2 * e + 1] = fe[6 + c];
        }
        if (Fe_Ex12) memcpy(Fe_Ex12 + 12 * e, fe, sizeof fe);
    }
}

/* ------------------------------------------------------------------------ */
/* (6) CPU baseline solver: 6x6-block BSR assembly on the eliminated system
 *     (element-order accumulation) + block-Jacobi PCG in FP64, OpenMP.
 *     This is the like-for-like CPU number for sizes where a sparse direct
 *     factorisation (what the reference's upstream library uses) does not
 *     finish on the host.  Also used by the tests as an independent check of
 *     the product's block pattern.                                           */

/* vertex-level pattern of the reduced system.  rowptr has nb+1 entries.
 * Pass colidx = NULL to size.  Columns sorted ascending, diagonal included.  */
ORC_API int64_t orc_bsr_pattern(int64_t V, int64_t E, const int32_t *elems_cm,
                                const int32_t *free_index, int64_t nb,
                                int64_t *rowptr, int32_t *colidx)
{
    int64_t *deg = (int64_t *)calloc((size_t)nb + 1, sizeof(int64_t));
    for (int64_t e = 0; e < E; ++e) {
        const int32_t fa = free_index[elems_cm[e]], fb = free_index[elems_cm[E + e]];
        if (fa >= 0 && fb >= 0 && fa != fb) { deg[fa]++; deg[fb]++; }
    }
    int64_t *st = (int64_t *)malloc(((size_t)nb + 1) * sizeof(int64_t));
    st[0] = 0;
    for (int64_t r = 0; r < nb; ++r) st[r + 1] = st[r] + deg[r] + 1;
    int32_t *tmp = (int32_t *)malloc((size_t)(st[nb] ? st[nb] : 1) * sizeof(int32_t));
    int64_t *pos = (int64_t *)malloc(((size_t)nb + 1) * sizeof(int64_t));
    for (int64_t r = 0; r < nb; ++r) { tmp[st[r]] = (int32_t)r; pos[r] = st[r] + 1; }
    for (int64_t e = 0; e < E; ++e) {
        const int32_t fa = free_index[elems_cm[e]], fb = free_index[elems_cm[E + e]];
        if (fa >= 0 && fb >= 0 && fa != fb) { tmp[pos[fa]++] = fb; tmp[pos[fb]++] = fa; }
    }
    int64_t nnzb = 0;
    rowptr[0] = 0;
    for (int64_t r = 0; r < nb; ++r) {
        const int64_t b = st[r], e2 = st[r + 1];
        for (int64_t i = b + 1; i < e2; ++i) {
            int32_t c = tmp[i]; int64_t j = i - 1;
            while (j >= b && tmp[j] > c) { tmp[j + 1] = tmp[j]; --j; }
            tmp[j + 1] = c;
        }
        for (int64_t i = b; i < e2; ++i)
            if (i == b || tmp[i] != tmp[i - 1]) { if (colidx) colidx[nnzb] = tmp[i]; ++nnzb; }
        rowptr[r + 1] = nnzb;
    }
    free(deg); free(st); free(tmp); free(pos);
    (void)V;
    return nnzb;
}

static int64_t find_block(const int64_t *rowptr, const int32_t *colidx, int64_t r, int32_t c)
{
    int64_t lo = rowptr[r], hi = rowptr[r + 1] - 1;
    while (lo <= hi) {
        int64_t m = (lo + hi) / 2;
        if (colidx[m] == c) return m;
        if (colidx[m] < c) lo = m + 1; else hi = m - 1;
    }
    return -1;
}

/* values: nnzb x 36, row-major 6x6 blocks, accumulated in element order.    */
ORC_API void orc_bsr_assemble(int64_t V, int64_t E, const double *nodes_cm,
                              const int32_t *elems_cm, const double *normals_cm,
                              const double *props4, const int32_t *free_index,
                              int64_t nb, const int64_t *rowptr, const int32_t *colidx,
                              double *vals)
{
    memset(vals, 0, (size_t)rowptr[nb] * 36 * sizeof(double));
    for (int64_t e = 0; e < E; ++e) {
        const int32_t nd[2] = {elems_cm[e], elems_cm[E + e]};
        const int32_t fr[2] = {free_index[nd[0]], free_index[nd[1]]};
        if (fr[0] < 0 && fr[1] < 0) continue;
        double p0[3] = {nodes_cm[nd[0]], nodes_cm[V + nd[0]], nodes_cm[2 * V + nd[0]]};
        double p1[3] = {nodes_cm[nd[1]], nodes_cm[V + nd[1]], nodes_cm[2 * V + nd[1]]};
        double nr[3] = {normals_cm[e], normals_cm[E + e], normals_cm[2 * E + e]};
        double Kl[144], A[144], Ke[144];
        orc_kelem(p0, p1, nr, props4 + 4 * e, Kl, A, Ke);
        for (int a = 0; a < 2; ++a)
            for (int b = 0; b < 2; ++b) {
                if (fr[a] < 0 || fr[b] < 0) continue;
                const int64_t k = find_block(rowptr, colidx, fr[a], fr[b]);
                double *dst = vals + 36 * k;
                for (int i = 0; i < 6; ++i)
                    for (int j = 0; j < 6; ++j) dst[6 * i + j] += Ke[(6 * a + i) * 12 + 6 * b + j];
            }
    }
}

static void inv6_spd(const double *Ain, double *inv)
{
    /* Gauss-Jordan without pivoting is fine for SPD diagonal blocks */
    double M[6][12];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) { M[i][j] = Ain[6 * i + j]; M[i][6 + j] = (i == j); }
    for (int k = 0; k < 6; ++k) {
        const double p = 1.0 / M[k][k];
        for (int j = 0; j < 12; ++j) M[k][j] *= p;
        for (int i = 0; i < 6; ++i) {
            if (i == k) continue;
            const double f = M[i][k];
            for (int j = 0; j < 12; ++j) M[i][j] -= f * M[k][j];
        }
    }
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) inv[6 * i + j] = M[i][6 + j];
}

ORC_API void orc_bsr_spmv(int64_t nb, const int64_t *rowptr, const int32_t *colidx,
                          const double *vals, const double *x, double *y)
{
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < nb; ++r) {
        double acc[6] = {0, 0, 0, 0, 0, 0};
        for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k) {
            const double *B = vals + 36 * k;
            const double *xx = x + 6 * (int64_t)colidx[k];
            for (int i = 0; i < 6; ++i)
                for (int j = 0; j < 6; ++j) acc[i] += B[6 * i + j] * xx[j];
        }
        for (int i = 0; i < 6; ++i) y[6 * r + i] = acc[i];
    }
}

/* Block-Jacobi PCG; returns iterations done.  stop: ||r|| <= tol*||f||.
 * max_iter bounds the work (the bench uses it to time a fixed sample).      */
/* optional wall-clock budget of orc_pcg_bj (seconds; <= 0: none): the bench's reference arm runs the CPU solve
 * to completion but must not hold the box for ever if the host is slow */
static double g_pcg_budget_s = 0.0;
ORC_API void orc_set_pcg_budget(double seconds) { g_pcg_budget_s = seconds; }

ORC_API int64_t orc_pcg_bj(int64_t nb, const int64_t *rowptr, const int32_t *colidx,
                           const double *vals, const double *f, double *x,
                           double tol, int64_t max_iter, double *rel_res_out)
{
    const int64_t n = 6 * nb;
    double *Minv = (double *)malloc((size_t)nb * 36 * sizeof(double));
    double *r = (double *)malloc((size_t)n * sizeof(double));
    double *z = (double *)malloc((size_t)n * sizeof(double));
    double *p = (double *)malloc((size_t)n * sizeof(double));
    double *Ap = (double *)malloc((size_t)n * sizeof(double));
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nb; ++i) {
        const int64_t k = find_block(rowptr, colidx, i, (int32_t)i);
        inv6_spd(vals + 36 * k, Minv + 36 * i);
    }
    double ff = 0, rz = 0;
#pragma omp parallel for schedule(static) reduction(+ : ff)
    for (int64_t i = 0; i < n; ++i) { x[i] = 0; r[i] = f[i]; ff += f[i] * f[i]; }
#pragma omp parallel for schedule(static) reduction(+ : rz)
    for (int64_t b = 0; b < nb; ++b)
        for (int i = 0; i < 6; ++i) {
            double s = 0;
            for (int j = 0; j < 6; ++j) s += Minv[36 * b + 6 * i + j] * r[6 * b + j];
            z[6 * b + i] = s; p[6 * b + i] = s; rz += r[6 * b + i] * s;
        }
    int64_t it = 0;
    double rr = ff;
    if (ff == 0) { if (rel_res_out) *rel_res_out = 0; goto done; }
    const double t_start = omp_get_wtime();
    while (it < max_iter && rr > tol * tol * ff) {
        if (g_pcg_budget_s > 0.0 && omp_get_wtime() - t_start > g_pcg_budget_s) break;
        orc_bsr_spmv(nb, rowptr, colidx, vals, p, Ap);
        double pAp = 0;
#pragma omp parallel for schedule(static) reduction(+ : pAp)
        for (int64_t i = 0; i < n; ++i) pAp += p[i] * Ap[i];
        const double alpha = rz / pAp;
        double rz_new = 0; rr = 0;
#pragma omp parallel for schedule(static) reduction(+ : rz_new, rr)
        for (int64_t b = 0; b < nb; ++b) {
            for (int i = 0; i < 6; ++i) {
                x[6 * b + i] += alpha * p[6 * b + i];
                r[6 * b + i] -= alpha * Ap[6 * b + i];
            }
            for (int i = 0; i < 6; ++i) {
                double s = 0;
                for (int j = 0; j < 6; ++j) s += Minv[36 * b + 6 * i + j] * r[6 * b + j];
                z[6 * b + i] = s;
                rz_new += r[6 * b + i] * s;
                rr += r[6 * b + i] * r[6 * b + i];
            }
        }
        const double beta = rz_new / rz;
        rz = rz_new;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
        ++it;
    }
    if (rel_res_out) *rel_res_out = sqrt(rr / ff);
done:
    free(Minv); free(r); free(z); free(p); free(Ap);
    return it;
}

ORC_API int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

ORC_API void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---------------------------------------------------------------------------
 * Extended-precision residual for the higher-precision reference solve
 * (oracle.solve_dd): r = f - A x with x = x_hi + x_lo and every product and sum
 * carried in double-double (error-free TwoProd / TwoSum), rounded to double at
 * the end.  A: scalar CSR.  The result is exact to ~1e-30 |A||x|, so iterative
 * refinement on top of the FP64 SuperLU factors converges to the solution of
 * the FP64-represented system far below the 1e-8 parity bar -- which is what
 * decides, on the ill-conditioned configs, whether the GPU solve or the plain
 * oracle solve is the one that is off.
 * ------------------------------------------------------------------------- */
static inline void dd_two_sum(double a, double b, double *s, double *e)
{
    const double t = a + b, bb = t - a;
    *s = t; *e = (a - (t - bb)) + (b - bb);
}
ORC_API void orc_csr_residual_dd(int64_t n, const int64_t *rowptr, const int32_t *col, const double *vals,
                                 const double *x_hi, const double *x_lo, const double *f, double *r)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double sh = 0.0, sl = 0.0;
        for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            const double a = vals[k], xh = x_hi[col[k]], xl = x_lo ? x_lo[col[k]] : 0.0;
            const double p = a * xh;
            double e = fma(a, xh, -p);      /* exact: a*xh = p + e */
            e += a * xl;
            double t, te;
            dd_two_sum(sh, p, &t, &te);
            te += sl + e;
            dd_two_sum(t, te, &sh, &sl);
        }
        double t, te;
        dd_two_sum(f[i], -sh, &t, &te);
        r[i] = t + (te - sl);
    }
}
