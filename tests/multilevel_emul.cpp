// Thread-level CPU emulation of the DRAFT multilevel kernels (csrc/multilevel.cu): every CUDA thread of a block is a
// host thread, __syncthreads and __shfl_down_sync are barriers over the block / the warp.  Checks the kernels' indexing
// and reductions against dense reference arithmetic on a small lattice -- the closest thing to running them that is
// possible without a GPU.  (They remain unvalidated on hardware.)
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <memory>
#include <random>
#include <thread>
#include <vector>

struct EmulDim { unsigned x; };
struct EmulThread {
    unsigned bx, tx, bdim, gdim;
    std::barrier<> *block_bar;
    std::barrier<> *warp_bar;
    double *warp_slots;
};
static thread_local EmulThread *emu = nullptr;
#define blockIdx (EmulDim{emu->bx})
#define threadIdx (EmulDim{emu->tx})
#define blockDim (EmulDim{emu->bdim})
#define gridDim (EmulDim{emu->gdim})
#define __global__
#define __shared__ static   /* one block at a time: block-shared == process-static */
#define __launch_bounds__(...)
#define __restrict__
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline void __syncthreads() { emu->block_bar->arrive_and_wait(); }
static inline double __shfl_down_sync(unsigned, double v, int o)
{
    const unsigned lane = emu->tx & 31;
    emu->warp_slots[lane] = v;
    emu->warp_bar->arrive_and_wait();
    const double r = lane + (unsigned)o < 32 ? emu->warp_slots[lane + o] : v;
    emu->warp_bar->arrive_and_wait();
    return r;
}
#define MBFEA_HOST_EMULATION 1
#include "multilevel.cu"

template <class... A, class... B>
static void emu_launch(void (*k)(A...), unsigned grid, unsigned block, B... args)
{
    for (unsigned b = 0; b < grid; ++b) {
        std::barrier<> bb((std::ptrdiff_t)block);
        const unsigned nw = (block + 31) / 32;
        std::vector<std::unique_ptr<std::barrier<>>> wb;
        std::vector<std::vector<double>> slots(nw, std::vector<double>(32, 0.0));
        for (unsigned w = 0; w < nw; ++w) wb.emplace_back(new std::barrier<>((std::ptrdiff_t)std::min(32u, block - 32 * w)));
        std::vector<std::thread> th;
        for (unsigned t = 0; t < block; ++t)
            th.emplace_back([&, t] {
                EmulThread me{b, t, block, grid, &bb, wb[t / 32].get(), slots[t / 32].data()};
                emu = &me;
                k(static_cast<A>(args)...);
                // a thread that left early must not strand the others at a later barrier
                me.block_bar->arrive_and_drop();
                me.warp_bar->arrive_and_drop();
            });
        for (auto &x : th) x.join();
    }
}

#include "mbfea_internal.hpp"
#include "multilevel_apply.hpp"
using namespace mbfea;
using namespace mbfea::dev;

static double maxrel(const std::vector<double> &a, const std::vector<double> &b)
{
    double e = 0, s = 0;
    for (size_t i = 0; i < a.size(); ++i) { e = std::max(e, std::fabs(a[i] - b[i])); s = std::max(s, std::fabs(b[i])); }
    return e / (s > 0 ? s : 1.0);
}

int main()
{
    // a 6 x 5 x 4 lattice, nothing fixed except vertex 0; random SPD-ish blocks on its pattern
    const int nx = 6, ny = 5, nz = 4; const int64_t V = nx * ny * nz;
    std::vector<int32_t> a, b; std::vector<double> pos;
    for (int i = 0; i < nx; ++i) for (int j = 0; j < ny; ++j) for (int k = 0; k < nz; ++k) {
        const int32_t v = (i * ny + j) * nz + k;
        if (i + 1 < nx) { a.push_back(v); b.push_back(v + ny * nz); }
        if (j + 1 < ny) { a.push_back(v); b.push_back(v + nz); }
        if (k + 1 < nz) { a.push_back(v); b.push_back(v + 1); }
    }
    const int64_t E = (int64_t)a.size();
    std::vector<int32_t> el((size_t)2 * E); for (int64_t e = 0; e < E; ++e) { el[(size_t)e] = a[(size_t)e]; el[(size_t)(E + e)] = b[(size_t)e]; }
    const int32_t fixed[1] = {0};
    HostPlan P; build_plan(V, E, el.data(), 1, fixed, 0, 1, false, P);
    const int64_t nb = P.nb();
    pos.assign((size_t)3 * nb, 0.0);
    for (int i = 0; i < nx; ++i) for (int j = 0; j < ny; ++j) for (int k = 0; k < nz; ++k) {
        const int32_t v = (i * ny + j) * nz + k, r = P.free_index[(size_t)v];
        if (r >= 0) { pos[(size_t)3 * r] = 0.1 * i; pos[(size_t)3 * r + 1] = 0.13 * j; pos[(size_t)3 * r + 2] = 0.09 * k; }
    }
    Hierarchy H; build_hierarchy(nb, P.rowptr.data(), P.colidx.data(), pos.data(), 0.21, 2.0, 3, 1, H);
    if (H.levels.empty()) { std::printf("no levels\n"); return 1; }
    const HierLevel &L = H.levels[0];
    std::mt19937_64 g(9); std::uniform_real_distribution<double> U(-1, 1);
    const int64_t nnz = P.rowptr[(size_t)nb];
    std::vector<double> vals((size_t)nnz * 36);
    std::vector<int32_t> frow((size_t)nnz);
    for (int64_t r = 0; r < nb; ++r) for (int32_t k = P.rowptr[(size_t)r]; k < P.rowptr[(size_t)r + 1]; ++k) frow[(size_t)k] = (int32_t)r;
    // symmetric and strongly diagonally dominant (=> SPD, so are its Galerkin products): A(b, a) = A(a, b)^T
    for (int64_t r = 0; r < nb; ++r) for (int32_t k = P.rowptr[(size_t)r]; k < P.rowptr[(size_t)r + 1]; ++k) {
        const int32_t cidx = P.colidx[(size_t)k];
        if (cidx < r) continue;
        double Bk[6][6];
        for (auto &row : Bk) for (double &x : row) x = U(g);
        if (cidx == r) { for (int i = 0; i < 6; ++i) for (int j = 0; j < i; ++j) Bk[i][j] = Bk[j][i]; for (int i = 0; i < 6; ++i) Bk[i][i] += 60.0; }
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) vals[(size_t)k * 36 + 6 * i + j] = Bk[i][j];
        if (cidx != r) {
            const int32_t *cb = P.colidx.data() + P.rowptr[(size_t)cidx], *ce = P.colidx.data() + P.rowptr[(size_t)cidx + 1];
            const int64_t m = std::lower_bound(cb, ce, (int32_t)r) - P.colidx.data();
            for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) vals[(size_t)m * 36 + 6 * i + j] = Bk[j][i];
        }
    }
    // dense reference: P (6 nb x 6 nc) from agg / off, A dense
    const int64_t nc = L.n_coarse, nf = 6 * nb, nn = 6 * nc;
    std::vector<double> Pd((size_t)(nf * nn), 0.0), Ad((size_t)(nf * nf), 0.0);
    for (int64_t v = 0; v < nb; ++v) {
        const double *d = &L.off[(size_t)3 * v];
        double Bv[6][6] = {};
        for (int i = 0; i < 6; ++i) Bv[i][i] = 1.0;
        Bv[0][4] = d[2]; Bv[0][5] = -d[1]; Bv[1][3] = -d[2]; Bv[1][5] = d[0]; Bv[2][3] = d[1]; Bv[2][4] = -d[0];
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) Pd[(size_t)((6 * v + i) * nn + 6 * L.agg[(size_t)v] + j)] = Bv[i][j];
    }
    for (int64_t r = 0; r < nb; ++r) for (int32_t k = P.rowptr[(size_t)r]; k < P.rowptr[(size_t)r + 1]; ++k)
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) Ad[(size_t)((6 * r + i) * nf + 6 * P.colidx[(size_t)k] + j)] = vals[(size_t)k * 36 + 6 * i + j];
    std::vector<double> AP((size_t)(nf * nn), 0.0), Ac((size_t)(nn * nn), 0.0);
    for (int64_t i = 0; i < nf; ++i) for (int64_t k = 0; k < nf; ++k) { const double x = Ad[(size_t)(i * nf + k)]; if (x != 0) for (int64_t j = 0; j < nn; ++j) AP[(size_t)(i * nn + j)] += x * Pd[(size_t)(k * nn + j)]; }
    for (int64_t k = 0; k < nf; ++k) for (int64_t i = 0; i < nn; ++i) { const double x = Pd[(size_t)(k * nn + i)]; if (x != 0) for (int64_t j = 0; j < nn; ++j) Ac[(size_t)(i * nn + j)] += x * AP[(size_t)(k * nn + j)]; }
    int bad = 0;
    // ---- k_ml_galerkin ----
    const int64_t ncb = (int64_t)L.colidx.size();
    std::vector<double> cvals((size_t)ncb * 36, 0.0), cref((size_t)ncb * 36);
    emu_launch(k_ml_galerkin, (unsigned)((ncb * 36 + 255) / 256), 256u, ncb, L.gal_ptr.data(), L.gal_src.data(), frow.data(), P.colidx.data(),
               vals.data(), L.off.data(), cvals.data());
    for (int64_t I = 0; I < nc; ++I) for (int32_t k = L.rowptr[(size_t)I]; k < L.rowptr[(size_t)I + 1]; ++k)
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) cref[(size_t)k * 36 + 6 * i + j] = Ac[(size_t)((6 * I + i) * nn + 6 * L.colidx[(size_t)k] + j)];
    double e = maxrel(cvals, cref); std::printf("galerkin      %.2e\n", e); bad += !(e < 1e-12);
    // ---- k_ml_restrict (scaled and unscaled) ----
    std::vector<double> r((size_t)nf), lfac((size_t)nb * 36, 0.0), rc((size_t)nn), rref((size_t)nn, 0.0), w((size_t)nf);
    for (double &x : r) x = U(g);
    for (int64_t v = 0; v < nb; ++v) for (int i = 0; i < 6; ++i) for (int j = 0; j <= i; ++j) lfac[(size_t)v * 36 + 6 * i + j] = (i == j ? 2.0 : 0.0) + U(g);
    for (int scaled = 0; scaled < 2; ++scaled) {
        for (int64_t v = 0; v < nb; ++v) for (int i = 0; i < 6; ++i) {
            double s = 0; if (scaled) for (int j = 0; j <= i; ++j) s += lfac[(size_t)v * 36 + 6 * i + j] * r[(size_t)(6 * v + j)]; else s = r[(size_t)(6 * v + i)];
            w[(size_t)(6 * v + i)] = s;
        }
        std::fill(rref.begin(), rref.end(), 0.0);
        for (int64_t k = 0; k < nf; ++k) for (int64_t j = 0; j < nn; ++j) rref[(size_t)j] += Pd[(size_t)(k * nn + j)] * w[(size_t)k];
        const unsigned grid = (unsigned)((nc * 32 + 255) / 256);
        if (scaled) emu_launch(k_ml_restrict<true>, grid, 256u, nc, L.agg_ptr.data(), L.agg_rows.data(), L.off.data(), lfac.data(), r.data(), rc.data());
        else emu_launch(k_ml_restrict<false>, grid, 256u, nc, L.agg_ptr.data(), L.agg_rows.data(), L.off.data(), (const double *)nullptr, r.data(), rc.data());
        e = maxrel(rc, rref); std::printf("restrict[%d]   %.2e\n", scaled, e); bad += !(e < 1e-12);
    }
    // ---- k_ml_smooth: z = S^T S r ----
    std::vector<double> z((size_t)nf), zref((size_t)nf);
    for (int64_t v = 0; v < nb; ++v) {
        double t[6];
        for (int i = 0; i < 6; ++i) { t[i] = 0; for (int j = 0; j <= i; ++j) t[i] += lfac[(size_t)v * 36 + 6 * i + j] * r[(size_t)(6 * v + j)]; }
        for (int i = 0; i < 6; ++i) { double s = 0; for (int j = i; j < 6; ++j) s += lfac[(size_t)v * 36 + 6 * j + i] * t[j]; zref[(size_t)(6 * v + i)] = s; }
    }
    emu_launch(k_ml_smooth, (unsigned)((nb + 255) / 256), 256u, nb, lfac.data(), r.data(), z.data());
    e = maxrel(z, zref); std::printf("smooth        %.2e\n", e); bad += !(e < 1e-12);
    // ---- k_ml_dense_matvec ----
    std::vector<double> M((size_t)(nn * nn)), xc((size_t)nn), yc((size_t)nn), yref((size_t)nn, 0.0);
    for (double &x : M) x = U(g);
    for (double &x : xc) x = U(g);
    for (int64_t i = 0; i < nn; ++i) for (int64_t j = 0; j < nn; ++j) yref[(size_t)i] += M[(size_t)(i * nn + j)] * xc[(size_t)j];
    emu_launch(k_ml_dense_matvec, (unsigned)((nn * 32 + 255) / 256), 256u, nn, M.data(), xc.data(), yc.data());
    e = maxrel(yc, yref); std::printf("dense matvec  %.2e\n", e); bad += !(e < 1e-12);
    // ---- k_ml_prolong: z += P zc ----
    std::vector<double> z0((size_t)nf), z1((size_t)nf);
    for (double &x : z0) x = U(g);
    z1 = z0;
    for (int64_t k = 0; k < nf; ++k) for (int64_t j = 0; j < nn; ++j) z1[(size_t)k] += Pd[(size_t)(k * nn + j)] * xc[(size_t)j];
    emu_launch(k_ml_prolong, (unsigned)((nb + 255) / 256), 256u, nb, L.agg.data(), L.off.data(), xc.data(), z0.data());
    e = maxrel(z0, z1); std::printf("prolong       %.2e\n", e); bad += !(e < 1e-12);
    // ---- k_ml_finish: zt = rt + L^T (P zc), partials of rt . zt ----
    std::vector<double> zt((size_t)nf), ztref((size_t)nf), part(4, 0.0);
    double dotref = 0;
    for (int64_t v = 0; v < nb; ++v) {
        double t[6];
        for (int i = 0; i < 6; ++i) { t[i] = 0; for (int64_t j = 0; j < nn; ++j) t[i] += Pd[(size_t)((6 * v + i) * nn + j)] * xc[(size_t)j]; }
        for (int i = 0; i < 6; ++i) { double s = r[(size_t)(6 * v + i)]; for (int j = i; j < 6; ++j) s += lfac[(size_t)v * 36 + 6 * j + i] * t[j]; ztref[(size_t)(6 * v + i)] = s; dotref += r[(size_t)(6 * v + i)] * s; }
    }
    emu_launch(k_ml_finish, 2u, 256u, nb, L.agg.data(), L.off.data(), lfac.data(), r.data(), xc.data(), zt.data(), part.data());
    e = maxrel(zt, ztref); std::printf("finish        %.2e, dot %.2e\n", e, std::fabs(part[0] + part[1] - dotref) / std::fabs(dotref));
    bad += !(e < 1e-12) + !(std::fabs(part[0] + part[1] - dotref) <= 1e-12 * std::fabs(dotref));
    // ---- the whole preconditioner through ml_apply_levels (the code the GPU driver runs), two coarse levels ----
    if (H.levels.size() >= 2) {
        const HierLevel &L1 = H.levels[1];
        const int64_t n1 = L.n_coarse, n2 = L1.n_coarse, m1 = 6 * n1, m2 = 6 * n2;
        // level-2 Galerkin matrix from the level-1 blocks (emulated kernel, as the driver does)
        std::vector<int32_t> frow1((size_t)ncb);
        for (int64_t I = 0; I < n1; ++I) for (int32_t k = L.rowptr[(size_t)I]; k < L.rowptr[(size_t)I + 1]; ++k) frow1[(size_t)k] = (int32_t)I;
        const int64_t ncb2 = (int64_t)L1.colidx.size();
        std::vector<double> cvals2((size_t)ncb2 * 36);
        emu_launch(k_ml_galerkin, (unsigned)((ncb2 * 36 + 255) / 256), 256u, ncb2, L1.gal_ptr.data(), L1.gal_src.data(), frow1.data(),
                   L.colidx.data(), cvals.data(), L1.off.data(), cvals2.data());
        // dense P1, A1 (= Ac above), A2 = P1^T A1 P1 and its inverse (Gauss-Jordan, tiny)
        std::vector<double> P1((size_t)(m1 * m2), 0.0);
        for (int64_t v = 0; v < n1; ++v) {
            const double *d = &L1.off[(size_t)3 * v];
            double Bv[6][6] = {};
            for (int i = 0; i < 6; ++i) Bv[i][i] = 1.0;
            Bv[0][4] = d[2]; Bv[0][5] = -d[1]; Bv[1][3] = -d[2]; Bv[1][5] = d[0]; Bv[2][3] = d[1]; Bv[2][4] = -d[0];
            for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) P1[(size_t)((6 * v + i) * m2 + 6 * L1.agg[(size_t)v] + j)] = Bv[i][j];
        }
        std::vector<double> T((size_t)(m1 * m2), 0.0), A2((size_t)(m2 * m2), 0.0), A2k((size_t)(m2 * m2), 0.0);
        for (int64_t i = 0; i < m1; ++i) for (int64_t k = 0; k < m1; ++k) { const double x = Ac[(size_t)(i * m1 + k)]; if (x != 0) for (int64_t j = 0; j < m2; ++j) T[(size_t)(i * m2 + j)] += x * P1[(size_t)(k * m2 + j)]; }
        for (int64_t k = 0; k < m1; ++k) for (int64_t i = 0; i < m2; ++i) { const double x = P1[(size_t)(k * m2 + i)]; if (x != 0) for (int64_t j = 0; j < m2; ++j) A2[(size_t)(i * m2 + j)] += x * T[(size_t)(k * m2 + j)]; }
        for (int64_t I = 0; I < n2; ++I) for (int32_t k = L1.rowptr[(size_t)I]; k < L1.rowptr[(size_t)I + 1]; ++k)
            for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) A2k[(size_t)((6 * I + i) * m2 + 6 * L1.colidx[(size_t)k] + j)] = cvals2[(size_t)k * 36 + 6 * i + j];
        e = maxrel(A2k, A2); std::printf("galerkin lvl2 %.2e\n", e); bad += !(e < 1e-12);
        std::vector<double> Ainv2((size_t)(m2 * m2), 0.0), G = A2;
        for (int64_t i = 0; i < m2; ++i) Ainv2[(size_t)(i * m2 + i)] = 1.0;
        for (int64_t cidx = 0; cidx < m2; ++cidx) {
            const double pv = G[(size_t)(cidx * m2 + cidx)];
            for (int64_t j = 0; j < m2; ++j) { G[(size_t)(cidx * m2 + j)] /= pv; Ainv2[(size_t)(cidx * m2 + j)] /= pv; }
            for (int64_t i = 0; i < m2; ++i) if (i != cidx) {
                const double fct = G[(size_t)(i * m2 + cidx)];
                if (fct != 0) for (int64_t j = 0; j < m2; ++j) { G[(size_t)(i * m2 + j)] -= fct * G[(size_t)(cidx * m2 + j)]; Ainv2[(size_t)(i * m2 + j)] -= fct * Ainv2[(size_t)(cidx * m2 + j)]; }
            }
        }
        // S of the level-1 diagonal blocks (what k_block_chol produces): D = Lc Lc^T, S = Lc^-1
        std::vector<double> sinv1((size_t)n1 * 36, 0.0), D1inv((size_t)n1 * 36, 0.0);
        for (int64_t I = 0; I < n1; ++I) {
            double Dm[6][6], Lc[6][6] = {}, Sm[6][6] = {};
            for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) Dm[i][j] = Ac[(size_t)((6 * I + i) * m1 + 6 * I + j)];
            for (int j = 0; j < 6; ++j) {
                double d = Dm[j][j]; for (int k = 0; k < j; ++k) d -= Lc[j][k] * Lc[j][k];
                Lc[j][j] = std::sqrt(d);
                for (int i = j + 1; i < 6; ++i) { double v = Dm[i][j]; for (int k = 0; k < j; ++k) v -= Lc[i][k] * Lc[j][k]; Lc[i][j] = v / Lc[j][j]; }
            }
            for (int j = 0; j < 6; ++j) for (int i = j; i < 6; ++i) {
                if (i == j) { Sm[i][j] = 1.0 / Lc[j][j]; continue; }
                double v = 0; for (int k = j; k < i; ++k) v -= Lc[i][k] * Sm[k][j];
                Sm[i][j] = v / Lc[i][i];
            }
            for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) {
                sinv1[(size_t)I * 36 + 6 * i + j] = Sm[i][j];
                double v = 0; for (int k = 0; k < 6; ++k) v += Sm[k][i] * Sm[k][j];
                D1inv[(size_t)I * 36 + 6 * i + j] = v;
            }
        }
        // dense reference: z = r + L^T P0 ( D1^-1 r1 + P1 A2^-1 P1^T r1 ),  r1 = P0^T L r
        std::vector<double> w0((size_t)nf), r1((size_t)m1, 0.0), r2((size_t)m2, 0.0), z2((size_t)m2, 0.0), z1v((size_t)m1, 0.0), zfull((size_t)nf);
        for (int64_t v = 0; v < nb; ++v) for (int i = 0; i < 6; ++i) { double sacc = 0; for (int j = 0; j <= i; ++j) sacc += lfac[(size_t)v * 36 + 6 * i + j] * r[(size_t)(6 * v + j)]; w0[(size_t)(6 * v + i)] = sacc; }
        for (int64_t k = 0; k < nf; ++k) for (int64_t j = 0; j < m1; ++j) r1[(size_t)j] += Pd[(size_t)(k * nn + j)] * w0[(size_t)k];
        for (int64_t k = 0; k < m1; ++k) for (int64_t j = 0; j < m2; ++j) r2[(size_t)j] += P1[(size_t)(k * m2 + j)] * r1[(size_t)k];
        for (int64_t i = 0; i < m2; ++i) for (int64_t j = 0; j < m2; ++j) z2[(size_t)i] += Ainv2[(size_t)(i * m2 + j)] * r2[(size_t)j];
        for (int64_t I = 0; I < n1; ++I) for (int i = 0; i < 6; ++i) { double v = 0; for (int j = 0; j < 6; ++j) v += D1inv[(size_t)I * 36 + 6 * i + j] * r1[(size_t)(6 * I + j)]; z1v[(size_t)(6 * I + i)] = v; }
        for (int64_t k = 0; k < m1; ++k) for (int64_t j = 0; j < m2; ++j) z1v[(size_t)k] += P1[(size_t)(k * m2 + j)] * z2[(size_t)j];
        double dot2 = 0;
        for (int64_t v = 0; v < nb; ++v) {
            double t[6];
            for (int i = 0; i < 6; ++i) { t[i] = 0; for (int64_t j = 0; j < m1; ++j) t[i] += Pd[(size_t)((6 * v + i) * nn + j)] * z1v[(size_t)j]; }
            for (int i = 0; i < 6; ++i) { double sacc = r[(size_t)(6 * v + i)]; for (int j = i; j < 6; ++j) sacc += lfac[(size_t)v * 36 + 6 * j + i] * t[j]; zfull[(size_t)(6 * v + i)] = sacc; dot2 += r[(size_t)(6 * v + i)] * sacc; }
        }
        // the shared level loop with emulated launches
        struct EmuLauncher {
            void restrict(int64_t n, const int32_t *ap, const int32_t *ar, const double *off, const double *lf, const double *rr, double *rcv)
            {
                const unsigned grid = (unsigned)((n * 32 + 255) / 256);
                if (lf) emu_launch(k_ml_restrict<true>, grid, 256u, n, ap, ar, off, lf, rr, rcv);
                else emu_launch(k_ml_restrict<false>, grid, 256u, n, ap, ar, off, lf, rr, rcv);
            }
            void smooth(int64_t n, const double *sv, const double *rr, double *zz) { emu_launch(k_ml_smooth, (unsigned)((n + 255) / 256), 256u, n, sv, rr, zz); }
            void dense_matvec(int64_t n, const double *Ai, const double *rr, double *zz) { emu_launch(k_ml_dense_matvec, (unsigned)((n * 32 + 255) / 256), 256u, n, Ai, rr, zz); }
            void prolong(int64_t n, const int32_t *ag, const double *off, const double *zc, double *zz) { emu_launch(k_ml_prolong, (unsigned)((n + 255) / 256), 256u, n, ag, off, zc, zz); }
            void finish(int grid, int64_t nbb, const int32_t *ag, const double *off, const double *lf, const double *rtv, const double *z1p, double *ztv, double *pp)
            { emu_launch(k_ml_finish, (unsigned)grid, 256u, nbb, ag, off, lf, rtv, z1p, ztv, pp); }
        };
        std::vector<double> rA((size_t)m1), zA((size_t)m1), rB((size_t)m2), zB((size_t)m2), ztl((size_t)nf), partl(3, 0.0);
        mbfea::MlLevelView views[2] = {
            {nb, n1, L.agg.data(), L.agg_ptr.data(), L.agg_rows.data(), L.off.data(), sinv1.data(), rA.data(), zA.data()},
            {n1, n2, L1.agg.data(), L1.agg_ptr.data(), L1.agg_rows.data(), L1.off.data(), nullptr, rB.data(), zB.data()}};
        mbfea::ml_apply_levels(EmuLauncher{}, views, 2, Ainv2.data(), nb, lfac.data(), r.data(), ztl.data(), partl.data(), 3);
        e = maxrel(ztl, zfull);
        const double de = std::fabs(partl[0] + partl[1] + partl[2] - dot2) / std::fabs(dot2);
        std::printf("level loop    %.2e, dot %.2e  (rows %lld -> %lld -> %lld)\n", e, de, (long long)nb, (long long)n1, (long long)n2);
        bad += !(e < 1e-11) + !(de < 1e-11);
    } else { std::printf("only one coarse level\n"); bad += 1; }
    std::printf("%s\n", bad ? "MISMATCH" : "all kernels match");
    return bad != 0;
}
