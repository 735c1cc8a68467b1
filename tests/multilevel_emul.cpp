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
    for (double &x : vals) x = U(g);
    std::vector<int32_t> frow((size_t)nnz);
    for (int64_t r = 0; r < nb; ++r) for (int32_t k = P.rowptr[(size_t)r]; k < P.rowptr[(size_t)r + 1]; ++k) frow[(size_t)k] = (int32_t)r;
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
    std::printf("%s\n", bad ? "MISMATCH" : "all kernels match");
    return bad != 0;
}
