// Sanitizer harness for the host passes of set_job (not part of the product): lattices through validation, section
// properties, the plan, the solver format and the hierarchy builder (started at the free map and gated on the pattern,
// as set_job's host path does), one and three ranks, with and without internal renumbering, shared host threads.
//   cd meshbeamfea_b200/csrc && for s in thread address,undefined; do g++ -std=c++17 -O1 -g -fsanitize=$s -pthread \
//     -I/usr/local/cuda/include -I. ../../tests/tools/host_sanitizer_harness.cpp host_prep.cpp hierarchy.cpp hostmem.cpp \
//     dense.cpp -o /tmp/san_$s && MBFEA_HOST_THREADS=6 /tmp/san_$s; done
// Last run (final build of round 2): no report from ThreadSanitizer, AddressSanitizer or UBSan.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <cmath>
#include "mbfea_internal.hpp"
#include <cuda_runtime.h>
// stubs: no device in this harness
extern "C" {
cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaErrorNoDevice; }
cudaError_t cudaSetDevice(int) { return cudaErrorNoDevice; }
cudaError_t cudaHostAlloc(void **, size_t, unsigned int) { return cudaErrorNoDevice; }
cudaError_t cudaFreeHost(void *) { return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
}
using namespace mbfea;
int main()
{
    std::mt19937 rng(7);
    for (int trial = 0; trial < 6; ++trial) {
        const int n = 18 + 5 * trial;   // lattice n^3
        const int64_t V = (int64_t)n * n * n;
        std::vector<double> nodes((size_t)3 * V);
        std::vector<int32_t> ea, eb;
        auto id = [&](int x, int y, int z) { return (int32_t)(x + n * (y + n * z)); };
        for (int z = 0; z < n; ++z) for (int y = 0; y < n; ++y) for (int x = 0; x < n; ++x) {
            const int32_t v = id(x, y, z);
            nodes[v] = 0.1 * x; nodes[V + v] = 0.1 * y; nodes[2 * V + v] = 0.1 * z;
            if (x + 1 < n) { ea.push_back(v); eb.push_back(id(x + 1, y, z)); }
            if (y + 1 < n) { ea.push_back(v); eb.push_back(id(x, y + 1, z)); }
            if (z + 1 < n) { ea.push_back(v); eb.push_back(id(x, y, z + 1)); }
        }
        const int64_t E = (int64_t)ea.size();
        std::vector<int32_t> elems((size_t)2 * E);
        for (int64_t e = 0; e < E; ++e) { elems[e] = ea[e]; elems[E + e] = eb[e]; }
        std::vector<int32_t> fixed;
        for (int i = 0; i < n * n; ++i) fixed.push_back(i);
        std::vector<float> dims((size_t)2 * E, 0.01f), mat((size_t)2 * E);
        for (int64_t e = 0; e < E; ++e) { mat[2 * e] = 0.3f; mat[2 * e + 1] = 210.f; }
        std::string msg;
        int64_t fnode = V - 1, fdof = 2;
        for (int world : {1, 3}) {
            host_threads_share(world);
            if (validate_job(V, E, elems.data(), dims.data(), mat.data(), (int64_t)fixed.size(), fixed.data(), 1, &fnode, &fdof, msg)) return 1;
            std::vector<float> props((size_t)4 * E);
            derive_props(E, dims.data(), mat.data(), props.data());
            for (int rank = 0; rank < world; ++rank) {
                HostPlan P;
                build_plan_begin(V, E, elems.data(), (int64_t)fixed.size(), fixed.data(), rank, world, P, trial % 2 ? 2 : 1, nodes.data());
                // hierarchy thread started at the free map, gated on the pattern
                std::mutex m; std::condition_variable cv; int state = 0;
                MlHierarchy H;
                std::vector<double> pos((size_t)3 * P.nb_global);
                std::thread th;
                if (world == 1) th = std::thread([&] {
                    for (int64_t v = 0; v < V; ++v) { const int32_t r = P.free_index[(size_t)v]; if (r >= 0) for (int a = 0; a < 3; ++a) pos[(size_t)3 * r + a] = nodes[(size_t)a * V + v]; }
                    const FinePatternGate gate = [&](const int32_t *&rp, const int32_t *&ci) {
                        std::unique_lock<std::mutex> g(m); cv.wait(g, [&] { return state != 0; });
                        rp = P.rowptr.data(); ci = P.colidx.data(); return state == 1; };
                    build_ml_hierarchy(P.nb_global, nullptr, nullptr, pos.data(), 0.2, 2.0, 12, 64, 1, H, nullptr, 1, &gate);
                });
                build_plan_pattern(E, elems.data(), true, P);
                { std::lock_guard<std::mutex> g(m); state = 1; } cv.notify_all();
                build_solver_format(P, true);
                if (th.joinable()) th.join();
                std::printf("n=%d world=%d rank=%d rows=%lld blocks=%lld ghosts=%zu levels=%zu balanced=%d\n", n, world, rank, (long long)P.nb(),
                            (long long)P.nnzb(), P.ghost_global.size(), H.levels(), (int)P.balanced);
            }
            host_threads_share(1);
        }
    }
    return 0;
}
