// Protocol-level simulation of the single-reduction CG exchange (Mailbox::ll, parity double-buffered) plus the
// neighbour halo tags: W "ranks" as threads with random delays; checks every value a rank consumes is the one
// its peer published for that iteration.
#include <atomic>
#include <cstdio>
#include <cstdint>
#include <random>
#include <thread>
#include <vector>
constexpr int W = 8, ITERS = 20000;
struct Mailbox { std::atomic<uint64_t> ll[2][2][W]; std::atomic<uint64_t> halo_tag[W]; std::atomic<uint64_t> ghost[W]; };
static Mailbox mb[W];
static uint64_t flag_of(uint64_t epoch, long it) { return 0x80000000u | ((epoch & 0x3f) << 25) | ((uint64_t)(it + 1) & 0x1ffffff); }
static uint32_t val_of(int q, long it) { return (uint32_t)(q * 2654435761u + it * 40503u); }
static std::atomic<long> errors{0};
void rank(int me) {
    std::mt19937 g(me * 77 + 1);
    const uint64_t epoch = 5;
    // neighbours in a chain (slab partition)
    std::vector<int> nb; if (me > 0) nb.push_back(me - 1); if (me + 1 < W) nb.push_back(me + 1);
    for (long it = 0; it < ITERS; ++it) {
        // SpMV(it): wait for the neighbours' halo of iteration it (pushed by their step(it-1)), read ghosts
        if (it > 0) for (int q : nb) { while (mb[me].halo_tag[q].load(std::memory_order_acquire) != (epoch << 32 | (uint64_t)it)) std::this_thread::yield();
                                       
                                        if (mb[me].ghost[q].load(std::memory_order_relaxed) != (uint64_t)val_of(q, it - 1) + 7) errors++; }
        if ((g() & 63) == 0) std::this_thread::yield();
        // last CTA: publish both dots to every mailbox (two words each in the real thing; one here)
        const int par = (int)(it & 1);
        for (int q = 0; q < W; ++q) for (int w = 0; w < 2; ++w)
            mb[q].ll[par][w][me].store(flag_of(epoch, it) << 32 | (val_of(me, it) + w), std::memory_order_relaxed);
        if ((g() & 31) == 0) std::this_thread::yield();
        // step(it): wait for everyone's dots of iteration it
        for (int q = 0; q < W; ++q) for (int w = 0; w < 2; ++w) {
            uint64_t x;
            while (((x = mb[me].ll[par][w][q].load(std::memory_order_relaxed)) >> 32) != flag_of(epoch, it)) std::this_thread::yield();
            if ((uint32_t)x != val_of(q, it) + w) errors++;
        }
        // ... update vectors, push boundary rows into the neighbours' ghost slots, then the halo tag
        for (int q : nb) { mb[q].ghost[me].store((uint64_t)val_of(me, it) + 7, std::memory_order_relaxed);
                           mb[q].halo_tag[me].store(epoch << 32 | (uint64_t)(it + 1), std::memory_order_release); }
    }
}
int main() {
    std::vector<std::thread> th; for (int r = 0; r < W; ++r) th.emplace_back(rank, r);
    for (auto &t : th) t.join();
    std::printf("W=%d, %d iterations: %ld inconsistencies\n", W, ITERS, errors.load());
    return errors.load() != 0;
}
