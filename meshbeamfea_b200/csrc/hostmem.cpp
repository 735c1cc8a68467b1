// Host arrays that travel to the device (plan, solver format, hierarchy): page-locked once a device context exists, so
// that their uploads run at PCIe rate (56 GB/s measured against ~10 GB/s from pageable memory) and asynchronously to
// the host code that builds the next array.  Without a device (host-only taps, CPU tests) they are plain heap memory.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <new>

#include "mbfea_internal.hpp"

namespace mbfea {

namespace {
constexpr uint64_t kHeap = 0x6d62666561686561ULL, kPinned = 0x6d62666561706e64ULL;
constexpr size_t kHeader = 64;              // keeps the payload 64-byte aligned
constexpr size_t kPinFrom = 64 << 10;       // smaller arrays are not worth a driver call
struct Header { uint64_t kind; uint64_t bytes; };
std::atomic<int> g_pin_device{-1};
}  // namespace

void host_arrays_pin_on(int device)
{
    if (const char *e = std::getenv("MBFEA_NO_PIN")) if (*e == '1') return;
    g_pin_device.store(device);
}

void *host_array_alloc(size_t bytes)
{
    void *raw = nullptr;
    uint64_t kind = kHeap;
    const int dev = g_pin_device.load();
    if (dev >= 0 && bytes >= kPinFrom) {
        // helper threads start on device 0: keep the allocation on this process's device (no stray context)
        int cur = -1;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != dev) cudaSetDevice(dev);
        if (cudaHostAlloc(&raw, bytes + kHeader, cudaHostAllocPortable) == cudaSuccess) kind = kPinned;
        else { raw = nullptr; cudaGetLastError(); }
    }
    if (!raw) {
        raw = std::aligned_alloc(kHeader, (bytes + kHeader + kHeader - 1) / kHeader * kHeader);
        if (!raw) throw std::bad_alloc();
    }
    Header *h = static_cast<Header *>(raw);
    h->kind = kind; h->bytes = bytes;
    return static_cast<unsigned char *>(raw) + kHeader;
}

void host_array_free(void *p) noexcept
{
    if (!p) return;
    void *raw = static_cast<unsigned char *>(p) - kHeader;
    if (static_cast<Header *>(raw)->kind == kPinned) cudaFreeHost(raw);
    else std::free(raw);
}

bool host_array_is_pinned(const void *p)
{
    return p && reinterpret_cast<const Header *>(static_cast<const unsigned char *>(p) - kHeader)->kind == kPinned;
}

}  // namespace mbfea
