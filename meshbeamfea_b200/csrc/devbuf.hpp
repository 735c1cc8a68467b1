// Device buffer + error plumbing shared by the host-side translation units of libmbfea.so.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace mbfea {

template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0, cap = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr; n = 0; cap = 0;
    }
    // keeps the allocation when it is large enough (and not wastefully so): a second
    // set_job on a same-sized mesh costs no cudaFree/cudaMalloc round trips
    // would alloc(count) keep the current allocation?
    bool fits(size_t count) const
    {
        if (count == 0) count = 1;
        return p && cap >= count && cap <= 2 * count + 1024;
    }
    cudaError_t alloc(size_t count)
    {
        if (count == 0) count = 1;
        if (fits(count)) { n = count; return cudaSuccess; }
        release();
        cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&p), count * sizeof(T));
        if (e == cudaSuccess) { n = count; cap = count; } else p = nullptr;
        return e;
    }
};

struct CudaFail { cudaError_t e; const char *what; int line; };
struct NcclFail { int e; const char *what; int line; };

#define CK(call)                                                        \
    do {                                                                \
        cudaError_t _e = (call);                                        \
        if (_e != cudaSuccess) throw CudaFail{_e, #call, __LINE__};     \
    } while (0)
#define NK(call)                                                        \
    do {                                                                \
        int _e = (call);                                                \
        if (_e != 0) throw NcclFail{_e, #call, __LINE__};               \
    } while (0)

}  // namespace mbfea
