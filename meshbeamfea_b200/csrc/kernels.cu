// Hand-written sm_100a kernels of libmbfea.so.
//
//   pack_*            input repacking (column-major Eigen images -> 32/64-byte records)
//   K3 free_map       constraint elimination as a vertex renumbering (flag + scan)
//   K1 elem_geom      local frame + stiffness coefficients of every beam (kelem: parity tap)
//   K2 assemble       deterministic scatter-free assembly into 6x6-block BSR
//   K5 block_chol     block-Jacobi as a symmetric scaling: L_r, S_r = L_r^-1 per diagonal block
//      scale_blocks   K~ = S K S^T in the solver layout (symmetric-half interleaved / full rows)
//   K4 spmv_*         y = K~ x with the dot-product partials in the epilogue
//                     (row-thread kernel <SYM, DOT2>; TMA bulk-copy staged kernel)
//   K6 cg_*           CG vector kernels on the scaled system, device-resident scalars,
//                     classic (update/pupdate) and single-reduction (cg2_step) recurrences,
//                     per-batch residual check, residual replacement, batched one-CTA-per-mesh CG
//   K7 element_forces beam end forces in the SimulationResults layout; colour-map pre-pass
//   multi-GPU         NVLink peer-memory mailboxes: halo push, dot all-gather (last-CTA pattern)
//
// Everything is FP64; tensor cores are not used (0.25 flop/byte sparse work).
// Reductions are two-level and fixed-order: per-CTA partials written to a
// buffer, then summed in a fixed order by every consumer CTA, so a solve is
// bit-reproducible run to run for a given grid.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "kernels.cuh"
#include "multilevel_math.hpp"

namespace mbfea { namespace dev {

// ---------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the CTA in a fixed order; result valid in thread 0 (and broadcast
// to all threads when BCAST).  scratch: >= 33 doubles of shared memory.
template <bool BCAST>
__device__ __forceinline__ double block_sum(double v, double *scratch)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();  // scratch may still be read from a previous call
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    double t = 0.0;
    if (w == 0) {
        t = (lane < nw) ? scratch[lane] : 0.0;
        t = warp_sum(t);
        if (BCAST && lane == 0) scratch[32] = t;
    }
    if (BCAST) {
        __syncthreads();
        t = scratch[32];
    }
    return t;
}

// Every CTA sums the same partials in the same order -> identical value.
__device__ __forceinline__ double sum_partials(PartialRef ref, double *scratch)
{
    double s = 0.0;
    for (int i = threadIdx.x; i < ref.n; i += blockDim.x) s += ref.buf[(size_t)i * ref.stride];
    return block_sum<true>(s, scratch);
}

// sums of n (a, b) pairs in a fixed order; one 16-byte load per pair, loads batched by the unroll
__device__ __forceinline__ void sum_partial_pairs(const double2 *pairs, int n, double *scratch, double &sa, double &sb)
{
    double s = 0.0, t = 0.0;
#pragma unroll 8
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double2 v = __ldcg(pairs + i);
        s += v.x; t += v.y;
    }
    sa = block_sum<true>(s, scratch);
    sb = block_sum<true>(t, scratch);
}

// ---- programmatic dependent launch (PDL) ------------------------------------------------------
// (Opt-in, MBFEA_PDL=1.)  The kernels of the CG loop can be launched with programmatic stream serialisation: a kernel lets
// its successor start launching at once (pdl_trigger) and itself waits for its predecessor's
// results with pdl_wait before touching them.  The successor's CTAs become resident as this
// kernel's CTAs retire, so launch latency and ramp-up overlap the previous kernel's tail.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// SM count of the device the context runs on (mbfea_create sets it): grids are sized in resident waves of it
static int g_sm_count = 148;
void set_sm_count(int n) { if (n > 0) g_sm_count = n; }
int sm_count() { return g_sm_count; }

static bool g_use_pdl = false;   // measured on B200 inside CUDA graphs: 304.8 vs 298.9 us/iteration -- no gain, off
void set_pdl(bool on) { g_use_pdl = on; }

template <class... KArgs, class... Args>
static void launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = g_use_pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- peer-memory signalling ---------------------------------------------------
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// flag of the fence-free exchange (Mailbox::ll): never 0, distinct from the slot's previous use (two iterations back)
__device__ __forceinline__ unsigned int ll_flag(unsigned long long epoch, long long iter)
{
    return 0x80000000u | (((unsigned int)epoch & 0x3fu) << 25) | ((unsigned int)(iter + 1) & 0x1ffffffu);
}
__device__ __forceinline__ void ll_store(unsigned long long *lo, unsigned long long *hi, double v, unsigned int flag)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v), f = (unsigned long long)flag << 32;
    st_relaxed_sys(lo, f | (b & 0xffffffffull));
    st_relaxed_sys(hi, f | (b >> 32));
}
// Bounded: a peer that never signals (a crashed rank) must not hang the GPU.  After ~10 s the
// waiter gives up and raises *timeout_flag; the host reports it as a failed solve.
__device__ int g_peer_timeout = 0;
__device__ __forceinline__ void spin_until(const unsigned long long *p, unsigned long long want)
{
    const long long t0 = clock64();
    while (ld_acquire_sys(p) != want) {
        __nanosleep(40);
        if (clock64() - t0 > 20000000000LL) { atomicExch(&g_peer_timeout, 1); break; }
    }
}
__device__ __forceinline__ double ll_wait(const unsigned long long *lo, const unsigned long long *hi, unsigned int flag)
{
    const long long t0 = clock64();
    unsigned long long a, b;
    for (;;) {
        a = ld_relaxed_sys(lo); b = ld_relaxed_sys(hi);
        if ((unsigned int)(a >> 32) == flag && (unsigned int)(b >> 32) == flag) break;
        __nanosleep(20);
        if (clock64() - t0 > 20000000000LL) { atomicExch(&g_peer_timeout, 1); break; }
    }
    return __longlong_as_double((long long)((a & 0xffffffffull) | (b << 32)));
}
__device__ __forceinline__ void trace_stamp(const PcgScalars *sc, long long iter, int slot)
{
    if (sc != nullptr && sc->trace != nullptr && iter < kTraceIters && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        sc->trace[iter * kTraceSlots + slot] = t;
    }
}

// stamp from whichever CTA calls it (thread 0): end-of-kernel events of the last CTA
__device__ __forceinline__ void trace_stamp_cta(const PcgScalars *sc, long long iter, int slot)
{
    if (sc != nullptr && sc->trace != nullptr && iter < kTraceIters && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        sc->trace[iter * kTraceSlots + slot] = t;
    }
}

// true in every thread of exactly one CTA of the grid: the last one to get here.  Call after the CTA's
// global writes; resets the ticket for the next launch.  ONE fence per CTA: the barrier orders every
// thread's writes before thread 0's fence, which is cumulative (the grid-sync pattern; a fence in every
// thread cost the step kernel 5 us per iteration on two GPUs).  sys: some thread of the CTA stored into
// peer memory, so the fence must reach system scope before the ticket is taken.
__device__ __forceinline__ bool last_cta_done(unsigned int *ticket, bool sys)
{
    __shared__ int is_last;
    const int any_sys = __syncthreads_or(sys ? 1 : 0);
    if (threadIdx.x == 0) {
        if (any_sys) __threadfence_system(); else __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
        if (is_last) *ticket = 0u;
    }
    __syncthreads();
    if (is_last) __threadfence();
    return is_last != 0;
}

__device__ __forceinline__ double2 ldg2(const double *p)
{
    return __ldg(reinterpret_cast<const double2 *>(p));
}

// streaming 16-byte load: matrix values are touched once per SpMV
__device__ __forceinline__ double2 ld_stream2(const double *p)
{
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

// ---------------------------------------------------------------------------
// packing
// ---------------------------------------------------------------------------
__global__ void k_pack_nodes(int64_t V, const double *__restrict__ nodes_cm, Node4 *__restrict__ node4)
{
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    Node4 n{nodes_cm[v], nodes_cm[V + v], nodes_cm[2 * V + v], 0.0};
    double2 *o = reinterpret_cast<double2 *>(node4 + v);
    o[0] = make_double2(n.x, n.y);
    o[1] = make_double2(n.z, 0.0);
}

__global__ void k_pack_elems(int64_t E, const int32_t *__restrict__ elems_cm,
                             const double *__restrict__ normals_cm, const float *__restrict__ props4,
                             ElemRec *__restrict__ rec)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    // EA, EIz, EIy, GJ are float values in the reference [REF src/beamsimulator.cpp:173-188]; the widening is exact
    const float4 pf = __ldg(reinterpret_cast<const float4 *>(props4) + e);
    const double2 p01 = make_double2((double)pf.x, (double)pf.y), p23 = make_double2((double)pf.z, (double)pf.w);
    const uint32_t a = (uint32_t)elems_cm[e], b = (uint32_t)elems_cm[E + e];
    const long long ab = (long long)(((unsigned long long)b << 32) | a);
    double2 *o = reinterpret_cast<double2 *>(rec + e);
    o[0] = make_double2(normals_cm[e], normals_cm[E + e]);
    o[1] = make_double2(normals_cm[2 * E + e], p01.x);
    o[2] = make_double2(p01.y, p23.x);
    o[3] = make_double2(p23.y, __longlong_as_double(ab));
}

void launch_pack_nodes(cudaStream_t st, int64_t V, const double *nodes_cm, Node4 *node4)
{
    if (V <= 0) return;
    k_pack_nodes<<<(unsigned)((V + 255) / 256), 256, 0, st>>>(V, nodes_cm, node4);
}

void launch_pack_elems(cudaStream_t st, int64_t E, const int32_t *elems_cm, const double *normals_cm,
                       const float *props4, ElemRec *rec)
{
    if (E <= 0) return;
    k_pack_elems<<<(unsigned)((E + 255) / 256), 256, 0, st>>>(E, elems_cm, normals_cm, props4, rec);
}

// ---------------------------------------------------------------------------
// K3: constraint map.  All six DOFs of a fixed vertex go
// [REF src/beamsimulator.cpp:139-150], so elimination is the vertex map
// free_index[v] = -1 (fixed) | rank of v among free vertices; the reduced DOF
// of (v,d) is 6*free_index[v]+d.  Integer only => bit-exact vs the oracle.
// ---------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;

__global__ void k_fill_i32(int64_t n, int32_t *a, int32_t v)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}

__global__ void k_mark_fixed(int64_t F, const int32_t *__restrict__ fixed, int32_t *flag)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < F) flag[fixed[i]] = 0;  // duplicates write the same value
}

// per-tile exclusive scan of the flags (in place) + tile totals
__global__ void __launch_bounds__(kScanThreads) k_scan_tiles(int64_t V, int32_t *flag_scan, int32_t *tile_sum,
                                                             int32_t *flag_keep)
{
    __shared__ int32_t wsum[kScanThreads / 32];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    int32_t v[kScanItems], s = 0;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
        v[j] = (base + j < V) ? flag_scan[base + j] : 0;
        s += v[j];
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int32_t inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    int32_t woff = 0;
    for (int j = 0; j < w; ++j) woff += wsum[j];
    int32_t run = woff + inc - s;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
        if (base + j < V) {
            flag_keep[base + j] = v[j];
            flag_scan[base + j] = run;
        }
        run += v[j];
    }
    if (threadIdx.x == kScanThreads - 1) tile_sum[blockIdx.x] = woff + inc;
}

// exclusive scan of the tile totals, one CTA, fixed order
__global__ void __launch_bounds__(kScanThreads) k_scan_tile_sums(int64_t ntiles, int32_t *tile_sum)
{
    __shared__ int32_t wsum[kScanThreads / 32];
    __shared__ int32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < ntiles; base += kScanThreads) {
        const int64_t i = base + threadIdx.x;
        const int32_t v = (i < ntiles) ? tile_sum[i] : 0;
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        int32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        int32_t woff = 0;
        for (int j = 0; j < w; ++j) woff += wsum[j];
        const int32_t carry = carry_s;
        if (i < ntiles) tile_sum[i] = carry + woff + inc - v;
        __syncthreads();
        if (threadIdx.x == kScanThreads - 1) carry_s = carry + woff + inc;
        __syncthreads();
    }
}

__global__ void k_scan_finish(int64_t V, int32_t *free_index, const int32_t *__restrict__ tile_sum,
                              const int32_t *__restrict__ flag_keep)
{
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    free_index[v] = flag_keep[v] ? free_index[v] + tile_sum[v / kScanTile] : -1;
}

// free_index[v] = perm[canonical rank] (internal block row), -1 stays -1
__global__ void k_apply_perm(int64_t V, const int32_t *__restrict__ canon, const int32_t *__restrict__ perm,
                             int32_t *__restrict__ out)
{
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const int32_t c = canon[v];
    out[v] = c >= 0 ? __ldg(perm + c) : -1;
}

void launch_apply_perm(cudaStream_t st, int64_t V, const int32_t *canon, const int32_t *perm, int32_t *out)
{
    if (V <= 0) return;
    k_apply_perm<<<(unsigned)((V + 255) / 256), 256, 0, st>>>(V, canon, perm, out);
}

void launch_free_map(cudaStream_t st, int64_t V, int64_t F, const int32_t *fixed, int32_t *free_index,
                     int32_t *scratch)
{
    if (V <= 0) return;
    const int64_t ntiles = (V + kScanTile - 1) / kScanTile;
    int32_t *flag_keep = scratch;          // V ints
    int32_t *tile_sum = scratch + V;       // ntiles ints
    k_fill_i32<<<(unsigned)((V + 255) / 256), 256, 0, st>>>(V, free_index, 1);
    if (F > 0) k_mark_fixed<<<(unsigned)((F + 255) / 256), 256, 0, st>>>(F, fixed, free_index);
    k_scan_tiles<<<(unsigned)ntiles, kScanThreads, 0, st>>>(V, free_index, tile_sum, flag_keep);
    k_scan_tile_sums<<<1, kScanThreads, 0, st>>>(ntiles, tile_sum);
    k_scan_finish<<<(unsigned)((V + 255) / 256), 256, 0, st>>>(V, free_index, tile_sum, flag_keep);
}

// ---------------------------------------------------------------------------
// K1: element stiffness in global axes, one thread per (beam, row of 12).
// Debug/parity tap: the production path never stores K_e (see k_assemble).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(192) k_kelem(int64_t E, const ElemRec *__restrict__ rec,
                                               const Node4 *__restrict__ node4, double *__restrict__ Ke)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t e = t / 12;
    const int r = (int)(t - e * 12);
    if (e >= E) return;
    Frame f; Stiff s; int32_t a, b;
    load_elem(rec, node4, e, f, s, a, b);
    const int half = r / 6, rr = r - 6 * half;
    double lo[6], hi[6];
    kelem_row(f, s, half * 2 + 0, rr, lo);
    kelem_row(f, s, half * 2 + 1, rr, hi);
    double2 *o = reinterpret_cast<double2 *>(Ke + e * 144 + r * 12);
    o[0] = make_double2(lo[0], lo[1]); o[1] = make_double2(lo[2], lo[3]); o[2] = make_double2(lo[4], lo[5]);
    o[3] = make_double2(hi[0], hi[1]); o[4] = make_double2(hi[2], hi[3]); o[5] = make_double2(hi[4], hi[5]);
}

void launch_kelem(cudaStream_t st, int64_t E, const ElemRec *rec, const Node4 *node4, double *Ke)
{
    if (E <= 0) return;
    const int64_t n = E * 12;
    k_kelem<<<(unsigned)((n + 191) / 192), 192, 0, st>>>(E, rec, node4, Ke);
}

// ---------------------------------------------------------------------------
// K1 (production): one thread per beam computes the local frame and the ten
// stiffness coefficients ONCE, register-resident from two 32-byte vertex
// records and one 64-byte beam record (vectorised 16-byte loads), and stores
// them as one 160-byte ElemGeom record.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_elem_geom(int64_t E, const ElemRec *__restrict__ rec,
                                                   const Node4 *__restrict__ node4, ElemGeom *__restrict__ geom)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    Frame f; Stiff s; int32_t a, b;
    load_elem(rec, node4, e, f, s, a, b);
    store_geom(geom, e, f, s, a, b);
}

void launch_elem_geom(cudaStream_t st, int64_t E, const ElemRec *rec, const Node4 *node4, ElemGeom *geom)
{
    if (E <= 0) return;
    k_elem_geom<<<(unsigned)((E + 127) / 128), 128, 0, st>>>(E, rec, node4, geom);
}

// ---------------------------------------------------------------------------
// K2: one thread per OUTPUT block.  The thread walks the block's contributor
// list (element, sub-block), which the host sorted by ascending element index
// -- the order in which the reference's setFromTriplets sums duplicates --
// forms the 6x6 sub-block Q^T S Q from the beam's 160-byte ElemGeom record
// with every row index a compile-time constant (no divergent row dispatch)
// and accumulates it in 36 registers.  No atomics, no scatter: every block is
// written exactly once, by one thread, in a fixed order, so the assembled
// matrix is bit-reproducible and bit-identical to the oracle's.
// (Measured alternatives: 6 lanes per block 1.72 ms -- divergent row dispatch;
// 2 threads per block at 6 CTAs/SM 1.61 ms -- the contributor walk is
// duplicated; this version 0.98 ms on the 100^3 lattice.)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_assemble(int64_t nnzb, const int32_t *__restrict__ contrib_ptr,
                                                  const int32_t *__restrict__ contrib,
                                                  const ElemGeom *__restrict__ geom, double *__restrict__ vals)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double acc[36];
#pragma unroll
    for (int j = 0; j < 36; ++j) acc[j] = 0.0;
    const int32_t c0 = k < nnzb ? __ldg(contrib_ptr + k) : 0, c1 = k < nnzb ? __ldg(contrib_ptr + k + 1) : 0;
    for (int32_t c = c0; c < c1; ++c) {
        const int32_t code = __ldg(contrib + c);
        Frame f; Stiff s;
        load_geom(geom, (int64_t)(code >> 2), f, s);
        const int which = code & 3;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            double row[6];
            kelem_row(f, s, which, r, row);
#pragma unroll
            for (int j = 0; j < 6; ++j) acc[6 * r + j] = ADD(acc[6 * r + j], row[j]);
        }
    }
    // the CTA's 128 blocks are one contiguous 36 KB span of `vals`: stage them in shared memory (row stride 37: no bank
    // conflicts) and store the span with consecutive threads on consecutive 16-byte chunks -- a thread storing its own
    // 288 bytes touches 32 different sectors per store instruction of the warp
    __shared__ double stage[128 * 37];
    if (k < nnzb) {
#pragma unroll
        for (int j = 0; j < 36; ++j) stage[threadIdx.x * 37 + j] = acc[j];
    }
    __syncthreads();
    const int64_t k0 = (int64_t)blockIdx.x * blockDim.x;
    const int64_t live = nnzb - k0 < 128 ? nnzb - k0 : 128;   // blocks of this CTA
    double2 *o = reinterpret_cast<double2 *>(vals + k0 * 36);
#pragma unroll
    for (int j = 0; j < 18; ++j) {
        const int c = threadIdx.x + 128 * j;          // 16-byte chunk of the span
        const int b = c / 18, w = (c - b * 18) * 2;   // block, first of its two entries
        if (b < live) o[c] = make_double2(stage[b * 37 + w], stage[b * 37 + w + 1]);
    }
}

void launch_assemble(cudaStream_t st, int64_t nnzb, const int32_t *contrib_ptr, const int32_t *contrib,
                     const ElemGeom *geom, double *vals)
{
    if (nnzb <= 0) return;
    k_assemble<<<(unsigned)((nnzb + 127) / 128), 128, 0, st>>>(nnzb, contrib_ptr, contrib, geom, vals);
}

// ---------------------------------------------------------------------------
// K5: block-Jacobi as a symmetric scaling.  One thread per block row factors the
// SPD diagonal block D_r = L_r L_r^T (Cholesky, registers) and stores L_r and
// S_r = L_r^{-1}.  The solver then runs plain CG on K~ = S K S^T, whose diagonal
// blocks are the identity -- the same iterates as block-Jacobi PCG on K, without
// reading a 288-byte inverse per row per iteration and without a z vector.
// A non-positive pivot raises *flag (isolated vertex, zero-stiffness beam...).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_block_chol(int64_t nb, const int32_t *__restrict__ diag_idx,
                                                    const double *__restrict__ vals, double *__restrict__ lfac,
                                                    double *__restrict__ sinv, int *flag)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nb) return;
    const double *D = vals + (int64_t)__ldg(diag_idx + r) * 36;
    double L[6][6], S[6][6];
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; j += 2) {
            const double2 v = ldg2(D + 6 * i + j);
            L[i][j] = v.x; L[i][j + 1] = v.y;
        }
    bool bad = false;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        double d = L[j][j];
#pragma unroll
        for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k];
        if (!(d > 0.0) || !isfinite(d)) { bad = true; d = 1.0; }
        const double ljj = sqrt(d), inv = 1.0 / ljj;
        L[j][j] = ljj;
#pragma unroll
        for (int i = j + 1; i < 6; ++i) {
            double v = L[i][j];
#pragma unroll
            for (int k = 0; k < j; ++k) v -= L[i][k] * L[j][k];
            L[i][j] = v * inv;
        }
    }
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = i + 1; j < 6; ++j) L[i][j] = 0.0;
    // S = L^{-1}, lower triangular, column by column
#pragma unroll
    for (int j = 0; j < 6; ++j) {
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            if (i < j) { S[i][j] = 0.0; continue; }
            if (i == j) { S[i][j] = 1.0 / L[j][j]; continue; }
            double v = 0.0;
#pragma unroll
            for (int k = j; k < i; ++k) v -= L[i][k] * S[k][j];
            S[i][j] = v / L[i][i];
        }
    }
    if (bad) atomicOr(flag, 1);
    double2 *ol = reinterpret_cast<double2 *>(lfac + r * 36), *os = reinterpret_cast<double2 *>(sinv + r * 36);
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; j += 2) {
            ol[3 * i + j / 2] = make_double2(L[i][j], L[i][j + 1]);
            os[3 * i + j / 2] = make_double2(S[i][j], S[i][j + 1]);
        }
}

void launch_block_chol(cudaStream_t st, int64_t nb, const int32_t *diag_idx, const double *vals, double *lfac,
                       double *sinv, int *flag)
{
    if (nb <= 0) return;
    k_block_chol<<<(unsigned)((nb + 127) / 128), 128, 0, st>>>(nb, diag_idx, vals, lfac, sinv, flag);
}

// K~_rc = S_r K_rc S_c^T for every block the solver format stores (one thread per block
// row; sinv holds the owned rows followed by the ghost rows).
__global__ void __launch_bounds__(128) k_scale_blocks(int64_t nb, const int32_t *__restrict__ uptr,
                                                      const int32_t *__restrict__ ucol,
                                                      const int32_t *__restrict__ usrc,
                                                      const double *__restrict__ vals,
                                                      const double *__restrict__ sinv, double *__restrict__ vals_s,
                                                      const int32_t *__restrict__ tile_base)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nb) return;
    double Sr[6][6];
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; j += 2) {
            const double2 v = ldg2(sinv + r * 36 + 6 * i + j);
            Sr[i][j] = v.x; Sr[i][j + 1] = v.y;
        }
    for (int32_t b = __ldg(uptr + r); b < __ldg(uptr + r + 1); ++b) {
        const double *K = vals + (int64_t)__ldg(usrc + b) * 36;
        const double *Sc = sinv + (int64_t)__ldg(ucol + b) * 36;
        double T[6][6];
#pragma unroll
        for (int j = 0; j < 6; j += 2) {
            double2 col[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) col[k] = ldg2(K + 6 * k + j);
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                double a = 0.0, c = 0.0;
#pragma unroll
                for (int k = 0; k <= i; ++k) { a = fma(Sr[i][k], col[k].x, a); c = fma(Sr[i][k], col[k].y, c); }
                T[i][j] = a; T[i][j + 1] = c;
            }
        }
        double out[6][6];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            double sc[6];
#pragma unroll
            for (int k = 0; k < 6; k += 2) { const double2 v = ldg2(Sc + 6 * j + k); sc[k] = v.x; sc[k + 1] = v.y; }
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                double a = 0.0;
#pragma unroll
                for (int k = 0; k <= j; ++k) a = fma(T[i][k], sc[k], a);
                out[i][j] = a;
            }
        }
        if (tile_base == nullptr) {   // plain BSR order
            double2 *o = reinterpret_cast<double2 *>(vals_s + (int64_t)b * 36);
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int j = 0; j < 6; j += 2) o[3 * i + j / 2] = make_double2(out[i][j], out[i][j + 1]);
        } else {                      // interleaved per 32-row tile (see HostPlan::tile_base)
            const int64_t k = b - __ldg(uptr + r);
            double2 *o = reinterpret_cast<double2 *>(vals_s) + (int64_t)__ldg(tile_base + (r >> 5)) + k * 576 + (r & 31) * 6;
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int m = 0; m < 3; ++m) o[m * 192 + i] = make_double2(out[i][2 * m], out[i][2 * m + 1]);
        }
    }
}

void launch_scale_blocks(cudaStream_t st, int64_t nb, const int32_t *uptr, const int32_t *ucol, const int32_t *usrc,
                         const double *vals, const double *sinv, double *vals_s, const int32_t *tile_base)
{
    if (nb <= 0) return;
    k_scale_blocks<<<(unsigned)((nb + 127) / 128), 128, 0, st>>>(nb, uptr, ucol, usrc, vals, sinv, vals_s, tile_base);
}

// ---------------------------------------------------------------------------
// K4 (a): row-thread SpMV.  192 threads = 32 block rows x 6 scalar rows; a
// thread owns one scalar row and streams its 48-byte slice of every block its
// row stores (3 x LDG.128), gathering the 48-byte x-block through the read-only
// path.  Consecutive threads read consecutive 48-byte slices, so a warp's
// three loads together cover a contiguous span of the value stream.
//   identity != 0 : a unit diagonal block is implied (scaled solver matrix)
//   lptr != null  : symmetric-half storage -- the row's remaining neighbours are
//                   TRANSPOSED references into other rows' stored blocks (thread
//                   i reads column i of the block: the 6 threads of a row
//                   together read the whole 288-byte block, which the owning
//                   row streamed from HBM moments before/after => L2 hit).
// Epilogue: partial[blockIdx] = sum over the CTA's rows of x_r . y_r.
// Multi-GPU (pc != null): waits for the neighbours' pushed ghost blocks first,
// and the last CTA all-gathers the rank's p.Ap through the NVLink mailboxes.
// ---------------------------------------------------------------------------
constexpr int kRowsPerCta = 32;
constexpr int kSpmvThreads = kRowsPerCta * 6;

// tuning knobs of the symmetric-half variant (A/B-tested on B200, see profiles/)
#ifndef MBFEA_SYM_MINB
#define MBFEA_SYM_MINB 6      // resident CTAs per SM the register allocation is pinned to
#endif
#ifndef MBFEA_SYM_UNROLL
#define MBFEA_SYM_UNROLL 2    // unroll factor of the block loops (0: the compiler's choice, 4; measured on B200:
                              // 1 -> 203.9 us SpMV / 294.4 us per iteration, 2 -> 208.2 / 292.5, 3 -> 276.7 / 351, 0 -> 216 / 302)
#endif
#ifndef MBFEA_SYM_L2PF
#define MBFEA_SYM_L2PF 0      // prefetch.global.L2 the value range of the CTA's tile after next (measured: slower)
#endif
#define MBFEA_PRAGMA_(x) _Pragma(#x)
#define MBFEA_UNROLL_(n) MBFEA_PRAGMA_(unroll n)
#if MBFEA_SYM_UNROLL > 0
#define MBFEA_UNROLL(n) MBFEA_UNROLL_(n)
#else
#define MBFEA_UNROLL(n)
#endif
#ifndef MBFEA_SYM_PREFETCH
#define MBFEA_SYM_PREFETCH 0  // fetch the next tile's row pointers while working on this one
#endif
#ifndef MBFEA_SYM_STREAM
#define MBFEA_SYM_STREAM 0    // value loads that bypass L1 allocation (ld.global.nc.L1::no_allocate): 1 the row's own blocks,
                              // 2 the transposed references too -- the values are read once per CTA, the x-blocks are reused
#endif
__device__ __forceinline__ double2 ldg_stream(const double2 *p)
{
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

// ROWS (32 or 16) block rows per CTA: the symmetric variant runs half-tiles on small systems (a rank's
// share on 8 GPUs is ~4 tiles per resident CTA: whole tiles quantise the makespan to 5 and leave too few
// independent load chains per SM); the interleaved layout is addressed by (row & 31) either way.
template <bool SYM, bool DOT2, int ROWS>
__global__ void __launch_bounds__(ROWS * 6, (SYM ? MBFEA_SYM_MINB : 8) * (32 / ROWS)) k_spmv_rowthread(
    int64_t nb, const int32_t *__restrict__ uptr, const int32_t *__restrict__ ucol,
    const double *__restrict__ vals, const int32_t *__restrict__ lptr, const int32_t *__restrict__ lcol,
    const int32_t *__restrict__ lofs, const int32_t *__restrict__ tile_base, int identity,
    const double *__restrict__ x, double *__restrict__ y, double *__restrict__ partial,
    double *__restrict__ partial2, const PcgScalars *sc, const PeerCtx *pc)
{
    constexpr int kThreads = ROWS * 6;
    __shared__ double red[33];
    __shared__ double ts[SYM ? kThreads * 6 : 1];
    pdl_wait();      // the previous kernel's results (and sc) are visible from here
    pdl_trigger();   // the next kernel may start launching
    if (sc != nullptr && sc->done != 0) return;
    if (sc != nullptr) trace_stamp(sc, sc->iter, 0);
    if (pc != nullptr) {
        // this iteration's ghost x-blocks were pushed into x[nb..) by the neighbours
        const unsigned long long tag = (sc->epoch << 32) | (unsigned long long)(unsigned int)sc->iter;
        if ((int)threadIdx.x < pc->n_neighbors) spin_until(&pc->mine->halo_tag[pc->neighbor[threadIdx.x]], tag);
        __syncthreads();
        trace_stamp(sc, sc->iter, 1);
    }
    const int lr = threadIdx.x / 6, i = threadIdx.x - 6 * lr;
    double dot = 0.0, dot2 = 0.0;
    const int64_t stride = (int64_t)gridDim.x * ROWS;
    // row pointers of the current tile; with MBFEA_SYM_PREFETCH the next tile's are requested before
    // this tile's blocks, so the pointer -> column -> x chain of the next tile is one level shorter
    int32_t kb = 0, ke = 0, lb = 0, le = 0, tb = 0;
    // [pb, pe): 16-byte units of the tile this CTA will process after the next one.  Its whole value
    // range is contiguous (interleaved layout), so the CTA can pull it into L2 a full tile ahead with
    // fire-and-forget prefetches: by the time the loads are issued they are L2 hits, not HBM misses.
    int32_t pb = 0, pe = 0;
    const int64_t ntiles = (nb + ROWS - 1) / ROWS;
    {
        const int64_t row = (int64_t)blockIdx.x * ROWS + lr;
        if (row < nb) {
            kb = __ldg(uptr + row); ke = __ldg(uptr + row + 1);
            if (SYM) { lb = __ldg(lptr + row); le = __ldg(lptr + row + 1); tb = __ldg(tile_base + (row >> 5)); }
        }
        if (SYM && MBFEA_SYM_L2PF) {
            const int64_t t1 = (int64_t)blockIdx.x + gridDim.x;
            if (t1 < ntiles) { pb = __ldg(tile_base + t1); pe = __ldg(tile_base + t1 + 1); }
        }
    }
    for (int64_t r0 = (int64_t)blockIdx.x * ROWS; r0 < nb; r0 += stride) {
        const int64_t row = r0 + lr;
        const bool live = row < nb;
        int32_t nkb = 0, nke = 0, nlb = 0, nle = 0, ntb = 0;
        if (MBFEA_SYM_PREFETCH) {
            const int64_t nrow = row + stride;
            if (nrow < nb) {
                nkb = __ldg(uptr + nrow); nke = __ldg(uptr + nrow + 1);
                if (SYM) { nlb = __ldg(lptr + nrow); nle = __ldg(lptr + nrow + 1); ntb = __ldg(tile_base + (nrow >> 5)); }
            }
        }
        if (SYM && MBFEA_SYM_L2PF) {
            const char *base = reinterpret_cast<const char *>(vals);
            for (int64_t u = (int64_t)pb * 16 + (int64_t)threadIdx.x * 128; u < (int64_t)pe * 16; u += kThreads * 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(base + u));
        }
        double xi = 0.0, acc = 0.0;
        if (live) {
            xi = __ldg(x + row * 6 + i);
            acc = identity ? xi : 0.0;
            // blocks this row stores: thread i streams ROW i of the block, y_r[i] += M[i][:] . x_c
            // SYM: interleaved tile layout -- the CTA's 192 threads read 192 consecutive 16-byte chunks
            const double2 *vt = SYM ? reinterpret_cast<const double2 *>(vals) + tb + (int)(row & 31) * 6 + i : nullptr;
            auto own_block = [&](int32_t k) {
                const int32_t c = __ldg(ucol + k);
                const double *xx = x + (int64_t)c * 6;
                double2 v0, v1, v2;
                if (SYM) {
                    const double2 *v = vt + (int64_t)(k - kb) * 576;
                    if (MBFEA_SYM_STREAM >= 1) { v0 = ldg_stream(v); v1 = ldg_stream(v + 192); v2 = ldg_stream(v + 384); }
                    else { v0 = __ldg(v); v1 = __ldg(v + 192); v2 = __ldg(v + 384); }
                } else {
                    const double *v = vals + (int64_t)k * 36 + i * 6;
                    v0 = ldg2(v); v1 = ldg2(v + 2); v2 = ldg2(v + 4);
                }
                const double2 x0 = ldg2(xx), x1 = ldg2(xx + 2), x2 = ldg2(xx + 4);
                acc = fma(v0.x, x0.x, acc); acc = fma(v0.y, x0.y, acc);
                acc = fma(v1.x, x1.x, acc); acc = fma(v1.y, x1.y, acc);
                acc = fma(v2.x, x2.x, acc); acc = fma(v2.y, x2.y, acc);
            };
            if (SYM) {   // the unroll factor is a knob of the symmetric variant only
MBFEA_UNROLL(MBFEA_SYM_UNROLL)
                for (int32_t k = kb; k < ke; ++k) own_block(k);
            } else {
                for (int32_t k = kb; k < ke; ++k) own_block(k);
            }
        }
        if (SYM) {
            // transposed references: block M = K~(c, r) lives in row c's storage and y_r += M^T x_c.
            // Thread i again streams ROW i of M (coalesced chunks of row c's tile) but needs only the
            // single entry x_c[i]: it accumulates M[i][k] * x_c[i] for all six outputs k, and the
            // six threads of the row exchange their partial vectors once per tile through shared
            // memory.  Half the register-fill traffic of gathering columns.
            double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0, t4 = 0.0, t5 = 0.0;
            if (live) {
MBFEA_UNROLL(MBFEA_SYM_UNROLL)
                for (int32_t l = lb; l < le; ++l) {
                    const double2 *v = reinterpret_cast<const double2 *>(vals) + __ldg(lofs + l) + i;
                    const double xj = __ldg(x + (int64_t)__ldg(lcol + l) * 6 + i);
                    double2 v0, v1, v2;
                    if (MBFEA_SYM_STREAM >= 2) { v0 = ldg_stream(v); v1 = ldg_stream(v + 192); v2 = ldg_stream(v + 384); }
                    else { v0 = __ldg(v); v1 = __ldg(v + 192); v2 = __ldg(v + 384); }
                    t0 = fma(v0.x, xj, t0); t1 = fma(v0.y, xj, t1);
                    t2 = fma(v1.x, xj, t2); t3 = fma(v1.y, xj, t3);
                    t4 = fma(v2.x, xj, t4); t5 = fma(v2.y, xj, t5);
                }
            }
            __syncthreads();  // ts may still be read from the previous tile
            double *mine = ts + threadIdx.x * 6;
            mine[0] = t0; mine[1] = t1; mine[2] = t2; mine[3] = t3; mine[4] = t4; mine[5] = t5;
            __syncthreads();
            const double *grp = ts + lr * 36 + i;   // entry i of the six partial vectors, fixed order
            acc += ((grp[0] + grp[6]) + (grp[12] + grp[18])) + (grp[24] + grp[30]);
        }
        if (live) {
            y[row * 6 + i] = acc;
            dot = fma(xi, acc, dot);
            if (DOT2) dot2 = fma(xi, xi, dot2);
        }
        if (MBFEA_SYM_PREFETCH) {
            kb = nkb; ke = nke; lb = nlb; le = nle; tb = ntb;
        } else {
            const int64_t nrow = row + stride;
            if (nrow < nb) {
                kb = __ldg(uptr + nrow); ke = __ldg(uptr + nrow + 1);
                if (SYM) { lb = __ldg(lptr + nrow); le = __ldg(lptr + nrow + 1); tb = __ldg(tile_base + (nrow >> 5)); }
            }
        }
        if (SYM && MBFEA_SYM_L2PF) {
            const int64_t t2 = (r0 >> 5) + 2 * (int64_t)gridDim.x;   // the tile after next
            pb = pe = 0;
            if (t2 < ntiles) { pb = __ldg(tile_base + t2); pe = __ldg(tile_base + t2 + 1); }
        }
    }
    if (DOT2 && pc != nullptr) {
        // multi-GPU single-reduction CG: both partials in one 16-byte slot of `partial`, so the last
        // CTA's serial reduction below needs half the loads (partial2 is not used on this path)
        const double s = block_sum<false>(dot, red), s2 = block_sum<false>(dot2, red);
        if (threadIdx.x == 0) reinterpret_cast<double2 *>(partial)[blockIdx.x] = make_double2(s, s2);
    } else {
        if (partial != nullptr) {
            const double s = block_sum<false>(dot, red);
            if (threadIdx.x == 0) partial[blockIdx.x] = s;
        }
        if (DOT2) {   // single-reduction CG: x.x rides along with x.Ax
            const double s = block_sum<false>(dot2, red);
            if (threadIdx.x == 0) partial2[blockIdx.x] = s;
        }
    }
    if (pc != nullptr && partial != nullptr && last_cta_done(pc->tickets + 0, false)) {
        // fused reduction + all-gather: the last CTA sums the partials in fixed order and
        // stores this rank's dot product(s) into every rank's mailbox over NVLink
        if (DOT2) trace_stamp_cta(sc, sc->iter, 6);
        double s, s2 = 0.0;
        if (DOT2) sum_partial_pairs(reinterpret_cast<const double2 *>(partial), (int)gridDim.x, red, s, s2);
        else s = sum_partials(PartialRef{partial, (int)gridDim.x, 1}, red);
        if (!DOT2) {
            if ((int)threadIdx.x < pc->world) {
                const unsigned long long tag = (sc->epoch << 32) | (unsigned long long)(unsigned int)sc->iter;
                Mailbox *m = pc->peer[threadIdx.x];
                m->dot_val[0][pc->rank] = s;
                st_release_sys(&m->dot_tag[0][pc->rank], tag + 1);   // release orders this thread's value store first
            }
        } else if ((int)threadIdx.x < 2 * pc->world) {
            // single-reduction CG: fence-free words, slots double-buffered by iteration parity;
            // thread q sends r.K~r to rank q, thread world + q sends r.r
            const int q = (int)threadIdx.x % pc->world, which = (int)threadIdx.x / pc->world;
            const int par = (int)(sc->iter & 1);
            Mailbox *m = pc->peer[q];
            ll_store(&m->ll[par][2 * which][pc->rank], &m->ll[par][2 * which + 1][pc->rank], which ? s2 : s,
                     ll_flag(sc->epoch, sc->iter));
        }
        if (DOT2) trace_stamp_cta(sc, sc->iter, 4);
    }
}

// rows per CTA: half-tiles when whole tiles would give a resident CTA fewer than kHalfTileBelow tiles
static int spmv_rows(int64_t nb, bool sym, bool peer)
{
    static const int forced = [] { const char *e = std::getenv("MBFEA_SPMV_ROWS"); return e ? std::atoi(e) : 0; }();
    if (!sym) return 32;
    // NVLink peer path: the last CTA reduces one partial per CTA serially -- twice the CTAs cost more there
    // (7.6 vs 2.8 us on two GPUs) than the finer tiles gain (and the paired partials fill `partial` at 148 * 8)
    if (peer) return 32;
    if (forced == 16 || forced == 32) return forced;
    constexpr int64_t kHalfTileBelow = 12;
    return (nb + 31) / 32 < (int64_t)g_sm_count * MBFEA_SYM_MINB * kHalfTileBelow ? 16 : 32;
}

// one full wave: SMs x resident CTAs per SM (8 at 40 registers, 6 at 56 for the symmetric variant)
int spmv_rowthread_grid(int64_t nb, bool sym, bool peer)
{
    const int rows = spmv_rows(nb, sym, peer);
    const int64_t tiles = (nb + rows - 1) / rows;
    int64_t cap = (int64_t)(sym ? g_sm_count * MBFEA_SYM_MINB : g_sm_count * 8) * (32 / rows);
    if (cap > kMaxPartials) cap = kMaxPartials;
    return (int)(tiles < 1 ? 1 : (tiles > cap ? cap : tiles));
}

void launch_spmv_rowthread(cudaStream_t st, int64_t nb, const int32_t *uptr, const int32_t *ucol, const double *vals,
                           const int32_t *lptr, const int32_t *lcol, const int32_t *lofs, const int32_t *tile_base,
                           int identity, const double *x, double *y, double *partial, double *partial2,
                           const PcgScalars *sc, const PeerCtx *pc)
{
    const int gs = spmv_rowthread_grid(nb, lptr != nullptr, pc != nullptr);
    const bool half_tiles = spmv_rows(nb, lptr != nullptr, pc != nullptr) == 16;
    const int32_t *null32 = nullptr;
    double *nulld = nullptr;
    if (lptr != nullptr && partial2 != nullptr && half_tiles)
        launch_pdl(k_spmv_rowthread<true, true, 16>, gs, 96, 0, st, nb, uptr, ucol, vals, lptr, lcol, lofs, tile_base,
                   identity, x, y, partial, partial2, sc, pc);
    else if (lptr != nullptr && partial2 != nullptr)
        launch_pdl(k_spmv_rowthread<true, true, 32>, gs, 192, 0, st, nb, uptr, ucol, vals, lptr, lcol, lofs, tile_base,
                   identity, x, y, partial, partial2, sc, pc);
    else if (lptr != nullptr && half_tiles)
        launch_pdl(k_spmv_rowthread<true, false, 16>, gs, 96, 0, st, nb, uptr, ucol, vals, lptr, lcol, lofs, tile_base,
                   identity, x, y, partial, nulld, sc, pc);
    else if (lptr != nullptr)
        launch_pdl(k_spmv_rowthread<true, false, 32>, gs, 192, 0, st, nb, uptr, ucol, vals, lptr, lcol, lofs, tile_base,
                   identity, x, y, partial, nulld, sc, pc);
    else if (partial2 != nullptr)
        launch_pdl(k_spmv_rowthread<false, true, 32>, gs, 192, 0, st, nb, uptr, ucol, vals, null32, null32, null32, null32,
                   identity, x, y, partial, partial2, sc, pc);
    else
        launch_pdl(k_spmv_rowthread<false, false, 32>, gs, 192, 0, st, nb, uptr, ucol, vals, null32, null32, null32, null32,
                   identity, x, y, partial, nulld, sc, pc);
}

// ---------------------------------------------------------------------------
// K4 (b): TMA-staged SpMV.  The values of a tile of consecutive block rows are
// ONE contiguous span of the value stream, so a single elected producer thread
// moves each tile with one cp.async.bulk (1-D TMA, SASS UBLKCP) into a ring of
// kTmaStages shared-memory stages, completion signalled on an mbarrier.  Six
// consumer warps (one thread per scalar row) read their 48-byte slices from
// shared memory, gather x through the read-only path, and hand the stage back
// through a second mbarrier.  Persistent CTAs, static tile -> CTA assignment
// (tile t to CTA t mod grid), so the partial dot products are reproducible.
// ---------------------------------------------------------------------------
constexpr int kTmaConsumers = kTmaMaxRows * 6;       // 192
constexpr int kTmaThreads = kTmaConsumers + 32;      // + producer warp
constexpr int kTmaStageBytes = kTmaCapBlocks * 288;  // 64512

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

__global__ void __launch_bounds__(kTmaThreads, 1) k_spmv_tma(
    int64_t ntiles, const int32_t *__restrict__ tile_row, const int32_t *__restrict__ rowptr,
    const int32_t *__restrict__ colidx, const double *__restrict__ vals, int identity,
    const double *__restrict__ x, double *__restrict__ y, double *__restrict__ partial, const PcgScalars *sc)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full_bar[kTmaStages], empty_bar[kTmaStages];
    __shared__ double red[33];
    if (sc != nullptr && sc->done != 0) return;
    double *stage_buf = reinterpret_cast<double *>(smem_raw);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kTmaStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kTmaConsumers / 32);  // one arrive per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5;
    double dot = 0.0;
    if (warp == kTmaConsumers / 32) {
        // ---- producer warp: one elected lane issues the bulk copies ----
        if ((threadIdx.x & 31) == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                const int32_t r0 = __ldg(tile_row + t), r1 = __ldg(tile_row + t + 1);
                const int32_t kb = __ldg(rowptr + r0), ke = __ldg(rowptr + r1);
                const uint32_t bytes = (uint32_t)(ke - kb) * 288u;
                if (bytes > 0) {
                    mbar_expect_tx(&full_bar[stage], bytes);
                    tma_bulk_g2s(stage_buf + (size_t)stage * (kTmaStageBytes / 8), vals + (int64_t)kb * 36, bytes,
                                 &full_bar[stage]);
                } else {
                    mbar_arrive(&full_bar[stage]);
                }
                if (++stage == kTmaStages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ---- consumers: thread = (local block row, scalar row) ----
        const int lr = threadIdx.x / 6, i = threadIdx.x - 6 * lr;
        int stage = 0;
        uint32_t phase = 0;
        for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const int32_t r0 = __ldg(tile_row + t), r1 = __ldg(tile_row + t + 1);
            const int32_t row = r0 + lr;
            const bool live = row < r1;
            int32_t kb = 0, ke = 0, k0 = 0;
            if (live) { kb = __ldg(rowptr + row); ke = __ldg(rowptr + row + 1); }
            k0 = __ldg(rowptr + r0);
            mbar_wait(&full_bar[stage], phase);
            if (live) {
                const double *sv = stage_buf + (size_t)stage * (kTmaStageBytes / 8) + (size_t)(kb - k0) * 36 + i * 6;
                const double xi = __ldg(x + (int64_t)row * 6 + i);
                double acc = identity ? xi : 0.0;
                for (int32_t k = kb; k < ke; ++k, sv += 36) {
                    const int32_t c = __ldg(colidx + k);
                    const double *xx = x + (int64_t)c * 6;
                    const double2 x0 = ldg2(xx), x1 = ldg2(xx + 2), x2 = ldg2(xx + 4);
                    const double2 v0 = *reinterpret_cast<const double2 *>(sv);
                    const double2 v1 = *reinterpret_cast<const double2 *>(sv + 2);
                    const double2 v2 = *reinterpret_cast<const double2 *>(sv + 4);
                    acc = fma(v0.x, x0.x, acc); acc = fma(v0.y, x0.y, acc);
                    acc = fma(v1.x, x1.x, acc); acc = fma(v1.y, x1.y, acc);
                    acc = fma(v2.x, x2.x, acc); acc = fma(v2.y, x2.y, acc);
                }
                y[(int64_t)row * 6 + i] = acc;
                dot = fma(xi, acc, dot);
            }
            __syncwarp();
            if ((threadIdx.x & 31) == 0) mbar_arrive(&empty_bar[stage]);
            if (++stage == kTmaStages) { stage = 0; phase ^= 1; }
        }
    }
    if (partial != nullptr) {
        const double s = block_sum<false>(dot, red);
        if (threadIdx.x == 0) partial[blockIdx.x] = s;
    }
}

int spmv_tma_grid(int64_t ntiles, int sm_count)
{
    int64_t g = ntiles < sm_count ? ntiles : sm_count;
    if (g > kMaxPartials) g = kMaxPartials;
    return (int)(g < 1 ? 1 : g);
}

size_t spmv_tma_smem_bytes() { return (size_t)kTmaStages * kTmaStageBytes; }

int spmv_tma_prepare()
{
    return (int)cudaFuncSetAttribute(k_spmv_tma, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)spmv_tma_smem_bytes());
}

void launch_spmv_tma(cudaStream_t st, int grid, int64_t ntiles, const int32_t *tile_row,
                     const int32_t *rowptr, const int32_t *colidx, const double *vals, int identity,
                     const double *x, double *y, double *partial, const PcgScalars *sc)
{
    k_spmv_tma<<<grid, kTmaThreads, spmv_tma_smem_bytes(), st>>>(ntiles, tile_row, rowptr, colidx, vals, identity, x,
                                                                 y, partial, sc);
}

// ---------------------------------------------------------------------------
// K6: CG vector kernels on the scaled system (unit block diagonal => no
// preconditioner apply, no z vector): 9 vector passes per iteration.  All
// scalars (alpha, beta, iteration count, stop flag) live on the device so a whole
// batch of iterations is one CUDA graph with no host sync.  The stop test
// ||f - K x||_2 <= tol ||f||_2 is evaluated on the UNSCALED residual r = L r~
// once per batch (k_cg_check), so the per-iteration kernels never read L.
// ---------------------------------------------------------------------------
int vec_grid(int64_t nb)
{
    const int64_t tiles = (nb + kRowsPerCta - 1) / kRowsPerCta;
    return (int)(tiles < 1 ? 1 : (tiles > kMaxPartials ? kMaxPartials : tiles));
}

// out_i = sum_{j<=i} M[row][i][j] * v[j]   (M lower triangular, full 6x6 storage)
__device__ __forceinline__ double lower_row_times(const double *__restrict__ M, int64_t row, int i, const double *v)
{
    const double *m = M + row * 36 + i * 6;
    const double2 m0 = ldg2(m), m1 = ldg2(m + 2), m2 = ldg2(m + 4);
    double s = m0.x * v[0];
    s = fma(m0.y, v[1], s); s = fma(m1.x, v[2], s); s = fma(m1.y, v[3], s);
    s = fma(m2.x, v[4], s); s = fma(m2.y, v[5], s);
    return s;  // entries above the diagonal are stored as exact zeros
}

// f~ = S f; y = 0; r~ = f~; p = f~; partials of rho = r~.r~ and of ||f||^2
__global__ void __launch_bounds__(kSpmvThreads) k_cg_init(
    int64_t nb, const double *__restrict__ f, const double *__restrict__ sinv, double *__restrict__ y,
    double *__restrict__ rt, double *__restrict__ ft, double *__restrict__ p, double *__restrict__ part_rho,
    double *__restrict__ part_ff)
{
    __shared__ double vs[kSpmvThreads];
    __shared__ double red[33];
    const int lr = threadIdx.x / 6, i = threadIdx.x - 6 * lr;
    double rho = 0.0, ff = 0.0;
    for (int64_t r0 = (int64_t)blockIdx.x * kRowsPerCta; r0 < nb; r0 += (int64_t)gridDim.x * kRowsPerCta) {
        const int64_t row = r0 + lr;
        const bool live = row < nb;
        const double fv = live ? f[row * 6 + i] : 0.0;
        __syncthreads();
        vs[threadIdx.x] = fv;
        __syncthreads();
        if (live) {
            const double t = lower_row_times(sinv, row, i, vs + lr * 6);
            y[row * 6 + i] = 0.0;
            rt[row * 6 + i] = t;
            ft[row * 6 + i] = t;
            p[row * 6 + i] = t;
            rho = fma(t, t, rho);
            ff = fma(fv, fv, ff);
        }
    }
    const double a = block_sum<false>(rho, red);
    const double b = block_sum<false>(ff, red);
    if (threadIdx.x == 0) { part_rho[blockIdx.x] = a; part_ff[blockIdx.x] = b; }
}

__global__ void __launch_bounds__(256) k_cg_init_finish(PartialRef rho, PartialRef ff, PcgScalars *sc, double tol,
                                                        long long max_iter, unsigned long long epoch,
                                                        unsigned long long *trace)
{
    __shared__ double red[33];
    const double a = sum_partials(rho, red);
    const double b = sum_partials(ff, red);
    if (threadIdx.x == 0) {
        sc->rz[0] = a; sc->rz[1] = 0.0; sc->rho0 = a;
        sc->ff = b; sc->rr = b; sc->pAp = 1.0;
        sc->iter = 0; sc->max_iter = max_iter; sc->tol2 = tol * tol;
        // f == 0 -> u == 0 is the solution; a non-finite load is a breakdown
        sc->done = (b == 0.0) ? 1 : ((isfinite(a) && isfinite(b)) ? 0 : 3);
        sc->pad = 0;
        sc->epoch = epoch;
        sc->tag_cur = 0;
        sc->checks = 0;
        sc->restart_iter = 0;
        sc->trace = trace;
    }
}

__global__ void __launch_bounds__(256, 4) k_cg_update(
    int64_t n, int parity, PartialRef pAp_ref, const double *__restrict__ p, const double *__restrict__ Ap,
    double *__restrict__ y, double *__restrict__ rt, double *__restrict__ part_rho, PcgScalars *sc,
    const PeerCtx *pc, int publish_rho)
{
    __shared__ double red[33];
    pdl_wait();      // the previous kernel's results (and sc) are visible from here
    pdl_trigger();   // the next kernel may start launching
    if (sc->done != 0) return;
    trace_stamp(sc, sc->iter, 2);
    double pAp;
    unsigned long long tag = 0;
    if (pc != nullptr) {
        // every rank's p.Ap arrives in my mailbox; summed in rank order => same bits on all ranks
        tag = ((sc->epoch << 32) | (unsigned long long)(unsigned int)sc->iter) + 1;
        if ((int)threadIdx.x < pc->world) spin_until(&pc->mine->dot_tag[0][threadIdx.x], tag);
        __syncthreads();
        trace_stamp(sc, sc->iter, 3);
        pAp = 0.0;
        for (int q = 0; q < pc->world; ++q) pAp += __ldcv(&pc->mine->dot_val[0][q]);
    } else {
        pAp = sum_partials(pAp_ref, red);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { sc->pAp = pAp; sc->tag_cur = tag; }
    if (!(pAp > 0.0) || !isfinite(pAp)) return;  // breakdown: k_cg_pupdate raises the flag
    const double alpha = sc->rz[parity] / pAp;
    double rho = 0.0;
    const int64_t n2 = n >> 1;  // n = 6*nb is even
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n2; j += (int64_t)gridDim.x * blockDim.x) {
        const double2 pv = reinterpret_cast<const double2 *>(p)[j], av = reinterpret_cast<const double2 *>(Ap)[j];
        double2 yv = reinterpret_cast<double2 *>(y)[j], rv = reinterpret_cast<double2 *>(rt)[j];
        yv.x = fma(alpha, pv.x, yv.x); yv.y = fma(alpha, pv.y, yv.y);
        rv.x = fma(-alpha, av.x, rv.x); rv.y = fma(-alpha, av.y, rv.y);
        reinterpret_cast<double2 *>(y)[j] = yv;
        reinterpret_cast<double2 *>(rt)[j] = rv;
        rho = fma(rv.x, rv.x, rho); rho = fma(rv.y, rv.y, rho);
    }
    const double a = block_sum<false>(rho, red);
    if (threadIdx.x == 0) part_rho[blockIdx.x] = a;
    // (with a preconditioner rho = r~.z~ comes from its last kernel, k_ml_finish0, which publishes it instead)
    if (pc != nullptr && publish_rho && last_cta_done(pc->tickets + 1, false)) {
        const double sa = sum_partials(PartialRef{part_rho, (int)gridDim.x, 1}, red);
        if ((int)threadIdx.x < pc->world) {
            Mailbox *m = pc->peer[threadIdx.x];
            m->dot_val[1][pc->rank] = sa;
            st_release_sys(&m->dot_tag[1][pc->rank], tag);
        }
    }
}

// Multilevel preconditioner, last kernel of the apply (fine prolongation + dot): z~_v = r~_v + T_v z1[agg v] with
// T_v = L_v^T B(d_v) in FP32 (see mlprec.cu); part[blockIdx] = sum of r~ . z~ over the CTA's entries.
// Thread = (row, scalar i): reads row i of T_v (24 B, consecutive threads -> consecutive addresses).
// Multi-GPU (pc != null, inside the CG iteration): the last CTA sums the partials and all-gathers this rank's
// r~.z~ through the NVLink mailboxes -- the slot and tag k_cg_update would have used for r~.r~.
template <bool H16>
__global__ void __launch_bounds__(256, 4) k_ml_finish0(int64_t nb, const int32_t *__restrict__ agg, const void *__restrict__ t0,
                                                       const double *__restrict__ rt, const double *__restrict__ z1,
                                                       double *__restrict__ zt, double *__restrict__ part,
                                                       const PcgScalars *sc, const PeerCtx *pc, double theta)
{
    __shared__ double red[33];
    if (sc != nullptr && sc->done != 0) return;
    double dot = 0.0;
    const int64_t n = 6 * nb;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = t / 6;
        const double *zc = z1 + (int64_t)__ldg(agg + v) * 6;
        float tr[6];
        double tsc;
        tblock_load_row<H16>(t0, v, (int)(t - v * 6), tr, tsc);
        const double ri = rt[t];
        double s = (double)tr[0] * zc[0];
        s = fma((double)tr[1], zc[1], s);
        s = fma((double)tr[2], zc[2], s); s = fma((double)tr[3], zc[3], s);
        s = fma((double)tr[4], zc[4], s); s = fma((double)tr[5], zc[5], s);
        s = fma(theta * tsc, s, ri);   // theta: weight of the coarse correction against the block-Jacobi term
        zt[t] = s;
        dot = fma(ri, s, dot);
    }
    const double a = block_sum<false>(dot, red);
    if (threadIdx.x == 0) part[blockIdx.x] = a;
    if (pc != nullptr && last_cta_done(pc->tickets + 1, false)) {
        const double sa = sum_partials(PartialRef{part, (int)gridDim.x, 1}, red);
        if ((int)threadIdx.x < pc->world) {
            Mailbox *m = pc->peer[threadIdx.x];
            m->dot_val[1][pc->rank] = sa;
            st_release_sys(&m->dot_tag[1][pc->rank], sc->tag_cur);
        }
    }
}

void launch_ml_finish0(cudaStream_t st, int grid, int64_t nb, const int32_t *agg, const void *t0, bool half, const double *rt,
                       const double *z1, double *zt, double *part, const PcgScalars *sc, const PeerCtx *pc, double theta)
{
    if (nb <= 0 || grid <= 0) return;
    if (half) k_ml_finish0<true><<<(unsigned)grid, 256, 0, st>>>(nb, agg, t0, rt, z1, zt, part, sc, pc, theta);
    else k_ml_finish0<false><<<(unsigned)grid, 256, 0, st>>>(nb, agg, t0, rt, z1, zt, part, sc, pc, theta);
}

// p = r~ + beta p.  Multi-GPU: the thread that produces a boundary block's entries also stores
// them straight into the neighbours' ghost regions (push_idx[row] = first entry of the row in
// push_dst | count << 28, or -1), and the last CTA publishes the halo tag -- the halo exchange
// costs no extra kernel and its NVLink stores overlap the rest of the update.
__global__ void __launch_bounds__(256, 4) k_cg_pupdate(int64_t n, int parity, PartialRef rho_ref,
                                                       const double *__restrict__ rt, double *__restrict__ p,
                                                       PcgScalars *sc, const PeerCtx *pc,
                                                       const int32_t *__restrict__ push_idx,
                                                       double *const *__restrict__ push_dst, unsigned int *ticket)
{
    __shared__ double red[33];
    pdl_wait();      // the previous kernel's results (and sc) are visible from here
    pdl_trigger();   // the next kernel may start launching
    // No CTA of this grid writes a scalar another CTA reads at entry: iter / done / rz are advanced by the
    // LAST CTA to retire (ticket), when every other CTA has finished -- no assumption about co-residency.
    if (sc->done != 0) return;
    const double pAp = sc->pAp;
    if (!(pAp > 0.0) || !isfinite(pAp)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) sc->done = 3;
        return;
    }
    const long long it0 = sc->iter;   // advanced only by the last CTA to retire
    trace_stamp(sc, it0, 4);
    double rho_new;
    if (pc != nullptr) {
        const unsigned long long tag = sc->tag_cur;   // written by k_cg_update, stable here
        if ((int)threadIdx.x < pc->world) spin_until(&pc->mine->dot_tag[1][threadIdx.x], tag);
        __syncthreads();
        trace_stamp(sc, it0, 5);
        rho_new = 0.0;
        for (int q = 0; q < pc->world; ++q) rho_new += __ldcv(&pc->mine->dot_val[1][q]);
    } else {
        rho_new = sum_partials(rho_ref, red);
    }
    const double beta = rho_new / sc->rz[parity];
    const int64_t n2 = n >> 1;
    bool pushed = false;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n2; j += (int64_t)gridDim.x * blockDim.x) {
        const double2 rv = reinterpret_cast<const double2 *>(rt)[j];
        double2 pv = reinterpret_cast<double2 *>(p)[j];
        pv.x = fma(beta, pv.x, rv.x); pv.y = fma(beta, pv.y, rv.y);
        reinterpret_cast<double2 *>(p)[j] = pv;
        if (push_idx != nullptr) {
            const int64_t row = j / 3;
            const int32_t code = __ldg(push_idx + row);
            if (code >= 0) {
                const int off = (int)(j - row * 3) * 2;
                const int32_t first = code & 0x0fffffff, cnt = code >> 28;
                for (int32_t q = 0; q < cnt; ++q) *reinterpret_cast<double2 *>(push_dst[first + q] + off) = pv;
                pushed = true;
            }
        }
    }
    if (last_cta_done(ticket, pushed)) {   // peer stores land (system-scope fence) before the ticket
        if (push_idx != nullptr && (int)threadIdx.x < pc->n_neighbors) {
            // tag_cur = (epoch << 32 | k) + 1: exactly the halo tag iteration k+1 waits for
            st_release_sys(&pc->peer[pc->neighbor[threadIdx.x]]->halo_tag[pc->rank], sc->tag_cur);
        }
        if (threadIdx.x == 0) {
            sc->rz[parity ^ 1] = rho_new;
            const long long it = sc->iter + 1;
            sc->iter = it;
            int done = 0;
            if (!isfinite(rho_new)) done = 3;
            else if (rho_new <= 1e-32 * sc->rho0) done = 1;   // below anything FP64 can resolve: exact convergence
            else if (it >= sc->max_iter) done = 2;
            sc->done = done;
        }
    }
}

// ---- single-reduction CG (Chronopoulos-Gear): ONE exchange per iteration ----------------
// The SpMV runs on the residual: w = K~ r, and its epilogue yields both gamma = r.r and
// delta = r.w.  One elementwise kernel then does everything else:
//   beta = gamma_i / gamma_{i-1};  alpha = gamma_i / (delta_i - beta gamma_i / alpha_{i-1})
//   d = r + beta d;  s = w + beta s (= K~ d);  y += alpha d;  r -= alpha s
// Two kernels and one cross-GPU exchange per iteration instead of three and two.
__global__ void __launch_bounds__(kSpmvThreads) k_cg2_init(int64_t nb, const double *__restrict__ f,
                                                           const double *__restrict__ sinv, double *__restrict__ y,
                                                           double *__restrict__ r, double *__restrict__ ft,
                                                           double *__restrict__ d, double *__restrict__ s,
                                                           double *__restrict__ part_ff)
{
    __shared__ double vs[kSpmvThreads];
    __shared__ double red[33];
    const int lr = threadIdx.x / 6, i = threadIdx.x - 6 * lr;
    double ff = 0.0;
    for (int64_t r0 = (int64_t)blockIdx.x * kRowsPerCta; r0 < nb; r0 += (int64_t)gridDim.x * kRowsPerCta) {
        const int64_t row = r0 + lr;
        const bool live = row < nb;
        const double fv = live ? f[row * 6 + i] : 0.0;
        __syncthreads();
        vs[threadIdx.x] = fv;
        __syncthreads();
        if (live) {
            const double t = lower_row_times(sinv, row, i, vs + lr * 6);
            y[row * 6 + i] = 0.0; r[row * 6 + i] = t; ft[row * 6 + i] = t;
            d[row * 6 + i] = 0.0; s[row * 6 + i] = 0.0;
            ff = fma(fv, fv, ff);
        }
    }
    const double b = block_sum<false>(ff, red);
    if (threadIdx.x == 0) part_ff[blockIdx.x] = b;
}

__global__ void __launch_bounds__(256, 4) k_cg2_step(
    int64_t n, PartialRef delta_ref, PartialRef gamma_ref, double *__restrict__ r, const double *__restrict__ w,
    double *__restrict__ d, double *__restrict__ s, double *__restrict__ y, PcgScalars *sc, const PeerCtx *pc,
    const int32_t *__restrict__ push_idx, double *const *__restrict__ push_dst,
    const int32_t *__restrict__ push_rows, int n_push, unsigned int *ticket)
{
    __shared__ double red[33];
    __shared__ double dots[2][kMaxWorld];
    pdl_wait();      // the previous kernel's results (and sc) are visible from here
    pdl_trigger();   // the next kernel may start launching
    if (sc->done != 0) return;
    const long long it = sc->iter;    // stable: advanced only by the last CTA to retire (ticket), after every CTA's reads
    trace_stamp(sc, it, 2);
    double delta, gamma;
    unsigned long long tag = 0;
    if (pc != nullptr) {
        tag = ((sc->epoch << 32) | (unsigned long long)(unsigned int)it) + 1;
        const int mp = (int)(it & 1);
        if ((int)threadIdx.x < 2 * pc->world) {   // thread q waits for rank q's r.K~r, thread world + q for its r.r
            const int q = (int)threadIdx.x % pc->world, which = (int)threadIdx.x / pc->world;
            const Mailbox *m = pc->mine;
            dots[which][q] = ll_wait(&m->ll[mp][2 * which][q], &m->ll[mp][2 * which + 1][q], ll_flag(sc->epoch, it));
        }
        __syncthreads();
        trace_stamp(sc, it, 3);
        delta = 0.0; gamma = 0.0;
        for (int q = 0; q < pc->world; ++q) { delta += dots[0][q]; gamma += dots[1][q]; }   // rank order: same bits everywhere
    } else {
        delta = sum_partials(delta_ref, red);
        gamma = sum_partials(gamma_ref, red);
    }
    const int par = (int)(it & 1);
    double beta = 0.0, den = delta;
    if (it > sc->restart_iter) {
        beta = gamma / sc->rz[par ^ 1];              // gamma_{i-1}
        den = delta - beta * gamma / sc->alpha[par ^ 1];
    }
    const bool exact = (it == 0) ? (gamma == 0.0) : (gamma <= 1e-32 * sc->rho0);
    // (rho0 is the very first gamma of the solve, kept across restarts)
    const bool bad = !exact && (!(den > 0.0) || !isfinite(den) || !isfinite(gamma));
    const double alpha = (exact || bad) ? 0.0 : gamma / den;
    bool pushed = false;
    if (!exact && !bad) {
        // d = r + beta d; s = w + beta s; y += alpha d; r -= alpha s on 16-byte chunk j; returns the new r
        auto step = [&](int64_t j) {
            double2 rv = reinterpret_cast<double2 *>(r)[j];
            const double2 wv = reinterpret_cast<const double2 *>(w)[j];
            double2 dv = reinterpret_cast<double2 *>(d)[j], sv = reinterpret_cast<double2 *>(s)[j];
            double2 yv = reinterpret_cast<double2 *>(y)[j];
            dv.x = fma(beta, dv.x, rv.x); dv.y = fma(beta, dv.y, rv.y);
            sv.x = fma(beta, sv.x, wv.x); sv.y = fma(beta, sv.y, wv.y);
            yv.x = fma(alpha, dv.x, yv.x); yv.y = fma(alpha, dv.y, yv.y);
            rv.x = fma(-alpha, sv.x, rv.x); rv.y = fma(-alpha, sv.y, rv.y);
            reinterpret_cast<double2 *>(d)[j] = dv;
            reinterpret_cast<double2 *>(s)[j] = sv;
            reinterpret_cast<double2 *>(y)[j] = yv;
            reinterpret_cast<double2 *>(r)[j] = rv;
            return rv;
        };
        const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
        // boundary blocks first: their stores into the neighbours' ghost regions cross NVLink while the
        // interior is updated, so the system-scope fence below finds them already acknowledged
        for (int64_t k = tid; k < 3 * (int64_t)n_push; k += nthr) {
            const int64_t kr = k / 3;
            const int32_t row = __ldg(push_rows + kr);
            const int off = (int)(k - kr * 3);
            const double2 rv = step((int64_t)row * 3 + off);
            const int32_t code = __ldg(push_idx + row);
            const int32_t first = code & 0x0fffffff, cnt = code >> 28;
            for (int32_t q = 0; q < cnt; ++q) *reinterpret_cast<double2 *>(push_dst[first + q] + 2 * off) = rv;
            pushed = true;
        }
        const int64_t n2 = n >> 1;
        for (int64_t j = tid; j < n2; j += nthr) {
            if (push_idx != nullptr && __ldg(push_idx + j / 3) >= 0) continue;   // done above
            step(j);
        }
    }
    if (last_cta_done(ticket, pushed)) {
        if (push_idx != nullptr) {   // the halo tag must be published even when nothing moved (peers wait on it)
            if ((int)threadIdx.x < pc->n_neighbors) {
                st_release_sys(&pc->peer[pc->neighbor[threadIdx.x]]->halo_tag[pc->rank], tag);
            }
            trace_stamp_cta(sc, it, 5);
        }
        if (threadIdx.x == 0) {
            sc->rz[par] = gamma;
            sc->alpha[par] = alpha;
            sc->pAp = den;
            if (it == 0) sc->rho0 = gamma;
            sc->iter = it + 1;
            int done = 0;
            if (bad) done = 3;
            else if (exact) done = 1;
            else if (it + 1 >= sc->max_iter) done = 2;
            sc->done = done;
        }
    }
}

void launch_cg2_init(cudaStream_t st, int64_t nb, const double *f, const double *sinv, double *y, double *r, double *ft,
                     double *d, double *s, double *part_ff)
{
    k_cg2_init<<<vec_grid(nb), kSpmvThreads, 0, st>>>(nb, f, sinv, y, r, ft, d, s, part_ff);
}

void launch_cg2_step(cudaStream_t st, int64_t nb, PartialRef delta, PartialRef gamma, double *r, const double *w,
                     double *d, double *s, double *y, PcgScalars *sc, const PeerCtx *pc, const int32_t *push_idx,
                     double *const *push_dst, const int32_t *push_rows, int n_push, unsigned int *ticket)
{
    launch_pdl(k_cg2_step, cg_vec_grid(nb), 256, 0, st, 6 * nb, delta, gamma, r, w, d, s, y, sc, pc, push_idx, push_dst,
               push_rows, push_idx != nullptr ? n_push : 0, ticket);
}

// per batch: partial sums of ||L (a - b)||^2 (b may be null): the unscaled residual norm
__global__ void __launch_bounds__(kSpmvThreads) k_cg_check(int64_t nb, const double *__restrict__ lfac,
                                                           const double *__restrict__ a, const double *__restrict__ b,
                                                           double *__restrict__ part, const PcgScalars *sc,
                                                           const PeerCtx *pc)
{
    __shared__ double vs[kSpmvThreads];
    __shared__ double red[33];
    pdl_wait();
    pdl_trigger();
    const int lr = threadIdx.x / 6, i = threadIdx.x - 6 * lr;
    double rr = 0.0;
    for (int64_t r0 = (int64_t)blockIdx.x * kRowsPerCta; r0 < nb; r0 += (int64_t)gridDim.x * kRowsPerCta) {
        const int64_t row = r0 + lr;
        const bool live = row < nb;
        double v = 0.0;
        if (live) v = a[row * 6 + i] - (b != nullptr ? b[row * 6 + i] : 0.0);
        __syncthreads();
        vs[threadIdx.x] = v;
        __syncthreads();
        if (live) {
            const double t = lower_row_times(lfac, row, i, vs + lr * 6);
            rr = fma(t, t, rr);
        }
    }
    const double s = block_sum<false>(rr, red);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
    if (pc != nullptr && last_cta_done(pc->tickets + 3, false)) {
        const double sa = sum_partials(PartialRef{part, (int)gridDim.x, 1}, red);
        if ((int)threadIdx.x < pc->world) {
            const unsigned long long tag = ((sc->epoch << 32) | (unsigned long long)(unsigned int)sc->checks) + 1;
            Mailbox *m = pc->peer[threadIdx.x];
            m->dot_val[2][pc->rank] = sa;
            st_release_sys(&m->dot_tag[2][pc->rank], tag);
        }
    }
}

// Residual replacement (restart) from the ASSEMBLED matrix: w = f - K x (x = S^T y, K the unscaled full-pattern
// matrix exactly as assembled -- not the scaled, symmetrised K~ the iteration multiplies with, whose entries carry
// a few ulp of rounding that an ill-conditioned system amplifies into the solution), r~ := S w, direction := r~
// (classic) or 0 (single-reduction); partials of ||w||^2 = ||f - K x||^2 and of rho = r~.r~
__global__ void __launch_bounds__(kSpmvThreads) k_cg_restart(int64_t nb, const double *__restrict__ sinv,
                                                             const double *__restrict__ f,
                                                             const double *__restrict__ Kx, double *__restrict__ res,
                                                             double *__restrict__ dir, double dir_scale,
                                                             double *__restrict__ aux_zero,
                                                             double *__restrict__ part_rr, double *__restrict__ part_rho)
{
    __shared__ double vs[kSpmvThreads];
    __shared__ double red[33];
    const int lr = threadIdx.x / 6, i = threadIdx.x - 6 * lr;
    double rr = 0.0, rho = 0.0;
    for (int64_t r0 = (int64_t)blockIdx.x * kRowsPerCta; r0 < nb; r0 += (int64_t)gridDim.x * kRowsPerCta) {
        const int64_t row = r0 + lr;
        const bool live = row < nb;
        double w = 0.0;
        if (live) w = f[row * 6 + i] - Kx[row * 6 + i];
        __syncthreads();
        vs[threadIdx.x] = w;
        __syncthreads();
        if (live) {
            const double v = lower_row_times(sinv, row, i, vs + lr * 6);
            rr = fma(w, w, rr);
            rho = fma(v, v, rho);
            if (res != nullptr) res[row * 6 + i] = v;
            if (dir != nullptr) dir[row * 6 + i] = dir_scale * v;
            if (aux_zero != nullptr) aux_zero[row * 6 + i] = 0.0;
        }
    }
    const double a = block_sum<false>(rr, red);
    const double b = block_sum<false>(rho, red);
    if (threadIdx.x == 0) { part_rr[blockIdx.x] = a; part_rho[blockIdx.x] = b; }
}

// one CTA: resume the iteration from the replaced residual under a new epoch
__global__ void __launch_bounds__(256) k_cg_restart_finish(PartialRef rr, PartialRef rho, PcgScalars *sc,
                                                           unsigned long long epoch, int clear_done)
{
    __shared__ double red[33];
    const double a = sum_partials(rr, red);
    const double b = sum_partials(rho, red);
    if (threadIdx.x == 0) {
        sc->rr = a;
        if (clear_done) {
            sc->rz[0] = b; sc->rz[1] = b;   // whichever parity the next iteration reads
            sc->restart_iter = sc->iter;     // single-reduction CG: next step runs with beta = 0
            sc->epoch = epoch;               // new tags: the halo / dot slots of this iteration index were used already
            sc->done = 0;
        }
    }
}

void launch_cg_restart(cudaStream_t st, int64_t nb, const double *sinv, const double *f, const double *Kx, double *res,
                       double *dir, double dir_scale, double *aux_zero, double *part_rr, double *part_rho)
{
    k_cg_restart<<<vec_grid(nb), kSpmvThreads, 0, st>>>(nb, sinv, f, Kx, res, dir, dir_scale, aux_zero, part_rr, part_rho);
}

void launch_cg_restart_finish(cudaStream_t st, PartialRef rr, PartialRef rho, PcgScalars *sc, unsigned long long epoch,
                              int clear_done)
{
    k_cg_restart_finish<<<1, 256, 0, st>>>(rr, rho, sc, epoch, clear_done);
}

// one CTA: total of the check partials -> sc->rr; converged => done = 1
__global__ void __launch_bounds__(256) k_cg_check_finish(PartialRef part, PcgScalars *sc, const PeerCtx *pc)
{
    __shared__ double red[33];
    pdl_wait();
    pdl_trigger();
    double rr;
    if (pc != nullptr) {
        const unsigned long long tag = ((sc->epoch << 32) | (unsigned long long)(unsigned int)sc->checks) + 1;
        if ((int)threadIdx.x < pc->world) spin_until(&pc->mine->dot_tag[2][threadIdx.x], tag);
        __syncthreads();
        rr = 0.0;
        for (int q = 0; q < pc->world; ++q) rr += __ldcv(&pc->mine->dot_val[2][q]);
    } else {
        rr = sum_partials(part, red);
    }
    if (threadIdx.x == 0) {
        sc->rr = rr;
        sc->checks = sc->checks + 1;
        if (sc->done == 0 && rr <= sc->tol2 * sc->ff) sc->done = 1;
        if (sc->done == 0 && !isfinite(rr)) sc->done = 3;
    }
}

// ---------------------------------------------------------------------------
// Batched CG: many small independent meshes (connected components with contiguous block rows)
// solved concurrently, ONE CTA PER COMPONENT.  The component's four CG vectors live in shared
// memory, its K~ rows (full-row storage, plain BSR) stream from L2; every component runs its own
// iteration count and stops on its own residual ||L r~|| <= tol ||f||.  No grid-wide
// synchronisation, no host polling: a mesh that needs 300 iterations does not wait for one that
// needs 30 000.
// ---------------------------------------------------------------------------
constexpr int kBatchThreads = 384;

__device__ __forceinline__ double batch_block_sum(double v, double *scratch)
{
    // all threads get the result; fixed order
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < kBatchThreads / 32; ++k) t += scratch[k];
    return t;
}

__global__ void __launch_bounds__(kBatchThreads) k_cg_batched(
    int ncomp, const int32_t *__restrict__ comp_ptr, const int32_t *__restrict__ uptr,
    const int32_t *__restrict__ ucol, const double *__restrict__ vals, const double *__restrict__ f,
    const double *__restrict__ sinv, const double *__restrict__ lfac, double *__restrict__ y_out,
    double tol, int max_iter, int check_every, int32_t *__restrict__ comp_iters, double *__restrict__ comp_rr,
    double *__restrict__ comp_ff, int32_t *__restrict__ comp_status)
{
    extern __shared__ double sm[];
    __shared__ double red[kBatchThreads / 32];
    const int b = blockIdx.x;
    if (b >= ncomp) return;
    const int r0 = comp_ptr[b], r1 = comp_ptr[b + 1];
    const int n = (r1 - r0) * 6;
    double *y = sm, *r = sm + n, *p = sm + 2 * n, *Ap = sm + 3 * n;
    // f~ = S f (row-local), ||f||^2
    double ff = 0.0, rho = 0.0;
    for (int j = threadIdx.x; j < n; j += kBatchThreads) {
        const int row = r0 + j / 6, i = j % 6;
        const double *fr = f + (int64_t)row * 6;
        const double fv[6] = {fr[0], fr[1], fr[2], fr[3], fr[4], fr[5]};
        const double t = lower_row_times(sinv, row, i, fv);
        y[j] = 0.0; r[j] = t; p[j] = t;
        rho = fma(t, t, rho);
        ff = fma(fv[i], fv[i], ff);
    }
    rho = batch_block_sum(rho, red);
    ff = batch_block_sum(ff, red);
    const double tol2ff = tol * tol * ff;
    int it = 0, status = (ff == 0.0) ? 1 : 0;
    double rr = ff;
    const double rho0 = rho;
    while (status == 0) {
        // Ap = K~ p (unit diagonal + the row's stored blocks; columns are inside the component)
        for (int j = threadIdx.x; j < n; j += kBatchThreads) {
            const int lrow = j / 6, i = j - 6 * lrow;
            const int row = r0 + lrow;
            double acc = p[j];
            const int32_t kb = __ldg(uptr + row), ke = __ldg(uptr + row + 1);
            for (int32_t k = kb; k < ke; ++k) {
                const double *xx = p + (__ldg(ucol + k) - r0) * 6;
                const double *v = vals + (int64_t)k * 36 + i * 6;
                const double2 v0 = ldg2(v), v1 = ldg2(v + 2), v2 = ldg2(v + 4);
                acc = fma(v0.x, xx[0], acc); acc = fma(v0.y, xx[1], acc);
                acc = fma(v1.x, xx[2], acc); acc = fma(v1.y, xx[3], acc);
                acc = fma(v2.x, xx[4], acc); acc = fma(v2.y, xx[5], acc);
            }
            Ap[j] = acc;
        }
        __syncthreads();
        double pAp = 0.0;
        for (int j = threadIdx.x; j < n; j += kBatchThreads) pAp = fma(p[j], Ap[j], pAp);
        pAp = batch_block_sum(pAp, red);
        if (!(pAp > 0.0) || !isfinite(pAp)) { status = 3; break; }
        const double alpha = rho / pAp;
        double rho_new = 0.0;
        for (int j = threadIdx.x; j < n; j += kBatchThreads) {
            y[j] = fma(alpha, p[j], y[j]);
            const double rv = fma(-alpha, Ap[j], r[j]);
            r[j] = rv;
            rho_new = fma(rv, rv, rho_new);
        }
        rho_new = batch_block_sum(rho_new, red);   // (also orders the writes of r before the reads below)
        ++it;
        if (it % check_every == 0 || rho_new <= 1e-32 * rho0 || it >= max_iter) {
            double t2 = 0.0;
            for (int j = threadIdx.x; j < n; j += kBatchThreads) {
                const int lrow = j / 6, i = j - 6 * lrow;
                const double t = lower_row_times(lfac, r0 + lrow, i, r + lrow * 6);
                t2 = fma(t, t, t2);
            }
            rr = batch_block_sum(t2, red);
            if (rr <= tol2ff) { status = 1; break; }
            if (rho_new <= 1e-32 * rho0) { status = 1; break; }
            if (!isfinite(rr)) { status = 3; break; }
            if (it >= max_iter) { status = 2; break; }
        }
        const double beta = rho_new / rho;
        rho = rho_new;
        for (int j = threadIdx.x; j < n; j += kBatchThreads) p[j] = fma(beta, p[j], r[j]);
        __syncthreads();
    }
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += kBatchThreads) y_out[(int64_t)r0 * 6 + j] = y[j];
    if (threadIdx.x == 0) { comp_iters[b] = it; comp_rr[b] = rr; comp_ff[b] = ff; comp_status[b] = status; }
}

int cg_batched_prepare(size_t smem_bytes)
{
    return (int)cudaFuncSetAttribute(k_cg_batched, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
}

void launch_cg_batched(cudaStream_t st, int ncomp, size_t smem_bytes, const int32_t *comp_ptr, const int32_t *uptr,
                       const int32_t *ucol, const double *vals, const double *f, const double *sinv, const double *lfac,
                       double *y, double tol, int max_iter, int check_every, int32_t *comp_iters, double *comp_rr,
                       double *comp_ff, int32_t *comp_status)
{
    if (ncomp <= 0) return;
    k_cg_batched<<<ncomp, kBatchThreads, smem_bytes, st>>>(ncomp, comp_ptr, uptr, ucol, vals, f, sinv, lfac, y, tol,
                                                          max_iter, check_every, comp_iters, comp_rr, comp_ff, comp_status);
}

// x = S^T y  (back to the unscaled unknowns)
__global__ void __launch_bounds__(kSpmvThreads) k_unscale(int64_t nb, const double *__restrict__ sinv,
                                                          const double *__restrict__ y, double *__restrict__ x)
{
    __shared__ double vs[kSpmvThreads];
    const int lr = threadIdx.x / 6, i = threadIdx.x - 6 * lr;
    for (int64_t r0 = (int64_t)blockIdx.x * kRowsPerCta; r0 < nb; r0 += (int64_t)gridDim.x * kRowsPerCta) {
        const int64_t row = r0 + lr;
        const bool live = row < nb;
        __syncthreads();
        vs[threadIdx.x] = live ? y[row * 6 + i] : 0.0;
        __syncthreads();
        if (live) {
            const double *S = sinv + row * 36;
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < 6; ++j) s = fma(__ldg(S + 6 * j + i), vs[lr * 6 + j], s);  // column i of S
            x[row * 6 + i] = s;
        }
    }
}

void launch_cg_init(cudaStream_t st, int64_t nb, const double *f, const double *sinv, double *y, double *rt,
                    double *ft, double *p, double *part_rho, double *part_ff)
{
    k_cg_init<<<vec_grid(nb), kSpmvThreads, 0, st>>>(nb, f, sinv, y, rt, ft, p, part_rho, part_ff);
}

void launch_cg_init_finish(cudaStream_t st, PartialRef rho, PartialRef ff, PcgScalars *sc, double tol,
                           long long max_iter, unsigned long long epoch, unsigned long long *trace)
{
    k_cg_init_finish<<<1, 256, 0, st>>>(rho, ff, sc, tol, max_iter, epoch, trace);
}

int cg_vec_grid(int64_t nb)
{
    const int64_t blocks = (3 * nb + 255) / 256;   // one double2 per thread and trip
    int64_t cap = (int64_t)g_sm_count * 4;          // one full wave at 4 CTAs of 256 threads per SM
    if (cap > kMaxPartials) cap = kMaxPartials;
    return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

void launch_cg_update(cudaStream_t st, int64_t nb, int parity, PartialRef pAp, const double *p, const double *Ap,
                      double *y, double *rt, double *part_rho, PcgScalars *sc, const PeerCtx *pc, int publish_rho)
{
    launch_pdl(k_cg_update, cg_vec_grid(nb), 256, 0, st, 6 * nb, parity, pAp, p, Ap, y, rt, part_rho, sc, pc, publish_rho);
}

void launch_cg_pupdate(cudaStream_t st, int64_t nb, int parity, PartialRef rho, const double *rt, double *p,
                       PcgScalars *sc, const PeerCtx *pc, const int32_t *push_idx, double *const *push_dst,
                       unsigned int *ticket)
{
    launch_pdl(k_cg_pupdate, cg_vec_grid(nb), 256, 0, st, 6 * nb, parity, rho, rt, p, sc, pc, push_idx, push_dst, ticket);
}

void launch_cg_check(cudaStream_t st, int64_t nb, const double *lfac, const double *a, const double *b, double *part,
                     const PcgScalars *sc, const PeerCtx *pc)
{
    launch_pdl(k_cg_check, vec_grid(nb), kSpmvThreads, 0, st, nb, lfac, a, b, part, sc, pc);
}

void launch_cg_check_finish(cudaStream_t st, PartialRef part, PcgScalars *sc, const PeerCtx *pc)
{
    launch_pdl(k_cg_check_finish, 1, 256, 0, st, part, sc, pc);
}

void launch_unscale(cudaStream_t st, int64_t nb, const double *sinv, const double *y, double *x)
{
    if (nb <= 0) return;
    k_unscale<<<vec_grid(nb), kSpmvThreads, 0, st>>>(nb, sinv, y, x);
}



// P2P halo exchange: every owned boundary x-block is stored straight into its
// slot of the destination rank's ghost region (NVLink peer stores, fire and
// forget); the last CTA then publishes the tag to every neighbour's mailbox.
__global__ void __launch_bounds__(256) k_push_halo(int64_t n_send, const int32_t *__restrict__ send_rows,
                                                   double *const *__restrict__ dst, const double *__restrict__ x,
                                                   const PcgScalars *sc, const PeerCtx *pc)
{
    if (sc->done != 0) return;
    if (sc->iter > 0) trace_stamp(sc, sc->iter - 1, 6);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_send * 3;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = t / 3;
        const int j = (int)(t - s * 3) * 2;
        const double2 v = *reinterpret_cast<const double2 *>(x + (int64_t)__ldg(send_rows + s) * 6 + j);
        *reinterpret_cast<double2 *>(dst[s] + j) = v;
    }
    if (last_cta_done(pc->tickets + 2, true)) {
        const unsigned long long tag = (sc->epoch << 32) | (unsigned long long)(unsigned int)sc->iter;
        if ((int)threadIdx.x < pc->n_neighbors) {
            st_release_sys(&pc->peer[pc->neighbor[threadIdx.x]]->halo_tag[pc->rank], tag);
        }
        if (sc->iter > 0 && sc->trace != nullptr && sc->iter - 1 < kTraceIters && threadIdx.x == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            sc->trace[(sc->iter - 1) * kTraceSlots + 7] = t;
        }
    }
}

int peer_timeout_flag()
{
    int v = 0;
    cudaMemcpyFromSymbol(&v, g_peer_timeout, sizeof(int));
    return v;
}

// start of every solve: a timeout of an earlier solve (or of another context) must not fail this one
void peer_timeout_clear(cudaStream_t st)
{
    const int zero = 0;
    cudaMemcpyToSymbolAsync(g_peer_timeout, &zero, sizeof(int), 0, cudaMemcpyHostToDevice, st);
}

void launch_push_halo(cudaStream_t st, int64_t n_send, const int32_t *send_rows, double *const *dst,
                      const double *x, const PcgScalars *sc, const PeerCtx *pc)
{
    int64_t g = (n_send * 3 + 255) / 256;
    if (g < 1) g = 1;
    if (g > g_sm_count) g = g_sm_count;
    k_push_halo<<<(unsigned)g, 256, 0, st>>>(n_send, send_rows, dst, x, sc, pc);
}

// collapse partial sums to one value each (fixed order); multi-GPU path
__global__ void __launch_bounds__(256) k_reduce_partials(const double *a, int n, double *out_a, const double *b,
                                                         double *out_b, const PcgScalars *sc)
{
    __shared__ double red[33];
    if (sc != nullptr && sc->done != 0) return;
    const double sa = sum_partials(PartialRef{a, n, 1}, red);
    if (threadIdx.x == 0) *out_a = sa;
    if (b != nullptr) {
        const double sb = sum_partials(PartialRef{b, n, 1}, red);
        if (threadIdx.x == 0) *out_b = sb;
    }
}

void launch_reduce_partials(cudaStream_t st, const double *a, int n, double *out_a, const double *b,
                            double *out_b, const PcgScalars *sc)
{
    k_reduce_partials<<<1, 256, 0, st>>>(a, n, out_a, b, out_b, sc);
}

__global__ void k_pack_halo(int64_t n_send, const int32_t *__restrict__ send_rows, const double *__restrict__ x,
                            double *__restrict__ sendbuf, int width)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = t / width;
    const int i = (int)(t - s * width);
    if (s >= n_send) return;
    sendbuf[t] = x[(int64_t)__ldg(send_rows + s) * width + i];
}

void launch_pack_halo(cudaStream_t st, int64_t n_send, const int32_t *send_rows, const double *x, double *sendbuf,
                      int width)
{
    if (n_send <= 0) return;
    const int64_t n = n_send * width;
    k_pack_halo<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n_send, send_rows, x, sendbuf, width);
}

__global__ void k_fill_f64(int64_t n, double *a, double v0, double dv, int period)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v0 + dv * (double)(i % period);
}

void launch_fill(cudaStream_t st, int64_t n, double *a, double v0, double dv, int period)
{
    if (n <= 0) return;
    k_fill_f64<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, a, v0, dv, period);
}

// load vector of the reduced system: f[6*free_index[node]+dof] += magnitude.
// Ordered variant (a job that loads some DOF more than once): one thread, fixed
// order, so duplicates accumulate in input order (deterministic).  Forces on fixed
// vertices are dropped, as elimination drops their equations.
__global__ void k_scatter_forces(int64_t nF, const int64_t *__restrict__ node, const int64_t *__restrict__ dof,
                                 const double *__restrict__ mag, const int32_t *__restrict__ free_index,
                                 int64_t row_begin, int64_t nb, double *f)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    for (int64_t i = 0; i < nF; ++i) {
        const int64_t g = free_index[node[i]];
        if (g < row_begin || g >= row_begin + nb) continue;
        f[(g - row_begin) * 6 + dof[i]] += mag[i];
    }
}

// the same with one thread per force -- only valid when no two forces hit the same (vertex, dof), which the
// host checks (upload_forces): then every destination is written once and the order does not matter
__global__ void k_scatter_forces_unique(int64_t nF, const int64_t *__restrict__ node, const int64_t *__restrict__ dof,
                                        const double *__restrict__ mag, const int32_t *__restrict__ free_index,
                                        int64_t row_begin, int64_t nb, double *f)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nF) return;
    const int64_t g = free_index[node[i]];
    if (g < row_begin || g >= row_begin + nb) return;
    f[(g - row_begin) * 6 + dof[i]] = mag[i];
}

void launch_scatter_forces(cudaStream_t st, int64_t nF, const int64_t *node, const int64_t *dof, const double *mag,
                           const int32_t *free_index, int64_t row_begin, int64_t nb, double *f, bool unique)
{
    if (nF <= 0) return;
    if (unique) k_scatter_forces_unique<<<(unsigned)((nF + 255) / 256), 256, 0, st>>>(nF, node, dof, mag, free_index, row_begin, nb, f);
    else k_scatter_forces<<<1, 32, 0, st>>>(nF, node, dof, mag, free_index, row_begin, nb, f);
}

// ---------------------------------------------------------------------------
// results
// ---------------------------------------------------------------------------
// U_cm: V x 6 column-major (Eigen::MatrixXd nodalDisplacements
// [REF src/beamsimulator.cpp:192]); U_rm: the same row-major for the K7 gather.
__global__ void k_expand(int64_t V, const int32_t *__restrict__ free_index, const double *__restrict__ xg,
                         double *__restrict__ U_cm, double *__restrict__ U_rm)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t v = t / 6;
    const int d = (int)(t - v * 6);
    if (v >= V) return;
    const int32_t fi = __ldg(free_index + v);
    const double u = fi >= 0 ? xg[(int64_t)fi * 6 + d] : 0.0;
    U_rm[t] = u;
    U_cm[(int64_t)d * V + v] = u;
}

void launch_expand(cudaStream_t st, int64_t V, const int32_t *free_index, const double *x_global,
                   double *U_cm, double *U_rm)
{
    if (V <= 0) return;
    const int64_t n = V * 6;
    k_expand<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(V, free_index, x_global, U_cm, U_rm);
}

__global__ void k_cm_to_rm(int64_t V, const double *__restrict__ U_cm, double *__restrict__ U_rm)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t v = t / 6;
    const int d = (int)(t - v * 6);
    if (v >= V) return;
    U_rm[t] = U_cm[(int64_t)d * V + v];
}

void launch_colmajor_to_rowmajor(cudaStream_t st, int64_t V, const double *U_cm, double *U_rm)
{
    if (V <= 0) return;
    const int64_t n = V * 6;
    k_cm_to_rm<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(V, U_cm, U_rm);
}

// K7: f_e = K_l (A u_e), one thread per beam, K_l and A rebuilt in registers.
// Output layout of SimulationResults::edgeForces [REF src/beamsimulator.cpp:198-207]:
// F[c*2E + 2e] = f_e[c], F[c*2E + 2e+1] = f_e[6+c]  -> one 16-byte store per component.
__global__ void __launch_bounds__(128) k_element_forces(int64_t E, const ElemRec *__restrict__ rec,
                                                        const Node4 *__restrict__ node4,
                                                        const double *__restrict__ U_rm,
                                                        double *__restrict__ F6x2E)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    Frame f; Stiff s; int32_t a, b;
    load_elem(rec, node4, e, f, s, a, b);
    double ue[12], ul[12], fe[12];
    const double *ua = U_rm + (int64_t)a * 6, *ub = U_rm + (int64_t)b * 6;
#pragma unroll
    for (int j = 0; j < 6; j += 2) {
        const double2 va = ldg2(ua + j), vb = ldg2(ub + j);
        ue[j] = va.x; ue[j + 1] = va.y; ue[6 + j] = vb.x; ue[6 + j + 1] = vb.y;
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const double *u3 = ue + 3 * g;
        ul[3 * g + 0] = ADD(ADD(MUL(f.nx[0], u3[0]), MUL(f.nx[1], u3[1])), MUL(f.nx[2], u3[2]));
        ul[3 * g + 1] = ADD(ADD(MUL(f.ny[0], u3[0]), MUL(f.ny[1], u3[1])), MUL(f.ny[2], u3[2]));
        ul[3 * g + 2] = ADD(ADD(MUL(f.nz[0], u3[0]), MUL(f.nz[1], u3[1])), MUL(f.nz[2], u3[2]));
    }
    local_forces(s, ul, fe);
#pragma unroll
    for (int c = 0; c < 6; ++c)
        *reinterpret_cast<double2 *>(F6x2E + (int64_t)c * 2 * E + 2 * e) = make_double2(fe[c], fe[6 + c]);
}

// ---------------------------------------------------------------------------
// Colour-mapping pre-pass (the step right after the hot path in the reference:
// igl::colormap(..., normalize=true) + the colour bar's min/max
// [REF src/gui.cpp:514-526; src/colorbar.cpp:79-87]): per force component the
// minimum and maximum over the 2E beam-end values, two-level, order-independent.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_force_minmax(int64_t n, const double *__restrict__ F, double *__restrict__ pmin,
                                                      double *__restrict__ pmax)
{
    __shared__ double smin[8], smax[8];
    const int c = blockIdx.y;
    const double *v = F + (int64_t)c * n;
    double lo = INFINITY, hi = -INFINITY;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        const double x = v[j];
        lo = fmin(lo, x); hi = fmax(hi, x);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_down_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_down_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = lo; smax[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { lo = fmin(lo, smin[w]); hi = fmax(hi, smax[w]); }
        pmin[c * gridDim.x + blockIdx.x] = lo; pmax[c * gridDim.x + blockIdx.x] = hi;
    }
}

__global__ void k_force_minmax_finish(int nparts, const double *__restrict__ pmin, const double *__restrict__ pmax,
                                      double *__restrict__ out /* 6 mins then 6 maxs */)
{
    const int c = threadIdx.x;
    if (c >= 6) return;
    double lo = INFINITY, hi = -INFINITY;
    for (int j = 0; j < nparts; ++j) { lo = fmin(lo, pmin[c * nparts + j]); hi = fmax(hi, pmax[c * nparts + j]); }
    out[c] = lo; out[6 + c] = hi;
}

// t = (v - min) / (max - min) per component (0 when the component is constant)
__global__ void k_force_normalize(int64_t n, const double *__restrict__ F, const double *__restrict__ range,
                                  double *__restrict__ out)
{
    const int c = blockIdx.y;
    const double lo = range[c], hi = range[6 + c];
    const double inv = hi > lo ? 1.0 / (hi - lo) : 0.0;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x)
        out[(int64_t)c * n + j] = (F[(int64_t)c * n + j] - lo) * inv;
}

int force_minmax_parts(int64_t n)
{
    const int64_t b = (n + 255) / 256;
    return (int)(b < 1 ? 1 : (b > g_sm_count ? g_sm_count : b));
}

void launch_force_ranges(cudaStream_t st, int64_t n, const double *F, double *pmin, double *pmax, double *range)
{
    const int parts = force_minmax_parts(n);
    k_force_minmax<<<dim3((unsigned)parts, 6), 256, 0, st>>>(n, F, pmin, pmax);
    k_force_minmax_finish<<<1, 32, 0, st>>>(parts, pmin, pmax, range);
}

void launch_force_normalize(cudaStream_t st, int64_t n, const double *F, const double *range, double *out)
{
    k_force_normalize<<<dim3((unsigned)force_minmax_parts(n), 6), 256, 0, st>>>(n, F, range, out);
}

void launch_element_forces(cudaStream_t st, int64_t E, const ElemRec *rec, const Node4 *node4,
                           const double *U_rm, double *F6x2E)
{
    if (E <= 0) return;
    k_element_forces<<<(unsigned)((E + 127) / 128), 128, 0, st>>>(E, rec, node4, U_rm, F6x2E);
}

}}  // namespace mbfea::dev
