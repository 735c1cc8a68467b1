// The block pattern of K, the contributor lists of K2 and the symmetric-half solver format built ON THE DEVICE
// (one rank, canonical numbering) -- the same arrays, entry for entry, as the host builders in host_prep.cpp
// (build_plan / build_solver_format), which remain the reference and the path for several ranks, renumbered or
// irregular jobs.  Everything here is integer work on sorted keys: no floating point, no order dependence in the
// results (atomics only hand out slots inside a row; the row is sorted afterwards), so the arrays are bit-identical to
// the host's and K stays bit-identical to the oracle's [REF src/beamsimulator.cpp:71-94: element order of the triplets].
//
//   entries   one 64-bit key per (row, contribution): column << 32 | element << 2 | sub-block; a beam with two free
//             ends contributes two entries to each of its two rows, with one free end a single diagonal entry
//   pattern   per row: keys sorted (column, element) -> unique columns = blocks; contributor list of block k =
//             the keys of that column, ascending element = the order Eigen's setFromTriplets sums duplicates in
//   format    upper rule: row r stores its blocks with column > r; the others are transposed references into the
//             storing row's interleaved tile (HostPlan::uptr ... lofs)
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.cuh"
#include "mbfea_internal.hpp"

namespace mbfea { namespace dev {

namespace {
constexpr int kT = 256;
inline unsigned blocks_for(int64_t n, int per = kT) { return (unsigned)((n + per - 1) / per); }

// ---- exclusive scan of n int32 values: out[0 .. n], out[n] = total (three phases, fixed order) ----
constexpr int kPsItems = 4, kPsTile = kT * kPsItems;

__global__ void __launch_bounds__(kT) k_ps_tiles(int64_t n, const int32_t *__restrict__ in, int32_t *__restrict__ out,
                                                 int32_t *__restrict__ tsum)
{
    __shared__ int32_t wsum[kT / 32];
    const int64_t base = (int64_t)blockIdx.x * kPsTile + (int64_t)threadIdx.x * kPsItems;
    int32_t v[kPsItems], s = 0;
#pragma unroll
    for (int j = 0; j < kPsItems; ++j) { v[j] = (base + j < n) ? in[base + j] : 0; s += v[j]; }
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
    for (int j = 0; j < kPsItems; ++j) {
        if (base + j < n) out[base + j] = run;
        run += v[j];
    }
    if (threadIdx.x == kT - 1) tsum[blockIdx.x] = woff + inc;
}

// exclusive scan of the tile totals in place (one CTA); tsum[ntiles] = grand total
__global__ void __launch_bounds__(kT) k_ps_sums(int64_t ntiles, int32_t *tsum)
{
    __shared__ int32_t wsum[kT / 32];
    __shared__ int32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < ntiles; base += kT) {
        const int64_t i = base + threadIdx.x;
        const int32_t v = (i < ntiles) ? tsum[i] : 0;
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
        if (i < ntiles) tsum[i] = carry + woff + inc - v;
        __syncthreads();
        if (threadIdx.x == kT - 1) carry_s = carry + woff + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) tsum[ntiles] = carry_s;
}

__global__ void __launch_bounds__(kT) k_ps_add(int64_t n, int32_t *__restrict__ out, const int32_t *__restrict__ tsum, int64_t ntiles)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] += tsum[i / kPsTile];
    if (i == n) out[n] = tsum[ntiles];
}

__device__ __forceinline__ unsigned long long pd_key(int32_t col, int32_t code)
{
    return ((unsigned long long)(unsigned int)col << 32) | (unsigned int)code;
}

// ---- entries per row ----
__global__ void __launch_bounds__(kT) k_pd_count(int64_t E, const int32_t *__restrict__ elems_cm, const int32_t *__restrict__ fm,
                                                 int32_t *cnt)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int32_t ra = __ldg(fm + __ldg(elems_cm + e)), rb = __ldg(fm + __ldg(elems_cm + E + e));
    if (ra >= 0) atomicAdd(cnt + ra, rb >= 0 ? 2 : 1);
    if (rb >= 0) atomicAdd(cnt + rb, ra >= 0 ? 2 : 1);
}

// the entries of every row, in arbitrary order inside the row (sorted next)
__global__ void __launch_bounds__(kT) k_pd_fill(int64_t E, const int32_t *__restrict__ elems_cm, const int32_t *__restrict__ fm,
                                                const int32_t *__restrict__ start, int32_t *cur, unsigned long long *__restrict__ ent)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int32_t ra = __ldg(fm + __ldg(elems_cm + e)), rb = __ldg(fm + __ldg(elems_cm + E + e));
    const int32_t code = (int32_t)(e << 2);
    if (ra >= 0) {
        const int32_t p = __ldg(start + ra) + atomicAdd(cur + ra, rb >= 0 ? 2 : 1);
        ent[p] = pd_key(ra, code | SUB_00);
        if (rb >= 0) ent[p + 1] = pd_key(rb, code | SUB_01);
    }
    if (rb >= 0) {
        const int32_t p = __ldg(start + rb) + atomicAdd(cur + rb, ra >= 0 ? 2 : 1);
        if (ra >= 0) { ent[p] = pd_key(ra, code | SUB_10); ent[p + 1] = pd_key(rb, code | SUB_11); }
        else ent[p] = pd_key(rb, code | SUB_11);
    }
}

// per row: sort the keys (column, then element: ascending element inside a block), count the blocks; a free vertex
// with no beam still owns a (zero) diagonal block so that the row exists (the solver reports it as singular)
constexpr int kPdLocal = 40;
__global__ void __launch_bounds__(128) k_pd_sort(int64_t nb, const int32_t *__restrict__ start, unsigned long long *ent,
                                                 int32_t *__restrict__ nblk)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nb) return;
    const int32_t b = start[r], n = start[r + 1] - b;
    unsigned long long *p = ent + b;
    int32_t uniq = 0;
    bool has_diag = false;
    if (n <= kPdLocal) {
        unsigned long long a[kPdLocal];
        for (int i = 0; i < n; ++i) a[i] = p[i];
        for (int i = 1; i < n; ++i) {
            const unsigned long long x = a[i];
            int j = i - 1;
            while (j >= 0 && a[j] > x) { a[j + 1] = a[j]; --j; }
            a[j + 1] = x;
        }
        for (int i = 0; i < n; ++i) {
            p[i] = a[i];
            const int32_t col = (int32_t)(a[i] >> 32);
            if (i == 0 || col != (int32_t)(a[i - 1] >> 32)) ++uniq;
            has_diag |= col == (int32_t)r;
        }
    } else {   // a vertex of very high valence: the same insertion sort in place
        for (int i = 1; i < n; ++i) {
            const unsigned long long x = p[i];
            int j = i - 1;
            while (j >= 0 && p[j] > x) { p[j + 1] = p[j]; --j; }
            p[j + 1] = x;
        }
        for (int i = 0; i < n; ++i) {
            const int32_t col = (int32_t)(p[i] >> 32);
            if (i == 0 || col != (int32_t)(p[i - 1] >> 32)) ++uniq;
            has_diag |= col == (int32_t)r;
        }
    }
    nblk[r] = uniq + (has_diag ? 0 : 1);
}

// per row: columns, diagonal position, contributor lists (the entry index IS the contributor index)
__global__ void __launch_bounds__(128) k_pd_emit(int64_t nb, const int32_t *__restrict__ start, const unsigned long long *__restrict__ ent,
                                                 const int32_t *__restrict__ rowptr, int32_t *__restrict__ colidx,
                                                 int32_t *__restrict__ diag_idx, int32_t *__restrict__ contrib_ptr,
                                                 int32_t *__restrict__ contrib)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nb) return;
    const int32_t b = start[r], e2 = start[r + 1], me = (int32_t)r;
    int32_t k = rowptr[r] - 1, prev = -1;
    bool diag_done = false;
    for (int32_t i = b; i < e2; ++i) {
        const unsigned long long key = ent[i];
        const int32_t col = (int32_t)(key >> 32);
        if (i == b || col != prev) {
            if (!diag_done && col > me) {   // isolated diagonal: empty block
                ++k; colidx[k] = me; diag_idx[r] = k; diag_done = true; contrib_ptr[k] = i;
            }
            ++k; colidx[k] = col;
            if (col == me) { diag_idx[r] = k; diag_done = true; }
            contrib_ptr[k] = i;
        }
        contrib[i] = (int32_t)(unsigned int)key;
        prev = col;
    }
    if (!diag_done) { ++k; colidx[k] = me; diag_idx[r] = k; contrib_ptr[k] = e2; }
    if (r == nb - 1) contrib_ptr[rowptr[nb]] = start[nb];
}

// ---- solver format, upper rule ----
// stored / referenced blocks per row: a row stores its ghost columns (local ids >= nb; several ranks) and its local
// columns above the diagonal; the local columns below it are transposed references
__global__ void __launch_bounds__(kT) k_pf_count(int64_t nb, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                                                 int32_t *__restrict__ nu, int32_t *__restrict__ nl)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nb) return;
    int32_t u = 0, l = 0;
    for (int32_t k = rowptr[r]; k < rowptr[r + 1]; ++k) {
        const int32_t c = colidx[k];
        if (c >= nb || c > (int32_t)r) ++u; else if (c < (int32_t)r) ++l;
    }
    nu[r] = u; nl[r] = l;
}

// per tile of 32 rows: 16-byte units of the interleaved layout (longest row x 3 chunks x 192 threads); totals of
// stored blocks and reserved slots (integer atomics: the sums do not depend on the order)
__global__ void __launch_bounds__(kT) k_pf_tiles(int64_t nb, int64_t ntiles, const int32_t *__restrict__ nu, int32_t *__restrict__ tunits,
                                                 unsigned long long *totals /* [0] stored, [1] slots */)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    int32_t mx = 0;
    unsigned long long st = 0;
    for (int64_t r = t * 32; r < nb && r < t * 32 + 32; ++r) { const int32_t u = nu[r]; st += (unsigned long long)u; mx = mx > u ? mx : u; }
    tunits[t] = mx * 3 * 192;
    atomicAdd(totals, st);
    atomicAdd(totals + 1, 32ull * (unsigned long long)mx);
}

__global__ void __launch_bounds__(128) k_pf_fill(int64_t nb, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                                                 const int32_t *__restrict__ diag_idx, const int32_t *__restrict__ uptr,
                                                 const int32_t *__restrict__ lptr, const int32_t *__restrict__ tile_base,
                                                 int32_t *__restrict__ ucol, int32_t *__restrict__ usrc, int32_t *__restrict__ lcol,
                                                 int32_t *__restrict__ lofs)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nb) return;
    int32_t u = uptr[r], l = lptr[r];
    for (int32_t k = rowptr[r]; k < rowptr[r + 1]; ++k) {
        const int32_t c = colidx[k];
        if (c == (int32_t)r) continue;
        if (c >= nb || c > (int32_t)r) { ucol[u] = c; usrc[u] = k; ++u; continue; }
        // transposed reference: the slot of (c, r) inside row c's stored list (present by symmetry of the local pattern).
        // Row c lists [ghosts of lower ranks][local columns below c][c][local columns above c][ghosts of higher ranks]
        // (ascending GLOBAL ids); it stores the ghosts and the local columns above c, in that order.
        int32_t low = rowptr[c];
        while (colidx[low] >= nb) ++low;             // leading ghost columns (the diagonal ends the scan at the latest)
        const int32_t first = diag_idx[c] + 1;
        int32_t lo = first, hi = rowptr[c + 1];      // ascending local ids behind the diagonal (ghost ids are >= nb)
        while (lo < hi) { const int32_t mid = (lo + hi) >> 1; if (colidx[mid] < (int32_t)r) lo = mid + 1; else hi = mid; }
        lcol[l] = c;
        lofs[l] = tile_base[c >> 5] + ((low - rowptr[c]) + (lo - first)) * 3 * 192 + (c & 31) * 6;
        ++l;
    }
}

// ---- the rank's share of a pattern built for the whole system (several ranks) ----
// ghost columns of the rows [r0, r1): global columns outside the range, marked (then ranked by a scan: ascending order)
__global__ void __launch_bounds__(kT) k_pl_mark(int64_t k0, int64_t k1, const int32_t *__restrict__ gcol, int32_t r0, int32_t r1, int32_t *mark)
{
    const int64_t k = k0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= k1) return;
    const int32_t g = gcol[k];
    if (g < r0 || g >= r1) mark[g] = 1;   // (every writer stores the same value)
}
__global__ void __launch_bounds__(kT) k_pl_ghosts(int64_t nbg, const int32_t *__restrict__ mark, const int32_t *__restrict__ gidx,
                                                  long long *__restrict__ ghost_global)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < nbg && mark[g]) ghost_global[gidx[g]] = g;
}
// local column ids: owned [0, nb), ghosts after (in ascending global order); the blocks of a row keep their global order
__global__ void __launch_bounds__(kT) k_pl_localize(int64_t k0, int64_t k1, const int32_t *__restrict__ gcol, int32_t r0, int32_t r1,
                                                    const int32_t *__restrict__ gidx, int32_t *__restrict__ colidx_l)
{
    const int64_t k = k0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= k1) return;
    const int32_t g = gcol[k];
    colidx_l[k - k0] = (g >= r0 && g < r1) ? g - r0 : (r1 - r0) + gidx[g];
}
__global__ void __launch_bounds__(kT) k_pl_shift(int64_t n, const int32_t *__restrict__ src, int32_t delta, int32_t *__restrict__ dst)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i] - delta;
}
}  // namespace

void launch_scan_i32(cudaStream_t st, int64_t n, const int32_t *in, int32_t *out, int32_t *tsum)
{
    const int64_t ntiles = (n + kPsTile - 1) / kPsTile;
    if (n > 0) k_ps_tiles<<<(unsigned)ntiles, kT, 0, st>>>(n, in, out, tsum);
    k_ps_sums<<<1, kT, 0, st>>>(ntiles, tsum);
    k_ps_add<<<blocks_for(n + 1), kT, 0, st>>>(n, out, tsum, ntiles);
}
int64_t scan_i32_scratch(int64_t n) { return (n + kPsTile - 1) / kPsTile + 2; }

void launch_pd_count(cudaStream_t st, int64_t E, const int32_t *elems_cm, const int32_t *fm, int32_t *cnt)
{
    if (E > 0) k_pd_count<<<blocks_for(E), kT, 0, st>>>(E, elems_cm, fm, cnt);
}
void launch_pd_fill(cudaStream_t st, int64_t E, const int32_t *elems_cm, const int32_t *fm, const int32_t *start, int32_t *cur,
                    unsigned long long *ent)
{
    if (E > 0) k_pd_fill<<<blocks_for(E), kT, 0, st>>>(E, elems_cm, fm, start, cur, ent);
}
void launch_pd_sort(cudaStream_t st, int64_t nb, const int32_t *start, unsigned long long *ent, int32_t *nblk)
{
    if (nb > 0) k_pd_sort<<<blocks_for(nb, 128), 128, 0, st>>>(nb, start, ent, nblk);
}
void launch_pd_emit(cudaStream_t st, int64_t nb, const int32_t *start, const unsigned long long *ent, const int32_t *rowptr,
                    int32_t *colidx, int32_t *diag_idx, int32_t *contrib_ptr, int32_t *contrib)
{
    if (nb > 0) k_pd_emit<<<blocks_for(nb, 128), 128, 0, st>>>(nb, start, ent, rowptr, colidx, diag_idx, contrib_ptr, contrib);
}
void launch_pf_count(cudaStream_t st, int64_t nb, const int32_t *rowptr, const int32_t *colidx, int32_t *nu, int32_t *nl)
{
    if (nb > 0) k_pf_count<<<blocks_for(nb), kT, 0, st>>>(nb, rowptr, colidx, nu, nl);
}
void launch_pl_mark(cudaStream_t st, int64_t k0, int64_t k1, const int32_t *gcol, int64_t r0, int64_t r1, int32_t *mark)
{
    if (k1 > k0) k_pl_mark<<<blocks_for(k1 - k0), kT, 0, st>>>(k0, k1, gcol, (int32_t)r0, (int32_t)r1, mark);
}
void launch_pl_ghosts(cudaStream_t st, int64_t nbg, const int32_t *mark, const int32_t *gidx, long long *ghost_global)
{
    if (nbg > 0) k_pl_ghosts<<<blocks_for(nbg), kT, 0, st>>>(nbg, mark, gidx, ghost_global);
}
void launch_pl_localize(cudaStream_t st, int64_t k0, int64_t k1, const int32_t *gcol, int64_t r0, int64_t r1, const int32_t *gidx,
                        int32_t *colidx_l)
{
    if (k1 > k0) k_pl_localize<<<blocks_for(k1 - k0), kT, 0, st>>>(k0, k1, gcol, (int32_t)r0, (int32_t)r1, gidx, colidx_l);
}
void launch_pl_shift(cudaStream_t st, int64_t n, const int32_t *src, int32_t delta, int32_t *dst)
{
    if (n > 0) k_pl_shift<<<blocks_for(n), kT, 0, st>>>(n, src, delta, dst);
}
void launch_pf_tiles(cudaStream_t st, int64_t nb, const int32_t *nu, int32_t *tunits, unsigned long long *totals)
{
    const int64_t ntiles = (nb + 31) / 32;
    if (ntiles > 0) k_pf_tiles<<<blocks_for(ntiles), kT, 0, st>>>(nb, ntiles, nu, tunits, totals);
}
void launch_pf_fill(cudaStream_t st, int64_t nb, const int32_t *rowptr, const int32_t *colidx, const int32_t *diag_idx,
                    const int32_t *uptr, const int32_t *lptr, const int32_t *tile_base, int32_t *ucol, int32_t *usrc, int32_t *lcol,
                    int32_t *lofs)
{
    if (nb > 0) k_pf_fill<<<blocks_for(nb, 128), 128, 0, st>>>(nb, rowptr, colidx, diag_idx, uptr, lptr, tile_base, ucol, usrc, lcol, lofs);
}

}}  // namespace mbfea::dev
