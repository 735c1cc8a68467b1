// Symbolic half of the multilevel preconditioner built ON THE DEVICE (one rank): brick aggregates of every level, the
// level-1 pattern and the Galerkin contributor lists of the fine level, and for every coarse-to-coarse transfer the
// patterns of P, P^T, A P and P^T A P -- the same arrays, entry for entry, as the host builder (hierarchy.cpp:
// build_ml_hierarchy), which stays the reference (CPU tests pin it against scipy's P^T K P) and the path for several
// ranks.  Integer work throughout, plus the centroids (sums in ascending member order, like the host).
//   aggregates   counting sort of the rows by brick (atomics hand out slots, each member list is sorted afterwards)
//   level-1 pattern, Galerkin lists   one thread per aggregate walks its members in ascending order: fixed order,
//                no atomics
//   P, A P, P^T A P   one CTA per row marks the candidate columns in a shared-memory bitmap; scanning the bitmap
//                yields the columns already sorted and unique (count pass, prefix sum, fill pass)
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.cuh"

namespace mbfea { namespace dev {

namespace {
constexpr int kT = 256;
inline unsigned blocks_for(int64_t n, int per = kT) { return (unsigned)((n + per - 1) / per); }

__global__ void __launch_bounds__(kT) k_h_pos(int64_t V, const Node4 *__restrict__ node4, const int32_t *__restrict__ fm,
                                              double *__restrict__ pos)
{
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const int32_t r = fm[v];
    if (r < 0) return;
    const Node4 p = node4[v];
    pos[3 * (int64_t)r] = p.x; pos[3 * (int64_t)r + 1] = p.y; pos[3 * (int64_t)r + 2] = p.z;
}

// bounding box (min / max: exact, order-free), two phases
__global__ void __launch_bounds__(kT) k_h_box_part(int64_t n, const double *__restrict__ pos, double *__restrict__ part)
{
    __shared__ double sm[6][kT / 32];
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x)
#pragma unroll
        for (int a = 0; a < 3; ++a) { const double x = pos[3 * v + a]; lo[a] = fmin(lo[a], x); hi[a] = fmax(hi[a], x); }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fmin(lo[a], __shfl_down_sync(0xffffffffu, lo[a], o));
            hi[a] = fmax(hi[a], __shfl_down_sync(0xffffffffu, hi[a], o));
        }
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int a = 0; a < 3; ++a) { sm[a][threadIdx.x >> 5] = lo[a]; sm[3 + a][threadIdx.x >> 5] = hi[a]; }
    __syncthreads();
    if (threadIdx.x < 6) {
        double r = sm[threadIdx.x][0];
        for (int w = 1; w < kT / 32; ++w) r = threadIdx.x < 3 ? fmin(r, sm[threadIdx.x][w]) : fmax(r, sm[threadIdx.x][w]);
        part[(int64_t)blockIdx.x * 6 + threadIdx.x] = r;
    }
}
__global__ void k_h_box_fin(int nparts, const double *__restrict__ part, double *__restrict__ box)
{
    if (threadIdx.x >= 6) return;
    double r = part[threadIdx.x];
    for (int p = 1; p < nparts; ++p) r = threadIdx.x < 3 ? fmin(r, part[p * 6 + threadIdx.x]) : fmax(r, part[p * 6 + threadIdx.x]);
    box[threadIdx.x] = r;
}

// parts (several ranks, fine level only): rows [bound[q], bound[q + 1]) belong to part q; a brick cut by a part boundary
// becomes one aggregate per part and the aggregates are numbered part-major (key += q * bricks of the grid)
struct BrickGrid { double lo[3]; double cell; int64_t cnt[3]; int32_t nparts; int64_t bound[kMaxWorld + 1]; };

// brick of every row (the host's formula, operation for operation) + rows per brick
__global__ void __launch_bounds__(kT) k_h_brick(int64_t n, const double *__restrict__ pos, BrickGrid g, int32_t *__restrict__ key,
                                                int32_t *bcnt)
{
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    int64_t c[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double q = __dadd_rn(__ddiv_rn(__dsub_rn(pos[3 * v + a], g.lo[a]), g.cell), 1e-9);
        int64_t k = (int64_t)floor(q);
        k = k < 0 ? 0 : k;
        c[a] = k > g.cnt[a] - 1 ? g.cnt[a] - 1 : k;
    }
    int64_t b64 = (c[0] * g.cnt[1] + c[1]) * g.cnt[2] + c[2];
    if (g.nparts > 1) {
        int q = 0;
        while (q + 1 < g.nparts && v >= g.bound[q + 1]) ++q;
        b64 += (int64_t)q * (g.cnt[0] * g.cnt[1] * g.cnt[2]);
    }
    const int32_t b = (int32_t)b64;
    key[v] = b;
    atomicAdd(bcnt + b, 1);
}
__global__ void __launch_bounds__(kT) k_h_occ(int64_t nk, const int32_t *__restrict__ bcnt, int32_t *__restrict__ occ)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nk) occ[b] = bcnt[b] > 0;
}
// aggregate of every row; member slots handed out by a cursor per brick (sorted afterwards)
__global__ void __launch_bounds__(kT) k_h_assign(int64_t n, const int32_t *__restrict__ key, const int32_t *__restrict__ aggid,
                                                 const int32_t *__restrict__ bstart, int32_t *bcur, int32_t *__restrict__ agg,
                                                 int32_t *__restrict__ agg_rows)
{
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const int32_t b = key[v];
    agg[v] = aggid[b];
    agg_rows[bstart[b] + atomicAdd(bcur + b, 1)] = (int32_t)v;
}
__global__ void __launch_bounds__(kT) k_h_aggptr(int64_t nk, const int32_t *__restrict__ bcnt, const int32_t *__restrict__ aggid,
                                                 const int32_t *__restrict__ bstart, int32_t *__restrict__ agg_ptr, int32_t n, int32_t na)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nk && bcnt[b] > 0) agg_ptr[aggid[b]] = bstart[b];
    if (b == 0) agg_ptr[na] = n;
}

// in-place ascending sort of a short int32 segment (insertion sort; shell gaps for long ones)
__device__ __forceinline__ void seg_sort(int32_t *p, int32_t n)
{
    for (int32_t gap = n > 64 ? 57 : 1; gap >= 1; gap = gap > 1 ? (gap > 10 ? gap / 3 : 1) : 0) {
        for (int32_t i = gap; i < n; ++i) {
            const int32_t x = p[i];
            int32_t j = i - gap;
            while (j >= 0 && p[j] > x) { p[j + gap] = p[j]; j -= gap; }
            p[j + gap] = x;
        }
        if (gap == 1) break;
    }
}
__global__ void __launch_bounds__(128) k_h_sort_members(int64_t na, const int32_t *__restrict__ agg_ptr, int32_t *agg_rows, int32_t *maxsz)
{
    const int64_t I = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= na) return;
    const int32_t b = agg_ptr[I], n = agg_ptr[I + 1] - b;
    seg_sort(agg_rows + b, n);
    atomicMax(maxsz, n);
}
// centroid = (sum of the member positions, ascending rows) / members
__global__ void __launch_bounds__(128) k_h_centroid(int64_t na, const int32_t *__restrict__ agg_ptr, const int32_t *__restrict__ agg_rows,
                                                    const double *__restrict__ pos, double *__restrict__ centroid)
{
    const int64_t I = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= na) return;
    double s[3] = {0.0, 0.0, 0.0};
    for (int32_t q = agg_ptr[I]; q < agg_ptr[I + 1]; ++q) {
        const int64_t v = agg_rows[q];
#pragma unroll
        for (int a = 0; a < 3; ++a) s[a] = __dadd_rn(s[a], pos[3 * v + a]);
    }
    const double m = (double)(agg_ptr[I + 1] - agg_ptr[I]);
#pragma unroll
    for (int a = 0; a < 3; ++a) centroid[3 * I + a] = __ddiv_rn(s[a], m);
}
__global__ void __launch_bounds__(kT) k_h_off(int64_t n, const double *__restrict__ pos, const int32_t *__restrict__ agg,
                                              const double *__restrict__ centroid, double *__restrict__ off)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3 * n) return;
    const int64_t v = t / 3;
    const int a = (int)(t - 3 * v);
    off[t] = __dsub_rn(pos[t], centroid[3 * (int64_t)agg[v] + a]);
}

// ---- level-1 pattern of the fine level: aggregate I couples to agg(col) of every block of its member rows ----
__global__ void __launch_bounds__(128) k_h0_ub(int64_t na, const int32_t *__restrict__ agg_ptr, const int32_t *__restrict__ agg_rows,
                                               const int32_t *__restrict__ rowptr, int32_t *__restrict__ ub)
{
    const int64_t I = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= na) return;
    int32_t s = 1;
    for (int32_t q = agg_ptr[I]; q < agg_ptr[I + 1]; ++q) { const int32_t v = agg_rows[q]; s += rowptr[v + 1] - rowptr[v]; }
    ub[I] = s;
}
// candidates of aggregate I into its segment, sorted, made unique in place; nuniq[I]
__global__ void __launch_bounds__(128) k_h0_cand(int64_t na, const int32_t *__restrict__ agg_ptr, const int32_t *__restrict__ agg_rows,
                                                 const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                                                 const int32_t *__restrict__ agg, const int32_t *__restrict__ cstart, int32_t *cand,
                                                 int32_t *__restrict__ nuniq)
{
    const int64_t I = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= na) return;
    int32_t *p = cand + cstart[I];
    int32_t n = 0;
    p[n++] = (int32_t)I;
    for (int32_t q = agg_ptr[I]; q < agg_ptr[I + 1]; ++q) {
        const int32_t v = agg_rows[q];
        for (int32_t k = rowptr[v]; k < rowptr[v + 1]; ++k) p[n++] = agg[colidx[k]];
    }
    seg_sort(p, n);
    int32_t u = 0;
    for (int32_t i = 0; i < n; ++i)
        if (i == 0 || p[i] != p[i - 1]) p[u++] = p[i];
    nuniq[I] = u;
}
__global__ void __launch_bounds__(128) k_h0_emit(int64_t na, const int32_t *__restrict__ cstart, const int32_t *__restrict__ cand,
                                                 const int32_t *__restrict__ rowptr1, int32_t *__restrict__ colidx1,
                                                 int32_t *__restrict__ diag1)
{
    const int64_t I = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= na) return;
    const int32_t *p = cand + cstart[I];
    const int32_t b = rowptr1[I], n = rowptr1[I + 1] - b;
    for (int32_t i = 0; i < n; ++i) {
        colidx1[b + i] = p[i];
        if (p[i] == (int32_t)I) diag1[I] = b + i;
    }
}
// Galerkin contributor lists: fine block k = (a, b) is summed into coarse block (agg a, agg b); per coarse block the fine
// blocks ascend.  The blocks of coarse row I receive from I's members only: one thread per aggregate, members and
// their blocks in ascending order, plain counters (FILL: cursors) on the row's own segment.
template <bool FILL>
__global__ void __launch_bounds__(128) k_h0_gal(int64_t na, const int32_t *__restrict__ agg_ptr, const int32_t *__restrict__ agg_rows,
                                                const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                                                const int32_t *__restrict__ agg, const int32_t *__restrict__ rowptr1,
                                                const int32_t *__restrict__ colidx1, int32_t *cnt_or_cur,
                                                const int32_t *__restrict__ gal_ptr, int32_t *__restrict__ gal_src)
{
    const int64_t I = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= na) return;
    const int32_t c0 = rowptr1[I], c1 = rowptr1[I + 1];
    for (int32_t q = agg_ptr[I]; q < agg_ptr[I + 1]; ++q) {
        const int32_t v = agg_rows[q];
        for (int32_t k = rowptr[v]; k < rowptr[v + 1]; ++k) {
            const int32_t J = agg[colidx[k]];
            int32_t lo = c0, hi = c1;
            while (lo < hi) { const int32_t mid = (lo + hi) >> 1; if (colidx1[mid] < J) lo = mid + 1; else hi = mid; }
            if (FILL) gal_src[gal_ptr[lo] + cnt_or_cur[lo]++] = k;
            else cnt_or_cur[lo]++;
        }
    }
}

// ---- coarse-to-coarse transfers: rows as sorted unique unions, through a shared-memory bitmap ----
// MODE 0  P:        row i -> agg[i] (and, smoothed, agg[j] of every block (i, j) of A)
// MODE 1  A P:      row i -> the P columns of every j with a block (i, j)
// MODE 2  P^T A P:  row I -> I and the A P columns of every fine row with a P block in column I
struct RowsArgs {
    int64_t nrows; int32_t ncols; int32_t smoothed;
    const int32_t *arow, *acol;     // A (modes 0, 1)
    const int32_t *agg;             // mode 0
    const int32_t *prow, *pcol;     // P (mode 1)
    const int32_t *rrow, *rfine;    // P^T (mode 2)
    const int32_t *aprow, *apcol;   // A P (mode 2)
};
template <int MODE, bool FILL>
__global__ void __launch_bounds__(128) k_hx_rows(RowsArgs a, int32_t *__restrict__ rowlen, const int32_t *__restrict__ rowptr_out,
                                                 int32_t *__restrict__ cols_out)
{
    extern __shared__ unsigned int bits[];
    __shared__ int wlo_s, whi_s, run_s;
    __shared__ int wsum[4];
    const int nwords = (a.ncols + 31) >> 5;
    for (int w = threadIdx.x; w < nwords; w += blockDim.x) bits[w] = 0u;
    for (int64_t row = blockIdx.x; row < a.nrows; row += gridDim.x) {
        if (threadIdx.x == 0) { wlo_s = nwords; whi_s = -1; run_s = 0; }
        __syncthreads();
        int wlo = nwords, whi = -1;
        auto mark = [&](int32_t c) {
            atomicOr(&bits[c >> 5], 1u << (c & 31));
            wlo = min(wlo, c >> 5); whi = max(whi, c >> 5);
        };
        if (MODE == 0) {
            if (threadIdx.x == 0) mark(a.agg[row]);
            if (a.smoothed)
                for (int32_t k = a.arow[row] + threadIdx.x; k < a.arow[row + 1]; k += blockDim.x) mark(a.agg[a.acol[k]]);
        } else if (MODE == 1) {
            for (int32_t k = a.arow[row] + threadIdx.x; k < a.arow[row + 1]; k += blockDim.x) {
                const int32_t j = a.acol[k];
                for (int32_t b = a.prow[j]; b < a.prow[j + 1]; ++b) mark(a.pcol[b]);
            }
        } else {
            if (threadIdx.x == 0) mark((int32_t)row);
            for (int32_t e = a.rrow[row] + threadIdx.x; e < a.rrow[row + 1]; e += blockDim.x) {
                const int32_t i = a.rfine[e];
                for (int32_t b = a.aprow[i]; b < a.aprow[i + 1]; ++b) mark(a.apcol[b]);
            }
        }
        if (whi >= 0) { atomicMin(&wlo_s, wlo); atomicMax(&whi_s, whi); }
        __syncthreads();
        const int w0 = wlo_s, w1 = whi_s;   // touched words [w0, w1]
        const int32_t base = FILL ? rowptr_out[row] : 0;
        // chunks of blockDim words: popcounts, exclusive scan inside the chunk, running offset across chunks
        for (int c0 = w0; c0 <= w1; c0 += blockDim.x) {
            const int w = c0 + threadIdx.x;
            const unsigned int word = w <= w1 ? bits[w] : 0u;
            const int pc = __popc(word);
            int inc = pc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if ((threadIdx.x & 31) >= o) inc += t; }
            if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = inc;
            __syncthreads();
            int woff = 0;
            for (int j = 0; j < (int)(threadIdx.x >> 5); ++j) woff += wsum[j];
            const int run = run_s;
            if (FILL && word) {
                int32_t o = base + run + woff + inc - pc;
                unsigned int m = word;
                while (m) { const int bit = __ffs(m) - 1; cols_out[o++] = (w << 5) + bit; m &= m - 1; }
            }
            if (w <= w1) bits[w] = 0u;   // leave the bitmap clean for the next row
            __syncthreads();
            if (threadIdx.x == blockDim.x - 1) run_s = run + woff + inc;
            __syncthreads();
        }
        if (!FILL && threadIdx.x == 0) rowlen[row] = run_s;
        __syncthreads();
    }
}

// ---- P^T: per coarse column the P blocks of that column (ascending) and their fine rows ----
__global__ void __launch_bounds__(kT) k_hx_tcount(int64_t nnzp, const int32_t *__restrict__ pcol, int32_t *rcnt)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nnzp) atomicAdd(rcnt + pcol[b], 1);
}
__global__ void __launch_bounds__(kT) k_hx_tfill(int64_t nnzp, const int32_t *__restrict__ pcol, const int32_t *__restrict__ rrow,
                                                 int32_t *rcur, int32_t *__restrict__ rblk)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nnzp) return;
    const int32_t c = pcol[b];
    rblk[rrow[c] + atomicAdd(rcur + c, 1)] = (int32_t)b;
}
__global__ void __launch_bounds__(128) k_hx_tsort(int64_t nc, const int32_t *__restrict__ rrow, int32_t *rblk,
                                                  const int32_t *__restrict__ prow_of, int32_t *__restrict__ rfine)
{
    const int64_t I = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= nc) return;
    const int32_t b = rrow[I], n = rrow[I + 1] - b;
    seg_sort(rblk + b, n);
    for (int32_t e = b; e < b + n; ++e) rfine[e] = prow_of[rblk[e]];
}
// position of the diagonal in every row (columns ascend, diagonal present)
__global__ void __launch_bounds__(kT) k_h_diag(int64_t n, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                                               int32_t *__restrict__ diag)
{
    const int64_t I = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= n) return;
    int32_t lo = rowptr[I], hi = rowptr[I + 1];
    while (lo < hi) { const int32_t mid = (lo + hi) >> 1; if (colidx[mid] < (int32_t)I) lo = mid + 1; else hi = mid; }
    diag[I] = lo;
}

template <int MODE>
void rows_pass(cudaStream_t st, const RowsArgs &a, bool fill, int32_t *rowlen, const int32_t *rowptr_out, int32_t *cols_out)
{
    if (a.nrows <= 0) return;
    const size_t smem = (size_t)((a.ncols + 31) / 32) * sizeof(unsigned int);
    const unsigned grid = (unsigned)(a.nrows < 148 * 16 ? a.nrows : 148 * 16);
    if (fill) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(k_hx_rows<MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_hx_rows<MODE, true><<<grid, 128, smem, st>>>(a, rowlen, rowptr_out, cols_out);
    } else {
        if (smem > 48 * 1024) cudaFuncSetAttribute(k_hx_rows<MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_hx_rows<MODE, false><<<grid, 128, smem, st>>>(a, rowlen, rowptr_out, cols_out);
    }
}
}  // namespace

void launch_h_pos(cudaStream_t st, int64_t V, const Node4 *node4, const int32_t *fm, double *pos)
{
    if (V > 0) k_h_pos<<<blocks_for(V), kT, 0, st>>>(V, node4, fm, pos);
}
// box[0..2] = min, box[3..5] = max of the n positions; part: 6 * 256 doubles of scratch
void launch_h_box(cudaStream_t st, int64_t n, const double *pos, double *part, double *box)
{
    const int grid = (int)(n < 256LL * kT ? (n + kT - 1) / kT : 256);
    k_h_box_part<<<(unsigned)(grid < 1 ? 1 : grid), kT, 0, st>>>(n, pos, part);
    k_h_box_fin<<<1, 32, 0, st>>>(grid < 1 ? 1 : grid, part, box);
}
void launch_h_brick(cudaStream_t st, int64_t n, const double *pos, const double lo[3], double cell, const int64_t cnt[3], int32_t *key,
                    int32_t *bcnt, const int64_t *part_bounds, int nparts)
{
    BrickGrid g;
    for (int a = 0; a < 3; ++a) { g.lo[a] = lo[a]; g.cnt[a] = cnt[a]; }
    g.cell = cell;
    g.nparts = (part_bounds != nullptr && nparts > 1) ? nparts : 1;
    for (int q = 0; q <= kMaxWorld; ++q) g.bound[q] = (g.nparts > 1 && q <= nparts) ? part_bounds[q] : 0;
    if (n > 0) k_h_brick<<<blocks_for(n), kT, 0, st>>>(n, pos, g, key, bcnt);
}
void launch_h_occ(cudaStream_t st, int64_t nk, const int32_t *bcnt, int32_t *occ)
{
    if (nk > 0) k_h_occ<<<blocks_for(nk), kT, 0, st>>>(nk, bcnt, occ);
}
void launch_h_assign(cudaStream_t st, int64_t n, const int32_t *key, const int32_t *aggid, const int32_t *bstart, int32_t *bcur,
                     int32_t *agg, int32_t *agg_rows)
{
    if (n > 0) k_h_assign<<<blocks_for(n), kT, 0, st>>>(n, key, aggid, bstart, bcur, agg, agg_rows);
}
void launch_h_aggptr(cudaStream_t st, int64_t nk, const int32_t *bcnt, const int32_t *aggid, const int32_t *bstart, int32_t *agg_ptr,
                     int64_t n, int64_t na)
{
    if (nk > 0) k_h_aggptr<<<blocks_for(nk), kT, 0, st>>>(nk, bcnt, aggid, bstart, agg_ptr, (int32_t)n, (int32_t)na);
}
void launch_h_sort_members(cudaStream_t st, int64_t na, const int32_t *agg_ptr, int32_t *agg_rows, int32_t *maxsz)
{
    if (na > 0) k_h_sort_members<<<blocks_for(na, 128), 128, 0, st>>>(na, agg_ptr, agg_rows, maxsz);
}
void launch_h_centroid_off(cudaStream_t st, int64_t n, int64_t na, const int32_t *agg_ptr, const int32_t *agg_rows, const int32_t *agg,
                           const double *pos, double *centroid, double *off)
{
    if (na > 0) k_h_centroid<<<blocks_for(na, 128), 128, 0, st>>>(na, agg_ptr, agg_rows, pos, centroid);
    if (n > 0) k_h_off<<<blocks_for(3 * n), kT, 0, st>>>(n, pos, agg, centroid, off);
}
void launch_h0_ub(cudaStream_t st, int64_t na, const int32_t *agg_ptr, const int32_t *agg_rows, const int32_t *rowptr, int32_t *ub)
{
    if (na > 0) k_h0_ub<<<blocks_for(na, 128), 128, 0, st>>>(na, agg_ptr, agg_rows, rowptr, ub);
}
void launch_h0_cand(cudaStream_t st, int64_t na, const int32_t *agg_ptr, const int32_t *agg_rows, const int32_t *rowptr,
                    const int32_t *colidx, const int32_t *agg, const int32_t *cstart, int32_t *cand, int32_t *nuniq)
{
    if (na > 0) k_h0_cand<<<blocks_for(na, 128), 128, 0, st>>>(na, agg_ptr, agg_rows, rowptr, colidx, agg, cstart, cand, nuniq);
}
void launch_h0_emit(cudaStream_t st, int64_t na, const int32_t *cstart, const int32_t *cand, const int32_t *rowptr1, int32_t *colidx1,
                    int32_t *diag1)
{
    if (na > 0) k_h0_emit<<<blocks_for(na, 128), 128, 0, st>>>(na, cstart, cand, rowptr1, colidx1, diag1);
}
void launch_h0_gal(cudaStream_t st, bool fill, int64_t na, const int32_t *agg_ptr, const int32_t *agg_rows, const int32_t *rowptr,
                   const int32_t *colidx, const int32_t *agg, const int32_t *rowptr1, const int32_t *colidx1, int32_t *cnt_or_cur,
                   const int32_t *gal_ptr, int32_t *gal_src)
{
    if (na <= 0) return;
    if (fill) k_h0_gal<true><<<blocks_for(na, 128), 128, 0, st>>>(na, agg_ptr, agg_rows, rowptr, colidx, agg, rowptr1, colidx1, cnt_or_cur, gal_ptr, gal_src);
    else k_h0_gal<false><<<blocks_for(na, 128), 128, 0, st>>>(na, agg_ptr, agg_rows, rowptr, colidx, agg, rowptr1, colidx1, cnt_or_cur, gal_ptr, gal_src);
}
void launch_hx_p_rows(cudaStream_t st, bool fill, int64_t n, int64_t ncols, bool smoothed, const int32_t *arow, const int32_t *acol,
                      const int32_t *agg, int32_t *rowlen, const int32_t *prow, int32_t *pcol)
{
    RowsArgs a{};
    a.nrows = n; a.ncols = (int32_t)ncols; a.smoothed = smoothed ? 1 : 0; a.arow = arow; a.acol = acol; a.agg = agg;
    rows_pass<0>(st, a, fill, rowlen, prow, pcol);
}
void launch_hx_ap_rows(cudaStream_t st, bool fill, int64_t n, int64_t ncols, const int32_t *arow, const int32_t *acol, const int32_t *prow,
                       const int32_t *pcol, int32_t *rowlen, const int32_t *aprow, int32_t *apcol)
{
    RowsArgs a{};
    a.nrows = n; a.ncols = (int32_t)ncols; a.arow = arow; a.acol = acol; a.prow = prow; a.pcol = pcol;
    rows_pass<1>(st, a, fill, rowlen, aprow, apcol);
}
void launch_hx_rap_rows(cudaStream_t st, bool fill, int64_t nc, const int32_t *rrow, const int32_t *rfine, const int32_t *aprow,
                        const int32_t *apcol, int32_t *rowlen, const int32_t *crow, int32_t *ccol)
{
    RowsArgs a{};
    a.nrows = nc; a.ncols = (int32_t)nc; a.rrow = rrow; a.rfine = rfine; a.aprow = aprow; a.apcol = apcol;
    rows_pass<2>(st, a, fill, rowlen, crow, ccol);
}
void launch_hx_transpose(cudaStream_t st, int phase, int64_t nnzp, int64_t nc, const int32_t *pcol, const int32_t *prow_of,
                         int32_t *rcnt_or_cur, const int32_t *rrow, int32_t *rblk, int32_t *rfine)
{
    if (phase == 0) { if (nnzp > 0) k_hx_tcount<<<blocks_for(nnzp), kT, 0, st>>>(nnzp, pcol, rcnt_or_cur); }
    else if (phase == 1) { if (nnzp > 0) k_hx_tfill<<<blocks_for(nnzp), kT, 0, st>>>(nnzp, pcol, rrow, rcnt_or_cur, rblk); }
    else if (nc > 0) k_hx_tsort<<<blocks_for(nc, 128), 128, 0, st>>>(nc, rrow, rblk, prow_of, rfine);
}
void launch_h_diag(cudaStream_t st, int64_t n, const int32_t *rowptr, const int32_t *colidx, int32_t *diag)
{
    if (n > 0) k_h_diag<<<blocks_for(n), kT, 0, st>>>(n, rowptr, colidx, diag);
}

}}  // namespace mbfea::dev
