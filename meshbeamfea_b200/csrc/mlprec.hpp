// Multilevel (rigid-body aggregate) preconditioner, host driver: owns the level hierarchy on the device, runs
// the numeric setup after every assembly and enqueues the apply z~ = M r~ on a stream (graph-capturable: no
// synchronisation, no allocation).  Kernels: mlprec.cu; symbolic structures: hierarchy.cpp (MlHierarchy).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <memory>
#include <vector>

#include "devbuf.hpp"
#include "kernels.cuh"
#include "mbfea_internal.hpp"
#include "nccl_dl.hpp"

namespace mbfea {

struct MlParams {
    double cell = 2.0;            // brick edge of the first coarse level, in mean beam lengths
    double ratio = 2.0;           // growth of the brick edge per level
    int64_t coarsest_rows = 512;  // stop coarsening at this many block rows (dense inverse of <= 6 x that)
    int smooth_from = 1;          // first level whose prolongator to the next level is smoothed (>= 1)
    int max_levels = 12;
    double omega = 0.6;           // damping of the prolongator smoother, P = (I - omega D^-1 A) T
    double theta = 1.0;           // weight of the coarse correction against the block-Jacobi term, z~ = r~ + theta P0~ (...)
};

struct MlCoarse {                 // one coarse level (l >= 1) and its transfer to level l + 1
    int64_t n = 0, nnz = 0;
    DevBuf<int32_t> rowptr, colidx, row_of, diag_idx;
    DevBuf<double> vals, lfac, sinv, dinv, r, z;
    bool has_next = false, smoothed = false;
    int64_t n_next = 0, nnzp = 0, nnzap = 0;
    DevBuf<int32_t> agg, prow, pcol, prow_of, rrow, rblk, rfine, aprow, apcol, aprow_of;
    DevBuf<double> off, pval, apval;
    DevBuf<float> pval32, dinv32;   // what the apply reads: the same rounded values in restriction and prolongation
    DevBuf<unsigned char> pvalh;    // P as 80-byte blocks of 16-bit floats + scale (multilevel_math.hpp), in place of pval32
    const void *p_apply(bool half16) const { return half16 ? (const void *)pvalh.p : (const void *)pval32.p; }
};

// Several ranks (vertex-block partition): the fine-level transfer is LOCAL to a rank -- aggregates never straddle a
// rank, every rank restricts / prolongs its own rows -- while the coarse levels (1 ..) are REPLICATED: each rank
// assembles its own rows of the level-1 matrix and its slice of the level-1 residual, the slices are exchanged over
// NCCL (matrix: once per assembly; residual: 48 B per aggregate per apply), and every rank then runs the identical
// coarse chain (same kernels, same inputs => same bits, so alpha / beta stay identical across ranks).
struct MlDist {
    int rank = 0, world = 1;
    int64_t row_begin = 0, row_end = 0;                   // own rows in the global internal numbering
    const std::vector<int64_t> *ghost_global = nullptr;   // global rows of the ghost columns (local id nb + g)
    const int32_t *loc_rowptr = nullptr, *loc_colidx = nullptr;   // the rank's local pattern (host; local column ids)
    const nccl::Api *nc = nullptr;
    nccl::Comm comm = nullptr;
};

class MlPrec {
public:
    MlParams par;
    bool symbolic_ok = false, numeric_ok = false;
    // Transfer blocks of the apply in the 16-bit scaled format (MBFEA_ML_HALF=1).  OFF by default: measured on the 119^3
    // lattice it saves 6 % of the solve at an unchanged iteration count (354), but on the thin gridshell (n = 1826) the
    // preconditioner stops working -- the coarse spaces must hold the rigid-body modes to ~1e-7: a relative error eps in T
    // adds eps^2 |K| of spurious energy to modes whose true energy is |K| / kappa, and kappa is 1e9 there.
    bool half16 = false;
    int64_t nb = 0;
    std::vector<int64_t> rows;    // rows of every level, fine first
    double ms_symbolic = 0.0;     // host wall time of the last build_symbolic (hierarchy + uploads)
    int64_t setup_launches = 0;   // kernels of one build_numeric

    void reset();
    // Level hierarchy of a one-rank job from the full block pattern (host arrays, local ids) and the positions of
    // the block rows; uploads every index array.  Returns false (and stays unusable) when the mesh does not coarsen.
    // dist != null: (h_rowptr, h_colidx, pos_rm) describe the GLOBAL reduced system (nb = global rows), dist the rank's share.
    bool build_symbolic(cudaStream_t st, int64_t nb, const int32_t *h_rowptr, const int32_t *h_colidx, const double *pos_rm,
                        double mean_beam_length, const MlDist *dist = nullptr);
    // the host half alone (what rank 0 runs for everybody on several ranks) ...
    void build_hierarchy_host(int64_t nb, const int32_t *h_rowptr, const int32_t *h_colidx, const double *pos_rm,
                              double mean_beam_length, int world, MlHierarchy &H, const FinePatternGate *fine_pattern = nullptr) const;
    // ... and the device half from a prebuilt hierarchy: fine_nnz = blocks of the (global) fine pattern, own_first_block =
    // index of the first block of the rank's first row in it (0 on one rank)
    bool upload_hierarchy(cudaStream_t st, const MlHierarchy &H, int64_t nb_global, int64_t fine_nnz, int64_t own_first_block,
                          const MlDist *dist);
    // One rank: the whole symbolic half built ON THE DEVICE (hier_dev.cu) from the mesh and the fine pattern that are
    // already there -- the same arrays as build_hierarchy_host + upload_hierarchy, entry for entry.  node4 / fm: vertex
    // records and the vertex -> block row map (-1 fixed); (d_rowptr, d_colidx): fine pattern, fine_nnz blocks.
    // false: nothing usable was built (mesh does not coarsen, brick grid too sparse for the counting sort, coarsest
    // level too large) -- the caller falls back to the host builder.
    bool build_hierarchy_device(cudaStream_t st, int64_t V, const dev::Node4 *d_node4, const int32_t *d_fm, int64_t nb,
                                const int32_t *d_rowptr, const int32_t *d_colidx, int64_t fine_nnz, double mean_beam_length,
                                int nparts = 1);
    // Several ranks, rank 0: after build_hierarchy_device(.., nparts = world) on the GLOBAL system, the flat image of
    // ml_image_pack assembled from the device arrays (device-to-device) for the broadcast; hd: its header (host copy)
    bool pack_image_device(cudaStream_t st, int world_, int64_t nbg, const int32_t *d_fine_rowptr, int64_t fine_nnz,
                           DevBuf<unsigned char> &dimg, MlImageHeader &hd);
    // parity tap: the device arrays of the current hierarchy against a host hierarchy (mismatch counts per array, see api.cu)
    void compare_with_host(cudaStream_t st, const MlHierarchy &H, int64_t fine_nnz, std::vector<int64_t> &bad) const;
    // several ranks: the device arrays straight from the device copy of rank 0's image (ml_image_pack): device-to-device
    // slices of level 0 (own rows, own aggregates, own level-1 block rows) and copies of the replicated coarse levels
    bool upload_from_image(cudaStream_t st, const unsigned char *d_img, const MlImageHeader &hd, const MlDist &dist);
    // Galerkin matrices, smoothers, prolongators, coarsest inverse from the assembled K (full pattern, unscaled) and
    // the Cholesky factors of its diagonal blocks.  *d_flag |= 1/2 on a non-SPD coarse block / coarsest pivot.
    void build_numeric(cudaStream_t st, const int32_t *d_rowptr, const int32_t *d_colidx, const double *d_vals,
                       const double *d_lfac, int *d_flag);
    // zt = M rt, part[0 .. grid) = partials of rt . zt
    // sc / pc (multi-GPU peer path, inside the CG iteration): the last kernel publishes this rank's r~.z~ through the mailboxes
    void apply(cudaStream_t st, const double *rt, double *zt, double *part, int grid, const dev::PcgScalars *sc = nullptr,
               const dev::PeerCtx *pc = nullptr);
    int kernels_per_apply() const { return lv.empty() ? 0 : (int)(2 * lv.size() + 1 + (world > 1 ? 1 : 0)); }
    size_t levels() const { return lv.size(); }

private:
    void alloc_level_buffers(cudaStream_t st, MlCoarse &c, bool has_next);
    // multi-rank: own slice of level 1 (aggregates [agg_begin[rank], agg_begin[rank + 1]), their matrix blocks)
    int rank = 0, world = 1;
    const nccl::Api *nc = nullptr;
    nccl::Comm comm = nullptr;
    std::vector<int64_t> agg_begin, blk_begin;   // per rank: first level-1 row / first level-1 matrix block (world + 1 entries)
    void exchange_slices(cudaStream_t st, double *buf, const std::vector<int64_t> &begin, int width);
    DevBuf<int32_t> agg0, agg_ptr0, agg_rows0, gal_ptr0, gal_src0, frow0;
    DevBuf<double> off0;
    DevBuf<int64_t> ghost_stage;
    // scratch of build_hierarchy_device, kept between jobs
    DevBuf<double> hd_pos[2], hd_part, hd_box;
    DevBuf<int32_t> hd_key, hd_bcnt, hd_bcur, hd_occ, hd_aggid, hd_bstart, hd_scan, hd_len, hd_aggptr, hd_aggrows, hd_cand, hd_small;
    DevBuf<float> t0;              // fine transfer T_v = L_v^T B(d_v), FP32 (with half16: scratch of the packing pass)
    DevBuf<unsigned char> t0h;     // the same as 80-byte 16-bit blocks: what the apply reads
    const void *t0_apply() const { return half16 ? (const void *)t0h.p : (const void *)t0.p; }
    int group0 = 32;
    std::vector<std::unique_ptr<MlCoarse>> lv;   // lv[m]: level m + 1
    std::vector<std::unique_ptr<MlCoarse>> spare; // level objects of the previous job, in level order (buffers keep their capacity)
    std::unique_ptr<MlCoarse> take_level();
    DevBuf<double> dense, gj_col;                // explicit inverse of the coarsest matrix (+ Gauss-Jordan scratch)
    DevBuf<float> dense32;                       // its symmetrised FP32 copy: what the apply multiplies with
    int64_t n_dense = 0;
};

}  // namespace mbfea
