// Multilevel (rigid-body aggregate) preconditioner, host driver: owns the level hierarchy on the device, runs
// the numeric setup after every assembly and enqueues the apply z~ = M r~ on a stream (graph-capturable: no
// synchronisation, no allocation).  Kernels: mlprec.cu; symbolic structures: hierarchy.cpp (MlHierarchy).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <memory>
#include <vector>

#include "devbuf.hpp"
#include "mbfea_internal.hpp"

namespace mbfea {

struct MlParams {
    double cell = 2.0;            // brick edge of the first coarse level, in mean beam lengths
    double ratio = 2.0;           // growth of the brick edge per level
    int64_t coarsest_rows = 512;  // stop coarsening at this many block rows (dense inverse of <= 6 x that)
    int smooth_from = 1;          // first level whose prolongator to the next level is smoothed (>= 1)
    int max_levels = 12;
    double omega = 0.6;           // damping of the prolongator smoother, P = (I - omega D^-1 A) T
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
};

class MlPrec {
public:
    MlParams par;
    bool symbolic_ok = false, numeric_ok = false;
    int64_t nb = 0;
    std::vector<int64_t> rows;    // rows of every level, fine first
    double ms_symbolic = 0.0;     // host wall time of the last build_symbolic (hierarchy + uploads)
    int64_t setup_launches = 0;   // kernels of one build_numeric

    void reset();
    // Level hierarchy of a one-rank job from the full block pattern (host arrays, local ids) and the positions of
    // the block rows; uploads every index array.  Returns false (and stays unusable) when the mesh does not coarsen.
    bool build_symbolic(cudaStream_t st, int64_t nb, const int32_t *h_rowptr, const int32_t *h_colidx, const double *pos_rm,
                        double mean_beam_length);
    // Galerkin matrices, smoothers, prolongators, coarsest inverse from the assembled K (full pattern, unscaled) and
    // the Cholesky factors of its diagonal blocks.  *d_flag |= 1/2 on a non-SPD coarse block / coarsest pivot.
    void build_numeric(cudaStream_t st, const int32_t *d_rowptr, const int32_t *d_colidx, const double *d_vals,
                       const double *d_lfac, int *d_flag);
    // zt = M rt, part[0 .. grid) = partials of rt . zt
    void apply(cudaStream_t st, const double *rt, double *zt, double *part, int grid);
    int kernels_per_apply() const { return lv.empty() ? 0 : (int)(2 * lv.size() + 1); }
    size_t levels() const { return lv.size(); }

private:
    DevBuf<int32_t> agg0, agg_ptr0, agg_rows0, gal_ptr0, gal_src0, frow0;
    DevBuf<double> off0;
    DevBuf<float> t0;
    int group0 = 32;
    std::vector<std::unique_ptr<MlCoarse>> lv;   // lv[m]: level m + 1
    DevBuf<double> dense, gj_col;                // explicit inverse of the coarsest matrix (+ Gauss-Jordan scratch)
    DevBuf<float> dense32;                       // its symmetrised FP32 copy: what the apply multiplies with
    int64_t n_dense = 0;
};

}  // namespace mbfea
