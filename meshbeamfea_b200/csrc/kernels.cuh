// Launch wrappers of the hand-written sm_100a kernels (definitions in
// kernels.cu).  All launches go to the caller's stream; none synchronises.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "element.cuh"

namespace mbfea { namespace dev {

constexpr int kVecThreads = 256;
constexpr int kMaxPartials = 148 * 16;  // fixed upper bound of reduction grids

// TMA-staged SpMV tile geometry (see kernels.cu)
constexpr int kTmaCapBlocks = 224;  // 6x6 blocks per stage: 224*288 B = 63 KB
constexpr int kTmaMaxRows = 32;     // block rows per tile  -> 192 consumer threads
constexpr int kTmaStages = 3;

constexpr int kMaxWorld = 16;
constexpr int kTBlock16Bytes = 80;   // 16-bit transfer block of the multilevel apply (multilevel_math.hpp)

// ---- multi-GPU over NVLink peer memory (no NCCL on the PCG data path) --------
// One Mailbox per rank, in device memory that every peer maps through CUDA IPC.
// Producers write values into the CONSUMER's mailbox with plain peer stores,
// fence at system scope, then store a tag; consumers spin on their own (local)
// mailbox with acquire loads.  tag = (solve epoch << 32) | iteration stamp.
struct Mailbox {
    unsigned long long halo_tag[kMaxWorld];    // [q]: rank q's boundary x-blocks have landed in my ghost region
    // classic CG: [0] p.Ap, [1] rho = r~.r~, [2] per-batch ||r||^2 -- two all-to-all exchanges per
    // iteration interlock, so one slot each is safe.  Single-reduction CG has ONE exchange per
    // iteration and a rank may run one exchange ahead of a non-neighbour's consumption: its slots are
    // double-buffered by iteration parity: tag [3 + par], values [3 + 2 par] (r.K~r) and [4 + 2 par] (r.r)
    unsigned long long dot_tag[5][kMaxWorld];
    double dot_val[7][kMaxWorld];
    // single-reduction CG, fence-free exchange: each double travels as two 8-byte words
    // (flag << 32 | half of the bit pattern); an 8-byte store is never torn, so a word whose flag
    // matches carries valid data and the sender needs no fence (and no round trip) before signalling.
    // [iteration parity][0 r.K~r low, 1 r.K~r high, 2 r.r low, 3 r.r high][rank]
    unsigned long long ll[2][4][kMaxWorld];
};

struct PeerCtx {          // device-resident, local memory
    int rank, world, n_neighbors, pad;
    int neighbor[kMaxWorld];
    Mailbox *mine;
    Mailbox *peer[kMaxWorld];   // peer[rank] == mine
    unsigned int *tickets;      // last-CTA-done tickets: [0] spmv [1] update [2] push_halo [3] check
};

struct PcgScalars {
    double rz[2];        // r.z, double-buffered by iteration parity
    double ff;           // f.f
    double rr;           // r.r of the last finished iteration
    double pAp;          // last p.Ap (diagnostic)
    long long iter;      // finished iterations
    long long max_iter;
    double tol2;         // tol^2
    int done;            // 0 running, 1 converged, 2 max_iter, 3 breakdown
    int pad;
    unsigned long long epoch;    // solve counter (multi-GPU tags)
    unsigned long long tag_cur;  // tag of the values produced in the current iteration
    unsigned long long *trace;   // optional: kTraceSlots globaltimer stamps per iteration (first kTraceIters)
    double rho0;                 // r~.r~ at iteration 0
    double alpha[2];             // single-reduction CG: alpha of the last two iterations (by parity)
    long long restart_iter;      // iteration index of the last residual replacement (its step uses beta = 0)
    long long checks;            // residual checks done (multi-GPU tag of the per-batch check)
};
constexpr int kTraceSlots = 8;   // 0 spmv in, 1 spmv past halo wait, 2 update in, 3 update past wait,
                                 // 4 pupdate in, 5 pupdate past wait, 6 push in, 7 push out (last CTA)
constexpr int kTraceIters = 256;

// where a kernel finds the partial sums of a dot product: n values, stride apart
struct PartialRef { const double *buf; int n; int stride; };

void launch_pack_nodes(cudaStream_t st, int64_t V, const double *nodes_cm, Node4 *node4);
void launch_pack_elems(cudaStream_t st, int64_t E, const int32_t *elems_cm, const double *normals_cm,
                       const float *props4, ElemRec *rec);
// K3: device constraint map (flags -> exclusive scan)
void launch_free_map(cudaStream_t st, int64_t V, int64_t F, const int32_t *fixed, int32_t *free_index,
                     int32_t *scratch /* >= V + V/1024 + 1 ints */);
// composes the canonical map with the internal renumbering of the reduced system
void launch_apply_perm(cudaStream_t st, int64_t V, const int32_t *canon, const int32_t *perm, int32_t *out);
// K1: E x 12 x 12 element matrices (parity tap)
void launch_kelem(cudaStream_t st, int64_t E, const ElemRec *rec, const Node4 *node4, double *Ke);
// K1: frame + stiffness coefficients of every beam -> ElemGeom records
void launch_elem_geom(cudaStream_t st, int64_t E, const ElemRec *rec, const Node4 *node4, ElemGeom *geom);
// K2: deterministic scatter-free BSR assembly from the ElemGeom records
void launch_assemble(cudaStream_t st, int64_t nnzb, const int32_t *contrib_ptr, const int32_t *contrib,
                     const ElemGeom *geom, double *vals);
// K5: Cholesky factor L and S = L^-1 of every diagonal block; *flag |= 1 on a non-positive pivot
void launch_block_chol(cudaStream_t st, int64_t nb, const int32_t *diag_idx, const double *vals, double *lfac,
                       double *sinv, int *flag);
// solver matrix K~ = S K S^T, stored blocks only (sinv: owned rows then ghost rows)
void launch_scale_blocks(cudaStream_t st, int64_t nb, const int32_t *uptr, const int32_t *ucol, const int32_t *usrc,
                         const double *vals, const double *sinv, double *vals_s, const int32_t *tile_base);
// K4: y = A x (x has nb + n_ghost blocks), partial[blockIdx] = sum x_own . y
//     identity: unit diagonal implied; lptr/lcol/lblk: transposed references (may be null)
int  spmv_rowthread_grid(int64_t nb, bool sym, bool peer = false);   // peer: launched with a PeerCtx
void launch_spmv_rowthread(cudaStream_t st, int64_t nb, const int32_t *uptr, const int32_t *ucol, const double *vals,
                           const int32_t *lptr, const int32_t *lcol, const int32_t *lofs, const int32_t *tile_base,
                           int identity, const double *x, double *y, double *partial, double *partial2,
                           const PcgScalars *sc, const PeerCtx *pc);
int  spmv_tma_grid(int64_t ntiles, int sm_count);
size_t spmv_tma_smem_bytes();
int  spmv_tma_prepare();  // opt in to the large dynamic shared memory; returns cudaError
void launch_spmv_tma(cudaStream_t st, int grid, int64_t ntiles, const int32_t *tile_row,
                     const int32_t *rowptr, const int32_t *colidx, const double *vals, int identity,
                     const double *x, double *y, double *partial, const PcgScalars *sc);
// K6: CG vector kernels on the scaled system
int  vec_grid(int64_t nb);
int  cg_vec_grid(int64_t nb);
void launch_cg_init(cudaStream_t st, int64_t nb, const double *f, const double *sinv, double *y, double *rt,
                    double *ft, double *p, double *part_rho, double *part_ff);
void launch_cg_init_finish(cudaStream_t st, PartialRef rho, PartialRef ff, PcgScalars *sc, double tol,
                           long long max_iter, unsigned long long epoch, unsigned long long *trace);
void launch_cg_update(cudaStream_t st, int64_t nb, int parity, PartialRef pAp, const double *p, const double *Ap,
                      double *y, double *rt, double *part_rho, PcgScalars *sc, const PeerCtx *pc,
                      int publish_rho = 1 /* peer path: the last CTA all-gathers r~.r~ (0: a preconditioner's rho follows) */);
void launch_cg_pupdate(cudaStream_t st, int64_t nb, int parity, PartialRef rho, const double *rt, double *p,
                       PcgScalars *sc, const PeerCtx *pc, const int32_t *push_idx, double *const *push_dst,
                       unsigned int *ticket /* last-CTA ticket: the CTA that retires last advances the scalars */);
// single-reduction (Chronopoulos-Gear) CG: SpMV on the residual yields r.r and r.K~r, one step kernel does the rest
void launch_cg2_init(cudaStream_t st, int64_t nb, const double *f, const double *sinv, double *y, double *r, double *ft,
                     double *d, double *s, double *part_ff);
void launch_cg2_step(cudaStream_t st, int64_t nb, PartialRef delta, PartialRef gamma, double *r, const double *w,
                     double *d, double *s, double *y, PcgScalars *sc, const PeerCtx *pc, const int32_t *push_idx,
                     double *const *push_dst, const int32_t *push_rows, int n_push, unsigned int *ticket);
// unscaled residual norm ||L (a - b)||^2 (b may be null) -> partials; finish: total -> sc->rr, stop test
void launch_cg_check(cudaStream_t st, int64_t nb, const double *lfac, const double *a, const double *b, double *part,
                     const PcgScalars *sc, const PeerCtx *pc);
void launch_cg_check_finish(cudaStream_t st, PartialRef part, PcgScalars *sc, const PeerCtx *pc);
// residual replacement from the assembled matrix: res := S (f - Kx), dir := dir_scale * res, aux_zero := 0;
// partials of ||f - Kx||^2 and res.res
void launch_cg_restart(cudaStream_t st, int64_t nb, const double *sinv, const double *f, const double *Kx, double *res,
                       double *dir, double dir_scale, double *aux_zero, double *part_rr, double *part_rho);
void launch_cg_restart_finish(cudaStream_t st, PartialRef rr, PartialRef rho, PcgScalars *sc, unsigned long long epoch,
                              int clear_done);
// batched CG: one CTA per connected component (contiguous block rows), vectors in shared memory
int  cg_batched_prepare(size_t smem_bytes);
void launch_cg_batched(cudaStream_t st, int ncomp, size_t smem_bytes, const int32_t *comp_ptr, const int32_t *uptr,
                       const int32_t *ucol, const double *vals, const double *f, const double *sinv, const double *lfac,
                       double *y, double tol, int max_iter, int check_every, int32_t *comp_iters, double *comp_rr,
                       double *comp_ff, int32_t *comp_status);
void launch_unscale(cudaStream_t st, int64_t nb, const double *sinv, const double *y, double *x);
void set_pdl(bool on);    // programmatic dependent launch of the CG-loop kernels (default off; MBFEA_PDL=1 turns it on)
int peer_timeout_flag();  // 1 once any peer wait gave up (a rank died)
void peer_timeout_clear(cudaStream_t st);
void set_sm_count(int n); // grids are sized in resident waves of the device's SM count
int  sm_count();
// P2P halo: owned boundary x-blocks -> the neighbours' ghost slots (peer stores), then the tag
void launch_push_halo(cudaStream_t st, int64_t n_send, const int32_t *send_rows, double *const *dst,
                      const double *x, const PcgScalars *sc, const PeerCtx *pc);
// collapse `n` partials to out[0] in fixed order (multi-GPU path, before the all-gather)
void launch_reduce_partials(cudaStream_t st, const double *a, int n, double *out_a, const double *b,
                            double *out_b, const PcgScalars *sc);
// halo: pack owned boundary blocks (width doubles each) into the send buffer
void launch_pack_halo(cudaStream_t st, int64_t n_send, const int32_t *send_rows, const double *x, double *sendbuf,
                      int width);
// deterministic test pattern a[i] = v0 + dv*(i % period)  (SpMV timing input)
void launch_fill(cudaStream_t st, int64_t n, double *a, double v0, double dv, int period);
// load vector of the reduced system (owned rows), duplicates accumulate in input order
void launch_scatter_forces(cudaStream_t st, int64_t nF, const int64_t *node, const int64_t *dof, const double *mag,
                           const int32_t *free_index, int64_t row_begin, int64_t nb, double *f,
                           bool unique /* no destination is hit twice: one thread per force */);
// results
void launch_expand(cudaStream_t st, int64_t V, const int32_t *free_index, const double *x_global,
                   double *U_cm, double *U_rm);
void launch_colmajor_to_rowmajor(cudaStream_t st, int64_t V, const double *U_cm, double *U_rm);
// K7: element end forces in the SimulationResults layout (6 x 2E)
void launch_element_forces(cudaStream_t st, int64_t E, const ElemRec *rec, const Node4 *node4,
                           const double *U_rm, double *F6x2E);

// colour-mapping pre-pass: per-component min/max of the 6 x n force table, and its normalisation
int  force_minmax_parts(int64_t n);
void launch_force_ranges(cudaStream_t st, int64_t n, const double *F, double *pmin, double *pmax, double *range);
void launch_force_normalize(cudaStream_t st, int64_t n, const double *F, const double *range, double *out);


// ---- block pattern, contributor lists and solver format built on the device (plan_dev.cu) -------------------
// exclusive scan of n int32 values: out[0 .. n] (out[n] = total); tsum: scan_i32_scratch(n) ints of scratch
void launch_scan_i32(cudaStream_t st, int64_t n, const int32_t *in, int32_t *out, int32_t *tsum);
int64_t scan_i32_scratch(int64_t n);
void launch_pd_count(cudaStream_t st, int64_t E, const int32_t *elems_cm, const int32_t *fm, int32_t *cnt);
void launch_pd_fill(cudaStream_t st, int64_t E, const int32_t *elems_cm, const int32_t *fm, const int32_t *start, int32_t *cur,
                    unsigned long long *ent);
void launch_pd_sort(cudaStream_t st, int64_t nb, const int32_t *start, unsigned long long *ent, int32_t *nblk);
void launch_pd_emit(cudaStream_t st, int64_t nb, const int32_t *start, const unsigned long long *ent, const int32_t *rowptr,
                    int32_t *colidx, int32_t *diag_idx, int32_t *contrib_ptr, int32_t *contrib);
void launch_pf_count(cudaStream_t st, int64_t nb, const int32_t *rowptr, const int32_t *colidx, int32_t *nu, int32_t *nl);
// the rank's share [r0, r1) of a pattern of the whole system: ghost marks over the blocks [k0, k1), ghost list, local ids
void launch_pl_mark(cudaStream_t st, int64_t k0, int64_t k1, const int32_t *gcol, int64_t r0, int64_t r1, int32_t *mark);
void launch_pl_ghosts(cudaStream_t st, int64_t nbg, const int32_t *mark, const int32_t *gidx, long long *ghost_global);
void launch_pl_localize(cudaStream_t st, int64_t k0, int64_t k1, const int32_t *gcol, int64_t r0, int64_t r1, const int32_t *gidx,
                        int32_t *colidx_l);
void launch_pl_shift(cudaStream_t st, int64_t n, const int32_t *src, int32_t delta, int32_t *dst);
void launch_pf_tiles(cudaStream_t st, int64_t nb, const int32_t *nu, int32_t *tunits, unsigned long long *totals);
void launch_pf_fill(cudaStream_t st, int64_t nb, const int32_t *rowptr, const int32_t *colidx, const int32_t *diag_idx,
                    const int32_t *uptr, const int32_t *lptr, const int32_t *tile_base, int32_t *ucol, int32_t *usrc, int32_t *lcol,
                    int32_t *lofs);

// ---- symbolic half of the multilevel preconditioner built on the device (hier_dev.cu) ------------------------
void launch_h_pos(cudaStream_t st, int64_t V, const Node4 *node4, const int32_t *fm, double *pos);
void launch_h_box(cudaStream_t st, int64_t n, const double *pos, double *part /* 6 * 256 */, double *box /* min[3], max[3] */);
void launch_h_brick(cudaStream_t st, int64_t n, const double *pos, const double lo[3], double cell, const int64_t cnt[3], int32_t *key,
                    int32_t *bcnt, const int64_t *part_bounds = nullptr, int nparts = 1);
void launch_h_occ(cudaStream_t st, int64_t nk, const int32_t *bcnt, int32_t *occ);
void launch_h_assign(cudaStream_t st, int64_t n, const int32_t *key, const int32_t *aggid, const int32_t *bstart, int32_t *bcur,
                     int32_t *agg, int32_t *agg_rows);
void launch_h_aggptr(cudaStream_t st, int64_t nk, const int32_t *bcnt, const int32_t *aggid, const int32_t *bstart, int32_t *agg_ptr,
                     int64_t n, int64_t na);
void launch_h_sort_members(cudaStream_t st, int64_t na, const int32_t *agg_ptr, int32_t *agg_rows, int32_t *maxsz);
void launch_h_centroid_off(cudaStream_t st, int64_t n, int64_t na, const int32_t *agg_ptr, const int32_t *agg_rows, const int32_t *agg,
                           const double *pos, double *centroid, double *off);
void launch_h0_ub(cudaStream_t st, int64_t na, const int32_t *agg_ptr, const int32_t *agg_rows, const int32_t *rowptr, int32_t *ub);
void launch_h0_cand(cudaStream_t st, int64_t na, const int32_t *agg_ptr, const int32_t *agg_rows, const int32_t *rowptr,
                    const int32_t *colidx, const int32_t *agg, const int32_t *cstart, int32_t *cand, int32_t *nuniq);
void launch_h0_emit(cudaStream_t st, int64_t na, const int32_t *cstart, const int32_t *cand, const int32_t *rowptr1, int32_t *colidx1,
                    int32_t *diag1);
void launch_h0_gal(cudaStream_t st, bool fill, int64_t na, const int32_t *agg_ptr, const int32_t *agg_rows, const int32_t *rowptr,
                   const int32_t *colidx, const int32_t *agg, const int32_t *rowptr1, const int32_t *colidx1, int32_t *cnt_or_cur,
                   const int32_t *gal_ptr, int32_t *gal_src);
// rows of P / A P / P^T A P as sorted unique unions (fill = false: row lengths; true: columns at rowptr)
void launch_hx_p_rows(cudaStream_t st, bool fill, int64_t n, int64_t ncols, bool smoothed, const int32_t *arow, const int32_t *acol,
                      const int32_t *agg, int32_t *rowlen, const int32_t *prow, int32_t *pcol);
void launch_hx_ap_rows(cudaStream_t st, bool fill, int64_t n, int64_t ncols, const int32_t *arow, const int32_t *acol, const int32_t *prow,
                       const int32_t *pcol, int32_t *rowlen, const int32_t *aprow, int32_t *apcol);
void launch_hx_rap_rows(cudaStream_t st, bool fill, int64_t nc, const int32_t *rrow, const int32_t *rfine, const int32_t *aprow,
                        const int32_t *apcol, int32_t *rowlen, const int32_t *crow, int32_t *ccol);
// phase 0: entries per column (rcnt), 1: block ids into the columns' segments (cursor), 2: sort each segment, fine rows
void launch_hx_transpose(cudaStream_t st, int phase, int64_t nnzp, int64_t nc, const int32_t *pcol, const int32_t *prow_of,
                         int32_t *rcnt_or_cur, const int32_t *rrow, int32_t *rblk, int32_t *rfine);
void launch_h_diag(cudaStream_t st, int64_t n, const int32_t *rowptr, const int32_t *colidx, int32_t *diag);

// ---- multilevel preconditioner, production kernels (mlprec.cu) ----------------------------------------------
void launch_ml_make_t0(cudaStream_t st, int64_t nb, const double *lfac, const double *off, float *t0);
void launch_ml_diag_inv(cudaStream_t st, int64_t n, const double *sinv, double *dinv, float *dinv32);
void launch_ml_galerkin0(cudaStream_t st, int64_t n_coarse_blocks, const int32_t *gal_ptr, const int32_t *gal_src,
                         const int32_t *frow, const int32_t *fcol, const double *fvals, const double *off, double *cvals);
void launch_ml_dense_to_f32(cudaStream_t st, int n, const double *a, float *out);
void launch_ml_dense_matvec(cudaStream_t st, int n, const float *ainv, const double *r, double *z);
void launch_ml_to_f32(cudaStream_t st, int64_t n, const double *a, float *out);
// restrictions: only the fine rows [fine_lo, fine_hi) contribute; prolongations: rows [row0, row0 + n)
void launch_ml_restrict_tent(cudaStream_t st, int64_t nc, const int32_t *rrow, const int32_t *rfine, const double *off,
                             const double *r, double *rc, int64_t fine_lo, int64_t fine_hi);
void launch_ml_smooth_prolong_tent(cudaStream_t st, int64_t row0, int64_t n, const float *dinv, const int32_t *agg, const double *off,
                                   const double *r, const double *zc, double *z);
void launch_ml_row_of(cudaStream_t st, int64_t n, const int32_t *rowptr, int32_t *row_of);
void launch_ml_shift_i32(cudaStream_t st, int64_t n, const int32_t *src, int32_t delta, int32_t *dst);
void launch_ml_gather3(cudaStream_t st, int64_t n, const double *src, const int64_t *idx, double *dst);
void launch_ml_make_p(cudaStream_t st, int64_t n, int64_t nnzp, const int32_t *prow, const int32_t *pcol, const int32_t *prow_of,
                      const int32_t *arow, const int32_t *acol, const double *avals, const int32_t *agg, const double *off,
                      const double *dinv, double omega, double *pval);
void launch_ml_ap(cudaStream_t st, int64_t n, int64_t nnzc, const int32_t *crow, const int32_t *ccol, const int32_t *crow_of,
                  const int32_t *arow, const int32_t *acol, const double *avals, const int32_t *prow, const int32_t *pcol,
                  const double *pval, double *cval);
void launch_ml_rap(cudaStream_t st, int64_t nc, int64_t nnzg, const int32_t *grow, const int32_t *gcol, const int32_t *grow_of,
                   const int32_t *rrow, const int32_t *rblk, const int32_t *rfine, const double *pval, const int32_t *crow,
                   const int32_t *ccol, const double *cval, double *gval);
void launch_ml_blocks_to_dense(cudaStream_t st, int64_t nrows, int64_t nnz, const int32_t *rowptr, const int32_t *colidx,
                               const int32_t *row_of, const double *vals, double *dense);
void launch_ml_dense_inverse(cudaStream_t st, int n, double *a, double *colk, int *flag);
// transfer blocks (t0, pval): half = 80-byte blocks of 16-bit floats + scale (launch_ml_pack16), else 144-byte FP32 blocks
void launch_ml_pack16(cudaStream_t st, int64_t nblk, const float *src, void *dst);
void launch_ml_pack16(cudaStream_t st, int64_t nblk, const double *src, void *dst);
void launch_ml_restrict0(cudaStream_t st, int64_t n_agg, int group, const int32_t *agg_ptr, const int32_t *agg_rows,
                         const void *t0, bool half, const double *r, double *rc);
void launch_ml_finish0(cudaStream_t st, int grid, int64_t nb, const int32_t *agg, const void *t0, bool half, const double *rt,
                       const double *z1, double *zt, double *part, const PcgScalars *sc = nullptr,
                       const PeerCtx *pc = nullptr /* kernels.cu: peer path publishes r~.z~ through the mailboxes */,
                       double theta = 1.0 /* weight of the coarse correction */);
void launch_ml_restrict(cudaStream_t st, int64_t nc, int64_t nnzp, const int32_t *rrow, const int32_t *rblk, const int32_t *rfine,
                        const void *pval, bool half, const double *r, double *rc, int64_t fine_lo, int64_t fine_hi);
void launch_ml_smooth_prolong(cudaStream_t st, int64_t row0, int64_t n, int64_t n_level, int64_t nnzp, const float *dinv,
                              const int32_t *prow, const int32_t *pcol, const void *pval, bool half, const double *r, const double *zc,
                              double *z);

}}  // namespace mbfea::dev
