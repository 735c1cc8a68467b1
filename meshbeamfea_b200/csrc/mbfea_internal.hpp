// Internal declarations shared by the host-prep, kernel, solver and C-ABI
// translation units of libmbfea.so.  Not installed; the public surface is
// include/mbfea.h.
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

#include "../../include/mbfea.h"

namespace mbfea {

// ---------------------------------------------------------------------------
// Host arrays that are uploaded (hostmem.cpp): page-locked once mbfea_create has found a device, plain heap otherwise.
// ---------------------------------------------------------------------------
void host_arrays_pin_on(int device);
void *host_array_alloc(size_t bytes);
void host_array_free(void *p) noexcept;
bool host_array_is_pinned(const void *p);
template <class T>
struct HostAlloc {
    using value_type = T;
    HostAlloc() = default;
    template <class U> HostAlloc(const HostAlloc<U> &) {}
    T *allocate(size_t n) { return static_cast<T *>(host_array_alloc(n * sizeof(T))); }
    void deallocate(T *p, size_t) noexcept { host_array_free(p); }
    template <class U> bool operator==(const HostAlloc<U> &) const { return true; }
    template <class U> bool operator!=(const HostAlloc<U> &) const { return false; }
};
template <class T> using hvec = std::vector<T, HostAlloc<T>>;

// ---------------------------------------------------------------------------
// Host-side plan: everything derived from connectivity + fixed set.  Pure C++,
// usable without a GPU (mbfea_plan_partition, gloo tests).
// ---------------------------------------------------------------------------
struct Neighbor {
    int32_t rank;
    int64_t recv_offset, recv_count;  // in ghost blocks (ghost region of x)
    int64_t send_offset, send_count;  // in send_rows
};

struct HostPlan {
    int64_t V = 0, E = 0;
    int32_t rank = 0, world = 1;
    int64_t nb_global = 0;            // free vertices
    int64_t row_begin = 0, row_end = 0;
    std::vector<int32_t> free_index;  // V, -1 fixed: INTERNAL block row of each vertex (K3 map composed with perm)
    // Internal renumbering of the reduced system (empty = identity).  The K3 map ranks free vertices
    // in vertex order ("canonical" reduced numbering, what the oracle and the reference use); when
    // that numbering has no locality (vertex ids in arbitrary order) the solver works on a breadth-
    // first renumbering instead: perm[canonical] = internal row, inv[internal] = canonical.  Purely
    // internal: every tap and result is reported in canonical / vertex numbering.
    std::vector<int32_t> perm, inv;
    hvec<int32_t> rowptr;      // nb+1
    hvec<int32_t> colidx;      // nnzb, LOCAL ids: owned [0,nb), ghosts after
    hvec<int32_t> diag_idx;    // nb: position of the diagonal block
    hvec<int32_t> contrib_ptr; // nnzb+1
    hvec<int32_t> contrib;     // (element<<2)|which, ascending element per block
    // Solver format (off-diagonal blocks of the block-Jacobi-scaled matrix, unit diagonal implied):
    // row r stores the blocks [uptr[r], uptr[r+1]) itself -- its ghost columns and, with
    // symmetric-half storage, one of every local pair {r, c} (the upper one, c > r, unless
    // `balanced`) -- and reaches the remaining local
    // columns as TRANSPOSED references [lptr[r], lptr[r+1]) into row c's stored blocks.
    hvec<int32_t> uptr, ucol, usrc;  // usrc: index of the block in the full pattern
    hvec<int32_t> lptr, lcol;
    std::vector<int32_t> lblk;       // lblk: stored-block index holding (lcol, r)
    bool sym_half = false;
    // build_plan scratch, kept between jobs: re-submitting a mesh of the same size then touches no fresh pages
    std::vector<int64_t> scratch_ent, scratch_start, scratch_pos;
    std::vector<int32_t> scratch_gcol, scratch_cnt, scratch_nblk;
    std::vector<std::vector<int32_t>> scratch_bucket;   // owned end points per (element chunk, row range)
    bool balanced = false;   // pairs oriented towards the emptier row instead of the upper one (irregular numberings)
    // Symmetric-half values are stored INTERLEAVED per tile of 32 block rows (sliced ELL): the
    // 16-byte chunk m (0..2) of scalar row i of the k-th stored block of local row lr sits at
    // 16-byte unit  tile_base[tile] + (3k + m) * 192 + 6 lr + i,  so the 192 threads of a CTA
    // read 192 consecutive chunks per load instruction.  lofs[l] = unit of (m = 0, i = 0) of the
    // block a transposed reference points to.  Plain BSR order is kept for full-row storage.
    hvec<int32_t> tile_base;         // #tiles + 1, in 16-byte units
    hvec<int32_t> lofs;
    // Connected components of the reduced system when each occupies a contiguous range of block
    // rows (a batch of independent meshes submitted as one job): comp_ptr[b]..comp_ptr[b+1].
    // Empty when the job is one mesh, the ranges interleave, or a component is too large.
    std::vector<int32_t> comp_ptr;
    int32_t comp_max_rows = 0;
    std::vector<int64_t> ghost_global;// ascending => grouped by owner
    std::vector<int32_t> ghost_owner;
    std::vector<int32_t> send_rows;   // local row ids grouped by destination rank
    std::vector<int32_t> send_dest;   // parallel to send_rows
    std::vector<Neighbor> neighbors;
    int64_t nb() const { return row_end - row_begin; }
    int64_t nnzb() const { return (int64_t)colidx.size(); }
};

// Which 6x6 sub-block of the 12x12 element matrix a contribution is.
enum : int { SUB_00 = 0, SUB_01 = 1, SUB_10 = 2, SUB_11 = 3 };

// Validation in the reference's order; returns MBFEA_* and fills msg.
int validate_job(int64_t V, int64_t E, const int32_t *elems_cm, const float *dims_bh,
                 const float *mat_nuE, int64_t F, const int32_t *fixed, int64_t nF,
                 const int64_t *f_node, const int64_t *f_dof, std::string &msg);
int validate_forces(int64_t V, int64_t nF, const int64_t *f_node, const int64_t *f_dof,
                    std::string &msg);

// Host threads of the calling thread's parallel passes: the node's cores divided by `world` ranks that share them.
void host_threads_share(int world);

// float-quirk section properties [REF src/beamsimulator.cpp:173-188]
// (the four products are float values in the reference; they are kept and uploaded as floats)
void derive_props(int64_t E, const float *dims_bh, const float *mat_nuE, float *props4);

// free_index + nb_global.  fixed must be validated already.
int64_t build_free_map(int64_t V, int64_t F, const int32_t *fixed, std::vector<int32_t> &free_index);

// Block pattern, contributor lists, halo plan for `rank` of `world`.
// with_contrib=false skips the contributor lists (plan-only callers).
// reorder: 0 never, 1 breadth-first when the canonical numbering has poor locality (auto), 2 breadth-first
// always, 3 Morton order of the vertex coordinates (needs nodes_cm; falls back to 2 without)
// The first part alone: free map, internal renumbering (perm / inv, empty = identity) and the rank's row range.
void build_plan_begin(int64_t V, int64_t E, const int32_t *elems_cm, int64_t F, const int32_t *fixed, int32_t rank, int32_t world,
                      HostPlan &plan, int reorder = 0, const double *nodes_cm = nullptr);
// The second part: block pattern, contributor lists and halo plan of the rank, from the state build_plan_begin left.
void build_plan_pattern(int64_t E, const int32_t *elems_cm, bool with_contrib, HostPlan &plan);
// Ghost owners, send lists and neighbour table from the rank's local pattern and its ascending ghost list.
void build_halo_lists(HostPlan &plan);
// after_free_map (optional): called once free_index, nb_global and the row range are final, before the pattern passes
void build_plan(int64_t V, int64_t E, const int32_t *elems_cm, int64_t F, const int32_t *fixed,
                int32_t rank, int32_t world, bool with_contrib, HostPlan &plan, int reorder = 0,
                const double *nodes_cm = nullptr, const std::function<void()> *after_free_map = nullptr);

// Solver format from the full pattern (see HostPlan::uptr).  half=false keeps every
// off-diagonal block in its own row (no transposed references).
void build_solver_format(HostPlan &plan, bool half);

// Fills plan.comp_ptr when the job is a batch of >= min_components independent meshes whose block
// rows are contiguous and no larger than max_rows (single rank only).
void find_components(HostPlan &plan, int32_t max_rows, int32_t min_components);

// Row tiles for the TMA-staged SpMV: tile t covers rows [tile_row[t], tile_row[t+1]).
void build_tiles(const hvec<int32_t> &rowptr, int cap_blocks, int max_rows,
                 hvec<int32_t> &tile_row);

// ---------------------------------------------------------------------------
// Aggregate hierarchy for the multilevel preconditioner (host half, see hierarchy.cpp)
// ---------------------------------------------------------------------------
struct HierLevel {
    int64_t n_fine = 0, n_coarse = 0;
    hvec<int32_t> agg;                 // n_fine: aggregate (= coarse block row) of each fine row
    hvec<double> off;                  // 3 * n_fine: fine position minus the centroid of its aggregate
    std::vector<double> centroid;      // 3 * n_coarse: positions of the coarse rows
    hvec<int32_t> agg_ptr, agg_rows;   // member rows of each aggregate, ascending
    hvec<int32_t> rowptr, colidx;      // coarse 6x6-block pattern (columns ascending, diagonal present)
    hvec<int32_t> gal_ptr, gal_src;    // per coarse block: the fine blocks summed into it, ascending
};
struct Hierarchy { std::vector<HierLevel> levels; };
// nb block rows with pattern (rowptr, colidx < nb) and positions pos_rm[3 * row]; level l bricks have edge
// cell0 * ratio^l; stops at max_levels levels (counting the fine one) or once a level has <= coarsest_rows rows.
void build_hierarchy(int64_t nb, const int32_t *rowptr, const int32_t *colidx, const double *pos_rm, double cell0,
                     double ratio, int max_levels, int64_t coarsest_rows, Hierarchy &H);

// ---------------------------------------------------------------------------
// Multilevel preconditioner, symbolic half (host, once per set_job; hierarchy.cpp).
// Level 0 is the reduced system; level l+1 rows are brick aggregates of level l carrying 6 rigid-body
// unknowns.  The transfer from the fine level is the TENTATIVE prolongator (one rigid block per row) and its
// Galerkin product runs over HierLevel's contributor lists.  Transfers between coarse levels are general
// 6x6-block sparse matrices P (tentative, or smoothed: P = (I - w D^-1 A) T, pattern = A's pattern mapped
// through the aggregates) with the structures the device kernels walk: P's transpose (restriction as a
// gather), the pattern of A P and of the next level's matrix P^T A P.  Columns ascend in every row.
// ---------------------------------------------------------------------------
struct MlTransfer {
    bool smoothed = false;
    int64_t n_fine = 0, n_coarse = 0;
    hvec<int32_t> prow, pcol;           // P: n_fine rows
    std::vector<int32_t> pown;          // n_fine: index in pcol of the block (i, agg i)
    hvec<int32_t> rrow, rblk, rfine;    // per coarse row: the P blocks of that column (ascending) and their fine rows
    hvec<int32_t> aprow, apcol;         // A P: n_fine rows
    hvec<int32_t> crow, ccol, cdiag;    // next level's matrix: n_coarse rows, diagonal position per row
};
struct MlHierarchy {
    std::vector<HierLevel> agg;       // agg[l]: aggregates of level l (l = 0: fine level, with Galerkin contributor lists)
    std::vector<MlTransfer> xfer;     // xfer[l] for l >= 1 (xfer[0] unused: the fine transfer is implicit in agg[0])
    size_t levels() const { return agg.size(); }   // number of coarse levels
};
// smooth_from: first level whose prolongator is smoothed (1 = every coarse-to-coarse transfer; large = none)
// part_bounds / nparts (multi-GPU): row ranges of the ranks; fine-level aggregates then never straddle a rank and are
// numbered rank-major (every rank owns a contiguous range of level-1 rows)
// fine_pattern (optional): the fine aggregates need the positions only, so the build may start before the fine block
// pattern exists; the gate is called once, when the pattern is first needed -- it blocks until the pattern is there,
// stores its arrays (which then replace rowptr / colidx) and returns true, or returns false to abandon the build.
using FinePatternGate = std::function<bool(const int32_t *&rowptr, const int32_t *&colidx)>;
void build_ml_hierarchy(int64_t nb, const int32_t *rowptr, const int32_t *colidx, const double *pos_rm, double cell0,
                        double ratio, int max_levels, int64_t coarsest_rows, int smooth_from, MlHierarchy &H,
                        const int64_t *part_bounds = nullptr, int nparts = 1, const FinePatternGate *fine_pattern = nullptr);

// Flat image of a hierarchy for several ranks: rank 0 builds the hierarchy with every host core, packs it once, the
// image travels H2D + one NCCL broadcast, and every rank fills its device arrays straight from the device copy
// (device-to-device slices: no host unpacking on the receivers).  Header first (host-readable), then the arrays,
// 16-byte aligned.  Level 0 arrays: agg, off, agg_ptr, agg_rows, level-1 rowptr / colidx / diagonal positions, gal_ptr,
// gal_src.  Level l >= 1 (transfer l -> l + 1): agg, off, prow, pcol, rrow, rblk, rfine, aprow, apcol, crow, ccol, cdiag.
constexpr int kMlImageMaxLevels = 16, kMlImageMaxWorld = 16, kMlImageArrays = 12;
struct MlImageHeader {
    int64_t magic, levels, world, group0, nbg, fine_nnz, total_bytes, pad;
    // per rank (world + 1 entries): first fine block of its rows, first level-1 row / block, agg_ptr / gal_ptr at those
    int64_t own_first[kMlImageMaxWorld + 1], agg_begin[kMlImageMaxWorld + 1], blk_begin[kMlImageMaxWorld + 1],
        aggptr_at[kMlImageMaxWorld + 1], galptr_at[kMlImageMaxWorld + 1];
    struct Level {
        int64_t n_fine, n_coarse, smoothed, pad;
        int64_t off[kMlImageArrays], cnt[kMlImageArrays];   // byte offset in the image, element count
    } lv[kMlImageMaxLevels];
};
constexpr int64_t kMlImageMagic = 0x6d6c696d67303031LL;
// fine_rowptr: block row pointer of the GLOBAL fine pattern (nbg + 1); false when the hierarchy does not fit the header
bool ml_image_pack(const MlHierarchy &H, int world, int64_t nbg, const int32_t *fine_rowptr, std::vector<unsigned char> &out);
// The same image as a layout: the header plus (source array, bytes, offset in the image) of every part -- rank 0 copies the
// (page-locked) hierarchy arrays straight into the device image, without assembling it on the host first.
struct MlImageLayout {
    struct Part { const void *p; size_t bytes; size_t at; };
    MlImageHeader hd;
    std::vector<Part> parts;
    hvec<int32_t> diag1;   // diagonal positions of the level-1 pattern (computed here; must outlive the copies)
};
bool ml_image_layout(const MlHierarchy &H, int world, int64_t nbg, const int32_t *fine_rowptr, MlImageLayout &lay);

// Explicit inverse of a dense SPD matrix (row-major n x n) for the coarsest level; 0, or 1 + the index of
// the first non-positive pivot (dense.cpp).
int dense_spd_inverse(int64_t n, const double *A, double *Ainv);

// Scenario files (config.json + ASCII PLY).
struct Scenario {
    int64_t V = 0, E = 0;
    std::vector<double> nodes_cm, normals_cm;
    std::vector<int32_t> elems_cm, fixed;
    std::vector<float> dims_bh, mat_nuE;
    std::vector<int64_t> f_node, f_dof;
    std::vector<double> f_mag;
};
int load_scenario_files(const char *json_path, const char *ply_override, Scenario &sc, std::string &msg);
int load_ply_edge_mesh(const char *ply_path, Scenario &sc, std::string &msg);

}  // namespace mbfea
