/*
 * mbfea.h -- C ABI of libmbfea.so, the B200-native replacement for the
 * simulation hot path of IasonManolas/MeshBeamFEA.
 *
 * The reference has no FFI: the path sits behind the C++ class BeamSimulator
 * [REF src/beamsimulator.hpp:78-95], whose executeSimulation() forwards to
 * fea::solve() of the un-vendored threed-beam-fea library
 * [REF src/beamsimulator.cpp:40-69].  This header is what a binding for that
 * seam binds: plain pointers and sizes, no C++/torch types.  The C++17 shim
 * meshbeamfea_b200/cpp/beamsimulator.hpp re-exposes the reference's own class
 * and struct names on top of it (INTEGRATION.md shows the two-line swap).
 *
 * Conventions
 *  - Matrices are COLUMN-MAJOR, exactly the memory image of the reference's
 *    Eigen members (Eigen::MatrixX3d nodes, MatrixX2i elements, MatrixX3d
 *    elementalNormals, [REF src/beamsimulator.hpp:35-43]).
 *  - DOF order per vertex: ux,uy,uz,rx,ry,rz; global DOF = 6*vertex+dof
 *    [REF src/beamsimulator.cpp:139-150].
 *  - Every function returns an mbfea_status; mbfea_last_error() gives text.
 *  - There is NO CPU fallback: without a usable CUDA device every compute entry
 *    point returns MBFEA_ECUDA.
 */
#ifndef MBFEA_H
#define MBFEA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define MBFEA_VERSION 100 /* 0.1.0 */

typedef enum mbfea_status {
    MBFEA_OK = 0,
    /* where the reference would hit a GSL Expects() and std::terminate:
     * empty job [REF src/beamsimulator.cpp:41], size mismatches [:73-75],
     * element index out of range [:81-83], b,h<=0 or nu outside [-1,0.5]
     * [REF src/beam.hpp:14,25], dof outside [0,6) [REF src/utilities.hpp:77] */
    MBFEA_EPRECOND = 1,
    /* std::runtime_error("Vertex index is outside of valid range.")
     * [REF src/beamsimulator.cpp:134-137] */
    MBFEA_ERANGE_FIXED = 2,
    /* std::range_error("Node index is out of valid range.")
     * [REF src/beamsimulator.cpp:164-166] */
    MBFEA_ERANGE_FORCE = 3,
    MBFEA_ECUDA = 4,      /* CUDA / NCCL runtime failure, or no device        */
    MBFEA_ESTATE = 5,     /* call order violated (e.g. execute before set_job) */
    MBFEA_ENOCONV = 6,    /* PCG hit max_iter before reaching tol (results are
                             still readable; stats carry the residual)         */
    MBFEA_ESINGULAR = 7,  /* breakdown: non-SPD system (no fixed vertex,
                             isolated vertex, zero-length beam ...)            */
    MBFEA_EIO = 8,        /* scenario / PLY / CSV file problem                 */
    MBFEA_EARG = 9        /* NULL pointer or nonsensical argument              */
} mbfea_status;

typedef struct mbfea_ctx mbfea_ctx; /* opaque: device buffers, stream, graphs, comm */

typedef struct mbfea_options {
    int32_t device;        /* CUDA ordinal; -1 = current device                */
    int32_t verbose;       /* reference sets opts.verbose=true [REF :57]; here
                              stdout logging is opt-in                         */
    double  tol;           /* PCG stop: ||f - K u||_2 <= tol*||f||_2 (1e-8); tested on
                              the recurrence once per batch, verified on the
                              recomputed residual at the end                   */
    int64_t max_iter;      /* 0 = 20*ndof                                       */
    int32_t check_every;   /* PCG iterations per captured CUDA graph = per residual
                              check; the host polls once per graph (0 = auto)  */
    int32_t spmv_kernel;   /* 0 auto (= 1), 1 row-thread LDG kernel, 2 TMA-staged */
    int32_t rank;          /* vertex-block partition: this process' rank ...   */
    int32_t world;         /* ... of `world` processes (1 = single GPU)        */
    int32_t use_graph;     /* 1 = capture the PCG iteration in a CUDA graph     */
    int32_t comm_mode;     /* world > 1: 0 auto (NVLink peer memory when every
                              peer is reachable, else NCCL), 1 NCCL send/recv +
                              all-gather, 2 peer memory required               */
    int32_t storage;       /* solver matrix: 0 auto, 1 every off-diagonal block in
                              its own row, 2 symmetric half (each block pair
                              stored once, reached transposed from the other row) */
    int32_t cg_variant;    /* 0 auto (1 on one GPU, 2 on several), 1 classic CG (two dot-product exchanges per
                              iteration), 2 single-reduction CG (Chronopoulos-Gear:
                              SpMV on the residual, one exchange per iteration)  */
    int32_t batch_mode;    /* a job that is a batch of >= 32 independent small meshes
                              (connected components with contiguous vertex ranges, each
                              <= 1024 free vertices) is solved one CTA per mesh, each
                              with its own iteration count: 0 auto, 1 never            */
    int32_t reorder;       /* internal breadth-first renumbering of the reduced system (results
                              and taps stay in vertex / canonical numbering): 0 auto (only when
                              the input vertex order has no locality), 1 never, 2 always,
                              3 Morton (Z-order) numbering from the vertex coordinates: 32-row
                              tiles become compact bricks of the mesh                          */
    int32_t precond;       /* preconditioner of the CG solve: 0 auto (multilevel on one rank for a single mesh of
                              >= 4096 free vertices, block-Jacobi otherwise), 1 6x6 block-Jacobi (the north-star
                              configuration the SpMV roofline is quoted on), 2 additive multilevel: block-Jacobi plus a
                              hierarchy of rigid-body aggregate coarse spaces                 */
    int32_t ml_smooth_from;/* multilevel: first coarse level whose prolongator to the next level is smoothed,
                              P = (I - w D^-1 A) T (0 = default 1; large = tentative prolongators everywhere)       */
    double  ml_cell;       /* multilevel: brick edge of the first coarse level in mean beam lengths (<= 0: 2)      */
    double  ml_ratio;      /* multilevel: growth of the brick edge per level (<= 1: 2)                              */
    int64_t ml_coarsest_rows; /* multilevel: stop coarsening at this many rows (<= 0: 512; dense inverse there)      */
} mbfea_options;

typedef struct mbfea_stats {
    int64_t n_vertices, n_beams, n_fixed;
    int64_t n_block_rows;       /* free vertices owned by this rank             */
    int64_t n_block_rows_global;
    int64_t nnz_blocks;         /* 6x6 blocks stored on this rank               */
    int64_t n_ghost_blocks;     /* halo x-blocks received per SpMV              */
    int64_t iterations;
    double  rel_residual;       /* recurrence residual ||r||/||f|| at exit      */
    double  true_rel_residual;  /* ||f-Ku||/||f|| recomputed after the solve    */
    /* host wall-clock, ms */
    double  ms_host_prep;       /* validate + float props + maps + pattern      */
    double  ms_h2d;
    /* device time (CUDA events on the ctx stream), ms */
    double  ms_pack, ms_assemble, ms_precond, ms_pcg, ms_recover, ms_d2h;
    double  ms_execute_total;
    /* algorithmic bytes (DESIGN.md "Algorithmic bytes") for this rank */
    double  bytes_spmv, bytes_pcg_iter, bytes_assemble, bytes_recover;
    int64_t kernel_launches;    /* kernels launched by the last execute()       */
    int32_t converged;          /* 1: recomputed residual <= tol; 2: recurrence met tol,
                                   recomputed residual stalls just above (FP64 limit of
                                   this system); 0: not converged                      */
    int32_t restarts;           /* residual replacements done by the solve             */
    int32_t n_components;       /* > 0: batched solve, one CTA per mesh; iterations = the
                                   slowest mesh's count                                 */
    int32_t precond;            /* preconditioner the last solve used: 1 block-Jacobi, 2 multilevel */
    int32_t ml_levels;          /* multilevel: coarse levels below the fine one                      */
    int32_t reserved;
    double  ms_ml_symbolic;     /* multilevel: host hierarchy build + uploads inside set_job (wall)  */
    double  ms_ml_setup;        /* multilevel: Galerkin products, smoothers, coarsest inverse (device time, part of ms_precond) */
    double  bytes_ml_apply;     /* multilevel: algorithmic bytes of the two fine-level transfer kernels per apply */
} mbfea_stats;

int          mbfea_version(void);
/* sizeof(mbfea_options), sizeof(mbfea_stats), sizeof(mbfea_plan_sizes) as compiled into the
 * library: lets a binding in another language verify its struct layouts at load time.       */
void         mbfea_abi_sizes(int64_t sizes3[3]);
void         mbfea_options_default(mbfea_options *opts);
int          mbfea_create(mbfea_ctx **out, const mbfea_options *opts);
void         mbfea_destroy(mbfea_ctx *ctx);
const char  *mbfea_last_error(const mbfea_ctx *ctx); /* ctx may be NULL: last create() error */
const char  *mbfea_status_string(int status);

/* ---- the simulation API: replaces BeamSimulator::setSimulation /
 *      FeaSimulationJob::{setElements,setNodes,setFixedNodes,setNodalForces}
 *      [REF src/beamsimulator.cpp:12-14,71-171] -------------------------------
 * nodes_cm   3V doubles  x[V],y[V],z[V]          (Eigen::MatrixX3d nodes)
 * elems_cm   2E int32    first[E],second[E]      (Eigen::MatrixX2i elements)
 * normals_cm 3E doubles                          (MatrixX3d elementalNormals)
 * dims_bh    2E floats   {b,h} per beam          (std::vector<BeamDimensions>)
 * mat_nuE    2E floats   {nu,E_GPa} per beam     (std::vector<BeamMaterial>)
 * fixed      F  int32                            (Eigen::VectorXi fixedNodes)
 * f_node/f_dof/f_mag  nF entries                 (std::vector<NodalForce>)
 * All inputs are copied; the caller may free them on return.  (Page-locked
 * input arrays make the uploads asynchronous to the host work; pageable ones work.)
 * Section properties EA,EIz,EIy,GJ are derived here, on the host, in IEEE
 * float with the reference's promotion rules [REF src/beamsimulator.cpp:173-188].
 * Everything else that is derived from the connectivity -- the block pattern of K,
 * the assembly's contributor lists in the element order of the reference's triplets
 * [REF :71-94], the solver format, the level hierarchy of the multilevel
 * preconditioner -- is built on the device from the uploaded mesh (host builders
 * of the same index structures serve internally renumbered jobs and are the
 * reference: mbfea_debug_compare_plan / mbfea_debug_compare_hierarchy compare the
 * two, entry for entry).  No arithmetic of the path runs on the CPU.
 */
int mbfea_set_job(mbfea_ctx *ctx, int64_t V, int64_t E,
                  const double *nodes_cm, const int32_t *elems_cm,
                  const double *normals_cm, const float *dims_bh,
                  const float *mat_nuE, int64_t F, const int32_t *fixed,
                  int64_t nF, const int64_t *f_node, const int64_t *f_dof,
                  const double *f_mag);

/* Replaces only the loads (same mesh, same constraints): the pattern, maps and
 * device mesh stay resident.  Same checks as FeaSimulationJob::setNodalForces. */
int mbfea_set_forces(mbfea_ctx *ctx, int64_t nF, const int64_t *f_node,
                     const int64_t *f_dof, const double *f_mag);

/* ---- replaces BeamSimulator::executeSimulation -> fea::solve
 *      [REF src/beamsimulator.cpp:40-69]: element stiffness + rotation,
 *      assembly, elimination, PCG, element-force recovery.  Results stay on
 *      the device until fetched.  Returns MBFEA_ENOCONV / MBFEA_ESINGULAR with
 *      results still readable.                                              */
int mbfea_execute(mbfea_ctx *ctx);

/* the three phases of execute(), separately callable (bench / profiling) */
int mbfea_assemble(mbfea_ctx *ctx); /* K1+K2 (+block-Jacobi setup K5)        */
int mbfea_solve(mbfea_ctx *ctx);    /* K4+K6 PCG                              */
int mbfea_recover(mbfea_ctx *ctx);  /* expand u, K7 element forces            */

/* ---- results: the memory image of SimulationResults
 *      [REF src/beamsimulator.hpp:45-52; src/beamsimulator.cpp:190-210] -----
 * U_cm   6V doubles, column-major V x 6 (Eigen::MatrixXd nodalDisplacements);
 *        rows of fixed vertices are 0.
 * F6x2E  12E doubles: component c in [0,6) occupies [c*2E,(c+1)*2E);
 *        entry 2e = end 0 of beam e, 2e+1 = end 1 (std::vector<VectorXd>
 *        edgeForces), local frame, order N,Ty,Tz,Mx,My,Mz.                   */
int mbfea_get_displacements(mbfea_ctx *ctx, double *U_cm);
int mbfea_get_edge_forces(mbfea_ctx *ctx, double *F6x2E);
int mbfea_get_stats(const mbfea_ctx *ctx, mbfea_stats *out);

/* ---- colour-mapping pre-pass: what the reference recomputes on the CPU right
 *      after the solve for its viewer and colour bar -- per force component the
 *      min / max over the 2E beam-end values (Colormap's labels
 *      [REF src/colorbar.cpp:79-111]) and the values normalised to [0,1]
 *      (igl::colormap(..., normalize = true) [REF src/gui.cpp:514-526]).      */
int mbfea_get_force_ranges(mbfea_ctx *ctx, double mins6[6], double maxs6[6]);
int mbfea_get_edge_forces_normalized(mbfea_ctx *ctx, double *T6x2E);

/* result files the reference has fea::solve write when a file name is set
 * [REF src/beamsimulator.cpp:46-54]: V x 6 and E x 12, ',' separated,
 * precision 14.  Opt-in here.                                               */
int mbfea_write_displacements_csv(mbfea_ctx *ctx, const char *path);
int mbfea_write_element_forces_csv(mbfea_ctx *ctx, const char *path);

/* ---- parity / debug taps (tests compare these with the oracle) ----------- */
int mbfea_get_props(mbfea_ctx *ctx, double *props4 /* E x {EA,EIz,EIy,GJ} */);
int mbfea_get_dofmap(mbfea_ctx *ctx, int32_t *free_index /* V; -1 = fixed */);
int mbfea_get_kelem(mbfea_ctx *ctx, double *Ke /* E x 12 x 12 row-major */);
int mbfea_get_bsr_sizes(mbfea_ctx *ctx, int64_t *n_block_rows, int64_t *nnz_blocks,
                        int64_t *row_offset /* first global block row owned */,
                        int64_t *n_ghost);
/* The block rows this rank owns, in CANONICAL reduced numbering (rank of the vertex among the
 * free vertices, what the reference / oracle use) whatever the internal renumbering: rows in
 * ascending canonical id (row_ids[nb]), rowptr nb+1, colidx nnzb GLOBAL canonical column ids,
 * ascending within a row, vals nnzb x 36 row-major 6x6.  Any pointer may be NULL.
 * On a single rank row_ids is 0..nb-1 and this is the reduced matrix of the whole job.   */
int mbfea_get_bsr(mbfea_ctx *ctx, int64_t *rowptr, int32_t *colidx, double *vals,
                  int64_t *row_ids);
int mbfea_get_load(mbfea_ctx *ctx, double *f_reduced /* 6*nb, owned rows, same row order */);
/* y = K x, single-rank contexts only: x,y 6*nb doubles (host), canonical numbering */
int mbfea_spmv(mbfea_ctx *ctx, const double *x, double *y);
/* same, with an explicit kernel choice (1 row-thread, 2 TMA-staged) */
int mbfea_spmv_with(mbfea_ctx *ctx, int32_t kernel, const double *x, double *y);
/* inject displacements (V x 6 column-major) and run only K7 on them */
int mbfea_recover_from(mbfea_ctx *ctx, const double *U_cm);

/* ---- measurement helpers (bench.py) --------------------------------------
 * Runs the SpMV kernel `reps` times on the ctx stream between two CUDA events
 * (after `warmup` untimed launches) and returns the mean ms per launch.      */
int mbfea_time_spmv(mbfea_ctx *ctx, int32_t kernel /* as options.spmv_kernel */,
                    int32_t warmup, int32_t reps, double *ms_per_launch);
int mbfea_time_assemble(mbfea_ctx *ctx, int32_t warmup, int32_t reps, double *ms_per_launch);

/* Tracing (opt-in: environment MBFEA_TRACE=1 at mbfea_solve): device globaltimer
 * stamps (ns) of the first 256 PCG iterations, *slots_per_iter (= 8) per iteration:
 * 0 SpMV entry, 1 SpMV past the halo wait, 2 update entry, 3 update past the dot wait,
 * 4 p-update entry, 5 p-update past the dot wait, 6 halo push entry, 7 halo push done.
 * Slots 1,3,5,6,7 are only written on the multi-GPU peer-memory path.            */
int mbfea_debug_trace(mbfea_ctx *ctx, uint64_t *stamps, int64_t max_iters, int64_t *slots_per_iter);

/* Parity tap of the set-up path.  On one rank the block pattern, the contributor lists of the assembly and the
 * solver format are built on the device; several ranks (and internally renumbered or irregular jobs) use the host
 * builders.  This call rebuilds all of them with the HOST builders from the mesh the device holds and compares them
 * with the device arrays, entry for entry.  mismatches[k] = number of differing entries (-1: different length) of
 * 0 rowptr, 1 colidx, 2 diag_idx, 3 contrib_ptr, 4 contrib, 5 uptr, 6 ucol, 7 usrc, 8 lptr, 9 lcol, 10 lofs,
 * 11 tile_base.  *built_on_device: 1 pattern on the device, 2 pattern and solver format on the device, 0 neither. */
int mbfea_debug_compare_plan(mbfea_ctx *ctx, int64_t mismatches[12], int32_t *built_on_device);
/* The same for the symbolic half of the multilevel preconditioner (one rank): rebuilds the level hierarchy with the HOST
 * builder (same parameters, same mean beam length) and compares it with the device arrays.  16 counters per level,
 * *levels levels (capacity: 16 * 16 entries): level 0: agg, off, agg_ptr, agg_rows, level-1 rowptr, colidx, diagonal
 * positions, gal_ptr, gal_src; level l >= 1: agg, off, prow, pcol, rrow, rblk, rfine, aprow, apcol, next level's rowptr,
 * colidx, diagonal positions, smoothed flag.  Counter = differing entries, -1 different length, -2 level count differs.
 * *built_on_device: 1 when the hierarchy of the current job was built on the device.  MBFEA_ESTATE without a hierarchy. */
int mbfea_debug_compare_hierarchy(mbfea_ctx *ctx, int64_t mismatches[256], int32_t *levels, int32_t *built_on_device);

/* ---- multi-GPU (one process per GPU; vertex-block partition) -------------
 * Rank 0 makes an id, the launcher broadcasts the 128 bytes (torch.distributed
 * in bench.py), every rank calls comm_init before set_job.  Collectives on the
 * data path: halo exchange of boundary x-blocks per SpMV and an all-gather of
 * per-rank dot-product partials (summed in fixed rank order).               */
int mbfea_comm_unique_id(uint8_t id128[128]);
int mbfea_comm_init(mbfea_ctx *ctx, const uint8_t id128[128], int32_t rank, int32_t world);

/* Host-only partition plan (no GPU needed; exercised by the gloo CPU tests):
 * which block rows `rank` owns and which ghost blocks it receives from whom.
 * Call with NULL arrays to size.  send_* describe what this rank SENDS.      */
typedef struct mbfea_plan_sizes {
    int64_t nb_global, row_begin, row_end, nnz_blocks, n_ghost, n_send, n_neighbors;
    /* batch detection (world == 1): independent block-row ranges of the reduced system when the
     * job is a batch of >= 32 small meshes (each <= 1024 rows), else 0                        */
    int64_t n_components, max_component_rows;
} mbfea_plan_sizes;
int mbfea_plan_partition(int64_t V, int64_t E, const int32_t *elems_cm,
                         int64_t F, const int32_t *fixed, int32_t rank, int32_t world,
                         mbfea_plan_sizes *sizes,
                         int64_t *rowptr /* nb_local+1 */, int32_t *colidx /* nnzb, local ids */,
                         int64_t *ghost_global /* n_ghost, ascending => grouped by owner */,
                         int32_t *ghost_owner /* n_ghost */,
                         int64_t *send_local_rows /* n_send local row ids, grouped by dest */,
                         int32_t *send_dest /* n_send */);

/* ---- scenario files ("next" row 1): config.json + ASCII PLY edge mesh -----
 * [REF src/utilities.hpp:50-95; src/edgemesh.cpp:115-135;
 *      Models/TwoLineSegments.ply].  Loads the scenario and calls set_job.    */
int mbfea_load_scenario(mbfea_ctx *ctx, const char *config_json_path,
                        const char *ply_path_override /* NULL = use plyFilename */);

/* The same files parsed into flat arrays without touching a GPU (host-only). */
typedef struct mbfea_scenario mbfea_scenario;
int  mbfea_scenario_open(const char *config_json_path, const char *ply_path_override,
                         mbfea_scenario **out, char *errbuf, int64_t errbuf_len);
void mbfea_scenario_sizes(const mbfea_scenario *sc, int64_t *V, int64_t *E, int64_t *F, int64_t *nF);
void mbfea_scenario_copy(const mbfea_scenario *sc, double *nodes_cm, int32_t *elems_cm,
                         double *normals_cm, float *dims_bh, float *mat_nuE, int32_t *fixed,
                         int64_t *f_node, int64_t *f_dof, double *f_mag);
void mbfea_scenario_close(mbfea_scenario *sc);

/* ---- host-only pieces of set_job, callable without a GPU (CPU tests) --------
 * validate: the checks of set_job in the reference's order, same return codes;
 *           msg (may be NULL) receives the text mbfea_last_error would give.
 * props   : EA,EIz,EIy,GJ per beam with the reference's float arithmetic
 *           [REF src/beamsimulator.cpp:173-188].
 * free_map: K3 on the host (the device kernel must agree bit for bit).        */
int mbfea_host_validate(int64_t V, int64_t E, const int32_t *elems_cm, const float *dims_bh,
                        const float *mat_nuE, int64_t F, const int32_t *fixed, int64_t nF,
                        const int64_t *f_node, const int64_t *f_dof, char *msg, int64_t msg_len);
int mbfea_host_props(int64_t E, const float *dims_bh, const float *mat_nuE, double *props4);
int mbfea_host_free_map(int64_t V, int64_t F, const int32_t *fixed, int32_t *free_index,
                        int64_t *n_free);
/* renumber: the internal breadth-first renumbering set_job would use (mode as options.reorder:
 *           0 auto, 1 never, 2 always; 3 needs coordinates and acts as 2 here).  perm[n_free]: canonical reduced id -> internal block row
 *           (identity when the canonical order is kept); *reordered says which.                */
int mbfea_host_renumber(int64_t V, int64_t E, const int32_t *elems_cm, int64_t F, const int32_t *fixed,
                        int32_t mode, int32_t *perm, int32_t *reordered);
/* contributors: the assembly plan of rank `rank` of `world` -- for block k of the rank's pattern
 *           (the order of mbfea_plan_partition's colidx) the entries [contrib_ptr[k], contrib_ptr[k+1])
 *           of contrib are (beam << 2) | sub-block (0: (a,a), 1: (a,b), 2: (b,a), 3: (b,b)) in ascending
 *           beam order = the order Eigen's setFromTriplets sums duplicates in.  sizes[2] = {blocks,
 *           entries}; arrays may be NULL (sizes only): contrib_ptr[blocks+1], contrib[entries].   */
int mbfea_host_contributors(int64_t V, int64_t E, const int32_t *elems_cm, int64_t F, const int32_t *fixed,
                            int32_t rank, int32_t world, int64_t sizes[2], int32_t *contrib_ptr,
                            int32_t *contrib);
/* solver_format: the symmetric-half solver format set_job would build on rank `rank` of `world`, in the
 *           INTERNAL numbering (local block rows; ghost columns have ids >= nb and are always stored) (mode as options.reorder; nodes_cm may be NULL unless mode == 3).
 *           sizes[8] = {block rows nb, off-diagonal blocks of the full pattern, stored blocks,
 *           transposed references, 16-byte units of the interleaved value array, balanced 0/1,
 *           longest stored row, renumbered 0/1}.  Arrays may be NULL (sizes only); otherwise
 *           uptr[nb+1], ucol[stored], lptr[nb+1], lcol[refs], lofs[refs] (16-byte unit of the
 *           referenced block's first chunk), tile_base[ceil(nb/32)+1], perm[nb].               */
int mbfea_host_solver_format(int64_t V, int64_t E, const double *nodes_cm, const int32_t *elems_cm, int64_t F,
                             const int32_t *fixed, int32_t mode, int32_t rank, int32_t world, int64_t sizes[8],
                             int32_t *uptr, int32_t *ucol, int32_t *lptr, int32_t *lcol, int32_t *lofs,
                             int32_t *tile_base, int32_t *perm);

/* hierarchy: aggregates and tentative Galerkin structures of the multilevel (rigid-body aggregate) preconditioner
 *           (options.precond = 2), exposed for the CPU tests.  Builds, for a
 *           one-rank job in the canonical reduced numbering, the chain of coarse levels: aggregates =
 *           bricks of edge cell0 * ratio^l over the free vertices' bounding box, 6 rigid-body degrees of
 *           freedom each.  cell0 <= 0 selects 4 x the mean beam length.
 *  levels  : number of coarse levels built (0: nothing to coarsen).
 *  sizes[6] of coarse level l (0-based): {fine rows, coarse rows, coarse blocks, fine blocks, 0, 0}.
 *  copy    : agg[fine rows], off[3 * fine rows] (fine position - aggregate centroid, x y z per row),
 *            rowptr[coarse rows + 1], colidx[coarse blocks], gal_ptr[coarse blocks + 1],
 *            gal_src[fine blocks] (indices into the finer level's block pattern, ascending per coarse
 *            block; level 0's finer pattern is mbfea_plan_partition's with world = 1).  Any may be NULL. */
/* ml_hierarchy: the level hierarchy mbfea_set_job builds for the multilevel preconditioner (options.precond = 2) of a
 *           one-rank job, sizes only (host-only; CPU tests and set_job timing).  *n_levels: coarse levels; sizes[6 * l ..]
 *           for the transfer from level l to l + 1: {rows of level l, rows of level l + 1, 6x6 blocks of level l + 1's
 *           matrix, blocks of the prolongator, blocks of A P (0 on the fine level: contributor lists), smoothed 0/1}.
 *           Arguments <= 0 select the defaults of mbfea_options.                                                        */
int mbfea_host_ml_hierarchy(int64_t V, int64_t E, const double *nodes_cm, const int32_t *elems_cm, int64_t F,
                            const int32_t *fixed, double cell_beam_lengths, double ratio, int64_t coarsest_rows,
                            int32_t smooth_from, int32_t max_levels, int32_t *n_levels, int64_t *sizes);
/* ml_image: the flat image of the hierarchy that rank 0 of a `world`-rank job builds and broadcasts (mbfea_set_job with
 *           options.precond = 2 on several ranks; csrc/mbfea_internal.hpp: MlImageHeader followed by the arrays).  Host-only,
 *           for the CPU tests of the multi-rank slicing.  Call with image = NULL to get *n_bytes, then with a buffer.     */
int mbfea_host_ml_image(int64_t V, int64_t E, const double *nodes_cm, const int32_t *elems_cm, int64_t F,
                        const int32_t *fixed, int32_t world, int32_t smooth_from, int64_t *n_bytes, unsigned char *image);
/* spd_inverse: explicit inverse of a dense symmetric positive definite matrix (row-major n x n), the
 *           coarsest-level solve of that preconditioner; MBFEA_ESINGULAR on a non-positive pivot.    */
int mbfea_host_spd_inverse(int64_t n, const double *A, double *Ainv);
typedef struct mbfea_hierarchy mbfea_hierarchy;
int mbfea_hierarchy_open(int64_t V, int64_t E, const double *nodes_cm, const int32_t *elems_cm, int64_t F,
                         const int32_t *fixed, double cell0, double ratio, int32_t max_levels,
                         int64_t coarsest_rows, mbfea_hierarchy **out);
int mbfea_hierarchy_levels(const mbfea_hierarchy *h);
int mbfea_hierarchy_sizes(const mbfea_hierarchy *h, int32_t level, int64_t sizes[6]);
int mbfea_hierarchy_copy(const mbfea_hierarchy *h, int32_t level, int32_t *agg, double *off, int32_t *rowptr,
                         int32_t *colidx, int32_t *gal_ptr, int32_t *gal_src);
void mbfea_hierarchy_close(mbfea_hierarchy *h);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* MBFEA_H */
