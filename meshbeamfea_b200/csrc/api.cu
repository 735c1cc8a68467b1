// libmbfea.so C ABI (include/mbfea.h): context, device residency, the CG driver
// (CUDA-graph batches, residual verification/replacement, batched small meshes),
// multi-GPU plumbing (NCCL bootstrap, CUDA-IPC peer memory) and the parity taps.
// Host code is C++17; all device work is in kernels.cu.  There is no CPU
// fallback: without a CUDA device every compute entry point fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "devbuf.hpp"
#include "kernels.cuh"
#include "mbfea_internal.hpp"
#include "mlprec.hpp"
#include "nccl_dl.hpp"

using namespace mbfea;
using namespace mbfea::dev;

namespace {

thread_local std::string g_create_error;

double now_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

}  // namespace

struct mbfea_ctx {
    mbfea_options opt{};
    mutable std::string err;
    int device = 0, sm_count = 148;
    cudaStream_t st = nullptr;
    cudaEvent_t ev[8] = {};

    // job (host)
    bool have_job = false, assembled = false, solved = false, recovered = false;
    int64_t V = 0, E = 0, F = 0;
    HostPlan plan;
    hvec<float> props_host;              // EA, EIz, EIy, GJ per beam: float values in the reference, uploaded as such
    hvec<int32_t> tile_row_host;    // TMA tiles over the full pattern (K, parity tap)
    hvec<int32_t> tile_row_s_host;  // TMA tiles over the solver format (full-row storage only)
    bool tma_ok = false;       // pattern fits the TMA tile geometry
    bool tma_prepared = false;

    // device: mesh
    DevBuf<Node4> node4;
    DevBuf<ElemRec> rec;
    DevBuf<ElemGeom> geom;
    DevBuf<int32_t> free_index, free_canon, fixed;   // free_canon: the K3 scan (canonical); free_index: composed with the internal renumbering
    // staging for set_job (kept: a re-submitted mesh of the same size costs no cudaMalloc/cudaFree)
    DevBuf<double> stg_nodes, stg_normals;
    DevBuf<float> stg_props;
    DevBuf<int32_t> stg_elems, stg_scratch, stg_perm;
    DevBuf<int64_t> stg_fnode, stg_fdof;
    DevBuf<double> stg_fmag;
    // device: pattern
    DevBuf<int32_t> rowptr, colidx, diag_idx, contrib_ptr, contrib, tile_row, tile_row_s, send_rows;
    DevBuf<int32_t> uptr, ucol, usrc, lptr, lcol, lofs, ilv_base;   // solver format (HostPlan::uptr ...)
    int64_t n_stored = 0, n_refs = 0, tile_units = 0;   // stored blocks, transposed references, 16-byte units of the interleaved value array
    // pattern / contributor lists / solver format built on the device (plan_dev.cu): scratch, kept between jobs
    bool device_plan = false;
    DevBuf<int32_t> pd_cnt, pd_cur, pd_start, pd_scan;
    DevBuf<int32_t> g_rowptr, g_colidx, g_diag, g_cptr, g_contrib;   // several ranks: pattern of the whole system (the rank's share is cut from it; rank 0: hierarchy build)
    DevBuf<long long> ghost_dev;
    DevBuf<unsigned long long> pd_ent, pd_tot;
    // device: matrix + vectors.  vals = assembled K (full pattern); vals_s = stored blocks of
    // K~ = S K S^T; lfac/sinv = Cholesky factor of the diagonal blocks and its inverse
    DevBuf<double> vals, vals_s, lfac, sinv, f, ft, x, y, rt, p, Ap, sdir, sendbuf;
    DevBuf<double> part_a, part_b, part_c, red, gath;
    DevBuf<PcgScalars> sc;
    DevBuf<int> flag;
    PcgScalars *sc_host = nullptr;  // pinned
    // device: results
    DevBuf<double> xg, U_cm, U_rm, F6;

    // multi-GPU
    const nccl::Api *nc = nullptr;
    nccl::Comm comm = nullptr;
    // NVLink peer-memory path (comm_mode 0/2): mailbox + ghost regions mapped through CUDA IPC
    bool p2p = false;
    bool comm_broken = false;   // a peer stopped responding: no further collective is attempted (destroy skips its rendezvous)
    DevBuf<Mailbox> mailbox;
    DevBuf<PeerCtx> peerctx;
    DevBuf<unsigned int> tickets;
    DevBuf<double *> push_dst;   // destination of every (boundary row, neighbour) pair, grouped by row
    DevBuf<double *> push_dst_send;  // the same in send-list order (standalone push kernel, iteration 0)
    DevBuf<int32_t> push_idx;    // per owned row: first entry in push_dst | count << 28, or -1
    DevBuf<int32_t> push_rows;   // the owned rows with push_idx >= 0, ascending (boundary-first pass of the step kernel)
    std::vector<Mailbox *> peer_mailbox;   // [world]; own entry = mailbox.p
    std::vector<double *> peer_p;          // [world]; peers' p vectors (IPC mappings), own = p.p
    unsigned long long epoch = 0;
    // batched solve (one CTA per connected component)
    DevBuf<int32_t> comp_ptr, comp_iters, comp_status;
    DevBuf<double> comp_rr, comp_ff;
    DevBuf<unsigned long long> trace;  // MBFEA_TRACE=1: globaltimer stamps of the first iterations

    // multilevel preconditioner (options.precond): hierarchy built in set_job, numeric setup in assemble
    MlPrec ml;
    MlPrec ml_global;   // several ranks, rank 0: builder of the hierarchy of the whole system (device path)
    bool use_ml = false;
    bool ml_on_device = false;   // the hierarchy of the current job was built on the device
    double ml_mean_len = 0.0;
    MlHierarchy ml_host;   // host half of the hierarchy, kept between jobs: its arrays keep their capacity
    HostPlan ml_gplan;                // several ranks, rank 0: pattern of the whole reduced system (the hierarchy is global)
    MlImageLayout ml_layout;          // several ranks, rank 0: where every hierarchy array sits in the broadcast image
    DevBuf<unsigned char> ml_dimg;    // several ranks: the device image
    std::vector<double> ml_pos;

    // PCG graph cache
    cudaGraphExec_t graph_exec = nullptr;
    int graph_iters = 0, graph_kernel = 0;
    int64_t launches = 0, graph_launches = 0;

    mbfea_stats stats{};

    int64_t nb() const { return plan.nb(); }
    int64_t ng() const { return (int64_t)plan.ghost_global.size(); }
    int world() const { return opt.world; }
    int64_t row_begin_of(int q) const { return plan.nb_global * q / opt.world; }
};

namespace {

int fail(mbfea_ctx *c, int code, const std::string &msg)
{
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}

template <class Fn>
int guarded(mbfea_ctx *c, Fn fn)
{
    try {
        return fn();
    } catch (const CudaFail &f) {
        char buf[512];
        std::snprintf(buf, sizeof buf, "CUDA error %d (%s) at api.cu:%d in %s", (int)f.e, cudaGetErrorString(f.e),
                      f.line, f.what);
        cudaGetLastError();
        return fail(c, MBFEA_ECUDA, buf);
    } catch (const NcclFail &f) {
        char buf[512];
        std::snprintf(buf, sizeof buf, "NCCL error %d (%s) at api.cu:%d in %s", f.e,
                      (c && c->nc) ? c->nc->GetErrorString(f.e) : "?", f.line, f.what);
        return fail(c, MBFEA_ECUDA, buf);
    } catch (const std::bad_alloc &) {
        return fail(c, MBFEA_EARG, "host allocation failed");
    }
}

template <class T>
void upload(mbfea_ctx *c, DevBuf<T> &d, const T *h, size_t n)
{
    CK(d.alloc(n));
    if (n) CK(cudaMemcpyAsync(d.p, h, n * sizeof(T), cudaMemcpyHostToDevice, c->st));
}

template <class T, class A>
void upload(mbfea_ctx *c, DevBuf<T> &d, const std::vector<T, A> &h) { upload(c, d, h.data(), h.size()); }

void drop_graph(mbfea_ctx *c)
{
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    c->graph_exec = nullptr; c->graph_iters = 0; c->graph_kernel = 0;
}

int pick_spmv(const mbfea_ctx *c, int requested)
{
    int k = requested ? requested : c->opt.spmv_kernel;
    if (k == 0) k = 1;  // measured on B200: the row-thread kernel reaches ~95 % of HBM peak, the TMA one ~70 %
    if (k == 2 && !c->tma_ok) k = 1;
    return k;
}

void prepare_tma(mbfea_ctx *c)
{
    if (!c->tma_prepared) {
        CK((cudaError_t)spmv_tma_prepare());
        c->tma_prepared = true;
    }
}

// parity tap: y = K x with the assembled (unscaled) matrix in its full pattern
void run_spmv_full(mbfea_ctx *c, int kernel, const double *x, double *y)
{
    const int64_t nb = c->nb();
    if (nb <= 0) return;
    if (kernel == 2) {
        prepare_tma(c);
        const int64_t ntiles = (int64_t)c->tile_row_host.size() - 1;
        launch_spmv_tma(c->st, spmv_tma_grid(ntiles, c->sm_count), ntiles, c->tile_row.p, c->rowptr.p, c->colidx.p,
                        c->vals.p, 0, x, y, nullptr, nullptr);
    } else {
        launch_spmv_rowthread(c->st, nb, c->rowptr.p, c->colidx.p, c->vals.p, nullptr, nullptr, nullptr, nullptr, 0, x,
                              y, nullptr, nullptr, nullptr, nullptr);
    }
    c->launches++;
}

// the solver's SpMV: y = K~ x (unit diagonal implied; x has nb + n_ghost blocks)
void run_spmv(mbfea_ctx *c, int kernel, const double *x, double *y, double *partial, const PcgScalars *sc,
              const PeerCtx *pc, double *partial2 = nullptr)
{
    const int64_t nb = c->nb();
    if (nb <= 0) return;
    if (kernel == 2 && !c->plan.sym_half && pc == nullptr && partial2 == nullptr) {
        prepare_tma(c);
        const int64_t ntiles = (int64_t)c->tile_row_s_host.size() - 1;
        launch_spmv_tma(c->st, spmv_tma_grid(ntiles, c->sm_count), ntiles, c->tile_row_s.p, c->uptr.p, c->ucol.p,
                        c->vals_s.p, 1, x, y, partial, sc);
    } else {
        const bool sym = c->plan.sym_half;
        launch_spmv_rowthread(c->st, nb, c->uptr.p, c->ucol.p, c->vals_s.p, sym ? c->lptr.p : nullptr,
                              sym ? c->lcol.p : nullptr, sym ? c->lofs.p : nullptr, sym ? c->ilv_base.p : nullptr, 1, x,
                              y, partial, partial2, sc, pc);
    }
    c->launches++;
}

// auto: classic CG on one GPU (same speed, the textbook recurrence); single-reduction CG on several
// (one cross-GPU exchange and two kernels per iteration instead of two and three: -10 % at 4 GPUs)
int cg_variant(const mbfea_ctx *c)
{
    if (c->use_ml) return 1;   // the preconditioned recurrence is the classic one (z~ = M r~ between update and p-update)
    if (c->opt.cg_variant == 1 || c->opt.cg_variant == 2) return c->opt.cg_variant;
    return c->opt.world > 1 ? 2 : 1;
}

int spmv_grid(const mbfea_ctx *c, int kernel)
{
    if (kernel == 2 && !c->plan.sym_half && !c->p2p && cg_variant(c) == 1)
        return spmv_tma_grid((int64_t)c->tile_row_s_host.size() - 1, c->sm_count);
    return spmv_rowthread_grid(c->nb(), c->plan.sym_half, c->p2p);
}

// boundary x-blocks -> neighbours' ghost slots (multi-GPU only)
// `width` doubles per block row: 6 for vectors, 36 for the S factors
void halo_exchange(mbfea_ctx *c, double *xvec, int width = 6)
{
    if (c->world() <= 1 || c->plan.neighbors.empty()) return;
    const int64_t nb = c->nb();
    launch_pack_halo(c->st, (int64_t)c->plan.send_rows.size(), c->send_rows.p, xvec, c->sendbuf.p, width);
    c->launches++;
    NK(c->nc->GroupStart());
    for (const Neighbor &q : c->plan.neighbors) {
        if (q.send_count)
            NK(c->nc->Send(c->sendbuf.p + width * q.send_offset, (size_t)(width * q.send_count), nccl::kFloat64,
                           q.rank, c->comm, c->st));
        if (q.recv_count)
            NK(c->nc->Recv(xvec + width * (nb + q.recv_offset), (size_t)(width * q.recv_count), nccl::kFloat64,
                           q.rank, c->comm, c->st));
    }
    NK(c->nc->GroupEnd());
}

// all-gather of `bytes` (multiple of 8) host bytes per rank through NCCL (setup paths only)
void allgather_host(mbfea_ctx *c, const void *in, void *out, size_t bytes)
{
    const int W = c->world();
    DevBuf<unsigned char> din, dout;
    CK(din.alloc(bytes));
    CK(dout.alloc(bytes * W));
    CK(cudaMemcpyAsync(din.p, in, bytes, cudaMemcpyHostToDevice, c->st));
    NK(c->nc->AllGather(din.p, dout.p, bytes / 8, nccl::kInt64, c->comm, c->st));
    CK(cudaMemcpyAsync(out, dout.p, bytes * W, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
}

void close_peer_vectors(mbfea_ctx *c)
{
    for (size_t q = 0; q < c->peer_p.size(); ++q)
        if ((int)q != c->opt.rank && c->peer_p[q]) cudaIpcCloseMemHandle(c->peer_p[q]);
    c->peer_p.clear();
}

// comm_init: decide whether every peer is reachable over NVLink peer memory and map the mailboxes
void setup_p2p_mailboxes(mbfea_ctx *c)
{
    const int W = c->world(), me = c->opt.rank;
    c->p2p = false;
    if (W <= 1 || c->opt.comm_mode == 1 || W > kMaxWorld) return;
    // (device ordinal, pid-independent) of every rank; all must be peer-accessible from here
    std::vector<int64_t> devs((size_t)W);
    int64_t mine = c->device;
    allgather_host(c, &mine, devs.data(), sizeof(int64_t));
    int64_t ok = 1;
    for (int q = 0; q < W; ++q) {
        if (q == me) continue;
        int can = 0;
        if (devs[q] == c->device || cudaDeviceCanAccessPeer(&can, c->device, (int)devs[q]) != cudaSuccess || !can) ok = 0;
    }
    cudaGetLastError();
    CK(c->mailbox.alloc(1));
    CK(c->peerctx.alloc(1));
    CK(cudaMemsetAsync(c->mailbox.p, 0, sizeof(Mailbox), c->st));
    CK(cudaMemsetAsync(c->tickets.p, 0, 8 * sizeof(unsigned int), c->st));
    CK(cudaStreamSynchronize(c->st));
    struct Rec { cudaIpcMemHandle_t h; int64_t ok; };
    static_assert(sizeof(Rec) % 8 == 0, "IPC record must be a multiple of 8 bytes");
    Rec r;
    std::memset(&r, 0, sizeof r);
    if (ok && cudaIpcGetMemHandle(&r.h, c->mailbox.p) != cudaSuccess) { ok = 0; cudaGetLastError(); }
    r.ok = ok;
    std::vector<Rec> all((size_t)W);
    allgather_host(c, &r, all.data(), sizeof(Rec));
    for (int q = 0; q < W; ++q) ok = ok && all[q].ok;
    c->peer_mailbox.assign((size_t)W, nullptr);
    if (ok) {
        for (int q = 0; q < W && ok; ++q) {
            if (q == me) { c->peer_mailbox[q] = c->mailbox.p; continue; }
            void *ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, all[q].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); }
            c->peer_mailbox[q] = static_cast<Mailbox *>(ptr);
        }
    }
    // every rank must take the same path
    int64_t okl = ok;
    std::vector<int64_t> oks((size_t)W);
    allgather_host(c, &okl, oks.data(), sizeof(int64_t));
    for (int q = 0; q < W; ++q) ok = ok && oks[q];
    c->p2p = ok != 0;
    if (!c->p2p && c->opt.comm_mode == 2) throw CudaFail{cudaErrorPeerAccessUnsupported, "comm_mode 2: peer memory not available on every rank", __LINE__};
    if (c->opt.verbose)
        std::fprintf(stderr, "[mbfea] rank %d/%d: PCG data path over %s\n", me, W,
                     c->p2p ? "NVLink peer memory (CUDA IPC mailboxes, no NCCL calls per iteration)" : "NCCL send/recv + all-gather");
}

// set_job: map the peers' p vectors and build the push list (where each owned boundary block goes)
void setup_p2p_halo(mbfea_ctx *c)
{
    if (!c->p2p) return;
    const int W = c->world(), me = c->opt.rank;
    const int64_t nb = c->nb();
    struct Rec { cudaIpcMemHandle_t h; int64_t nb; int64_t recv_off[kMaxWorld]; };
    static_assert(sizeof(Rec) % 8 == 0, "IPC record must be a multiple of 8 bytes");
    Rec r;
    std::memset(&r, 0, sizeof r);
    CK(cudaIpcGetMemHandle(&r.h, c->p.p));
    r.nb = nb;
    for (int q = 0; q < W; ++q) r.recv_off[q] = -1;
    for (const Neighbor &n : c->plan.neighbors) r.recv_off[n.rank] = n.recv_offset;
    std::vector<Rec> all((size_t)W);
    allgather_host(c, &r, all.data(), sizeof(Rec));
    if (c->peer_p.empty()) {   // (else: the mappings of the previous job are still valid, see set_job)
        c->peer_p.assign((size_t)W, nullptr);
        for (int q = 0; q < W; ++q) {
            if (q == me) { c->peer_p[q] = c->p.p; continue; }
            void *ptr = nullptr;
            CK(cudaIpcOpenMemHandle(&ptr, all[q].h, cudaIpcMemLazyEnablePeerAccess));
            c->peer_p[q] = static_cast<double *>(ptr);
        }
    }
    // what rank q receives from me occupies q's ghost slots [recv_off_q[me], ...) in my send order
    std::vector<double *> dst(c->plan.send_rows.size());
    for (const Neighbor &n : c->plan.neighbors)
        for (int64_t j = 0; j < n.send_count; ++j)
            dst[(size_t)(n.send_offset + j)] = c->peer_p[n.rank] + 6 * (all[n.rank].nb + all[n.rank].recv_off[me] + j);
    upload(c, c->push_dst_send, dst);
    // the same destinations grouped by row, for the push fused into k_cg_pupdate
    {
        std::vector<int32_t> cnt((size_t)nb, 0), idx((size_t)nb, -1);
        for (int32_t r : c->plan.send_rows) cnt[(size_t)r]++;
        std::vector<int32_t> first((size_t)nb + 1, 0);
        for (int64_t r = 0; r < nb; ++r) first[(size_t)r + 1] = first[(size_t)r] + cnt[(size_t)r];
        std::vector<double *> byrow(dst.size());
        std::vector<int32_t> cur(first.begin(), first.end() - 1);
        for (size_t i = 0; i < dst.size(); ++i) byrow[(size_t)cur[(size_t)c->plan.send_rows[i]]++] = dst[i];
        for (int64_t r = 0; r < nb; ++r)
            if (cnt[(size_t)r]) idx[(size_t)r] = first[(size_t)r] | (cnt[(size_t)r] << 28);
        std::vector<int32_t> rows;
        for (int64_t r = 0; r < nb; ++r)
            if (cnt[(size_t)r]) rows.push_back((int32_t)r);
        upload(c, c->push_dst, byrow);
        upload(c, c->push_idx, idx);
        upload(c, c->push_rows, rows);
    }
    PeerCtx pc;
    std::memset(&pc, 0, sizeof pc);
    pc.rank = me; pc.world = W; pc.n_neighbors = (int)c->plan.neighbors.size();
    for (int j = 0; j < pc.n_neighbors; ++j) pc.neighbor[j] = c->plan.neighbors[(size_t)j].rank;
    pc.mine = c->mailbox.p;
    for (int q = 0; q < W; ++q) pc.peer[q] = c->peer_mailbox[(size_t)q];
    pc.tickets = c->tickets.p;
    CK(cudaMemcpyAsync(c->peerctx.p, &pc, sizeof pc, cudaMemcpyHostToDevice, c->st));
    CK(cudaStreamSynchronize(c->st));
}

// One CG iteration (parity = iteration & 1) enqueued on the stream.
// Dot products: single GPU -> consumers read the per-CTA partials directly.
// NVLink peer-memory path -> 4 kernels, no NCCL call: the SpMV waits for the neighbours' pushed
// ghost blocks and its last CTA all-gathers p.Ap through the mailboxes; update does the same for
// rho; push_halo stores the new boundary blocks of p into the neighbours' ghost regions.
// NCCL path -> each rank collapses its partials, all-gathers the per-rank values, and every rank
// sums the W entries in rank order (identical bits on every rank => same stop decision).
void enqueue_iteration(mbfea_ctx *c, int kernel, int parity)
{
    const int64_t nb = c->nb();
    const int W = c->world();
    const int gs = spmv_grid(c, kernel), gv = cg_vec_grid(nb);
    if (c->p2p) {
        run_spmv(c, 1, c->p.p, c->Ap.p, c->part_a.p, c->sc.p, c->peerctx.p);
        launch_cg_update(c->st, nb, parity, PartialRef{c->part_a.p, gs, 1}, c->p.p, c->Ap.p, c->y.p, c->rt.p,
                         c->part_b.p, c->sc.p, c->peerctx.p, c->use_ml ? 0 : 1);
        const double *zdir = c->rt.p;
        if (c->use_ml) {   // z~ = M r~; its last kernel all-gathers r~.z~ through the mailboxes in place of r~.r~
            c->ml.apply(c->st, c->rt.p, c->sdir.p, c->part_b.p, gv, c->sc.p, c->peerctx.p);
            c->launches += c->ml.kernels_per_apply();
            zdir = c->sdir.p;
        }
        launch_cg_pupdate(c->st, nb, parity, PartialRef{c->part_b.p, gv, 1}, zdir, c->p.p, c->sc.p, c->peerctx.p,
                          c->push_idx.p, c->push_dst.p, c->tickets.p + 2);
        c->launches += 2;
        return;
    }
    halo_exchange(c, c->p.p);
    run_spmv(c, kernel, c->p.p, c->Ap.p, c->part_a.p, c->sc.p, nullptr);
    PartialRef pAp{c->part_a.p, gs, 1};
    if (W > 1) {
        launch_reduce_partials(c->st, c->part_a.p, gs, c->red.p, nullptr, nullptr, c->sc.p);
        NK(c->nc->AllGather(c->red.p, c->gath.p, 1, nccl::kFloat64, c->comm, c->st));
        pAp = PartialRef{c->gath.p, W, 1};
        c->launches++;
    }
    launch_cg_update(c->st, nb, parity, pAp, c->p.p, c->Ap.p, c->y.p, c->rt.p, c->part_b.p, c->sc.p, nullptr);
    if (c->use_ml) {
        // z~ = M r~ (into sdir), rho = r~ . z~; the direction update takes z~ where plain CG takes r~
        c->ml.apply(c->st, c->rt.p, c->sdir.p, c->part_b.p, gv);
        PartialRef rz{c->part_b.p, gv, 1};
        if (W > 1) {
            launch_reduce_partials(c->st, c->part_b.p, gv, c->red.p + 2, nullptr, nullptr, c->sc.p);
            NK(c->nc->AllGather(c->red.p + 2, c->gath.p + W, 1, nccl::kFloat64, c->comm, c->st));
            rz = PartialRef{c->gath.p + W, W, 1};
            c->launches++;
        }
        launch_cg_pupdate(c->st, nb, parity, rz, c->sdir.p, c->p.p, c->sc.p, nullptr, nullptr, nullptr, c->tickets.p + 2);
        c->launches += 2 + c->ml.kernels_per_apply();
        return;
    }
    PartialRef rho{c->part_b.p, gv, 1};
    if (W > 1) {
        launch_reduce_partials(c->st, c->part_b.p, gv, c->red.p + 2, nullptr, nullptr, c->sc.p);
        NK(c->nc->AllGather(c->red.p + 2, c->gath.p + W, 1, nccl::kFloat64, c->comm, c->st));
        rho = PartialRef{c->gath.p + W, W, 1};
        c->launches++;
    }
    launch_cg_pupdate(c->st, nb, parity, rho, c->rt.p, c->p.p, c->sc.p, nullptr, nullptr, nullptr, c->tickets.p + 2);
    c->launches += 2;
}

// Single-reduction CG (cg_variant 2): c->p holds the RESIDUAL (ghosted, peer-mapped: it is the SpMV
// input), c->rt the direction d, c->sdir = K~ d, c->Ap = w = K~ r.  Dots of the SpMV that produced w
// (delta = r.w in part_a, gamma = r.r in part_b) are consumed by the next step kernel.
void enqueue_spmv2(mbfea_ctx *c, int kernel)
{
    const int W = c->world();
    const int gs = spmv_grid(c, kernel);
    if (c->p2p) {
        run_spmv(c, 1, c->p.p, c->Ap.p, c->part_a.p, c->sc.p, c->peerctx.p, c->part_b.p);
        return;
    }
    halo_exchange(c, c->p.p);
    run_spmv(c, kernel, c->p.p, c->Ap.p, c->part_a.p, c->sc.p, nullptr, c->part_b.p);
    if (W > 1) {
        launch_reduce_partials(c->st, c->part_a.p, gs, c->red.p, c->part_b.p, c->red.p + 1, c->sc.p);
        NK(c->nc->AllGather(c->red.p, c->gath.p, 2, nccl::kFloat64, c->comm, c->st));
        c->launches++;
    }
}

void enqueue_iteration2(mbfea_ctx *c, int kernel)
{
    const int64_t nb = c->nb();
    const int W = c->world();
    const int gs = spmv_grid(c, kernel);
    PartialRef delta{c->part_a.p, gs, 1}, gamma{c->part_b.p, gs, 1};
    if (W > 1 && !c->p2p) { delta = PartialRef{c->gath.p, W, 2}; gamma = PartialRef{c->gath.p + 1, W, 2}; }
    launch_cg2_step(c->st, nb, delta, gamma, c->p.p, c->Ap.p, c->rt.p, c->sdir.p, c->y.p, c->sc.p,
                    c->p2p ? c->peerctx.p : nullptr, c->p2p ? c->push_idx.p : nullptr, c->p2p ? c->push_dst.p : nullptr,
                    c->p2p ? c->push_rows.p : nullptr, c->p2p ? (int)c->push_rows.n : 0, c->tickets.p + 2);
    c->launches++;
    enqueue_spmv2(c, kernel);
}

// Once per batch: the unscaled residual norm ||L r~||^2 and the stop test (never skipped, so
// sc->rr is current when the loop ends).
void enqueue_check(mbfea_ctx *c)
{
    const int64_t nb = c->nb();
    const int W = c->world();
    const int gv = vec_grid(nb);
    const double *res = cg_variant(c) == 2 ? c->p.p : c->rt.p;   // where the scaled residual lives
    double *part = c->part_c.p;   // part_a / part_b may hold dots the next step still needs
    if (c->p2p) {
        launch_cg_check(c->st, nb, c->lfac.p, res, nullptr, part, c->sc.p, c->peerctx.p);
        launch_cg_check_finish(c->st, PartialRef{part, gv, 1}, c->sc.p, c->peerctx.p);
    } else {
        launch_cg_check(c->st, nb, c->lfac.p, res, nullptr, part, c->sc.p, nullptr);
        PartialRef rr{part, gv, 1};
        if (W > 1) {
            launch_reduce_partials(c->st, part, gv, c->red.p + 4, nullptr, nullptr, nullptr);
            NK(c->nc->AllGather(c->red.p + 4, c->gath.p + 2 * W, 1, nccl::kFloat64, c->comm, c->st));
            rr = PartialRef{c->gath.p + 2 * W, W, 1};
            c->launches++;
        }
        launch_cg_check_finish(c->st, rr, c->sc.p, nullptr);
    }
    c->launches += 2;
}

void do_assemble(mbfea_ctx *c)
{
    launch_elem_geom(c->st, c->E, c->rec.p, c->node4.p, c->geom.p);                                  // K1
    launch_assemble(c->st, c->plan.nnzb(), c->contrib_ptr.p, c->contrib.p, c->geom.p, c->vals.p);  // K2
    c->launches += 2;
}

// K5: block-Jacobi as a symmetric scaling: factor the diagonal blocks, fetch the ghost rows'
// factors (multi-GPU, once), and form the stored blocks of K~ = S K S^T.
void do_precond(mbfea_ctx *c)
{
    CK(cudaMemsetAsync(c->flag.p, 0, sizeof(int), c->st));
    launch_block_chol(c->st, c->nb(), c->diag_idx.p, c->vals.p, c->lfac.p, c->sinv.p, c->flag.p);
    halo_exchange(c, c->sinv.p, 36);
    launch_scale_blocks(c->st, c->nb(), c->uptr.p, c->ucol.p, c->usrc.p, c->vals.p, c->sinv.p, c->vals_s.p,
                        c->plan.sym_half ? c->ilv_base.p : nullptr);
    c->launches += 2;
}

float elapsed(mbfea_ctx *c, int a, int b)
{
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, c->ev[a], c->ev[b]));
    return ms;
}

int need_job(mbfea_ctx *c)
{
    if (!c) return MBFEA_EARG;
    if (!c->have_job) return fail(c, MBFEA_ESTATE, "no job: call mbfea_set_job first");
    return MBFEA_OK;
}

void upload_forces(mbfea_ctx *c, int64_t nF, const int64_t *f_node, const int64_t *f_dof, const double *f_mag)
{
    const int64_t nb = c->nb();
    CK(cudaMemsetAsync(c->f.p, 0, (size_t)std::max<int64_t>(1, 6 * nb) * sizeof(double), c->st));
    if (nF > 0) {
        // one thread per force unless some (vertex, dof) is loaded twice: then the ordered one-thread kernel
        // keeps the accumulation order of the input (a bitmap of 6V bits finds out; validated indices)
        bool unique = true;
        {
            std::vector<uint64_t> seen(((size_t)6 * (size_t)c->V + 63) / 64, 0);
            for (int64_t i = 0; i < nF && unique; ++i) {
                const size_t bit = (size_t)6 * (size_t)f_node[i] + (size_t)f_dof[i];
                uint64_t &w = seen[bit >> 6];
                const uint64_t m = 1ull << (bit & 63);
                if (w & m) unique = false;
                w |= m;
            }
        }
        DevBuf<int64_t> &dn = c->stg_fnode, &dd = c->stg_fdof;   // staging kept between jobs (no cudaMalloc / cudaFree per call)
        DevBuf<double> &dm = c->stg_fmag;
        upload(c, dn, f_node, (size_t)nF);
        upload(c, dd, f_dof, (size_t)nF);
        upload(c, dm, f_mag, (size_t)nF);
        launch_scatter_forces(c->st, nF, dn.p, dd.p, dm.p, c->free_index.p, c->plan.row_begin, nb, c->f.p, unique);
        CK(cudaStreamSynchronize(c->st));
    }
    c->solved = c->recovered = false;
}

// Block pattern (and, when asked, the contributor lists) of the WHOLE reduced system on the device (plan_dev.cu), from
// the mesh that is already there; fm = vertex -> block row (-1 fixed), nb rows.  Returns the number of blocks.
int64_t device_pattern(mbfea_ctx *c, int64_t E, const int32_t *fm, int64_t nb, DevBuf<int32_t> &rowptr, DevBuf<int32_t> &colidx,
                       DevBuf<int32_t> &diag_idx, DevBuf<int32_t> &contrib_ptr, DevBuf<int32_t> &contrib)
{
    // scratch: entries per row / cursor, entry start, blocks per row, scan scratch, 64-bit entries
    CK(c->pd_cnt.alloc((size_t)nb + 1)); CK(c->pd_cur.alloc((size_t)nb + 1)); CK(c->pd_start.alloc((size_t)nb + 1));
    CK(c->pd_scan.alloc((size_t)scan_i32_scratch(nb + 1)));
    CK(cudaMemsetAsync(c->pd_cnt.p, 0, ((size_t)nb + 1) * sizeof(int32_t), c->st));
    CK(cudaMemsetAsync(c->pd_cur.p, 0, ((size_t)nb + 1) * sizeof(int32_t), c->st));
    launch_pd_count(c->st, E, c->stg_elems.p, fm, c->pd_cnt.p);
    launch_scan_i32(c->st, nb, c->pd_cnt.p, c->pd_start.p, c->pd_scan.p);
    int32_t nent = 0;
    CK(cudaMemcpyAsync(&nent, c->pd_start.p + nb, sizeof nent, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    CK(c->pd_ent.alloc((size_t)nent + 1));
    launch_pd_fill(c->st, E, c->stg_elems.p, fm, c->pd_start.p, c->pd_cur.p, c->pd_ent.p);
    launch_pd_sort(c->st, nb, c->pd_start.p, c->pd_ent.p, c->pd_cnt.p);   // pd_cnt now: blocks per row
    CK(rowptr.alloc((size_t)nb + 1));
    launch_scan_i32(c->st, nb, c->pd_cnt.p, rowptr.p, c->pd_scan.p);
    int32_t nnz32 = 0;
    CK(cudaMemcpyAsync(&nnz32, rowptr.p + nb, sizeof nnz32, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    CK(colidx.alloc((size_t)nnz32)); CK(diag_idx.alloc((size_t)nb));
    CK(contrib_ptr.alloc((size_t)nnz32 + 1)); CK(contrib.alloc((size_t)nent));
    launch_pd_emit(c->st, nb, c->pd_start.p, c->pd_ent.p, rowptr.p, colidx.p, diag_idx.p, contrib_ptr.p, contrib.p);
    return nnz32;
}

// mean beam length (the brick aggregates of the multilevel hierarchy are a few beam lengths wide): fixed chunks,
// summed in chunk order -- reproducible, and the same value on the host and the device path of the hierarchy
double mean_beam_length(int64_t V, int64_t E, const double *nodes_cm, const int32_t *elems_cm)
{
    if (E <= 0) return 0.0;
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(16, E >> 16));
    std::vector<double> part((size_t)nt, 0.0);
    std::vector<std::thread> th;
    auto body = [&](int t) {
        double len = 0.0;
        for (int64_t e = E * t / nt; e < E * (t + 1) / nt; ++e) {
            double d2 = 0.0;
            for (int a = 0; a < 3; ++a) {
                const double d = nodes_cm[(size_t)a * V + elems_cm[e]] - nodes_cm[(size_t)a * V + elems_cm[E + e]];
                d2 += d * d;
            }
            len += std::sqrt(d2);
        }
        part[(size_t)t] = len;
    };
    if (nt == 1) body(0);
    else {
        for (int t = 0; t < nt; ++t) th.emplace_back(body, t);
        for (auto &x : th) x.join();
    }
    double len = 0.0;
    for (int t = 0; t < nt; ++t) len += part[(size_t)t];
    return len / (double)E;
}

MlParams ml_params_of(const mbfea_ctx *c)
{
    MlParams mp;
    if (c->opt.ml_cell > 0.0) mp.cell = c->opt.ml_cell;
    if (c->opt.ml_ratio > 1.0) mp.ratio = c->opt.ml_ratio;
    if (c->opt.ml_coarsest_rows > 0) mp.coarsest_rows = std::min<int64_t>(c->opt.ml_coarsest_rows, 512);
    if (c->opt.ml_smooth_from > 0) mp.smooth_from = c->opt.ml_smooth_from;
    if (const char *e = std::getenv("MBFEA_ML_OMEGA")) { const double v = std::atof(e); if (v > 0.0 && v < 2.0) mp.omega = v; }
    if (const char *e = std::getenv("MBFEA_ML_THETA")) { const double v = std::atof(e); if (v > 0.0) mp.theta = v; }
    return mp;
}

}  // namespace

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

int mbfea_version(void) { return MBFEA_VERSION; }

void mbfea_abi_sizes(int64_t sizes3[3])
{
    if (!sizes3) return;
    sizes3[0] = (int64_t)sizeof(mbfea_options);
    sizes3[1] = (int64_t)sizeof(mbfea_stats);
    sizes3[2] = (int64_t)sizeof(mbfea_plan_sizes);
}

void mbfea_options_default(mbfea_options *o)
{
    if (!o) return;
    std::memset(o, 0, sizeof *o);
    o->device = -1;
    o->verbose = 0;
    o->tol = 1e-8;
    o->max_iter = 0;
    o->check_every = 0;
    o->spmv_kernel = 0;
    o->rank = 0;
    o->world = 1;
    o->use_graph = 1;
    o->precond = 0;
    o->ml_smooth_from = 0;
    o->ml_cell = 0.0;
    o->ml_ratio = 0.0;
    o->ml_coarsest_rows = 0;
}

const char *mbfea_status_string(int s)
{
    switch (s) {
        case MBFEA_OK: return "ok";
        case MBFEA_EPRECOND: return "precondition violated (the reference would Expects()-terminate)";
        case MBFEA_ERANGE_FIXED: return "Vertex index is outside of valid range.";
        case MBFEA_ERANGE_FORCE: return "Node index is out of valid range.";
        case MBFEA_ECUDA: return "CUDA/NCCL failure or no device";
        case MBFEA_ESTATE: return "call order violated";
        case MBFEA_ENOCONV: return "PCG did not converge within max_iter";
        case MBFEA_ESINGULAR: return "singular or non-SPD system";
        case MBFEA_EIO: return "file error";
        case MBFEA_EARG: return "bad argument";
        default: return "unknown status";
    }
}

const char *mbfea_last_error(const mbfea_ctx *c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int mbfea_create(mbfea_ctx **out, const mbfea_options *opts)
{
    if (!out) return fail(nullptr, MBFEA_EARG, "out is NULL");
    *out = nullptr;
    mbfea_options o;
    if (opts) o = *opts; else mbfea_options_default(&o);
    if (o.world < 1 || o.rank < 0 || o.rank >= o.world) return fail(nullptr, MBFEA_EARG, "bad rank/world");
    if (!(o.tol > 0.0)) o.tol = 1e-8;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return fail(nullptr, MBFEA_ECUDA,
                    std::string("no usable CUDA device (libmbfea has no CPU fallback): ") + cudaGetErrorString(e));
    }
    std::unique_ptr<mbfea_ctx> c(new (std::nothrow) mbfea_ctx);
    if (!c) return fail(nullptr, MBFEA_EARG, "host allocation failed");
    c->opt = o;
    int rc = guarded(nullptr, [&]() -> int {
        if (o.device >= 0) CK(cudaSetDevice(o.device));
        CK(cudaGetDevice(&c->device));
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, c->device));
        if (prop.major < 10) {
            return fail(nullptr, MBFEA_ECUDA, std::string("device '") + prop.name +
                                                  "' is not sm_100-class; libmbfea ships sm_100a code only");
        }
        c->sm_count = prop.multiProcessorCount;
        if (const char *e = std::getenv("MBFEA_PDL")) set_pdl(*e != '0');
        CK(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
        for (auto &ev : c->ev) CK(cudaEventCreate(&ev));
        CK(cudaMallocHost(reinterpret_cast<void **>(&c->sc_host), sizeof(PcgScalars)));
        set_sm_count(c->sm_count);
        CK(c->sc.alloc(1));
        CK(c->flag.alloc(1));
        CK(c->tickets.alloc(8));
        CK(cudaMemsetAsync(c->tickets.p, 0, 8 * sizeof(unsigned int), c->st));
        CK(c->part_a.alloc(kMaxPartials));
        CK(c->part_b.alloc(kMaxPartials));
        CK(c->part_c.alloc(kMaxPartials));
        CK(c->red.alloc(8));
        CK(c->gath.alloc((size_t)8 * o.world));
        return MBFEA_OK;
    });
    if (rc != MBFEA_OK) return rc;
    host_arrays_pin_on(c->device);   // from now on the host arrays that travel are page-locked (hostmem.cpp)
    *out = c.release();
    return MBFEA_OK;
}

void mbfea_destroy(mbfea_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->st) cudaStreamSynchronize(c->st);
    drop_graph(c);
    close_peer_vectors(c);
    for (size_t q = 0; q < c->peer_mailbox.size(); ++q)
        if ((int)q != c->opt.rank && c->peer_mailbox[q]) cudaIpcCloseMemHandle(c->peer_mailbox[q]);
    if (c->p2p && c->comm && c->nc && !c->comm_broken) {
        // rendezvous before the exported allocations (p, mailbox) are freed: every peer has unmapped them by
        // the time the all-gather completes (all ranks destroy their contexts collectively, like set_job)
        try {
            int64_t token = 0;
            std::vector<int64_t> tokens((size_t)c->world());
            allgather_host(c, &token, tokens.data(), sizeof token);
        } catch (...) { cudaGetLastError(); }
    }
    if (c->comm && c->nc) c->nc->CommDestroy(c->comm);
    for (auto &ev : c->ev)
        if (ev) cudaEventDestroy(ev);
    if (c->sc_host) cudaFreeHost(c->sc_host);
    if (c->st) cudaStreamDestroy(c->st);
    delete c;
}

int mbfea_comm_unique_id(uint8_t id128[128])
{
    if (!id128) return MBFEA_EARG;
    std::string msg;
    const nccl::Api *nc = nccl::load(msg);
    if (!nc) return fail(nullptr, MBFEA_ECUDA, msg);
    nccl::UniqueId id;
    if (int e = nc->GetUniqueId(&id)) return fail(nullptr, MBFEA_ECUDA, std::string("ncclGetUniqueId: ") + nc->GetErrorString(e));
    std::memcpy(id128, id.internal, 128);
    return MBFEA_OK;
}

int mbfea_comm_init(mbfea_ctx *c, const uint8_t id128[128], int32_t rank, int32_t world)
{
    if (!c || !id128) return MBFEA_EARG;
    if (world < 1 || rank < 0 || rank >= world) return fail(c, MBFEA_EARG, "bad rank/world");
    if (c->have_job) return fail(c, MBFEA_ESTATE, "comm_init must precede set_job");
    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        std::string msg;
        c->nc = nccl::load(msg);
        if (!c->nc) return fail(c, MBFEA_ECUDA, msg);
        nccl::UniqueId id;
        std::memcpy(id.internal, id128, 128);
        NK(c->nc->CommInitRank(&c->comm, world, id, rank));
        c->opt.rank = rank; c->opt.world = world;
        CK(c->gath.alloc((size_t)8 * world));
        setup_p2p_mailboxes(c);
        // warm the point-to-point connections now: the first ncclSend/Recv between two ranks sets up the
        // channel (seconds on 8 GPUs), which otherwise lands inside the first timed assemble (ghost S factors)
        if (world > 1) {
            DevBuf<double> w;
            CK(w.alloc((size_t)2 * world));
            NK(c->nc->GroupStart());
            for (int q = 0; q < world; ++q) {
                if (q == rank) continue;
                NK(c->nc->Send(w.p + rank, 1, nccl::kFloat64, q, c->comm, c->st));
                NK(c->nc->Recv(w.p + world + q, 1, nccl::kFloat64, q, c->comm, c->st));
            }
            NK(c->nc->GroupEnd());
            NK(c->nc->AllReduce(c->flag.p, c->flag.p, 1, nccl::kInt32, nccl::kMax, c->comm, c->st));
            NK(c->nc->Broadcast(w.p, w.p, 1, nccl::kFloat64, 0, c->comm, c->st));
            CK(cudaStreamSynchronize(c->st));
        }
        return MBFEA_OK;
    });
}

int mbfea_set_job(mbfea_ctx *c, int64_t V, int64_t E, const double *nodes_cm, const int32_t *elems_cm,
                  const double *normals_cm, const float *dims_bh, const float *mat_nuE, int64_t F,
                  const int32_t *fixed, int64_t nF, const int64_t *f_node, const int64_t *f_dof,
                  const double *f_mag)
{
    if (!c) return MBFEA_EARG;
    if (F < 0 || nF < 0) return fail(c, MBFEA_EARG, "negative count");
    if ((V > 0 && !nodes_cm) || (E > 0 && (!elems_cm || !normals_cm || !dims_bh || !mat_nuE)) || (F > 0 && !fixed) ||
        (nF > 0 && (!f_node || !f_dof || !f_mag)))
        return fail(c, MBFEA_EARG, "NULL input array");
    if (c->world() > 1 && !c->comm) return fail(c, MBFEA_ESTATE, "world > 1 requires mbfea_comm_init first");
    c->have_job = c->assembled = c->solved = c->recovered = false;
    std::string msg;
    // one process per GPU on one node: each rank's host passes take their share of the cores (8 ranks x 16 threads on
    // 16 cores only thrash)
    struct Share { explicit Share(int w) { host_threads_share(w); } ~Share() { host_threads_share(1); } } share(c->opt.world);
    const double tv0 = now_ms();
    if (int rc = validate_job(V, E, elems_cm, dims_bh, mat_nuE, F, fixed, nF, f_node, f_dof, msg)) return fail(c, rc, msg);
    if (std::getenv("MBFEA_PREP_TRACE")) std::fprintf(stderr, "[mbfea set_job] %-28s %8.2f ms\n", "validation", now_ms() - tv0);

    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        drop_graph(c);
        std::memset(&c->stats, 0, sizeof c->stats);
        const double t0 = now_ms();
        c->V = V; c->E = E; c->F = F;
        const bool ptrace = std::getenv("MBFEA_PREP_TRACE") != nullptr;
        double tl = t0, host_ms = 0.0;
        auto lap = [&](const char *what) {
            if (!ptrace) return;
            cudaStreamSynchronize(c->st);
            const double n = now_ms();
            std::fprintf(stderr, "[mbfea set_job] %-28s %8.2f ms\n", what, n - tl);
            tl = n;
        };
        // The host arrays that travel are page-locked (hostmem.cpp), so every upload below is asynchronous: the device
        // copies one array while the host builds the next.
        // mesh -> device, repacked into aligned records; K3 (canonical map)
        {
            DevBuf<double> &d_nodes = c->stg_nodes, &d_normals = c->stg_normals;
            DevBuf<int32_t> &d_elems = c->stg_elems, &d_scratch = c->stg_scratch;
            upload(c, d_nodes, nodes_cm, (size_t)3 * V);
            upload(c, d_elems, elems_cm, (size_t)2 * E);
            upload(c, d_normals, normals_cm, (size_t)3 * E);
            upload(c, c->fixed, fixed, (size_t)F);
            CK(c->node4.alloc((size_t)V));
            CK(c->rec.alloc((size_t)E));
            CK(c->geom.alloc((size_t)E));
            CK(c->free_index.alloc((size_t)V));
            CK(c->free_canon.alloc((size_t)V));
            CK(d_scratch.alloc((size_t)V + (size_t)V / 1024 + 2));
            launch_pack_nodes(c->st, V, d_nodes.p, c->node4.p);
            launch_free_map(c->st, V, F, c->fixed.p, c->free_canon.p, d_scratch.p);
            lap("mesh upload + pack + K3");
        }
        const double tp0 = now_ms();
        // internal renumbering: 0 auto (only when the input vertex order has no locality), 1 never, 2 always
        const int reorder = c->opt.reorder == 1 ? 0 : (c->opt.reorder == 2 ? 2 : (c->opt.reorder == 3 ? 3 : 1));
        // The host half of the multilevel hierarchy runs on its own thread, next to the pattern build: its first phase
        // (positions of the block rows, brick aggregates, centroids) needs the free map only and starts as soon as
        // build_plan has it; the thread then waits at the gate for the block pattern.  (Several ranks: rank 0 builds the
        // hierarchy for everybody -- the ranks of a box share the host cores, W redundant builds would each get 1 / W of
        // them -- from its own pattern of the whole system.)
        c->use_ml = false;
        c->ml_on_device = false;
        c->ml.reset();
        struct MlHostJob {
            double mean_len = 0.0;
            bool failed = false;
            double ms = 0.0;
            std::mutex m;
            std::condition_variable cv;
            int pattern_state = 0;   // 0 under construction, 1 ready, 2 abandoned (batched job, or set_job failed)
            void signal(int st) { { std::lock_guard<std::mutex> g(m); if (pattern_state == 0) pattern_state = st; } cv.notify_all(); }
        } mlh;
        std::thread ml_thread;
        bool ml_started = false;
        // destruction order: the gate is released first (an exception on this thread must not leave the other one waiting), then joined
        struct JoinGuard { std::thread &t; ~JoinGuard() { if (t.joinable()) t.join(); } } ml_join{ml_thread};
        struct GateGuard { MlHostJob &j; ~GateGuard() { j.signal(2); } } ml_gate{mlh};
        const std::function<void()> start_ml_thread = [&]() {
            const int64_t nbg = c->plan.nb_global;
            const bool wanted = c->opt.precond == 2 || (c->opt.precond == 0 && nbg >= 4096);
            const bool possible = nbg >= 64 && (c->opt.world == 1 || c->plan.nb() > 0);
            if (!wanted || !possible || !(c->opt.world == 1 || c->opt.rank == 0)) return;
            c->ml.par = ml_params_of(c);
            ml_started = true;
            ml_thread = std::thread([&, V, E, F, nbg]() {
                try {
                    cudaSetDevice(c->device);   // page-locked allocations of this thread belong to this context's device
                    const double tw = now_ms();
                    const int W = c->opt.world;
                    // positions of the block rows (internal numbering; all of them: the coarse levels are global) and the
                    // mean beam length: the aggregates are bricks of a few beam lengths
                    c->ml_pos.resize((size_t)3 * nbg);
                    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(16, V >> 16));
                    std::vector<std::thread> th;
                    for (int t = 0; t < nt; ++t)
                        th.emplace_back([&, t]() {
                            for (int64_t v = V * t / nt; v < V * (t + 1) / nt; ++v) {
                                const int32_t r = c->plan.free_index[(size_t)v];
                                if (r >= 0) for (int a = 0; a < 3; ++a) c->ml_pos[(size_t)3 * r + a] = nodes_cm[(size_t)a * V + v];
                            }
                        });
                    for (auto &x : th) x.join();
                    mlh.mean_len = mean_beam_length(V, E, nodes_cm, elems_cm);
                    if (W == 1) {
                        const FinePatternGate gate = [&](const int32_t *&rp, const int32_t *&ci) {
                            std::unique_lock<std::mutex> g(mlh.m);
                            mlh.cv.wait(g, [&] { return mlh.pattern_state != 0; });
                            if (mlh.pattern_state != 1) return false;
                            rp = c->plan.rowptr.data(); ci = c->plan.colidx.data();
                            return true;
                        };
                        c->ml.build_hierarchy_host(nbg, nullptr, nullptr, c->ml_pos.data(), mlh.mean_len, 1, c->ml_host, &gate);
                        if (c->ml_host.levels() == 0) mlh.failed = true;
                    } else {
                        HostPlan &G = c->ml_gplan;   // kept between jobs (page-locked arrays: allocation is expensive)
                        build_plan(V, E, elems_cm, F, fixed, 0, 1, false, G, reorder, nodes_cm);
                        c->ml.build_hierarchy_host(nbg, G.rowptr.data(), G.colidx.data(), c->ml_pos.data(), mlh.mean_len, W, c->ml_host);
                        if (!ml_image_layout(c->ml_host, W, nbg, G.rowptr.data(), c->ml_layout)) mlh.failed = true;
                    }
                    mlh.ms = now_ms() - tw;
                } catch (...) {
                    mlh.failed = true;
                }
            });
        };
        build_plan_begin(V, E, elems_cm, F, fixed, c->opt.rank, c->opt.world, c->plan, reorder, nodes_cm);
        // section properties (float arithmetic of the reference, on the host) -> beam records
        c->props_host.resize((size_t)4 * E);
        derive_props(E, dims_bh, mat_nuE, c->props_host.data());
        upload(c, c->stg_props, c->props_host);
        launch_pack_elems(c->st, E, c->stg_elems.p, c->stg_normals.p, c->stg_props.p, c->rec.p);
        // One rank, canonical numbering: the block pattern, the contributor lists and the solver format are built on the
        // DEVICE (plan_dev.cu) from the mesh that is already there -- the same arrays as the host builders', entry for
        // entry -- and only the pattern comes back (the hierarchy build and the taps read it).  Several ranks, an
        // internally renumbered job or MBFEA_HOST_PLAN=1: the host builders and an upload.
        static const bool force_host_plan = [] { const char *e = std::getenv("MBFEA_HOST_PLAN"); return e && *e == '1'; }();
        const bool device_plan = !force_host_plan && c->opt.world == 1 && c->plan.perm.empty() && c->plan.nb_global > 0 &&
                                 4 * E < (int64_t)INT32_MAX - 8;
        const bool device_plan_multi = !force_host_plan && c->opt.world > 1 && c->plan.perm.empty() && c->plan.nb_global > 0 && c->plan.nb() > 0 &&
                                       4 * E < (int64_t)INT32_MAX - 8;
        int64_t g_nnz = -1;   // several ranks: blocks of the global pattern once it is on the device
        c->device_plan = device_plan || device_plan_multi;
        // ... and so is the symbolic half of the multilevel preconditioner (hier_dev.cu; MBFEA_HOST_HIER=1: host thread)
        static const bool force_host_hier = [] { const char *e = std::getenv("MBFEA_HOST_HIER"); return e && *e == '1'; }();
        const bool ml_dev_try = device_plan && !force_host_hier && c->plan.nb_global >= 64 &&
                                (c->opt.precond == 2 || (c->opt.precond == 0 && c->plan.nb_global >= 4096));
        // several ranks: rank 0 builds the hierarchy of the WHOLE system for everybody -- on its device too (global pattern,
        // hierarchy with rank-local fine aggregates, flat image assembled device-to-device), else on the host thread
        const bool ml_dev_multi = c->opt.world > 1 && c->opt.rank == 0 && !force_host_plan && !force_host_hier && c->plan.perm.empty() &&
                                  c->plan.nb_global >= 64 && c->plan.nb() > 0 && 4 * E < (int64_t)INT32_MAX - 8 &&
                                  (c->opt.precond == 2 || (c->opt.precond == 0 && c->plan.nb_global >= 4096));
        if (!ml_dev_try && !ml_dev_multi) start_ml_thread();
        const int64_t nb = c->nb(), ng_dev = 0;
        (void)ng_dev;
        int64_t nnzb_dev = 0;
        if (device_plan) {
            HostPlan &P = c->plan;
            nnzb_dev = device_pattern(c, E, c->free_canon.p, nb, c->rowptr, c->colidx, c->diag_idx, c->contrib_ptr, c->contrib);
            // the pattern on the host: hierarchy build, component detection, taps
            P.rowptr.resize((size_t)nb + 1); P.colidx.resize((size_t)nnzb_dev);
            CK(cudaMemcpyAsync(P.rowptr.data(), c->rowptr.p, ((size_t)nb + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, c->st));
            CK(cudaMemcpyAsync(P.colidx.data(), c->colidx.p, (size_t)nnzb_dev * sizeof(int32_t), cudaMemcpyDeviceToHost, c->st));
            CK(cudaStreamSynchronize(c->st));
            P.diag_idx.clear(); P.contrib_ptr.clear(); P.contrib.clear();
            P.ghost_global.clear(); P.ghost_owner.clear(); P.send_rows.clear(); P.send_dest.clear(); P.neighbors.clear();
            CK(c->send_rows.alloc(0));
            lap("pattern on the device + download");
        } else if (device_plan_multi) {
            // several ranks: every rank builds the pattern of the WHOLE system on its device (it holds the whole mesh; a few
            // milliseconds) and takes its share: own block rows, local column ids (ghosts after the owned columns, ascending
            // global order), own contributor lists.  Rank 0 keeps the global pattern for the hierarchy.
            HostPlan &P = c->plan;
            const int64_t nbg = P.nb_global, r0 = P.row_begin, r1 = P.row_end;
            g_nnz = device_pattern(c, E, c->free_canon.p, nbg, c->g_rowptr, c->g_colidx, c->g_diag, c->g_cptr, c->g_contrib);
            int32_t k0 = 0, k1 = 0, c0 = 0, c1 = 0;
            CK(cudaMemcpyAsync(&k0, c->g_rowptr.p + r0, sizeof k0, cudaMemcpyDeviceToHost, c->st));
            CK(cudaMemcpyAsync(&k1, c->g_rowptr.p + r1, sizeof k1, cudaMemcpyDeviceToHost, c->st));
            CK(cudaStreamSynchronize(c->st));
            CK(cudaMemcpyAsync(&c0, c->g_cptr.p + k0, sizeof c0, cudaMemcpyDeviceToHost, c->st));
            CK(cudaMemcpyAsync(&c1, c->g_cptr.p + k1, sizeof c1, cudaMemcpyDeviceToHost, c->st));
            int32_t *mark = c->pd_cnt.p, *gidx = c->pd_start.p;   // scratch of the pattern pass (nbg + 1 entries), free again
            CK(cudaMemsetAsync(mark, 0, ((size_t)nbg + 1) * sizeof(int32_t), c->st));
            launch_pl_mark(c->st, k0, k1, c->g_colidx.p, r0, r1, mark);
            launch_scan_i32(c->st, nbg, mark, gidx, c->pd_scan.p);
            int32_t ng32 = 0;
            CK(cudaMemcpyAsync(&ng32, gidx + nbg, sizeof ng32, cudaMemcpyDeviceToHost, c->st));
            CK(cudaStreamSynchronize(c->st));
            nnzb_dev = (int64_t)k1 - k0;
            CK(c->ghost_dev.alloc((size_t)ng32));
            launch_pl_ghosts(c->st, nbg, mark, gidx, c->ghost_dev.p);
            CK(c->rowptr.alloc((size_t)nb + 1)); CK(c->diag_idx.alloc((size_t)nb)); CK(c->colidx.alloc((size_t)nnzb_dev));
            CK(c->contrib_ptr.alloc((size_t)nnzb_dev + 1)); CK(c->contrib.alloc((size_t)(c1 - c0)));
            launch_pl_shift(c->st, nb + 1, c->g_rowptr.p + r0, k0, c->rowptr.p);
            launch_pl_shift(c->st, nb, c->g_diag.p + r0, k0, c->diag_idx.p);
            launch_pl_localize(c->st, k0, k1, c->g_colidx.p, r0, r1, gidx, c->colidx.p);
            launch_pl_shift(c->st, nnzb_dev + 1, c->g_cptr.p + k0, c0, c->contrib_ptr.p);
            if (c1 > c0) CK(cudaMemcpyAsync(c->contrib.p, c->g_contrib.p + c0, (size_t)(c1 - c0) * sizeof(int32_t), cudaMemcpyDeviceToDevice, c->st));
            // the local pattern and the ghost list on the host: halo plan, hierarchy slices, taps
            P.rowptr.resize((size_t)nb + 1); P.colidx.resize((size_t)nnzb_dev); P.ghost_global.resize((size_t)ng32);
            static_assert(sizeof(long long) == sizeof(int64_t), "ghost list is copied as is");
            CK(cudaMemcpyAsync(P.rowptr.data(), c->rowptr.p, ((size_t)nb + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, c->st));
            CK(cudaMemcpyAsync(P.colidx.data(), c->colidx.p, (size_t)nnzb_dev * sizeof(int32_t), cudaMemcpyDeviceToHost, c->st));
            if (ng32) CK(cudaMemcpyAsync(P.ghost_global.data(), c->ghost_dev.p, (size_t)ng32 * sizeof(int64_t), cudaMemcpyDeviceToHost, c->st));
            CK(cudaStreamSynchronize(c->st));
            P.diag_idx.clear(); P.contrib_ptr.clear(); P.contrib.clear();
            build_halo_lists(P);
            upload(c, c->send_rows, P.send_rows);
            lap("pattern on the device (rank's share) + download");
        } else {
            build_plan_pattern(E, elems_cm, true, c->plan);
        }
        if (c->opt.batch_mode != 1) find_components(c->plan, 1024, 32);
        else { c->plan.comp_ptr.clear(); c->plan.comp_max_rows = 0; }
        const bool batched = !c->plan.comp_ptr.empty();   // the batched kernel streams full plain rows
        mlh.signal(batched ? 2 : 1);   // the hierarchy thread may read the block pattern now (a batch has no use for it)
        bool ml_on_device = false, ml_image_on_device = false;
        if (ml_dev_multi) {
            const int64_t nbg = c->plan.nb_global;
            if (g_nnz < 0) g_nnz = device_pattern(c, E, c->free_canon.p, nbg, c->g_rowptr, c->g_colidx, c->g_diag, c->g_cptr, c->g_contrib);
            // (a builder object of its own: the global arrays are W times the size of the rank's share that c->ml holds
            //  afterwards -- one object for both would free and reallocate every level buffer twice per job)
            c->ml.par = ml_params_of(c);
            c->ml_global.par = c->ml.par;
            const double ml_len = mean_beam_length(V, E, nodes_cm, elems_cm);
            ml_image_on_device = c->ml_global.build_hierarchy_device(c->st, V, c->node4.p, c->free_canon.p, nbg, c->g_rowptr.p, c->g_colidx.p,
                                                                     g_nnz, ml_len, c->opt.world) &&
                                 c->ml_global.pack_image_device(c->st, c->opt.world, nbg, c->g_rowptr.p, g_nnz, c->ml_dimg, c->ml_layout.hd);
            c->ml_mean_len = ml_len;
            lap("global hierarchy + image on the device");
            if (!ml_image_on_device) start_ml_thread();
        }
        if (ml_dev_try && !batched) {
            c->ml.par = ml_params_of(c);
            const double ml_len = mean_beam_length(V, E, nodes_cm, elems_cm);
            ml_on_device = c->ml.build_hierarchy_device(c->st, V, c->node4.p, c->free_canon.p, nb, c->rowptr.p, c->colidx.p, nnzb_dev, ml_len);
            c->ml_mean_len = ml_len;
            lap("multilevel hierarchy on the device");
            if (!ml_on_device) start_ml_thread();   // (does not coarsen on the device's terms: the host builder decides)
        }
        // preconditioner: multilevel for a single large mesh (auto) or on request
        const bool ml_possible = !batched && c->plan.nb_global >= 64 && (c->opt.world == 1 || c->plan.nb() > 0);
        const bool ml_wanted = c->opt.precond == 2 || (c->opt.precond == 0 && c->plan.nb_global >= 4096);
        host_ms += now_ms() - tp0;
        const int64_t ng = c->ng(), nnzb = c->plan.nnzb();
        // the reduced numbering the kernels use: K3's map, composed with the internal renumbering when there is one
        if (c->plan.perm.empty()) {
            CK(cudaMemcpyAsync(c->free_index.p, c->free_canon.p, (size_t)V * sizeof(int32_t), cudaMemcpyDeviceToDevice, c->st));
        } else {
            DevBuf<int32_t> &d_perm = c->stg_perm;
            upload(c, d_perm, c->plan.perm);
            launch_apply_perm(c->st, V, c->free_canon.p, d_perm.p, c->free_index.p);
            CK(cudaStreamSynchronize(c->st));
        }
        if (!device_plan && !device_plan_multi) {
            upload(c, c->rowptr, c->plan.rowptr);
            upload(c, c->colidx, c->plan.colidx);
            upload(c, c->diag_idx, c->plan.diag_idx);
            upload(c, c->contrib_ptr, c->plan.contrib_ptr);
            upload(c, c->contrib, c->plan.contrib);
            upload(c, c->send_rows, c->plan.send_rows);
            lap("pattern upload");
        }
        // solver format: symmetric-half storage unless full rows are asked for (the TMA kernel
        // streams whole rows and has no transposed-reference path)
        const double tp1 = now_ms();
        const bool half = !batched && (c->opt.storage == 2 || (c->opt.storage == 0 && c->opt.spmv_kernel != 2));
        bool format_on_device = false;
        if ((device_plan || device_plan_multi) && half) {
            // upper rule on the device; an irregular numbering that pads the 32-row tiles by more than 5 % under it
            // (or a value array beyond 2^31 units) goes to the host builder, which balances the pair orientation
            const int64_t ntiles = (nb + 31) / 32;
            CK(c->uptr.alloc((size_t)nb + 1)); CK(c->lptr.alloc((size_t)nb + 1)); CK(c->ilv_base.alloc((size_t)ntiles + 1));
            CK(c->pd_tot.alloc(2));
            CK(cudaMemsetAsync(c->pd_tot.p, 0, 2 * sizeof(unsigned long long), c->st));
            int32_t *nu = c->pd_cnt.p, *nl = c->pd_cur.p, *tunits = c->pd_start.p;   // scratch of the pattern pass, free again
            launch_pf_count(c->st, nb, c->rowptr.p, c->colidx.p, nu, nl);
            launch_pf_tiles(c->st, nb, nu, tunits, c->pd_tot.p);
            unsigned long long tot[2] = {0, 0};
            CK(cudaMemcpyAsync(tot, c->pd_tot.p, sizeof tot, cudaMemcpyDeviceToHost, c->st));
            CK(cudaStreamSynchronize(c->st));
            static const bool no_balance = std::getenv("MBFEA_NO_BALANCE") != nullptr;
            const bool pads = tot[1] * 20 > tot[0] * 21 && !no_balance;
            if (!pads && tot[1] * 18 <= (unsigned long long)INT32_MAX) {
                launch_scan_i32(c->st, nb, nu, c->uptr.p, c->pd_scan.p);
                launch_scan_i32(c->st, nb, nl, c->lptr.p, c->pd_scan.p);
                launch_scan_i32(c->st, ntiles, tunits, c->ilv_base.p, c->pd_scan.p);
                c->n_stored = (int64_t)tot[0];
                c->n_refs = nnzb - nb - c->n_stored;
                c->tile_units = (int64_t)(tot[1] * 18);
                CK(c->ucol.alloc((size_t)c->n_stored)); CK(c->usrc.alloc((size_t)c->n_stored));
                CK(c->lcol.alloc((size_t)c->n_refs)); CK(c->lofs.alloc((size_t)c->n_refs));
                launch_pf_fill(c->st, nb, c->rowptr.p, c->colidx.p, c->diag_idx.p, c->uptr.p, c->lptr.p, c->ilv_base.p, c->ucol.p, c->usrc.p,
                               c->lcol.p, c->lofs.p);
                HostPlan &P = c->plan;
                P.sym_half = true; P.balanced = false;
                P.uptr.clear(); P.ucol.clear(); P.usrc.clear(); P.lptr.clear(); P.lcol.clear(); P.lblk.clear(); P.lofs.clear(); P.tile_base.clear();
                format_on_device = true;
                lap("solver format on the device");
            }
        }
        if (!format_on_device) {
            build_solver_format(c->plan, half);
            c->n_stored = (int64_t)c->plan.ucol.size();
            c->n_refs = (int64_t)c->plan.lcol.size();
            c->tile_units = c->plan.sym_half ? (int64_t)c->plan.tile_base.back() : 0;
            upload(c, c->uptr, c->plan.uptr);
            upload(c, c->ucol, c->plan.ucol);
            upload(c, c->usrc, c->plan.usrc);
            upload(c, c->lptr, c->plan.lptr);
            upload(c, c->lcol, c->plan.lcol);
            upload(c, c->lofs, c->plan.lofs);
            upload(c, c->ilv_base, c->plan.tile_base);
        }
        host_ms += now_ms() - tp1;
        const double tp2 = now_ms();
        build_tiles(c->plan.rowptr, kTmaCapBlocks, kTmaMaxRows, c->tile_row_host);
        if (!format_on_device) build_tiles(c->plan.uptr, kTmaCapBlocks, kTmaMaxRows, c->tile_row_s_host);
        else c->tile_row_s_host.clear();   // (tiles over the solver format serve the full-row TMA kernel only)
        c->tma_ok = true;
        for (size_t t = 0; t + 1 < c->tile_row_host.size(); ++t)
            if (c->plan.rowptr[c->tile_row_host[t + 1]] - c->plan.rowptr[c->tile_row_host[t]] > kTmaCapBlocks)
                c->tma_ok = false;  // a single row longer than a stage: row-thread kernel only
        host_ms += now_ms() - tp2;
        upload(c, c->tile_row, c->tile_row_host);
        upload(c, c->tile_row_s, c->tile_row_s_host);
        if (batched) {
            const size_t nc = c->plan.comp_ptr.size() - 1;
            upload(c, c->comp_ptr, c->plan.comp_ptr);
            CK(c->comp_iters.alloc(nc)); CK(c->comp_status.alloc(nc));
            CK(c->comp_rr.alloc(nc)); CK(c->comp_ff.alloc(nc));
        }
        lap("solver format upload");
        c->stats.ms_host_prep = host_ms;
        const double t1 = t0 + host_ms;   // ms_h2d below = everything that is not host preparation
        CK(c->sendbuf.alloc((size_t)36 * c->plan.send_rows.size()));
        CK(c->vals.alloc((size_t)nnzb * 36));
        // symmetric-half: interleaved tiles (16-byte units, padded to the longest row of each tile)
        CK(c->vals_s.alloc(c->plan.sym_half ? (size_t)c->tile_units * 2 : (size_t)c->n_stored * 36));
        CK(c->lfac.alloc((size_t)nb * 36));
        CK(c->sinv.alloc((size_t)(nb + ng) * 36));
        CK(c->f.alloc((size_t)6 * nb));
        CK(c->x.alloc((size_t)6 * nb));
        CK(c->y.alloc((size_t)6 * nb));
        CK(c->rt.alloc((size_t)6 * nb));
        CK(c->ft.alloc((size_t)6 * nb));
        CK(c->sdir.alloc((size_t)6 * nb));
        CK(c->Ap.alloc((size_t)6 * nb));
        if (c->p2p) {
            // The peers map my p (CUDA IPC) and I map theirs.  The mappings are KEPT from job to job as long as nobody's p
            // has to be reallocated (a re-submitted mesh: cudaIpcOpen/CloseMemHandle cost ~5 ms per peer, 35 ms on eight
            // ranks); when any rank must reallocate, everybody unmaps first -- the second all-gather is the rendezvous
            // that lets the owners free -- and setup_p2p_halo maps afresh.
            const int W = c->world();
            int64_t flag = (c->p.fits((size_t)6 * (size_t)(nb + ng)) && !c->peer_p.empty()) ? 0 : 1;
            std::vector<int64_t> flags((size_t)W);
            allgather_host(c, &flag, flags.data(), sizeof flag);
            bool any = false;
            for (int q = 0; q < W; ++q) any = any || flags[(size_t)q] != 0;
            if (any) {
                close_peer_vectors(c);
                int64_t token = 0;
                allgather_host(c, &token, flags.data(), sizeof token);
            }
        }
        CK(c->p.alloc((size_t)6 * (nb + ng)));
        CK(cudaMemsetAsync(c->p.p, 0, (size_t)std::max<int64_t>(1, 6 * (nb + ng)) * sizeof(double), c->st));
        CK(c->xg.alloc((size_t)6 * c->plan.nb_global));
        CK(c->U_cm.alloc((size_t)6 * V));
        CK(c->U_rm.alloc((size_t)6 * V));
        CK(c->F6.alloc((size_t)12 * E));
        lap("matrix/vector allocation");
        upload_forces(c, nF, f_node, f_dof, f_mag);
        CK(cudaStreamSynchronize(c->st));
        lap("load vector");
        if (ml_possible && ml_wanted) {
            const double tw = now_ms();
            const int64_t nbg = c->plan.nb_global;
            if (ml_thread.joinable()) ml_thread.join();
            c->ml_on_device = ml_on_device;
            if (ml_on_device) {
                c->use_ml = true;
            } else if (c->opt.world == 1) {
                c->use_ml = ml_started && !mlh.failed && c->ml.upload_hierarchy(c->st, c->ml_host, nbg, (int64_t)c->plan.rowptr[(size_t)nb], 0, nullptr);
            } else {
                // rank 0's flat image goes through the GPUs to everybody; each rank then uploads its own share of level 0 and
                // the replicated coarse levels
                const int W = c->opt.world;
                int64_t bytes = (mlh.failed || c->opt.rank != 0) ? 0 : c->ml_layout.hd.total_bytes;
                std::vector<int64_t> all((size_t)W);
                allgather_host(c, &bytes, all.data(), sizeof bytes);
                bytes = all[0];
                if (bytes > 0) {
                    DevBuf<unsigned char> &dimg = c->ml_dimg;   // kept between jobs
                    CK(dimg.alloc((size_t)bytes));
                    if (c->opt.rank == 0 && !ml_image_on_device) {
                        // the image is assembled on the device: header + every (page-locked) hierarchy array at its offset
                        CK(cudaMemcpyAsync(dimg.p, &c->ml_layout.hd, sizeof(MlImageHeader), cudaMemcpyHostToDevice, c->st));
                        for (const MlImageLayout::Part &pt : c->ml_layout.parts)
                            CK(cudaMemcpyAsync(dimg.p + pt.at, pt.p, pt.bytes, cudaMemcpyHostToDevice, c->st));
                    }
                    NK(c->nc->Broadcast(dimg.p, dimg.p, (size_t)bytes / 8, nccl::kInt64, 0, c->comm, c->st));
                    MlImageHeader hd;
                    CK(cudaMemcpyAsync(&hd, dimg.p, sizeof hd, cudaMemcpyDeviceToHost, c->st));
                    CK(cudaStreamSynchronize(c->st));
                    MlDist dist;
                    dist.rank = c->opt.rank; dist.world = W;
                    dist.row_begin = c->plan.row_begin; dist.row_end = c->plan.row_end;
                    dist.ghost_global = &c->plan.ghost_global;
                    dist.loc_rowptr = c->plan.rowptr.data(); dist.loc_colidx = c->plan.colidx.data();
                    dist.nc = c->nc; dist.comm = c->comm;
                    if (hd.magic != kMlImageMagic || hd.world != W ||
                        hd.own_first[c->opt.rank + 1] - hd.own_first[c->opt.rank] != (int64_t)c->plan.rowptr[(size_t)nb])
                        return fail(c, MBFEA_ECUDA, "multilevel hierarchy: the image of rank 0 does not match this rank's block pattern");
                    c->use_ml = c->ml.upload_from_image(c->st, dimg.p, hd, dist);
                }
                // every rank must take the same path
                int64_t ok = c->use_ml ? 1 : 0;
                allgather_host(c, &ok, all.data(), sizeof ok);
                for (int q = 0; q < W; ++q) if (!all[(size_t)q]) c->use_ml = false;
            }
            if (!ml_on_device) { c->ml.ms_symbolic = mlh.ms + (now_ms() - tw); c->ml_mean_len = mlh.mean_len; }
            if (!c->use_ml && c->opt.precond == 2 && c->opt.verbose)
                std::fprintf(stderr, "[mbfea] multilevel preconditioner requested but the mesh does not coarsen: block-Jacobi\n");
            c->stats.ms_ml_symbolic = c->ml.ms_symbolic;
            lap("multilevel hierarchy");
        }
        setup_p2p_halo(c);
        lap("peer mappings + push lists");
        c->stats.ms_h2d = now_ms() - t1;

        mbfea_stats &s = c->stats;
        s.n_vertices = V; s.n_beams = E; s.n_fixed = V - c->plan.nb_global;
        s.n_block_rows = nb; s.n_block_rows_global = c->plan.nb_global;
        s.nnz_blocks = nnzb; s.n_ghost_blocks = ng;
        s.n_components = batched ? (int32_t)(c->plan.comp_ptr.size() - 1) : 0;
        // algorithmic bytes: the solver SpMV streams 288 B + a 4 B column id per STORED
        // block, 8 B per transposed reference, and per row 48 B x + 48 B y + two 4 B row pointers;
        // a CG iteration adds 9 vector passes of 48 B per row
        s.bytes_spmv = 292.0 * (double)c->n_stored + 8.0 * (double)c->n_refs + 104.0 * (double)nb;
        s.bytes_pcg_iter = s.bytes_spmv + 432.0 * (double)nb;
        // multilevel: restriction reads r~ (48) + T (144, FP32) + member id (4); prolongation reads r~, T, aggregate id and
        // writes z~ (48); the p-update reads z~ in place of r~ (no change)
        s.bytes_ml_apply = c->use_ml ? (c->ml.half16 ? 312.0 : 440.0) * (double)nb : 0.0;   // T blocks: 80 B (16-bit + scale) or 144 B (FP32)
        if (c->use_ml) s.bytes_pcg_iter += s.bytes_ml_apply;
        s.precond = c->use_ml ? 2 : 1;
        s.ml_levels = c->use_ml ? (int32_t)c->ml.levels() : 0;
        s.bytes_assemble = 112.0 * (double)E + 288.0 * (double)nnzb;
        s.bytes_recover = 208.0 * (double)E + 96.0 * (double)E;
        c->have_job = true;
        if (c->opt.verbose)
            std::fprintf(stderr, "[mbfea] rank %d/%d: V=%lld E=%lld free=%lld rows=%lld blocks=%lld ghosts=%lld prep=%.1f ms h2d=%.1f ms\n",
                         c->opt.rank, c->opt.world, (long long)V, (long long)E, (long long)c->plan.nb_global,
                         (long long)nb, (long long)nnzb, (long long)ng, s.ms_host_prep, s.ms_h2d);
        if (c->opt.verbose && c->use_ml) {
            std::string rows;
            for (int64_t r : c->ml.rows) rows += " " + std::to_string(r);
            std::fprintf(stderr, "[mbfea] multilevel preconditioner: rows per level%s (hierarchy %.1f ms)\n", rows.c_str(), c->ml.ms_symbolic);
        }
        return MBFEA_OK;
    });
}

int mbfea_set_forces(mbfea_ctx *c, int64_t nF, const int64_t *f_node, const int64_t *f_dof, const double *f_mag)
{
    if (int rc = need_job(c)) return rc;
    if (nF < 0 || (nF > 0 && (!f_node || !f_dof || !f_mag))) return fail(c, MBFEA_EARG, "bad force arrays");
    std::string msg;
    if (int rc = validate_forces(c->V, nF, f_node, f_dof, msg)) return fail(c, rc, msg);
    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        upload_forces(c, nF, f_node, f_dof, f_mag);
        return MBFEA_OK;
    });
}

int mbfea_assemble(mbfea_ctx *c)
{
    if (int rc = need_job(c)) return rc;
    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        CK(cudaEventRecord(c->ev[0], c->st));
        do_assemble(c);
        CK(cudaEventRecord(c->ev[1], c->st));
        do_precond(c);
        CK(cudaEventRecord(c->ev[6], c->st));
        if (c->use_ml) {
            c->ml.build_numeric(c->st, c->rowptr.p, c->colidx.p, c->vals.p, c->lfac.p, c->flag.p);
            c->launches += c->ml.setup_launches;
        }
        CK(cudaEventRecord(c->ev[2], c->st));
        // every rank takes the same exit: a non-SPD block on one rank fails the job on all of them (otherwise
        // the others would go on into the solve's collectives and hang)
        if (c->world() > 1) NK(c->nc->AllReduce(c->flag.p, c->flag.p, 1, nccl::kInt32, nccl::kMax, c->comm, c->st));
        int flag = 0;
        CK(cudaMemcpyAsync(&flag, c->flag.p, sizeof(int), cudaMemcpyDeviceToHost, c->st));
        CK(cudaStreamSynchronize(c->st));
        c->stats.ms_assemble = elapsed(c, 0, 1);
        c->stats.ms_precond = elapsed(c, 1, 2);
        c->stats.ms_ml_setup = c->use_ml ? elapsed(c, 6, 2) : 0.0;
        c->assembled = true; c->solved = c->recovered = false;
        if (flag & 1) return fail(c, MBFEA_ESINGULAR,
                                  "a diagonal 6x6 block is not positive definite (isolated vertex, zero-length or zero-stiffness beam)");
        if (flag & 2) {
            if (c->opt.precond != 0)
                return fail(c, MBFEA_ESINGULAR, "the coarsest matrix of the multilevel preconditioner is not positive definite");
            c->use_ml = false;   // auto: block-Jacobi for this job (the flag was all-reduced: every rank decides alike)
            c->stats.precond = 1; c->stats.ml_levels = 0;
            drop_graph(c);
        }
        return MBFEA_OK;
    });
}

static int solve_once(mbfea_ctx *c);

int mbfea_solve(mbfea_ctx *c)
{
    if (int rc = need_job(c)) return rc;
    if (!c->assembled) return fail(c, MBFEA_ESTATE, "solve before assemble");
    int rc = solve_once(c);
    // options.precond = 0 (auto) promised a solve, not a particular preconditioner: if the multilevel-preconditioned
    // iteration breaks down or stalls (every rank sees the same scalars, so every rank decides alike), fall back to the
    // block-Jacobi iteration once
    if ((rc == MBFEA_ESINGULAR || rc == MBFEA_ENOCONV) && c->use_ml && c->opt.precond == 0) {
        if (c->opt.verbose) std::fprintf(stderr, "[mbfea] rank %d: multilevel PCG failed (%s); retrying with block-Jacobi\n", c->opt.rank, c->err.c_str());
        c->use_ml = false;
        c->stats.precond = 1; c->stats.ml_levels = 0;
        drop_graph(c);
        rc = solve_once(c);
    }
    return rc;
}

static int solve_once(mbfea_ctx *c)
{
    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        const int64_t nb = c->nb();
        const int W = c->world();
        const int kernel = c->p2p ? 1 : pick_spmv(c, 0);
        const int gv = vec_grid(nb);
        c->epoch++;
        // a solve that ended abnormally (timeout, breakdown mid-kernel) must not leak its tickets or its
        // timeout flag into this one
        CK(cudaMemsetAsync(c->tickets.p, 0, 8 * sizeof(unsigned int), c->st));
        if (c->p2p) peer_timeout_clear(c->st);
        long long max_iter = c->opt.max_iter;
        if (max_iter <= 0) max_iter = std::min<long long>(20LL * 6 * std::max<int64_t>(1, c->plan.nb_global), 5000000LL);
        int check = c->opt.check_every;
        if (check <= 0) {
            // aim at ~4 ms of device work between residual checks / host polls
            const double est_us = c->stats.bytes_pcg_iter / 5.0e6 + 8.0;
            check = (int)std::min(32.0, std::max(2.0, 4000.0 / est_us));
        }
        check += check & 1;  // parity alternates: keep batches even

        if (!c->plan.comp_ptr.empty()) {
            // ---- batched path: one CTA per mesh, own iteration count, own stop test ----
            const int nc = (int)c->plan.comp_ptr.size() - 1;
            const size_t smem = (size_t)c->plan.comp_max_rows * 6 * 4 * sizeof(double);
            CK((cudaError_t)cg_batched_prepare(smem));
            CK(cudaEventRecord(c->ev[0], c->st));
            launch_cg_batched(c->st, nc, smem, c->comp_ptr.p, c->uptr.p, c->ucol.p, c->vals_s.p, c->f.p, c->sinv.p,
                              c->lfac.p, c->y.p, c->opt.tol, (int)std::min<long long>(max_iter, INT32_MAX), 16,
                              c->comp_iters.p, c->comp_rr.p, c->comp_ff.p, c->comp_status.p);
            launch_unscale(c->st, nb, c->sinv.p, c->y.p, c->x.p);
            CK(cudaEventRecord(c->ev[1], c->st));
            c->launches += 2;
            // recomputed residual over the whole job: f~ = S f into ft first
            launch_cg_init(c->st, nb, c->f.p, c->sinv.p, c->rt.p, c->rt.p, c->ft.p, c->p.p, c->part_b.p, c->part_c.p);
            CK(cudaMemcpyAsync(c->p.p, c->y.p, (size_t)6 * nb * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
            run_spmv(c, 1, c->p.p, c->Ap.p, nullptr, nullptr, nullptr);
            launch_cg_check(c->st, nb, c->lfac.p, c->ft.p, c->Ap.p, c->part_a.p, nullptr, nullptr);
            launch_reduce_partials(c->st, c->part_a.p, gv, c->red.p, nullptr, nullptr, nullptr);
            c->launches += 4;
            std::vector<int32_t> its((size_t)nc), stat((size_t)nc);
            std::vector<double> crr((size_t)nc), cff((size_t)nc);
            double rr_true = 0.0;
            CK(cudaMemcpyAsync(its.data(), c->comp_iters.p, its.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, c->st));
            CK(cudaMemcpyAsync(stat.data(), c->comp_status.p, stat.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, c->st));
            CK(cudaMemcpyAsync(crr.data(), c->comp_rr.p, crr.size() * sizeof(double), cudaMemcpyDeviceToHost, c->st));
            CK(cudaMemcpyAsync(cff.data(), c->comp_ff.p, cff.size() * sizeof(double), cudaMemcpyDeviceToHost, c->st));
            CK(cudaMemcpyAsync(&rr_true, c->red.p, sizeof(double), cudaMemcpyDeviceToHost, c->st));
            CK(cudaStreamSynchronize(c->st));
            long long it_max = 0, it_sum = 0;
            double rr = 0.0, ff = 0.0, worst = 0.0;
            int n_bad = 0, n_max = 0;
            for (int b = 0; b < nc; ++b) {
                it_max = std::max<long long>(it_max, its[(size_t)b]); it_sum += its[(size_t)b];
                rr += crr[(size_t)b]; ff += cff[(size_t)b];
                if (cff[(size_t)b] > 0) worst = std::max(worst, std::sqrt(crr[(size_t)b] / cff[(size_t)b]));
                n_bad += stat[(size_t)b] == 3; n_max += stat[(size_t)b] == 2;
            }
            mbfea_stats &s = c->stats;
            s.ms_pcg = elapsed(c, 0, 1);
            s.iterations = it_max;
            s.rel_residual = worst;   // the worst mesh's own ||r|| / ||f||
            s.true_rel_residual = ff > 0 ? std::sqrt(rr_true / ff) : 0.0;
            s.restarts = 0;
            s.converged = (n_bad == 0 && n_max == 0) ? 1 : 0;
            c->solved = true; c->recovered = false;
            if (c->opt.verbose)
                std::fprintf(stderr, "[mbfea] batched CG: %d meshes, iterations max %lld mean %.1f, worst rel res %.3e, job true rel res %.3e, %.2f ms\n",
                             nc, it_max, (double)it_sum / nc, worst, s.true_rel_residual, s.ms_pcg);
            if (n_bad) return fail(c, MBFEA_ESINGULAR, "CG breakdown in " + std::to_string(n_bad) + " mesh(es) of the batch (not SPD: is every mesh held by a fixed vertex?)");
            if (n_max) return fail(c, MBFEA_ENOCONV, std::to_string(n_max) + " mesh(es) of the batch reached max_iter before the requested tolerance");
            if (!(s.true_rel_residual <= std::max(1e3 * c->opt.tol, 1e-6)))
                return fail(c, MBFEA_ESINGULAR, "recomputed residual disagrees with the CG recurrence: a mesh of the batch is singular");
            return MBFEA_OK;
        }
        const int variant = cg_variant(c);
        CK(cudaEventRecord(c->ev[0], c->st));
        // f~ = S f, y = 0, r~ = p = f~  (CG on K~ = S K S^T  ==  block-Jacobi PCG on K)
        if (variant == 2) {
            launch_cg2_init(c->st, nb, c->f.p, c->sinv.p, c->y.p, c->p.p, c->ft.p, c->rt.p, c->sdir.p, c->part_c.p);
            CK(cudaMemsetAsync(c->part_b.p, 0, sizeof(double) * kMaxPartials, c->st));
        } else {
            launch_cg_init(c->st, nb, c->f.p, c->sinv.p, c->y.p, c->rt.p, c->ft.p, c->p.p, c->part_b.p, c->part_c.p);
            if (c->use_ml) {   // p0 = z~0 = M r~0, rho0 = r~0 . z~0
                c->ml.apply(c->st, c->rt.p, c->sdir.p, c->part_b.p, gv);
                CK(cudaMemcpyAsync(c->p.p, c->sdir.p, (size_t)6 * nb * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
                c->launches += c->ml.kernels_per_apply();
            }
        }
        PartialRef rho{c->part_b.p, gv, 1}, ff{c->part_c.p, gv, 1};
        if (W > 1) {
            launch_reduce_partials(c->st, c->part_b.p, gv, c->red.p + 2, c->part_c.p, c->red.p + 3, nullptr);
            NK(c->nc->AllGather(c->red.p + 2, c->gath.p + W, 2, nccl::kFloat64, c->comm, c->st));
            rho = PartialRef{c->gath.p + W, W, 2};
            ff = PartialRef{c->gath.p + W + 1, W, 2};
        }
        unsigned long long *trace = nullptr;
        if (const char *tr = std::getenv("MBFEA_TRACE"); tr && *tr == '1') {
            CK(c->trace.alloc((size_t)kTraceSlots * kTraceIters));
            CK(cudaMemsetAsync(c->trace.p, 0, sizeof(unsigned long long) * kTraceSlots * kTraceIters, c->st));
            trace = c->trace.p;
        }
        launch_cg_init_finish(c->st, rho, ff, c->sc.p, c->opt.tol, max_iter, c->epoch, trace);
        c->launches += 2;
        if (c->p2p) {  // ghost blocks of p for iteration 0
            launch_push_halo(c->st, (int64_t)c->plan.send_rows.size(), c->send_rows.p, c->push_dst_send.p, c->p.p,
                             c->sc.p, c->peerctx.p);
            c->launches++;
        }
        if (variant == 2) enqueue_spmv2(c, kernel);   // w0 = K~ r0, gamma0, delta0

        bool use_graph = c->opt.use_graph != 0;
        if (use_graph && (c->graph_exec == nullptr || c->graph_iters != check || c->graph_kernel != kernel + 16 * variant)) {
            drop_graph(c);
            if (kernel == 2) prepare_tma(c);
            cudaGraph_t g = nullptr;
            const int64_t l0 = c->launches;
            CK(cudaStreamBeginCapture(c->st, cudaStreamCaptureModeThreadLocal));
            try {
                for (int i = 0; i < check; ++i) {
                    if (variant == 2) enqueue_iteration2(c, kernel); else enqueue_iteration(c, kernel, i & 1);
                }
                enqueue_check(c);
            } catch (...) {
                cudaStreamEndCapture(c->st, &g);
                if (g) cudaGraphDestroy(g);
                throw;
            }
            CK(cudaStreamEndCapture(c->st, &g));
            c->graph_launches = c->launches - l0;
            c->launches = l0;
            cudaError_t e = cudaGraphInstantiate(&c->graph_exec, g, 0);
            cudaGraphDestroy(g);
            CK(e);
            c->graph_iters = check; c->graph_kernel = kernel + 16 * variant;
        }
        // Outer loop: run batches until the recurrence says "converged", then RECOMPUTE the residual
        // r~ = f~ - K~ y.  If the recomputed residual misses the tolerance (CG's residual gap on
        // ill-conditioned systems) replace the recurrence residual by it and resume -- so the tolerance
        // holds for the true residual ||f - K u||, as far as FP64 allows.
        double rr_true = 0.0, prev_true = INFINITY;
        int restarts = 0;
        for (;;) {
            for (;;) {
                if (use_graph) {
                    CK(cudaGraphLaunch(c->graph_exec, c->st));
                    c->launches += c->graph_launches;
                } else {
                    for (int i = 0; i < check; ++i) {
                        if (variant == 2) enqueue_iteration2(c, kernel); else enqueue_iteration(c, kernel, i & 1);
                    }
                    enqueue_check(c);
                }
                CK(cudaMemcpyAsync(c->sc_host, c->sc.p, sizeof(PcgScalars), cudaMemcpyDeviceToHost, c->st));
                CK(cudaStreamSynchronize(c->st));
                if (c->sc_host->done != 0) break;
                if (c->p2p && peer_timeout_flag()) {
                    c->comm_broken = true;
                    return fail(c, MBFEA_ECUDA, "a peer rank stopped signalling during the CG loop (NVLink mailbox wait timed out)");
                }
            }
            // f - K x with the ASSEMBLED matrix (unscaled, full pattern, bit-identical to the oracle's K) and x = S^T y,
            // through the ghosted buffer (its previous content -- direction or residual -- is rebuilt by the
            // replacement below if the iteration resumes).  The iteration itself multiplies with K~ = S K S^T, whose
            // entries are rounded: on an ill-conditioned system (planar 100 x 100 cantilever, kappa ~ 1e10) that
            // perturbation alone moves the solution by 4e-8, which only a residual against K itself can see.
            launch_unscale(c->st, nb, c->sinv.p, c->y.p, c->p.p);
            halo_exchange(c, c->p.p);
            run_spmv_full(c, 1, c->p.p, c->Ap.p);
            c->launches++;
            if (variant == 2)
                launch_cg_restart(c->st, nb, c->sinv.p, c->f.p, c->Ap.p, c->p.p, c->rt.p, 0.0, c->sdir.p, c->part_a.p, c->part_b.p);
            else
                launch_cg_restart(c->st, nb, c->sinv.p, c->f.p, c->Ap.p, c->rt.p, c->p.p, 1.0, nullptr, c->part_a.p, c->part_b.p);
            launch_reduce_partials(c->st, c->part_a.p, gv, c->red.p, c->part_b.p, c->red.p + 1, nullptr);
            c->launches += 3;
            PartialRef trr{c->red.p, 1, 1}, trho{c->red.p + 1, 1, 1};
            rr_true = 0.0;
            if (W > 1) {
                NK(c->nc->AllGather(c->red.p, c->gath.p, 2, nccl::kFloat64, c->comm, c->st));
                trr = PartialRef{c->gath.p, W, 2}; trho = PartialRef{c->gath.p + 1, W, 2};
                std::vector<double> g((size_t)2 * W);
                CK(cudaMemcpyAsync(g.data(), c->gath.p, g.size() * sizeof(double), cudaMemcpyDeviceToHost, c->st));
                CK(cudaStreamSynchronize(c->st));
                for (int q = 0; q < W; ++q) rr_true += g[(size_t)2 * q];
            } else {
                CK(cudaMemcpyAsync(&rr_true, c->red.p, sizeof(double), cudaMemcpyDeviceToHost, c->st));
                CK(cudaStreamSynchronize(c->st));
            }
            const PcgScalars &hh = *c->sc_host;
            const bool met_true = rr_true <= hh.tol2 * hh.ff;
            if (hh.done != 1 || met_true || restarts >= 6 || !(rr_true < 0.25 * prev_true) || hh.iter >= hh.max_iter) break;
            prev_true = rr_true;
            ++restarts;
            c->epoch++;
            if (c->use_ml) {   // resume from the replaced residual: p = z~ = M r~, rho = r~ . z~
                c->ml.apply(c->st, c->rt.p, c->sdir.p, c->part_b.p, gv);
                CK(cudaMemcpyAsync(c->p.p, c->sdir.p, (size_t)6 * nb * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
                launch_reduce_partials(c->st, c->part_b.p, gv, c->red.p + 1, nullptr, nullptr, nullptr);
                if (W > 1) NK(c->nc->AllGather(c->red.p, c->gath.p, 2, nccl::kFloat64, c->comm, c->st));   // trr unchanged, trho = r~.z~
                c->launches += 1 + c->ml.kernels_per_apply();
            }
            launch_cg_restart_finish(c->st, trr, trho, c->sc.p, c->epoch, 1);
            c->launches++;
            if (c->p2p) {
                launch_push_halo(c->st, (int64_t)c->plan.send_rows.size(), c->send_rows.p, c->push_dst_send.p, c->p.p,
                                 c->sc.p, c->peerctx.p);
                c->launches++;
            }
            if (variant == 2) enqueue_spmv2(c, kernel);
        }
        launch_unscale(c->st, nb, c->sinv.p, c->y.p, c->x.p);   // x = S^T y
        CK(cudaEventRecord(c->ev[1], c->st));
        CK(cudaStreamSynchronize(c->st));
        const PcgScalars &h = *c->sc_host;
        mbfea_stats &s = c->stats;
        s.ms_pcg = elapsed(c, 0, 1);
        s.iterations = h.iter;
        s.rel_residual = h.ff > 0 ? std::sqrt(h.rr / h.ff) : 0.0;
        s.true_rel_residual = h.ff > 0 ? std::sqrt(rr_true / h.ff) : 0.0;
        // done == 1 is either the residual test or exact convergence (rho below FP64 resolution):
        // either way the recurrence residual at exit decides
        // converged = 1: the RECOMPUTED residual meets the tolerance.  2: the recurrence met it but the
        // recomputed residual stalls slightly above (attainable FP64 accuracy of this system).
        // On a singular (under-constrained) system the iterates blow up while the recurrence residual
        // keeps shrinking; the recomputed residual exposes it.
        const bool met_rec = h.ff == 0.0 || h.rr <= h.tol2 * h.ff || restarts > 0;
        const bool met_true = h.ff == 0.0 || rr_true <= h.tol2 * h.ff;
        const bool near_true = s.true_rel_residual <= std::max(1e3 * c->opt.tol, 1e-6);
        s.converged = (h.done == 1 && met_true) ? 1 : ((h.done == 1 && met_rec && near_true) ? 2 : 0);
        const bool diverged = h.done == 1 && met_rec && !near_true;
        s.restarts = restarts;
        c->solved = true; c->recovered = false;
        if (c->opt.verbose)
            std::fprintf(stderr, "[mbfea] rank %d: %s %lld its, rel res %.3e (true %.3e), %.2f ms, spmv kernel %d (%s storage), batch %d\n",
                         c->opt.rank, c->use_ml ? "multilevel PCG" : "CG", (long long)h.iter, s.rel_residual, s.true_rel_residual, s.ms_pcg, kernel,
                         c->plan.sym_half ? (c->plan.balanced ? "symmetric-half, balanced" : "symmetric-half") : "full-row", check);
        if (c->opt.verbose && restarts)
            std::fprintf(stderr, "[mbfea] rank %d: %d residual replacement(s)\n", c->opt.rank, restarts);
        if (c->p2p && peer_timeout_flag()) {
            c->comm_broken = true;
            return fail(c, MBFEA_ECUDA, "a peer rank stopped signalling during the PCG loop (NVLink mailbox wait timed out)");
        }
        if (h.done == 3) return fail(c, MBFEA_ESINGULAR, "CG breakdown: p.Ap <= 0 or non-finite (system not SPD: is any vertex fixed?)");
        if (diverged)
            return fail(c, MBFEA_ESINGULAR, "recomputed residual ||f - K u|| disagrees with the CG recurrence: the system is singular "
                                            "(is every connected component held by a fixed vertex?)");
        if (h.done == 2) return fail(c, MBFEA_ENOCONV, "CG reached max_iter before the requested tolerance");
        if (!s.converged) return fail(c, MBFEA_ENOCONV, "CG stagnated above the requested tolerance (residual at FP64 resolution)");
        return MBFEA_OK;
    });
}

int mbfea_recover(mbfea_ctx *c)
{
    if (int rc = need_job(c)) return rc;
    if (!c->solved) return fail(c, MBFEA_ESTATE, "recover before solve");
    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        const int64_t nb = c->nb();
        CK(cudaEventRecord(c->ev[0], c->st));
        if (c->world() > 1) {
            for (int q = 0; q < c->world(); ++q) {
                const int64_t b = c->row_begin_of(q), e = c->row_begin_of(q + 1);
                if (e > b)
                    NK(c->nc->Broadcast(c->x.p, c->xg.p + 6 * b, (size_t)(6 * (e - b)), nccl::kFloat64, q, c->comm, c->st));
            }
        } else if (nb > 0) {
            CK(cudaMemcpyAsync(c->xg.p, c->x.p, (size_t)6 * nb * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
        }
        launch_expand(c->st, c->V, c->free_index.p, c->xg.p, c->U_cm.p, c->U_rm.p);
        launch_element_forces(c->st, c->E, c->rec.p, c->node4.p, c->U_rm.p, c->F6.p);
        c->launches += 2;
        CK(cudaEventRecord(c->ev[1], c->st));
        CK(cudaStreamSynchronize(c->st));
        c->stats.ms_recover = elapsed(c, 0, 1);
        c->recovered = true;
        return MBFEA_OK;
    });
}

int mbfea_execute(mbfea_ctx *c)
{
    if (int rc = need_job(c)) return rc;
    const double t0 = now_ms();
    c->launches = 0;
    int rc = mbfea_assemble(c);
    if (rc != MBFEA_OK) return rc;
    int rs = mbfea_solve(c);
    if (rs != MBFEA_OK && rs != MBFEA_ENOCONV && rs != MBFEA_ESINGULAR) return rs;
    const std::string keep = c->err;
    rc = mbfea_recover(c);
    if (rc != MBFEA_OK) return rc;
    c->stats.ms_execute_total = now_ms() - t0;
    c->stats.kernel_launches = c->launches;
    if (rs != MBFEA_OK) c->err = keep;
    return rs;
}

int mbfea_get_displacements(mbfea_ctx *c, double *U_cm)
{
    if (int rc = need_job(c)) return rc;
    if (!U_cm) return fail(c, MBFEA_EARG, "U_cm is NULL");
    if (!c->recovered) return fail(c, MBFEA_ESTATE, "no results: call mbfea_execute first");
    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        const double t0 = now_ms();
        CK(cudaMemcpyAsync(U_cm, c->U_cm.p, (size_t)6 * c->V * sizeof(double), cudaMemcpyDeviceToHost, c->st));
        CK(cudaStreamSynchronize(c->st));
        c->stats.ms_d2h += now_ms() - t0;
        return MBFEA_OK;
    });
}

int mbfea_get_edge_forces(mbfea_ctx *c, double *F6x2E)
{
    if (int rc = need_job(c)) return rc;
    if (!F6x2E) return fail(c, MBFEA_EARG, "F6x2E is NULL");
    if (!c->recovered) return fail(c, MBFEA_ESTATE, "no results: call mbfea_execute first");
    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        const double t0 = now_ms();
        CK(cudaMemcpyAsync(F6x2E, c->F6.p, (size_t)12 * c->E * sizeof(double), cudaMemcpyDeviceToHost, c->st));
        CK(cudaStreamSynchronize(c->st));
        c->stats.ms_d2h += now_ms() - t0;
        return MBFEA_OK;
    });
}

// ---- colour-mapping pre-pass ------------------------------------
static int force_ranges_device(mbfea_ctx *c, DevBuf<double> &range)
{
    const int64_t n = 2 * c->E;
    DevBuf<double> pmin, pmax;
    CK(pmin.alloc((size_t)6 * force_minmax_parts(n)));
    CK(pmax.alloc((size_t)6 * force_minmax_parts(n)));
    CK(range.alloc(12));
    launch_force_ranges(c->st, n, c->F6.p, pmin.p, pmax.p, range.p);
    CK(cudaStreamSynchronize(c->st));
    return MBFEA_OK;
}

int mbfea_get_force_ranges(mbfea_ctx *c, double mins6[6], double maxs6[6])
{
    if (int rc = need_job(c)) return rc;
    if (!mins6 || !maxs6) return fail(c, MBFEA_EARG, "NULL");
    if (!c->recovered) return fail(c, MBFEA_ESTATE, "no results: call mbfea_execute first");
    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        DevBuf<double> range;
        force_ranges_device(c, range);
        double h[12];
        CK(cudaMemcpy(h, range.p, sizeof h, cudaMemcpyDeviceToHost));
        std::copy(h, h + 6, mins6);
        std::copy(h + 6, h + 12, maxs6);
        return MBFEA_OK;
    });
}

int mbfea_get_edge_forces_normalized(mbfea_ctx *c, double *T6x2E)
{
    if (int rc = need_job(c)) return rc;
    if (!T6x2E) return fail(c, MBFEA_EARG, "NULL");
    if (!c->recovered) return fail(c, MBFEA_ESTATE, "no results: call mbfea_execute first");
    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        DevBuf<double> range, out;
        force_ranges_device(c, range);
        CK(out.alloc((size_t)12 * c->E));
        launch_force_normalize(c->st, 2 * c->E, c->F6.p, range.p, out.p);
        CK(cudaMemcpyAsync(T6x2E, out.p, (size_t)12 * c->E * sizeof(double), cudaMemcpyDeviceToHost, c->st));
        CK(cudaStreamSynchronize(c->st));
        return MBFEA_OK;
    });
}

int mbfea_get_stats(const mbfea_ctx *c, mbfea_stats *out)
{
    if (!c || !out) return MBFEA_EARG;
    *out = c->stats;
    return MBFEA_OK;
}

// ---- result files ----------------------------------------------------------
static int write_csv(mbfea_ctx *c, const char *path, const std::vector<double> &rowmajor, int64_t rows, int cols)
{
    std::FILE *fp = std::fopen(path, "w");
    if (!fp) return fail(c, MBFEA_EIO, std::string("cannot open ") + path);
    for (int64_t r = 0; r < rows; ++r) {
        for (int j = 0; j < cols; ++j)
            std::fprintf(fp, j + 1 < cols ? "%.14g," : "%.14g\n", rowmajor[(size_t)r * cols + j]);
    }
    std::fclose(fp);
    return MBFEA_OK;
}

int mbfea_write_displacements_csv(mbfea_ctx *c, const char *path)
{
    if (int rc = need_job(c)) return rc;
    if (!path) return fail(c, MBFEA_EARG, "path is NULL");
    if (!c->recovered) return fail(c, MBFEA_ESTATE, "no results");
    std::vector<double> U((size_t)6 * c->V);
    int rc = guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        CK(cudaMemcpyAsync(U.data(), c->U_rm.p, U.size() * sizeof(double), cudaMemcpyDeviceToHost, c->st));
        CK(cudaStreamSynchronize(c->st));
        return MBFEA_OK;
    });
    if (rc) return rc;
    return write_csv(c, path, U, c->V, 6);
}

int mbfea_write_element_forces_csv(mbfea_ctx *c, const char *path)
{
    if (int rc = need_job(c)) return rc;
    if (!path) return fail(c, MBFEA_EARG, "path is NULL");
    if (!c->recovered) return fail(c, MBFEA_ESTATE, "no results");
    const int64_t E = c->E;
    std::vector<double> F((size_t)12 * E), rows((size_t)12 * E);
    int rc = guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        CK(cudaMemcpyAsync(F.data(), c->F6.p, F.size() * sizeof(double), cudaMemcpyDeviceToHost, c->st));
        CK(cudaStreamSynchronize(c->st));
        return MBFEA_OK;
    });
    if (rc) return rc;
    for (int64_t e = 0; e < E; ++e)
        for (int k = 0; k < 6; ++k) {
            rows[(size_t)e * 12 + k] = F[(size_t)k * 2 * E + 2 * e];
            rows[(size_t)e * 12 + 6 + k] = F[(size_t)k * 2 * E + 2 * e + 1];
        }
    return write_csv(c, path, rows, E, 12);
}

// ---- parity / debug taps ---------------------------------------------------
int mbfea_get_props(mbfea_ctx *c, double *props4)
{
    if (int rc = need_job(c)) return rc;
    if (!props4) return fail(c, MBFEA_EARG, "NULL");
    for (size_t i = 0; i < c->props_host.size(); ++i) props4[i] = (double)c->props_host[i];
    return MBFEA_OK;
}

int mbfea_get_dofmap(mbfea_ctx *c, int32_t *free_index)
{
    if (int rc = need_job(c)) return rc;
    if (!free_index) return fail(c, MBFEA_EARG, "NULL");
    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        CK(cudaMemcpyAsync(free_index, c->free_canon.p, (size_t)c->V * sizeof(int32_t), cudaMemcpyDeviceToHost, c->st));
        CK(cudaStreamSynchronize(c->st));
        return MBFEA_OK;
    });
}

int mbfea_get_kelem(mbfea_ctx *c, double *Ke)
{
    if (int rc = need_job(c)) return rc;
    if (!Ke) return fail(c, MBFEA_EARG, "NULL");
    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        DevBuf<double> d;
        CK(d.alloc((size_t)144 * c->E));
        launch_kelem(c->st, c->E, c->rec.p, c->node4.p, d.p);
        CK(cudaMemcpyAsync(Ke, d.p, (size_t)144 * c->E * sizeof(double), cudaMemcpyDeviceToHost, c->st));
        CK(cudaStreamSynchronize(c->st));
        return MBFEA_OK;
    });
}

int mbfea_get_bsr_sizes(mbfea_ctx *c, int64_t *nbr, int64_t *nnzb, int64_t *row_offset, int64_t *n_ghost)
{
    if (int rc = need_job(c)) return rc;
    if (nbr) *nbr = c->nb();
    if (nnzb) *nnzb = c->plan.nnzb();
    if (row_offset) *row_offset = c->plan.row_begin;
    if (n_ghost) *n_ghost = c->ng();
    return MBFEA_OK;
}

// canonical reduced id of an internal global row / of a local column id
static inline int64_t canon_of(const mbfea_ctx *c, int64_t internal_global)
{
    return c->plan.inv.empty() ? internal_global : (int64_t)c->plan.inv[(size_t)internal_global];
}

// owned rows in ascending canonical order: order[j] = local row
static std::vector<int32_t> canonical_row_order(const mbfea_ctx *c)
{
    const int64_t nb = c->nb();
    std::vector<int32_t> order((size_t)nb);
    for (int64_t r = 0; r < nb; ++r) order[(size_t)r] = (int32_t)r;
    if (!c->plan.inv.empty())
        std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
            return c->plan.inv[(size_t)(c->plan.row_begin + a)] < c->plan.inv[(size_t)(c->plan.row_begin + b)];
        });
    return order;
}

int mbfea_get_bsr(mbfea_ctx *c, int64_t *rowptr, int32_t *colidx, double *vals, int64_t *row_ids)
{
    if (int rc = need_job(c)) return rc;
    if (vals && !c->assembled) return fail(c, MBFEA_ESTATE, "values requested before assemble");
    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        const int64_t nb = c->nb(), nnzb = c->plan.nnzb(), r0 = c->plan.row_begin;
        const hvec<int32_t> &rp = c->plan.rowptr, &ci = c->plan.colidx;
        std::vector<double> v;
        if (vals) {
            v.resize((size_t)nnzb * 36);
            CK(cudaMemcpyAsync(v.data(), c->vals.p, v.size() * sizeof(double), cudaMemcpyDeviceToHost, c->st));
            CK(cudaStreamSynchronize(c->st));
        }
        const std::vector<int32_t> order = canonical_row_order(c);
        int64_t out = 0;
        std::vector<std::pair<int64_t, int32_t>> cols;   // (canonical column, position)
        for (int64_t j = 0; j < nb; ++j) {
            const int32_t r = order[(size_t)j];
            if (rowptr) rowptr[j] = out;
            if (row_ids) row_ids[j] = canon_of(c, r0 + r);
            cols.clear();
            for (int32_t k = rp[(size_t)r]; k < rp[(size_t)r + 1]; ++k) {
                const int32_t lc = ci[(size_t)k];
                const int64_t g = lc < nb ? r0 + lc : c->plan.ghost_global[(size_t)(lc - nb)];
                cols.emplace_back(canon_of(c, g), k);
            }
            std::sort(cols.begin(), cols.end());
            for (const auto &pc : cols) {
                if (colidx) colidx[out] = (int32_t)pc.first;
                if (vals) std::memcpy(vals + out * 36, v.data() + (size_t)pc.second * 36, 36 * sizeof(double));
                ++out;
            }
        }
        if (rowptr) rowptr[nb] = out;
        return MBFEA_OK;
    });
}

int mbfea_get_load(mbfea_ctx *c, double *f_reduced)
{
    if (int rc = need_job(c)) return rc;
    if (!f_reduced) return fail(c, MBFEA_EARG, "NULL");
    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        const int64_t nb = c->nb();
        std::vector<double> f((size_t)6 * nb);
        CK(cudaMemcpyAsync(f.data(), c->f.p, f.size() * sizeof(double), cudaMemcpyDeviceToHost, c->st));
        CK(cudaStreamSynchronize(c->st));
        const std::vector<int32_t> order = canonical_row_order(c);
        for (int64_t j = 0; j < nb; ++j) std::memcpy(f_reduced + 6 * j, f.data() + 6 * (size_t)order[(size_t)j], 6 * sizeof(double));
        return MBFEA_OK;
    });
}

int mbfea_spmv_with(mbfea_ctx *c, int32_t kernel, const double *x, double *y)
{
    if (int rc = need_job(c)) return rc;
    if (!x || !y) return fail(c, MBFEA_EARG, "NULL");
    if (!c->assembled) return fail(c, MBFEA_ESTATE, "spmv before assemble");
    if (c->world() > 1) return fail(c, MBFEA_ESTATE, "mbfea_spmv is a single-rank tap");
    if (kernel == 2 && !c->tma_ok) return fail(c, MBFEA_EARG, "pattern has a row longer than a TMA stage");
    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        const int64_t nb = c->nb();
        const size_t n = (size_t)6 * nb * sizeof(double);
        // x and y are in canonical reduced numbering; the device works on the internal one
        std::vector<double> xi, yi;
        const double *src = x;
        if (!c->plan.perm.empty()) {
            xi.resize((size_t)6 * nb); yi.resize((size_t)6 * nb);
            for (int64_t i = 0; i < nb; ++i) std::memcpy(xi.data() + 6 * (size_t)c->plan.perm[(size_t)i], x + 6 * i, 6 * sizeof(double));
            src = xi.data();
        }
        CK(cudaMemcpyAsync(c->p.p, src, n, cudaMemcpyHostToDevice, c->st));
        run_spmv_full(c, pick_spmv(c, kernel), c->p.p, c->Ap.p);
        CK(cudaMemcpyAsync(c->plan.perm.empty() ? y : yi.data(), c->Ap.p, n, cudaMemcpyDeviceToHost, c->st));
        CK(cudaStreamSynchronize(c->st));
        if (!c->plan.perm.empty())
            for (int64_t i = 0; i < nb; ++i) std::memcpy(y + 6 * i, yi.data() + 6 * (size_t)c->plan.perm[(size_t)i], 6 * sizeof(double));
        c->solved = c->recovered = false;  // p / Ap were overwritten
        return MBFEA_OK;
    });
}

int mbfea_spmv(mbfea_ctx *c, const double *x, double *y) { return mbfea_spmv_with(c, 0, x, y); }

int mbfea_recover_from(mbfea_ctx *c, const double *U_cm)
{
    if (int rc = need_job(c)) return rc;
    if (!U_cm) return fail(c, MBFEA_EARG, "NULL");
    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        CK(cudaMemcpyAsync(c->U_cm.p, U_cm, (size_t)6 * c->V * sizeof(double), cudaMemcpyHostToDevice, c->st));
        launch_colmajor_to_rowmajor(c->st, c->V, c->U_cm.p, c->U_rm.p);
        launch_element_forces(c->st, c->E, c->rec.p, c->node4.p, c->U_rm.p, c->F6.p);
        CK(cudaStreamSynchronize(c->st));
        c->recovered = true;
        return MBFEA_OK;
    });
}

// ---- measurement helpers -----------------------------------------------------
int mbfea_time_spmv(mbfea_ctx *c, int32_t kernel, int32_t warmup, int32_t reps, double *ms_per_launch)
{
    if (int rc = need_job(c)) return rc;
    if (!ms_per_launch || reps <= 0) return fail(c, MBFEA_EARG, "bad arguments");
    if (!c->assembled) return fail(c, MBFEA_ESTATE, "time_spmv before assemble");
    if (kernel == 2 && !c->tma_ok) return fail(c, MBFEA_EARG, "pattern has a row longer than a TMA stage");
    if (kernel == 2 && c->plan.sym_half)
        return fail(c, MBFEA_EARG, "the TMA kernel streams whole rows: create the context with options.storage = 1 (or spmv_kernel = 2)");
    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        const int k = pick_spmv(c, kernel);
        launch_fill(c->st, 6 * (c->nb() + c->ng()), c->p.p, 0.5, 0.03125, 17);
        for (int i = 0; i < warmup; ++i) run_spmv(c, k, c->p.p, c->Ap.p, c->part_a.p, nullptr, nullptr);
        CK(cudaEventRecord(c->ev[4], c->st));
        for (int i = 0; i < reps; ++i) run_spmv(c, k, c->p.p, c->Ap.p, c->part_a.p, nullptr, nullptr);
        CK(cudaEventRecord(c->ev[5], c->st));
        CK(cudaStreamSynchronize(c->st));
        *ms_per_launch = (double)elapsed(c, 4, 5) / reps;
        c->solved = c->recovered = false;
        return MBFEA_OK;
    });
}

int mbfea_time_assemble(mbfea_ctx *c, int32_t warmup, int32_t reps, double *ms_per_launch)
{
    if (int rc = need_job(c)) return rc;
    if (!ms_per_launch || reps <= 0) return fail(c, MBFEA_EARG, "bad arguments");
    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        for (int i = 0; i < warmup; ++i) do_assemble(c);
        CK(cudaEventRecord(c->ev[4], c->st));
        for (int i = 0; i < reps; ++i) do_assemble(c);
        CK(cudaEventRecord(c->ev[5], c->st));
        CK(cudaStreamSynchronize(c->st));
        *ms_per_launch = (double)elapsed(c, 4, 5) / reps;
        return MBFEA_OK;
    });
}

// stamps[iteration][slot], kTraceSlots per iteration, ns (globaltimer); see kernels.cuh
int mbfea_debug_trace(mbfea_ctx *c, uint64_t *stamps, int64_t max_iters, int64_t *slots_per_iter)
{
    if (!c || !stamps) return MBFEA_EARG;
    if (slots_per_iter) *slots_per_iter = kTraceSlots;
    if (!c->trace.p) return fail(c, MBFEA_ESTATE, "no trace: set MBFEA_TRACE=1 before mbfea_solve");
    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        const int64_t n = std::min<int64_t>(max_iters, kTraceIters);
        CK(cudaMemcpy(stamps, c->trace.p, (size_t)n * kTraceSlots * sizeof(uint64_t), cudaMemcpyDeviceToHost));
        return MBFEA_OK;
    });
}

// ---- host-only partition plan --------------------------------------------------
int mbfea_debug_compare_plan(mbfea_ctx *c, int64_t mismatches[12], int32_t *built_on_device)
{
    if (int rc = need_job(c)) return rc;
    if (!mismatches) return fail(c, MBFEA_EARG, "NULL");
    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        CK(cudaStreamSynchronize(c->st));
        std::vector<int32_t> elems((size_t)2 * (size_t)c->E), fixed((size_t)c->F);
        if (c->E) CK(cudaMemcpy(elems.data(), c->stg_elems.p, elems.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
        if (c->F) CK(cudaMemcpy(fixed.data(), c->fixed.p, fixed.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
        HostPlan H;
        H.perm = c->plan.perm; H.inv = c->plan.inv;
        // same numbering as the job: the renumbering decision is not repeated (it needs the coordinates), its result is reused
        build_plan_begin(c->V, c->E, elems.data(), c->F, fixed.data(), c->opt.rank, c->opt.world, H, 0, nullptr);
        if (!c->plan.perm.empty()) {
            H.perm = c->plan.perm; H.inv = c->plan.inv;
            for (int64_t v = 0; v < c->V; ++v)
                if (H.free_index[(size_t)v] >= 0) H.free_index[(size_t)v] = H.perm[(size_t)H.free_index[(size_t)v]];
        }
        build_plan_pattern(c->E, elems.data(), true, H);
        build_solver_format(H, c->plan.sym_half);
        auto cmp = [&](const DevBuf<int32_t> &d, const hvec<int32_t> &h, size_t n_dev) -> int64_t {
            if (n_dev != h.size()) return -1;
            std::vector<int32_t> t(h.size());
            if (!t.empty()) CK(cudaMemcpy(t.data(), d.p, t.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
            int64_t bad = 0;
            for (size_t i = 0; i < t.size(); ++i) bad += t[i] != h[i];
            return bad;
        };
        const size_t nb = (size_t)c->nb(), nnzb = (size_t)c->plan.nnzb();
        const size_t ntile1 = c->plan.sym_half ? (nb + 31) / 32 + 1 : 0;
        mismatches[0] = cmp(c->rowptr, H.rowptr, nb + 1);
        mismatches[1] = cmp(c->colidx, H.colidx, nnzb);
        mismatches[2] = cmp(c->diag_idx, H.diag_idx, nb);
        mismatches[3] = cmp(c->contrib_ptr, H.contrib_ptr, nnzb + 1);
        mismatches[4] = cmp(c->contrib, H.contrib, H.contrib.size());
        mismatches[5] = cmp(c->uptr, H.uptr, nb + 1);
        mismatches[6] = cmp(c->ucol, H.ucol, (size_t)c->n_stored);
        mismatches[7] = cmp(c->usrc, H.usrc, (size_t)c->n_stored);
        mismatches[8] = cmp(c->lptr, H.lptr, nb + 1);
        mismatches[9] = cmp(c->lcol, H.lcol, (size_t)c->n_refs);
        mismatches[10] = cmp(c->lofs, H.lofs, c->plan.sym_half ? (size_t)c->n_refs : 0);
        mismatches[11] = cmp(c->ilv_base, H.tile_base, ntile1);
        if (built_on_device) *built_on_device = c->device_plan ? (c->plan.uptr.empty() && c->n_stored > 0 ? 2 : 1) : 0;
        return MBFEA_OK;
    });
}

int mbfea_debug_compare_hierarchy(mbfea_ctx *c, int64_t mismatches[256], int32_t *levels, int32_t *built_on_device)
{
    if (int rc = need_job(c)) return rc;
    if (!mismatches || !levels) return fail(c, MBFEA_EARG, "NULL");
    if (!c->use_ml || c->opt.world != 1) return fail(c, MBFEA_ESTATE, "no multilevel hierarchy on this (one-rank) job");
    return guarded(c, [&]() -> int {
        CK(cudaSetDevice(c->device));
        CK(cudaStreamSynchronize(c->st));
        const int64_t V = c->V, nbg = c->plan.nb_global;
        std::vector<double> nodes((size_t)3 * (size_t)V), pos((size_t)3 * (size_t)nbg);
        CK(cudaMemcpy(nodes.data(), c->stg_nodes.p, nodes.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (int64_t v = 0; v < V; ++v) {
            const int32_t r = c->plan.free_index[(size_t)v];
            if (r >= 0) for (int a = 0; a < 3; ++a) pos[(size_t)3 * r + a] = nodes[(size_t)a * V + v];
        }
        MlHierarchy H;
        c->ml.build_hierarchy_host(nbg, c->plan.rowptr.data(), c->plan.colidx.data(), pos.data(), c->ml_mean_len, 1, H);
        std::vector<int64_t> bad;
        c->ml.compare_with_host(c->st, H, (int64_t)c->plan.rowptr[(size_t)nbg], bad);
        std::memset(mismatches, 0, 256 * sizeof(int64_t));
        for (size_t i = 0; i < bad.size() && i < 256; ++i) mismatches[i] = bad[i];
        *levels = (int32_t)std::min<size_t>(16, bad.size() / 16);
        if (built_on_device) *built_on_device = c->ml_on_device ? 1 : 0;
        return MBFEA_OK;
    });
}

int mbfea_plan_partition(int64_t V, int64_t E, const int32_t *elems_cm, int64_t F, const int32_t *fixed,
                         int32_t rank, int32_t world, mbfea_plan_sizes *sizes, int64_t *rowptr, int32_t *colidx,
                         int64_t *ghost_global, int32_t *ghost_owner, int64_t *send_local_rows, int32_t *send_dest)
{
    if (V <= 0 || E <= 0 || !elems_cm || (F > 0 && !fixed) || world < 1 || rank < 0 || rank >= world) return MBFEA_EARG;
    for (int64_t e = 0; e < 2 * E; ++e)
        if (elems_cm[e] < 0 || elems_cm[e] >= V) return MBFEA_EPRECOND;
    for (int64_t i = 0; i < F; ++i)
        if (fixed[i] < 0 || fixed[i] >= V) return MBFEA_ERANGE_FIXED;
    HostPlan P;
    try {
        build_plan(V, E, elems_cm, F, fixed, rank, world, false, P);
    } catch (const std::bad_alloc &) {
        return MBFEA_EARG;
    }
    if (sizes) {
        sizes->nb_global = P.nb_global; sizes->row_begin = P.row_begin; sizes->row_end = P.row_end;
        sizes->nnz_blocks = P.nnzb(); sizes->n_ghost = (int64_t)P.ghost_global.size();
        sizes->n_send = (int64_t)P.send_rows.size(); sizes->n_neighbors = (int64_t)P.neighbors.size();
        find_components(P, 1024, 32);
        sizes->n_components = P.comp_ptr.empty() ? 0 : (int64_t)P.comp_ptr.size() - 1;
        sizes->max_component_rows = P.comp_max_rows;
    }
    if (rowptr) for (size_t i = 0; i < P.rowptr.size(); ++i) rowptr[i] = P.rowptr[i];
    if (colidx) std::copy(P.colidx.begin(), P.colidx.end(), colidx);
    if (ghost_global) std::copy(P.ghost_global.begin(), P.ghost_global.end(), ghost_global);
    if (ghost_owner) std::copy(P.ghost_owner.begin(), P.ghost_owner.end(), ghost_owner);
    if (send_local_rows) for (size_t i = 0; i < P.send_rows.size(); ++i) send_local_rows[i] = P.send_rows[i];
    if (send_dest) std::copy(P.send_dest.begin(), P.send_dest.end(), send_dest);
    return MBFEA_OK;
}

int mbfea_host_contributors(int64_t V, int64_t E, const int32_t *elems_cm, int64_t F, const int32_t *fixed,
                            int32_t rank, int32_t world, int64_t sizes[2], int32_t *contrib_ptr, int32_t *contrib)
{
    if (V <= 0 || E <= 0 || !elems_cm || !sizes || (F > 0 && !fixed) || world < 1 || rank < 0 || rank >= world) return MBFEA_EARG;
    for (int64_t e = 0; e < 2 * E; ++e)
        if (elems_cm[e] < 0 || elems_cm[e] >= V) return MBFEA_EPRECOND;
    for (int64_t i = 0; i < F; ++i)
        if (fixed[i] < 0 || fixed[i] >= V) return MBFEA_ERANGE_FIXED;
    HostPlan P;
    try {
        build_plan(V, E, elems_cm, F, fixed, rank, world, true, P);
    } catch (const std::bad_alloc &) {
        return MBFEA_EARG;
    }
    sizes[0] = P.nnzb(); sizes[1] = (int64_t)P.contrib.size();
    if (contrib_ptr) std::copy(P.contrib_ptr.begin(), P.contrib_ptr.end(), contrib_ptr);
    if (contrib) std::copy(P.contrib.begin(), P.contrib.end(), contrib);
    return MBFEA_OK;
}

// ---- scenario files ------------------------------------------------------------
int mbfea_load_scenario(mbfea_ctx *c, const char *config_json_path, const char *ply_path_override)
{
    if (!c || !config_json_path) return MBFEA_EARG;
    Scenario sc;
    std::string msg;
    try {
        if (int rc = load_scenario_files(config_json_path, ply_path_override, sc, msg)) return fail(c, rc, msg);
    } catch (const std::exception &e) {   // e.g. bad_alloc: nothing may unwind through the C ABI
        return fail(c, MBFEA_EIO, std::string("scenario reader: ") + e.what());
    }
    return mbfea_set_job(c, sc.V, sc.E, sc.nodes_cm.data(), sc.elems_cm.data(), sc.normals_cm.data(),
                         sc.dims_bh.data(), sc.mat_nuE.data(), (int64_t)sc.fixed.size(), sc.fixed.data(),
                         (int64_t)sc.f_node.size(), sc.f_node.data(), sc.f_dof.data(), sc.f_mag.data());
}

struct mbfea_scenario { Scenario sc; };

int mbfea_scenario_open(const char *config_json_path, const char *ply_path_override, mbfea_scenario **out,
                        char *errbuf, int64_t errbuf_len)
{
    if (!config_json_path || !out) return MBFEA_EARG;
    *out = nullptr;
    std::unique_ptr<mbfea_scenario> h(new (std::nothrow) mbfea_scenario);
    if (!h) return MBFEA_EARG;
    std::string msg;
    int rc;
    try {
        rc = load_scenario_files(config_json_path, ply_path_override, h->sc, msg);
    } catch (const std::exception &e) {   // e.g. bad_alloc: nothing may unwind through the C ABI
        rc = MBFEA_EIO; msg = std::string("scenario reader: ") + e.what();
    }
    if (errbuf && errbuf_len > 0) std::snprintf(errbuf, (size_t)errbuf_len, "%s", msg.c_str());
    if (rc != MBFEA_OK) return rc;
    *out = h.release();
    return MBFEA_OK;
}

void mbfea_scenario_sizes(const mbfea_scenario *h, int64_t *V, int64_t *E, int64_t *F, int64_t *nF)
{
    if (!h) return;
    if (V) *V = h->sc.V;
    if (E) *E = h->sc.E;
    if (F) *F = (int64_t)h->sc.fixed.size();
    if (nF) *nF = (int64_t)h->sc.f_node.size();
}

void mbfea_scenario_copy(const mbfea_scenario *h, double *nodes_cm, int32_t *elems_cm, double *normals_cm,
                         float *dims_bh, float *mat_nuE, int32_t *fixed, int64_t *f_node, int64_t *f_dof,
                         double *f_mag)
{
    if (!h) return;
    const Scenario &s = h->sc;
    if (nodes_cm) std::copy(s.nodes_cm.begin(), s.nodes_cm.end(), nodes_cm);
    if (elems_cm) std::copy(s.elems_cm.begin(), s.elems_cm.end(), elems_cm);
    if (normals_cm) std::copy(s.normals_cm.begin(), s.normals_cm.end(), normals_cm);
    if (dims_bh) std::copy(s.dims_bh.begin(), s.dims_bh.end(), dims_bh);
    if (mat_nuE) std::copy(s.mat_nuE.begin(), s.mat_nuE.end(), mat_nuE);
    if (fixed) std::copy(s.fixed.begin(), s.fixed.end(), fixed);
    if (f_node) std::copy(s.f_node.begin(), s.f_node.end(), f_node);
    if (f_dof) std::copy(s.f_dof.begin(), s.f_dof.end(), f_dof);
    if (f_mag) std::copy(s.f_mag.begin(), s.f_mag.end(), f_mag);
}

void mbfea_scenario_close(mbfea_scenario *h) { delete h; }

// ---- host-only pieces ----------------------------------------------------------
int mbfea_host_validate(int64_t V, int64_t E, const int32_t *elems_cm, const float *dims_bh, const float *mat_nuE,
                        int64_t F, const int32_t *fixed, int64_t nF, const int64_t *f_node, const int64_t *f_dof,
                        char *msg, int64_t msg_len)
{
    std::string m;
    const int rc = validate_job(V, E, elems_cm, dims_bh, mat_nuE, F, fixed, nF, f_node, f_dof, m);
    if (msg && msg_len > 0) std::snprintf(msg, (size_t)msg_len, "%s", m.c_str());
    return rc;
}

int mbfea_host_props(int64_t E, const float *dims_bh, const float *mat_nuE, double *props4)
{
    if (E < 0 || (E > 0 && (!dims_bh || !mat_nuE || !props4))) return MBFEA_EARG;
    std::vector<float> pf((size_t)4 * (size_t)E);
    derive_props(E, dims_bh, mat_nuE, pf.data());
    for (size_t i = 0; i < pf.size(); ++i) props4[i] = (double)pf[i];
    return MBFEA_OK;
}

int mbfea_host_free_map(int64_t V, int64_t F, const int32_t *fixed, int32_t *free_index, int64_t *n_free)
{
    if (V <= 0 || !free_index || (F > 0 && !fixed)) return MBFEA_EARG;
    for (int64_t i = 0; i < F; ++i)
        if (fixed[i] < 0 || fixed[i] > V - 1) return MBFEA_ERANGE_FIXED;
    std::vector<int32_t> fm;
    const int64_t n = build_free_map(V, F, fixed, fm);
    std::copy(fm.begin(), fm.end(), free_index);
    if (n_free) *n_free = n;
    return MBFEA_OK;
}

int mbfea_host_renumber(int64_t V, int64_t E, const int32_t *elems_cm, int64_t F, const int32_t *fixed, int32_t mode,
                        int32_t *perm, int32_t *reordered)
{
    if (V <= 0 || E <= 0 || !elems_cm || !perm || (F > 0 && !fixed)) return MBFEA_EARG;
    for (int64_t e = 0; e < 2 * E; ++e)
        if (elems_cm[e] < 0 || elems_cm[e] >= V) return MBFEA_EPRECOND;
    for (int64_t i = 0; i < F; ++i)
        if (fixed[i] < 0 || fixed[i] >= V) return MBFEA_ERANGE_FIXED;
    HostPlan P;
    try {
        build_plan(V, E, elems_cm, F, fixed, 0, 1, false, P, mode == 1 ? 0 : (mode == 2 ? 2 : (mode == 3 ? 3 : 1)));
    } catch (const std::bad_alloc &) {
        return MBFEA_EARG;
    }
    if (reordered) *reordered = P.perm.empty() ? 0 : 1;
    for (int64_t i = 0; i < P.nb_global; ++i) perm[i] = P.perm.empty() ? (int32_t)i : P.perm[(size_t)i];
    return MBFEA_OK;
}

int mbfea_host_solver_format(int64_t V, int64_t E, const double *nodes_cm, const int32_t *elems_cm, int64_t F,
                             const int32_t *fixed, int32_t mode, int32_t rank, int32_t world, int64_t sizes[8],
                             int32_t *uptr, int32_t *ucol, int32_t *lptr, int32_t *lcol, int32_t *lofs,
                             int32_t *tile_base, int32_t *perm)
{
    if (V <= 0 || E <= 0 || !elems_cm || !sizes || (F > 0 && !fixed) || world < 1 || rank < 0 || rank >= world) return MBFEA_EARG;
    for (int64_t e = 0; e < 2 * E; ++e)
        if (elems_cm[e] < 0 || elems_cm[e] >= V) return MBFEA_EPRECOND;
    for (int64_t i = 0; i < F; ++i)
        if (fixed[i] < 0 || fixed[i] >= V) return MBFEA_ERANGE_FIXED;
    HostPlan P;
    try {
        build_plan(V, E, elems_cm, F, fixed, rank, world, false, P, mode == 1 ? 0 : (mode == 2 ? 2 : (mode == 3 ? 3 : 1)), nodes_cm);
        build_solver_format(P, true);
    } catch (const std::bad_alloc &) {
        return MBFEA_EARG;
    }
    const int64_t nb = P.nb();
    int32_t longest = 0;
    for (int64_t r = 0; r < nb; ++r) longest = std::max(longest, P.uptr[(size_t)r + 1] - P.uptr[(size_t)r]);
    sizes[0] = nb; sizes[1] = P.nnzb() - nb; sizes[2] = (int64_t)P.ucol.size(); sizes[3] = (int64_t)P.lcol.size();
    sizes[4] = P.tile_base.empty() ? 0 : P.tile_base.back(); sizes[5] = P.balanced ? 1 : 0; sizes[6] = longest;
    sizes[7] = P.perm.empty() ? 0 : 1;
    if (!P.sym_half) return MBFEA_EARG;   // value array beyond 2^31 units: full rows (not reachable at host-test sizes)
    auto put = [](int32_t *dst, const auto &src) { if (dst && !src.empty()) std::memcpy(dst, src.data(), src.size() * sizeof(int32_t)); };
    put(uptr, P.uptr); put(ucol, P.ucol); put(lptr, P.lptr); put(lcol, P.lcol); put(lofs, P.lofs); put(tile_base, P.tile_base);
    if (perm) for (int64_t i = 0; i < P.nb_global; ++i) perm[i] = P.perm.empty() ? (int32_t)i : P.perm[(size_t)i];
    return MBFEA_OK;
}

int mbfea_host_ml_hierarchy(int64_t V, int64_t E, const double *nodes_cm, const int32_t *elems_cm, int64_t F,
                            const int32_t *fixed, double cell_beam_lengths, double ratio, int64_t coarsest_rows,
                            int32_t smooth_from, int32_t max_levels, int32_t *n_levels, int64_t *sizes)
{
    if (V <= 0 || E <= 0 || !nodes_cm || !elems_cm || !n_levels || (F > 0 && !fixed) || max_levels < 1) return MBFEA_EARG;
    for (int64_t e = 0; e < 2 * E; ++e)
        if (elems_cm[e] < 0 || elems_cm[e] >= V) return MBFEA_EPRECOND;
    for (int64_t i = 0; i < F; ++i)
        if (fixed[i] < 0 || fixed[i] >= V) return MBFEA_ERANGE_FIXED;
    try {
        HostPlan P;
        build_plan(V, E, elems_cm, F, fixed, 0, 1, false, P);
        const int64_t nb = P.nb();
        std::vector<double> pos((size_t)3 * std::max<int64_t>(nb, 1));
        for (int64_t v = 0; v < V; ++v)
            if (P.free_index[(size_t)v] >= 0)
                for (int a = 0; a < 3; ++a) pos[(size_t)3 * P.free_index[(size_t)v] + a] = nodes_cm[(size_t)a * V + v];
        double sum = 0.0;
        for (int64_t e = 0; e < E; ++e) {
            double d2 = 0.0;
            for (int a = 0; a < 3; ++a) {
                const double d = nodes_cm[(size_t)a * V + elems_cm[e]] - nodes_cm[(size_t)a * V + elems_cm[E + e]];
                d2 += d * d;
            }
            sum += std::sqrt(d2);
        }
        MlParams mp;
        if (cell_beam_lengths > 0.0) mp.cell = cell_beam_lengths;
        if (ratio > 1.0) mp.ratio = ratio;
        if (coarsest_rows > 0) mp.coarsest_rows = coarsest_rows;
        if (smooth_from > 0) mp.smooth_from = smooth_from;
        MlHierarchy H;
        build_ml_hierarchy(nb, P.rowptr.data(), P.colidx.data(), pos.data(), mp.cell * sum / (double)E, mp.ratio, mp.max_levels,
                           mp.coarsest_rows, mp.smooth_from, H);
        *n_levels = (int32_t)H.levels();
        for (size_t l = 0; l < H.levels() && (int32_t)l < max_levels && sizes; ++l) {
            const bool gen = l >= 1;
            sizes[6 * l + 0] = H.agg[l].n_fine;
            sizes[6 * l + 1] = H.agg[l].n_coarse;
            sizes[6 * l + 2] = gen ? (int64_t)H.xfer[l].ccol.size() : (int64_t)H.agg[l].colidx.size();   // blocks of the next level's matrix
            sizes[6 * l + 3] = gen ? (int64_t)H.xfer[l].pcol.size() : H.agg[l].n_fine;                    // blocks of P
            sizes[6 * l + 4] = gen ? (int64_t)H.xfer[l].apcol.size() : 0;                                 // blocks of A P
            sizes[6 * l + 5] = gen && H.xfer[l].smoothed ? 1 : 0;
        }
    } catch (const std::bad_alloc &) {
        return MBFEA_EARG;
    }
    return MBFEA_OK;
}

int mbfea_host_ml_image(int64_t V, int64_t E, const double *nodes_cm, const int32_t *elems_cm, int64_t F,
                        const int32_t *fixed, int32_t world, int32_t smooth_from, int64_t *n_bytes, unsigned char *image)
{
    if (V <= 0 || E <= 0 || !nodes_cm || !elems_cm || !n_bytes || (F > 0 && !fixed) || world < 1 || world > kMlImageMaxWorld) return MBFEA_EARG;
    for (int64_t e = 0; e < 2 * E; ++e)
        if (elems_cm[e] < 0 || elems_cm[e] >= V) return MBFEA_EPRECOND;
    for (int64_t i = 0; i < F; ++i)
        if (fixed[i] < 0 || fixed[i] >= V) return MBFEA_ERANGE_FIXED;
    try {
        HostPlan P;
        build_plan(V, E, elems_cm, F, fixed, 0, 1, false, P);
        const int64_t nb = P.nb();
        std::vector<double> pos((size_t)3 * std::max<int64_t>(nb, 1));
        for (int64_t v = 0; v < V; ++v)
            if (P.free_index[(size_t)v] >= 0)
                for (int a = 0; a < 3; ++a) pos[(size_t)3 * P.free_index[(size_t)v] + a] = nodes_cm[(size_t)a * V + v];
        double sum = 0.0;
        for (int64_t e = 0; e < E; ++e) {
            double d2 = 0.0;
            for (int a = 0; a < 3; ++a) {
                const double d = nodes_cm[(size_t)a * V + elems_cm[e]] - nodes_cm[(size_t)a * V + elems_cm[E + e]];
                d2 += d * d;
            }
            sum += std::sqrt(d2);
        }
        MlPrec ml;
        if (smooth_from > 0) ml.par.smooth_from = smooth_from;
        MlHierarchy H;
        ml.build_hierarchy_host(nb, P.rowptr.data(), P.colidx.data(), pos.data(), sum / (double)E, world, H);
        std::vector<unsigned char> img;
        if (!ml_image_pack(H, world, nb, P.rowptr.data(), img)) return MBFEA_ESTATE;
        if (image && *n_bytes >= (int64_t)img.size()) std::memcpy(image, img.data(), img.size());
        *n_bytes = (int64_t)img.size();
    } catch (const std::bad_alloc &) {
        return MBFEA_EARG;
    }
    return MBFEA_OK;
}

int mbfea_host_spd_inverse(int64_t n, const double *A, double *Ainv)
{
    if (n <= 0 || !A || !Ainv) return MBFEA_EARG;
    try {
        return dense_spd_inverse(n, A, Ainv) == 0 ? MBFEA_OK : MBFEA_ESINGULAR;
    } catch (const std::bad_alloc &) {
        return MBFEA_EARG;
    }
}

struct mbfea_hierarchy { Hierarchy H; };

int mbfea_hierarchy_open(int64_t V, int64_t E, const double *nodes_cm, const int32_t *elems_cm, int64_t F,
                         const int32_t *fixed, double cell0, double ratio, int32_t max_levels,
                         int64_t coarsest_rows, mbfea_hierarchy **out)
{
    if (V <= 0 || E <= 0 || !nodes_cm || !elems_cm || !out || (F > 0 && !fixed) || max_levels < 1 || !(ratio > 1.0)) return MBFEA_EARG;
    *out = nullptr;
    for (int64_t e = 0; e < 2 * E; ++e)
        if (elems_cm[e] < 0 || elems_cm[e] >= V) return MBFEA_EPRECOND;
    for (int64_t i = 0; i < F; ++i)
        if (fixed[i] < 0 || fixed[i] >= V) return MBFEA_ERANGE_FIXED;
    try {
        HostPlan P;
        build_plan(V, E, elems_cm, F, fixed, 0, 1, false, P);
        const int64_t nb = P.nb();
        std::vector<double> pos((size_t)3 * std::max<int64_t>(nb, 1));
        for (int64_t v = 0; v < V; ++v)
            if (P.free_index[(size_t)v] >= 0)
                for (int a = 0; a < 3; ++a) pos[(size_t)3 * P.free_index[(size_t)v] + a] = nodes_cm[(size_t)a * V + v];
        if (!(cell0 > 0.0)) {   // 4 x the mean beam length
            double sum = 0.0;
            for (int64_t e = 0; e < E; ++e) {
                double d2 = 0.0;
                for (int a = 0; a < 3; ++a) {
                    const double d = nodes_cm[(size_t)a * V + elems_cm[e]] - nodes_cm[(size_t)a * V + elems_cm[E + e]];
                    d2 += d * d;
                }
                sum += std::sqrt(d2);
            }
            cell0 = 4.0 * sum / (double)E;
        }
        std::unique_ptr<mbfea_hierarchy> h(new mbfea_hierarchy);
        if (cell0 > 0.0)
            build_hierarchy(nb, P.rowptr.data(), P.colidx.data(), pos.data(), cell0, ratio, max_levels, coarsest_rows, h->H);
        *out = h.release();
    } catch (const std::bad_alloc &) {
        return MBFEA_EARG;
    }
    return MBFEA_OK;
}

int mbfea_hierarchy_levels(const mbfea_hierarchy *h) { return h ? (int)h->H.levels.size() : 0; }

int mbfea_hierarchy_sizes(const mbfea_hierarchy *h, int32_t level, int64_t sizes[6])
{
    if (!h || !sizes || level < 0 || level >= (int32_t)h->H.levels.size()) return MBFEA_EARG;
    const HierLevel &L = h->H.levels[(size_t)level];
    sizes[0] = L.n_fine; sizes[1] = L.n_coarse; sizes[2] = (int64_t)L.colidx.size(); sizes[3] = (int64_t)L.gal_src.size();
    sizes[4] = sizes[5] = 0;
    return MBFEA_OK;
}

int mbfea_hierarchy_copy(const mbfea_hierarchy *h, int32_t level, int32_t *agg, double *off, int32_t *rowptr,
                         int32_t *colidx, int32_t *gal_ptr, int32_t *gal_src)
{
    if (!h || level < 0 || level >= (int32_t)h->H.levels.size()) return MBFEA_EARG;
    const HierLevel &L = h->H.levels[(size_t)level];
    if (agg) std::copy(L.agg.begin(), L.agg.end(), agg);
    if (off) std::copy(L.off.begin(), L.off.end(), off);
    if (rowptr) std::copy(L.rowptr.begin(), L.rowptr.end(), rowptr);
    if (colidx) std::copy(L.colidx.begin(), L.colidx.end(), colidx);
    if (gal_ptr) std::copy(L.gal_ptr.begin(), L.gal_ptr.end(), gal_ptr);
    if (gal_src) std::copy(L.gal_src.begin(), L.gal_src.end(), gal_src);
    return MBFEA_OK;
}

void mbfea_hierarchy_close(mbfea_hierarchy *h) { delete h; }

}  // extern "C"
