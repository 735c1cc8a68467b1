// Host-side preparation: validation in the reference's order, the float
// section-property quirk, the constraint map, the 6x6-block CSR pattern with
// per-block contributor lists (ascending element order => deterministic,
// scatter-free assembly on the device) and the vertex-block partition plan.
// Pure C++17: nothing here touches CUDA.
#include <algorithm>
#include <chrono>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <memory>
#include <thread>

#include "mbfea_internal.hpp"

namespace mbfea {

// ---------------------------------------------------------------------------
// Validation.  Order follows FeaSimulationJob(job): setElements() first, then
// setNodes() -> setFixedNodes() -> setNodalForces()
// [REF src/beamsimulator.cpp:107-111].
// ---------------------------------------------------------------------------
template <class Fn>
static void parallel_ranges(int64_t n, int64_t min_per_thread, Fn fn);   // (defined below)

int validate_forces(int64_t V, int64_t nF, const int64_t *f_node, const int64_t *f_dof,
                    std::string &msg)
{
    for (int64_t i = 0; i < nF; ++i) {
        // `force.index > nodes.size() - 1` compares a ptrdiff_t with a size_t, so a
        // negative index wraps and throws as well [REF src/beamsimulator.cpp:164-166]
        if (f_node[i] < 0 || f_node[i] > V - 1) {
            msg = "Node index is out of valid range.";
            return MBFEA_ERANGE_FORCE;
        }
    }
    for (int64_t i = 0; i < nF; ++i) {
        // the scenario loader Expects(dof in [0,6)) [REF src/utilities.hpp:77]
        if (f_dof[i] < 0 || f_dof[i] > 5) {
            msg = "precondition violated: force dof must be in [0,6)";
            return MBFEA_EPRECOND;
        }
    }
    return MBFEA_OK;
}

int validate_job(int64_t V, int64_t E, const int32_t *elems_cm, const float *dims_bh,
                 const float *mat_nuE, int64_t F, const int32_t *fixed, int64_t nF,
                 const int64_t *f_node, const int64_t *f_dof, std::string &msg)
{
    // Expects(!simulationJob.isEmpty()) [REF src/beamsimulator.cpp:41,113-115]
    if (V <= 0 || E <= 0) {
        msg = "precondition violated: empty simulation job (no nodes or no elements)";
        return MBFEA_EPRECOND;
    }
    if (V > (int64_t)INT32_MAX / 8 || E > (int64_t)(1u << 29) - 1) {
        msg = "job too large for 32-bit block indices";
        return MBFEA_EARG;
    }
    // per beam, in the reference's order of checks; chunks are scanned in parallel and the FIRST offending beam decides
    // the message, as in a serial scan
    //   1 the reference's range check [REF :81-83] compares against 3V and lets negatives through; anything it would
    //     let through out of [0,V) is UB there, so this build rejects it as the same class of error
    //   3, 4 BeamDimensions / BeamMaterial constructors [REF src/beam.hpp:14,25]
    auto check = [=](int64_t e) -> int {
        const int32_t a = elems_cm[e], b = elems_cm[E + e];
        if (a < 0 || b < 0 || a >= V || b >= V) return 1;
        if (a == b) return 2;
        const float bb = dims_bh[2 * e], hh = dims_bh[2 * e + 1];
        const float nu = mat_nuE[2 * e];
        if (!(bb > 0 && hh > 0)) return 3;
        if (!(nu <= 0.5f && nu >= -1.0f)) return 4;
        return 0;
    };
    int64_t first_bad[32];
    for (auto &x : first_bad) x = -1;
    int64_t *fbp = first_bad;
    parallel_ranges(E, 1 << 16, [=](int64_t lo, int64_t hi, int t) {
        for (int64_t e = lo; e < hi; ++e)
            if (check(e) != 0) { fbp[t] = e; return; }
    });
    int64_t bad = -1;
    for (int64_t x : first_bad) if (x >= 0 && (bad < 0 || x < bad)) bad = x;
    if (bad >= 0) {
        static const char *const what[] = {"", "precondition violated: element node index out of range",
                                           "precondition violated: zero-length beam (both ends on one vertex)",
                                           "precondition violated: beam dimensions must be positive",
                                           "precondition violated: Poisson's ratio outside [-1, 0.5]"};
        msg = what[check(bad)];
        return MBFEA_EPRECOND;
    }
    for (int64_t i = 0; i < F; ++i) {
        if (fixed[i] < 0 || (int64_t)fixed[i] > V - 1) {
            msg = "Vertex index is outside of valid range.";  // [REF :134-137]
            return MBFEA_ERANGE_FIXED;
        }
    }
    return validate_forces(V, nF, f_node, f_dof, msg);
}

// ---------------------------------------------------------------------------
// Section properties.  Every member of BeamSimulationProperties is `float`
// [REF src/beamsimulator.hpp:13-27]; the expressions below keep the reference's
// promotions: std::pow(float,int) -> double, float*double -> double, result
// narrowed to float on assignment [REF src/beamsimulator.cpp:173-188].
// ---------------------------------------------------------------------------
static inline void props_one(float b, float h, float nu, float Egpa, float *out)
{
    const float crossArea = b * h;
    const float I2 = static_cast<float>(b * std::pow(static_cast<double>(h), 3) / 12);
    const float I3 = static_cast<float>(h * std::pow(static_cast<double>(b), 3) / 12);
    const float polarInertia = I2 + I3;
    const float youngsModulusPascal = static_cast<float>(Egpa * std::pow(10.0, 9));
    const float G = youngsModulusPascal / (2 * (1 + nu));
    const float EA = youngsModulusPascal * crossArea;
    const float EIy = youngsModulusPascal * I3;
    const float EIz = youngsModulusPascal * I2;
    const float GJ = G * polarInertia;
    out[0] = EA;   // fea::Props(EA, EIz, EIy, GJ, normal) [REF :89-90]
    out[1] = EIz;
    out[2] = EIy;
    out[3] = GJ;
}

// Host threads of this process: the node's cores divided by the ranks that share it (one process per GPU, all on
// one node: 8 ranks x 16 threads on 16 cores made set_job slower on 8 GPUs than on one).  MBFEA_HOST_THREADS overrides.
static thread_local int tl_rank_share = 1;
struct RankShare {
    int saved;
    explicit RankShare(int world) : saved(tl_rank_share) { tl_rank_share = std::max(1, world); }
    ~RankShare() { tl_rank_share = saved; }
};
// the calling thread's passes share the node's cores with `world` ranks from here on (set_job; 1 = all cores)
void host_threads_share(int world) { tl_rank_share = std::max(1, world); }
static int64_t host_threads()
{
    static const int forced = [] { const char *e = std::getenv("MBFEA_HOST_THREADS"); return e ? std::atoi(e) : 0; }();
    if (forced > 0) return std::min(forced, 32);
    const unsigned hw = std::thread::hardware_concurrency();
    return std::max<int64_t>(1, std::min<int64_t>(32, (int64_t)(hw ? hw : 1) / tl_rank_share));
}

template <class Fn>
static void parallel_ranges(int64_t n, int64_t min_per_thread, Fn fn)
{
    const int64_t nt = std::max<int64_t>(1, std::min<int64_t>(host_threads(), n / std::max<int64_t>(1, min_per_thread)));
    if (nt <= 1) { fn(0, n, 0); return; }
    std::vector<std::thread> th;
    for (int64_t t = 0; t < nt; ++t)
        th.emplace_back([=] { fn(n * t / nt, n * (t + 1) / nt, (int)t); });
    for (auto &x : th) x.join();
}

void derive_props(int64_t E, const float *dims_bh, const float *mat_nuE, float *props4)
{
    parallel_ranges(E, 1 << 16, [=](int64_t lo, int64_t hi, int) {
        // most meshes repeat one section: memoise the last one
        float lb = NAN, lh = NAN, ln = NAN, le = NAN;
        float last[4] = {0, 0, 0, 0};
        for (int64_t e = lo; e < hi; ++e) {
            const float b = dims_bh[2 * e], h = dims_bh[2 * e + 1], nu = mat_nuE[2 * e], Eg = mat_nuE[2 * e + 1];
            if (!(b == lb && h == lh && nu == ln && Eg == le)) {
                props_one(b, h, nu, Eg, last);
                lb = b; lh = h; ln = nu; le = Eg;
            }
            std::memcpy(props4 + 4 * e, last, sizeof last);
        }
    });
}

// ---------------------------------------------------------------------------
// K3 (host copy): all six DOFs of a fixed vertex are removed
// [REF src/beamsimulator.cpp:139-150], so elimination is a vertex renumbering.
// ---------------------------------------------------------------------------
int64_t build_free_map(int64_t V, int64_t F, const int32_t *fixed, std::vector<int32_t> &fm)
{
    fm.resize((size_t)V);
    int32_t *m = fm.data();
    const int64_t nt = std::max<int64_t>(1, std::min<int64_t>(host_threads(), V >> 16));
    if (nt <= 1) {
        std::fill(m, m + V, 0);
        for (int64_t i = 0; i < F; ++i) m[fixed[i]] = -1;
        int32_t k = 0;
        for (int64_t v = 0; v < V; ++v)
            if (m[v] == 0) m[v] = k++;
        return k;
    }
    // chunked: clear, mark (duplicates store the same value), count the free vertices per chunk, rank them from the
    // chunk's prefix -- the same map as the serial scan above
    parallel_ranges(V, 1 << 16, [=](int64_t lo, int64_t hi, int) { std::fill(m + lo, m + hi, 0); });
    if (F >= (1 << 18)) parallel_ranges(F, 1 << 16, [=](int64_t lo, int64_t hi, int) { for (int64_t i = lo; i < hi; ++i) __atomic_store_n(&m[fixed[i]], -1, __ATOMIC_RELAXED); });
    else for (int64_t i = 0; i < F; ++i) m[fixed[i]] = -1;
    std::vector<int64_t> first((size_t)nt + 1, 0);
    int64_t *fp = first.data();
    parallel_ranges(V, 1 << 16, [=](int64_t lo, int64_t hi, int t) {
        int64_t k = 0;
        for (int64_t v = lo; v < hi; ++v) k += m[v] == 0;
        fp[t + 1] = k;
    });
    for (int64_t t = 0; t < nt; ++t) first[(size_t)t + 1] += first[(size_t)t];
    parallel_ranges(V, 1 << 16, [=](int64_t lo, int64_t hi, int t) {
        int32_t k = (int32_t)fp[t];
        for (int64_t v = lo; v < hi; ++v)
            if (m[v] == 0) m[v] = k++;
    });
    return first[(size_t)nt];
}

static inline int32_t owner_of(int64_t g, int64_t nb_global, int32_t world)
{
    // inverse of row_begin(r) = nb_global*r/world
    int32_t r = (int32_t)(((g + 1) * world - 1) / std::max<int64_t>(nb_global, 1));
    if (r >= world) r = world - 1;
    while (r > 0 && nb_global * r / world > g) --r;
    while (r + 1 < world && nb_global * (r + 1) / world <= g) ++r;
    return r;
}

static double wall_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}
struct PrepTrace {
    bool on = std::getenv("MBFEA_PREP_TRACE") != nullptr;
    double t = wall_ms();
    void lap(const char *what)
    {
        if (!on) return;
        const double n = wall_ms();
        std::fprintf(stderr, "[mbfea prep] %-28s %8.2f ms\n", what, n - t);
        t = n;
    }
};

// Breadth-first renumbering of the free vertices (Cuthill-McKee order without the degree sort):
// neighbours end up within one BFS level of each other, so a row's transposed references and its
// x-gathers land in recently streamed (L2-resident) memory whatever the input vertex ids were.
// Components are exhausted one after the other, so a batch of independent meshes stays a sequence
// of contiguous row ranges.  Returns false (perm left empty) when the canonical order is kept.
// Morton (Z-order) renumbering from the vertex coordinates: consecutive 32 rows form a compact brick
// of the mesh, so both rows of most pairs share an SpMV tile (the pair's second use hits L1) while
// the curve keeps the remaining neighbours within L2 reach.  10 bits per axis, LSD radix sort.
static void morton_renumber(int64_t V, const double *nodes_cm, const std::vector<int32_t> &fc, int64_t nbg,
                            std::vector<int32_t> &perm, std::vector<int32_t> &inv)
{
    double lo[3], hi[3];
    for (int a = 0; a < 3; ++a) { lo[a] = INFINITY; hi[a] = -INFINITY; }
    for (int64_t v = 0; v < V; ++v)
        if (fc[v] >= 0)
            for (int a = 0; a < 3; ++a) {
                const double x = nodes_cm[(size_t)a * V + v];
                lo[a] = std::min(lo[a], x); hi[a] = std::max(hi[a], x);
            }
    const double span = std::max(std::max(hi[0] - lo[0], hi[1] - lo[1]), std::max(hi[2] - lo[2], 1e-300));
    auto spread = [](uint32_t x) {   // 10 bits -> every third bit
        x &= 0x3ffu;
        x = (x | (x << 16)) & 0x030000ffu;
        x = (x | (x << 8)) & 0x0300f00fu;
        x = (x | (x << 4)) & 0x030c30c3u;
        x = (x | (x << 2)) & 0x09249249u;
        return x;
    };
    std::vector<uint32_t> key((size_t)nbg), key2((size_t)nbg);
    std::vector<int32_t> id((size_t)nbg), id2((size_t)nbg);
    for (int64_t v = 0; v < V; ++v) {
        const int32_t c = fc[v];
        if (c < 0) continue;
        uint32_t k = 0;
        for (int a = 0; a < 3; ++a) {
            const double t = (nodes_cm[(size_t)a * V + v] - lo[a]) / span;
            const uint32_t q = (uint32_t)std::min(1023.0, std::max(0.0, t * 1023.0 + 0.5));
            k |= spread(q) << a;
        }
        key[(size_t)c] = k; id[(size_t)c] = c;
    }
    for (int pass = 0; pass < 4; ++pass) {   // stable LSD radix sort, 8 bits per pass
        const int sh = 8 * pass;
        size_t cnt[257] = {0};
        for (int64_t i = 0; i < nbg; ++i) cnt[((key[(size_t)i] >> sh) & 0xff) + 1]++;
        for (int b = 0; b < 256; ++b) cnt[b + 1] += cnt[b];
        for (int64_t i = 0; i < nbg; ++i) {
            const size_t d = cnt[(key[(size_t)i] >> sh) & 0xff]++;
            key2[d] = key[(size_t)i]; id2[d] = id[(size_t)i];
        }
        key.swap(key2); id.swap(id2);
    }
    perm.resize((size_t)nbg); inv.resize((size_t)nbg);
    for (int64_t i = 0; i < nbg; ++i) { inv[(size_t)i] = id[(size_t)i]; perm[(size_t)id[(size_t)i]] = (int32_t)i; }
}

static bool bfs_renumber(int64_t V, int64_t E, const int32_t *elems_cm, const std::vector<int32_t> &fc, int64_t nbg,
                         int mode, const double *nodes_cm, std::vector<int32_t> &perm, std::vector<int32_t> &inv)
{
    perm.clear(); inv.clear();
    if (mode == 0 || nbg < 2) return false;
    if (mode == 3 && nodes_cm != nullptr) { morton_renumber(V, nodes_cm, fc, nbg, perm, inv); return true; }
    if (mode == 3) mode = 2;
    if (mode == 1) {
        // locality of a numbering: share of beams whose two rows are further apart than ~16 MB of
        // the value stream (~900 B per row)
        const int64_t window = std::max<int64_t>(4096, (int64_t)(16e6 / 900.0));
        auto poor = [&](const int32_t *pm) {
            int64_t cnt[64] = {0};   // per thread: far, total (integer counts: the split does not change the result)
            int64_t *cp = cnt;
            const int32_t *fcp = fc.data();
            parallel_ranges(E, 1 << 16, [=](int64_t lo, int64_t hi, int t) {
                int64_t far = 0, tot = 0;
                for (int64_t e = lo; e < hi; ++e) {
                    int32_t a = fcp[elems_cm[e]], b = fcp[elems_cm[E + e]];
                    if (a < 0 || b < 0) continue;
                    if (pm) { a = pm[a]; b = pm[b]; }
                    ++tot;
                    far += std::llabs((long long)a - (long long)b) > window;
                }
                cp[2 * t] = far; cp[2 * t + 1] = tot;
            });
            int64_t far = 0, tot = 0;
            for (int t = 0; t < 32; ++t) { far += cnt[2 * t]; tot += cnt[2 * t + 1]; }
            return tot != 0 && far * 20 >= tot;   // >= 5 % far pairs
        };
        if (!poor(nullptr)) return false;   // keep the input order
        // (Z-order, mode 3, costs a fifth of the traversal on the host but its SpMV measured 5 % slower
        // than breadth-first + balanced storage on B200: profiles/README.md -- so auto stays breadth-first)
    }
    (void)V;
    // adjacency of the free-vertex graph (CSR)
    std::vector<int32_t> ptr((size_t)nbg + 1, 0);
    for (int64_t e = 0; e < E; ++e) {
        const int32_t a = fc[elems_cm[e]], b = fc[elems_cm[E + e]];
        if (a < 0 || b < 0 || a == b) continue;
        ptr[a + 1]++; ptr[b + 1]++;
    }
    for (int64_t i = 0; i < nbg; ++i) ptr[i + 1] += ptr[i];
    std::vector<int32_t> adj((size_t)ptr[nbg]), cur(ptr.begin(), ptr.end() - 1);
    for (int64_t e = 0; e < E; ++e) {
        const int32_t a = fc[elems_cm[e]], b = fc[elems_cm[E + e]];
        if (a < 0 || b < 0 || a == b) continue;
        adj[cur[a]++] = b; adj[cur[b]++] = a;
    }
    perm.assign((size_t)nbg, -1);
    inv.resize((size_t)nbg);
    int32_t next = 0;
    for (int64_t seed = 0; seed < nbg; ++seed) {
        if (perm[seed] >= 0) continue;
        perm[seed] = next; inv[next] = (int32_t)seed; ++next;
        for (int32_t head = perm[seed]; head < next; ++head) {   // inv[head..next) is the BFS queue
            const int32_t u = inv[head];
            for (int32_t k = ptr[u]; k < ptr[u + 1]; ++k) {
                const int32_t w = adj[k];
                if (perm[w] < 0) { perm[w] = next; inv[next] = w; ++next; }
            }
        }
    }
    return true;
}

void build_plan_begin(int64_t V, int64_t E, const int32_t *elems_cm, int64_t F, const int32_t *fixed, int32_t rank, int32_t world,
                      HostPlan &P, int reorder, const double *nodes_cm)
{
    RankShare share(world);
    P.V = V; P.E = E; P.rank = rank; P.world = world;
    P.nb_global = build_free_map(V, F, fixed, P.free_index);
    const int64_t nbg = P.nb_global;
    if (bfs_renumber(V, E, elems_cm, P.free_index, nbg, reorder, nodes_cm, P.perm, P.inv)) {
        for (int64_t v = 0; v < V; ++v)
            if (P.free_index[v] >= 0) P.free_index[v] = P.perm[P.free_index[v]];
    }
    P.row_begin = nbg * rank / world;
    P.row_end = nbg * (rank + 1) / world;
}

void build_plan(int64_t V, int64_t E, const int32_t *elems_cm, int64_t F, const int32_t *fixed,
                int32_t rank, int32_t world, bool with_contrib, HostPlan &P, int reorder, const double *nodes_cm,
                const std::function<void()> *after_free_map)
{
    PrepTrace tr;
    build_plan_begin(V, E, elems_cm, F, fixed, rank, world, P, reorder, nodes_cm);
    tr.lap("free map");
    if (after_free_map && *after_free_map) (*after_free_map)();
    build_plan_pattern(E, elems_cm, with_contrib, P);
}

void build_plan_pattern(int64_t E, const int32_t *elems_cm, bool with_contrib, HostPlan &P)
{
    PrepTrace tr;
    const int32_t world = P.world;
    RankShare share(world);
    const int64_t r0 = P.row_begin, nb = P.row_end - P.row_begin;
    const int32_t *fm = P.free_index.data();

    // pass 1: entries per owned row (diagonal contributions + off-diagonal ones).  The owned rows are
    // split into one range per thread.  Every end point that falls into an owned row is first dropped
    // into the bucket (element chunk, row range) it belongs to -- one parallel sweep over the beams -- and
    // each row-range thread then walks its buckets in chunk order, i.e. in ascending element order, so
    // each row's entries stay in the order Eigen's setFromTriplets would sum them.  O(E) work in total
    // (O(E / world) after the sweep on several ranks) instead of one full scan of the beams per thread.
    std::vector<int32_t> &cnt = P.scratch_cnt;
    cnt.assign((size_t)nb + 1, 0);
    const int64_t n_ranges = std::max<int64_t>(1, std::min<int64_t>(host_threads(), nb >> 15));
    const int64_t n_chunks = std::max<int64_t>(1, std::min<int64_t>(host_threads(), E >> 16));
    int64_t bound[34];   // row range t = [bound[t], bound[t + 1])
    for (int64_t t = 0; t <= n_ranges; ++t) bound[t] = nb * t / n_ranges;
    const double to_range = (double)n_ranges / (double)std::max<int64_t>(1, nb);
    auto range_of = [&](int64_t lrow) {   // estimate by multiplication, settle against the bounds
        int64_t t = std::min<int64_t>(n_ranges - 1, (int64_t)((double)lrow * to_range));
        while (lrow >= bound[t + 1]) ++t;
        while (lrow < bound[t]) --t;
        return t;
    };
    std::vector<std::vector<int32_t>> &bucket = P.scratch_bucket;   // (e << 1) | (0: first end, 1: second end)
    if (bucket.size() < (size_t)(n_chunks * n_ranges)) bucket.resize((size_t)(n_chunks * n_ranges));
    auto run_threads = [](int64_t n, auto &&body) {
        if (n == 1) { body((int64_t)0); return; }
        std::vector<std::thread> th;
        for (int64_t t = 0; t < n; ++t) th.emplace_back([&body, t] { body(t); });
        for (auto &x : th) x.join();
    };
    run_threads(n_chunks, [&](int64_t ch) {
        const int64_t lo = E * ch / n_chunks, hi = E * (ch + 1) / n_chunks;
        std::vector<int32_t> *mine = &bucket[(size_t)(ch * n_ranges)];
        for (int64_t t = 0; t < n_ranges; ++t) mine[t].clear();   // capacity survives between jobs
        for (int64_t e = lo; e < hi; ++e) {
            const int64_t la = (int64_t)fm[elems_cm[e]] - r0, lb2 = (int64_t)fm[elems_cm[E + e]] - r0;
            if (la >= 0 && la < nb) mine[range_of(la)].push_back((int32_t)(e << 1));
            if (lb2 >= 0 && lb2 < nb) mine[range_of(lb2)].push_back((int32_t)((e << 1) | 1));
        }
    });
    tr.lap("end points -> buckets");
    // visit(local row, element, is first end, free row of first end, of second end) for every owned end
    // point of row range t, ascending in the element index
    auto scan = [&](int64_t t, auto &&visit) {
        for (int64_t ch = 0; ch < n_chunks; ++ch)
            for (const int32_t code : bucket[(size_t)(ch * n_ranges + t)]) {
                const int64_t e = code >> 1;
                const int64_t fa = fm[elems_cm[e]], fb = fm[elems_cm[E + e]];
                if (code & 1) visit(fb - r0, e, false, fa, fb);
                else visit(fa - r0, e, true, fa, fb);
            }
    };
    run_threads(n_ranges, [&](int64_t t) {
        scan(t, [&](int64_t lrow, int64_t, bool first, int64_t fa, int64_t fb) {
            cnt[lrow] += ((first ? fb : fa) >= 0) ? 2 : 1;
        });
    });
    tr.lap("count pass");
    std::vector<int64_t> &start = P.scratch_start;
    start.resize((size_t)nb + 1);
    start[0] = 0;
    for (int64_t r = 0; r < nb; ++r) start[r + 1] = start[r] + cnt[r];
    const int64_t nent = start[nb];
    struct Ent { int32_t col; int32_t code; };  // global column, (e<<2)|which
    static_assert(sizeof(Ent) == sizeof(int64_t), "Ent is stored in the int64 scratch");
    if (P.scratch_ent.size() < (size_t)nent + 1) P.scratch_ent.resize((size_t)nent + 1);
    Ent *const ent = reinterpret_cast<Ent *>(P.scratch_ent.data());
    {
        std::vector<int64_t> &pos = P.scratch_pos;
        pos.assign(start.begin(), start.end() - 1);
        run_threads(n_ranges, [&](int64_t t) {
            scan(t, [&](int64_t lrow, int64_t e, bool first, int64_t fa, int64_t fb) {
                const int32_t code = (int32_t)(e << 2);
                if (first) {
                    ent[pos[lrow]++] = {(int32_t)fa, code | SUB_00};
                    if (fb >= 0) ent[pos[lrow]++] = {(int32_t)fb, code | SUB_01};
                } else {
                    if (fa >= 0) ent[pos[lrow]++] = {(int32_t)fa, code | SUB_10};
                    ent[pos[lrow]++] = {(int32_t)fb, code | SUB_11};
                }
            });
        });
    }
    tr.lap("fill pass");
    // a free vertex with no beam still owns a (zero) diagonal block so that the
    // row exists; the solver reports it as singular.
    // per row: stable sort by column (keeps ascending element order), count blocks
    P.rowptr.assign((size_t)nb + 1, 0);
    std::vector<int32_t> &nblk = P.scratch_nblk;
    nblk.assign((size_t)nb, 0);
    parallel_ranges(nb, 1 << 14, [&](int64_t lo, int64_t hi, int) {
        for (int64_t r = lo; r < hi; ++r) {
            Ent *b = ent + start[r], *e2 = ent + start[r + 1];
            if (e2 - b <= 24) {
                for (Ent *i = b + 1; i < e2; ++i) {
                    Ent x = *i; Ent *j = i - 1;
                    while (j >= b && j->col > x.col) { j[1] = *j; --j; }
                    j[1] = x;
                }
            } else {
                std::stable_sort(b, e2, [](const Ent &x, const Ent &y) { return x.col < y.col; });
            }
            int32_t n = 0; bool has_diag = false;
            for (Ent *i = b; i < e2; ++i) {
                if (i == b || i->col != i[-1].col) ++n;
                if (i->col == r0 + r) has_diag = true;
            }
            nblk[r] = n + (has_diag ? 0 : 1);
        }
    });
    tr.lap("row sort + count");
    for (int64_t r = 0; r < nb; ++r) P.rowptr[r + 1] = P.rowptr[r] + nblk[r];
    const int64_t nnzb = P.rowptr[nb];
    // every entry of these is overwritten below: resize (no refill when the size is unchanged)
    P.colidx.resize((size_t)nnzb);
    P.diag_idx.resize((size_t)nb);
    if (P.scratch_gcol.size() < (size_t)nnzb + 1) P.scratch_gcol.resize((size_t)nnzb + 1);   // global column ids first
    int32_t *const gcol = P.scratch_gcol.data();
    if (with_contrib) { P.contrib_ptr.resize((size_t)nnzb + 1); P.contrib.resize((size_t)nent); }
    else { P.contrib_ptr.clear(); P.contrib.clear(); }
    parallel_ranges(nb, 1 << 14, [&](int64_t lo, int64_t hi, int) {
        for (int64_t r = lo; r < hi; ++r) {
            const Ent *b = ent + start[r], *e2 = ent + start[r + 1];
            int64_t k = P.rowptr[r] - 1;
            const int32_t me = (int32_t)(r0 + r);
            bool diag_done = false;
            // number of contributions before this row == start[r] (every entry is one)
            int64_t c = start[r];
            for (const Ent *i = b; i < e2; ++i) {
                if (i == b || i->col != i[-1].col) {
                    if (!diag_done && i->col > me) {  // isolated vertex: empty diagonal block
                        ++k; gcol[k] = me; P.diag_idx[r] = (int32_t)k; diag_done = true;
                        if (with_contrib) P.contrib_ptr[k] = (int32_t)c;
                    }
                    ++k; gcol[k] = i->col;
                    if (i->col == me) { P.diag_idx[r] = (int32_t)k; diag_done = true; }
                    if (with_contrib) P.contrib_ptr[k] = (int32_t)c;
                }
                if (with_contrib) P.contrib[c] = i->code;
                ++c;
            }
            if (!diag_done) {
                ++k; gcol[k] = me; P.diag_idx[r] = (int32_t)k;
                if (with_contrib) P.contrib_ptr[k] = (int32_t)c;
            }
        }
    });
    if (with_contrib) P.contrib_ptr[nnzb] = (int32_t)nent;
    tr.lap("pattern + contributor fill");

    // ghosts: global columns outside the owned range, ascending
    P.ghost_global.clear();
    if (world > 1)
        for (int64_t k = 0; k < nnzb; ++k)
            if (gcol[k] < r0 || gcol[k] >= r0 + nb) P.ghost_global.push_back(gcol[k]);
    std::sort(P.ghost_global.begin(), P.ghost_global.end());
    P.ghost_global.erase(std::unique(P.ghost_global.begin(), P.ghost_global.end()), P.ghost_global.end());
    parallel_ranges(nnzb, 1 << 18, [&](int64_t lo, int64_t hi, int) {
        for (int64_t k = lo; k < hi; ++k) {
            const int64_t g = gcol[k];
            if (g >= r0 && g < r0 + nb) P.colidx[k] = (int32_t)(g - r0);
            else {
                const auto it = std::lower_bound(P.ghost_global.begin(), P.ghost_global.end(), g);
                P.colidx[k] = (int32_t)(nb + (it - P.ghost_global.begin()));
            }
        }
    });
    tr.lap("ghosts + local column ids");
    build_halo_lists(P);
    tr.lap("send lists");
}

// Owners of the ghost columns, send lists and neighbour table of a rank from its local pattern (rowptr, colidx with
// local ids) and its ascending ghost list -- the last step of build_plan_pattern, also used when the pattern was
// built on the device.
void build_halo_lists(HostPlan &P)
{
    const int32_t rank = P.rank, world = P.world;
    const int64_t nbg = P.nb_global, nb = P.row_end - P.row_begin;
    const int64_t ng = (int64_t)P.ghost_global.size();
    P.ghost_owner.resize((size_t)ng);
    for (int64_t i = 0; i < ng; ++i) P.ghost_owner[i] = owner_of(P.ghost_global[i], nbg, world);
    // sends: the pattern is symmetric, so rank q needs owned row i from us iff
    // row i has a column owned by q.  Sorted by (dest, row) == the order in which
    // q lists them among its ghosts (ascending global id).
    std::vector<std::pair<int32_t, int32_t>> snd;  // (dest, local row)
    for (int64_t r = 0; world > 1 && r < nb; ++r) {
        int32_t last = -1;
        for (int64_t k = P.rowptr[r]; k < P.rowptr[r + 1]; ++k) {
            if (P.colidx[k] < nb) continue;
            const int32_t q = P.ghost_owner[P.colidx[k] - nb];
            if (q != last) { snd.emplace_back(q, (int32_t)r); last = q; }
        }
    }
    std::sort(snd.begin(), snd.end());
    snd.erase(std::unique(snd.begin(), snd.end()), snd.end());
    P.send_rows.resize(snd.size()); P.send_dest.resize(snd.size());
    for (size_t i = 0; i < snd.size(); ++i) { P.send_dest[i] = snd[i].first; P.send_rows[i] = snd[i].second; }
    // neighbour table
    P.neighbors.clear();
    for (int32_t q = 0; q < world; ++q) {
        if (q == rank) continue;
        Neighbor nbq{q, 0, 0, 0, 0};
        const auto gl = std::lower_bound(P.ghost_owner.begin(), P.ghost_owner.end(), q);
        const auto gh = std::upper_bound(P.ghost_owner.begin(), P.ghost_owner.end(), q);
        nbq.recv_offset = gl - P.ghost_owner.begin(); nbq.recv_count = gh - gl;
        const auto sl = std::lower_bound(P.send_dest.begin(), P.send_dest.end(), q);
        const auto sh = std::upper_bound(P.send_dest.begin(), P.send_dest.end(), q);
        nbq.send_offset = sl - P.send_dest.begin(); nbq.send_count = sh - sl;
        if (nbq.recv_count || nbq.send_count) P.neighbors.push_back(nbq);
    }
}

void build_solver_format(HostPlan &P, bool half)
{
    PrepTrace tr;
    RankShare share(P.world);
    const int64_t nb = P.nb();
    P.sym_half = half;
    P.uptr.assign((size_t)nb + 1, 0);
    P.lptr.assign((size_t)nb + 1, 0);
    // Which row stores the block of a local pair {r, c}?  Default: the upper one (r < c) -- on
    // translation-invariant numberings (lattices, grids in scan order) every row then stores the
    // same number of blocks and the transposed references of a tile stay coalesced.  On irregular
    // numberings the upper counts vary row by row and the 32-row tiles of the interleaved layout
    // pad to their longest row; there the pairs are oriented greedily towards the emptier row
    // (rows ascending, ties to the upper row), which brings every row to ~deg/2 stored blocks.
    std::vector<uint8_t> own;   // per BSR entry: 1 = stored in its own row (balanced orientation only)
    P.balanced = false;
    if (half) {
        int64_t stored = 0, slots = 0;
        {
            const int64_t nt32 = (nb + 31) / 32;
            std::vector<int64_t> part_stored(64, 0), part_slots(64, 0);   // per thread (parallel_ranges uses <= 32)
            parallel_ranges(nt32, 1 << 9, [&](int64_t lo, int64_t hi, int th) {
                int64_t st = 0, sl = 0;
                for (int64_t t = lo; t < hi; ++t) {
                    int32_t mx = 0;
                    for (int64_t r = t * 32; r < std::min<int64_t>(nb, t * 32 + 32); ++r) {
                        int32_t nu = 0;
                        for (int32_t k = P.rowptr[r]; k < P.rowptr[r + 1]; ++k) nu += P.colidx[k] > r;
                        st += nu; mx = std::max(mx, nu);
                    }
                    sl += 32 * (int64_t)mx;
                }
                part_stored[(size_t)th] = st; part_slots[(size_t)th] = sl;
            });
            for (int t = 0; t < 64; ++t) { stored += part_stored[(size_t)t]; slots += part_slots[(size_t)t]; }
        }
        static const bool no_balance = std::getenv("MBFEA_NO_BALANCE") != nullptr;   // A/B knob
        if (slots * 20 > stored * 21 && !no_balance) {   // > 5 % padding under the upper rule
            P.balanced = true;
            own.assign(P.colidx.size(), 1);
            std::vector<int32_t> cnt((size_t)nb, 0), cursor(P.rowptr.begin(), P.rowptr.end() - 1);
            // a row lists its columns in GLOBAL order: ghosts owned by lower ranks (local ids >= nb) come first
            for (int64_t r = 0; r < nb; ++r)
                while (cursor[r] < P.rowptr[r + 1] && P.colidx[cursor[r]] >= nb) ++cursor[r];
            for (int64_t r = 0; r < nb; ++r)
                for (int32_t k = P.rowptr[r]; k < P.rowptr[r + 1]; ++k) {
                    const int32_t c = P.colidx[k];
                    if (c >= nb) { ++cnt[r]; continue; }   // ghost column: always stored here
                    if (c <= r) continue;
                    // the mirrored entry (c, r): row c's local lower entries are met in ascending r
                    const int32_t m = cursor[c]++;
                    const bool mine = cnt[r] <= cnt[c];
                    own[(size_t)k] = mine; own[(size_t)m] = !mine;
                    ++cnt[mine ? r : c];
                }
            // keep it only when it pays (a small regular mesh pads at its faces under either rule)
            int64_t slots_bal = 0;
            for (int64_t t = 0; t * 32 < nb; ++t)
                slots_bal += 32 * (int64_t)*std::max_element(cnt.begin() + t * 32, cnt.begin() + std::min<int64_t>(nb, t * 32 + 32));
            if (slots_bal * 100 > slots * 97) { P.balanced = false; own.clear(); }
        }
    }
    auto stored_here = [&](int64_t r, int32_t k, int32_t c) {
        if (!half || c >= nb) return true;
        return P.balanced ? own[(size_t)k] != 0 : c > r;
    };
    // pass 1: counts.  (r,c) is stored in row r, or reached through row c's block (c,r).
    parallel_ranges(nb, 1 << 14, [&](int64_t lo, int64_t hi, int) {
        for (int64_t r = lo; r < hi; ++r) {
            int32_t nu = 0, nl = 0;
            for (int32_t k = P.rowptr[r]; k < P.rowptr[r + 1]; ++k) {
                const int32_t c = P.colidx[k];
                if (c == r) continue;
                if (stored_here(r, k, c)) ++nu; else ++nl;
            }
            P.uptr[r + 1] = nu; P.lptr[r + 1] = nl;   // counts first, prefix sums below
        }
    });
    for (int64_t r = 0; r < nb; ++r) { P.uptr[r + 1] += P.uptr[r]; P.lptr[r + 1] += P.lptr[r]; }
    P.ucol.resize((size_t)P.uptr[nb]);
    P.usrc.resize((size_t)P.uptr[nb]);
    P.lcol.resize((size_t)P.lptr[nb]);
    P.lblk.resize((size_t)P.lptr[nb]);
    parallel_ranges(nb, 1 << 14, [&](int64_t lo, int64_t hi, int) {
        for (int64_t r = lo; r < hi; ++r) {
            int32_t u = P.uptr[r];
            for (int32_t k = P.rowptr[r]; k < P.rowptr[r + 1]; ++k) {
                const int32_t c = P.colidx[k];
                if (c == r) continue;
                if (!stored_here(r, k, c)) continue;
                P.ucol[u] = c; P.usrc[u] = k; ++u;
            }
        }
    });
    P.tile_base.clear(); P.lofs.clear();
    tr.lap("solver format: stored blocks");
    if (!half) return;
    // interleaved layout: every tile of 32 rows reserves max-row-length block slots per row
    const int64_t ntiles = (nb + 31) / 32;
    P.tile_base.assign((size_t)ntiles + 1, 0);
    for (int64_t t = 0; t < ntiles; ++t) {
        int32_t mx = 0;
        for (int64_t r = t * 32; r < std::min<int64_t>(nb, t * 32 + 32); ++r) mx = std::max(mx, P.uptr[r + 1] - P.uptr[r]);
        const int64_t next = (int64_t)P.tile_base[t] + (int64_t)mx * 3 * 192;
        if (next > INT32_MAX) { P.tile_base.clear(); P.sym_half = false; build_solver_format(P, false); return; }
        P.tile_base[t + 1] = (int32_t)next;
    }
    // pass 2: transposed references.  Row r's list holds its local neighbours whose block lives in the
    // other row, ascending in c (the order of the pattern row); the slot of (c, r) inside row c's
    // short stored list is found by scanning it.
    P.lofs.resize(P.lcol.size());
    parallel_ranges(nb, 1 << 14, [&](int64_t lo, int64_t hi, int) {
        for (int64_t r = lo; r < hi; ++r) {
            int32_t l = P.lptr[r];
            for (int32_t k = P.rowptr[r]; k < P.rowptr[r + 1]; ++k) {
                const int32_t c = P.colidx[k];
                if (c == r || stored_here(r, k, c)) continue;
                int32_t b = P.uptr[c];
                while (P.ucol[b] != (int32_t)r) ++b;   // present by symmetry of the local pattern
                P.lcol[l] = c; P.lblk[l] = b;
                P.lofs[l] = P.tile_base[c / 32] + (b - P.uptr[c]) * 3 * 192 + (int32_t)(c % 32) * 6;
                ++l;
            }
        }
    });
    tr.lap("solver format: transposed refs");
}

void find_components(HostPlan &P, int32_t max_rows, int32_t min_components)
{
    P.comp_ptr.clear(); P.comp_max_rows = 0;
    const int64_t nb = P.nb();
    if (P.world != 1 || nb < min_components) return;
    // Independent row ranges: sweep the rows keeping the furthest column reached so far; a range
    // closes at row r when nothing in it couples beyond r (the pattern is symmetric, so nothing
    // later couples back into it either).  A range may hold several connected components (a mesh
    // whose free vertices fall apart once its fixed vertices are removed stays one range).
    std::vector<int32_t> starts{0};
    int32_t reach = 0, mx = 0;
    for (int64_t r = 0; r < nb; ++r) {
        if (P.rowptr[r + 1] > P.rowptr[r]) reach = std::max(reach, P.colidx[P.rowptr[r + 1] - 1]);  // columns ascending
        if (reach <= (int32_t)r) {
            mx = std::max(mx, (int32_t)(r + 1) - starts.back());
            starts.push_back((int32_t)(r + 1));
        }
    }
    if (starts.back() != (int32_t)nb) return;   // cannot happen for a symmetric pattern
    if ((int64_t)starts.size() - 1 < min_components || mx > max_rows) return;
    P.comp_ptr = std::move(starts);
    P.comp_max_rows = mx;
}

void build_tiles(const hvec<int32_t> &rowptr, int cap_blocks, int max_rows,
                 hvec<int32_t> &tile_row)
{
    const int64_t nb = (int64_t)rowptr.size() - 1;
    tile_row.clear();
    tile_row.push_back(0);
    int64_t r = 0;
    while (r < nb) {
        int64_t r1 = r + 1;  // a tile always holds at least one row (long rows are chunked)
        while (r1 < nb && r1 - r < max_rows && rowptr[r1 + 1] - rowptr[r] <= cap_blocks) ++r1;
        tile_row.push_back((int32_t)r1);
        r = r1;
    }
}

}  // namespace mbfea
