// gc_sparse_api.cu -- host side of the C ABI for general sparse graphs and label images
// (include/medpy_b200_graphcut.h, sections "general sparse graphs" and "label images").
//
// mgc_sparse keeps what the reference's Graph<> keeps on the host while a graph is assembled -- the arc pairs in
// insertion order with accumulated capacities (graph.h:427-480) and tr_cap / flow per add_tweights (graph.h:415-425);
// these are O(1) updates per call exactly like the reference's -- and runs the max-flow on the device
// (gc_sparse.cuh).  mgc_labels keeps a label image resident in HBM and reduces voxel-scale data to the region
// adjacency graph there (gc_labels.cuh); only per-region / per-region-pair results ever reach the host.
// The stable radix sort that orders the contributions by key is cub::DeviceRadixSort (CUDA toolkit).
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <unordered_map>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>

#include "../../include/medpy_b200_graphcut.h"
#include "gc_labels.cuh"
#include "gc_sparse.cuh"

namespace {

thread_local std::string g_sp_create_error;
thread_local std::string g_lab_create_error;

// device allocations of one call, released together
struct DevScope {
    std::vector<void*> ptrs;
    ~DevScope() { for (void* p : ptrs) cudaFree(p); }
    template <typename T>
    cudaError_t alloc(T** out, size_t count)
    {
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, (count ? count : 1) * sizeof(T));
        if (e == cudaSuccess) ptrs.push_back(p);
        *out = (T*)p;
        return e;
    }
};

int bits_for(unsigned long long v)   // number of bits needed to represent values in [0, v]
{
    int b = 0;
    while (v) { ++b; v >>= 1; }
    return b ? b : 1;
}

unsigned grid_for(long long n, int cap = 1184 * 4)
{
    long long b = (n + LAB_BLOCK - 1) / LAB_BLOCK;
    if (b < 1) b = 1;
    return (unsigned)(b < cap ? b : cap);
}

}  // namespace

// =================================================================================================================
// sparse graph
// =================================================================================================================
struct mgc_sparse {
    int device = 0;
    int64_t n = 0;
    std::vector<double> tr;                            // net terminal capacity per node
    double flow_const = 0.0;                           // sum of the add_tweights minima (graph.h:423)
    std::unordered_map<uint64_t, int64_t> pair_of;     // (lo << 32 | hi) -> index into the pair arrays (built on demand)
    bool indexed = true;                               // pair_of covers every pair
    std::vector<int32_t> plo, phi;                     // node pairs in insertion order, lo < hi
    std::vector<double> cap_lh, cap_hl;                // capacity lo->hi, hi->lo
    bool solved = false;
    double energy = 0.0;
    std::vector<uint8_t> mask;
    int push_steps = 4;                                // push steps per node and launch
    int sweeps_per_round = 16;                         // push launches between two global relabels
    int relax_batch = 8;                               // relaxation launches per "changed" read-back
    int64_t max_rounds = 1000000;
    double max_seconds = 600.0;                        // wall-clock cap of one solve (MEDPY_GC_SPARSE_TIMEOUT overrides)
    mgc_stats st{};
    std::string err;
};

#define SPCK(call)                                                                                 \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            g->err = std::string(#call) + ": " + cudaGetErrorString(_e);                           \
            cudaGetLastError();                                                                    \
            return MGC_E_CUDA;                                                                     \
        }                                                                                          \
    } while (0)

#define SPFAIL(code, msg)                                                                          \
    do {                                                                                           \
        g->err = (msg);                                                                            \
        return (code);                                                                             \
    } while (0)

namespace {

void sparse_ensure_index(mgc_sparse* g)
{
    if (g->indexed) return;
    g->pair_of.clear();
    g->pair_of.reserve(g->plo.size() * 2);
    for (size_t p = 0; p < g->plo.size(); ++p)
        g->pair_of.emplace(((uint64_t)(uint32_t)g->plo[p] << 32) | (uint32_t)g->phi[p], (int64_t)p);
    g->indexed = true;
}

int sparse_solve(mgc_sparse* g)
{
    SPCK(cudaSetDevice(g->device));
    const int n = (int)g->n;
    const int64_t np = (int64_t)g->plo.size();
    if (2 * np >= (int64_t)INT32_MAX) SPFAIL(MGC_E_ARG, "too many arcs for 32-bit arc ids");
    const int m2 = (int)(2 * np);
    // CSR in insertion order of the pairs (the order add_edge appends arcs to a node's list, graph.h:443-452)
    std::vector<int> row((size_t)n + 1, 0), head((size_t)m2), sis((size_t)m2);
    std::vector<double> cap((size_t)m2);
    for (int64_t p = 0; p < np; ++p) { row[(size_t)g->plo[p] + 1]++; row[(size_t)g->phi[p] + 1]++; }
    for (int v = 0; v < n; ++v) row[(size_t)v + 1] += row[(size_t)v];
    {
        std::vector<int> fill(row.begin(), row.end() - 1);
        for (int64_t p = 0; p < np; ++p) {
            const int a = fill[(size_t)g->plo[p]]++, b = fill[(size_t)g->phi[p]]++;
            head[(size_t)a] = g->phi[p]; head[(size_t)b] = g->plo[p];
            sis[(size_t)a] = b; sis[(size_t)b] = a;
            cap[(size_t)a] = g->cap_lh[p]; cap[(size_t)b] = g->cap_hl[p];
        }
    }
    DevScope dev;
    SparseState S{};
    int *d_row, *d_head, *d_sis, *d_height, *d_flags;
    double *d_cap, *d_tr, *d_excess, *d_sunk, *d_abs;
    uint8_t* d_mask;
    unsigned long long* d_count;
    SPCK(dev.alloc(&d_row, (size_t)n + 1));
    SPCK(dev.alloc(&d_head, (size_t)m2));
    SPCK(dev.alloc(&d_sis, (size_t)m2));
    SPCK(dev.alloc(&d_cap, (size_t)m2));
    SPCK(dev.alloc(&d_tr, (size_t)n));
    SPCK(dev.alloc(&d_excess, (size_t)n));
    SPCK(dev.alloc(&d_sunk, (size_t)n));
    SPCK(dev.alloc(&d_height, (size_t)n));
    SPCK(dev.alloc(&d_mask, (size_t)n));
    SPCK(dev.alloc(&d_flags, 2));
    SPCK(dev.alloc(&d_abs, 1));
    SPCK(dev.alloc(&d_count, 1));
    SPCK(cudaMemcpy(d_row, row.data(), ((size_t)n + 1) * sizeof(int), cudaMemcpyHostToDevice));
    if (m2) {
        SPCK(cudaMemcpy(d_head, head.data(), (size_t)m2 * sizeof(int), cudaMemcpyHostToDevice));
        SPCK(cudaMemcpy(d_sis, sis.data(), (size_t)m2 * sizeof(int), cudaMemcpyHostToDevice));
        SPCK(cudaMemcpy(d_cap, cap.data(), (size_t)m2 * sizeof(double), cudaMemcpyHostToDevice));
    }
    SPCK(cudaMemcpy(d_tr, g->tr.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
    S.n = n; S.m2 = m2; S.row = d_row; S.head = d_head; S.sis = d_sis; S.cap = d_cap; S.tr = d_tr;
    S.excess = d_excess; S.sunk = d_sunk; S.height = d_height;

    cudaEvent_t ev0, ev1;
    SPCK(cudaEventCreate(&ev0));
    SPCK(cudaEventCreate(&ev1));
    SPCK(cudaEventRecord(ev0, 0));
    const unsigned blocks = grid_for(n);
    k_sp_init<<<blocks, 256>>>(S);
    g->st.kernel_launches++;
    int64_t rounds = 0;
    long long active = 0;
    const auto t_start = std::chrono::steady_clock::now();
    for (;;) {
        // exact global relabel: backward BFS from the sink by in-place relaxation
        k_sp_relabel_init<<<blocks, 256>>>(S);
        g->st.kernel_launches++;
        g->st.global_relabels++;
        for (;;) {
            SPCK(cudaMemsetAsync(d_flags, 0, sizeof(int), 0));
            for (int r = 0; r < g->relax_batch; ++r) k_sp_relax<<<blocks, 256>>>(S, d_flags);
            g->st.kernel_launches += g->relax_batch;
            g->st.relabel_sweeps += g->relax_batch;
            int changed = 0;
            SPCK(cudaMemcpy(&changed, d_flags, sizeof(int), cudaMemcpyDeviceToHost));
            if (!changed) break;
        }
        // stop test, only ever right after an exact relabel
        SPCK(cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), 0));
        k_sp_count_active<<<blocks, 256>>>(S, d_count);
        g->st.kernel_launches++;
        unsigned long long c = 0;
        SPCK(cudaMemcpy(&c, d_count, sizeof(c), cudaMemcpyDeviceToHost));
        active = (long long)c;
        if (!active) break;
        const double elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
        if (++rounds > g->max_rounds || elapsed > g->max_seconds) {
            cudaEventDestroy(ev0); cudaEventDestroy(ev1);
            SPFAIL(MGC_E_NOCONV, "sparse push-relabel did not converge within " + std::to_string(rounds) + " rounds / " +
                                     std::to_string(elapsed) + " s (" + std::to_string(active) + " active nodes left)");
        }
        for (int s = 0; s < g->sweeps_per_round; ++s) k_sp_push<<<blocks, 256>>>(S, g->push_steps, d_flags + 1);
        g->st.kernel_launches += g->sweeps_per_round;
        g->st.push_sweeps += g->sweeps_per_round;
    }
    k_sp_readout<<<1, 256>>>(S, d_mask, d_abs);
    g->st.kernel_launches++;
    SPCK(cudaEventRecord(ev1, 0));
    SPCK(cudaGetLastError());
    g->mask.assign((size_t)n, 0);
    double absorbed = 0.0;
    SPCK(cudaMemcpy(g->mask.data(), d_mask, (size_t)n, cudaMemcpyDeviceToHost));
    SPCK(cudaMemcpy(&absorbed, d_abs, sizeof(double), cudaMemcpyDeviceToHost));
    float ms = 0.f;
    SPCK(cudaEventElapsedTime(&ms, ev0, ev1));
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    g->energy = g->flow_const + absorbed;
    g->solved = true;
    g->st.n_voxels = n;
    g->st.ms_solve = ms;
    g->st.active_last = active;
    g->st.flow_const = g->flow_const;
    g->st.energy = g->energy;
    g->st.device_bytes = (int64_t)((size_t)m2 * 16 + (size_t)n * 33);
    return MGC_OK;
}

}  // namespace

extern "C" {

int mgc_sparse_create(int64_t n_nodes, int32_t device, mgc_sparse** out)
{
    if (!out) return MGC_E_ARG;
    *out = nullptr;
    if (n_nodes < 1 || n_nodes >= (int64_t)INT32_MAX) { g_sp_create_error = "node count must be in [1, 2^31-2]"; return MGC_E_ARG; }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count < 1) {
        cudaGetLastError();
        g_sp_create_error = "no CUDA device available (this library has no CPU solver)";
        return MGC_E_CUDA;
    }
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
    if (device >= count) { g_sp_create_error = "invalid CUDA device ordinal"; return MGC_E_ARG; }
    mgc_sparse* g = new mgc_sparse();
    g->device = device;
    g->n = n_nodes;
    g->tr.assign((size_t)n_nodes, 0.0);
    if (const char* e = std::getenv("MEDPY_GC_SPARSE_TIMEOUT")) { const double v = std::atof(e); if (v > 0) g->max_seconds = v; }
    if (const char* e = std::getenv("MEDPY_GC_SPARSE_SWEEPS")) { const int v = std::atoi(e); if (v > 0) g->sweeps_per_round = v; }
    *out = g;
    return MGC_OK;
}

void mgc_sparse_destroy(mgc_sparse* g) { delete g; }

int mgc_sparse_reset(mgc_sparse* g)
{
    if (!g) return MGC_E_ARG;
    std::fill(g->tr.begin(), g->tr.end(), 0.0);
    g->flow_const = 0.0;
    g->pair_of.clear();
    g->indexed = true;
    g->plo.clear(); g->phi.clear(); g->cap_lh.clear(); g->cap_hl.clear();
    g->solved = false;
    g->energy = 0.0;
    g->mask.clear();
    g->st = mgc_stats{};
    return MGC_OK;
}

const char* mgc_sparse_last_error(const mgc_sparse* g) { return g ? g->err.c_str() : g_sp_create_error.c_str(); }

int mgc_sparse_sum_edges(mgc_sparse* g, int64_t count, const int32_t* i, const int32_t* j, const double* cap, const double* rev_cap)
{
    if (!g) return MGC_E_ARG;
    if (count < 0 || (count > 0 && (!i || !j || !cap || !rev_cap))) SPFAIL(MGC_E_ARG, "null edge arrays");
    for (int64_t k = 0; k < count; ++k) {
        const int64_t a = i[k], b = j[k];
        if (a < 0 || b < 0 || a >= g->n || b >= g->n)
            SPFAIL(MGC_E_ARG, "Invalid node id in edge (" + std::to_string(a) + ", " + std::to_string(b) + "). Valid values are 0 to " +
                                  std::to_string(g->n - 1) + ".");
        if (a == b) SPFAIL(MGC_E_ARG, "The node_from (" + std::to_string(a) + ") can not be equal to the node_to (" + std::to_string(b) + ") (self-connections are forbidden in graph-cuts).");
    }
    // A batch of distinct pairs in strictly increasing (i, j) order with i < j -- what the region adjacency reduction
    // delivers (mgc_labels_fetch_edges) -- landing in an empty graph needs no look-ups: append, index later if ever needed.
    if (g->plo.empty() && count > 0) {
        bool sorted_unique = true;
        uint64_t prev = 0;
        for (int64_t k = 0; k < count && sorted_unique; ++k) {
            const uint64_t key = ((uint64_t)(uint32_t)i[k] << 32) | (uint32_t)j[k];
            sorted_unique = i[k] < j[k] && (k == 0 || key > prev);
            prev = key;
        }
        if (sorted_unique) {
            g->plo.assign(i, i + count);
            g->phi.assign(j, j + count);
            g->cap_lh.assign(cap, cap + count);
            g->cap_hl.assign(rev_cap, rev_cap + count);
            g->pair_of.clear();
            g->indexed = false;
            g->solved = false;
            return MGC_OK;
        }
    }
    sparse_ensure_index(g);
    for (int64_t k = 0; k < count; ++k) {
        const bool fwd = i[k] < j[k];
        const int32_t lo = fwd ? i[k] : j[k], hi = fwd ? j[k] : i[k];
        const uint64_t key = ((uint64_t)(uint32_t)lo << 32) | (uint32_t)hi;
        auto it = g->pair_of.find(key);
        const double c_lh = fwd ? cap[k] : rev_cap[k], c_hl = fwd ? rev_cap[k] : cap[k];
        if (it == g->pair_of.end()) {
            g->pair_of.emplace(key, (int64_t)g->plo.size());
            g->plo.push_back(lo); g->phi.push_back(hi);
            g->cap_lh.push_back(c_lh); g->cap_hl.push_back(c_hl);      // add_edge: r_cap = cap (graph.h:449-450)
        } else {
            g->cap_lh[(size_t)it->second] += c_lh;                       // sum_edge: r_cap += cap (graph.h:472-476)
            g->cap_hl[(size_t)it->second] += c_hl;
        }
    }
    if (count) g->solved = false;
    return MGC_OK;
}

int mgc_sparse_add_tweights(mgc_sparse* g, int64_t count, const int32_t* nodes, const double* src, const double* snk)
{
    if (!g) return MGC_E_ARG;
    if (count < 0 || (count > 0 && (!src || !snk))) SPFAIL(MGC_E_ARG, "null t-weight arrays");
    if (!nodes && count > g->n) SPFAIL(MGC_E_ARG, "more t-weights than nodes");
    if (nodes)
        for (int64_t k = 0; k < count; ++k)
            if (nodes[k] < 0 || nodes[k] >= g->n)
                SPFAIL(MGC_E_ARG, "Invalid node id of " + std::to_string(nodes[k]) + ". Valid values are 0 to " + std::to_string(g->n - 1) + ".");
    for (int64_t k = 0; k < count; ++k) {
        const size_t v = (size_t)(nodes ? nodes[k] : k);
        double s = src[k], t = snk[k];
        const double delta = g->tr[v];                 // graph.h:418-424
        if (delta > 0) s += delta; else t -= delta;
        g->flow_const += (s < t) ? s : t;
        g->tr[v] = s - t;
    }
    if (count) g->solved = false;
    return MGC_OK;
}

int mgc_sparse_maxflow(mgc_sparse* g, double* energy)
{
    if (!g) return MGC_E_ARG;
    if (!g->solved) {
        int rc = sparse_solve(g);
        if (rc) return rc;
    }
    if (energy) *energy = g->energy;
    return MGC_OK;
}

int mgc_sparse_get_mask(mgc_sparse* g, uint8_t* out)
{
    if (!g || !out) return MGC_E_ARG;
    if (!g->solved) { int rc = sparse_solve(g); if (rc) return rc; }
    std::memcpy(out, g->mask.data(), (size_t)g->n);
    return MGC_OK;
}

int mgc_sparse_what_segment(mgc_sparse* g, int64_t node, int32_t* segment)
{
    if (!g || !segment) return MGC_E_ARG;
    if (node < 0 || node >= g->n) SPFAIL(MGC_E_ARG, "node id out of range");
    if (!g->solved) { int rc = sparse_solve(g); if (rc) return rc; }
    *segment = g->mask[(size_t)node] ? MGC_SOURCE : MGC_SINK;
    return MGC_OK;
}

int mgc_sparse_get_edge(const mgc_sparse* g, int64_t i, int64_t j, double* cap)
{
    if (!g || !cap) return MGC_E_ARG;
    sparse_ensure_index(const_cast<mgc_sparse*>(g));   // the index is a cache: building it does not change the graph
    *cap = 0.0;
    if (i < 0 || j < 0 || i >= g->n || j >= g->n || i == j) return MGC_OK;
    const bool fwd = i < j;
    const uint64_t key = ((uint64_t)(uint32_t)(fwd ? i : j) << 32) | (uint32_t)(fwd ? j : i);
    auto it = g->pair_of.find(key);
    if (it != g->pair_of.end()) *cap = fwd ? g->cap_lh[(size_t)it->second] : g->cap_hl[(size_t)it->second];
    return MGC_OK;
}

int mgc_sparse_get_trcap(const mgc_sparse* g, int64_t node, double* trcap)
{
    if (!g || !trcap || node < 0 || node >= g->n) return MGC_E_ARG;
    *trcap = g->tr[(size_t)node];
    return MGC_OK;
}

int mgc_sparse_get_node_num(const mgc_sparse* g, int64_t* n) { if (!g || !n) return MGC_E_ARG; *n = g->n; return MGC_OK; }
int mgc_sparse_get_arc_num(const mgc_sparse* g, int64_t* n) { if (!g || !n) return MGC_E_ARG; *n = 2 * (int64_t)g->plo.size(); return MGC_OK; }
int mgc_sparse_get_stats(const mgc_sparse* g, mgc_stats* out) { if (!g || !out) return MGC_E_ARG; *out = g->st; return MGC_OK; }

}  // extern "C"

// =================================================================================================================
// label images
// =================================================================================================================
struct mgc_labels {
    int device = 0;
    LabGeom G{};
    int64_t shape[4] = {1, 1, 1, 1};
    int* labels = nullptr;          // dense int32, C order
    bool owns_labels = false;
    int64_t k = 0;                  // regions
    std::vector<int32_t> ei, ej;    // last boundary result, sorted by (i, j)
    std::vector<double> ew, er;
    int64_t kernel_launches = 0;
    std::string err;
};

#undef SPCK
#undef SPFAIL
#define LBCK(call)                                                                                 \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            l->err = std::string(#call) + ": " + cudaGetErrorString(_e);                           \
            cudaGetLastError();                                                                    \
            return MGC_E_CUDA;                                                                     \
        }                                                                                          \
    } while (0)
#define LBFAIL(code, msg)                                                                          \
    do {                                                                                           \
        l->err = (msg);                                                                            \
        return (code);                                                                             \
    } while (0)

namespace {

// Dense, C-ordered device copy of `a` over the handle's shape.  *out points into `a` itself when it already is a dense
// device array, otherwise into memory owned by `scope`.
template <typename E>
int lab_stage(mgc_labels* l, const mgc_array* a, DevScope& scope, const E** out)
{
    if (!a || !a->data) LBFAIL(MGC_E_ARG, "null array");
    const size_t es = sizeof(E);
    LabStrides st{};
    bool contiguous = true;
    long long span = (long long)es, expect = (long long)es;
    for (int d = l->G.nd - 1; d >= 0; --d) {
        long long s = (long long)a->strides[d];
        if (l->G.dim[d] > 1) {
            if (s <= 0) LBFAIL(MGC_E_ARG, "array strides must be positive (pass a contiguous copy)");
            if (s != expect) contiguous = false;
            span += (l->G.dim[d] - 1) * s;
        } else {
            s = 0;
        }
        st.s[d] = s;
        expect *= l->G.dim[d];
    }
    const size_t bytes = (size_t)l->G.n * es;
    if (contiguous && a->mem == MGC_MEM_DEVICE) { *out = (const E*)a->data; return MGC_OK; }
    E* dst = nullptr;
    LBCK(scope.alloc(&dst, (size_t)l->G.n));
    if (contiguous) {
        LBCK(cudaMemcpy(dst, a->data, bytes, cudaMemcpyHostToDevice));
        *out = dst;
        return MGC_OK;
    }
    const char* src = (const char*)a->data;
    if (a->mem == MGC_MEM_HOST) {
        char* raw = nullptr;
        LBCK(scope.alloc(&raw, (size_t)span));
        LBCK(cudaMemcpy(raw, a->data, (size_t)span, cudaMemcpyHostToDevice));
        src = raw;
    }
    k_lab_gather<E><<<(unsigned)((l->G.n + LAB_BLOCK - 1) / LAB_BLOCK), LAB_BLOCK>>>(l->G, src, st, dst);
    l->kernel_launches++;
    LBCK(cudaGetLastError());
    *out = dst;
    return MGC_OK;
}

template <typename E, int MODE>
int lab_boundary_run(mgc_labels* l, const E* grad, double directedness)
{
    DevScope dev;
    const long long items = (long long)l->G.nd * l->G.n;
    const long long nb = (items + LAB_BLOCK - 1) / LAB_BLOCK;
    if (nb >= (long long)INT32_MAX) LBFAIL(MGC_E_ARG, "label image too large");
    unsigned* block_count;
    unsigned long long* block_off;
    LBCK(dev.alloc(&block_count, (size_t)nb));
    LBCK(dev.alloc(&block_off, (size_t)nb + 1));
    k_lab_pair_count<<<(unsigned)nb, LAB_BLOCK>>>(l->G, l->labels, MODE == 2 ? 1 : 0, block_count);
    k_lab_scan_blocks<<<1, 1024>>>(block_count, nb, block_off);
    l->kernel_launches += 2;
    LBCK(cudaGetLastError());
    unsigned long long m_u = 0;
    LBCK(cudaMemcpy(&m_u, block_off + nb, sizeof(m_u), cudaMemcpyDeviceToHost));
    const long long m = (long long)m_u;
    l->ei.clear(); l->ej.clear(); l->ew.clear(); l->er.clear();
    if (m == 0) return MGC_OK;
    if (m >= (1ll << 32)) LBFAIL(MGC_E_ARG, "more than 2^32 border voxel pairs");
    unsigned long long *keys, *keys_sorted;
    unsigned *perm = nullptr, *perm_sorted = nullptr;
    double *wf = nullptr, *wr = nullptr, *wf_s = nullptr, *wr_s = nullptr;
    LBCK(dev.alloc(&keys, (size_t)m));
    LBCK(dev.alloc(&keys_sorted, (size_t)m));
    if (MODE >= 1) { LBCK(dev.alloc(&wf, (size_t)m)); LBCK(dev.alloc(&wf_s, (size_t)m)); }
    if (MODE == 2) { LBCK(dev.alloc(&wr, (size_t)m)); LBCK(dev.alloc(&wr_s, (size_t)m)); }
    k_lab_pair_emit<E, MODE><<<(unsigned)nb, LAB_BLOCK>>>(l->G, l->labels, grad, directedness < 0 ? -directedness : directedness,
                                                          directedness < 0 ? 1 : 0, block_off, keys, wf, wr);
    l->kernel_launches++;
    LBCK(cudaGetLastError());
    // stable sort by key: contributions of one region pair stay in the reference's order
    const unsigned mblocks = (unsigned)((m + LAB_BLOCK - 1) / LAB_BLOCK);
    const int end_bit = 32 + bits_for((unsigned long long)l->k);
    if (MODE >= 1) {
        LBCK(dev.alloc(&perm, (size_t)m));
        LBCK(dev.alloc(&perm_sorted, (size_t)m));
        k_lab_iota<<<mblocks, LAB_BLOCK>>>(perm, m);
        size_t tb = 0;
        LBCK(cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, keys_sorted, perm, perm_sorted, m, 0, end_bit, 0));
        char* tmp;
        LBCK(dev.alloc(&tmp, tb));
        LBCK(cub::DeviceRadixSort::SortPairs(tmp, tb, keys, keys_sorted, perm, perm_sorted, m, 0, end_bit, 0));
        k_lab_permute<<<mblocks, LAB_BLOCK>>>(wf, perm_sorted, m, wf_s);
        if (MODE == 2) k_lab_permute<<<mblocks, LAB_BLOCK>>>(wr, perm_sorted, m, wr_s);
        l->kernel_launches += 3;
    } else {
        size_t tb = 0;
        LBCK(cub::DeviceRadixSort::SortKeys(nullptr, tb, keys, keys_sorted, m, 0, end_bit, 0));
        char* tmp;
        LBCK(dev.alloc(&tmp, tb));
        LBCK(cub::DeviceRadixSort::SortKeys(tmp, tb, keys, keys_sorted, m, 0, end_bit, 0));
    }
    LBCK(cudaGetLastError());
    // one output slot per run of equal keys, in key order (heads per block -> scan -> slot)
    unsigned* head_count;
    unsigned long long* head_off;
    LBCK(dev.alloc(&head_count, (size_t)mblocks));
    LBCK(dev.alloc(&head_off, (size_t)mblocks + 1));
    k_lab_seg_head_count<<<mblocks, LAB_BLOCK>>>(keys_sorted, m, head_count);
    k_lab_scan_blocks<<<1, 1024>>>(head_count, (long long)mblocks, head_off);
    unsigned long long u = 0;
    LBCK(cudaMemcpy(&u, head_off + mblocks, sizeof(u), cudaMemcpyDeviceToHost));
    unsigned long long* out_key;
    double *out_f, *out_r;
    LBCK(dev.alloc(&out_key, (size_t)u));
    LBCK(dev.alloc(&out_f, (size_t)u));
    LBCK(dev.alloc(&out_r, (size_t)u));
    k_lab_seg_reduce<<<mblocks, LAB_BLOCK>>>(keys_sorted, wf_s, wr_s, m, head_off, out_key, out_f, out_r);
    l->kernel_launches += 3;
    LBCK(cudaGetLastError());
    std::vector<unsigned long long> hk((size_t)u);
    l->ew.resize((size_t)u); l->er.resize((size_t)u);
    LBCK(cudaMemcpy(hk.data(), out_key, (size_t)u * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    LBCK(cudaMemcpy(l->ew.data(), out_f, (size_t)u * sizeof(double), cudaMemcpyDeviceToHost));
    LBCK(cudaMemcpy(l->er.data(), out_r, (size_t)u * sizeof(double), cudaMemcpyDeviceToHost));
    l->ei.resize((size_t)u); l->ej.resize((size_t)u);
    for (size_t t = 0; t < (size_t)u; ++t) {
        l->ei[t] = (int32_t)(hk[t] >> 32);
        l->ej[t] = (int32_t)(hk[t] & 0xffffffffull);
    }
    return MGC_OK;
}

template <typename E>
int lab_boundary_dispatch(mgc_labels* l, int kind, const mgc_array* values, double directedness)
{
    DevScope scope;
    const E* grad = nullptr;
    int rc = lab_stage<E>(l, values, scope, &grad);
    if (rc) return rc;
    if (kind == MGC_LABELS_STAWIASKI) return lab_boundary_run<E, 1>(l, grad, 0.0);
    return lab_boundary_run<E, 2>(l, grad, directedness);
}

// V = element type the sums are formed in (double: bincount; float/double: numpy.sum of a float array)
template <typename E, typename V, bool PAIRWISE>
int lab_region_sums_run(mgc_labels* l, const mgc_array* values, double* sums, int64_t* counts)
{
    DevScope dev;
    const E* vals_in = nullptr;
    int rc = lab_stage<E>(l, values, dev, &vals_in);
    if (rc) return rc;
    const long long n = l->G.n;
    if (n >= (long long)INT32_MAX) LBFAIL(MGC_E_ARG, "label image too large");
    unsigned *keys, *keys_sorted;
    V *vals, *vals_sorted;
    double* d_sums;
    long long* d_counts;
    LBCK(dev.alloc(&keys, (size_t)n));
    LBCK(dev.alloc(&keys_sorted, (size_t)n));
    LBCK(dev.alloc(&vals, (size_t)n));
    LBCK(dev.alloc(&vals_sorted, (size_t)n));
    LBCK(dev.alloc(&d_sums, (size_t)l->k));
    LBCK(dev.alloc(&d_counts, (size_t)l->k));
    const unsigned blocks = (unsigned)((n + LAB_BLOCK - 1) / LAB_BLOCK);
    k_lab_region_items<E, V><<<blocks, LAB_BLOCK>>>(l->labels, vals_in, n, keys, vals);
    size_t tb = 0;
    const int end_bit = bits_for((unsigned long long)l->k);
    LBCK(cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, keys_sorted, vals, vals_sorted, n, 0, end_bit, 0));
    char* tmp;
    LBCK(dev.alloc(&tmp, tb));
    LBCK(cub::DeviceRadixSort::SortPairs(tmp, tb, keys, keys_sorted, vals, vals_sorted, n, 0, end_bit, 0));
    if (PAIRWISE) k_lab_region_reduce_pairwise<V><<<blocks, LAB_BLOCK>>>(keys_sorted, vals_sorted, n, d_sums, d_counts);
    else          k_lab_region_reduce<<<blocks, LAB_BLOCK>>>(keys_sorted, (const double*)vals_sorted, n, d_sums, d_counts);
    l->kernel_launches += 2;
    LBCK(cudaGetLastError());
    std::vector<long long> hc((size_t)l->k);
    LBCK(cudaMemcpy(sums, d_sums, (size_t)l->k * sizeof(double), cudaMemcpyDeviceToHost));
    LBCK(cudaMemcpy(hc.data(), d_counts, (size_t)l->k * sizeof(long long), cudaMemcpyDeviceToHost));
    if (counts) for (int64_t r = 0; r < l->k; ++r) counts[r] = (int64_t)hc[(size_t)r];
    return MGC_OK;
}

}  // namespace

extern "C" {

int mgc_labels_create(int32_t ndim, const int64_t* shape, const mgc_array* labels, int32_t device, mgc_labels** out)
{
    if (!out) return MGC_E_ARG;
    *out = nullptr;
    if (ndim < 1 || ndim > MGC_MAX_NDIM || !shape || !labels) { g_lab_create_error = "label images must have 1 to 4 dimensions"; return MGC_E_ARG; }
    if (labels->dtype != MGC_I32) { g_lab_create_error = "label image must be int32"; return MGC_E_ARG; }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count < 1) {
        cudaGetLastError();
        g_lab_create_error = "no CUDA device available (this library has no CPU path)";
        return MGC_E_CUDA;
    }
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
    if (device >= count) { g_lab_create_error = "invalid CUDA device ordinal"; return MGC_E_ARG; }
    mgc_labels* l = new mgc_labels();
    l->device = device;
    l->G.nd = ndim;
    long long n = 1;
    for (int d = 0; d < 4; ++d) { l->G.dim[d] = 1; l->G.stride[d] = 1; }
    for (int d = 0; d < ndim; ++d) {
        if (shape[d] < 1) { g_lab_create_error = "empty label image"; delete l; return MGC_E_LABELS; }
        l->G.dim[d] = shape[d];
        l->shape[d] = shape[d];
        n *= shape[d];
    }
    if (n >= (long long)INT32_MAX) { g_lab_create_error = "label image too large (2^31 voxels)"; delete l; return MGC_E_ARG; }
    l->G.n = n;
    long long acc = 1;
    for (int d = ndim - 1; d >= 0; --d) { l->G.stride[d] = acc; acc *= l->G.dim[d]; }
    auto fail = [&](int code) { g_lab_create_error = l->err; mgc_labels_destroy(l); return code; };
    if (cudaSetDevice(device) != cudaSuccess) { l->err = "cudaSetDevice failed"; return fail(MGC_E_CUDA); }
    {
        DevScope scope;
        const int* p = nullptr;
        int rc = lab_stage<int>(l, labels, scope, &p);
        if (rc) return fail(rc);
        if (!scope.ptrs.empty()) {
            // the dense copy is the last block the staging allocated: keep it, release the rest with the scope
            l->labels = (int*)p;
            l->owns_labels = true;
            scope.ptrs.erase(std::find(scope.ptrs.begin(), scope.ptrs.end(), (void*)p));
        } else {
            l->labels = (int*)p;        // dense device array of the caller: read in place
        }
        // __check_label_image: min == 1 and every id up to max present
        int* mm = nullptr;
        unsigned long long* cnt = nullptr;
        if (scope.alloc(&mm, 2) != cudaSuccess || scope.alloc(&cnt, 1) != cudaSuccess) { l->err = "device allocation failed"; return fail(MGC_E_NOMEM); }
        const int init[2] = {INT32_MAX, INT32_MIN};
        cudaMemcpy(mm, init, sizeof(init), cudaMemcpyHostToDevice);
        k_lab_minmax<<<grid_for(n), LAB_BLOCK>>>(l->labels, n, mm);
        int got[2] = {0, 0};
        if (cudaMemcpy(got, mm, sizeof(got), cudaMemcpyDeviceToHost) != cudaSuccess) { l->err = std::string("label scan failed: ") + cudaGetErrorString(cudaGetLastError()); return fail(MGC_E_CUDA); }
        const char* msg = "The supplied label image does either not contain any regions or they are not labeled consecutively starting from 1.";
        if (got[0] != 1 || got[1] < 1) { l->err = msg; return fail(MGC_E_LABELS); }
        uint8_t* present = nullptr;
        if (scope.alloc(&present, (size_t)got[1]) != cudaSuccess) { l->err = "device allocation failed"; return fail(MGC_E_NOMEM); }
        cudaMemset(present, 0, (size_t)got[1]);
        cudaMemset(cnt, 0, sizeof(unsigned long long));
        k_lab_presence<<<grid_for(n), LAB_BLOCK>>>(l->labels, n, present);
        k_lab_count_u8<<<grid_for(got[1]), LAB_BLOCK>>>(present, got[1], cnt);
        unsigned long long c = 0;
        if (cudaMemcpy(&c, cnt, sizeof(c), cudaMemcpyDeviceToHost) != cudaSuccess) { l->err = std::string("label scan failed: ") + cudaGetErrorString(cudaGetLastError()); return fail(MGC_E_CUDA); }
        l->kernel_launches += 3;
        if ((long long)c != (long long)got[1]) { l->err = msg; return fail(MGC_E_LABELS); }
        l->k = got[1];
    }
    *out = l;
    return MGC_OK;
}

void mgc_labels_destroy(mgc_labels* l)
{
    if (!l) return;
    if (l->owns_labels && l->labels) { cudaSetDevice(l->device); cudaFree(l->labels); }
    delete l;
}

const char* mgc_labels_last_error(const mgc_labels* l) { return l ? l->err.c_str() : g_lab_create_error.c_str(); }

int mgc_labels_region_count(const mgc_labels* l, int64_t* k) { if (!l || !k) return MGC_E_ARG; *k = l->k; return MGC_OK; }

int mgc_labels_boundary(mgc_labels* l, int32_t kind, const mgc_array* values, double directedness, int64_t* n_edges)
{
    if (!l) return MGC_E_ARG;
    LBCK(cudaSetDevice(l->device));
    int rc;
    if (kind == MGC_LABELS_ADJACENCY) {
        rc = lab_boundary_run<float, 0>(l, nullptr, 0.0);
    } else if (kind == MGC_LABELS_STAWIASKI || kind == MGC_LABELS_STAWIASKI_DIRECTED) {
        if (!values) LBFAIL(MGC_E_ARG, "the boundary term needs the gradient image");
        switch (values->dtype) {
            case MGC_F32: rc = lab_boundary_dispatch<float>(l, kind, values, directedness); break;
            case MGC_F64: rc = lab_boundary_dispatch<double>(l, kind, values, directedness); break;
            case MGC_U8: rc = lab_boundary_dispatch<uint8_t>(l, kind, values, directedness); break;
            case MGC_I16: rc = lab_boundary_dispatch<int16_t>(l, kind, values, directedness); break;
            case MGC_I32: rc = lab_boundary_dispatch<int32_t>(l, kind, values, directedness); break;
            default: LBFAIL(MGC_E_ARG, "unsupported dtype");
        }
    } else {
        LBFAIL(MGC_E_ARG, "unknown label boundary term");
    }
    if (rc) return rc;
    if (n_edges) *n_edges = (int64_t)l->ei.size();
    return MGC_OK;
}

int mgc_labels_fetch_edges(const mgc_labels* l, int32_t* i, int32_t* j, double* w_ij, double* w_ji)
{
    if (!l) return MGC_E_ARG;
    const size_t u = l->ei.size();
    if (u && (!i || !j)) return MGC_E_ARG;
    if (u) {
        std::memcpy(i, l->ei.data(), u * sizeof(int32_t));
        std::memcpy(j, l->ej.data(), u * sizeof(int32_t));
        if (w_ij) std::memcpy(w_ij, l->ew.data(), u * sizeof(double));
        if (w_ji) std::memcpy(w_ji, l->er.data(), u * sizeof(double));
    }
    return MGC_OK;
}

int mgc_labels_region_sums(mgc_labels* l, const mgc_array* values, int32_t mode, double* sums, int64_t* counts)
{
    if (!l || !values || !sums) return MGC_E_ARG;
    LBCK(cudaSetDevice(l->device));
    const bool pw = (mode == MGC_SUM_PAIRWISE);
    switch (values->dtype) {
        case MGC_F32: return pw ? lab_region_sums_run<float, float, true>(l, values, sums, counts)
                                : lab_region_sums_run<float, double, false>(l, values, sums, counts);
        case MGC_F64: return pw ? lab_region_sums_run<double, double, true>(l, values, sums, counts)
                                : lab_region_sums_run<double, double, false>(l, values, sums, counts);
        case MGC_U8: return lab_region_sums_run<uint8_t, double, false>(l, values, sums, counts);
        case MGC_I16: return lab_region_sums_run<int16_t, double, false>(l, values, sums, counts);
        case MGC_I32: return lab_region_sums_run<int32_t, double, false>(l, values, sums, counts);
        default: LBFAIL(MGC_E_ARG, "unsupported dtype");
    }
}

int mgc_labels_region_flags(mgc_labels* l, const mgc_array* markers, uint8_t* flags)
{
    if (!l || !markers || !flags) return MGC_E_ARG;
    if (markers->dtype != MGC_U8) LBFAIL(MGC_E_ARG, "markers must be uint8 / bool");
    LBCK(cudaSetDevice(l->device));
    DevScope dev;
    const uint8_t* m = nullptr;
    int rc = lab_stage<uint8_t>(l, markers, dev, &m);
    if (rc) return rc;
    uint8_t* d_flags;
    LBCK(dev.alloc(&d_flags, (size_t)l->k));
    LBCK(cudaMemset(d_flags, 0, (size_t)l->k));
    k_lab_region_flags<<<grid_for(l->G.n), LAB_BLOCK>>>(l->labels, m, l->G.n, d_flags);
    l->kernel_launches++;
    LBCK(cudaGetLastError());
    LBCK(cudaMemcpy(flags, d_flags, (size_t)l->k, cudaMemcpyDeviceToHost));
    return MGC_OK;
}

int mgc_labels_apply(mgc_labels* l, const uint8_t* per_region, uint8_t* out, int32_t out_mem)
{
    if (!l || !per_region || !out) return MGC_E_ARG;
    LBCK(cudaSetDevice(l->device));
    DevScope dev;
    uint8_t *d_reg, *d_out = out;
    LBCK(dev.alloc(&d_reg, (size_t)l->k));
    LBCK(cudaMemcpy(d_reg, per_region, (size_t)l->k, cudaMemcpyHostToDevice));
    if (out_mem == MGC_MEM_HOST) LBCK(dev.alloc(&d_out, (size_t)l->G.n));
    k_lab_apply<<<grid_for(l->G.n), LAB_BLOCK>>>(l->labels, d_reg, l->G.n, d_out);
    l->kernel_launches++;
    LBCK(cudaGetLastError());
    if (out_mem == MGC_MEM_HOST) LBCK(cudaMemcpy(out, d_out, (size_t)l->G.n, cudaMemcpyDeviceToHost));
    else LBCK(cudaDeviceSynchronize());
    return MGC_OK;
}

}  // extern "C"
