// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Bulk C-ABI driver around the *unmodified* reference Boykov-Kolmogorov solver
// (lib/maxflow/src/{graph.h,graph.cpp,maxflow.cpp,block.h,instances.inc}).  The reference
// sources are compiled where they lie under /root/reference by oracle/Makefile; nothing from
// them is copied into this repository.  The output lands in oracle/_ref/libbkref.so
// (git-ignored, but it travels to the GPU box with gpurun).
//
// The driver replays exactly the call sequence the reference Python layer issues
// (medpy/graphcut/generate.py:120-174):
//   GraphDouble(N, E); add_node(N)                         graph.py:305-306
//   regional:  add_tweights(v, src[v], snk[v])  v=0..N-1   graph.py:551-552 -> graph.h:415-425
//   boundary:  sum_edge(p, p+stride_d, w, w') axis by axis energy_voxel.py:637-664 -> graph.h:456-480
//   fg:        add_tweights(v, 65535, 0)                   graph.py:341-344
//   bg:        add_tweights(v, 0, 65535)                   graph.py:377-380
//   maxflow()                                              maxflow.cpp:471-604
//   what_segment(v)                                        graph.h:560-571
// so the returned flow is bit-identical to what the reference's own Python path returns for the
// same fp64 weights (same arc insertion order => same augmentation order).
//
// Only tests/, bench.py's cpu_baseline / --impl reference arm and __graft_entry__.smoke() may load it.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <chrono>
#include "graph.h"

typedef Graph<double, double, double> GD;

extern "C" {

// Lattice solve.  shape[ndim] logical C-order shape.  For axis d the arrays wf[d] / wb[d] hold, for
// every voxel p (C-order flat index, N entries) the capacity p -> p+stride_d (wf) and
// p+stride_d -> p (wb); entries of voxels on the last plane of axis d are ignored.  Either of
// src/snk may be NULL (no regional term).  fg/bg are uint8[N] or NULL.  A NULL wf[d] skips the axis.
// use_sum_edge != 0 uses sum_edge (what GCGraph.set_nweight calls); 0 uses add_edge (faster fill,
// identical arc order when every pair is set once).
// mask_out: uint8[N], 1 unless what_segment == SINK (bin/medpy_graphcut_voxel.py:178-181).
// times_out[3]: seconds for {graph fill, maxflow(), read-out}.
// Returns 0, or -1 when the reference's int32 ids cannot hold the instance (graph.h:62,82).
int bkref_lattice_solve(int ndim, const int64_t* shape,
                        const double* const* wf, const double* const* wb,
                        const double* src, const double* snk,
                        const uint8_t* fg, const uint8_t* bg,
                        int use_sum_edge,
                        uint8_t* mask_out, double* flow_out, double* times_out)
{
    int64_t N = 1;
    for (int d = 0; d < ndim; ++d) N *= shape[d];
    int64_t E = 0;
    for (int d = 0; d < ndim; ++d) if (shape[d] > 1) E += (N / shape[d]) * (shape[d] - 1);
    if (N >= (int64_t(1) << 31) - 1 || 2 * E >= (int64_t(1) << 31) - 1) return -1;

    auto t0 = std::chrono::steady_clock::now();
    GD* g = new GD((int)N, (int)(E > 0 ? E : 1), NULL);
    g->add_node((int)N);
    if (src && snk)
        for (int64_t v = 0; v < N; ++v) g->add_tweights((int)v, src[v], snk[v]);
    int64_t stride = N;
    for (int d = 0; d < ndim; ++d) {
        stride /= shape[d];
        if (shape[d] < 2 || !wf[d]) continue;
        const double* f = wf[d];
        const double* b = wb[d];
        const int64_t block = stride * shape[d];
        for (int64_t p = 0; p < N; ++p) {
            if ((p % block) / stride == shape[d] - 1) continue;
            if (use_sum_edge) g->sum_edge((int)p, (int)(p + stride), f[p], b[p]);
            else              g->add_edge((int)p, (int)(p + stride), f[p], b[p]);
        }
    }
    if (fg) for (int64_t v = 0; v < N; ++v) if (fg[v]) g->add_tweights((int)v, 65535.0, 0.0);
    if (bg) for (int64_t v = 0; v < N; ++v) if (bg[v]) g->add_tweights((int)v, 0.0, 65535.0);
    auto t1 = std::chrono::steady_clock::now();
    double flow = g->maxflow();
    auto t2 = std::chrono::steady_clock::now();
    if (mask_out)
        for (int64_t v = 0; v < N; ++v) mask_out[v] = (g->what_segment((int)v) == GD::SINK) ? 0 : 1;
    auto t3 = std::chrono::steady_clock::now();
    if (flow_out) *flow_out = flow;
    if (times_out) {
        times_out[0] = std::chrono::duration<double>(t1 - t0).count();
        times_out[1] = std::chrono::duration<double>(t2 - t1).count();
        times_out[2] = std::chrono::duration<double>(t3 - t2).count();
    }
    delete g;
    return 0;
}

// General sparse graph (region adjacency graphs of graph_from_labels, generate.py:177-338, and graphs assembled edge by
// edge through GCGraph, graph.py:382-498).  Replays, in the given order, n_tw add_tweights calls (graph.h:415-425)
// and m sum_edge calls (graph.h:456-480; the first call for a pair creates the arc pair, later ones accumulate), which
// is all the reference's Python layer ever does to a graph; t-link and n-link state are independent, so the relative
// order of the two groups does not matter.  mask_out: uint8[n], 1 unless what_segment == SINK.
int bkref_sparse_solve(int n, int64_t m, const int32_t* ei, const int32_t* ej, const double* cap, const double* rev,
                       int64_t n_tw, const int32_t* tw_node, const double* tw_src, const double* tw_snk,
                       uint8_t* mask_out, double* flow_out, double* maxflow_seconds)
{
    if (n < 1) return -1;
    GD* g = new GD(n, (int)(m > 0 ? m : 1), NULL);
    g->add_node(n);
    for (int64_t k = 0; k < n_tw; ++k) g->add_tweights(tw_node[k], tw_src[k], tw_snk[k]);
    for (int64_t k = 0; k < m; ++k) g->sum_edge(ei[k], ej[k], cap[k], rev[k]);
    auto t0 = std::chrono::steady_clock::now();
    const double flow = g->maxflow();
    auto t1 = std::chrono::steady_clock::now();
    if (mask_out)
        for (int v = 0; v < n; ++v) mask_out[v] = (g->what_segment(v) == GD::SINK) ? 0 : 1;
    if (flow_out) *flow_out = flow;
    if (maxflow_seconds) *maxflow_seconds = std::chrono::duration<double>(t1 - t0).count();
    delete g;
    return 0;
}

// Incremental handle API (used by tests that mirror GCGraph call by call).
void* bkref_new(int nodes, int edges) { GD* g = new GD(nodes, edges > 0 ? edges : 1, NULL); g->add_node(nodes); return g; }
void bkref_delete(void* h) { delete (GD*)h; }
void bkref_add_tweights(void* h, int i, double s, double t) { ((GD*)h)->add_tweights(i, s, t); }
void bkref_sum_edge(void* h, int i, int j, double c, double r) { ((GD*)h)->sum_edge(i, j, c, r); }
void bkref_add_edge(void* h, int i, int j, double c, double r) { ((GD*)h)->add_edge(i, j, c, r); }
double bkref_maxflow(void* h) { return ((GD*)h)->maxflow(); }
int bkref_what_segment(void* h, int i) { return (int)((GD*)h)->what_segment(i); }
double bkref_get_edge(void* h, int i, int j) { return ((GD*)h)->get_edge(i, j); }
double bkref_get_trcap(void* h, int i) { return ((GD*)h)->get_trcap(i); }

}  // extern "C"
