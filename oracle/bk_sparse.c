/* oracle/bk_sparse.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the reference's graph assembly and max-flow for GENERAL graphs: the region adjacency
 * graphs of graph_from_labels (medpy/graphcut/generate.py:177-338) and graphs built call by call through GCGraph
 * (graph.py:382-498).  Same algorithm as oracle/bk_lattice.c (Boykov-Kolmogorov, lib/maxflow/src/maxflow.cpp), here
 * on explicit arc arrays:
 *   add_tweights  graph.h:415-425        sum_edge / add_edge / get_arc  graph.h:427-509
 *   maxflow_init  maxflow.cpp:118-156    augment  :243-311    process_*_orphan  :315-467    maxflow  :471-604
 * Arcs hang off their tail node in a singly linked list with the newest arc first (add_edge pushes at the head,
 * graph.h:443-446) and every loop walks that list front to back, so the order of growth, augmentation and adoption --
 * and with it the returned flow, bit for bit -- equals the reference solver's for the same call sequence.
 *
 * Parity status: PINNED by tests/test_oracle_labels.py against oracle/_ref (the real reference solver compiled from
 * /root/reference) on random graphs and on the golden region graphs of tests/golden/golden_labels_v1.npz.
 *
 * Only tests/, bench.py's cpu_baseline leg and __graft_entry__.smoke() may load this.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define A_NONE (-1)      /* parent == NULL            */
#define A_TERMINAL (-2)  /* maxflow.cpp:11 TERMINAL   */
#define A_ORPHAN (-3)    /* maxflow.cpp:12 ORPHAN     */
#define INFINITE_D 0x7fffffff

typedef struct {
    int n;
    int64_t n_arcs, arc_cap;
    int32_t *first;    /* [n] newest arc of the node, -1 if none   */
    int32_t *head;     /* [arc] node the arc points to             */
    int32_t *anext;    /* [arc] next (older) arc of the same tail  */
    double *rcap;      /* [arc] residual capacity; sister = arc^1  */
    double *tr;        /* [n] tr_cap                               */
    int32_t *parent;   /* [n] arc towards the parent, or A_*       */
    uint8_t *is_sink;
    int64_t *ts;
    int32_t *dist;
    int32_t *qnext;    /* active list link: -1 not listed, self = last */
    int32_t qfirst[2], qlast[2];
    int32_t *o_node;
    int64_t *o_next;
    int64_t o_cap, o_free, o_first, o_last;
    int64_t TIME;
    double flow;
} g_t;

/* ---- assembly ------------------------------------------------------------------------------------------------ */
static void add_tweights(g_t *g, int i, double s, double t)       /* graph.h:415-425 */
{
    double delta = g->tr[i];
    if (delta > 0) s += delta; else t -= delta;
    g->flow += (s < t) ? s : t;
    g->tr[i] = s - t;
}

static int add_edge(g_t *g, int i, int j, double cap, double rev)  /* graph.h:427-454 */
{
    if (g->n_arcs + 2 > g->arc_cap) {
        int64_t nc = g->arc_cap ? 2 * g->arc_cap : 1024;
        g->head = (int32_t *)realloc(g->head, (size_t)nc * sizeof(int32_t));
        g->anext = (int32_t *)realloc(g->anext, (size_t)nc * sizeof(int32_t));
        g->rcap = (double *)realloc(g->rcap, (size_t)nc * sizeof(double));
        if (!g->head || !g->anext || !g->rcap) return -1;
        g->arc_cap = nc;
    }
    const int32_t a = (int32_t)g->n_arcs, ar = a + 1;      /* sisters are consecutive: sister(x) == x ^ 1 */
    g->n_arcs += 2;
    g->anext[a] = g->first[i]; g->first[i] = a;
    g->anext[ar] = g->first[j]; g->first[j] = ar;
    g->head[a] = j; g->head[ar] = i;
    g->rcap[a] = cap; g->rcap[ar] = rev;
    return 0;
}

static int sum_edge(g_t *g, int i, int j, double cap, double rev)  /* graph.h:456-480 with get_arc :499-509 */
{
    for (int32_t a = g->first[i]; a >= 0; a = g->anext[a])
        if (g->head[a] == j) { g->rcap[a] += cap; g->rcap[a ^ 1] += rev; return 0; }
    return add_edge(g, i, j, cap, rev);
}

/* ---- active / orphan lists (maxflow.cpp:33-101) ------------------------------------------------------------------ */
static void set_active(g_t *g, int32_t i)
{
    if (g->qnext[i] < 0) {
        if (g->qlast[1] >= 0) g->qnext[g->qlast[1]] = i; else g->qfirst[1] = i;
        g->qlast[1] = i;
        g->qnext[i] = i;
    }
}

static int32_t next_active(g_t *g)
{
    for (;;) {
        int32_t i = g->qfirst[0];
        if (i < 0) {
            g->qfirst[0] = i = g->qfirst[1];
            g->qlast[0] = g->qlast[1];
            g->qfirst[1] = g->qlast[1] = -1;
            if (i < 0) return -1;
        }
        if (g->qnext[i] == i) g->qfirst[0] = g->qlast[0] = -1; else g->qfirst[0] = g->qnext[i];
        g->qnext[i] = -1;
        if (g->parent[i] != A_NONE) return i;
    }
}

static int64_t o_alloc(g_t *g)
{
    if (g->o_free < 0) {
        int64_t nc = g->o_cap ? g->o_cap * 2 : 1024;
        g->o_node = (int32_t *)realloc(g->o_node, (size_t)nc * sizeof(int32_t));
        g->o_next = (int64_t *)realloc(g->o_next, (size_t)nc * sizeof(int64_t));
        for (int64_t c = g->o_cap; c < nc; ++c) g->o_next[c] = (c + 1 < nc) ? c + 1 : -1;
        g->o_free = g->o_cap;
        g->o_cap = nc;
    }
    int64_t c = g->o_free;
    g->o_free = g->o_next[c];
    return c;
}

static void set_orphan_front(g_t *g, int32_t i)
{
    g->parent[i] = A_ORPHAN;
    int64_t c = o_alloc(g);
    g->o_node[c] = i;
    g->o_next[c] = g->o_first;
    g->o_first = c;
}

static void set_orphan_rear(g_t *g, int32_t i)
{
    g->parent[i] = A_ORPHAN;
    int64_t c = o_alloc(g);
    g->o_node[c] = i;
    if (g->o_last >= 0) g->o_next[g->o_last] = c; else g->o_first = c;
    g->o_last = c;
    g->o_next[c] = -1;
}

/* ---- max-flow ----------------------------------------------------------------------------------------------------- */
static void maxflow_init(g_t *g)                                    /* maxflow.cpp:118-156 */
{
    g->qfirst[0] = g->qlast[0] = g->qfirst[1] = g->qlast[1] = -1;
    g->o_first = g->o_last = -1;
    g->TIME = 0;
    for (int32_t i = 0; i < g->n; ++i) {
        g->qnext[i] = -1;
        g->ts[i] = 0;
        if (g->tr[i] > 0) { g->is_sink[i] = 0; g->parent[i] = A_TERMINAL; set_active(g, i); g->dist[i] = 1; }
        else if (g->tr[i] < 0) { g->is_sink[i] = 1; g->parent[i] = A_TERMINAL; set_active(g, i); g->dist[i] = 1; }
        else g->parent[i] = A_NONE;
    }
}

/* maxflow.cpp:243-311; `mid` runs from the source tree (its tail) to the sink tree (its head) */
static void augment(g_t *g, int32_t mid)
{
    const int32_t u = g->head[mid ^ 1], v = g->head[mid];
    int32_t i, a;
    double bottleneck = g->rcap[mid];
    for (i = u;; i = g->head[a]) {                 /* 1a: source tree, residual parent -> child = sister of the parent arc */
        a = g->parent[i];
        if (a == A_TERMINAL) break;
        if (bottleneck > g->rcap[a ^ 1]) bottleneck = g->rcap[a ^ 1];
    }
    if (bottleneck > g->tr[i]) bottleneck = g->tr[i];
    for (i = v;; i = g->head[a]) {                 /* 1b: sink tree, child -> parent arcs */
        a = g->parent[i];
        if (a == A_TERMINAL) break;
        if (bottleneck > g->rcap[a]) bottleneck = g->rcap[a];
    }
    if (bottleneck > -g->tr[i]) bottleneck = -g->tr[i];

    g->rcap[mid ^ 1] += bottleneck;                /* 2a */
    g->rcap[mid] -= bottleneck;
    for (i = u;;) {
        a = g->parent[i];
        if (a == A_TERMINAL) break;
        g->rcap[a] += bottleneck;
        g->rcap[a ^ 1] -= bottleneck;
        const int32_t cur = i;
        i = g->head[a];                            /* advance before set_orphan_front overwrites parent[cur] */
        if (!g->rcap[a ^ 1]) set_orphan_front(g, cur);
    }
    g->tr[i] -= bottleneck;
    if (!g->tr[i]) set_orphan_front(g, i);
    for (i = v;;) {                                /* 2b */
        a = g->parent[i];
        if (a == A_TERMINAL) break;
        g->rcap[a ^ 1] += bottleneck;
        g->rcap[a] -= bottleneck;
        const int32_t cur = i;
        i = g->head[a];
        if (!g->rcap[a]) set_orphan_front(g, cur);
    }
    g->tr[i] += bottleneck;
    if (!g->tr[i]) set_orphan_front(g, i);
    g->flow += bottleneck;
}

/* maxflow.cpp:315-390 (source tree) and :392-467 (sink tree) */
static void process_orphan(g_t *g, int32_t i, int sink)
{
    int32_t a0_min = A_NONE;
    int32_t d_min = INFINITE_D;
    for (int32_t a0 = g->first[i]; a0 >= 0; a0 = g->anext[a0]) {
        const double c = sink ? g->rcap[a0] : g->rcap[a0 ^ 1];
        if (!c) continue;
        const int32_t j = g->head[a0];
        if (g->is_sink[j] != sink || g->parent[j] == A_NONE) continue;
        int32_t d = 0, jj = j;
        for (;;) {                                 /* checking the origin of j */
            if (g->ts[jj] == g->TIME) { d += g->dist[jj]; break; }
            const int32_t a = g->parent[jj];
            d++;
            if (a == A_TERMINAL) { g->ts[jj] = g->TIME; g->dist[jj] = 1; break; }
            if (a == A_ORPHAN) { d = INFINITE_D; break; }
            jj = g->head[a];
        }
        if (d < INFINITE_D) {
            if (d < d_min) { a0_min = a0; d_min = d; }
            for (jj = j; g->ts[jj] != g->TIME; jj = g->head[g->parent[jj]]) { g->ts[jj] = g->TIME; g->dist[jj] = d--; }
        }
    }
    if (a0_min != A_NONE) {
        g->parent[i] = a0_min;
        g->ts[i] = g->TIME;
        g->dist[i] = d_min + 1;
    } else {
        g->parent[i] = A_NONE;
        for (int32_t a0 = g->first[i]; a0 >= 0; a0 = g->anext[a0]) {
            const int32_t j = g->head[a0];
            const int32_t a = g->parent[j];
            if (g->is_sink[j] == sink && a != A_NONE) {
                const double c = sink ? g->rcap[a0] : g->rcap[a0 ^ 1];
                if (c) set_active(g, j);
                if (a != A_TERMINAL && a != A_ORPHAN && g->head[a] == i) set_orphan_rear(g, j);
            }
        }
    }
}

static double run_maxflow(g_t *g)                                   /* maxflow.cpp:471-604, reuse_trees == false */
{
    int32_t current = -1;
    maxflow_init(g);
    for (;;) {
        int32_t i = current;
        if (i >= 0) {
            g->qnext[i] = -1;
            if (g->parent[i] == A_NONE) i = -1;
        }
        if (i < 0) {
            i = next_active(g);
            if (i < 0) break;
        }
        int32_t mid = A_NONE;
        if (!g->is_sink[i]) {                      /* grow the source tree */
            for (int32_t a = g->first[i]; a >= 0; a = g->anext[a]) {
                if (!g->rcap[a]) continue;
                const int32_t j = g->head[a];
                if (g->parent[j] == A_NONE) {
                    g->is_sink[j] = 0; g->parent[j] = a ^ 1; g->ts[j] = g->ts[i]; g->dist[j] = g->dist[i] + 1;
                    set_active(g, j);
                } else if (g->is_sink[j]) {
                    mid = a; break;
                } else if (g->ts[j] <= g->ts[i] && g->dist[j] > g->dist[i]) {
                    g->parent[j] = a ^ 1; g->ts[j] = g->ts[i]; g->dist[j] = g->dist[i] + 1;
                }
            }
        } else {                                   /* grow the sink tree */
            for (int32_t a = g->first[i]; a >= 0; a = g->anext[a]) {
                if (!g->rcap[a ^ 1]) continue;
                const int32_t j = g->head[a];
                if (g->parent[j] == A_NONE) {
                    g->is_sink[j] = 1; g->parent[j] = a ^ 1; g->ts[j] = g->ts[i]; g->dist[j] = g->dist[i] + 1;
                    set_active(g, j);
                } else if (!g->is_sink[j]) {
                    mid = a ^ 1; break;            /* the middle arc is the sister j -> i */
                } else if (g->ts[j] <= g->ts[i] && g->dist[j] > g->dist[i]) {
                    g->parent[j] = a ^ 1; g->ts[j] = g->ts[i]; g->dist[j] = g->dist[i] + 1;
                }
            }
        }
        g->TIME++;
        if (mid != A_NONE) {
            g->qnext[i] = i;                       /* set active flag */
            current = i;
            augment(g, mid);
            while (g->o_first >= 0) {              /* adoption, maxflow.cpp:572-590 */
                int64_t np = g->o_first;
                const int64_t np_next = g->o_next[np];
                g->o_next[np] = -1;
                while ((np = g->o_first) >= 0) {
                    g->o_first = g->o_next[np];
                    const int32_t nd = g->o_node[np];
                    g->o_next[np] = g->o_free; g->o_free = np;
                    if (g->o_first < 0) g->o_last = -1;
                    process_orphan(g, nd, g->is_sink[nd]);
                }
                g->o_first = np_next;
            }
        } else {
            current = -1;
        }
    }
    return g->flow;
}

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* Same contract as bkref_sparse_solve (oracle/ref_driver.cpp): n_tw add_tweights calls, then m sum_edge calls, in
 * order; mask_out[v] = 1 unless what_segment(v) == SINK (graph.h:560-571). */
int bk_sparse_solve(int n, int64_t m, const int32_t *ei, const int32_t *ej, const double *cap, const double *rev,
                    int64_t n_tw, const int32_t *tw_node, const double *tw_src, const double *tw_snk,
                    uint8_t *mask_out, double *flow_out, double *maxflow_seconds)
{
    if (n < 1) return -1;
    g_t g;
    memset(&g, 0, sizeof g);
    g.n = n;
    g.first = (int32_t *)malloc((size_t)n * sizeof(int32_t));
    g.tr = (double *)calloc((size_t)n, sizeof(double));
    g.parent = (int32_t *)malloc((size_t)n * sizeof(int32_t));
    g.is_sink = (uint8_t *)calloc((size_t)n, 1);
    g.ts = (int64_t *)calloc((size_t)n, sizeof(int64_t));
    g.dist = (int32_t *)calloc((size_t)n, sizeof(int32_t));
    g.qnext = (int32_t *)malloc((size_t)n * sizeof(int32_t));
    g.o_free = -1;
    if (!g.first || !g.tr || !g.parent || !g.is_sink || !g.ts || !g.dist || !g.qnext) return -1;
    for (int i = 0; i < n; ++i) g.first[i] = -1;
    for (int64_t k = 0; k < n_tw; ++k) {
        if (tw_node[k] < 0 || tw_node[k] >= n) return -2;
        add_tweights(&g, tw_node[k], tw_src[k], tw_snk[k]);
    }
    for (int64_t k = 0; k < m; ++k) {
        if (ei[k] < 0 || ej[k] < 0 || ei[k] >= n || ej[k] >= n || ei[k] == ej[k]) return -2;
        if (sum_edge(&g, ei[k], ej[k], cap[k], rev[k])) return -1;
    }
    const double t0 = now_s();
    const double flow = run_maxflow(&g);
    const double t1 = now_s();
    if (mask_out)
        for (int v = 0; v < n; ++v) mask_out[v] = (g.parent[v] != A_NONE && g.is_sink[v]) ? 0 : 1;
    if (flow_out) *flow_out = flow;
    if (maxflow_seconds) *maxflow_seconds = t1 - t0;
    free(g.first); free(g.head); free(g.anext); free(g.rcap); free(g.tr); free(g.parent); free(g.is_sink);
    free(g.ts); free(g.dist); free(g.qnext); free(g.o_node); free(g.o_next);
    return 0;
}
