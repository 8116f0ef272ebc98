/* oracle/bk_lattice.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the reference's max-flow for the voxel path: the Boykov-Kolmogorov
 * augmenting-path algorithm of lib/maxflow/src/maxflow.cpp, re-expressed for an IMPLICIT
 * 2*ndim-connected lattice with dense structure-of-arrays residual capacities (no node/arc
 * objects, no pointers: arc (v,k) leaves voxel v towards offset off[k], its sister is (v+off[k], k^1)).
 * Each function names the reference lines it follows.  Arcs of a voxel are visited in the order the
 * reference's adjacency list yields for a lattice built by energy_voxel.__skeleton_base
 * (last inserted first: axis ndim-1 "+", axis ndim-1 "-", ..., axis 0 "+", axis 0 "-"; see
 * graph.h:427-454 where add_edge pushes at the list head, and energy_voxel.py:637-664 for the
 * insertion order), so the augmentation order -- and therefore the returned flow, bit for bit --
 * equals that of the reference solver on the same float64 capacities.
 *
 * Parity status: PINNED by tests/test_oracle.py against oracle/_ref (the real reference solver,
 * compiled from /root/reference) and against the reference's own fixtures
 * (tests/graphcut_/cut.py:32-50 maxflow==3; tests/graphcut_/energy_voxel.py:55-149 masks).
 *
 * Only tests/, bench.py's cpu_baseline / --impl reference arm and __graft_entry__.smoke() may
 * load this; the product library never links or calls it.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define BK_MAXDIM 8
#define P_NONE (-1)      /* maxflow.cpp: parent == NULL  */
#define P_TERMINAL 100   /* maxflow.cpp:11  TERMINAL     */
#define P_ORPHAN 101     /* maxflow.cpp:12  ORPHAN       */
#define INFINITE_D 0x7fffffff /* maxflow.cpp:15 */

typedef struct {
    int ndim, K;
    int64_t N;
    int64_t off[2 * BK_MAXDIM];
    double *rcap;      /* [K][N] residual capacity of arc (v,k)                 */
    double *tr;        /* [N] tr_cap: >0 residual source link, <0 sink link     */
    uint8_t *amask;    /* [N] bit k set iff arc (v,k) exists                    */
    int8_t *parent;    /* [N] arc index towards the parent, or P_*              */
    uint8_t *is_sink;  /* [N]                                                   */
    int64_t *ts;       /* [N] time stamp                                        */
    int32_t *dist;     /* [N] distance to terminal                              */
    int64_t *next;     /* [N] active list link; -1 = not in list; self = last   */
    int64_t qfirst[2], qlast[2];
    /* orphan list (maxflow.cpp:80-101): singly linked cells from a growable pool */
    int64_t *o_node, *o_next;
    int64_t o_cap, o_free, o_first, o_last;
    int64_t TIME;
    double flow;
} bk_t;

#define RCAP(g, v, k) ((g)->rcap[(int64_t)(k) * (g)->N + (v)])

/* maxflow.cpp:33-44 set_active */
static void set_active(bk_t *g, int64_t i)
{
    if (g->next[i] < 0) {
        if (g->qlast[1] >= 0) g->next[g->qlast[1]] = i;
        else g->qfirst[1] = i;
        g->qlast[1] = i;
        g->next[i] = i;
    }
}

/* maxflow.cpp:51-75 next_active */
static int64_t next_active(bk_t *g)
{
    for (;;) {
        int64_t i = g->qfirst[0];
        if (i < 0) {
            g->qfirst[0] = i = g->qfirst[1];
            g->qlast[0] = g->qlast[1];
            g->qfirst[1] = g->qlast[1] = -1;
            if (i < 0) return -1;
        }
        if (g->next[i] == i) g->qfirst[0] = g->qlast[0] = -1;
        else g->qfirst[0] = g->next[i];
        g->next[i] = -1;
        if (g->parent[i] != P_NONE) return i; /* active iff it has a parent */
    }
}

static int64_t o_alloc(bk_t *g)
{
    if (g->o_free < 0) {
        int64_t ncap = g->o_cap ? g->o_cap * 2 : 1024;
        g->o_node = (int64_t *)realloc(g->o_node, (size_t)ncap * sizeof(int64_t));
        g->o_next = (int64_t *)realloc(g->o_next, (size_t)ncap * sizeof(int64_t));
        for (int64_t c = g->o_cap; c < ncap; ++c) g->o_next[c] = (c + 1 < ncap) ? c + 1 : -1;
        g->o_free = g->o_cap;
        g->o_cap = ncap;
    }
    int64_t c = g->o_free;
    g->o_free = g->o_next[c];
    return c;
}

/* maxflow.cpp:80-88 set_orphan_front */
static void set_orphan_front(bk_t *g, int64_t i)
{
    g->parent[i] = P_ORPHAN;
    int64_t c = o_alloc(g);
    g->o_node[c] = i;
    g->o_next[c] = g->o_first;
    g->o_first = c;
}

/* maxflow.cpp:90-101 set_orphan_rear */
static void set_orphan_rear(bk_t *g, int64_t i)
{
    g->parent[i] = P_ORPHAN;
    int64_t c = o_alloc(g);
    g->o_node[c] = i;
    if (g->o_last >= 0) g->o_next[g->o_last] = c;
    else g->o_first = c;
    g->o_last = c;
    g->o_next[c] = -1;
}

/* maxflow.cpp:118-156 maxflow_init */
static void maxflow_init(bk_t *g)
{
    g->qfirst[0] = g->qlast[0] = g->qfirst[1] = g->qlast[1] = -1;
    g->o_first = -1;
    g->o_last = -1;
    g->TIME = 0;
    for (int64_t i = 0; i < g->N; ++i) {
        g->next[i] = -1;
        g->ts[i] = 0;
        if (g->tr[i] > 0) {
            g->is_sink[i] = 0; g->parent[i] = P_TERMINAL; set_active(g, i); g->dist[i] = 1;
        } else if (g->tr[i] < 0) {
            g->is_sink[i] = 1; g->parent[i] = P_TERMINAL; set_active(g, i); g->dist[i] = 1;
        } else {
            g->parent[i] = P_NONE;
        }
    }
}

/* maxflow.cpp:243-311 augment.  The middle arc is (u,k): u in the source tree, head in the sink tree. */
static void augment(bk_t *g, int64_t u, int k)
{
    const int64_t v = u + g->off[k];
    int64_t i;
    int a;
    double bottleneck = RCAP(g, u, k);
    /* 1a: source tree -- the residual that matters is parent->child, i.e. the sister of the parent arc */
    for (i = u;; i += g->off[a]) {
        a = g->parent[i];
        if (a == P_TERMINAL) break;
        double c = RCAP(g, i + g->off[a], a ^ 1);
        if (bottleneck > c) bottleneck = c;
    }
    if (bottleneck > g->tr[i]) bottleneck = g->tr[i];
    /* 1b: sink tree -- child->parent arcs */
    for (i = v;; i += g->off[a]) {
        a = g->parent[i];
        if (a == P_TERMINAL) break;
        double c = RCAP(g, i, a);
        if (bottleneck > c) bottleneck = c;
    }
    if (bottleneck > -g->tr[i]) bottleneck = -g->tr[i];

    /* 2a */
    RCAP(g, v, k ^ 1) += bottleneck;
    RCAP(g, u, k) -= bottleneck;
    for (i = u;;) {
        a = g->parent[i];
        if (a == P_TERMINAL) break;
        int64_t p = i + g->off[a];
        RCAP(g, i, a) += bottleneck;
        RCAP(g, p, a ^ 1) -= bottleneck;
        int sat = !RCAP(g, p, a ^ 1);
        int64_t cur = i;
        i = p; /* advance before set_orphan_front overwrites parent[cur] */
        if (sat) set_orphan_front(g, cur);
    }
    g->tr[i] -= bottleneck;
    if (!g->tr[i]) set_orphan_front(g, i);
    /* 2b */
    for (i = v;;) {
        a = g->parent[i];
        if (a == P_TERMINAL) break;
        int64_t p = i + g->off[a];
        RCAP(g, p, a ^ 1) += bottleneck;
        RCAP(g, i, a) -= bottleneck;
        int sat = !RCAP(g, i, a);
        int64_t cur = i;
        i = p;
        if (sat) set_orphan_front(g, cur);
    }
    g->tr[i] += bottleneck;
    if (!g->tr[i]) set_orphan_front(g, i);

    g->flow += bottleneck;
}

/* maxflow.cpp:315-390 (source) and :392-467 (sink) -- one routine, the tree selected by `sink`. */
static void process_orphan(bk_t *g, int64_t i, int sink)
{
    int a0_min = -1;
    int32_t d_min = INFINITE_D;
    const uint8_t am = g->amask[i];
    for (int k = g->K - 1; k >= 0; --k) {
        if (!(am & (1u << k))) continue;
        int64_t j = i + g->off[k];
        /* source orphan: arc parent->child = sister of (i,k); sink orphan: arc (i,k) itself */
        double c = sink ? RCAP(g, i, k) : RCAP(g, j, k ^ 1);
        if (!c) continue;
        if (g->is_sink[j] != sink || g->parent[j] == P_NONE) continue;
        /* checking the origin of j */
        int32_t d = 0;
        int64_t jj = j;
        for (;;) {
            if (g->ts[jj] == g->TIME) { d += g->dist[jj]; break; }
            int a = g->parent[jj];
            d++;
            if (a == P_TERMINAL) { g->ts[jj] = g->TIME; g->dist[jj] = 1; break; }
            if (a == P_ORPHAN) { d = INFINITE_D; break; }
            jj += g->off[a];
        }
        if (d < INFINITE_D) {
            if (d < d_min) { a0_min = k; d_min = d; }
            for (jj = j; g->ts[jj] != g->TIME; jj += g->off[g->parent[jj]]) {
                g->ts[jj] = g->TIME;
                g->dist[jj] = d--;
            }
        }
    }
    if (a0_min >= 0) {
        g->parent[i] = (int8_t)a0_min;
        g->ts[i] = g->TIME;
        g->dist[i] = d_min + 1;
    } else {
        g->parent[i] = P_NONE;
        for (int k = g->K - 1; k >= 0; --k) {
            if (!(am & (1u << k))) continue;
            int64_t j = i + g->off[k];
            int a = g->parent[j];
            if (g->is_sink[j] == sink && a != P_NONE) {
                double c = sink ? RCAP(g, i, k) : RCAP(g, j, k ^ 1);
                if (c) set_active(g, j);
                if (a != P_TERMINAL && a != P_ORPHAN && j + g->off[a] == i) set_orphan_rear(g, j);
            }
        }
    }
}

/* maxflow.cpp:471-604 maxflow (reuse_trees == false path) */
static double run_maxflow(bk_t *g)
{
    int64_t current = -1;
    maxflow_init(g);
    for (;;) {
        int64_t i = current;
        if (i >= 0) {
            g->next[i] = -1; /* remove active flag */
            if (g->parent[i] == P_NONE) i = -1;
        }
        if (i < 0) {
            i = next_active(g);
            if (i < 0) break;
        }
        /* growth */
        int found_k = -1;
        int64_t found_u = -1;
        const uint8_t am = g->amask[i];
        if (!g->is_sink[i]) {
            for (int k = g->K - 1; k >= 0; --k) {
                if (!(am & (1u << k))) continue;
                if (!RCAP(g, i, k)) continue;
                int64_t j = i + g->off[k];
                if (g->parent[j] == P_NONE) {
                    g->is_sink[j] = 0; g->parent[j] = (int8_t)(k ^ 1);
                    g->ts[j] = g->ts[i]; g->dist[j] = g->dist[i] + 1;
                    set_active(g, j);
                } else if (g->is_sink[j]) {
                    found_u = i; found_k = k; break;
                } else if (g->ts[j] <= g->ts[i] && g->dist[j] > g->dist[i]) {
                    g->parent[j] = (int8_t)(k ^ 1);
                    g->ts[j] = g->ts[i]; g->dist[j] = g->dist[i] + 1;
                }
            }
        } else {
            for (int k = g->K - 1; k >= 0; --k) {
                if (!(am & (1u << k))) continue;
                int64_t j = i + g->off[k];
                if (!RCAP(g, j, k ^ 1)) continue;
                if (g->parent[j] == P_NONE) {
                    g->is_sink[j] = 1; g->parent[j] = (int8_t)(k ^ 1);
                    g->ts[j] = g->ts[i]; g->dist[j] = g->dist[i] + 1;
                    set_active(g, j);
                } else if (!g->is_sink[j]) {
                    found_u = j; found_k = k ^ 1; break; /* middle arc is the sister j -> i */
                } else if (g->ts[j] <= g->ts[i] && g->dist[j] > g->dist[i]) {
                    g->parent[j] = (int8_t)(k ^ 1);
                    g->ts[j] = g->ts[i]; g->dist[j] = g->dist[i] + 1;
                }
            }
        }
        g->TIME++;
        if (found_k >= 0) {
            g->next[i] = i; /* set active flag */
            current = i;
            augment(g, found_u, found_k);
            /* adoption (maxflow.cpp:572-590) */
            while (g->o_first >= 0) {
                int64_t np = g->o_first;
                int64_t np_next = g->o_next[np];
                g->o_next[np] = -1;
                while ((np = g->o_first) >= 0) {
                    g->o_first = g->o_next[np];
                    int64_t n = g->o_node[np];
                    g->o_next[np] = g->o_free; g->o_free = np; /* Delete(np) */
                    if (g->o_first < 0) g->o_last = -1;
                    process_orphan(g, n, g->is_sink[n]);
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

/* Solve a lattice instance.
 *   wf[d] / wb[d]: N doubles per axis (capacity p -> p+stride_d / p+stride_d -> p; last plane ignored),
 *                  or NULL to leave axis d without arcs.
 *   tr:            N doubles, the net terminal capacity after all add_tweights calls (graph.h:415-425).
 *   flow_const:    the minima accumulated by those add_tweights calls (graph.h:423).
 *   mask_out:      uint8[N]; 1 unless the voxel ends in the sink tree (graph.h:560-571 + CLI :178-181).
 *   times_out[2]:  seconds {setup, maxflow}.
 * Returns 0 on success, -1 on bad arguments / allocation failure. */
int bk_lattice_solve(int ndim, const int64_t *shape, const double *const *wf, const double *const *wb,
                     const double *tr, double flow_const, uint8_t *mask_out, double *flow_out,
                     double *times_out)
{
    if (ndim < 1 || ndim > 4) return -1;
    double t0 = now_s();
    bk_t g;
    memset(&g, 0, sizeof g);
    g.ndim = ndim;
    g.K = 2 * ndim;
    g.N = 1;
    for (int d = 0; d < ndim; ++d) g.N *= shape[d];
    const int64_t N = g.N;
    int64_t stride[BK_MAXDIM];
    {
        int64_t s = N;
        for (int d = 0; d < ndim; ++d) { s /= shape[d]; stride[d] = s; g.off[2 * d] = -s; g.off[2 * d + 1] = s; }
    }
    g.rcap = (double *)calloc((size_t)(g.K * N), sizeof(double));
    g.tr = (double *)malloc((size_t)N * sizeof(double));
    g.amask = (uint8_t *)calloc((size_t)N, 1);
    g.parent = (int8_t *)malloc((size_t)N);
    g.is_sink = (uint8_t *)calloc((size_t)N, 1);
    g.ts = (int64_t *)calloc((size_t)N, sizeof(int64_t));
    g.dist = (int32_t *)calloc((size_t)N, sizeof(int32_t));
    g.next = (int64_t *)malloc((size_t)N * sizeof(int64_t));
    g.o_free = -1;
    if (!g.rcap || !g.tr || !g.amask || !g.parent || !g.is_sink || !g.ts || !g.dist || !g.next) return -1;
    memcpy(g.tr, tr, (size_t)N * sizeof(double));
    for (int d = 0; d < ndim; ++d) {
        if (shape[d] < 2 || !wf || !wf[d]) continue;
        const int64_t s = stride[d], block = s * shape[d];
        for (int64_t p = 0; p < N; ++p) {
            if ((p % block) / s == shape[d] - 1) continue;
            RCAP(&g, p, 2 * d + 1) = wf[d][p];
            RCAP(&g, p + s, 2 * d) = wb[d][p];
            g.amask[p] |= (uint8_t)(1u << (2 * d + 1));
            g.amask[p + s] |= (uint8_t)(1u << (2 * d));
        }
    }
    g.flow = flow_const;
    double t1 = now_s();
    double flow = run_maxflow(&g);
    double t2 = now_s();
    if (mask_out)
        for (int64_t v = 0; v < N; ++v)
            mask_out[v] = (g.parent[v] != P_NONE && g.is_sink[v]) ? 0 : 1;
    if (flow_out) *flow_out = flow;
    if (times_out) { times_out[0] = t1 - t0; times_out[1] = t2 - t1; }
    free(g.rcap); free(g.tr); free(g.amask); free(g.parent); free(g.is_sink);
    free(g.ts); free(g.dist); free(g.next); free(g.o_node); free(g.o_next);
    return 0;
}
