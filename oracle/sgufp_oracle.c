/*
 * sgufp_oracle.c — ORACLE B.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded restatement of the reference's scenario-cut routine
 *   GuroSolver::solveSubProblem(path)            /root/reference/grb.cpp:139-159
 *   GuroSolver::solveSubProblem(y) scenario loop /root/reference/grb.cpp:162-360
 * over the instance model of /root/reference/Network.cpp:10-186 and the result format of
 * /root/reference/Cut.h:342-344,406-421.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file.  The product (sgufp_solver_b200/) never links, imports or executes it.
 *
 * PARITY UNPINNED for the LP half: the reference delegates the LP arithmetic to Gurobi 11.0.3
 * (grb.cpp:231-235,241-278,304-344; CMakeLists.txt:33-41), which is proprietary and absent from
 * /root/reference and from this image, and none of the reference's tests pins a cut value.  The
 * dual LP is degenerate, so "the reference's cut" is whatever vertex Gurobi's pivoting reaches.
 * This oracle therefore fixes ONE optimal dual solution by a written-down rule (SPEC-LP, see
 * DESIGN.md §3) and is pinned instead by solver-independent checks against HiGHS
 * (tests/test_oracle_lp.py): per-scenario status, per-scenario optimal objective, feasibility of
 * the lifted dual in every row of grb.cpp:49-123, and RHS + sum coef*y == mean objective.
 *
 * SPEC-LP in one paragraph.  With y fixed the primal (StochasticModel.h:103-197) is a
 * maximum-reward flow in which every matched (in-arc, out-arc) pair at a V-bar node carries
 * equal flow and every unmatched arc at a V-bar node is closed.  Arcs are therefore grouped into
 * CHAINS through matched pairs; a chain is an arc of the CONTRACTED graph whose nodes are the
 * remaining nodes, with all conservation-free nodes (no in-arcs, no out-arcs, node 0, node n-1:
 * grb.cpp:134-135) merged into one ROOT of potential 0.  Any optimal flow x* is found (here:
 * successive shortest paths).  The node potentials are then the algorithm-independent extreme
 * point "shortest residual distance from the root" (unique for every optimal x*), completed for
 * nodes the root cannot reach by the least labels consistent with the labelled ones, and by a
 * zero-rooted completion for nodes cut off both ways.  Inside a chain the capacity multiplier
 * sits on the FIRST arc of least capacity and the lower-bound multiplier on the LAST arc of
 * greatest lower bound; it is booked as sigma if that arc enters a V-bar node, else phi if it
 * leaves one, else gamma (the strongest valid cut).  lambda-mu of a matched pair is its wire
 * potential minus alpha_q, alpha_q being the wire potential of q's first matched pair.
 * Everything is an integer until the reference's own fold `(cap / S) * dual` (grb.cpp:241-278).
 */
#include <limits.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define INF_D (INT_MAX / 4)

typedef struct OrcNet {
    int n, m, S;
    int *tail, *head, *rew; /* rew = rewards[0] (grb.cpp:53,71,89) */
    int *u, *l;             /* [m][S], arc-major as NetworkArc stores them (Network.h:32-34) */
    int *in_ptr, *in_arc, *out_ptr, *out_arc; /* per-node arc lists in file order (Network.cpp:45-48) */
    int *out_index;                           /* position of arc b in outgoingArcs(tail(b)) */
    int nvbar, *vbar;                         /* after shuffleVBarNodes (Network.cpp:132-186) */
    unsigned char *is_vbar, *active, *is_root;
    int L, *layer_arc, *arc_layer; /* processingOrder (Network.cpp:94-118) */
    int T, *slot_base, *slot_in, *slot_out, *slot_sorted;
    int nc, *cn; /* contracted node ids, 0 = root */
    /* per-path plan */
    int nch, *ch_ptr, *ch_arcs, *ch_sv, *ch_ev, *ch_qs, *ch_qe, *ch_r;
    int *arc_chain, *arc_pos, *match_out, *match_in;
} OrcNet;

static void *xcalloc(size_t n, size_t sz) {
    void *p = calloc(n ? n : 1, sz);
    if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
    return p;
}

void orc_free(OrcNet *N) {
    if (!N) return;
    free(N->tail); free(N->head); free(N->rew); free(N->u); free(N->l);
    free(N->in_ptr); free(N->in_arc); free(N->out_ptr); free(N->out_arc); free(N->out_index);
    free(N->vbar); free(N->is_vbar); free(N->active); free(N->is_root);
    free(N->layer_arc); free(N->arc_layer); free(N->slot_base); free(N->slot_in); free(N->slot_out);
    free(N->slot_sorted); free(N->cn);
    free(N->ch_ptr); free(N->ch_arcs); free(N->ch_sv); free(N->ch_ev); free(N->ch_qs); free(N->ch_qe);
    free(N->ch_r); free(N->arc_chain); free(N->arc_pos); free(N->match_out); free(N->match_in);
    free(N);
}

/* Network::shuffleVBarNodes (Network.cpp:132-186): demand points that are in V-bar first, then
 * breadth-first over parents.  Duplicates in the reference's parentsVec never change the order of
 * first occurrence, so they are dropped here.  Returns -1 where the reference would spin forever. */
static int shuffle_vbar(OrcNet *N) {
    int n = N->n, k = N->nvbar;
    int *neworder = xcalloc(k, sizeof(int)), nn = 0;
    unsigned char *pending = xcalloc(n, 1), *seen = xcalloc(n, 1);
    int *childs = xcalloc(n, sizeof(int)), nchild = 0, *parents = xcalloc(n, sizeof(int));
    int left = 0;
    for (int i = 0; i < k; i++) if (!pending[N->vbar[i]]) { pending[N->vbar[i]] = 1; left++; }
    for (int v = 0; v < n; v++)
        if (N->out_ptr[v + 1] - N->out_ptr[v] == 1 && N->head[N->out_arc[N->out_ptr[v]]] == n - 1) {
            childs[nchild++] = v;
            if (pending[v]) { neworder[nn++] = v; pending[v] = 0; left--; }
        }
    while (left > 0) {
        int np = 0;
        memset(seen, 0, n);
        for (int c = 0; c < nchild; c++)
            for (int e = N->in_ptr[childs[c]]; e < N->in_ptr[childs[c] + 1]; e++) {
                int p = N->tail[N->in_arc[e]];
                if (!seen[p]) { seen[p] = 1; parents[np++] = p; }
            }
        if (np <= 0) { free(neworder); free(pending); free(seen); free(childs); free(parents); return -1; }
        for (int i = 0; i < np; i++)
            if (pending[parents[i]]) { neworder[nn++] = parents[i]; pending[parents[i]] = 0; left--; }
        memcpy(childs, parents, (size_t)np * sizeof(int));
        nchild = np;
    }
    memcpy(N->vbar, neworder, (size_t)nn * sizeof(int));
    N->nvbar = nn;
    free(neworder); free(pending); free(seen); free(childs); free(parents);
    return 0;
}

static const OrcNet *g_sort_net;
static int cmp_slot(const void *a, const void *b) {
    const OrcNet *N = g_sort_net;
    int sa = *(const int *)a, sb = *(const int *)b;
    int ia = N->tail[N->slot_in[sa]], ib = N->tail[N->slot_in[sb]];
    if (ia != ib) return ia < ib ? -1 : 1;
    int qa = N->head[N->slot_in[sa]], qb = N->head[N->slot_in[sb]];
    if (qa != qb) return qa < qb ? -1 : 1;
    int ja = N->head[N->slot_out[sa]], jb = N->head[N->slot_out[sb]];
    return ja < jb ? -1 : (ja > jb);
}

/* Network::Network (Network.cpp:10-129) from arrays instead of a file.  u,l are [m][S]. */
OrcNet *orc_create(int n, int m, int S, const int *tail, const int *head, const int *u, const int *l,
                   const int *r0, const int *vbar, int nvbar) {
    OrcNet *N = xcalloc(1, sizeof(OrcNet));
    N->n = n; N->m = m; N->S = S;
    N->tail = xcalloc(m, sizeof(int)); N->head = xcalloc(m, sizeof(int)); N->rew = xcalloc(m, sizeof(int));
    memcpy(N->tail, tail, m * sizeof(int)); memcpy(N->head, head, m * sizeof(int)); memcpy(N->rew, r0, m * sizeof(int));
    N->u = xcalloc((size_t)m * S, sizeof(int)); N->l = xcalloc((size_t)m * S, sizeof(int));
    memcpy(N->u, u, (size_t)m * S * sizeof(int)); memcpy(N->l, l, (size_t)m * S * sizeof(int));
    N->in_ptr = xcalloc(n + 1, sizeof(int)); N->out_ptr = xcalloc(n + 1, sizeof(int));
    N->in_arc = xcalloc(m, sizeof(int)); N->out_arc = xcalloc(m, sizeof(int)); N->out_index = xcalloc(m, sizeof(int));
    for (int a = 0; a < m; a++) { N->in_ptr[head[a] + 1]++; N->out_ptr[tail[a] + 1]++; }
    for (int v = 0; v < n; v++) { N->in_ptr[v + 1] += N->in_ptr[v]; N->out_ptr[v + 1] += N->out_ptr[v]; }
    int *fi = xcalloc(n, sizeof(int)), *fo = xcalloc(n, sizeof(int));
    for (int a = 0; a < m; a++) {
        N->in_arc[N->in_ptr[head[a]] + fi[head[a]]++] = a;
        N->out_index[a] = fo[tail[a]];
        N->out_arc[N->out_ptr[tail[a]] + fo[tail[a]]++] = a;
    }
    free(fi); free(fo);
    N->vbar = xcalloc(nvbar, sizeof(int)); memcpy(N->vbar, vbar, nvbar * sizeof(int)); N->nvbar = nvbar;
    N->is_vbar = xcalloc(n, 1); N->active = xcalloc(n, 1); N->is_root = xcalloc(n, 1);
    for (int i = 0; i < nvbar; i++) N->is_vbar[vbar[i]] = 1; /* isVbar (Network.cpp:61) */
    if (shuffle_vbar(N) != 0) { orc_free(N); return NULL; }
    for (int v = 0; v < n; v++) {
        int ind = N->in_ptr[v + 1] - N->in_ptr[v], outd = N->out_ptr[v + 1] - N->out_ptr[v];
        /* no conservation row => no alpha (A1/A2 rows, grb.cpp:54,72); alpha[0]=alpha[n-1]=0 (134-135) */
        N->is_root[v] = (ind == 0 || outd == 0 || v == 0 || v == n - 1);
        /* lambda/mu/sigma/phi exist only where a V-bar node has both in- and out-arcs (grb.cpp:55-60,73-78,91-117) */
        N->active[v] = N->is_vbar[v] && ind > 0 && outd > 0;
    }
    /* processingOrder / totalLayers (Network.cpp:94-121) and the T coefficient keys (grb.h:60-68) */
    N->L = 0;
    for (int i = 0; i < N->nvbar; i++) N->L += N->in_ptr[N->vbar[i] + 1] - N->in_ptr[N->vbar[i]];
    N->layer_arc = xcalloc(N->L, sizeof(int)); N->arc_layer = xcalloc(m, sizeof(int)); N->slot_base = xcalloc(N->L + 1, sizeof(int));
    for (int a = 0; a < m; a++) N->arc_layer[a] = -1;
    int ell = 0, T = 0;
    for (int i = 0; i < N->nvbar; i++) {
        int q = N->vbar[i];
        for (int e = N->in_ptr[q]; e < N->in_ptr[q + 1]; e++) {
            N->layer_arc[ell] = N->in_arc[e]; N->arc_layer[N->in_arc[e]] = ell;
            N->slot_base[ell] = T; T += N->out_ptr[q + 1] - N->out_ptr[q]; ell++;
        }
    }
    N->slot_base[N->L] = T; N->T = T;
    N->slot_in = xcalloc(T, sizeof(int)); N->slot_out = xcalloc(T, sizeof(int)); N->slot_sorted = xcalloc(T, sizeof(int));
    for (ell = 0; ell < N->L; ell++) {
        int a = N->layer_arc[ell], q = N->head[a];
        for (int k = 0; k < N->out_ptr[q + 1] - N->out_ptr[q]; k++) {
            N->slot_in[N->slot_base[ell] + k] = a; N->slot_out[N->slot_base[ell] + k] = N->out_arc[N->out_ptr[q] + k];
        }
    }
    for (int s = 0; s < T; s++) N->slot_sorted[s] = s;
    g_sort_net = N; qsort(N->slot_sorted, T, sizeof(int), cmp_slot); /* std::map order (i,q,j), Cut.h:75 */
    N->cn = xcalloc(n, sizeof(int)); N->nc = 1;
    for (int v = 0; v < n; v++) N->cn[v] = N->is_root[v] ? 0 : (N->active[v] ? -1 : N->nc++);
    N->ch_ptr = xcalloc(m + 1, sizeof(int)); N->ch_arcs = xcalloc(m, sizeof(int));
    N->ch_sv = xcalloc(m, sizeof(int)); N->ch_ev = xcalloc(m, sizeof(int)); N->ch_qs = xcalloc(m, sizeof(int));
    N->ch_qe = xcalloc(m, sizeof(int)); N->ch_r = xcalloc(m, sizeof(int));
    N->arc_chain = xcalloc(m, sizeof(int)); N->arc_pos = xcalloc(m, sizeof(int));
    N->match_out = xcalloc(m, sizeof(int)); N->match_in = xcalloc(m, sizeof(int));
    return N;
}

int orc_L(const OrcNet *N) { return N->L; }
int orc_T(const OrcNet *N) { return N->T; }
int orc_nvbar(const OrcNet *N) { return N->nvbar; }
void orc_get_vbar(const OrcNet *N, int *out) { memcpy(out, N->vbar, N->nvbar * sizeof(int)); }
void orc_get_layers(const OrcNet *N, int *out) { memcpy(out, N->layer_arc, N->L * sizeof(int)); }
/* slot s (processingOrder-major, out-arc minor) -> node ids (i,q,j) */
void orc_get_slots(const OrcNet *N, int *si, int *sq, int *sj) {
    for (int s = 0; s < N->T; s++) { si[s] = N->tail[N->slot_in[s]]; sq[s] = N->head[N->slot_in[s]]; sj[s] = N->head[N->slot_out[s]]; }
}

/* path -> y-bar (grb.cpp:141-150) -> chains.  A decision that is not an out-arc of the layer's
 * V-bar node sets a y entry no constraint reads, i.e. it behaves as -1.  Returns -2 if two
 * in-arcs claim one out-arc (never produced by the DD: DD.cpp:3666-3668). */
static int build_plan(OrcNet *N, const int16_t *path, int plen) {
    int m = N->m;
    for (int a = 0; a < m; a++) { N->match_out[a] = -1; N->match_in[a] = -1; }
    for (int ell = 0; ell < plen && ell < N->L; ell++) {
        int b = path[ell];
        if (b == -1) continue;
        if (b < 0 || b >= m) return -1; /* the reference would index out of bounds (grb.cpp:147) */
        int a = N->layer_arc[ell], q = N->head[a];
        if (!N->active[q]) continue;
        /* y-bar is keyed by NODE ids (grb.cpp:145-148): the decision means "the out-arc of q whose
         * head is head(path[a])"; if q has no such out-arc the entry is never read. */
        int j = N->head[b]; b = -1;
        for (int e = N->out_ptr[q]; e < N->out_ptr[q + 1]; e++) if (N->head[N->out_arc[e]] == j) { b = N->out_arc[e]; break; }
        if (b < 0) continue;
        if (N->match_in[b] != -1) return -2;
        N->match_out[a] = b; N->match_in[b] = a;
    }
    int nch = 0, fill = 0;
    for (int a0 = 0; a0 < m; a0++) {
        if (N->active[N->tail[a0]] && N->match_in[a0] != -1) continue; /* not a chain start */
        int c = nch++, a = a0, pos = 0, r = 0;
        N->ch_ptr[c] = fill;
        N->ch_qs[c] = N->active[N->tail[a0]] ? N->tail[a0] : -1;
        N->ch_sv[c] = N->active[N->tail[a0]] ? -1 : N->cn[N->tail[a0]];
        for (;;) {
            N->ch_arcs[fill++] = a; N->arc_chain[a] = c; N->arc_pos[a] = pos++; r += N->rew[a];
            if (N->active[N->head[a]] && N->match_out[a] != -1) a = N->match_out[a]; else break;
        }
        N->ch_qe[c] = N->active[N->head[a]] ? N->head[a] : -1;
        N->ch_ev[c] = N->active[N->head[a]] ? -1 : N->cn[N->head[a]];
        N->ch_r[c] = r;
    }
    N->ch_ptr[nch] = fill; N->nch = nch;
    if (fill != m) return -3; /* a cycle made only of matched pairs: instance is not a DAG */
    return 0;
}

/* ---- optimal flow on the contracted graph: successive shortest paths --------------------- */
typedef struct {
    int nc, nch;
    const int *sv, *ev, *r;
    int *lo, *up, *x;         /* per chain; closed chains have up=0 and are skipped */
    unsigned char *open;
    int *dist, *pred, *exc;   /* nc+1 nodes: index nc is the root seen as a path END */
} Flow;

/* Bellman-Ford from `src` over residual arcs; strict improvement => predecessor tree. */
static void bf_paths(Flow *F, int src) {
    int nn = F->nc + 1;
    for (int v = 0; v < nn; v++) { F->dist[v] = INF_D; F->pred[v] = -1; }
    F->dist[src] = 0;
    for (int pass = 0; pass <= nn; pass++) {
        int changed = 0;
        for (int c = 0; c < F->nch; c++) {
            if (!F->open[c]) continue;
            int a = F->sv[c], b = F->ev[c], bt = b == 0 ? F->nc : b, at = a == 0 ? F->nc : a;
            if (F->x[c] < F->up[c] && F->dist[a] < INF_D && F->dist[a] - F->r[c] < F->dist[bt]) {
                F->dist[bt] = F->dist[a] - F->r[c]; F->pred[bt] = 2 * c; changed = 1;
            }
            if (F->x[c] > F->lo[c] && F->dist[b] < INF_D && F->dist[b] + F->r[c] < F->dist[at]) {
                F->dist[at] = F->dist[b] + F->r[c]; F->pred[at] = 2 * c + 1; changed = 1;
            }
        }
        if (!changed) return;
    }
    fprintf(stderr, "oracle: negative residual cycle (instance is not a DAG?)\n"); abort();
}

static int path_bottleneck(Flow *F, int src, int dst, int limit) {
    int v = dst, guard = 0;
    while (v != src) {
        int p = F->pred[v], c = p >> 1, res;
        if (p & 1) { res = F->x[c] - F->lo[c]; v = F->ev[c]; } else { res = F->up[c] - F->x[c]; v = F->sv[c]; }
        if (res < limit) limit = res;
        if (++guard > 4 * (F->nc + 2)) { fprintf(stderr, "oracle: predecessor loop\n"); abort(); }
    }
    return limit;
}
static void path_push(Flow *F, int src, int dst, int d) {
    int v = dst;
    while (v != src) {
        int p = F->pred[v], c = p >> 1;
        if (p & 1) { F->x[c] -= d; v = F->ev[c]; } else { F->x[c] += d; v = F->sv[c]; }
    }
}

/* 0 = optimal, 1 = infeasible */
static int solve_flow(Flow *F) {
    int nc = F->nc;
    for (int v = 0; v <= nc; v++) F->exc[v] = 0;
    for (int c = 0; c < F->nch; c++) {
        if (!F->open[c]) { if (F->lo[c] > 0) return 1; F->x[c] = 0; continue; }
        if (F->lo[c] > F->up[c]) return 1;
        F->x[c] = F->lo[c];
        if (F->ev[c] > 0) F->exc[F->ev[c]] += F->lo[c];
        if (F->sv[c] > 0) F->exc[F->sv[c]] -= F->lo[c];
    }
    /* phase A: route forced flow (lower bounds) along shortest residual paths */
    for (int v = 1; v < nc; v++)
        while (F->exc[v] > 0) {
            bf_paths(F, v);
            int best = -1;
            if (F->dist[nc] < INF_D) best = nc;
            for (int w = 1; w < nc; w++)
                if (F->exc[w] < 0 && F->dist[w] < INF_D && (best < 0 || F->dist[w] < F->dist[best])) best = w;
            if (best < 0) return 1;
            int lim = F->exc[v];
            if (best != nc && -F->exc[best] < lim) lim = -F->exc[best];
            int d = path_bottleneck(F, v, best, lim);
            path_push(F, v, best, d);
            F->exc[v] -= d; if (best != nc) F->exc[best] += d;
        }
    for (int v = 1; v < nc; v++)
        while (F->exc[v] < 0) {
            bf_paths(F, 0);
            if (F->dist[v] >= INF_D) return 1;
            int d = path_bottleneck(F, 0, v, -F->exc[v]);
            path_push(F, 0, v, d);
            F->exc[v] += d;
        }
    /* phase B: profitable root->root cycles */
    for (;;) {
        bf_paths(F, 0);
        if (F->dist[nc] >= 0) break;
        int d = path_bottleneck(F, 0, nc, INT_MAX);
        path_push(F, 0, nc, d);
    }
    return 0;
}

/* SPEC-LP potentials: d = shortest residual distance from the root (phase 1); least consistent
 * labels for nodes the root cannot reach (phase 2); zero-rooted completion (phase 3).  P = -d. */
static void canonical_potentials(Flow *F, int *P) {
    int nc = F->nc, *d = F->dist;
    unsigned char *lab = xcalloc(nc, 1);
    for (int v = 0; v < nc; v++) d[v] = INF_D;
    d[0] = 0;
    for (int pass = 0;; pass++) {
        int changed = 0;
        for (int c = 0; c < F->nch; c++) {
            if (!F->open[c]) continue;
            int a = F->sv[c], b = F->ev[c];
            if (F->x[c] < F->up[c] && d[a] < INF_D && b != 0 && d[a] - F->r[c] < d[b]) { d[b] = d[a] - F->r[c]; changed = 1; }
            if (F->x[c] > F->lo[c] && d[b] < INF_D && a != 0 && d[b] + F->r[c] < d[a]) { d[a] = d[b] + F->r[c]; changed = 1; }
        }
        if (!changed) break;
        if (pass > nc + 1) { fprintf(stderr, "oracle: potentials diverge\n"); abort(); }
    }
    int unl = 0;
    for (int v = 0; v < nc; v++) { lab[v] = d[v] < INF_D; unl += !lab[v]; }
    if (unl) {
        /* phase 2: for residual arc v->w (cost k) with v unlabelled: d[v] >= d[w] - k; take the least */
        for (int v = 0; v < nc; v++) if (!lab[v]) d[v] = -INF_D;
        for (int pass = 0;; pass++) {
            int changed = 0;
            for (int c = 0; c < F->nch; c++) {
                if (!F->open[c]) continue;
                int a = F->sv[c], b = F->ev[c];
                if (F->x[c] < F->up[c] && !lab[a] && d[b] > -INF_D && d[b] + F->r[c] > d[a]) { d[a] = d[b] + F->r[c]; changed = 1; }
                if (F->x[c] > F->lo[c] && !lab[b] && d[a] > -INF_D && d[a] - F->r[c] > d[b]) { d[b] = d[a] - F->r[c]; changed = 1; }
            }
            if (!changed) break;
            if (pass > nc + 1) { fprintf(stderr, "oracle: potentials diverge (2)\n"); abort(); }
        }
        /* phase 3: nodes cut off both ways start at 0 and are relaxed from everything labelled */
        int iso = 0;
        for (int v = 0; v < nc; v++) if (d[v] == -INF_D) { d[v] = 0; lab[v] = 2; iso++; }
        if (iso)
            for (int pass = 0;; pass++) {
                int changed = 0;
                for (int c = 0; c < F->nch; c++) {
                    if (!F->open[c]) continue;
                    int a = F->sv[c], b = F->ev[c];
                    if (F->x[c] < F->up[c] && lab[b] == 2 && d[a] - F->r[c] < d[b]) { d[b] = d[a] - F->r[c]; changed = 1; }
                    if (F->x[c] > F->lo[c] && lab[a] == 2 && d[b] + F->r[c] < d[a]) { d[a] = d[b] + F->r[c]; changed = 1; }
                }
                if (!changed) break;
                if (pass > nc + 1) { fprintf(stderr, "oracle: potentials diverge (3)\n"); abort(); }
            }
    }
    for (int v = 0; v < nc; v++) P[v] = -d[v];
    free(lab);
}

typedef struct { /* one scenario's dual solution over the USED variables of grb.h:44-51 */
    int *alpha;                        /* [n] */
    int *beta, *gamma, *sigma, *phi;   /* [m], indexed by arc (= node pair: the graph is simple) */
    int *lambda, *mu;                  /* [T] */
} Duals;

/* Per-scenario solve + SPEC-LP lifting.  Returns status (0 optimal, 1 infeasible); *obj = primal optimum. */
static int solve_scenario(const OrcNet *N, int s, Flow *F, int *P, int *ph, int *alpha_q, Duals *D, long long *obj) {
    int m = N->m, S = N->S;
    for (int c = 0; c < N->nch; c++) {
        int lo = 0, up = INT_MAX;
        for (int e = N->ch_ptr[c]; e < N->ch_ptr[c + 1]; e++) {
            int a = N->ch_arcs[e], ua = N->u[(size_t)a * S + s], la = N->l[(size_t)a * S + s];
            if (ua < up) up = ua;
            if (la > lo) lo = la;
        }
        F->open[c] = N->ch_sv[c] >= 0 && N->ch_ev[c] >= 0;
        F->lo[c] = lo; F->up[c] = F->open[c] ? up : 0;
    }
    if (solve_flow(F)) return 1;
    long long o = 0;
    for (int c = 0; c < N->nch; c++) if (F->open[c]) o += (long long)N->ch_r[c] * F->x[c];
    *obj = o;
    if (!D) return 0;
    canonical_potentials(F, P);
    memset(D->alpha, 0, N->n * sizeof(int));
    memset(D->beta, 0, m * sizeof(int)); memset(D->gamma, 0, m * sizeof(int));
    memset(D->sigma, 0, m * sizeof(int)); memset(D->phi, 0, m * sizeof(int));
    memset(D->lambda, 0, N->T * sizeof(int)); memset(D->mu, 0, N->T * sizeof(int));
    for (int v = 0; v < N->n; v++) if (N->cn[v] > 0) D->alpha[v] = P[N->cn[v]];
    /* 1. head potentials ph[a] of every arc whose head is a wire or an anchored node */
    for (int c = 0; c < N->nch; c++) {
        int b0 = N->ch_ptr[c], b1 = N->ch_ptr[c + 1], k = b1 - b0;
        if (F->open[c]) {
            int Dp = P[N->ch_ev[c]] - P[N->ch_sv[c]];
            int G = N->ch_r[c] - Dp > 0 ? N->ch_r[c] - Dp : 0, B = Dp - N->ch_r[c] > 0 ? Dp - N->ch_r[c] : 0;
            int pstar = -1, pcirc = -1;
            for (int e = b0; e < b1; e++) {
                int a = N->ch_arcs[e];
                if (pstar < 0 && N->u[(size_t)a * S + s] == F->up[c]) pstar = e;
                if (N->l[(size_t)a * S + s] == F->lo[c]) pcirc = e;
            }
            int acc = P[N->ch_sv[c]];
            for (int e = b0; e < b1; e++) {
                int a = N->ch_arcs[e];
                acc += N->rew[a];
                if (e == pstar) {
                    acc -= G;
                    if (N->active[N->head[a]]) D->sigma[a] = G; else if (N->active[N->tail[a]]) D->phi[a] = G; else D->gamma[a] = G;
                }
                if (e == pcirc) { acc += B; D->beta[a] = B; }
                ph[a] = acc;
            }
        } else if (N->ch_sv[c] >= 0) { /* anchored start, dangling end: tight rows from the anchor */
            int acc = P[N->ch_sv[c]];
            for (int e = b0; e < b1; e++) { acc += N->rew[N->ch_arcs[e]]; ph[N->ch_arcs[e]] = acc; }
        } else if (N->ch_ev[c] >= 0) { /* dangling start, anchored end: tight rows backwards */
            int acc = P[N->ch_ev[c]];
            for (int e = b1 - 1; e >= b0; e--) { ph[N->ch_arcs[e]] = acc; acc -= N->rew[N->ch_arcs[e]]; }
        } else { /* dangling both ways: wires as if the start potential were 0 */
            int acc = 0;
            for (int e = b0; e < b1; e++) { acc += N->rew[N->ch_arcs[e]]; ph[N->ch_arcs[e]] = acc; }
        }
        (void)k;
    }
    /* 2. alpha_q of active V-bar nodes */
    for (int v = 0; v < N->n; v++) {
        if (!N->active[v]) continue;
        int found = 0, val = 0;
        for (int e = N->in_ptr[v]; e < N->in_ptr[v + 1] && !found; e++)
            if (N->match_out[N->in_arc[e]] != -1) { val = ph[N->in_arc[e]]; found = 1; }
        if (!found) {
            for (int e = N->out_ptr[v]; e < N->out_ptr[v + 1]; e++) {
                int c = N->arc_chain[N->out_arc[e]];
                if (N->ch_ev[c] < 0) continue;
                int cand = P[N->ch_ev[c]] - N->ch_r[c];
                if (!found || cand < val) { val = cand; found = 1; }
            }
        }
        alpha_q[v] = found ? val : 0;
        D->alpha[v] = alpha_q[v];
    }
    /* 3. closed arcs (free sigma / phi) and the lambda/mu of matched pairs */
    for (int c = 0; c < N->nch; c++) {
        if (F->open[c]) continue;
        int b0 = N->ch_ptr[c], b1 = N->ch_ptr[c + 1];
        int first = N->ch_arcs[b0], last = N->ch_arcs[b1 - 1];
        if (N->ch_qs[c] >= 0 && N->ch_qe[c] >= 0 && b1 - b0 == 1) {
            int v = N->rew[first] - (alpha_q[N->ch_qe[c]] - alpha_q[N->ch_qs[c]]);
            D->sigma[first] = v > 0 ? v : 0;
            continue;
        }
        if (N->ch_qs[c] >= 0) {
            int v = N->rew[first] - (ph[first] - alpha_q[N->ch_qs[c]]);
            D->phi[first] = v > 0 ? v : 0;
        }
        if (N->ch_qe[c] >= 0) {
            int pt = b1 - b0 > 1 ? ph[N->ch_arcs[b1 - 2]] : P[N->ch_sv[c]];
            int v = N->rew[last] - (alpha_q[N->ch_qe[c]] - pt);
            D->sigma[last] = v > 0 ? v : 0;
        }
    }
    for (int a = 0; a < m; a++) {
        int b = N->match_out[a];
        if (b < 0) continue;
        int slot = N->slot_base[N->arc_layer[a]] + N->out_index[b];
        int dl = ph[a] - alpha_q[N->head[a]];
        if (dl > 0) D->lambda[slot] = dl; else D->mu[slot] = -dl;
    }
    return 0;
}

static Flow *flow_alloc(const OrcNet *N) {
    Flow *F = xcalloc(1, sizeof(Flow));
    F->nc = N->nc; F->nch = N->nch; F->sv = N->ch_sv; F->ev = N->ch_ev; F->r = N->ch_r;
    F->lo = xcalloc(N->m, sizeof(int)); F->up = xcalloc(N->m, sizeof(int)); F->x = xcalloc(N->m, sizeof(int));
    F->open = xcalloc(N->m, 1);
    F->dist = xcalloc(N->nc + 1, sizeof(int)); F->pred = xcalloc(N->nc + 1, sizeof(int)); F->exc = xcalloc(N->nc + 1, sizeof(int));
    return F;
}
static void flow_free(Flow *F) { free(F->lo); free(F->up); free(F->x); free(F->open); free(F->dist); free(F->pred); free(F->exc); free(F); }
static Duals *duals_alloc(const OrcNet *N) {
    Duals *D = xcalloc(1, sizeof(Duals));
    D->alpha = xcalloc(N->n, sizeof(int));
    D->beta = xcalloc(N->m, sizeof(int)); D->gamma = xcalloc(N->m, sizeof(int)); D->sigma = xcalloc(N->m, sizeof(int)); D->phi = xcalloc(N->m, sizeof(int));
    D->lambda = xcalloc(N->T, sizeof(int)); D->mu = xcalloc(N->T, sizeof(int));
    return D;
}
static void duals_free(Duals *D) { free(D->alpha); free(D->beta); free(D->gamma); free(D->sigma); free(D->phi); free(D->lambda); free(D->mu); free(D); }

/* One scenario's dual solution, for the row-by-row check against grb.cpp:49-123 in the tests. */
int orc_scenario_duals(OrcNet *N, const int16_t *path, int plen, int s, int *alpha, int *beta, int *gamma,
                       int *sigma, int *phi, int *lambda, int *mu, double *obj) {
    int rc = build_plan(N, path, plen);
    if (rc) return rc;
    Flow *F = flow_alloc(N);
    Duals D = {alpha, beta, gamma, sigma, phi, lambda, mu};
    int *P = xcalloc(N->nc + 1, sizeof(int)), *ph = xcalloc(N->m, sizeof(int)), *aq = xcalloc(N->n, sizeof(int));
    long long o = 0;
    int st = solve_scenario(N, s, F, P, ph, aq, &D, &o);
    *obj = (double)o;
    free(P); free(ph); free(aq); flow_free(F);
    return st;
}

/* The feasibility ray (replaces GRB_DoubleAttr_UnbdRay, grb.cpp:304-344).  SPEC-LP: if some arc
 * has l > u~ (u~ = 0 for a closed arc) the ray is beta = 1 and the capacity multiplier = 1 on the
 * lowest-index such arc.  Otherwise it is the indicator of the MINIMAL minimum cut of the
 * lower-bound feasibility network (the node set reachable from the super-source in the residual
 * graph of any maximum flow), shifted so that the root has potential 0. */
static int ray_scenario(const OrcNet *N, int s, Duals *D);

/* GuroSolver::solveSubProblem(path) (grb.cpp:139-159): the whole call.
 * Outputs: cut_type (0 OPTIMALITY / 1 FEASIBILITY, Cut.h:22-25), rhs, coef_dense[T] in slot order,
 * the Inavap::Cut pairs (keys/vals/nnz; Cut.h:406-421), per-scenario objective and status
 * (status 2 = not evaluated: the loop breaks at the first infeasible scenario, grb.cpp:350),
 * and the exact integer sums before the 1/S scaling (isum[0] = RHS, isum[1+slot]). */
int orc_solve_path(OrcNet *N, const int16_t *path, int plen, int *cut_type, double *rhs_out, double *coef_dense,
                   uint64_t *keys, double *vals, int *nnz, double *obj, unsigned char *status, long long *isum,
                   int *first_infeasible) {
    int rc = build_plan(N, path, plen);
    if (rc) return rc;
    int n = N->n, m = N->m, S = N->S, T = N->T;
    Flow *F = flow_alloc(N);
    Duals *D = duals_alloc(N);
    int *P = xcalloc(N->nc + 1, sizeof(int)), *ph = xcalloc(m, sizeof(int)), *aq = xcalloc(n, sizeof(int));
    double scenarios = S; /* `double scenarios` (grb.cpp:169) */
    double rhs = 0.0;
    double *coef = xcalloc(T, sizeof(double));
    long long *is = xcalloc(T + 1, sizeof(long long));
    int type = 0;
    *first_infeasible = -1;
    if (status) memset(status, 2, S);
    for (int s = 0; s < S; s++) {
        long long o = 0;
        int st = solve_scenario(N, s, F, P, ph, aq, D, &o);
        if (status) status[s] = (unsigned char)st;
        if (st == 0) {
            if (obj) obj[s] = (double)o;
            /* first term (grb.cpp:238-244) */
            for (int q = 0; q < n; q++)
                for (int e = N->out_ptr[q]; e < N->out_ptr[q + 1]; e++) {
                    int a = N->out_arc[e];
                    rhs += (N->u[(size_t)a * S + s] / scenarios) * D->gamma[a];
                    rhs -= (N->l[(size_t)a * S + s] / scenarios) * D->beta[a];
                    is[0] += (long long)N->u[(size_t)a * S + s] * D->gamma[a] - (long long)N->l[(size_t)a * S + s] * D->beta[a];
                }
            /* second term (grb.cpp:246-259) */
            for (int iv = 0; iv < N->nvbar; iv++) {
                int q = N->vbar[iv];
                for (int e = N->in_ptr[q]; e < N->in_ptr[q + 1]; e++) {
                    int ain = N->in_arc[e];
                    for (int f = N->out_ptr[q]; f < N->out_ptr[q + 1]; f++) {
                        int aout = N->out_arc[f], slot = N->slot_base[N->arc_layer[ain]] + (f - N->out_ptr[q]);
                        rhs += (N->u[(size_t)ain * S + s] / scenarios) * D->lambda[slot];
                        rhs += (N->u[(size_t)aout * S + s] / scenarios) * D->mu[slot];
                        coef[slot] -= (N->u[(size_t)ain * S + s] / scenarios) * D->lambda[slot];
                        coef[slot] -= (N->u[(size_t)aout * S + s] / scenarios) * D->mu[slot];
                        long long t = (long long)N->u[(size_t)ain * S + s] * D->lambda[slot] + (long long)N->u[(size_t)aout * S + s] * D->mu[slot];
                        is[0] += t; is[1 + slot] -= t;
                    }
                }
            }
            /* third term (grb.cpp:261-270): sigma_iq is added to EVERY j */
            for (int iv = 0; iv < N->nvbar; iv++) {
                int q = N->vbar[iv];
                for (int e = N->in_ptr[q]; e < N->in_ptr[q + 1]; e++) {
                    int ain = N->in_arc[e];
                    unsigned u_iq = (unsigned)N->u[(size_t)ain * S + s]; /* `uint u_iq` (grb.cpp:264) */
                    for (int f = N->out_ptr[q]; f < N->out_ptr[q + 1]; f++) {
                        int slot = N->slot_base[N->arc_layer[ain]] + (f - N->out_ptr[q]);
                        coef[slot] += (u_iq / scenarios) * D->sigma[ain];
                        is[1 + slot] += (long long)u_iq * D->sigma[ain];
                    }
                }
            }
            /* fourth term (grb.cpp:272-281): phi_qj is added to EVERY i */
            for (int iv = 0; iv < N->nvbar; iv++) {
                int q = N->vbar[iv];
                for (int f = N->out_ptr[q]; f < N->out_ptr[q + 1]; f++) {
                    int aout = N->out_arc[f];
                    unsigned u_qj = (unsigned)N->u[(size_t)aout * S + s];
                    for (int e = N->in_ptr[q]; e < N->in_ptr[q + 1]; e++) {
                        int slot = N->slot_base[N->arc_layer[N->in_arc[e]]] + (f - N->out_ptr[q]);
                        coef[slot] += (u_qj / scenarios) * D->phi[aout];
                        is[1 + slot] += (long long)u_qj * D->phi[aout];
                    }
                }
            }
        } else {
            /* feasibility cut from THIS scenario only, no 1/S (grb.cpp:284-351) */
            rhs = 0; memset(coef, 0, T * sizeof(double)); memset(is, 0, (T + 1) * sizeof(long long));
            ray_scenario(N, s, D);
            for (int q = 0; q < n; q++)
                for (int e = N->out_ptr[q]; e < N->out_ptr[q + 1]; e++) {
                    int a = N->out_arc[e];
                    rhs += N->u[(size_t)a * S + s] * (double)D->gamma[a];
                    rhs -= N->l[(size_t)a * S + s] * (double)D->beta[a];
                    is[0] += (long long)N->u[(size_t)a * S + s] * D->gamma[a] - (long long)N->l[(size_t)a * S + s] * D->beta[a];
                }
            for (int iv = 0; iv < N->nvbar; iv++) {
                int q = N->vbar[iv];
                for (int e = N->in_ptr[q]; e < N->in_ptr[q + 1]; e++) {
                    int ain = N->in_arc[e];
                    for (int f = N->out_ptr[q]; f < N->out_ptr[q + 1]; f++) {
                        int aout = N->out_arc[f], slot = N->slot_base[N->arc_layer[ain]] + (f - N->out_ptr[q]);
                        double t = N->u[(size_t)ain * S + s] * (double)D->lambda[slot] + N->u[(size_t)aout * S + s] * (double)D->mu[slot];
                        rhs += t; coef[slot] -= t;
                        long long ti = (long long)N->u[(size_t)ain * S + s] * D->lambda[slot] + (long long)N->u[(size_t)aout * S + s] * D->mu[slot];
                        is[0] += ti; is[1 + slot] -= ti;
                        coef[slot] += (double)(unsigned)N->u[(size_t)ain * S + s] * D->sigma[ain];
                        coef[slot] += (double)(unsigned)N->u[(size_t)aout * S + s] * D->phi[aout];
                        is[1 + slot] += (long long)N->u[(size_t)ain * S + s] * D->sigma[ain] + (long long)N->u[(size_t)aout * S + s] * D->phi[aout];
                    }
                }
            }
            type = 1; *first_infeasible = s;
            break; /* grb.cpp:350 */
        }
    }
    *cut_type = type; *rhs_out = rhs;
    if (coef_dense) memcpy(coef_dense, coef, T * sizeof(double));
    if (isum) memcpy(isum, is, (T + 1) * sizeof(long long));
    /* cutToCut (Cut.h:406-421): (i,q,j)-lexicographic, exact zeros dropped, key = q | i<<16 | j<<32 */
    int k = 0;
    for (int t = 0; t < T; t++) {
        int slot = N->slot_sorted[t];
        if (coef[slot] == 0) continue;
        if (keys) {
            uint64_t i = (uint64_t)N->tail[N->slot_in[slot]], q = (uint64_t)N->head[N->slot_in[slot]], j = (uint64_t)N->head[N->slot_out[slot]];
            keys[k] = q | (i << 16) | (j << 32);
            vals[k] = coef[slot];
        }
        k++;
    }
    if (nnz) *nnz = k;
    free(coef); free(is); free(P); free(ph); free(aq); flow_free(F); duals_free(D);
    return 0;
}

/* ---- feasibility ray -------------------------------------------------------------------- */
/* Works on the SPLIT graph (every original arc; a wire node per matched pair) so that chains
 * whose own arcs conflict (l_a > u_b) need no special case. */
static int ray_scenario(const OrcNet *N, int s, Duals *D) {
    int n = N->n, m = N->m, S = N->S;
    memset(D->alpha, 0, n * sizeof(int));
    memset(D->beta, 0, m * sizeof(int)); memset(D->gamma, 0, m * sizeof(int));
    memset(D->sigma, 0, m * sizeof(int)); memset(D->phi, 0, m * sizeof(int));
    memset(D->lambda, 0, N->T * sizeof(int)); memset(D->mu, 0, N->T * sizeof(int));
    /* split nodes: 0 = root, cn[v] for plain nodes, then one wire per matched in-arc, then one
     * private stub per dangling arc end (closed arcs carry no flow; the stub keeps them inert) */
    int nn = N->nc, *wire = xcalloc(m, sizeof(int));
    for (int a = 0; a < m; a++) wire[a] = N->match_out[a] >= 0 ? nn++ : -1;
    int *ts = xcalloc(m, sizeof(int)), *hs = xcalloc(m, sizeof(int)), *cap = xcalloc(m, sizeof(int));
    unsigned char *closed = xcalloc(m, 1);
    for (int a = 0; a < m; a++) {
        int t = N->tail[a], h = N->head[a];
        int ua = N->u[(size_t)a * S + s], la = N->l[(size_t)a * S + s];
        if (N->active[t]) { if (N->match_in[a] >= 0) ts[a] = wire[N->match_in[a]]; else { ts[a] = -1; closed[a] = 1; } } else ts[a] = N->cn[t];
        if (N->active[h]) { if (N->match_out[a] >= 0) hs[a] = wire[a]; else { hs[a] = -1; closed[a] = 1; } } else hs[a] = N->cn[h];
        int ut = closed[a] ? 0 : ua;
        if (la > ut) { /* single-arc certificate, lowest arc id; a closed arc uses its FREE multiplier */
            D->beta[a] = 1;
            if (closed[a]) { if (hs[a] < 0) D->sigma[a] = 1; else D->phi[a] = 1; }
            else if (N->active[h]) D->sigma[a] = 1; else if (N->active[t]) D->phi[a] = 1; else D->gamma[a] = 1;
            free(wire); free(ts); free(hs); free(cap); free(closed);
            return 0;
        }
        cap[a] = ut - la;
    }
    /* max flow from super-source (excess nodes) to super-sink (deficit nodes); the root conserves
     * automatically, so it is an ordinary node here.  Plain augmenting paths by BFS. */
    long long *b = xcalloc(nn, sizeof(long long));
    for (int a = 0; a < m; a++) {
        if (closed[a]) continue;
        int la = N->l[(size_t)a * S + s];
        b[hs[a]] += la; b[ts[a]] -= la;
    }
    int *f = xcalloc(m, sizeof(int)), *pred = xcalloc(nn, sizeof(int)), *queue = xcalloc(nn, sizeof(int));
    long long *srcl = xcalloc(nn, sizeof(long long)), *snkl = xcalloc(nn, sizeof(long long));
    for (int v = 0; v < nn; v++) { srcl[v] = b[v] > 0 ? b[v] : 0; snkl[v] = b[v] < 0 ? -b[v] : 0; }
    unsigned char *vis = xcalloc(nn, 1);
    for (;;) {
        int qh = 0, qt = 0, found = -1;
        memset(vis, 0, nn);
        for (int v = 0; v < nn; v++) if (srcl[v] > 0) { vis[v] = 1; pred[v] = -1; queue[qt++] = v; }
        while (qh < qt && found < 0) {
            int v = queue[qh++];
            if (snkl[v] > 0) { found = v; break; }
            for (int a = 0; a < m; a++) { /* O(m) scan per node: the ray is built once per call */
                if (closed[a]) continue;
                if (ts[a] == v && f[a] < cap[a] && !vis[hs[a]]) { vis[hs[a]] = 1; pred[hs[a]] = 2 * a; queue[qt++] = hs[a]; }
                if (hs[a] == v && f[a] > 0 && !vis[ts[a]]) { vis[ts[a]] = 1; pred[ts[a]] = 2 * a + 1; queue[qt++] = ts[a]; }
            }
        }
        if (found < 0) break;
        long long d = snkl[found];
        int v = found;
        while (pred[v] >= 0) {
            int a = pred[v] >> 1;
            if (pred[v] & 1) { if (f[a] < d) d = f[a]; v = hs[a]; } else { if (cap[a] - f[a] < d) d = cap[a] - f[a]; v = ts[a]; }
        }
        if (srcl[v] < d) d = srcl[v];
        srcl[v] -= d; snkl[found] -= d;
        v = found;
        while (pred[v] >= 0) {
            int a = pred[v] >> 1;
            if (pred[v] & 1) { f[a] -= (int)d; v = hs[a]; } else { f[a] += (int)d; v = ts[a]; }
        }
    }
    /* vis[] now marks X = the minimal min-cut source side.  Potential = 1[X] - 1[root in X]. */
    int shift = vis[0] ? 1 : 0;
    int *pot = xcalloc(nn, sizeof(int));
    for (int v = 0; v < nn; v++) pot[v] = (vis[v] ? 1 : 0) - shift;
    for (int v = 0; v < n; v++) if (N->cn[v] > 0) D->alpha[v] = pot[N->cn[v]];
    /* alpha_q of an active V-bar node: potential of its first matched wire, else 0 */
    for (int v = 0; v < n; v++) {
        if (!N->active[v]) continue;
        for (int e = N->in_ptr[v]; e < N->in_ptr[v + 1]; e++)
            if (N->match_out[N->in_arc[e]] >= 0) { D->alpha[v] = pot[wire[N->in_arc[e]]]; break; }
    }
    for (int a = 0; a < m; a++) {
        int t = N->tail[a], h = N->head[a];
        int pt = ts[a] >= 0 ? pot[ts[a]] : D->alpha[t], phd = hs[a] >= 0 ? pot[hs[a]] : D->alpha[h];
        int need = -(phd - pt); /* ray rows are homogeneous: ph - pt - beta + Gamma >= 0 */
        if (need > 0) {
            if (closed[a]) { if (hs[a] < 0) D->sigma[a] = need; else D->phi[a] = need; }
            else if (N->active[h]) D->sigma[a] = need; else if (N->active[t]) D->phi[a] = need; else D->gamma[a] = need;
        }
        else if (need < 0 && !closed[a]) D->beta[a] = -need;
        if (N->match_out[a] >= 0) {
            int slot = N->slot_base[N->arc_layer[a]] + N->out_index[N->match_out[a]];
            int dl = pot[wire[a]] - D->alpha[h];
            if (dl > 0) D->lambda[slot] = dl; else D->mu[slot] = -dl;
        }
    }
    free(wire); free(ts); free(hs); free(cap); free(closed); free(b); free(f); free(pred); free(queue);
    free(srcl); free(snkl); free(vis); free(pot);
    return 0;
}

/* The ray alone, for the row-by-row check in the tests. */
int orc_scenario_ray(OrcNet *N, const int16_t *path, int plen, int s, int *alpha, int *beta, int *gamma,
                     int *sigma, int *phi, int *lambda, int *mu) {
    int rc = build_plan(N, path, plen);
    if (rc) return rc;
    Duals D = {alpha, beta, gamma, sigma, phi, lambda, mu};
    return ray_scenario(N, s, &D);
}

/* cpu_baseline leg: statuses + objectives + cut of scenarios [s0,s1) only, accumulated as exact
 * integers (what one host thread of a scenario-sharded CPU run would do). */
int orc_solve_range(OrcNet *N, const int16_t *path, int plen, int s0, int s1, long long *isum, int *n_infeasible) {
    int rc = build_plan(N, path, plen);
    if (rc) return rc;
    int n = N->n, m = N->m, S = N->S;
    Flow *F = flow_alloc(N);
    Duals *D = duals_alloc(N);
    int *P = xcalloc(N->nc + 1, sizeof(int)), *ph = xcalloc(m, sizeof(int)), *aq = xcalloc(n, sizeof(int));
    int bad = 0;
    for (int s = s0; s < s1; s++) {
        long long o = 0;
        if (solve_scenario(N, s, F, P, ph, aq, D, &o)) { bad++; continue; }
        for (int a = 0; a < m; a++) {
            long long ua = N->u[(size_t)a * S + s], la = N->l[(size_t)a * S + s];
            isum[0] += ua * D->gamma[a] - la * D->beta[a];
            if (N->arc_layer[a] >= 0 && N->active[N->head[a]]) {
                int q = N->head[a], base = N->slot_base[N->arc_layer[a]];
                for (int f = 0; f < N->out_ptr[q + 1] - N->out_ptr[q]; f++) {
                    int aout = N->out_arc[N->out_ptr[q] + f];
                    long long t = ua * D->lambda[base + f] + (long long)N->u[(size_t)aout * S + s] * D->mu[base + f];
                    isum[0] += t;
                    isum[1 + base + f] += -t + ua * D->sigma[a] + (long long)N->u[(size_t)aout * S + s] * D->phi[aout];
                }
            }
        }
    }
    *n_infeasible = bad;
    free(P); free(ph); free(aq); flow_free(F); duals_free(D);
    return 0;
}
