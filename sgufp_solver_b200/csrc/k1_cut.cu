// k1_cut.cu — K1: batched per-scenario Benders subproblem evaluation + cut folding (sm_100a).
//
// Replaces the scenario loop of GuroSolver::solveSubProblem (/root/reference/grb.cpp:162-360):
// for every (candidate k, scenario s) it solves the second-stage LP exactly and adds the
// scenario's dual contribution cap x dual (grb.cpp:238-281) into exact integer accumulators.
//
// Mapping: one warp per (candidate, scenario); a persistent grid of
// sm_count x CTAS_PER_SM CTAs strides over the K*S work items scenario-minor, so neighbouring
// warps stream neighbouring rows of the scenario-major fp64 capacity arrays with coalesced
// 128-bit loads.  The candidate's contracted graph (chains) is read-only and
// stays in L1; all per-scenario state (chain capacities, flows, labels, potentials) lives in the
// warp's slice of shared memory.  No tensor cores: this is integer graph work.
//
// Per work item:
//   1. stream u_s, l_s; segmented min/max into chain capacities (shared-memory atomics, packed
//      with the position of the FIRST least-capacity / LAST greatest-lower-bound arc)
//   2. optimal flow by successive shortest paths; labels are (distance, hops) packed in one
//      int so that a warp-wide atomicMin relaxation yields an acyclic predecessor structure
//   3. SPEC-LP potentials (algorithm-independent) by label correction from the root
//   4. lift to (gamma, beta, sigma, phi, lambda, mu), multiply by the capacities and add to the
//      candidate's accumulators; write objective + status
#include "k1_cut.cuh"

#include <climits>

#include "model.hpp"

namespace sgufp {

namespace {

constexpr int HB = 10;                 // hop bits of a label
constexpr int LAB_BIAS = 1 << 20;      // distance bias so that labels are positive ints
constexpr int LAB_INF = 0x7fffffff;
constexpr int NEG_INF = INT_MIN / 4;
constexpr int WARPS = 8;               // warps per CTA
constexpr unsigned FULL = 0xffffffffu;

struct PlanView {
    const PlanHeader *h;
    const int32_t *arc_cp, *arc_info, *arc_pre, *ch_ends, *ch_r, *ch_ptr, *ch_arcs, *ch_q, *av_ptr, *av_arcs, *fb_ptr, *fb_ch;
    __device__ explicit PlanView(const int32_t *base) {
        h = reinterpret_cast<const PlanHeader *>(base);
        arc_cp = base + h->o_arc_cp; arc_info = base + h->o_arc_info; arc_pre = base + h->o_arc_pre;
        ch_ends = base + h->o_ch_ends; ch_r = base + h->o_ch_r; ch_ptr = base + h->o_ch_ptr; ch_arcs = base + h->o_ch_arcs;
        ch_q = base + h->o_ch_q; av_ptr = base + h->o_av_ptr; av_arcs = base + h->o_av_arcs; fb_ptr = base + h->o_fb_ptr; fb_ch = base + h->o_fb_ch;
    }
};

struct WarpMem {  // this warp's slice of shared memory
    int *up, *lo, *x, *lab, *pred, *pot, *aq, *exc;
};

__device__ __forceinline__ int lab_dist(int lab) { return (lab >> HB) - LAB_BIAS; }

// Label-correcting shortest paths from `src` over residual arcs.  MERGED: the root is one node
// (index 0) that is never relabelled; otherwise arcs entering the root end at index nc.
template <bool MERGED>
__device__ void shortest_paths(int src, const PlanView &P, int nopen, int nc, int lane, WarpMem &w) {
    for (int v = lane; v <= nc; v += 32) w.lab[v] = LAB_INF;
    __syncwarp();
    if (lane == 0) w.lab[src] = LAB_BIAS << HB;
    __syncwarp();
    bool changed;
    do {
        changed = false;
        for (int c = lane; c < nopen; c += 32) {
            const int e = P.ch_ends[c], sv = (e & 0xffff) - 1, ev = (e >> 16) - 1, r = P.ch_r[c];
            const int xc = w.x[c], up = w.up[c] >> HB, lo = w.lo[c] >> HB;
            if (xc < up && !(MERGED && ev == 0)) {
                const int lu = w.lab[sv];
                if (lu != LAB_INF) {
                    const int cand = lu - r * (1 << HB) + 1;
                    const int t = (!MERGED && ev == 0) ? nc : ev;
                    if (cand < atomicMin(&w.lab[t], cand)) changed = true;
                }
            }
            if (xc > lo && !(MERGED && sv == 0)) {
                const int lu = w.lab[ev];
                if (lu != LAB_INF) {
                    const int cand = lu + r * (1 << HB) + 1;
                    const int t = (!MERGED && sv == 0) ? nc : sv;
                    if (cand < atomicMin(&w.lab[t], cand)) changed = true;
                }
            }
        }
        __syncwarp();
        changed = __any_sync(FULL, changed);
    } while (changed);
}

// One tight residual arc per labelled node; hops strictly decrease along it, so following
// predecessors always reaches the source.
__device__ void mark_predecessors(const PlanView &P, int nopen, int nc, int lane, WarpMem &w) {
    for (int c = lane; c < nopen; c += 32) {
        const int e = P.ch_ends[c], sv = (e & 0xffff) - 1, ev = (e >> 16) - 1, r = P.ch_r[c];
        const int xc = w.x[c], up = w.up[c] >> HB, lo = w.lo[c] >> HB;
        if (xc < up) {
            const int lu = w.lab[sv], t = ev == 0 ? nc : ev;
            if (lu != LAB_INF && lu - r * (1 << HB) + 1 == w.lab[t]) w.pred[t] = 2 * c;
        }
        if (xc > lo) {
            const int lu = w.lab[ev], t = sv == 0 ? nc : sv;
            if (lu != LAB_INF && lu + r * (1 << HB) + 1 == w.lab[t]) w.pred[t] = 2 * c + 1;
        }
    }
    __syncwarp();
}

// lane 0 only: bottleneck of the predecessor path src -> dst, then push it
__device__ int augment(int src, int dst, int limit, const PlanView &P, WarpMem &w) {
    int v = dst, d = limit;
    while (v != src) {
        const int p = w.pred[v], c = p >> 1, e = P.ch_ends[c];
        int res;
        if (p & 1) { res = w.x[c] - (w.lo[c] >> HB); v = (e >> 16) - 1; } else { res = (w.up[c] >> HB) - w.x[c]; v = (e & 0xffff) - 1; }
        d = min(d, res);
    }
    v = dst;
    while (v != src) {
        const int p = w.pred[v], c = p >> 1, e = P.ch_ends[c];
        if (p & 1) { w.x[c] -= d; v = (e >> 16) - 1; } else { w.x[c] += d; v = (e & 0xffff) - 1; }
    }
    return d;
}

__device__ __forceinline__ long long warp_sum(long long v) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
    for (int o = 16; o; o >>= 1) { unsigned long long t = __shfl_xor_sync(FULL, v, o); v = t < v ? t : v; }
    return v;
}

// Forced flow from lower bounds (rare): route every excess / deficit along shortest residual
// paths.  Returns false if some forced flow cannot be routed (scenario infeasible).
__device__ bool route_lower_bounds(const PlanView &P, int nopen, int nc, int lane, WarpMem &w) {
    for (int v = lane; v <= nc; v += 32) w.exc[v] = 0;
    __syncwarp();
    for (int c = lane; c < nopen; c += 32) {
        const int lo = w.lo[c] >> HB;
        if (lo > 0) {
            const int e = P.ch_ends[c], sv = (e & 0xffff) - 1, ev = (e >> 16) - 1;
            if (ev > 0) atomicAdd(&w.exc[ev], lo);
            if (sv > 0) atomicSub(&w.exc[sv], lo);
        }
    }
    __syncwarp();
    for (int v = 1; v < nc; v++) {
        while (w.exc[v] > 0) {
            shortest_paths<false>(v, P, nopen, nc, lane, w);
            unsigned long long best = ~0ull;
            for (int t = lane; t <= nc; t += 32)
                if ((t == nc || (t > 0 && w.exc[t] < 0)) && w.lab[t] != LAB_INF) {
                    unsigned long long key = ((unsigned long long)(unsigned)w.lab[t] << 32) | (unsigned)t;
                    best = key < best ? key : best;
                }
            best = warp_min_u64(best);
            if (best == ~0ull) return false;
            const int t = (int)(best & 0xffffffffu);
            mark_predecessors(P, nopen, nc, lane, w);
            if (lane == 0) {
                int lim = w.exc[v];
                if (t != nc) lim = min(lim, -w.exc[t]);
                const int d = augment(v, t, lim, P, w);
                w.exc[v] -= d;
                if (t != nc) w.exc[t] += d;
            }
            __syncwarp();
        }
    }
    for (int v = 1; v < nc; v++) {
        while (w.exc[v] < 0) {
            shortest_paths<false>(0, P, nopen, nc, lane, w);
            if (w.lab[v] == LAB_INF) return false;
            mark_predecessors(P, nopen, nc, lane, w);
            if (lane == 0) w.exc[v] += augment(0, v, -w.exc[v], P, w);
            __syncwarp();
        }
    }
    return true;
}

// SPEC-LP potentials: pot[v] = -(shortest residual distance from the root);
// nodes the root cannot reach get the least labels consistent with the labelled ones; nodes cut
// off both ways get a zero-rooted completion.
__device__ void canonical_potentials(const PlanView &P, int nopen, int nc, int lane, WarpMem &w) {
    shortest_paths<true>(0, P, nopen, nc, lane, w);
    bool missing = false;
    for (int v = lane; v < nc; v += 32) {
        const int l = w.lab[v];
        if (l == LAB_INF) { missing = true; w.pot[v] = NEG_INF; w.pred[v] = 0; } else { w.pot[v] = lab_dist(l); w.pred[v] = 1; }
    }
    __syncwarp();
    if (__any_sync(FULL, missing)) {
        // pot[] holds d here; pred[] is the state: 1 labelled by phase 1, 0 not yet, 2 isolated
        bool changed;
        do {
            changed = false;
            for (int c = lane; c < nopen; c += 32) {
                const int e = P.ch_ends[c], sv = (e & 0xffff) - 1, ev = (e >> 16) - 1, r = P.ch_r[c];
                const int xc = w.x[c], up = w.up[c] >> HB, lo = w.lo[c] >> HB;
                if (xc < up && w.pred[sv] == 0) { const int db = w.pot[ev]; if (db > NEG_INF && db + r > atomicMax(&w.pot[sv], db + r)) changed = true; }
                if (xc > lo && w.pred[ev] == 0) { const int da = w.pot[sv]; if (da > NEG_INF && da - r > atomicMax(&w.pot[ev], da - r)) changed = true; }
            }
            __syncwarp();
            changed = __any_sync(FULL, changed);
        } while (changed);
        bool iso = false;
        for (int v = lane; v < nc; v += 32)
            if (w.pot[v] == NEG_INF) { w.pot[v] = 0; w.pred[v] = 2; iso = true; }
        __syncwarp();
        if (__any_sync(FULL, iso)) {
            do {
                changed = false;
                for (int c = lane; c < nopen; c += 32) {
                    const int e = P.ch_ends[c], sv = (e & 0xffff) - 1, ev = (e >> 16) - 1, r = P.ch_r[c];
                    const int xc = w.x[c], up = w.up[c] >> HB, lo = w.lo[c] >> HB;
                    if (xc < up && w.pred[ev] == 2) { const int cand = w.pot[sv] - r; if (cand < atomicMin(&w.pot[ev], cand)) changed = true; }
                    if (xc > lo && w.pred[sv] == 2) { const int cand = w.pot[ev] + r; if (cand < atomicMin(&w.pot[sv], cand)) changed = true; }
                }
                __syncwarp();
                changed = __any_sync(FULL, changed);
            } while (changed);
        }
    }
    for (int v = lane; v < nc; v += 32) w.pot[v] = -w.pot[v];
    __syncwarp();
}

// wire potential at the HEAD of arc a (a matched in-arc, or any arc of a chain)
__device__ __forceinline__ int head_potential(int a, const PlanView &P, int nopen, const WarpMem &w) {
    const int cp = P.arc_cp[a], c = cp >> 10, pos = cp & 1023, pre = P.arc_pre[a];
    const int e = P.ch_ends[c], sv = (e & 0xffff) - 1, ev = (e >> 16) - 1;
    if (c < nopen) {
        const int r = P.ch_r[c], dp = w.pot[ev] - w.pot[sv];
        const int g = max(0, r - dp), b = max(0, dp - r);
        return w.pot[sv] + pre - (pos >= (w.up[c] & 1023) ? g : 0) + (pos >= (w.lo[c] & 1023) ? b : 0);
    }
    if (sv >= 0) return w.pot[sv] + pre;
    if (ev >= 0) return w.pot[ev] - (P.ch_r[c] - pre);
    return pre;
}

__global__ void __launch_bounds__(WARPS * 32) k1_cut_eval(K1Launch p, int words_per_warp) {
    extern __shared__ int smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    WarpMem w;
    {
        int *base = smem + (size_t)wid * words_per_warp;
        w.up = base; base += p.max_nch;
        w.lo = base; base += p.max_nch;
        w.x = base; base += p.max_nopen;
        w.lab = base; base += p.nc + 1;
        w.pred = base; base += p.nc + 1;
        w.pot = base; base += p.nc + 1;
        w.exc = base; base += p.nc + 1;
        w.aq = base;
    }
    const long long items = (long long)p.K * p.S;
    const long long stride = (long long)gridDim.x * WARPS;
    for (long long item = (long long)blockIdx.x * WARPS + wid; item < items; item += stride) {
        const int k = (int)(item / p.S), s = (int)(item - (long long)k * p.S);
        const PlanView P(p.plans + p.plan_off[k]);
        const int nch = P.h->nch, nopen = P.h->nopen, nc = P.h->nc, nav = P.h->nav, m = p.m;
        unsigned long long *sums = p.sums + (size_t)k * p.W;
        const double *row_u = p.cap_u + (size_t)s * p.m_pad, *row_l = p.cap_l + (size_t)s * p.m_pad;

        // 1. chain capacities
        for (int c = lane; c < nch; c += 32) { w.up[c] = LAB_INF; w.lo[c] = 0; }
        __syncwarp();
        {
            const double2 *ru = reinterpret_cast<const double2 *>(row_u), *rl = reinterpret_cast<const double2 *>(row_l);
            for (int a2 = lane; a2 < p.m_pad / 2; a2 += 32) {
                const double2 u2 = __ldg(ru + a2), l2 = __ldg(rl + a2);
                const int a = 2 * a2;
                {
                    const int cp = P.arc_cp[a], c = cp >> 10, pos = cp & 1023;
                    atomicMin(&w.up[c], ((int)u2.x << HB) | pos);
                    atomicMax(&w.lo[c], ((int)l2.x << HB) | pos);
                }
                if (a + 1 < m) {
                    const int cp = P.arc_cp[a + 1], c = cp >> 10, pos = cp & 1023;
                    atomicMin(&w.up[c], ((int)u2.y << HB) | pos);
                    atomicMax(&w.lo[c], ((int)l2.y << HB) | pos);
                }
            }
        }
        __syncwarp();
        bool bad = false, forced = false;
        for (int c = lane; c < nch; c += 32) {
            const int lo = w.lo[c] >> HB;
            if (c < nopen) { w.x[c] = lo; bad |= lo > (w.up[c] >> HB); forced |= lo > 0; } else bad |= lo > 0;
        }
        __syncwarp();
        bad = __any_sync(FULL, bad);
        forced = __any_sync(FULL, forced);
        // 2. optimal flow
        if (!bad && forced) bad = !route_lower_bounds(P, nopen, nc, lane, w);
        if (bad) {
            if (lane == 0) {
                atomicMin(p.first_inf + k, p.scen_offset + s);
                if (p.status) p.status[(size_t)k * p.S + s] = 1;
                if (p.obj) p.obj[(size_t)k * p.S + s] = 0.0;
            }
            continue;
        }
        for (;;) {
            shortest_paths<false>(0, P, nopen, nc, lane, w);
            const int lt = w.lab[nc];
            if (lt == LAB_INF || lab_dist(lt) >= 0) break;
            mark_predecessors(P, nopen, nc, lane, w);
            if (lane == 0) augment(0, nc, INT_MAX, P, w);
            __syncwarp();
        }
        // 3. potentials
        canonical_potentials(P, nopen, nc, lane, w);
        // 4. lifting + folding
        long long rhs = 0, objv = 0;
        for (int c = lane; c < nopen; c += 32) {
            const int e = P.ch_ends[c], sv = (e & 0xffff) - 1, ev = (e >> 16) - 1, r = P.ch_r[c];
            const int dp = w.pot[ev] - w.pot[sv];
            const int g = r - dp, up = w.up[c] >> HB, lo = w.lo[c] >> HB;
            objv += (long long)r * w.x[c];
            if (g > 0) {
                const int a = P.ch_arcs[P.ch_ptr[c] + (w.up[c] & 1023)];
                const int info = P.arc_info[a];
                const long long v = (long long)up * g;
                if ((info & 3) == KIND_GAMMA) rhs += v; else atomicAdd(sums + 1 + p.L + a, (unsigned long long)v);
            } else if (g < 0) rhs -= (long long)lo * (-g);
        }
        for (int i = lane; i < nav; i += 32) {
            const int b0 = P.av_ptr[i], b1 = P.av_ptr[i + 1];
            int alpha = 0;
            if (b1 > b0) {
                alpha = head_potential(P.av_arcs[b0], P, nopen, w);
                for (int t = b0 + 1; t < b1; t++) {
                    const int a = P.av_arcs[t];
                    const int dl = head_potential(a, P, nopen, w) - alpha;
                    if (dl != 0) {
                        long long v;
                        if (dl > 0) v = (long long)(int)row_u[a] * dl;
                        else {
                            const int cp = P.arc_cp[a];
                            const int bnext = P.ch_arcs[P.ch_ptr[cp >> 10] + (cp & 1023) + 1];
                            v = (long long)(int)row_u[bnext] * (-dl);
                        }
                        rhs += v;
                        atomicAdd(sums + 1 + (P.arc_info[a] >> 2) - 1, (unsigned long long)v);
                    }
                }
            } else {
                bool found = false;
                for (int t = P.fb_ptr[i]; t < P.fb_ptr[i + 1]; t++) {
                    const int c = P.fb_ch[t], ev = (P.ch_ends[c] >> 16) - 1;
                    const int cand = w.pot[ev] - P.ch_r[c];
                    if (!found || cand < alpha) { alpha = cand; found = true; }
                }
            }
            w.aq[i] = alpha;
        }
        __syncwarp();
        for (int c = nopen + lane; c < nch; c += 32) {
            const int b0 = P.ch_ptr[c], b1 = P.ch_ptr[c + 1], first = P.ch_arcs[b0], last = P.ch_arcs[b1 - 1];
            const int e = P.ch_ends[c], sv = (e & 0xffff) - 1, ev = (e >> 16) - 1;
            const int q = P.ch_q[c], qs = (q & 0xffff) - 1, qe = (q >> 16) - 1;
            const int rfirst = P.arc_pre[first];
            if (qs >= 0 && qe >= 0 && b1 - b0 == 1) {
                const int v = rfirst - (w.aq[qe] - w.aq[qs]);
                if (v > 0) atomicAdd(sums + 1 + p.L + first, (unsigned long long)((long long)(int)row_u[first] * v));
                continue;
            }
            if (qs >= 0) {
                const int phf = ev >= 0 ? w.pot[ev] - (P.ch_r[c] - rfirst) : rfirst;
                const int v = rfirst - (phf - w.aq[qs]);
                if (v > 0) atomicAdd(sums + 1 + p.L + first, (unsigned long long)((long long)(int)row_u[first] * v));
            }
            if (qe >= 0) {
                int pt, rlast;
                if (b1 - b0 > 1) {
                    const int prev = P.ch_arcs[b1 - 2];
                    pt = (sv >= 0 ? w.pot[sv] : 0) + P.arc_pre[prev];
                    rlast = P.ch_r[c] - P.arc_pre[prev];
                } else { pt = w.pot[sv]; rlast = P.ch_r[c]; }
                const int v = rlast - (w.aq[qe] - pt);
                if (v > 0) atomicAdd(sums + 1 + p.L + last, (unsigned long long)((long long)(int)row_u[last] * v));
            }
        }
        rhs = warp_sum(rhs);
        objv = warp_sum(objv);
        if (lane == 0) {
            if (rhs) atomicAdd(sums, (unsigned long long)rhs);
            if (p.status) p.status[(size_t)k * p.S + s] = 0;
            if (p.obj) p.obj[(size_t)k * p.S + s] = (double)objv;
        }
        __syncwarp();
    }
}

// ---- feasibility ray of ONE scenario (replaces GRB_DoubleAttr_UnbdRay, grb.cpp:304-344) --------
// Cold path: runs once per call that meets an infeasible scenario.  One thread, split graph
// (DESIGN.md §3): minimal min-cut of the lower-bound feasibility network.
__global__ void k1_ray(RayLaunch p) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int m = p.m, nn = p.nn;
    const double *row_u = p.cap_u + (size_t)p.s_local * p.m_pad, *row_l = p.cap_l + (size_t)p.s_local * p.m_pad;
    long long *src = reinterpret_cast<long long *>(p.scratch), *snk = src + nn;   // scratch is 8-byte aligned
    int *cap = reinterpret_cast<int *>(snk + nn), *f = cap + m, *pred = f + m, *queue = pred + nn, *vis = queue + nn, *pot = vis + nn;
    unsigned long long *sums = p.sums;
    auto book = [&](int a, long long val) {  // capacity multiplier of arc a (sigma/phi -> per-arc slot, gamma -> RHS)
        if ((p.arc_info[a] & 3) == KIND_GAMMA) atomicAdd(sums, (unsigned long long)val); else atomicAdd(sums + 1 + p.L + a, (unsigned long long)val);
    };
    for (int a = 0; a < m; a++) {
        const bool closed = p.arc_ts[a] < 0 || p.arc_hs[a] < 0;
        const int ua = (int)row_u[a], la = (int)row_l[a], ut = closed ? 0 : ua;
        if (la > ut) {  // single-arc certificate: beta = 1, capacity multiplier = 1
            atomicAdd(sums, (unsigned long long)(-(long long)la));
            book(a, (long long)ua);
            return;
        }
        cap[a] = ut - la; f[a] = 0;
    }
    for (int v = 0; v < nn; v++) { src[v] = 0; snk[v] = 0; }
    for (int a = 0; a < m; a++) {
        if (p.arc_ts[a] < 0 || p.arc_hs[a] < 0) continue;
        const int la = (int)row_l[a];
        src[p.arc_hs[a]] += la; snk[p.arc_ts[a]] += la;
    }
    for (int v = 0; v < nn; v++) { const long long b = src[v] - snk[v]; src[v] = b > 0 ? b : 0; snk[v] = b < 0 ? -b : 0; }
    for (;;) {
        int qh = 0, qt = 0, found = -1;
        for (int v = 0; v < nn; v++) { vis[v] = 0; if (src[v] > 0) { vis[v] = 1; pred[v] = -1; queue[qt++] = v; } }
        while (qh < qt && found < 0) {
            const int v = queue[qh++];
            if (snk[v] > 0) { found = v; break; }
            for (int a = 0; a < m; a++) {
                const int ts = p.arc_ts[a], hs = p.arc_hs[a];
                if (ts < 0 || hs < 0) continue;
                if (ts == v && f[a] < cap[a] && !vis[hs]) { vis[hs] = 1; pred[hs] = 2 * a; queue[qt++] = hs; }
                if (hs == v && f[a] > 0 && !vis[ts]) { vis[ts] = 1; pred[ts] = 2 * a + 1; queue[qt++] = ts; }
            }
        }
        if (found < 0) break;
        long long d = snk[found];
        int v = found;
        while (pred[v] >= 0) {
            const int a = pred[v] >> 1;
            if (pred[v] & 1) { d = min(d, (long long)f[a]); v = p.arc_hs[a]; } else { d = min(d, (long long)(cap[a] - f[a])); v = p.arc_ts[a]; }
        }
        d = min(d, src[v]);
        src[v] -= d; snk[found] -= d;
        v = found;
        while (pred[v] >= 0) {
            const int a = pred[v] >> 1;
            if (pred[v] & 1) { f[a] -= (int)d; v = p.arc_hs[a]; } else { f[a] += (int)d; v = p.arc_ts[a]; }
        }
    }
    const int shift = vis[0] ? 1 : 0;
    for (int v = 0; v < nn; v++) pot[v] = (vis[v] ? 1 : 0) - shift;
    for (int a = 0; a < m; a++) {
        const int ts = p.arc_ts[a], hs = p.arc_hs[a], q = p.arc_q[a], qt = (q & 0xffff) - 1, qh = (q >> 16) - 1;
        const int alpha_t = (qt >= 0 && p.av_first_wire[qt] >= 0) ? pot[p.av_first_wire[qt]] : 0;
        const int alpha_h = (qh >= 0 && p.av_first_wire[qh] >= 0) ? pot[p.av_first_wire[qh]] : 0;
        const int pt = ts >= 0 ? pot[ts] : alpha_t, ph = hs >= 0 ? pot[hs] : alpha_h;
        const int need = pt - ph;
        const bool closed = ts < 0 || hs < 0;
        if (need > 0) book(a, (long long)(int)row_u[a] * need);
        else if (need < 0 && !closed) atomicAdd(sums, (unsigned long long)(-(long long)(int)row_l[a] * (-need)));
        if (p.arc_pair_layer[a] >= 0) {
            const int dl = pot[hs] - alpha_h;
            if (dl != 0) {
                const long long v = dl > 0 ? (long long)(int)row_u[a] * dl : (long long)(int)row_u[p.arc_next[a]] * (-dl);
                atomicAdd(sums, (unsigned long long)v);
                atomicAdd(sums + 1 + p.arc_pair_layer[a], (unsigned long long)v);
            }
        }
    }
}

// [m][S] int32 (the reference's per-arc vectors) -> [S][m_pad] fp64, 32x32 tiles through shared memory
__global__ void relayout_caps(const int32_t *__restrict__ src, double *__restrict__ dst, int m, int S, int m_pad) {
    __shared__ int tile[32][33];
    const int a0 = blockIdx.y * 32, s0 = blockIdx.x * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int a = a0 + r, s = s0 + threadIdx.x;
        tile[r][threadIdx.x] = (a < m && s < S) ? src[(size_t)a * S + s] : 0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int s = s0 + r, a = a0 + threadIdx.x;
        if (s < S && a < m_pad) dst[(size_t)s * m_pad + a] = (double)tile[threadIdx.x][r];
    }
}

}  // namespace

cudaError_t k1_launch(const K1Launch &p, cudaStream_t st, int sm_count, int *launches) {
    const int words = 2 * p.max_nch + p.max_nopen + 4 * (p.nc + 1) + p.nav + 2;
    const size_t smem = (size_t)WARPS * words * sizeof(int);
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(k1_cut_eval, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k1_cut_eval, WARPS * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const long long items = (long long)p.K * p.S;
    long long want = (items + WARPS - 1) / WARPS;
    long long grid = (long long)sm_count * per_sm;   // persistent: a whole number of CTAs per SM
    if (want < grid) grid = want;
    if (grid < 1) grid = 1;
    k1_cut_eval<<<(unsigned)grid, WARPS * 32, smem, st>>>(p, words);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

cudaError_t ray_launch(const RayLaunch &p, cudaStream_t st, int *launches) {
    k1_ray<<<1, 32, 0, st>>>(p);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

cudaError_t relayout_launch(const int32_t *src, double *dst, int m, int S, int m_pad, cudaStream_t st, int *launches) {
    dim3 grid((S + 31) / 32, (m_pad + 31) / 32), block(32, 8);
    relayout_caps<<<grid, block, 0, st>>>(src, dst, m, S, m_pad);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

}  // namespace sgufp
