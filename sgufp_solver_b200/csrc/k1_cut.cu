// k1_cut.cu — K1: batched per-scenario Benders subproblem evaluation + cut folding (sm_100a).
//
// Replaces the scenario loop of GuroSolver::solveSubProblem (/root/reference/grb.cpp:162-360):
// for every (candidate k, scenario s) it solves the second-stage LP exactly and adds the
// scenario's dual contribution cap x dual (grb.cpp:238-281) into exact integer accumulators.
//
// Mapping: one TILE of 8, 16 or 32 lanes per (candidate, scenario) — small
// contracted graphs put several scenarios in one warp; a persistent grid of
// sm_count x resident-CTAs strides over the K*S work items scenario-minor, so neighbouring tiles
// stream neighbouring rows of the scenario-major fp64 capacity arrays with coalesced 128-bit
// loads.  The candidate's contracted graph (chains + head-sorted residual slots) is
// read-only and stays in L1; all per-scenario state lives in the tile's slice of shared memory.
// No tensor cores: this is integer graph work.
//
// Per work item:
//   1. stream u_s, l_s; segmented min/max into chain capacities (shared-memory atomics, packed
//      with the position of the FIRST least-capacity / LAST greatest-lower-bound arc)
//   2. optimal flow, one shortest-path LEVEL at a time: a pull-style label correction over the
//      head-sorted slots (each lane reduces a contiguous slot range in registers and issues one
//      atomicMin per head it touches; labels are (distance, hops) in one int), then a blocking
//      flow on the tight sub-graph by a backward depth-first search with current-arc pointers
//      (every tight slot is scanned once per level, every path found is pushed immediately)
//   3. SPEC-LP potentials (algorithm-independent) by label correction from the root
//   4. lift to (gamma, beta, sigma, phi, lambda, mu), multiply by the capacities and add to the
//      candidate's accumulators; write objective + status
#include "k1_cut.cuh"

#include <climits>
#include <cstdlib>

#include "model.hpp"

namespace sgufp {

#ifdef SGUFP_K1_STATS
__device__ unsigned long long g_k1_stats[4];   // passes, label computations
#endif

namespace {

constexpr int HB = 10;                 // hop bits of a label
constexpr int LAB_BIAS = 1 << 19;      // |distance| < 2^18 (checked at create) => 0 < distance + bias < 2^20
constexpr int LAB_INF = 0x3fffffff;    // every real label is < 2^30; label + increment never overflows an int
constexpr int NEG_INF = INT_MIN / 4;
constexpr int WARPS = 8;               // warps per CTA

struct PlanView {
    const PlanHeader *h;
    const int32_t *arc_cp, *arc_info, *arc_pre, *ch_ends, *ch_r, *ch_ptr, *ch_arcs, *ch_q, *av_ptr, *av_arcs, *fb_ptr, *fb_ch;
    const int32_t *slot_th, *slot_cs, *slot_ch, *ch_slots, *node_in;
    const int2 *ch_sr;
    const int4 *slot_pk4;          // padded residual slots, two per int4
    const int32_t *node_in4;
    __device__ explicit PlanView(const int32_t *base) {
        h = reinterpret_cast<const PlanHeader *>(base);
        arc_cp = base + h->o_arc_cp; arc_info = base + h->o_arc_info; arc_pre = base + h->o_arc_pre;
        ch_ends = base + h->o_ch_ends; ch_r = base + h->o_ch_r; ch_ptr = base + h->o_ch_ptr; ch_arcs = base + h->o_ch_arcs;
        ch_q = base + h->o_ch_q; av_ptr = base + h->o_av_ptr; av_arcs = base + h->o_av_arcs; fb_ptr = base + h->o_fb_ptr; fb_ch = base + h->o_fb_ch;
        slot_th = base + h->o_slot_th; slot_cs = base + h->o_slot_cs; slot_ch = base + h->o_slot_ch; ch_slots = base + h->o_ch_slots;
        node_in = base + h->o_node_in;
        ch_sr = reinterpret_cast<const int2 *>(base + h->o_ch_sr);
        slot_pk4 = reinterpret_cast<const int4 *>(base + h->o_slot_pk);
        node_in4 = base + h->o_node_in4;
    }
};

struct TileMem {  // this tile's slice of shared memory
    int *up, *lo, *x, *res, *lab, *pred, *pot, *exc, *aq;
};

template <int TILE>
struct Lanes {    // the TILE lanes that work on one scenario
    unsigned mask;
    int tl;
    __device__ Lanes() {
        const int lane = threadIdx.x & 31;
        tl = lane & (TILE - 1);
        mask = TILE == 32 ? 0xffffffffu : (((1u << TILE) - 1u) << (lane & ~(TILE - 1)));
    }
    __device__ __forceinline__ void sync() const { __syncwarp(mask); }
    __device__ __forceinline__ bool any(bool p) const { return __any_sync(mask, p) != 0; }
    __device__ __forceinline__ long long sum(long long v) const {
        for (int o = TILE / 2; o; o >>= 1) v += __shfl_xor_sync(mask, v, o, TILE);
        return v;
    }
    __device__ __forceinline__ unsigned long long min_u64(unsigned long long v) const {
        for (int o = TILE / 2; o; o >>= 1) { const unsigned long long t = __shfl_xor_sync(mask, v, o, TILE); v = t < v ? t : v; }
        return v;
    }
};

__device__ __forceinline__ int lab_dist(int lab) { return (lab >> HB) - LAB_BIAS; }

// Label-correcting shortest paths from `src` over the residual arcs of the contracted graph
// (push style: every lane relaxes its chains, shared-memory atomicMin on the packed label).
// MERGED: the root is one node (index 0) that is never relabelled; otherwise arcs entering the
// root end at index nc.
template <int TILE, bool MERGED>
__device__ void shortest_paths(int src, const PlanView &P, int nopen, int nc, const Lanes<TILE> &T, TileMem &w, int &fuel) {
    for (int v = T.tl; v <= nc; v += TILE) w.lab[v] = LAB_INF;
    T.sync();
    if (T.tl == 0) w.lab[src] = LAB_BIAS << HB;
    T.sync();
    bool changed;
    do {
        changed = false;
        for (int c = T.tl; c < nopen; c += TILE) {
            const int2 sr = P.ch_sr[c];            // static: ends | reward, one 64-bit load
            const int f = w.res[c];                // dynamic: bit 0 forward residual, bit 1 backward residual
            const int sv = (sr.x & 0xffff) - 1, ev = (sr.x >> 16) - 1;
            if ((f & 1) && !(MERGED && ev == 0)) {
                const int lu = w.lab[sv];
                if (lu != LAB_INF) {
                    const int cand = lu - sr.y * (1 << HB) + 1;
                    const int t = (!MERGED && ev == 0) ? nc : ev;
                    if (cand < atomicMin(&w.lab[t], cand)) { changed = true; w.pred[t] = 2 * c; }
                }
            }
            if ((f & 2) && !(MERGED && sv == 0)) {
                const int lu = w.lab[ev];
                if (lu != LAB_INF) {
                    const int cand = lu + sr.y * (1 << HB) + 1;
                    const int t = (!MERGED && sv == 0) ? nc : sv;
                    if (cand < atomicMin(&w.lab[t], cand)) { changed = true; w.pred[t] = 2 * c + 1; }
                }
            }
        }
        T.sync();
        changed = T.any(changed) && --fuel > 0;
#ifdef SGUFP_K1_STATS
        if (T.tl == 0) atomicAdd(&g_k1_stats[0], 1ull);
#endif
    } while (changed);
#ifdef SGUFP_K1_STATS
    if (T.tl == 0) atomicAdd(&g_k1_stats[1], 1ull);
#endif
}

// One tight residual arc per labelled node (pred = 2*chain + direction).  Only needed when the
// predecessor written next to a label update lost a race with a better update of the same pass.
template <int TILE>
__device__ void mark_predecessors(const PlanView &P, int nopen, int nc, const Lanes<TILE> &T, TileMem &w) {
    for (int c = T.tl; c < nopen; c += TILE) {
        const int2 sr = P.ch_sr[c];
        const int f = w.res[c];
        const int sv = (sr.x & 0xffff) - 1, ev = (sr.x >> 16) - 1;
        if (f & 1) {
            const int lu = w.lab[sv], t = ev == 0 ? nc : ev;
            if (lu != LAB_INF && lu - sr.y * (1 << HB) + 1 == w.lab[t]) w.pred[t] = 2 * c;
        }
        if (f & 2) {
            const int lu = w.lab[ev], t = sv == 0 ? nc : sv;
            if (lu != LAB_INF && lu + sr.y * (1 << HB) + 1 == w.lab[t]) w.pred[t] = 2 * c + 1;
        }
    }
    T.sync();
}

// Lane 0 of the tile: follow the predecessors dst -> src, then push the bottleneck.  A label is
// (distance, hops), so a predecessor is usable iff its arc still has residual capacity and is
// TIGHT (label(tail) + increment == label(head)); the hop count then drops by one per step and the
// walk can only end at src.  Returns the amount pushed, or -1 if some predecessor is not tight
// (stale, or overwritten by a racing lane): the caller re-marks and walks again.
__device__ int augment(int src, int dst, int limit, const PlanView &P, int nc, TileMem &w) {
    int v = dst, d = limit;
    while (v != src) {
        const int p = w.pred[v], c = p >> 1;
        if ((unsigned)c >= (unsigned)P.h->nopen) return -1;
        const int2 sr = P.ch_sr[c];
        const int sv = (sr.x & 0xffff) - 1, ev = (sr.x >> 16) - 1;
        int u, res, inc;
        int head;   // pred[v] may be a leftover of another candidate's plan: the arc must really end at v
        if (p & 1) { u = ev; head = sv == 0 ? nc : sv; res = w.x[c] - (w.lo[c] >> HB); inc = sr.y * (1 << HB) + 1; }
        else { u = sv; head = ev == 0 ? nc : ev; res = (w.up[c] >> HB) - w.x[c]; inc = -sr.y * (1 << HB) + 1; }
        if (head != v || res <= 0 || w.lab[u] == LAB_INF || w.lab[u] + inc != w.lab[v]) return -1;
        d = min(d, res);
        v = u;
    }
    v = dst;
    while (v != src) {
        const int p = w.pred[v], c = p >> 1;
        const int2 sr = P.ch_sr[c];
        int xc;
        if (p & 1) { xc = (w.x[c] -= d); v = (sr.x >> 16) - 1; } else { xc = (w.x[c] += d); v = (sr.x & 0xffff) - 1; }
        w.res[c] = (xc < (w.up[c] >> HB) ? 1 : 0) | (xc > (w.lo[c] >> HB) ? 2 : 0);
    }
    return d;
}

// augment, falling back to one marking pass when a captured predecessor is not usable
template <int TILE>
__device__ void push_path(int src, int dst, int limit, const PlanView &P, int nopen, int nc, const Lanes<TILE> &T, TileMem &w, int *pushed) {
    int d = 0;
    if (T.tl == 0) d = augment(src, dst, limit, P, nc, w);
    d = __shfl_sync(T.mask, d, 0, TILE);
    if (d < 0) {
        mark_predecessors<TILE>(P, nopen, nc, T, w);
        if (T.tl == 0) d = augment(src, dst, limit, P, nc, w);
        d = __shfl_sync(T.mask, d, 0, TILE);
    }
    T.sync();
    if (pushed) *pushed = d;
}

// Forced flow from lower bounds (rare): route every excess / deficit along shortest residual
// paths.  Returns false if some forced flow cannot be routed (scenario infeasible).
template <int TILE>
__device__ bool route_lower_bounds(const PlanView &P, int nopen, int nc, const Lanes<TILE> &T, TileMem &w, int &fuel) {
    for (int v = T.tl; v <= nc; v += TILE) w.exc[v] = 0;
    T.sync();
    for (int c = T.tl; c < nopen; c += TILE) {
        const int lo = w.lo[c] >> HB;
        if (lo > 0) {
            const int e = P.ch_ends[c], sv = (e & 0xffff) - 1, ev = (e >> 16) - 1;
            if (ev > 0) atomicAdd(&w.exc[ev], lo);
            if (sv > 0) atomicSub(&w.exc[sv], lo);
        }
    }
    T.sync();
    for (int v = 1; v < nc; v++) {
        while (w.exc[v] > 0 && --fuel > 0) {
            shortest_paths<TILE, false>(v, P, nopen, nc, T, w, fuel);
            unsigned long long best = ~0ull;
            for (int t = T.tl; t <= nc; t += TILE)
                if ((t == nc || (t > 0 && w.exc[t] < 0)) && w.lab[t] != LAB_INF) {
                    const unsigned long long key = ((unsigned long long)(unsigned)w.lab[t] << 32) | (unsigned)t;
                    best = key < best ? key : best;
                }
            best = T.min_u64(best);
            if (best == ~0ull) return false;
            const int t = (int)(best & 0xffffffffu);
            int lim = w.exc[v], d = 0;
            if (t != nc) lim = min(lim, -w.exc[t]);
            push_path<TILE>(v, t, lim, P, nopen, nc, T, w, &d);
            if (T.tl == 0 && d > 0) { w.exc[v] -= d; if (t != nc) w.exc[t] += d; }
            T.sync();
            if (d <= 0) return false;
        }
    }
    for (int v = 1; v < nc; v++) {
        while (w.exc[v] < 0 && --fuel > 0) {
            shortest_paths<TILE, false>(0, P, nopen, nc, T, w, fuel);
            if (w.lab[v] == LAB_INF) return false;
            int d = 0;
            push_path<TILE>(0, v, -w.exc[v], P, nopen, nc, T, w, &d);
            if (T.tl == 0 && d > 0) w.exc[v] += d;
            T.sync();
            if (d <= 0) return false;
        }
    }
    return true;
}

// SPEC-LP potentials: pot[v] = -(shortest residual distance from the root);
// nodes the root cannot reach get the least labels consistent with the labelled ones; nodes cut
// off both ways get a zero-rooted completion.
template <int TILE>
__device__ void canonical_potentials(const PlanView &P, int nopen, int nc, const Lanes<TILE> &T, TileMem &w, int &fuel) {
    shortest_paths<TILE, true>(0, P, nopen, nc, T, w, fuel);
    bool missing = false;
    for (int v = T.tl; v < nc; v += TILE) {
        const int l = w.lab[v];
        if (l == LAB_INF) { missing = true; w.pot[v] = NEG_INF; w.pred[v] = 0; } else { w.pot[v] = lab_dist(l); w.pred[v] = 1; }
    }
    T.sync();
    if (T.any(missing)) {
        // pot[] holds d here; cur[] is the state: 1 labelled by phase 1, 0 not yet, 2 isolated
        bool changed;
        do {
            changed = false;
            for (int c = T.tl; c < nopen; c += TILE) {
                const int e = P.ch_ends[c], sv = (e & 0xffff) - 1, ev = (e >> 16) - 1, r = P.ch_r[c];
                const int xc = w.x[c], up = w.up[c] >> HB, lo = w.lo[c] >> HB;
                if (xc < up && w.pred[sv] == 0) { const int db = w.pot[ev]; if (db > NEG_INF && db + r > atomicMax(&w.pot[sv], db + r)) changed = true; }
                if (xc > lo && w.pred[ev] == 0) { const int da = w.pot[sv]; if (da > NEG_INF && da - r > atomicMax(&w.pot[ev], da - r)) changed = true; }
            }
            T.sync();
            changed = T.any(changed) && --fuel > 0;
        } while (changed);
        bool iso = false;
        for (int v = T.tl; v < nc; v += TILE)
            if (w.pot[v] == NEG_INF) { w.pot[v] = 0; w.pred[v] = 2; iso = true; }
        T.sync();
        if (T.any(iso)) {
            do {
                changed = false;
                for (int c = T.tl; c < nopen; c += TILE) {
                    const int e = P.ch_ends[c], sv = (e & 0xffff) - 1, ev = (e >> 16) - 1, r = P.ch_r[c];
                    const int xc = w.x[c], up = w.up[c] >> HB, lo = w.lo[c] >> HB;
                    if (xc < up && w.pred[ev] == 2) { const int cand = w.pot[sv] - r; if (cand < atomicMin(&w.pot[ev], cand)) changed = true; }
                    if (xc > lo && w.pred[sv] == 2) { const int cand = w.pot[ev] + r; if (cand < atomicMin(&w.pot[sv], cand)) changed = true; }
                }
                T.sync();
                changed = T.any(changed) && --fuel > 0;
            } while (changed);
        }
    }
    for (int v = T.tl; v < nc; v += TILE) w.pot[v] = -w.pot[v];
    T.sync();
}

// wire potential at the HEAD of arc a (a matched in-arc, or any arc of a chain)
__device__ __forceinline__ int head_potential(int a, const PlanView &P, int nopen, const TileMem &w) {
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

template <int TILE, int NW>
__global__ void __launch_bounds__(NW * 32, NW == 8 ? 4 : 1) k1_cut_eval(K1Launch p, int words_per_tile) {
#ifdef SGUFP_K1_EMULATE
    int *smem = sgufp_emul_smem;   // tests/cpp/k1_emul.cpp: the kernel body compiled for the host, TILE = 1
#else
    extern __shared__ int smem[];
#endif
    const Lanes<TILE> T;
    constexpr int TILES_PER_CTA = NW * 32 / TILE;
    const int tile_in_cta = threadIdx.x / TILE;
    TileMem w;
    {
        int *base = smem + (size_t)tile_in_cta * words_per_tile;
        w.up = base; base += p.max_nch;
        w.lo = base; base += p.max_nch;
        w.x = base; base += p.max_nopen;
        w.res = base; base += p.max_nopen;
        w.lab = base; base += p.nc + 2;
        w.pred = base; base += p.nc + 2;
        w.pot = base; base += p.nc + 2;
        w.exc = base; base += p.nc + 2;
        w.aq = base;
    }
    const long long items = (long long)p.K * p.S;
    const long long stride = (long long)gridDim.x * TILES_PER_CTA;
    for (long long item = (long long)blockIdx.x * TILES_PER_CTA + tile_in_cta; item < items; item += stride) {
        const int k = (int)(item / p.S), s = (int)(item - (long long)k * p.S);
        if (*reinterpret_cast<volatile long long *>(p.first_inf + k) < 0) continue;   // this candidate was aborted: drain
        int fuel = 1 << 20;   // passes + levels a work item may spend (tile-uniform)
        const PlanView P(p.plans + p.plan_off[k]);
        const int nch = P.h->nch, nopen = P.h->nopen, nc = P.h->nc, nav = P.h->nav, m = p.m;
        unsigned long long *sums = p.sums + (size_t)k * p.W;
        const double *row_u = p.cap_u + (size_t)s * p.m_pad, *row_l = p.cap_l + (size_t)s * p.m_pad;

        // 1. chain capacities
        for (int c = T.tl; c < nch; c += TILE) { w.up[c] = INT_MAX; w.lo[c] = 0; }
        T.sync();
        {
            const double2 *ru = reinterpret_cast<const double2 *>(row_u), *rl = reinterpret_cast<const double2 *>(row_l);
            for (int a2 = T.tl; a2 < p.m_pad / 2; a2 += TILE) {
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
        T.sync();
        bool bad = false, forced = false;
        for (int c = T.tl; c < nch; c += TILE) {
            const int lo = w.lo[c] >> HB;
            if (c < nopen) {
                const int up = w.up[c] >> HB;
                w.x[c] = lo; bad |= lo > up; forced |= lo > 0;
                w.res[c] = lo < up ? 1 : 0;          // x == lo: forward residual only
            } else bad |= lo > 0;
        }
        T.sync();
        bad = T.any(bad);
        forced = T.any(forced);
        // 2. optimal flow
        if (!bad && forced) bad = !route_lower_bounds<TILE>(P, nopen, nc, T, w, fuel);
        if (bad) {
            if (T.tl == 0) {
                atomicMin(p.first_inf + k, p.scen_offset + s);
                if (p.status) p.status[(size_t)k * p.S + s] = 1;
                if (p.obj) p.obj[(size_t)k * p.S + s] = 0.0;
            }
            T.sync();
            continue;
        }
        while (fuel > 0) {   // one label computation per iteration, then as many pushes as its tight arcs allow
            shortest_paths<TILE, false>(0, P, nopen, nc, T, w, fuel);
            const int lt = w.lab[nc];
            if (lt == LAB_INF || lab_dist(lt) >= 0) break;
            int d = 0;
            push_path<TILE>(0, nc, INT_MAX, P, nopen, nc, T, w, &d);
            if (d <= 0) fuel = 0;   // cannot happen: the labels were just computed
        }
        // 3. potentials
        canonical_potentials<TILE>(P, nopen, nc, T, w, fuel);
        if (fuel <= 0) {   // a bound that no valid instance reaches: refuse to answer rather than spin
            if (T.tl == 0) atomicMin(p.first_inf + k, -1LL);
            T.sync();
            continue;
        }
        // 4. lifting + folding
        long long rhs = 0, objv = 0;
        for (int c = T.tl; c < nopen; c += TILE) {
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
        for (int i = T.tl; i < nav; i += TILE) {
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
        T.sync();
        for (int c = nopen + T.tl; c < nch; c += TILE) {
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
        rhs = T.sum(rhs);
        objv = T.sum(objv);
        if (T.tl == 0) {
            if (rhs) atomicAdd(sums, (unsigned long long)rhs);
            if (p.status) p.status[(size_t)k * p.S + s] = 0;
            if (p.obj) p.obj[(size_t)k * p.S + s] = (double)objv;
        }
        T.sync();
    }
}

// =================================================================================================
// Lane-per-scenario variant (small contracted graphs).
//
// One warp works on LW = 32 consecutive scenarios of ONE candidate: lane = scenario.  All lanes walk
// the candidate's chains in the same order, so the static half of every chain (ends, reward) is
// loaded once per warp-instruction and serves 32 scenarios, and every piece of per-scenario state
// sits in shared memory as a column [index][lane] — each lane only ever touches its own bank:
// no conflicts, no atomics, and label updates are sequential per lane (Gauss-Seidel in the
// topological order of the chains, exact predecessors).  Lanes differ only in data, never in the
// program counter, except for the per-lane path walks.  Same SPEC-LP results as the warp variant.
// =================================================================================================
struct LaneMem {   // column-major slices of this warp's shared memory, LW lanes wide
    unsigned short *x, *up, *lo, *pred;   // [nopen], [nopen], [nopen], [nc+1]
    unsigned char *res;                   // [nopen] residual flags: bit 0 forward, bit 1 backward
    int *lab, *exc, *aq;                  // [nc+1], [nc+1], [nav]   (lab doubles as the potential after the flow phase)
};

template <int LW>
struct LaneCtx {
    int lane;
    unsigned mask;
    __device__ LaneCtx() { lane = LW == 1 ? 0 : (int)(threadIdx.x & 31); mask = 0xffffffffu; }
    __device__ __forceinline__ bool any(bool p) const { return LW == 1 ? p : __any_sync(mask, p) != 0; }
    __device__ __forceinline__ long long sum(long long v) const {
        if (LW > 1) for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(mask, v, o);
        return v;
    }
};

#define LX(a, i) (a)[(i) * LW + L.lane]

// Label correction from a per-lane source, PULL style and sequential per lane: nodes are visited in
// topological order, each takes the minimum over its in-slots (the loads of one node are independent
// of each other: that is the instruction-level parallelism of this kernel) and keeps the slot that
// gave it as its exact predecessor.  Residual status comes from one flag byte per chain.
template <int LW, bool MERGED>
__device__ void lane_shortest_paths(int src, const PlanView &P, int nc, const LaneCtx<LW> &L, LaneMem &w, int &fuel) {
    for (int v = 0; v <= nc + 1; v++) LX(w.lab, v) = LAB_INF;      // index nc+1: the tail of padding slots, never reached
    LX(w.lab, src) = LAB_BIAS << HB;
    const int vend = MERGED ? nc : nc + 1;
    bool changed;
    do {
        changed = false;
        for (int v = 1; v < vend; v++) {
            const int s1 = P.node_in4[v + 1];
            int best = LX(w.lab, v), bs = -1;
            for (int sl = P.node_in4[v]; sl < s1; sl += 4) {          // slot lists are padded to a multiple of 4
                const int4 q0 = P.slot_pk4[sl >> 1], q1 = P.slot_pk4[(sl >> 1) + 1];   // {tail | (2*chain+dir) << 16, increment} x 4
                const int d0 = q0.x >> 16, d1 = q0.z >> 16, d2 = q1.x >> 16, d3 = q1.z >> 16;
                const int f0 = LX(w.res, d0 >> 1), f1 = LX(w.res, d1 >> 1), f2 = LX(w.res, d2 >> 1), f3 = LX(w.res, d3 >> 1);
                const int l0 = LX(w.lab, q0.x & 0xffff), l1 = LX(w.lab, q0.z & 0xffff), l2 = LX(w.lab, q1.x & 0xffff), l3 = LX(w.lab, q1.z & 0xffff);
                const int c0 = l0 + q0.y, c1 = l1 + q0.w, c2 = l2 + q1.y, c3 = l3 + q1.w;
                if (((f0 >> (d0 & 1)) & 1) && l0 != LAB_INF && c0 < best) { best = c0; bs = sl; }
                if (((f1 >> (d1 & 1)) & 1) && l1 != LAB_INF && c1 < best) { best = c1; bs = sl + 1; }
                if (((f2 >> (d2 & 1)) & 1) && l2 != LAB_INF && c2 < best) { best = c2; bs = sl + 2; }
                if (((f3 >> (d3 & 1)) & 1) && l3 != LAB_INF && c3 < best) { best = c3; bs = sl + 3; }
            }
            if (bs >= 0) { LX(w.lab, v) = best; LX(w.pred, v) = (unsigned short)bs; changed = true; }
        }
        changed = L.any(changed) && --fuel > 0;
    } while (changed);
}

// per-lane walk dst -> src along the exact predecessor slots, then push; lanes with go == false idle
template <int LW>
__device__ int lane_push(bool go, int src, int dst, int limit, const PlanView &P, const LaneCtx<LW> &L, LaneMem &w) {
    int v = dst, d = limit;
    while (L.any(go && v != src)) {
        if (go && v != src) {
            const int2 pk = reinterpret_cast<const int2 *>(P.slot_pk4)[LX(w.pred, v)];
            const int cd = pk.x >> 16, c = cd >> 1;
            d = min(d, (cd & 1) ? (int)LX(w.x, c) - (int)LX(w.lo, c) : (int)LX(w.up, c) - (int)LX(w.x, c));
            v = pk.x & 0xffff;
        }
    }
    v = dst;
    while (L.any(go && v != src)) {
        if (go && v != src) {
            const int2 pk = reinterpret_cast<const int2 *>(P.slot_pk4)[LX(w.pred, v)];
            const int cd = pk.x >> 16, c = cd >> 1;
            const int xc = (int)LX(w.x, c) + ((cd & 1) ? -d : d);
            LX(w.x, c) = (unsigned short)xc;
            LX(w.res, c) = (unsigned char)((xc < (int)LX(w.up, c) ? 1 : 0) | (xc > (int)LX(w.lo, c) ? 2 : 0));
            v = pk.x & 0xffff;
        }
    }
    return go ? d : 0;
}

// potential at the HEAD of arc a for this lane's scenario (cf. head_potential of the warp variant);
// the position tests "pos >= first least-capacity arc" / "pos >= last greatest-lower-bound arc" are
// answered by re-reading the few capacities of the chain
template <int LW>
__device__ int lane_head_potential(int a, const PlanView &P, int nopen, const double *row_u, const double *row_l, const LaneCtx<LW> &L, const LaneMem &w) {
    const int cp = P.arc_cp[a], c = cp >> 10, pos = cp & 1023, pre = P.arc_pre[a];
    const int e = P.ch_ends[c], sv = (e & 0xffff) - 1, ev = (e >> 16) - 1;
    if (c < nopen) {
        const int r = P.ch_r[c], psv = LX(w.lab, sv), dp = LX(w.lab, ev) - psv;
        const int g = max(0, r - dp), b = max(0, dp - r);
        int val = psv + pre;
        const int b0 = P.ch_ptr[c], b1 = P.ch_ptr[c + 1];
        if (L.any(g > 0)) {
            bool found = false;
            const int up = LX(w.up, c);
            for (int k = b0; k <= b0 + pos; k++) found |= (int)row_u[P.ch_arcs[k]] == up;
            if (found) val -= g;
        }
        if (L.any(b > 0)) {
            bool later = false;
            const int lo = LX(w.lo, c);
            for (int k = b0 + pos + 1; k < b1; k++) later |= (int)row_l[P.ch_arcs[k]] == lo;
            if (!later) val += b;
        }
        return val;
    }
    if (sv >= 0) return LX(w.lab, sv) + pre;
    if (ev >= 0) return LX(w.lab, ev) - (P.ch_r[c] - pre);
    return pre;
}

template <int LW>
__global__ void __launch_bounds__(32) k1_lane_eval(K1Launch p, int bytes_per_warp) {
#ifdef SGUFP_K1_EMULATE
    unsigned char *smem = reinterpret_cast<unsigned char *>(sgufp_emul_smem);
#else
    extern __shared__ unsigned char smem_l[];
    unsigned char *smem = smem_l;
#endif
    const LaneCtx<LW> L;
    const int nc_max = p.nc + 2;
    LaneMem w;
    {
        unsigned char *base = smem;   // one warp per CTA
        (void)bytes_per_warp;
        w.lab = reinterpret_cast<int *>(base); base += (size_t)nc_max * LW * 4;
        w.exc = reinterpret_cast<int *>(base); base += (size_t)nc_max * LW * 4;
        w.aq = reinterpret_cast<int *>(base); base += (size_t)(p.nav + 1) * LW * 4;
        w.x = reinterpret_cast<unsigned short *>(base); base += (size_t)p.max_nopen * LW * 2;
        w.up = reinterpret_cast<unsigned short *>(base); base += (size_t)p.max_nopen * LW * 2;
        w.lo = reinterpret_cast<unsigned short *>(base); base += (size_t)p.max_nopen * LW * 2;
        w.pred = reinterpret_cast<unsigned short *>(base); base += (size_t)nc_max * LW * 2;
        w.res = base;
    }
    const int bps = (p.S + LW - 1) / LW;                     // scenario blocks per candidate
    const long long nblk = (long long)p.K * bps;
    for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const int k = (int)(blk / bps), s_raw = (int)(blk - (long long)k * bps) * LW + L.lane;
        const bool active = s_raw < p.S;
        const int s = active ? s_raw : p.S - 1;             // idle lanes shadow the last scenario; they never write results
        if (*reinterpret_cast<volatile long long *>(p.first_inf + k) < 0) continue;
        int fuel = 1 << 20;
        const PlanView P(p.plans + p.plan_off[k]);
        const int nch = P.h->nch, nopen = P.h->nopen, nc = P.h->nc, nav = P.h->nav, m = p.m;
        unsigned long long *sums = p.sums + (size_t)k * p.W;
        const double *row_u = p.cap_u + (size_t)s * p.m_pad, *row_l = p.cap_l + (size_t)s * p.m_pad;

        // 1. chain capacities, sequential per lane
        for (int c = 0; c < nopen; c++) { LX(w.up, c) = 0xffff; LX(w.lo, c) = 0; }
        bool bad = false;
        for (int a = 0; a < m; a++) {
            const int c = P.arc_cp[a] >> 10;
            const int ua = (int)row_u[a], la = (int)row_l[a];
            if (c < nopen) {
                if (ua < (int)LX(w.up, c)) LX(w.up, c) = (unsigned short)ua;
                if (la > (int)LX(w.lo, c)) LX(w.lo, c) = (unsigned short)la;
            } else bad |= la > 0;
        }
        bool forced = false;
        for (int c = 0; c < nopen; c++) {
            const int lo = LX(w.lo, c), up = LX(w.up, c);
            LX(w.x, c) = (unsigned short)lo;
            LX(w.res, c) = (unsigned char)(lo < up ? 1 : 0);
            bad |= lo > up; forced |= lo > 0;
        }
        // 2a. forced flow of the lower bounds (rare)
        if (L.any(forced && !bad)) {
            for (int v = 0; v <= nc; v++) LX(w.exc, v) = 0;
            for (int c = 0; c < nopen; c++) {
                const int lo = LX(w.lo, c), e = P.ch_ends[c], sv = (e & 0xffff) - 1, ev = (e >> 16) - 1;
                if (lo > 0 && !bad) { if (ev > 0) LX(w.exc, ev) += lo; if (sv > 0) LX(w.exc, sv) -= lo; }
            }
            for (;;) {   // excess nodes: to the root or to a deficit node, whichever is nearer
                int src = -1;
                for (int v = 1; v < nc; v++) if (src < 0 && LX(w.exc, v) > 0) src = v;
                const bool live = !bad && src >= 0;
                if (!L.any(live) || fuel <= 0) break;
                lane_shortest_paths<LW, false>(live ? src : 0, P, nc, L, w, fuel);
                int best = -1, bl = LAB_INF;
                for (int t = 1; t <= nc; t++) {
                    const int lt = LX(w.lab, t);
                    if ((t == nc || LX(w.exc, t) < 0) && lt < bl) { bl = lt; best = t; }
                }
                if (live && best < 0) bad = true;
                const bool go = live && !bad;
                int lim = go ? LX(w.exc, src) : 0;
                if (go && best != nc) lim = min(lim, -LX(w.exc, best));
                const int d = lane_push<LW>(go, src, best, lim, P, L, w);
                if (go) { LX(w.exc, src) -= d; if (best != nc) LX(w.exc, best) += d; }
                fuel--;
            }
            for (;;) {   // deficit nodes: from the root
                int dst = -1;
                for (int v = 1; v < nc; v++) if (dst < 0 && LX(w.exc, v) < 0) dst = v;
                const bool live = !bad && dst >= 0;
                if (!L.any(live) || fuel <= 0) break;
                lane_shortest_paths<LW, false>(0, P, nc, L, w, fuel);
                if (live && LX(w.lab, dst) == LAB_INF) bad = true;
                const bool go = live && !bad;
                const int d = lane_push<LW>(go, 0, go ? dst : 0, go ? -LX(w.exc, dst) : 0, P, L, w);
                if (go) LX(w.exc, dst) += d;
                fuel--;
            }
        }
        // 2b. successive shortest paths, all lanes in step
        while (fuel > 0) {
            lane_shortest_paths<LW, false>(0, P, nc, L, w, fuel);
            const int lt = LX(w.lab, nc);
            const bool go = !bad && lt != LAB_INF && lab_dist(lt) < 0;
            if (!L.any(go)) break;
            lane_push<LW>(go, 0, nc, INT_MAX, P, L, w);
            fuel--;
        }
        // 3. potentials (lab becomes the potential)
        lane_shortest_paths<LW, true>(0, P, nc, L, w, fuel);
        bool missing = false;
        for (int v = 0; v < nc; v++) {
            const int l = LX(w.lab, v);
            if (l == LAB_INF) { missing = true; LX(w.lab, v) = NEG_INF; LX(w.pred, v) = 0; } else { LX(w.lab, v) = lab_dist(l); LX(w.pred, v) = 1; }
        }
        if (L.any(missing)) {
            bool changed;
            do {   // least labels consistent with the labelled nodes
                changed = false;
                for (int c = 0; c < nopen; c++) {
                    const int e = P.ch_ends[c], sv = (e & 0xffff) - 1, ev = (e >> 16) - 1, r = P.ch_r[c];
                    const int xc = LX(w.x, c), up = LX(w.up, c), lo = LX(w.lo, c);
                    if (xc < up && LX(w.pred, sv) == 0) { const int db = LX(w.lab, ev); if (db > NEG_INF && db + r > LX(w.lab, sv)) { LX(w.lab, sv) = db + r; changed = true; } }
                    if (xc > lo && LX(w.pred, ev) == 0) { const int da = LX(w.lab, sv); if (da > NEG_INF && da - r > LX(w.lab, ev)) { LX(w.lab, ev) = da - r; changed = true; } }
                }
                changed = L.any(changed) && --fuel > 0;
            } while (changed);
            bool iso = false;
            for (int v = 0; v < nc; v++) if (LX(w.lab, v) == NEG_INF) { LX(w.lab, v) = 0; LX(w.pred, v) = 2; iso = true; }
            if (L.any(iso)) {
                do {   // zero-rooted completion of the nodes cut off both ways
                    changed = false;
                    for (int c = 0; c < nopen; c++) {
                        const int e = P.ch_ends[c], sv = (e & 0xffff) - 1, ev = (e >> 16) - 1, r = P.ch_r[c];
                        const int xc = LX(w.x, c), up = LX(w.up, c), lo = LX(w.lo, c);
                        if (xc < up && LX(w.pred, ev) == 2) { const int cand = LX(w.lab, sv) - r; if (cand < LX(w.lab, ev)) { LX(w.lab, ev) = cand; changed = true; } }
                        if (xc > lo && LX(w.pred, sv) == 2) { const int cand = LX(w.lab, ev) + r; if (cand < LX(w.lab, sv)) { LX(w.lab, sv) = cand; changed = true; } }
                    }
                    changed = L.any(changed) && --fuel > 0;
                } while (changed);
            }
        }
        for (int v = 0; v < nc; v++) LX(w.lab, v) = -LX(w.lab, v);
        if (fuel <= 0) {
            if (L.lane == 0) atomicMin(p.first_inf + k, -1LL);
            continue;
        }
        const bool ok = active && !bad;      // lanes whose scenario contributes to the cut
        // 4. lifting + folding: one reduction over the 32 scenarios per accumulator touched
        long long rhs = 0, objv = 0;
        for (int c = 0; c < nopen; c++) {
            const int e = P.ch_ends[c], sv = (e & 0xffff) - 1, ev = (e >> 16) - 1, r = P.ch_r[c];
            const int dp = LX(w.lab, ev) - LX(w.lab, sv), g = r - dp;
            const int up = LX(w.up, c), lo = LX(w.lo, c);
            objv += (long long)r * (int)LX(w.x, c);
            if (ok && g < 0) rhs -= (long long)lo * (-g);
            if (L.any(ok && g > 0)) {
                bool found = false;
                for (int kk = P.ch_ptr[c]; kk < P.ch_ptr[c + 1]; kk++) {   // the FIRST arc of least capacity carries the multiplier
                    const int a = P.ch_arcs[kk];
                    const bool hit = ok && g > 0 && !found && (int)row_u[a] == up;
                    found |= hit;
                    const long long tot = L.sum(hit ? (long long)up * g : 0);
                    if (tot != 0 && L.lane == 0) {
                        if ((P.arc_info[a] & 3) == KIND_GAMMA) atomicAdd(sums, (unsigned long long)tot);
                        else atomicAdd(sums + 1 + p.L + a, (unsigned long long)tot);
                    }
                }
            }
        }
        for (int i = 0; i < nav; i++) {
            const int b0 = P.av_ptr[i], b1 = P.av_ptr[i + 1];
            int alpha = 0;
            if (b1 > b0) {
                alpha = lane_head_potential<LW>(P.av_arcs[b0], P, nopen, row_u, row_l, L, w);
                for (int t = b0 + 1; t < b1; t++) {
                    const int a = P.av_arcs[t];
                    const int dl = lane_head_potential<LW>(a, P, nopen, row_u, row_l, L, w) - alpha;
                    if (L.any(ok && dl != 0)) {
                        long long v = 0;
                        if (ok && dl > 0) v = (long long)(int)row_u[a] * dl;
                        else if (ok && dl < 0) {
                            const int cp = P.arc_cp[a];
                            v = (long long)(int)row_u[P.ch_arcs[P.ch_ptr[cp >> 10] + (cp & 1023) + 1]] * (-dl);
                        }
                        const long long tot = L.sum(v);
                        rhs += v;
                        if (tot != 0 && L.lane == 0) atomicAdd(sums + 1 + (P.arc_info[a] >> 2) - 1, (unsigned long long)tot);
                    }
                }
            } else {
                bool found = false;
                for (int t = P.fb_ptr[i]; t < P.fb_ptr[i + 1]; t++) {
                    const int c = P.fb_ch[t], ev = (P.ch_ends[c] >> 16) - 1;
                    const int cand = LX(w.lab, ev) - P.ch_r[c];
                    if (!found || cand < alpha) { alpha = cand; found = true; }
                }
            }
            LX(w.aq, i) = alpha;
        }
        for (int c = nopen; c < nch; c++) {
            const int b0 = P.ch_ptr[c], b1 = P.ch_ptr[c + 1], first = P.ch_arcs[b0], last = P.ch_arcs[b1 - 1];
            const int e = P.ch_ends[c], sv = (e & 0xffff) - 1, ev = (e >> 16) - 1;
            const int q = P.ch_q[c], qs = (q & 0xffff) - 1, qe = (q >> 16) - 1;
            const int rfirst = P.arc_pre[first];
            if (qs >= 0 && qe >= 0 && b1 - b0 == 1) {
                const int v = rfirst - (LX(w.aq, qe) - LX(w.aq, qs));
                const long long tot = L.sum(ok && v > 0 ? (long long)(int)row_u[first] * v : 0);
                if (tot != 0 && L.lane == 0) atomicAdd(sums + 1 + p.L + first, (unsigned long long)tot);
                continue;
            }
            if (qs >= 0) {
                const int phf = ev >= 0 ? LX(w.lab, ev) - (P.ch_r[c] - rfirst) : rfirst;
                const int v = rfirst - (phf - LX(w.aq, qs));
                const long long tot = L.sum(ok && v > 0 ? (long long)(int)row_u[first] * v : 0);
                if (tot != 0 && L.lane == 0) atomicAdd(sums + 1 + p.L + first, (unsigned long long)tot);
            }
            if (qe >= 0) {
                int pt, rlast;
                if (b1 - b0 > 1) {
                    const int prev = P.ch_arcs[b1 - 2];
                    pt = (sv >= 0 ? LX(w.lab, sv) : 0) + P.arc_pre[prev];
                    rlast = P.ch_r[c] - P.arc_pre[prev];
                } else { pt = LX(w.lab, sv); rlast = P.ch_r[c]; }
                const int v = rlast - (LX(w.aq, qe) - pt);
                const long long tot = L.sum(ok && v > 0 ? (long long)(int)row_u[last] * v : 0);
                if (tot != 0 && L.lane == 0) atomicAdd(sums + 1 + p.L + last, (unsigned long long)tot);
            }
        }
        const long long rtot = L.sum(ok ? rhs : 0);
        if (rtot != 0 && L.lane == 0) atomicAdd(sums, (unsigned long long)rtot);
        if (active) {
            if (bad) atomicMin(p.first_inf + k, p.scen_offset + s);
            if (p.status) p.status[(size_t)k * p.S + s] = bad ? 1 : 0;
            if (p.obj) p.obj[(size_t)k * p.S + s] = bad ? 0.0 : (double)objv;
        }
    }
}
#undef LX

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

#ifndef SGUFP_K1_EMULATE
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

#endif  // SGUFP_K1_EMULATE

}  // namespace

#ifndef SGUFP_K1_EMULATE
template <int TILE, int NW>
static cudaError_t launch_tile_nw(const K1Launch &p, cudaStream_t st, int sm_count) {
    const int words = 2 * p.max_nch + 2 * p.max_nopen + 4 * (p.nc + 2) + p.nav + 2;
    constexpr int tiles = NW * 32 / TILE;
    const size_t smem = (size_t)tiles * words * sizeof(int);
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(k1_cut_eval<TILE, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k1_cut_eval<TILE, NW>, NW * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const long long items = (long long)p.K * p.S;
    long long want = (items + tiles - 1) / tiles;
    long long grid = (long long)sm_count * per_sm;   // persistent: a whole number of CTAs per SM
    if (want < grid) grid = want;
    if (grid < 1) grid = 1;
    k1_cut_eval<TILE, NW><<<(unsigned)grid, NW * 32, smem, st>>>(p, words);
    return cudaGetLastError();
}

// 8 warps per CTA; large networks (state of one scenario in the tens of KB) fall back to 2 or 1
template <int TILE>
static cudaError_t launch_tile(const K1Launch &p, cudaStream_t st, int sm_count) {
    cudaError_t e = launch_tile_nw<TILE, WARPS>(p, st, sm_count);
    if (e == cudaErrorInvalidConfiguration) { cudaGetLastError(); e = launch_tile_nw<TILE, 2>(p, st, sm_count); }
    if (e == cudaErrorInvalidConfiguration) { cudaGetLastError(); e = launch_tile_nw<TILE, 1>(p, st, sm_count); }
    return e;
}

int k1_tile_for(int max_nopen) {
    if (const char *e = getenv("SGUFP_K1_TILE")) { const int t = atoi(e); if (t == 8 || t == 16 || t == 32) return t; }
    (void)max_nopen;                                  // measured (profiles/): sub-warp tiles lose to divergence between tiles
    return 32;
}

static size_t lane_bytes_per_warp(const K1Launch &p) {
    return ((size_t)2 * (p.nc + 2) * 4 + (size_t)(p.nav + 1) * 4 + (size_t)3 * p.max_nopen * 2 + (size_t)(p.nc + 2) * 2 + (size_t)p.max_nopen) * 32;
}

// lane-per-scenario variant: one warp per CTA, as many CTAs per SM as the shared memory allows
static cudaError_t launch_lane(const K1Launch &p, cudaStream_t st, int sm_count) {
    const size_t smem = lane_bytes_per_warp(p);
    cudaError_t e = cudaFuncSetAttribute(k1_lane_eval<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k1_lane_eval<32>, 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const long long nblk = (long long)p.K * ((p.S + 31) / 32);
    long long grid = (long long)sm_count * per_sm;
    if (nblk < grid) grid = nblk;
    if (grid < 1) grid = 1;
    k1_lane_eval<32><<<(unsigned)grid, 32, smem, st>>>(p, (int)smem);
    return cudaGetLastError();
}

// which variant: the warp-per-scenario kernel unless SGUFP_K1_MODE=lane asks for the lane-per-scenario
// one (measured slower on C2 and C4, profiles/r01_k1_variants.md; kept as an independently written
// second implementation that the parity tests also run)
bool k1_use_lane_variant(const K1Launch &p) {
    const bool fits = p.max_cap < 65536 && lane_bytes_per_warp(p) <= 200 * 1024;
    if (const char *e = getenv("SGUFP_K1_MODE")) {
        if (e[0] == 'l') return fits;
        if (e[0] == 'w') return false;
    }
    return false;
}

cudaError_t k1_launch(const K1Launch &p, cudaStream_t st, int sm_count, int *launches) {
    if (launches) (*launches)++;
    if (k1_use_lane_variant(p)) return launch_lane(p, st, sm_count);
    int tile = k1_tile_for(p.max_nopen);
    // fall back to a wider tile (fewer tiles per CTA) if the per-CTA shared memory does not fit
    for (;;) {
        cudaError_t e = tile == 8 ? launch_tile<8>(p, st, sm_count) : tile == 16 ? launch_tile<16>(p, st, sm_count) : launch_tile<32>(p, st, sm_count);
        if (e == cudaErrorInvalidConfiguration && tile < 32) { tile *= 2; cudaGetLastError(); continue; }
        return e;
    }
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

#endif  // SGUFP_K1_EMULATE

}  // namespace sgufp

#if defined(SGUFP_K1_STATS) && !defined(SGUFP_K1_EMULATE)
// debug build only (-DSGUFP_K1_STATS): relaxation passes / label computations since the last call
extern "C" int sgufp_debug_k1_stats(unsigned long long *out4) {
    unsigned long long z[4] = {0, 0, 0, 0};
    cudaError_t e = cudaMemcpyFromSymbol(out4, sgufp::g_k1_stats, sizeof(z));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(sgufp::g_k1_stats, z, sizeof(z));
    return e == cudaSuccess ? 0 : -6;
}
#endif

