// k1_cut.cu — K1: batched per-scenario Benders subproblem evaluation + cut folding (sm_100a).
//
// Replaces the scenario loop of GuroSolver::solveSubProblem (/root/reference/grb.cpp:162-360):
// for every (candidate k, scenario s) it solves the second-stage LP exactly and adds the
// scenario's dual contribution cap x dual (grb.cpp:238-281) into exact integer accumulators.
//
// Mapping: one warp per (candidate, scenario); a persistent grid of
// sm_count x resident-CTAs strides over the K*S work items scenario-minor, so neighbouring warps
// stream neighbouring rows of the scenario-major fp64 capacity arrays with coalesced 128-bit
// loads.  The candidate's contracted graph (chains) is read-only and stays in L1;
// all per-scenario state lives in the warp's slice of shared memory.
// No tensor cores: this is integer graph work.
//
// Per work item:
//   1. stream u_s, l_s; segmented min/max into chain capacities (shared-memory atomics, packed
//      with the position of the FIRST least-capacity / LAST greatest-lower-bound arc)
//   2. optimal flow, primal-dual: a push-style label correction gives the shortest residual
//      distances (one per DISTINCT path length); the chains that are tight under those labels are
//      compacted once, and breadth-first searches over that short list (the reached set is one
//      register, one warp reduction per level) find and saturate every shortest path of that
//      length before the labels are computed again
//   3. SPEC-LP potentials (algorithm-independent) by label correction from the root
//   4. lift to (gamma, beta, sigma, phi, lambda, mu), multiply by the capacities and add to the
//      candidate's accumulators; write objective + status
#include "k1_cut.cuh"

#include <atomic>
#include <climits>
#include <type_traits>
#include <cstdlib>

#include "model.hpp"
#include "k1_common.cuh"

namespace sgufp {

#if defined(SGUFP_K1_STATS) && defined(SGUFP_K1_SMALL_TU)
#undef SGUFP_K1_STATS                       // the counters live in the main translation unit (no relocatable device code)
#endif
#ifdef SGUFP_K1_STATS
// SM clocks of the first lane per phase.  Flow kernel: 0 link (flow carried over), 1 row + start, 2 warm_init / the one label
// computation + the first tight list, 3 searches, 4 pushes, 5 dual updates (+ tight lists), 6 flow and state written, 7 evaluations;
// cut kernel: 8 row + flow read, 9 SPEC-LP potentials, 10 chains + alphas, 11 per-arc lifting, 12 closed chains + sums, 13 evaluations;
// 14 / 15: warm / cold evaluations of the flow kernel
__device__ unsigned long long g_k1_clk[16];
#define K1_CLK(i) do { if (T.tl == 0) { const long long now_ = clock64(); atomicAdd(&g_k1_clk[i], (unsigned long long)(now_ - clk_)); clk_ = now_; } } while (0)
__device__ unsigned long long g_k1_stats[8];   // relaxation passes, label computations, searches, work items, list sweeps, list entries swept, searches that reached dst, chains per pass
#endif

#ifndef K1_CLK
#define K1_CLK(i) do {} while (0)
#endif
#ifndef SGUFP_K1_EMULATE
#define K1_CNT(i, n) do {} while (0)
#endif
// No loop of the kernel is unrolled: a warp goes through every phase once per evaluation and the warps of an SM are in
// different phases, so the kernel is bound by its instruction-cache footprint, not by loop overhead (unrolled: 105 KB of SASS
// and a no_instruction stall of 3.9 warps per issue on C4; profiles/r02_k1_warm.md).  -DSGUFP_K1_UNROLL leaves it to the compiler.
// K1_LOOP: the loops of the flow phase (label computation, tight chains, search, push, dual update); K1_LOOPB: the others
// (streaming, potentials, lifting).  -DSGUFP_K1_UNROLL_FLOW / -DSGUFP_K1_UNROLL_BODY leave that group to the compiler.
#if defined(SGUFP_K1_UNROLL) || defined(SGUFP_K1_EMULATE)
#define SGUFP_K1_UNROLL_FLOW
#define SGUFP_K1_UNROLL_BODY
#endif
#ifdef SGUFP_K1_UNROLL_FLOW
#define K1_LOOP
#else
#define K1_LOOP _Pragma("unroll 1")
#endif
#ifdef SGUFP_K1_UNROLL_BODY
#define K1_LOOPB
#else
#define K1_LOOPB _Pragma("unroll 1")
#endif
// K1_LOOP1: loops that run once per evaluation over a handful of words (state rows, the link of a warm start, targets):
// never unrolled, in either size class
#ifdef SGUFP_K1_EMULATE
#define K1_LOOP1
#else
#define K1_LOOP1 _Pragma("unroll 1")
#endif
#if defined(SGUFP_K1_EMULATE) || defined(SGUFP_K1_UNROLL2)
#define K1_LOOP2
#else
#define K1_LOOP2 _Pragma("unroll 1")
#endif

namespace {

// This tile's slice of shared memory, as word OFFSETS into the CTA's buffer: indexing the
// __shared__ symbol directly keeps every access a plain LDS/STS/ATOMS with a register + immediate
// address (pointers kept in a struct are generic and cost an address computation per use).
template <bool BIG_>
struct TileMemT {
    static constexpr bool BIG = BIG_;     // contracted graph of more than SMALL_NC nodes (compile time: each kernel carries one search)
    int tin, tout, hist, path, tab;       // small graphs: bit sets of the tight residual graph, search levels, path, pair table
    int tc;                               // larger graphs: the tight-chain list (same place)
    int rw;                               // reached set of the last search
    int tg;                               // warm start: the set of nodes a search may end at (deficits; index nc = the root)
    int up, lo, x, res, lab, pred, pot, exc, aq;
};
#ifdef SGUFP_K1_EMULATE
#define k1_smem sgufp_emul_smem           // tests/cpp/k1_emul.cpp: the kernel body compiled for the host
static long long sgufp_emul_warm[8];   // warm starts taken / given up / saturated chains / searches / dual updates / sources
static long long sgufp_emul_cnt[16];   // tools/proto/flow_census.py: searches (repair ok / failed, to the sink ok / failed), chunk visits, ... (K1_CNT below)
#define K1_CNT(i, n) (sgufp_emul_cnt[i] += (n))
#else
extern __shared__ int k1_smem[];
#endif
#define SI(off) (k1_smem[off])
#define SU(off) (reinterpret_cast<unsigned *>(k1_smem)[off])
#define SH(off) (reinterpret_cast<unsigned short *>(k1_smem)[off])
#define SB(off) (reinterpret_cast<unsigned char *>(k1_smem)[off])
// residual / tight flags of chain c: one byte each for the larger graphs (shared memory decides how many tiles an SM
// holds: 3 -> 4 CTAs of 8 warps on C4), one word each for the small ones (byte accesses cost C2 7 %)
#define TM_BIG(w) (std::remove_cv_t<std::remove_reference_t<decltype(w)>>::BIG)
#define RGET(c) (TM_BIG(w) ? (int)SB(4 * w.res + (c)) : SI(w.res + (c)))
#define RSET(c, v) do { if (TM_BIG(w)) SB(4 * w.res + (c)) = (unsigned char)(v); else SI(w.res + (c)) = (v); } while (0)

__host__ __device__ inline int reach_words(int nc) { return (nc + 1 + 31) >> 5; }   // bit sets over the label indices 0..nc

// int32 words of shared memory per tile (the launcher and the host emulation size the buffer with it)
constexpr int SMALL_NC = 31;              // label indices 0..nc fit one 32-bit set
__host__ __device__ inline int k1_search_words(int nc, int max_nopen) {   // bit sets + levels + path + pair table, or the tight-chain list
    return nc <= SMALL_NC ? 32 + 32 + 34 + 34 + 32 * 32 / 2 : (max_nopen + 1) / 2;   // list entries are 16 bits
}
__host__ __device__ inline int k1_words_per_tile(const K1Launch &p) {
    return 3 * p.max_nopen + (p.nc <= SMALL_NC ? p.max_nopen : (p.max_nopen + 3) / 4) + 3 * (p.nc + 2) + p.nav + 2 + 2 * reach_words(p.nc) + k1_search_words(p.nc, p.max_nopen);
}

template <int TILE>
struct Lanes {    // the TILE lanes that work on one scenario (32 on the GPU; 1 in the host emulation)
    unsigned mask, lt;
    int tl, base;
    __device__ Lanes() {
        const int lane = TILE == 1 ? 0 : (int)(threadIdx.x & 31);
        tl = lane & (TILE - 1);
        base = lane & ~(TILE - 1);
        mask = TILE == 32 ? 0xffffffffu : (((1u << TILE) - 1u) << base);
        lt = (1u << tl) - 1u;
    }
    __device__ __forceinline__ void sync() const { __syncwarp(mask); }
    __device__ __forceinline__ bool any(bool p) const { return __any_sync(mask, p) != 0; }
    __device__ __forceinline__ unsigned ballot(bool p) const { return (__ballot_sync(mask, p) & mask) >> base; }
    __device__ __forceinline__ unsigned reduce_or(unsigned v) const {
        if constexpr (TILE == 32) return __reduce_or_sync(mask, v);
        else { for (int o = TILE / 2; o; o >>= 1) v |= __shfl_xor_sync(mask, v, o, TILE); return v; }
    }
    __device__ __forceinline__ long long sum(long long v) const {
        K1_LOOP
        for (int o = TILE / 2; o; o >>= 1) v += __shfl_xor_sync(mask, v, o, TILE);
        return v;
    }
    __device__ __forceinline__ int min_i32(int v) const {
        if constexpr (TILE == 32) return __reduce_min_sync(mask, v);
        else { for (int o = TILE / 2; o; o >>= 1) { const int t = __shfl_xor_sync(mask, v, o, TILE); v = t < v ? t : v; } return v; }
    }
    __device__ __forceinline__ unsigned long long min_u64(unsigned long long v) const {
        K1_LOOP
        for (int o = TILE / 2; o; o >>= 1) { const unsigned long long t = __shfl_xor_sync(mask, v, o, TILE); v = t < v ? t : v; }
        return v;
    }
};

// Label-correcting shortest distances from `src` over the residual arcs of the contracted graph
// (push style: every lane relaxes its chains, shared-memory atomicMin on the label).  One 64-bit
// load brings the static half of a chain, one word of flags its residual state.
// MERGED: the root is one node (index 0) that is never relabelled; otherwise arcs entering the
// root end at index nc.
// WARM: the nodes of the set w.rw (those a failed search could still reach over tight residual
// arcs) keep their labels — they are still exact, see reach_* — and only the others start over.
// zero: every node starts at 0 (a source of its own): the labels are then feasible potentials of the residual graph that are
// finite at EVERY node, which is what routing imbalances needs (forced flow of lower bounds).
template <int TILE, bool MERGED, bool WARM, class TM>
__device__ void shortest_paths(int src, const PlanView &P, int nopen, int nc, const Lanes<TILE> &T, TM &w, int &fuel, bool zero = false) {
    K1_LOOP
    for (int v = T.tl; v <= nc; v += TILE)
        if (!WARM || !((SU(w.rw + (v >> 5)) >> (v & 31)) & 1)) SI(w.lab + v) = zero ? 0 : LAB_INF;
    T.sync();
    if (!WARM && T.tl == 0) SI(w.lab + src) = 0;
    T.sync();
    bool changed;
    do {
        changed = false;
        K1_LOOP
        for (int c = T.tl; c < nopen; c += TILE) {
            const int f = RGET(c);                // bit 0 forward residual, bit 1 backward residual
            const ChainEnds e(P.ch_st[c]);
            const int ls = SI(w.lab + e.sv), le = SI(w.lab + e.ev);
            if ((f & 1) && ls != LAB_INF && !(MERGED && e.hf == nc)) {
                const int cand = ls - e.r;
                if (cand < SI(w.lab + e.hf)) { atomicMin(&SI(w.lab + e.hf), cand); changed = true; }
            }
            if ((f & 2) && le != LAB_INF && !(MERGED && e.hb == nc)) {
                const int cand = le + e.r;
                if (cand < SI(w.lab + e.hb)) { atomicMin(&SI(w.lab + e.hb), cand); changed = true; }
            }
        }
        T.sync();
        changed = T.any(changed) && --fuel > 0;
#ifdef SGUFP_K1_STATS
        if (T.tl == 0) { atomicAdd(&g_k1_stats[0], 1ull); atomicAdd(&g_k1_stats[7], (unsigned long long)nopen); }
#endif
    } while (changed);
#ifdef SGUFP_K1_STATS
    if (T.tl == 0) atomicAdd(&g_k1_stats[1], 1ull);
#endif
}

// ---- shortest paths of one length: searches over the TIGHT residual arcs ------------------------
// An arc is tight when label(tail) + cost == label(head).  Tightness depends on the labels only, so
// it is computed once per label computation (bits 2 and 3 of a chain's flag word) and stays valid
// while flow is pushed along tight arcs: the labels remain feasible potentials, every tight
// residual path is a shortest path, and a node that a complete search still reaches keeps its
// exact label (its tight path is a shortest path; no path can be shorter than the potential).
//
// Contracted graphs of up to 31 nodes (SMALL, config C2): the tight residual graph is two rows of
// bit sets per node (tout[u] = heads, tin[v] = tails) plus a table tab[v][u] with one usable arc per
// pair, all kept up to date under pushes.  A breadth-first level is then one OR-reduction over the
// warp, the path is read back one level per hop, and the hops are updated in parallel.
// Larger graphs: a compacted list of the tight chains, scanned once per level.

template <int TILE, class TM>
__device__ void tight_small(const PlanView &P, int nopen, const Lanes<TILE> &T, TM &w) {
    K1_LOOP
    for (int i = T.tl; i < 64; i += TILE) SU(w.tin + i) = 0u;    // tin[32] and tout[32] are adjacent
    T.sync();
    K1_LOOP
    for (int c = T.tl; c < nopen; c += TILE) {
        const ChainEnds e(P.ch_st[c]);
        const int ls = SI(w.lab + e.sv), le = SI(w.lab + e.ev), f = RGET(c) & 3;
        int tf = 0;
        if (ls != LAB_INF && ls - e.r == SI(w.lab + e.hf)) tf |= 1;
        if (le != LAB_INF && le + e.r == SI(w.lab + e.hb)) tf |= 2;
        RSET(c, f | (tf << 2));
        if (f & tf & 1) { atomicOr(&SU(w.tin + e.hf), 1u << e.sv); atomicOr(&SU(w.tout + e.sv), 1u << e.hf); SH(2 * w.tab + e.hf * 32 + e.sv) = (unsigned short)(2 * c); }
        if (f & tf & 2) { atomicOr(&SU(w.tin + e.hb), 1u << e.ev); atomicOr(&SU(w.tout + e.ev), 1u << e.hb); SH(2 * w.tab + e.hb * 32 + e.ev) = (unsigned short)(2 * c + 1); }
    }
    T.sync();
}

// Breadth-first search src -> dst; hist[l] = the nodes first reached at level l.  Returns the level
// of dst, or 0 if it cannot be reached — then SU(w.rw + 0) holds the complete reached set.
// dst < 0: the search ends at any node of the target set w.tg and *hit tells which (warm start).
template <int TILE, class TM>
__device__ int reach_small(int src, int dst, int nc, const Lanes<TILE> &T, TM &w, int *hit) {
    const unsigned dm = dst >= 0 ? 1u << dst : SU(w.tg);
    unsigned R = 1u << src, F = R;
    int lev = 0;
    if (T.tl == 0) SU(w.hist + 0) = F;
    K1_LOOP
    for (;;) {
        unsigned nb = 0;
        K1_LOOP
        for (int u = T.tl; u <= nc; u += TILE)
            if ((F >> u) & 1) nb |= SU(w.tout + u);
        nb = T.reduce_or(nb) & ~R;
        if (!nb) { if (T.tl == 0) SU(w.rw + 0) = R; T.sync(); return 0; }
        R |= nb; F = nb; lev++;
        if (T.tl == 0) SU(w.hist + lev) = F;
        if (R & dm) { *hit = __ffs(R & dm) - 1; T.sync(); return lev; }
    }
}

// Push along a path dst -> src that drops one level per hop.  Every lane walks it (broadcast
// loads) for the bottleneck; then one lane per hop moves the flow and updates the bit sets.
template <int TILE, class TM>
__device__ int push_small(int dst, int limit, int lev, int nc, const Lanes<TILE> &T, TM &w) {
    int v = dst, d = limit, h = 0;
    K1_LOOP
    for (int l = lev; l > 0; l--, h++) {
        const unsigned m = SU(w.tin + v) & SU(w.hist + l - 1);
        if (!m) { d = 0; break; }              // cannot happen: v was reached from level l-1
        const int u = __ffs(m) - 1, s = SH(2 * w.tab + v * 32 + u), c = s >> 1, xc = SI(w.x + c);
        d = min(d, (s & 1) ? xc - (SI(w.lo + c) >> HB) : (SI(w.up + c) >> HB) - xc);
        if (T.tl == 0) SI(w.path + h) = s | (u << 16) | (v << 21);
        v = u;
    }
    T.sync();
    if (d <= 0) return 0;
    K1_LOOP
    for (int i = T.tl; i < h; i += TILE) {
        const int rec = SI(w.path + i), s = rec & 0xffff, u = (rec >> 16) & 31, vv = (rec >> 21) & 31, c = s >> 1, dir = s & 1;
        const int xc = SI(w.x + c) + (dir ? -d : d);
        SI(w.x + c) = xc;
        const int nf = (xc < (SI(w.up + c) >> HB) ? 1 : 0) | (xc > (SI(w.lo + c) >> HB) ? 2 : 0) | (RGET(c) & 12);
        RSET(c, nf);
        if (!((nf >> dir) & 1)) { atomicAnd(&SU(w.tin + vv), ~(1u << u)); atomicAnd(&SU(w.tout + u), ~(1u << vv)); }   // saturated
        if ((nf >> (2 + (dir ^ 1))) & 1) {     // the reverse arc has residual capacity now and is tight
            const int rt = vv == nc ? 0 : vv, rh = u == 0 ? nc : u;
            atomicOr(&SU(w.tin + rh), 1u << rt); atomicOr(&SU(w.tout + rt), 1u << rh);
            SH(2 * w.tab + rh * 32 + rt) = (unsigned short)(s ^ 1);
        }
    }
    T.sync();
    return d;
}

// list variant: the tight chains compacted into w.tc as 16-bit entries chain | flags << 14 (a tile whose chains
// fit shared memory has far fewer than 16 384 of them)
template <int TILE, class TM>
__device__ int tight_list(const PlanView &P, int nopen, const Lanes<TILE> &T, TM &w) {
    int n = 0;
    K1_LOOP
    for (int c0 = 0; c0 < nopen; c0 += TILE) {     // tile-uniform trip count
        const int c = c0 + T.tl;
        int tf = 0;
        if (c < nopen) {
            const ChainEnds e(P.ch_st[c]);
            const int ls = SI(w.lab + e.sv), le = SI(w.lab + e.ev);
            if (ls != LAB_INF && ls - e.r == SI(w.lab + e.hf)) tf |= 1;
            if (le != LAB_INF && le + e.r == SI(w.lab + e.hb)) tf |= 2;
        }
        const unsigned b = T.ballot(tf != 0);
        if (tf) SH(2 * w.tc + n + __popc(b & T.lt)) = (unsigned short)(c | (tf << 14));
        n += __popc(b);
    }
    T.sync();
    return n;
}

// list variant of the search: sweeps over the tight-chain list that mark heads IN PLACE, so one sweep carries the
// reached set along every run of list entries whose tails come before their heads (a level-synchronous scan needs one
// sweep per level: 59 % of K1's instructions on C4, profiles/r01c_summary.md).  The lane whose atomicOr sets a
// node's bit is the only one to write its predecessor (tail | (2*chain+dir) << 10), and that tail's bit was set
// before: the predecessors form a tree rooted at src.  Any tight path will do — the duals do not depend on which
// optimal flow is found.  If dst is not reached, w.rw holds the complete reached set.
// Every (sweep, chunk) visit has a number q; a node's predecessor word also carries the visit at which it was reached
// (7 bits, saturating: tail | (2*chain+dir) << 10 | min(q, 127) << 25).  A node reached at visit q hangs on nodes reached at visits
// <= q, so after a push the nodes reached BEFORE the first visit that used a now saturated arc keep their place in the set and
// the next search starts in the middle of a sweep, at that visit's chunk (c0, q), instead of from the root.
template <int TILE, class TM>
__device__ bool reach_list(int src, int dst, const PlanView &P, int ntc, int nword, const Lanes<TILE> &T, TM &w, bool keep, int c0, int &q, int *hit) {
    const int R = w.rw;
    // keep: the search goes on from a reached set that is still valid — after a dual update (the nodes keep their labels and
    // their predecessors, only more arcs are tight), or after a push (the part of the set reached before the saturated arcs)
    if (!keep) for (int i = T.tl; i < nword; i += TILE) SU(R + (i)) = i == (src >> 5) ? 1u << (src & 31) : 0u;
    T.sync();
    bool partial = c0 > 0;                                  // the first sweep starts at chunk c0: it can find dst, it cannot prove a fixpoint
    K1_LOOP
    for (;;) {
        bool grew = false;
#ifdef SGUFP_K1_SKIP_CONFIRM
        bool back = false;          // a backward arc fired, or a usable backward entry is still waiting for its tail
#endif
        K1_LOOP
        for (int i0 = partial ? c0 * TILE : 0; i0 < ntc; i0 += TILE, q++) {           // tile-uniform trip count
            const int i = i0 + T.tl;
#ifdef SGUFP_K1_EMULATE
            {   // census: chunk visits, and those made after the target was already in the set
                bool there = false;
                if (dst >= 0) there = (SU(R + (dst >> 5)) >> (dst & 31)) & 1;
                else for (int j = 0; j < nword; j++) there |= (SU(R + j) & SU(w.tg + j)) != 0;
                K1_CNT(4, 1); if (there) K1_CNT(6, 1);
            }
#endif
            const int stamp = min(q, 127) << 25;
            int c = 0, f = 0;
            if (i < ntc) { const int en = SH(2 * w.tc + i); c = en & 0x3fff; f = RGET(c) & (en >> 14); }
            const ChainEnds e(f ? P.ch_st[c] : make_int2(0, 0));
            // the 32 entries of a chunk stay in registers until none of them can grow the set any more: entries of one
            // chunk feed each other (the list is in topological order of the tails), and a re-test is two loads per direction
            bool again;
            do {
                bool g = false;
                if ((f & 1) && ((SU(R + (e.sv >> 5)) >> (e.sv & 31)) & 1)) {
                    const unsigned bit = 1u << (e.hf & 31);
                    if (!((SU(R + (e.hf >> 5)) & bit)) && !(atomicOr(&SU(R + (e.hf >> 5)), bit) & bit)) { SI(w.pred + e.hf) = e.sv | ((2 * c) << 10) | stamp; g = true; }
                    f &= ~1;
                }
                if ((f & 2) && ((SU(R + (e.ev >> 5)) >> (e.ev & 31)) & 1)) {
                    const unsigned bit = 1u << (e.hb & 31);
                    if (!((SU(R + (e.hb >> 5)) & bit)) && !(atomicOr(&SU(R + (e.hb >> 5)), bit) & bit)) {
                        SI(w.pred + e.hb) = e.ev | ((2 * c + 1) << 10) | stamp; g = true;
#ifdef SGUFP_K1_SKIP_CONFIRM
                        back = true;
#endif
                    }
                    f &= ~2;
                }
                T.sync();
                again = T.any(g);
                grew |= again;
                K1_CNT(5, 1);
            } while (again);
#ifdef SGUFP_K1_SKIP_CONFIRM
            back |= (f & 2) != 0;
#endif
        }
        T.sync();
#ifdef SGUFP_K1_STATS
        if (T.tl == 0) { atomicAdd(&g_k1_stats[4], 1ull); atomicAdd(&g_k1_stats[5], (unsigned long long)(ntc - (partial ? c0 * TILE : 0))); }
#endif
        K1_CNT(7, 1);
        if (dst >= 0) {
            if ((SU(R + (dst >> 5)) >> (dst & 31)) & 1) {
#ifdef SGUFP_K1_STATS
                if (T.tl == 0) atomicAdd(&g_k1_stats[6], 1ull);
#endif
                *hit = dst;
                return true;
            }
        } else {                                             // any node of the target set (warm start): the lowest one reached
            int found = INT_MAX;
            K1_LOOP
            for (int i = T.tl; i < nword; i += TILE) { const unsigned both = SU(R + i) & SU(w.tg + i); if (both && found == INT_MAX) found = 32 * i + __ffs(both) - 1; }
            found = T.min_i32(found);
            if (found != INT_MAX) { *hit = found; return true; }
        }
        if (partial) { partial = false; continue; }          // full sweeps decide
        if (!T.any(grew)) return false;
#ifdef SGUFP_K1_SKIP_CONFIRM
        // The list is sorted by the depth of the tails (model.cpp: build_plan), so a forward entry can only be overtaken by a
        // backward arc that fires after it, and a waiting backward entry by any later arc: without either, this sweep already
        // reached the fixpoint.
        if (!T.any(back)) return false;
#endif
    }
}

// list variant of the push: every lane walks the predecessor tree dst -> src, then a hop per lane moves the flow.
// *restart: the visit number from which the reached set has to be searched again (the least visit stamp among the heads of the
// arcs this push saturates; the nodes reached at or after it are taken out of the set here), or -1: search from the root.
template <int TILE, class TM>
__device__ int push_list(int src, int dst, int limit, int nc, const Lanes<TILE> &T, TM &w, int *restart) {
    int v = dst, d = limit, hops = 0;
    *restart = -1;
#ifdef SGUFP_K1_PUSH_PAR
    int mine = 0;                          // the hop this lane will update (hop number == lane number): its predecessor word
    bool have = false;                     // (a word with a visit stamp of 64 or more is negative: the flag cannot live in its sign)
#endif
    K1_LOOP
    while (v != src) {
        const int p = SI(w.pred + v), s = (p >> 10) & 0x7fff, c = s >> 1, xc = SI(w.x + c);
        d = min(d, (s & 1) ? xc - (SI(w.lo + c) >> HB) : (SI(w.up + c) >> HB) - xc);
#ifdef SGUFP_K1_PUSH_PAR
        if (hops == T.tl) { mine = p; have = true; }
#endif
        v = p & 1023;
        if (++hops > nc + 1) { d = 0; break; }
    }
    K1_CNT(8, hops); K1_CNT(9, 1);
    T.sync();                              // every lane has its bottleneck before lane 0 moves the flow
    if (d <= 0) return 0;
#ifdef SGUFP_K1_PUSH_PAR
    // a path of at most TILE hops is updated in one step, a hop per lane (the chains of a simple path are distinct)
    if (hops <= TILE) {
        int sat = 255;                     // visit stamp of this hop's head if the push saturates its arc
        if (have) {
            const int s = (mine >> 10) & 0x7fff, c = s >> 1;
            const int xc = SI(w.x + c) + ((s & 1) ? -d : d);
            SI(w.x + c) = xc;
            const int fl = (xc < (SI(w.up + c) >> HB) ? 1 : 0) | (xc > (SI(w.lo + c) >> HB) ? 2 : 0);
            RSET(c, fl);
            if (!((fl >> (s & 1)) & 1)) sat = (int)((unsigned)mine >> 25);
        }
#ifndef SGUFP_K1_NO_RESTART
        const int qs = T.min_i32(sat);     // at least one arc saturates (d is the bottleneck) unless `limit` bound the push
        if (qs > 0 && qs < 127) {
            if constexpr (TILE == 32) {   // a word of the set per step: the lanes vote on its 32 nodes
                K1_LOOP
                for (int u = T.tl; u < 32 * reach_words(nc); u += 32) {
                    const bool in = u <= nc && ((SU(w.rw + (u >> 5)) >> (u & 31)) & 1);
                    const unsigned word = T.ballot(in && (u == src || (int)((unsigned)SI(w.pred + u) >> 25) < qs));
                    if (T.tl == 0) SU(w.rw + (u >> 5)) = word;
                }
            } else
            K1_LOOP
            for (int u = T.tl; u <= nc; u += TILE)
                if (u != src && ((SU(w.rw + (u >> 5)) >> (u & 31)) & 1) && (int)((unsigned)SI(w.pred + u) >> 25) >= qs) atomicAnd(&SU(w.rw + (u >> 5)), ~(1u << (u & 31)));
            *restart = qs;
        }
#endif
        T.sync();
        return d;
    }
#endif
    v = dst;
    K1_LOOP
    while (v != src) {
        const int p = SI(w.pred + v), s = (p >> 10) & 0x7fff, c = s >> 1;
        if (T.tl == 0) {
            const int xc = SI(w.x + c) + ((s & 1) ? -d : d);
            SI(w.x + c) = xc;
            RSET(c, (xc < (SI(w.up + c) >> HB) ? 1 : 0) | (xc > (SI(w.lo + c) >> HB) ? 2 : 0));
        }
        v = p & 1023;
    }
    T.sync();
    return d;
}

// What one label computation allows: search + push over its tight arcs, until none is left.
template <int TILE, class TM>
struct TightPaths {
    const PlanView &P;
    const Lanes<TILE> &T;
    TM &w;
    int nopen, nc, nword, ntc;
    int rq, q;                            // list search: visit to restart from after a push (-1: from the root), visit counter
    int hit;                              // the node the last successful search ended at
#ifdef SGUFP_K1_STATS
    long long clk_ = 0;                   // K1_CLK inside the flow loop (handed over by the kernel around solve)
#endif
    static constexpr bool small = !TM::BIG;
    __device__ TightPaths(const PlanView &P_, const Lanes<TILE> &T_, TM &w_, int nopen_, int nc_)
        : P(P_), T(T_), w(w_), nopen(nopen_), nc(nc_), nword(reach_words(nc_)), ntc(0), rq(-1), q(0), hit(0) {}
    __device__ void prepare() {
        if constexpr (small) tight_small<TILE>(P, nopen, T, w); else ntc = tight_list<TILE>(P, nopen, T, w);
        K1_CNT(11, 1); K1_CNT(12, ntc);
        rq = -1;                          // a new list: chunk positions of the old one mean nothing
    }
    __device__ __forceinline__ bool in_r(int v) const {
        if constexpr (small) return (SU(w.rw) >> v) & 1u; else return (SU(w.rw + (v >> 5)) >> (v & 31)) & 1u;
    }
    // Primal-dual step after a FAILED search from the root (w.rw = the complete reached set R): the labels of the nodes outside R
    // rise by delta = the least slack of a residual arc that leaves R.  They stay feasible potentials (every residual arc keeps
    // a non-negative reduced cost), the labels inside R stay exact, at least one more arc becomes tight: ONE pass over the
    // chains instead of a label computation of 3.6 - 3.9 relaxation passes (profiles/r02_summary.md).  Returns false when the
    // flow is optimal: no residual arc leaves R (maximum flow), or the sink's label has reached 0 (no profitable path left).
    // to_sink = false (repair of a warm start, any source): only "no residual arc leaves R" ends it; root_in: the search
    // started at the root, whose other face (index nc, the root as a path end) then belongs to R as well.
    __device__ bool dual_update(bool to_sink = true, bool root_in = false) {
        if (root_in) {
            if (T.tl == 0) { if constexpr (small) SU(w.rw) |= 1u << nc; else SU(w.rw + (nc >> 5)) |= 1u << (nc & 31); }
            T.sync();
        }
        int best = INT_MAX;
        K1_CNT(10, 1);
        K1_LOOP
        for (int c = T.tl; c < nopen; c += TILE) {
            const int f = RGET(c) & 3;
            if (!f) continue;
            const ChainEnds e(P.ch_st[c]);
            if ((f & 1) && in_r(e.sv) && !in_r(e.hf)) { const int lh = SI(w.lab + e.hf); if (lh != LAB_INF) best = min(best, SI(w.lab + e.sv) - e.r - lh); }
            if ((f & 2) && in_r(e.ev) && !in_r(e.hb)) { const int lh = SI(w.lab + e.hb); if (lh != LAB_INF) best = min(best, SI(w.lab + e.ev) + e.r - lh); }
        }
        int delta = T.min_i32(best);
#ifdef SGUFP_K1_STATS
        if (T.tl == 0) atomicAdd(&g_k1_stats[1], 1ull);
#endif
        // to the sink: never past the root's label (a smaller step keeps the potentials feasible too), and when no residual arc
        // leaves R — any step does — exactly up to it: a finished flow leaves lab[nc] == lab[0], the root's ONE potential,
        // which is what the next candidate of the run starts from
        if (to_sink) delta = min(delta, SI(w.lab + 0) - SI(w.lab + nc));
        if (delta == INT_MAX) return false;
        T.sync();
        K1_LOOP
        for (int v = T.tl; v <= nc; v += TILE)
            if (!in_r(v)) { const int l = SI(w.lab + v); if (l != LAB_INF) SI(w.lab + v) = l + delta; }
        T.sync();
        if (to_sink && SI(w.lab + nc) >= SI(w.lab + 0)) return false;
        prepare();
        return true;
    }
    // The whole flow phase in ONE loop, so that the kernel holds one copy of the search, the push and the dual update (with a
    // copy per use it is instruction-cache bound: profiles/r02_k1_warm.md).  Stages: (0) imbalances, if the caller booked any
    // in w.exc with the deficits' bits in w.tg — every excess to the nearest deficit or to the root (index nc), then the
    // root's own excess (index 0) to the deficits that are left: successive shortest paths for a pseudoflow, the labels stay
    // feasible potentials throughout; (1) from the root to the sink along tight residual paths until the sink's label reaches
    // the root's or no residual arc leaves the reached set.  A failed search is followed by a dual update.
    // Returns 0: optimal flow; 1: an imbalance cannot be routed (forced flow: the scenario is infeasible; warm start: cannot
    // happen on a valid state, the caller starts over from zero flow); 2: out of fuel.
    __device__ int solve(bool imbalances, int &fuel) {
        int stage = imbalances ? 0 : 2;       // 0: nodes with an excess, 1: the root's excess, 2: to the sink (not begun), 3: to the sink
        int v0 = 0, src = 0, need = 0;
        unsigned todo = 0;
        bool keep = false;
        K1_LOOP
        for (;;) {
            if (need <= 0) {                  // the next source
                if (stage == 0) {
                    K1_LOOP
                    while (!todo && v0 < nc) {
                        const int v = v0 + T.tl;
                        todo = T.ballot(v > 0 && v < nc && SI(w.exc + v) > 0);
                        v0 += TILE;
                    }
                    if (todo) { src = v0 - TILE + __ffs(todo) - 1; todo &= todo - 1; need = SI(w.exc + src); }
                    else stage = 1;
                }
                if (stage == 1) {             // the root is one node: what reached its end face (nc) leaves from its start face (0)
                    if (T.tl == 0) SU(w.tg + (nc >> 5)) &= ~(1u << (nc & 31));
                    T.sync();
                    stage = 2; src = 0; need = SI(w.exc + 0);
                }
                if (stage == 2 && need <= 0) {
                    const int lt = SI(w.lab + nc);
                    if (lt == LAB_INF || lt >= SI(w.lab + 0)) return 0;
                    stage = 3; src = 0; need = INT_MAX;
                }
                rq = -1; keep = false;
#ifdef SGUFP_K1_EMULATE
                if (stage < 3) sgufp_emul_warm[5]++;
#endif
            }
            if (--fuel <= 0) return 2;
            const int d = augment(src, stage == 3 ? nc : -1, need, keep);
#ifdef SGUFP_K1_EMULATE
            if (stage < 3) { sgufp_emul_warm[3]++; if (d <= 0) sgufp_emul_warm[4]++; }
#endif
            keep = false;
            K1_CNT((stage < 3 ? 0 : 2) + (d > 0 ? 0 : 1), 1);
            if (d > 0) {
                if (stage < 3) {
                    need -= d;
                    T.sync();
                    if (T.tl == 0) {                   // books: the source, the node it ended at (its target bit goes with its deficit)
                        const int t = hit == nc ? 0 : hit, left = SI(w.exc + t) + d;
                        SI(w.exc + src) = need; SI(w.exc + t) = left;
                        if (!left && hit != nc) SU(w.tg + (hit >> 5)) &= ~(1u << (hit & 31));
                    }
                    T.sync();
                }
                continue;
            }
            const bool more = dual_update(stage == 3, stage < 3 && src == 0);
            K1_CLK(5);
            if (!more) { if (stage == 3) return 0; return 1; }
            keep = TM::BIG;
        }
    }
    // one search and, if dst is reachable, one push of at most `limit`; returns the amount pushed (0: not reachable)
    // dst < 0 (warm start): to the first node of the target set w.tg the search reaches — at most its deficit; the root (index nc)
    // takes any amount; hit = that node
    __device__ int augment(int src, int dst, int limit, bool keep = false) {
#ifdef SGUFP_K1_STATS
        if (T.tl == 0) atomicAdd(&g_k1_stats[2], 1ull);
#endif
        if constexpr (small) {
            const int lev = reach_small<TILE>(src, dst, nc, T, w, &hit);
            K1_CLK(3);
            if (!lev) return 0;
            if (dst < 0 && hit != nc) limit = min(limit, -SI(w.exc + hit));
            const int d = push_small<TILE>(hit, limit, lev, nc, T, w);
            K1_CLK(4);
            return d;
        } else {
            const int nchunk = (ntc + TILE - 1) / TILE;
            int c0 = 0;
            if (rq > 0 && nchunk > 0) { c0 = rq % nchunk; q = rq; keep = true; }      // after a push: from the middle of a sweep
            else if (keep && nchunk > 0) q = (q + nchunk - 1) / nchunk * nchunk;     // after a dual update: a new sweep, visit numbers go on
            else { q = 0; keep = false; }
            rq = -1;
            const bool found = reach_list<TILE>(src, dst, P, ntc, nword, T, w, keep, c0, q, &hit);
            K1_CLK(3);
            if (!found) return 0;
            if (dst < 0 && hit != nc) limit = min(limit, -SI(w.exc + hit));
            const int d = push_list<TILE>(src, hit, limit, nc, T, w, &rq);
            K1_CLK(4);
            return d;
        }
    }
};

// Target set of an imbalance routing (TightPaths::solve): the nodes with a deficit, and the root (as a path end: index nc)
// whatever its own balance — it is ONE node, so what an excess sends there goes on to the remaining deficits from its other
// face (index 0) at the end.  exc[0] is the root's own imbalance (the books include it: the imbalances sum to zero).
template <int TILE, class TM>
__device__ void deficit_targets(int nc, const Lanes<TILE> &T, TM &w) {
    K1_LOOP1
    for (int i = T.tl; i < reach_words(nc); i += TILE) SU(w.tg + i) = 0u;
    T.sync();
    K1_LOOP1
    for (int v = T.tl; v <= nc; v += TILE)
        if (v == nc || (v > 0 && SI(w.exc + v) < 0)) atomicOr(&SU(w.tg + (v >> 5)), 1u << (v & 31));
    T.sync();
}

// Forced flow from lower bounds (rare): x = lo leaves an excess at the head and a deficit at the tail of every chain with a
// positive lower bound; the labels are feasible potentials that are finite everywhere.  TightPaths::solve routes the
// imbalances; one that cannot be routed means the scenario is infeasible.
template <int TILE, class TM>
__device__ void forced_flow_init(const PlanView &P, int nopen, int nc, const Lanes<TILE> &T, TM &w, int &fuel) {
    K1_LOOP1
    for (int v = T.tl; v <= nc; v += TILE) SI(w.exc + v) = 0;
    T.sync();
    K1_LOOP1
    for (int c = T.tl; c < nopen; c += TILE) {
        const int lo = SI(w.lo + c) >> HB;
        if (lo > 0) {
            const ChainEnds e(P.ch_st[c]);
            atomicAdd(&SI(w.exc + e.ev), lo);
            atomicSub(&SI(w.exc + e.sv), lo);
        }
    }
    T.sync();
    deficit_targets<TILE>(nc, T, w);
}

// SPEC-LP potentials: pot[v] = -(shortest residual distance from the root);
// nodes the root cannot reach get the least labels consistent with the labelled ones; nodes cut
// off both ways get a zero-rooted completion.
template <int TILE, class TM>
__device__ void canonical_potentials(const PlanView &P, int nopen, int nc, const Lanes<TILE> &T, TM &w, int &fuel) {
    shortest_paths<TILE, true, false>(0, P, nopen, nc, T, w, fuel);
    bool missing = false;
    K1_LOOP2
    for (int v = T.tl; v < nc; v += TILE) {
        const int l = SI(w.lab + v);
        if (l == LAB_INF) { missing = true; SI(w.pot + v) = NEG_INF; SI(w.pred + v) = 0; } else { SI(w.pot + v) = l; SI(w.pred + v) = 1; }
    }
    T.sync();
    if (T.any(missing)) {
        // pot[] holds d here; cur[] is the state: 1 labelled by phase 1, 0 not yet, 2 isolated
        bool changed;
        do {
            changed = false;
            K1_LOOP2
            for (int c = T.tl; c < nopen; c += TILE) {
                const int e = P.ch_ends[c], sv = (e & 0xffff) - 1, ev = (e >> 16) - 1, r = P.ch_r[c];
                const int xc = SI(w.x + c), up = SI(w.up + c) >> HB, lo = SI(w.lo + c) >> HB;
                if (xc < up && SI(w.pred + sv) == 0) { const int db = SI(w.pot + ev); if (db > NEG_INF && db + r > atomicMax(&SI(w.pot + sv), db + r)) changed = true; }
                if (xc > lo && SI(w.pred + ev) == 0) { const int da = SI(w.pot + sv); if (da > NEG_INF && da - r > atomicMax(&SI(w.pot + ev), da - r)) changed = true; }
            }
            T.sync();
            changed = T.any(changed) && --fuel > 0;
        } while (changed);
        bool iso = false;
        K1_LOOP2
        for (int v = T.tl; v < nc; v += TILE)
            if (SI(w.pot + v) == NEG_INF) { SI(w.pot + v) = 0; SI(w.pred + v) = 2; iso = true; }
        T.sync();
        if (T.any(iso)) {
            do {
                changed = false;
                K1_LOOP2
                for (int c = T.tl; c < nopen; c += TILE) {
                    const int e = P.ch_ends[c], sv = (e & 0xffff) - 1, ev = (e >> 16) - 1, r = P.ch_r[c];
                    const int xc = SI(w.x + c), up = SI(w.up + c) >> HB, lo = SI(w.lo + c) >> HB;
                    if (xc < up && SI(w.pred + ev) == 2) { const int cand = SI(w.pot + sv) - r; if (cand < atomicMin(&SI(w.pot + ev), cand)) changed = true; }
                    if (xc > lo && SI(w.pred + sv) == 2) { const int cand = SI(w.pot + ev) + r; if (cand < atomicMin(&SI(w.pot + sv), cand)) changed = true; }
                }
                T.sync();
                changed = T.any(changed) && --fuel > 0;
            } while (changed);
        }
    }
    K1_LOOP2
    for (int v = T.tl; v < nc; v += TILE) SI(w.pot + v) = -SI(w.pot + v);
    T.sync();
}

// Warm start (one work item = a run of consecutive candidates on ONE scenario): x holds the previous candidate's optimal
// flow on the chains this candidate still has (0 on the new ones), lab its potentials (all finite), exc the imbalances the
// removed chains left behind.  A chain whose reduced cost has the wrong sign for its flow is saturated / emptied (more
// imbalance); TightPaths::solve then routes every excess to a deficit or the root and the root's excess to the deficits left,
// along tight residual paths with dual updates: the labels stay feasible potentials throughout, so the circulation that is
// left when the imbalances are gone is optimal.  Consecutive paths of the Benders loop differ in a few layers: ~16 pushes
// instead of ~100 on the C4 network (profiles/r02_k1_warm.md).  Instances without forced flow only (lo == 0).
template <int TILE, class TM>
__device__ void warm_init(const PlanView &P, int nopen, int nc, const Lanes<TILE> &T, TM &w) {
    K1_LOOP2
    for (int c = T.tl; c < nopen; c += TILE) {
        const ChainEnds e(P.ch_st[c]);
        const int up = SI(w.up + c) >> HB;
        int xc = min(SI(w.x + c), up);
        const int rc = SI(w.lab + e.sv) - e.r - SI(w.lab + e.ev);      // reduced cost of the forward arc (root = index 0 on both sides here)
        int d = 0;
        if (rc < 0 && xc < up) { d = up - xc; xc = up; }
        else if (rc > 0 && xc > 0) { d = -xc; xc = 0; }
        if (d) { atomicSub(&SI(w.exc + e.sv), d); atomicAdd(&SI(w.exc + e.ev), d); }
#ifdef SGUFP_K1_EMULATE
        if (d) sgufp_emul_warm[2]++;
#endif
        SI(w.x + c) = xc;
        RSET(c, (xc < up ? 1 : 0) | (xc > 0 ? 2 : 0));
    }
    T.sync();
    deficit_targets<TILE>(nc, T, w);
}

// wire potential at the HEAD of arc a (a matched in-arc, or any arc of a chain)
template <class TM>
__device__ __forceinline__ int head_potential(int a, const PlanView &P, int nopen, const TM &w) {
    const int cp = P.arc_cp[a], c = cp >> 10, pos = cp & 1023, pre = P.arc_pre[a];
    const int e = P.ch_ends[c], sv = (e & 0xffff) - 1, ev = (e >> 16) - 1;
    if (c < nopen) {
        const int r = P.ch_r[c], dp = SI(w.pot + ev) - SI(w.pot + sv);
        const int g = max(0, r - dp), b = max(0, dp - r);
        return SI(w.pot + sv) + pre - (pos >= (SI(w.up + c) & 1023) ? g : 0) + (pos >= (SI(w.lo + c) & 1023) ? b : 0);
    }
    if (sv >= 0) return SI(w.pot + sv) + pre;
    if (ev >= 0) return SI(w.pot + ev) - (P.ch_r[c] - pre);
    return pre;
}

// this tile's slice of the CTA's shared memory (both kernels lay it out alike; k1_words_per_tile sizes it)
template <class TM>
__device__ __forceinline__ void tile_mem(TM &w, int tile_in_cta, int words_per_tile, const K1Launch &p) {
    int base = tile_in_cta * words_per_tile;   // the fixed-size arrays first: constant offsets from the tile base
    w.tin = base; w.tout = base + 32; w.hist = base + 64; w.path = base + 98; w.tab = base + 132;
    w.tc = base; base += k1_search_words(p.nc, p.max_nopen);
    w.rw = base; base += reach_words(p.nc);
    w.tg = base; base += reach_words(p.nc);
    w.up = base; base += p.max_nopen;
    w.lo = base; base += p.max_nopen;
    w.x = base; base += p.max_nopen;
    w.res = base; base += TM::BIG ? (p.max_nopen + 3) / 4 : p.max_nopen;
    w.lab = base; base += p.nc + 2;
    w.pred = base; base += p.nc + 2;
    w.pot = base; w.exc = base; base += p.nc + 2;   // imbalances (flow kernel) and potentials (cut kernel) share a place
    w.aq = base;
}

// 1. chain capacities of one scenario (open chains only: a closed chain carries no flow, it is infeasible iff one of its arcs
//    has a positive lower bound, and its multipliers are read from the capacity row directly): the row is streamed with 128-bit
//    loads, segmented min / max into up[c] = (least capacity << 10 | position of its FIRST arc) and lo[c] = (greatest lower bound
//    << 10 | position of its LAST arc) — with no positive lower bound on the chain that is (0, last position), known up front, so
//    only positive lower bounds (rare) need the atomic.  Returns this lane's share of "a closed chain has a positive lower bound".
template <int TILE, class TM>
__device__ __forceinline__ bool stream_row(const PlanView &P, int nopen, int m, int m_pad, const double *row_u, const double *row_l, const Lanes<TILE> &T, TM &w) {
    K1_LOOP2
    for (int c = T.tl; c < nopen; c += TILE) { SI(w.up + c) = INT_MAX; SI(w.lo + c) = P.ch_ptr[c + 1] - P.ch_ptr[c] - 1; }
    T.sync();
    bool bad = false;
    const double2 *ru = reinterpret_cast<const double2 *>(row_u), *rl = reinterpret_cast<const double2 *>(row_l);
    K1_LOOPB
    for (int a2 = T.tl; a2 < m_pad / 2; a2 += TILE) {
        const double2 u2 = __ldg(ru + a2), l2 = __ldg(rl + a2);
        const int a = 2 * a2;
        {
            const int cp = P.arc_cp[a], c = cp >> 10, pos = cp & 1023;
            if (c < nopen) { atomicMin(&SI(w.up + c), ((int)u2.x << HB) | pos); if ((int)l2.x > 0) atomicMax(&SI(w.lo + c), ((int)l2.x << HB) | pos); }
            else bad |= (int)l2.x > 0;
        }
        if (a + 1 < m) {
            const int cp = P.arc_cp[a + 1], c = cp >> 10, pos = cp & 1023;
            if (c < nopen) { atomicMin(&SI(w.up + c), ((int)u2.y << HB) | pos); if ((int)l2.y > 0) atomicMax(&SI(w.lo + c), ((int)l2.y << HB) | pos); }
            else bad |= (int)l2.y > 0;
        }
    }
    T.sync();
    return bad;
}

// K1 is two kernels per launch (profiles/r02_k1_warm.md):
//   k1_cut_eval (the FLOW kernel): a warp takes a run of consecutive candidates on one scenario, finds an optimal flow of each
//     (the first from zero flow, the others from the one before) and writes it to xout[k][s][chain];
//   k1_cut_fold (the CUT kernel): a warp per (candidate, scenario), candidate-major, so that the warps of an SM read ONE plan and
//     run the same few hundred instructions: SPEC-LP potentials of the flow's residual graph, lifting, and the fold into the
//     candidate's accumulators.
// BIG: the contracted graph has more than SMALL_NC nodes (list searches, flag bytes); the launcher picks the instantiation.
template <int TILE, int NW, bool BIG>
__global__ void __launch_bounds__(NW * 32, NW == 8 ? SGUFP_K1_MINBLOCKS : NW == 4 ? 2 * SGUFP_K1_MINBLOCKS : 1) k1_cut_eval(K1Launch p, int words_per_tile) {
    const Lanes<TILE> T;
    constexpr int TILES_PER_CTA = NW * 32 / TILE;
    const int tile_in_cta = threadIdx.x / TILE;
    using TM = TileMemT<BIG>;
    TM w;
    tile_mem(w, tile_in_cta, words_per_tile, p);
    // A work item is a RUN of `group` consecutive candidates on one scenario: the first is solved from zero flow, each of the
    // others from the optimal flow and potentials of the one before it (warm_repair), when the host linked the two plans.
    const int group = p.group > 1 && p.link_off ? p.group : 1, ngroups = (p.K + group - 1) / group;
    const long long items = (long long)ngroups * p.S;
    const long long stride = (long long)gridDim.x * TILES_PER_CTA;
    // Work items come from a queue (one atomic per item, taken by the tile's first lane) when the launch carries one: an item
    // takes 0.1 - 0.3 ms and varies with the scenario, so with a few items per warp (a rank of an 8-GPU partition of C4 has
    // 2.1) a fixed assignment leaves most of the GPU waiting for the warps that drew three long ones.  Scenario-minor order
    // either way: neighbouring tiles stream neighbouring rows.
    long long item = (long long)blockIdx.x * TILES_PER_CTA + tile_in_cta - stride;
    K1_LOOPB
    for (;;) {
        if (p.work) {
            unsigned long long nxt = 0;
            if (T.tl == 0) nxt = atomicAdd(p.work, 1ull);
            item = (long long)__shfl_sync(T.mask, nxt, T.base, 32);
        } else item += stride;
        if (item >= items) break;
        const int grp = (int)(item / p.S), s = (int)(item - (long long)grp * p.S);
        const double *row_u = p.cap_u + (size_t)s * p.m_pad, *row_l = p.cap_l + (size_t)s * p.m_pad;
        bool holds = false;                     // this tile holds the optimal flow (x) and the labels (lab: feasible potentials, finite everywhere) of the candidate before k
        int32_t *srow = p.state && ngroups == 1 ? p.state + (size_t)s * p.state_stride : nullptr;   // this scenario's row of the handle's state
        if (srow && (p.state_io & 1) && srow[0]) {         // left by the last candidate of the previous launch on this handle
            const int nprev = (p.plans + p.link_off[p.order ? p.order[0] : 0])[1];
            K1_LOOP1
            for (int v = T.tl; v <= p.nc; v += TILE) SI(w.lab + v) = srow[1 + v];
            K1_LOOP1
            for (int c = T.tl; c < nprev; c += TILE) SI(w.x + c) = srow[2 + p.nc + c];
            holds = true;
        }
        T.sync();
        if (srow && (p.state_io & 2) && T.tl == 0) srow[0] = 0;   // until the last candidate of this launch has filled it
        K1_LOOPB
        for (int j = grp * group, j_end = min(p.K, j + group); j < j_end; j++) {
        const int k = p.order ? p.order[j] : j;            // the run takes the candidates in the host's order (model.hpp: order_batch)
        const bool carried = holds;
        holds = false;
        if (*reinterpret_cast<volatile long long *>(p.first_inf + k) < 0) continue;   // this candidate was aborted: drain
        int fuel = 1 << 20;   // passes + levels a work item may spend (tile-uniform)
        const PlanView P(p.plans + p.plan_off[k]);
        const int nopen = P.h->nopen, nc = P.h->nc, m = p.m;
        const int32_t *link = nullptr;          // model.hpp: link_plans
        if (carried) { const int lo = p.link_off[k]; if (lo >= 0) link = p.plans + lo; }
#ifdef SGUFP_K1_STATS
        long long clk_ = clock64();
#endif
        if (link) {
            // the labels stay (the previous candidate's final potentials: finite at every node, the root's two faces equal);
            // imbalances of the chains that are gone, and the flow carried over to this candidate's chain numbering (through
            // w.up, which is rebuilt below)
            K1_LOOP1
            for (int v = T.tl; v <= nc; v += TILE) SI(w.exc + v) = 0;
            T.sync();
            const int nrem = link[0];
            const int32_t *prev_of = link + 2, *removed = prev_of + nopen, *ends = removed + nrem;
            K1_LOOP1
            for (int i = T.tl; i < nrem; i += TILE) {
                const int f = SI(w.x + removed[i]);
                if (f > 0) { const int e = ends[i]; atomicAdd(&SI(w.exc + (e & 1023)), f); atomicSub(&SI(w.exc + (e >> 10)), f); }
            }
            K1_LOOP1
            for (int c = T.tl; c < nopen; c += TILE) { const int pc = prev_of[c]; SI(w.up + c) = pc >= 0 ? SI(w.x + pc) : 0; }
            T.sync();
            K1_LOOP1
            for (int c = T.tl; c < nopen; c += TILE) SI(w.x + c) = SI(w.up + c);
            T.sync();
        }

        K1_CLK(0);
        bool bad = stream_row<TILE>(P, nopen, m, p.m_pad, row_u, row_l, T, w), forced = false;
        K1_LOOP2
        for (int c = T.tl; c < nopen; c += TILE) {
            const int lo = SI(w.lo + c) >> HB, up = SI(w.up + c) >> HB;
            bad |= lo > up; forced |= lo > 0;
            if (!link) { SI(w.x + c) = lo; RSET(c, lo < up ? 1 : 0); }   // x == lo: forward residual only
        }
        T.sync();
        bad = T.any(bad);
        forced = T.any(forced);
        // 2. optimal flow
        TightPaths<TILE, TM> TP(P, T, w, nopen, nc);
        K1_CLK(1);
        int *xrow = p.xout + ((size_t)k * p.S + s) * p.xstride;
        if (bad) {
            if (T.tl == 0) {
                atomicMin(p.first_inf + k, p.scen_offset + s);
                if (p.status) p.status[(size_t)k * p.S + s] = 1;
                if (p.obj) p.obj[(size_t)k * p.S + s] = 0.0;
                xrow[0] = INT_MIN;
            }
            T.sync();
            continue;
        }
        if (link && forced) {                        // a run goes on from zero flow where a scenario has forced flow
            K1_LOOPB
            for (int c = T.tl; c < nopen; c += TILE) { const int lo = SI(w.lo + c) >> HB; SI(w.x + c) = lo; RSET(c, lo < (SI(w.up + c) >> HB) ? 1 : 0); }
            T.sync();
        }
        // ONE call of the flow loop serves the three starts (a second round only when a warm start gets stuck, which no valid
        // state does): warm (the previous candidate's flow and potentials), forced (x = lo), from zero flow.
        bool warm = link && !forced;
        int rc;
        K1_LOOPB
        for (;;) {
            if (warm) warm_init<TILE>(P, nopen, nc, T, w);
            else { if (forced) forced_flow_init<TILE>(P, nopen, nc, T, w, fuel); shortest_paths<TILE, false, false>(0, P, nopen, nc, T, w, fuel, true);   // the ONE label computation: feasible potentials at the starting flow, finite at every node
            }
            TP.prepare();
            K1_CLK(2);
#ifdef SGUFP_K1_STATS
            TP.clk_ = clk_;
#endif
            K1_CNT(13, nopen); K1_CNT(14, 1); K1_CNT(15, nc);
            rc = TP.solve(warm || forced, fuel);
#ifdef SGUFP_K1_STATS
            clk_ = TP.clk_;
#endif
#ifdef SGUFP_K1_EMULATE
            if (warm) sgufp_emul_warm[rc ? 1 : 0]++;   // tests: warm starts taken / given up
#endif
#ifdef SGUFP_K1_STATS
            if (T.tl == 0) { atomicAdd(&g_k1_clk[warm && !rc ? 14 : 15], 1ull); atomicAdd(&g_k1_clk[7], 1ull); }
#endif
            if (!warm || rc != 1) break;
            K1_LOOPB
            for (int c = T.tl; c < nopen; c += TILE) { SI(w.x + c) = 0; RSET(c, (SI(w.up + c) >> HB) > 0 ? 1 : 0); }   // from zero flow (lo == 0 here)
            T.sync();
            warm = false; fuel = 1 << 20;
        }
        if (rc == 1) {                               // forced flow that cannot be routed: infeasible
            if (T.tl == 0) {
                atomicMin(p.first_inf + k, p.scen_offset + s);
                if (p.status) p.status[(size_t)k * p.S + s] = 1;
                if (p.obj) p.obj[(size_t)k * p.S + s] = 0.0;
                xrow[0] = INT_MIN;
            }
            T.sync();
            continue;
        }
        if (rc == 2 || fuel <= 0) {   // a bound that no valid instance reaches: refuse to answer rather than spin
            if (T.tl == 0) atomicMin(p.first_inf + k, -1LL);
            T.sync();
            continue;
        }
        // the optimal flow goes to the cut kernel
        if (nopen == 0 && T.tl == 0) xrow[0] = 0;
        K1_LOOP1
        for (int c = T.tl; c < nopen; c += TILE) xrow[c] = SI(w.x + c);
        holds = !forced;                                    // x and lab of this candidate serve the next one of the run
        if (srow && (p.state_io & 2) && j == p.K - 1 && holds) {   // ... and the first one of the next launch on this handle
            K1_LOOP1
            for (int v = T.tl; v <= nc; v += TILE) srow[1 + v] = SI(w.lab + v);
            K1_LOOP1
            for (int c = T.tl; c < nopen; c += TILE) srow[2 + nc + c] = SI(w.x + c);
            T.sync();
            if (T.tl == 0) srow[0] = 1;
        }
        K1_CLK(6);
        }
    }
}

// The CUT kernel: SPEC-LP potentials, lifting, fold — of the optimal flow the flow kernel left in xout.  A warp per (candidate,
// scenario), candidate-major and scenario-minor: neighbouring warps stream neighbouring rows and read the same plan.
template <int TILE, int NW, bool BIG>
__global__ void __launch_bounds__(NW * 32, NW == 8 ? SGUFP_K1_MINBLOCKS : NW == 4 ? 2 * SGUFP_K1_MINBLOCKS : 1) k1_cut_fold(K1Launch p, int words_per_tile) {
    const Lanes<TILE> T;
    constexpr int TILES_PER_CTA = NW * 32 / TILE;
    const int tile_in_cta = threadIdx.x / TILE;
    using TM = TileMemT<BIG>;
    TM w;
    tile_mem(w, tile_in_cta, words_per_tile, p);
    const long long items = (long long)p.K * p.S;
    const long long stride = (long long)gridDim.x * TILES_PER_CTA;
    K1_LOOPB
    for (long long item = (long long)blockIdx.x * TILES_PER_CTA + tile_in_cta; item < items; item += stride) {
        const int k = (int)(item / p.S), s = (int)(item - (long long)k * p.S);
        if (*reinterpret_cast<volatile long long *>(p.first_inf + k) < 0) continue;   // this candidate was aborted
        const int *xrow = p.xout + ((size_t)k * p.S + s) * p.xstride;
        if (xrow[0] == INT_MIN) continue;                                             // infeasible scenario: no flow, no cut term
        int fuel = 1 << 20;
        const PlanView P(p.plans + p.plan_off[k]);
        const int nch = P.h->nch, nopen = P.h->nopen, nc = P.h->nc, nav = P.h->nav, m = p.m;
        unsigned long long *sums = p.sums + (size_t)k * p.W;
        const double *row_u = p.cap_u + (size_t)s * p.m_pad, *row_l = p.cap_l + (size_t)s * p.m_pad;
#ifdef SGUFP_K1_STATS
        long long clk_ = clock64();
#endif
        stream_row<TILE>(P, nopen, m, p.m_pad, row_u, row_l, T, w);
        K1_LOOP2
        for (int c = T.tl; c < nopen; c += TILE) {
            const int xc = xrow[c];
            SI(w.x + c) = xc;
            RSET(c, (xc < (SI(w.up + c) >> HB) ? 1 : 0) | (xc > (SI(w.lo + c) >> HB) ? 2 : 0));
        }
        T.sync();
        K1_CLK(8);
        // 3. potentials
        canonical_potentials<TILE>(P, nopen, nc, T, w, fuel);
        K1_CLK(9);
        if (fuel <= 0) {   // a bound that no valid instance reaches: refuse to answer rather than spin
            if (T.tl == 0) atomicMin(p.first_inf + k, -1LL);
            T.sync();
            continue;
        }
        // 4. lifting + folding
        long long rhs = 0, objv = 0;
        // (a) per open chain: objective, the lower-bound multiplier's term, and r - dp kept in place of the flow (which is done):
        //     its positive part is the capacity multiplier of the chain, booked below by the arc that carries it
        K1_LOOPB
        for (int c = T.tl; c < nopen; c += TILE) {
            const int e = P.ch_ends[c], sv = (e & 0xffff) - 1, ev = (e >> 16) - 1, r = P.ch_r[c];
            const int g = r - (SI(w.pot + ev) - SI(w.pot + sv));
            objv += (long long)r * SI(w.x + c);
            if (g < 0) rhs -= (long long)(SI(w.lo + c) >> HB) * (-g);
            SI(w.x + c) = g;
        }
        // (b) alpha of every active V-bar node: the wire potential of its first matched pair, else the least start potential of
        //     the end-anchored chains leaving it, else 0 (rule 5)
        K1_LOOPB
        for (int i = T.tl; i < nav; i += TILE) {
            const int b0 = P.av_ptr[i], b1 = P.av_ptr[i + 1];
            int alpha = 0;
            if (b1 > b0) alpha = head_potential(P.av_arcs[b0], P, nopen, w);
            else {
                bool found = false;
                K1_LOOPB
                for (int t = P.fb_ptr[i]; t < P.fb_ptr[i + 1]; t++) {
                    const int c = P.fb_ch[t], ev = (P.ch_ends[c] >> 16) - 1;
                    const int cand = SI(w.pot + ev) - P.ch_r[c];
                    if (!found || cand < alpha) { alpha = cand; found = true; }
                }
            }
            SI(w.aq + i) = alpha;
        }
        T.sync();
        K1_CLK(10);
        // (c) per ARC, a lane each (coalesced plan reads, every lane busy; the per-node loop over matched pairs this replaces ran
        //     with 9 - 14 of 32 lanes behind chains of dependent loads): the capacity multiplier of an open chain sits on its first
        //     arc of least capacity; lambda - mu of a matched pair is its wire potential minus the node's alpha
        K1_LOOPB
        for (int a = T.tl; a < m; a += TILE) {
            const int cp = P.arc_cp[a], c = cp >> 10, pos = cp & 1023;
            const int2 av = P.arc_av[a];
            const int q = av.x;
            if (c < nopen) {
                const int g = SI(w.x + c), upw = SI(w.up + c);
                if (g > 0 && pos == (upw & 1023)) {
                    const long long v = (long long)(upw >> HB) * g;
                    if ((P.arc_info[a] & 3) == KIND_GAMMA) rhs += v; else atomicAdd(sums + 1 + p.L + a, (unsigned long long)v);
                }
            }
            if (q >= 0) {
                const int dl = head_potential(a, P, nopen, w) - SI(w.aq + q);
                if (dl != 0) {
                    long long v;
                    if (dl > 0) v = (long long)(int)row_u[a] * dl;
                    else v = (long long)(int)row_u[av.y] * (-dl);      // the matched out-arc
                    rhs += v;
                    atomicAdd(sums + 1 + (P.arc_info[a] >> 2) - 1, (unsigned long long)v);
                }
            }
        }
        K1_CLK(11);
        K1_LOOPB
        for (int c = nopen + T.tl; c < nch; c += TILE) {
            const int b0 = P.ch_ptr[c], b1 = P.ch_ptr[c + 1], first = P.ch_arcs[b0], last = P.ch_arcs[b1 - 1];
            const int e = P.ch_ends[c], sv = (e & 0xffff) - 1, ev = (e >> 16) - 1;
            const int q = P.ch_q[c], qs = (q & 0xffff) - 1, qe = (q >> 16) - 1;
            const int rfirst = P.arc_pre[first];
            if (qs >= 0 && qe >= 0 && b1 - b0 == 1) {
                const int v = rfirst - (SI(w.aq + qe) - SI(w.aq + qs));
                if (v > 0) atomicAdd(sums + 1 + p.L + first, (unsigned long long)((long long)(int)row_u[first] * v));
                continue;
            }
            if (qs >= 0) {
                const int phf = ev >= 0 ? SI(w.pot + ev) - (P.ch_r[c] - rfirst) : rfirst;
                const int v = rfirst - (phf - SI(w.aq + qs));
                if (v > 0) atomicAdd(sums + 1 + p.L + first, (unsigned long long)((long long)(int)row_u[first] * v));
            }
            if (qe >= 0) {
                int pt, rlast;
                if (b1 - b0 > 1) {
                    const int prev = P.ch_arcs[b1 - 2];
                    pt = (sv >= 0 ? SI(w.pot + sv) : 0) + P.arc_pre[prev];
                    rlast = P.ch_r[c] - P.arc_pre[prev];
                } else { pt = SI(w.pot + sv); rlast = P.ch_r[c]; }
                const int v = rlast - (SI(w.aq + qe) - pt);
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
        K1_CLK(12);
#ifdef SGUFP_K1_STATS
        if (T.tl == 0) atomicAdd(&g_k1_clk[13], 1ull);
#endif
    }
}

#ifndef SGUFP_K1_SMALL_TU
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
// a slab of na arcs, [na][S] int32 (the reference's per-arc vectors) -> columns a0 .. a0+na of [S][m_pad] fp64, 32x32 tiles through shared memory
__global__ void relayout_caps(const int32_t *__restrict__ src, double *__restrict__ dst, int na, int S, int m_pad, int a0) {
    __shared__ int tile[32][33];
    const int b0 = blockIdx.y * 32, s0 = blockIdx.x * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int a = b0 + r, s = s0 + threadIdx.x;
        tile[r][threadIdx.x] = (a < na && s < S) ? src[(size_t)a * S + s] : 0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int s = s0 + r, a = b0 + threadIdx.x;
        if (s < S && a < na) dst[(size_t)s * m_pad + a0 + a] = (double)tile[threadIdx.x][r];
    }
}

#endif  // SGUFP_K1_EMULATE
#endif  // SGUFP_K1_SMALL_TU

}  // namespace

#ifndef SGUFP_K1_EMULATE
template <int NW, bool BIG>
static cudaError_t launch_warp_nw(const K1Launch &p, cudaStream_t st, int sm_count, int *resident_warps, bool dry) {
    const int words = k1_words_per_tile(p);
    const size_t smem = (size_t)NW * words * sizeof(int);
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    // The dynamic shared-memory limit of a kernel is PROCESS-wide state (per device): host threads with batches of different
    // sizes must not lower it under each other (17 threads, one handle each: tests/test_cache_gpu.py).  It is raised once per
    // device to the most an SM offers; a launch then asks for what its batch needs.
    static std::atomic<unsigned long long> limit_raised{0};
    static thread_local int known_words = -1, known_per_sm = 0, known_dev = -1;   // per instantiation, thread and device: occupancy known
    int per_sm = 1, dev = 0;
    cudaGetDevice(&dev);
    if (!((limit_raised.load(std::memory_order_acquire) >> (dev & 63)) & 1ull)) {
        cudaError_t e = cudaFuncSetAttribute(k1_cut_eval<32, NW, BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k1_cut_fold<32, NW, BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        limit_raised.fetch_or(1ull << (dev & 63), std::memory_order_release);
    }
    if (known_words == words && known_dev == dev) per_sm = known_per_sm;
    else {
        const cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k1_cut_eval<32, NW, BIG>, NW * 32, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorInvalidConfiguration;
        known_words = words; known_per_sm = per_sm; known_dev = dev;
    }
    *resident_warps = per_sm * NW;
    if (dry) return cudaSuccess;
    const long long items = (long long)p.K * p.S;
    long long want = (items + NW - 1) / NW;
    long long grid = (long long)sm_count * per_sm;   // persistent: a whole number of CTAs per SM
    if (want < grid) grid = want;
    if (grid < 1) grid = 1;
    k1_cut_eval<32, NW, BIG><<<(unsigned)grid, NW * 32, smem, st>>>(p, words);
    if (const cudaError_t e = cudaGetLastError()) return e;
    k1_cut_fold<32, NW, BIG><<<(unsigned)grid, NW * 32, smem, st>>>(p, words);
    return cudaGetLastError();
}

// The kernels of the small size class are compiled in their own translation unit (k1_cut_small.cu: this file once more, with the
// compiler's own loop unrolling — measured faster there, while the larger graphs' kernels are faster without any: K1_LOOP).
#ifdef SGUFP_K1_SMALL_TU
cudaError_t k1_launch_small_class(int nw, const K1Launch &p, cudaStream_t st, int sm_count, int *rw, bool dry) {
    switch (nw) {
        case 8: return launch_warp_nw<8, false>(p, st, sm_count, rw, dry);
        case 4: return launch_warp_nw<4, false>(p, st, sm_count, rw, dry);
        case 2: return launch_warp_nw<2, false>(p, st, sm_count, rw, dry);
        default: return launch_warp_nw<1, false>(p, st, sm_count, rw, dry);
    }
}
#else
cudaError_t k1_launch_small_class(int nw, const K1Launch &p, cudaStream_t st, int sm_count, int *rw, bool dry);

static cudaError_t launch_warp_dispatch(int nw, const K1Launch &p, cudaStream_t st, int sm_count, int *rw, bool dry) {
    if (p.nc <= SMALL_NC) return k1_launch_small_class(nw, p, st, sm_count, rw, dry);
    switch (nw) {
        case 8: return launch_warp_nw<8, true>(p, st, sm_count, rw, dry);
        case 4: return launch_warp_nw<4, true>(p, st, sm_count, rw, dry);
        case 2: return launch_warp_nw<2, true>(p, st, sm_count, rw, dry);
        default: return launch_warp_nw<1, true>(p, st, sm_count, rw, dry);
    }
}

// Warps per CTA: the count that keeps most warps resident per SM (shared memory per warp grows with
// the network: 4.7 KB at C2, 6.8 KB at C4, tens of KB beyond); ties go to the larger CTA.
static cudaError_t launch_warp(const K1Launch &p, cudaStream_t st, int sm_count) {
    // the choice depends on the shared memory per warp alone: remember it (a one-candidate call is ~0.1 ms of kernel,
    // eight occupancy queries per call would show)
    static thread_local int cached_words = -1, cached_nw = 0, cached_dev = -1, cached_big = -1;
    const int words = k1_words_per_tile(p), big = p.nc > SMALL_NC;   // the size class picks the kernel instantiation: part of the key
    int rw = 0, dev = 0;
    cudaGetDevice(&dev);
    if (cached_words == words && cached_dev == dev && cached_big == big && cached_nw) {
        const cudaError_t e = launch_warp_dispatch(cached_nw, p, st, sm_count, &rw, false);
        if (e != cudaErrorInvalidConfiguration) return e;
        cudaGetLastError(); cached_nw = 0;      // probe again below
    }
    int best_nw = 0, best = 0;
    for (int nw = WARPS; nw >= 1; nw >>= 1) {
        const cudaError_t e = launch_warp_dispatch(nw, p, st, sm_count, &rw, true);
        if (e == cudaErrorInvalidConfiguration) { cudaGetLastError(); continue; }
        if (e != cudaSuccess) return e;
        if (rw > best) { best = rw; best_nw = nw; }
    }
    if (!best_nw) return cudaErrorInvalidConfiguration;
    cached_words = words; cached_nw = best_nw; cached_dev = dev; cached_big = big;
    return launch_warp_dispatch(best_nw, p, st, sm_count, &rw, false);
}

// Candidates per work item of the warp kernel (K1Launch::group).  A warm-started candidate costs about a third of one
// solved from zero flow, so long runs save work; but a launch ends with its slowest warp, and with A = K * S / (resident
// warps) evaluations per warp the time follows (A / g + 0.45) * (0.65 + 0.35 g) — the 0.45 is the measured cost of the last,
// partly filled round of items (C2, C4 on a B200: profiles/r02_k1_warm.md).  Least at g = sqrt(4 A); at most 32, in runs of
// equal length.  SGUFP_K1_GROUP=n overrides (1 = every candidate from zero flow).
int k1_group(int K, int S, int sm_count) {
    if (const char *e = getenv("SGUFP_K1_GROUP")) { const int g = atoi(e); return g < 1 ? 1 : g > K ? K : g; }
    const double A = (double)K * S / ((double)(sm_count > 0 ? sm_count : 1) * 32.0);
    int g = 1;
    while (g < 32 && g < K && (double)(g + 1) * (g + 1) <= 4.0 * A + (double)g + 1.0) g++;      // (g + 1/2)^2 <= 4 A: round to nearest
    const int runs = (K + g - 1) / g;
    return (K + runs - 1) / runs;
}

// which kernel: the lane-per-scenario one (k1_lane.cu) for batches it accepts, else the warp-per-scenario one
cudaError_t k1_launch(const K1Launch &p, cudaStream_t st, int sm_count, int *launches, bool *state_kept) {
    if (launches) (*launches)++;
    if (state_kept) *state_kept = false;     // only the warp kernels read and write K1Launch::state
    if (k1_lane_eligible(p, sm_count)) {
        const cudaError_t e = k1_lane_launch(p, st, sm_count);
        if (e != cudaErrorInvalidConfiguration) return e;
        cudaGetLastError();          // the state of this batch does not fit one SM's shared memory: the warp kernel takes it
    }
    if (launches) (*launches)++;     // the warp path is two kernels: flow, cut
    if (state_kept) *state_kept = p.state != nullptr && (p.state_io & 2);
    return launch_warp(p, st, sm_count);
}

cudaError_t ray_launch(const RayLaunch &p, cudaStream_t st, int *launches) {
    k1_ray<<<1, 32, 0, st>>>(p);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

cudaError_t relayout_launch(const int32_t *src, double *dst, int na, int S, int m_pad, int a0, cudaStream_t st, int *launches) {
    dim3 grid((S + 31) / 32, (na + 31) / 32), block(32, 8);
    relayout_caps<<<grid, block, 0, st>>>(src, dst, na, S, m_pad, a0);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

#endif  // SGUFP_K1_SMALL_TU
#endif  // SGUFP_K1_EMULATE

}  // namespace sgufp

#if defined(SGUFP_K1_STATS) && !defined(SGUFP_K1_EMULATE) && !defined(SGUFP_K1_SMALL_TU)
// debug build only (-DSGUFP_K1_STATS): relaxation passes / label computations since the last call
extern "C" int sgufp_debug_k1_clk(unsigned long long *out16) {
    unsigned long long z[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    cudaError_t e = cudaMemcpyFromSymbol(out16, sgufp::g_k1_clk, sizeof(z));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(sgufp::g_k1_clk, z, sizeof(z));
    return e == cudaSuccess ? 0 : -6;
}
extern "C" int sgufp_debug_k1_stats(unsigned long long *out8) {
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaError_t e = cudaMemcpyFromSymbol(out8, sgufp::g_k1_stats, sizeof(z));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(sgufp::g_k1_stats, z, sizeof(z));
    return e == cudaSuccess ? 0 : -6;
}
#endif

