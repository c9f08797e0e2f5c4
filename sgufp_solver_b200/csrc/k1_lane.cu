// k1_lane.cu — K1, lane-per-scenario kernel: one warp evaluates 32 consecutive scenarios of ONE candidate, lane = scenario.
//
// Replaces the scenario loop of GuroSolver::solveSubProblem (/root/reference/grb.cpp:162-360) for instances without
// positive lower bounds (the second-stage LP is then always feasible: zero flow).  Same SPEC-LP duals as the
// warp-per-scenario kernel of k1_cut.cu and as Oracle B, bit for bit (the duals depend on the optimal flow
// only through its residual graph, not on how the flow was found).
//
// Why a lane per scenario: every lane walks the candidate's chains in the same order, so the static half of a chain is
// one uniform load per warp-instruction for 32 scenarios, the per-scenario state sits in conflict-free shared-memory
// columns [index][lane], and nothing needs atomics or warp reductions until the fold.  What makes it pay is the flow
// algorithm, which has only two kinds of steps, both with warp-uniform control flow:
//   * a SWEEP over the open chains (uniform loop, per-lane predicated updates), and
//   * a STEP of a per-lane state machine (one in-slot examined / one hop of a path walked).
// Primal-dual with dual updates instead of label computations:
//   1. labels = shortest distances from the root at zero flow: ONE sweep (the chains are in topological order);
//   2. depth-first search BACKWARDS from the sink over tight residual in-arcs with current-arc cursors; a path that
//      reaches the root is saturated (two walks: bottleneck, push) and the search goes on from the head of the last
//      saturated arc; when the sink's cursor runs out the visited nodes D are exactly the nodes that reach the sink
//      over tight arcs but that the root cannot reach;
//   3. one sweep finds delta = the least slack of a residual arc entering D, the labels of D rise by delta (they stay
//      feasible potentials: every residual arc keeps a non-negative reduced cost), the search starts over;
//      done when the sink's label is >= 0 (no profitable path) or no residual arc enters D (maximum flow);
//   4. SPEC-LP potentials by label correction from the merged root over the final residual graph, lifting, fold
//      (one warp reduction per accumulator touched: 32 scenarios per atomic).
// Measured (profiles/r02_k1_lane.md): bit-identical cuts, the sweeps are as cheap as planned (33 per block of 32 scenarios on
// C2), but the per-lane search steps serialize (three modes, 124 warp-instructions per step, 15 of 32 lanes active) and a
// block waits for its slowest lane at every dual update, so the kernel is SLOWER than the warp-per-scenario one (C2 4.7 ms
// against 2.4, C4 70 against 28) and runs only on request (SGUFP_K1_MODE=lane).  No tensor cores: integer graph work.
#include "k1_cut.cuh"

#include <atomic>
#include <climits>
#include <cstdlib>
#include <type_traits>

#include "k1_common.cuh"
#include "model.hpp"

namespace sgufp {

#ifdef SGUFP_K1_STATS
__device__ unsigned long long g_k1_lane_stats[8];   // blocks, sweeps, steps (warp-level), dual updates (lane-level), pushes (lane-level), scenarios
#endif

namespace {

#ifdef SGUFP_K1_EMULATE
#define k1l_smem (reinterpret_cast<unsigned char *>(sgufp_emul_smem))
#else
extern __shared__ __align__(16) unsigned char k1l_smem_[];
#define k1l_smem k1l_smem_
#endif

// element types of the per-lane state: capacities, labels (packed: label * 4 | flags), cursors, path slots
template <class CT_, class LT_, class CUT_, class PST_>
struct LaneCfg { using CT = CT_; using LT = LT_; using CUT = CUT_; using PST = PST_; };
using CfgSmall = LaneCfg<uint8_t, int16_t, uint8_t, uint8_t>;      // C2 class: capacities < 256, at most 255 slots
using CfgMid = LaneCfg<uint8_t, int16_t, uint8_t, uint16_t>;       // C4 class
using CfgWide = LaneCfg<uint16_t, int32_t, uint16_t, uint16_t>;    // capacities < 65536, any rewards within the packing limits

template <class LT> struct LabLim;
template <> struct LabLim<int16_t> { static constexpr int inf = 0x7ffc, neg = -0x8000, max_sum_r = 8000; };
template <> struct LabLim<int32_t> { static constexpr int inf = 0x3ffffffc, neg = -0x40000000, max_sum_r = 1 << 18; };
template <class CT> struct CapLim;
template <> struct CapLim<uint8_t> { static constexpr int max = 255; };
template <> struct CapLim<uint16_t> { static constexpr int max = 65535; };

constexpr int F_DEAD = 1, F_ONP = 2;   // flag bits under a packed label: searched without finding the root / on the current path

struct LaneLayout { int o_lab, o_ps, o_cur, o_aq, o_rf, o_rb, bytes; };

template <class Cfg>
__host__ __device__ inline LaneLayout lane_layout(int max_nopen, int nc, int nav, int lw) {
    auto up4 = [](int x) { return (x + 3) & ~3; };
    LaneLayout y;
    const int nn = nc + 2;
    int off = 0;
    y.o_lab = off; off += up4(nn * (int)sizeof(typename Cfg::LT) * lw);
    // flow phase: path slots + cursors; lifting phase: alpha of the active V-bar nodes (same place)
    const int a = up4(nn * (int)sizeof(typename Cfg::PST) * lw), b = up4(nn * (int)sizeof(typename Cfg::CUT) * lw);
    const int q = up4((nav + 1) * (int)sizeof(typename Cfg::LT) * lw);
    y.o_ps = off; y.o_cur = off + a; y.o_aq = off;
    off += (a + b > q ? a + b : q);
    y.o_rf = off; off += up4(max_nopen * (int)sizeof(typename Cfg::CT) * lw);
    y.o_rb = off; off += up4(max_nopen * (int)sizeof(typename Cfg::CT) * lw);
    y.bytes = (off + 15) & ~15;
    return y;
}

template <int LW>
struct LaneWarp {
    int lane;
    __device__ LaneWarp() { lane = LW == 1 ? 0 : (int)(threadIdx.x & 31); }
    __device__ __forceinline__ bool any(bool p) const { return LW == 1 ? p : __any_sync(0xffffffffu, p) != 0; }
    __device__ __forceinline__ long long sum(long long v) const {
        if (LW > 1) for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
};

enum { M_RUN = 0, M_MIN = 1, M_PUSH = 2, M_STUCK = 3, M_DONE = 4 };

#define LX(a, i) (a)[(i) * LW]

template <int LW, class Cfg>
__global__ void __launch_bounds__(32) k1_lane_pd(K1Launch p) {
    using CT = typename Cfg::CT;
    using LT = typename Cfg::LT;
    using CUT = typename Cfg::CUT;
    using PST = typename Cfg::PST;
    constexpr int LINF = LabLim<LT>::inf, LNEG = LabLim<LT>::neg;
    const LaneWarp<LW> L;
    const LaneLayout lay = lane_layout<Cfg>(p.max_nopen, p.nc, p.nav, LW);
    unsigned char *sm = k1l_smem;
    LT *lab = reinterpret_cast<LT *>(sm + lay.o_lab) + L.lane;      // [nc+2] packed label * 4 | flags; plain potential after the flow phase
    PST *psl = reinterpret_cast<PST *>(sm + lay.o_ps) + L.lane;     // [nc+2] the in-slot through which the search entered a node
    CUT *cur = reinterpret_cast<CUT *>(sm + lay.o_cur) + L.lane;    // [nc+2] cursor into a node's in-slots
    LT *aq = reinterpret_cast<LT *>(sm + lay.o_aq) + L.lane;        // [nav]  alpha of the active V-bar nodes (lifting)
    CT *rf = reinterpret_cast<CT *>(sm + lay.o_rf) + L.lane;        // [nopen] residual capacity forward  (up - x)
    CT *rb = reinterpret_cast<CT *>(sm + lay.o_rb) + L.lane;        // [nopen] residual capacity backward (x - lo, lo = 0)

    const int bps = (p.S + LW - 1) / LW;                            // scenario blocks per candidate
    const long long nblk = (long long)p.K * bps;
    for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const int k = (int)(blk / bps), s_raw = (int)(blk - (long long)k * bps) * LW + L.lane;
        const bool active = s_raw < p.S;
        const int s = active ? s_raw : p.S - 1;                     // idle lanes shadow the last scenario; they never write results
        if (*reinterpret_cast<volatile long long *>(p.first_inf + k) < 0) continue;
        int fuel = 1 << 24;                                         // steps + sweeps a block may spend (warp-uniform)
        const PlanView P(p.plans + p.plan_off[k]);
        const int nch = P.h->nch, nopen = P.h->nopen, nc = P.h->nc, nav = P.h->nav, m = p.m;
        unsigned long long *sums = p.sums + (size_t)k * p.W;
        const double *row_u = p.cap_u + (size_t)s * p.m_pad;

        // 1. chain capacities: the lane streams its own row front to back (128-bit loads; a 32-byte sector is used up by
        //    two consecutive iterations); closed chains carry no flow
        for (int c = 0; c < nopen; c++) LX(rf, c) = (CT)CapLim<CT>::max;
        {
            const double2 *ru = reinterpret_cast<const double2 *>(row_u);
            const int2 *cp2 = reinterpret_cast<const int2 *>(P.arc_cp);
            for (int a2 = 0; a2 < p.m_pad / 2; a2++) {
                const double2 u2 = __ldg(ru + a2);
                const int2 cp = cp2[a2];                            // uniform; arc_cp is padded to an even count
                const int c0 = cp.x >> 10, c1 = cp.y >> 10;
                if (c0 < nopen) { const int u = (int)u2.x, o = LX(rf, c0); if (u < o) LX(rf, c0) = (CT)u; }
                if (2 * a2 + 1 < m && c1 < nopen) { const int u = (int)u2.y, o = LX(rf, c1); if (u < o) LX(rf, c1) = (CT)u; }
            }
        }
        // 2. zero flow; labels at zero flow in one sweep (chains are sorted by the depth of their tails: model.cpp)
        for (int v = 0; v <= nc + 1; v++) { LX(lab, v) = (LT)LINF; LX(cur, v) = 0; }
        LX(lab, 0) = 0;
        for (int c = 0; c < nopen; c++) {
            const ChainEnds e(P.ch_st[c]);
            LX(rb, c) = 0;
            const int ls = LX(lab, e.sv);
            if (LX(rf, c) > 0 && ls < LINF) { const int cand = ls - 4 * e.r; if (cand < LX(lab, e.hf)) LX(lab, e.hf) = (LT)cand; }
        }
        // 3. optimal flow
        int mode = LX(lab, nc) < 0 ? M_RUN : M_DONE;               // an unreachable sink carries LINF > 0
        int v = nc, w = 0, d = 0, hop = 0, jlast = 0;
#ifdef SGUFP_K1_STATS
        unsigned long long st_sweeps = 0, st_steps = 0, st_dual = 0, st_push = 0;
#endif
        while (fuel > 0) {
            while (L.any(mode <= M_PUSH) && --fuel > 0) {
#ifdef SGUFP_K1_STATS
                st_steps++;
#endif
                if (mode == M_RUN) {
                    if (v == 0) { mode = M_MIN; w = 0; d = INT_MAX; hop = 0; jlast = 0; }   // the root: a tight path root -> sink stands in psl[]
                    else {
                        const int pd = P.in_pd[v], b0 = pd & 0xffff, deg = pd >> 16, e = LX(cur, v);
                        if (e >= deg) {                             // no tight in-arc left: v cannot be reached from the root
                            LX(lab, v) = (LT)((LX(lab, v) & ~3) | F_DEAD);
                            if (v == nc) mode = M_STUCK;
                            else { v = P.slots[LX(psl, v)].y & 0xfff; LX(cur, v) = (CUT)(LX(cur, v) + 1); }   // back to the head it was entered from
                        } else {
                            const int2 rec = P.slots[b0 + e];
                            const int t = rec.x & 0xffff, cd = (int)((unsigned)rec.x >> 16), c = cd >> 1;
                            const int res = (cd & 1) ? LX(rb, c) : LX(rf, c);
                            const int plt = LX(lab, t);
                            if (res > 0 && plt < LINF && !(plt & 3) && plt + ((rec.y >> 12) << 2) == (LX(lab, v) & ~3)) {
                                LX(psl, t) = (PST)(b0 + e);
                                if (t) LX(lab, t) = (LT)(plt | F_ONP);
                                v = t;
                            } else LX(cur, v) = (CUT)(e + 1);
                        }
                    }
                } else if (mode == M_MIN) {                         // walk root -> sink: bottleneck, and the LAST arc that attains it
                    const int2 rec = P.slots[LX(psl, w)];
                    const int cd = (int)((unsigned)rec.x >> 16), c = cd >> 1;
                    const int res = (cd & 1) ? LX(rb, c) : LX(rf, c);
                    if (res <= d) { d = res; jlast = hop; }
                    hop++; w = rec.y & 0xfff;
                    if (w == nc) { mode = M_PUSH; w = 0; hop = 0; }
                } else if (mode == M_PUSH) {                        // walk again: move d units; the search resumes at the head of arc jlast
                    const int2 rec = P.slots[LX(psl, w)];
                    const int cd = (int)((unsigned)rec.x >> 16), c = cd >> 1;
                    int f = LX(rf, c), b = LX(rb, c);
                    if (cd & 1) { b -= d; f += d; } else { f -= d; b += d; }
                    LX(rf, c) = (CT)f; LX(rb, c) = (CT)b;
                    if (hop <= jlast && w != 0) LX(lab, w) = (LT)(LX(lab, w) & ~F_ONP);   // the nodes up to that arc leave the path
                    w = rec.y & 0xfff;
                    if (hop == jlast) v = w;
                    hop++;
                    if (w == nc) {
                        mode = M_RUN;
#ifdef SGUFP_K1_STATS
                        st_push++;
#endif
                    }
                }
            }
            if (!L.any(mode == M_STUCK)) break;
            // dual update: D = the nodes flagged dead (the sink among them); delta = least slack of a residual arc entering D
            int delta = INT_MAX;
            const int pl_sink = LX(lab, nc);
            for (int c = 0; c < nopen; c++) {
                const ChainEnds e(P.ch_st[c]);
                const int f = LX(rf, c), b = LX(rb, c), ps = LX(lab, e.sv), pe = LX(lab, e.ev);
                const int hf = e.ev ? pe : pl_sink, hb = e.sv ? ps : pl_sink;           // an arc entering the root ends at the sink
                if (f > 0 && ps < LINF && !(ps & F_DEAD) && (hf & F_DEAD)) delta = min(delta, (ps & ~3) - 4 * e.r - (hf & ~3));
                if (b > 0 && pe < LINF && !(pe & F_DEAD) && (hb & F_DEAD)) delta = min(delta, (pe & ~3) + 4 * e.r - (hb & ~3));
            }
            --fuel;
#ifdef SGUFP_K1_STATS
            st_sweeps++;
            if (mode == M_STUCK) st_dual++;
#endif
            const bool go = mode == M_STUCK && delta != INT_MAX;
            if (mode == M_STUCK && !go) mode = M_DONE;              // no residual arc enters D: the flow is maximum
            for (int u = 1; u <= nc; u++) {
                if (go) {
                    int pl = LX(lab, u);
                    if (pl & F_DEAD) pl += delta;
                    LX(lab, u) = (LT)(pl & ~3);
                    LX(cur, u) = 0;
                }
            }
            if (go) { v = nc; mode = LX(lab, nc) < 0 ? M_RUN : M_DONE; }
        }
        if (fuel <= 0) {   // a bound that no valid instance reaches: refuse to answer rather than spin
            if (L.lane == 0) atomicMin(p.first_inf + k, -1LL);
            continue;
        }
        // 4. SPEC-LP potentials: shortest residual distances from the MERGED root (arcs entering the root
        //    are not relaxed), the two completions for nodes the root cannot reach; lab[] becomes the plain potential
        for (int u = 1; u <= nc + 1; u++) LX(lab, u) = (LT)LINF;
        {
            bool changed;
            do {
                changed = false;
                for (int c = 0; c < nopen; c++) {
                    const ChainEnds e(P.ch_st[c]);
                    const int ls = LX(lab, e.sv), le = LX(lab, e.ev);
                    if (e.ev && LX(rf, c) > 0 && ls < LINF) { const int cand = ls - 4 * e.r; if (cand < le) { LX(lab, e.ev) = (LT)cand; changed = true; } }
                    if (e.sv && LX(rb, c) > 0 && le < LINF) { const int cand = le + 4 * e.r; if (cand < LX(lab, e.sv)) { LX(lab, e.sv) = (LT)cand; changed = true; } }
                }
#ifdef SGUFP_K1_STATS
                st_sweeps++;
#endif
                changed = L.any(changed) && --fuel > 0;
            } while (changed);
        }
        bool missing = false;
        for (int u = 0; u < nc; u++) {
            const int l = LX(lab, u);
            if (l >= LINF) { missing = true; LX(lab, u) = (LT)LNEG; } else LX(lab, u) = (LT)(l | 1);   // flag 1: labelled by phase 1
        }
        if (L.any(missing)) {
            bool changed;
            do {   // least labels consistent with the labelled nodes
                changed = false;
                for (int c = 0; c < nopen; c++) {
                    const int e = P.ch_ends[c], sv = (e & 0xffff) - 1, ev = (e >> 16) - 1, r4 = 4 * P.ch_r[c];
                    const int ps = LX(lab, sv), pe = LX(lab, ev);
                    if (LX(rf, c) > 0 && !(ps & 3) && pe > LNEG) { const int cand = (pe & ~3) + r4; if (cand > ps) { LX(lab, sv) = (LT)cand; changed = true; } }
                    const int ps2 = LX(lab, sv);
                    if (LX(rb, c) > 0 && !(pe & 3) && ps2 > LNEG) { const int cand = (ps2 & ~3) - r4; if (cand > pe) { LX(lab, ev) = (LT)cand; changed = true; } }
                }
                changed = L.any(changed) && --fuel > 0;
            } while (changed);
            bool iso = false;
            for (int u = 0; u < nc; u++)
                if (LX(lab, u) == (LT)LNEG) { LX(lab, u) = 2; iso = true; }          // label 0, flag 2: cut off both ways
            if (L.any(iso)) {
                do {   // zero-rooted completion
                    changed = false;
                    for (int c = 0; c < nopen; c++) {
                        const int e = P.ch_ends[c], sv = (e & 0xffff) - 1, ev = (e >> 16) - 1, r4 = 4 * P.ch_r[c];
                        const int ps = LX(lab, sv), pe = LX(lab, ev);
                        if (LX(rf, c) > 0 && (pe & 2)) { const int cand = (ps & ~3) - r4; if (cand < (pe & ~3)) { LX(lab, ev) = (LT)(cand | 2); changed = true; } }
                        const int pe2 = LX(lab, ev);
                        if (LX(rb, c) > 0 && (ps & 2)) { const int cand = (pe2 & ~3) + r4; if (cand < (ps & ~3)) { LX(lab, sv) = (LT)(cand | 2); changed = true; } }
                    }
                    changed = L.any(changed) && --fuel > 0;
                } while (changed);
            }
        }
        for (int u = 0; u < nc; u++) LX(lab, u) = (LT)(-(LX(lab, u) >> 2));   // potential = -distance (arithmetic shift: the flags fall off)
        if (fuel <= 0) {
            if (L.lane == 0) atomicMin(p.first_inf + k, -1LL);
            continue;
        }
        // 5. lifting + folding: one reduction over the 32 scenarios per accumulator touched
        const bool ok = active;
        auto head_potential = [&](int a) -> int {   // wire potential at the HEAD of arc a (cf. head_potential in k1_cut.cu)
            const int cp = P.arc_cp[a], c = cp >> 10, pos = cp & 1023, pre = P.arc_pre[a];
            const int e = P.ch_ends[c], sv = (e & 0xffff) - 1, ev = (e >> 16) - 1;
            if (c < nopen) {
                const int r = P.ch_r[c], psv = LX(lab, sv), dp = LX(lab, ev) - psv;
                const int g = max(0, r - dp), b = max(0, dp - r);
                int val = psv + pre;
                const int b0 = P.ch_ptr[c], b1 = P.ch_ptr[c + 1];
                if (L.any(g > 0)) {    // the capacity multiplier sits on the FIRST arc of least capacity
                    bool found = false;
                    const int up = (int)LX(rf, c) + (int)LX(rb, c);
                    for (int kk = b0; kk <= b0 + pos; kk++) found |= (int)row_u[P.ch_arcs[kk]] == up;
                    if (found) val -= g;
                }
                if (b > 0 && b0 + pos == b1 - 1) val += b;   // the lower-bound multiplier on the LAST arc of greatest lower bound: all are 0 here
                return val;
            }
            if (sv >= 0) return LX(lab, sv) + pre;
            if (ev >= 0) return LX(lab, ev) - (P.ch_r[c] - pre);
            return pre;
        };
        long long rhs = 0, objv = 0;
        for (int c = 0; c < nopen; c++) {
            const int e = P.ch_ends[c], sv = (e & 0xffff) - 1, ev = (e >> 16) - 1, r = P.ch_r[c];
            const int g = r - (LX(lab, ev) - LX(lab, sv));
            const int up = (int)LX(rf, c) + (int)LX(rb, c);
            objv += (long long)r * (int)LX(rb, c);
            if (L.any(ok && g > 0)) {
                bool found = false;
                for (int kk = P.ch_ptr[c]; kk < P.ch_ptr[c + 1]; kk++) {
                    const int a = P.ch_arcs[kk];
                    const bool hit = ok && g > 0 && !found && (int)row_u[a] == up;
                    found |= hit;
                    if (L.any(hit)) {
                        const long long tot = L.sum(hit ? (long long)up * g : 0);
                        if (L.lane == 0) {
                            if ((P.arc_info[a] & 3) == KIND_GAMMA) atomicAdd(sums, (unsigned long long)tot);
                            else atomicAdd(sums + 1 + p.L + a, (unsigned long long)tot);
                        }
                    }
                }
            }
        }
        for (int i = 0; i < nav; i++) {
            const int b0 = P.av_ptr[i], b1 = P.av_ptr[i + 1];
            int alpha = 0;
            if (b1 > b0) {
                alpha = head_potential(P.av_arcs[b0]);
                for (int t = b0 + 1; t < b1; t++) {
                    const int a = P.av_arcs[t];
                    const int dl = head_potential(a) - alpha;
                    if (L.any(ok && dl != 0)) {
                        long long val = 0;
                        if (ok && dl > 0) val = (long long)(int)row_u[a] * dl;
                        else if (ok && dl < 0) {
                            const int cp = P.arc_cp[a];
                            val = (long long)(int)row_u[P.ch_arcs[P.ch_ptr[cp >> 10] + (cp & 1023) + 1]] * (-dl);
                        }
                        const long long tot = L.sum(val);
                        rhs += val;
                        if (tot != 0 && L.lane == 0) atomicAdd(sums + 1 + (P.arc_info[a] >> 2) - 1, (unsigned long long)tot);
                    }
                }
            } else {
                bool found = false;
                for (int t = P.fb_ptr[i]; t < P.fb_ptr[i + 1]; t++) {
                    const int c = P.fb_ch[t], ev = (P.ch_ends[c] >> 16) - 1;
                    const int cand = LX(lab, ev) - P.ch_r[c];
                    if (!found || cand < alpha) { alpha = cand; found = true; }
                }
            }
            LX(aq, i) = (LT)alpha;
        }
        for (int c = nopen; c < nch; c++) {
            const int b0 = P.ch_ptr[c], b1 = P.ch_ptr[c + 1], first = P.ch_arcs[b0], last = P.ch_arcs[b1 - 1];
            const int e = P.ch_ends[c], sv = (e & 0xffff) - 1, ev = (e >> 16) - 1;
            const int q = P.ch_q[c], qs = (q & 0xffff) - 1, qe = (q >> 16) - 1;
            const int rfirst = P.arc_pre[first];
            if (qs >= 0 && qe >= 0 && b1 - b0 == 1) {
                const int val = rfirst - (LX(aq, qe) - LX(aq, qs));
                if (L.any(ok && val > 0)) {
                    const long long tot = L.sum(ok && val > 0 ? (long long)(int)row_u[first] * val : 0);
                    if (L.lane == 0) atomicAdd(sums + 1 + p.L + first, (unsigned long long)tot);
                }
                continue;
            }
            if (qs >= 0) {
                const int phf = ev >= 0 ? LX(lab, ev) - (P.ch_r[c] - rfirst) : rfirst;
                const int val = rfirst - (phf - LX(aq, qs));
                if (L.any(ok && val > 0)) {
                    const long long tot = L.sum(ok && val > 0 ? (long long)(int)row_u[first] * val : 0);
                    if (L.lane == 0) atomicAdd(sums + 1 + p.L + first, (unsigned long long)tot);
                }
            }
            if (qe >= 0) {
                int pt, rlast;
                if (b1 - b0 > 1) {
                    const int prev = P.ch_arcs[b1 - 2];
                    pt = (sv >= 0 ? LX(lab, sv) : 0) + P.arc_pre[prev];
                    rlast = P.ch_r[c] - P.arc_pre[prev];
                } else { pt = LX(lab, sv); rlast = P.ch_r[c]; }
                const int val = rlast - (LX(aq, qe) - pt);
                if (L.any(ok && val > 0)) {
                    const long long tot = L.sum(ok && val > 0 ? (long long)(int)row_u[last] * val : 0);
                    if (L.lane == 0) atomicAdd(sums + 1 + p.L + last, (unsigned long long)tot);
                }
            }
        }
        const long long rtot = L.sum(ok ? rhs : 0);
        if (rtot != 0 && L.lane == 0) atomicAdd(sums, (unsigned long long)rtot);
        if (active) {
            if (p.status) p.status[(size_t)k * p.S + s] = 0;
            if (p.obj) p.obj[(size_t)k * p.S + s] = (double)objv;
        }
#ifdef SGUFP_K1_STATS
        {
            const long long dsum = L.sum((long long)st_dual), psum = L.sum((long long)st_push);
            if (L.lane == 0) {
                atomicAdd(&g_k1_lane_stats[0], 1ull); atomicAdd(&g_k1_lane_stats[1], st_sweeps); atomicAdd(&g_k1_lane_stats[2], st_steps);
                atomicAdd(&g_k1_lane_stats[3], (unsigned long long)dsum); atomicAdd(&g_k1_lane_stats[4], (unsigned long long)psum);
                atomicAdd(&g_k1_lane_stats[5], (unsigned long long)min(LW, p.S - (int)(blk - (long long)k * bps) * LW));
            }
        }
#endif
    }
}
#undef LX

// which state widths a batch needs; 0 = the lane kernel does not take it
template <class Cfg>
__host__ inline bool lane_cfg_fits(const K1Launch &p) {
    return p.max_cap <= CapLim<typename Cfg::CT>::max && p.sum_abs_r < LabLim<typename Cfg::LT>::max_sum_r &&
           p.max_indeg < (1 << (8 * (int)sizeof(typename Cfg::CUT))) - 1 && 2 * p.max_nopen < (1 << (8 * (int)sizeof(typename Cfg::PST)));
}

}  // namespace

#ifndef SGUFP_K1_EMULATE
static int k1_mode_env() {   // 0 auto, 1 warp kernel only, 2 lane kernel whenever it accepts the batch
    const char *e = getenv("SGUFP_K1_MODE");
    if (!e) return 0;
    return e[0] == 'w' ? 1 : e[0] == 'l' ? 2 : 0;
}

bool k1_lane_tables_wanted() { return k1_mode_env() == 2; }

bool k1_lane_eligible(const K1Launch &p, int sm_count) {
    const int mode = k1_mode_env();
    if (mode == 1 || p.has_lower || !p.lane_tables || p.S < 1) return false;
    if (!lane_cfg_fits<CfgSmall>(p) && !lane_cfg_fits<CfgMid>(p) && !lane_cfg_fits<CfgWide>(p)) return false;
    // Measured (profiles/r02_k1_lane.md): the per-lane search steps diverge (15 of 32 lanes active, 38 % issue-active on C2;
    // 5 single-warp CTAs per SM on C4) and the kernel is slower than the warp-per-scenario one on C2, C4 and C5, so it runs
    // only when asked for; it stays as the second, independently written K1 that the parity tests compare against.
    (void)sm_count;
    return mode == 2;
}

template <class Cfg>
static cudaError_t lane_launch_cfg(const K1Launch &p, cudaStream_t st, int sm_count) {
    const LaneLayout lay = lane_layout<Cfg>(p.max_nopen, p.nc, p.nav, 32);
    if (lay.bytes > 227 * 1024) return cudaErrorInvalidConfiguration;
    static std::atomic<unsigned long long> limit_raised{0};    // the shared-memory limit of a kernel is process-wide state: raised once per device (k1_cut.cu)
    static thread_local int known_bytes = -1, known_per_sm = 0, known_dev = -1;   // per instantiation, thread and device
    int per_sm = 1, dev = 0;
    cudaGetDevice(&dev);
    if (!((limit_raised.load(std::memory_order_acquire) >> (dev & 63)) & 1ull)) {
        const cudaError_t e = cudaFuncSetAttribute(k1_lane_pd<32, Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        limit_raised.fetch_or(1ull << (dev & 63), std::memory_order_release);
    }
    if (known_bytes == lay.bytes && known_dev == dev) per_sm = known_per_sm;
    else {
        const cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k1_lane_pd<32, Cfg>, 32, (size_t)lay.bytes);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorInvalidConfiguration;
        known_bytes = lay.bytes; known_per_sm = per_sm; known_dev = dev;
    }
    const long long nblk = (long long)p.K * ((p.S + 31) / 32);
    long long grid = (long long)sm_count * per_sm;   // persistent: single-warp CTAs, as many per SM as the shared memory holds
    if (nblk < grid) grid = nblk;
    if (grid < 1) grid = 1;
    k1_lane_pd<32, Cfg><<<(unsigned)grid, 32, (size_t)lay.bytes, st>>>(p);
    return cudaGetLastError();
}

cudaError_t k1_lane_launch(const K1Launch &p, cudaStream_t st, int sm_count) {
    if (lane_cfg_fits<CfgSmall>(p)) return lane_launch_cfg<CfgSmall>(p, st, sm_count);
    if (lane_cfg_fits<CfgMid>(p)) return lane_launch_cfg<CfgMid>(p, st, sm_count);
    if (lane_cfg_fits<CfgWide>(p)) return lane_launch_cfg<CfgWide>(p, st, sm_count);
    return cudaErrorInvalidConfiguration;
}
#endif  // SGUFP_K1_EMULATE

}  // namespace sgufp

#if defined(SGUFP_K1_STATS) && !defined(SGUFP_K1_EMULATE)
extern "C" int sgufp_debug_k1_lane_stats(unsigned long long *out8) {
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaError_t e = cudaMemcpyFromSymbol(out8, sgufp::g_k1_lane_stats, sizeof(z));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(sgufp::g_k1_lane_stats, z, sizeof(z));
    return e == cudaSuccess ? 0 : -6;
}
#endif
