// model.cpp — see model.hpp.  Host-only, no CUDA.
#include "model.hpp"

#include <algorithm>
#include <climits>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <cstring>
#include <functional>
#include <numeric>

#include "../../include/sgufp_b200.h"

namespace sgufp {

// Same visiting order as Network::shuffleVBarNodes (Network.cpp:132-186): V-bar demand points
// first (a demand point has exactly one out-arc, into node n-1), then breadth-first towards the
// sources, a node taking its place the first time it shows up as a parent.
static int order_vbar(const Model &M, std::vector<int32_t> &vb, std::string &err) {
    const int n = M.n;
    std::vector<uint8_t> waiting(n, 0);
    int remaining = 0;
    for (int v : vb)
        if (!waiting[v]) { waiting[v] = 1; remaining++; }
    std::vector<int32_t> order, frontier;
    for (int v = 0; v < n; v++) {
        if (M.outdeg(v) == 1 && M.head[M.out_arc[M.out_ptr[v]]] == n - 1) {
            frontier.push_back(v);
            if (waiting[v]) { order.push_back(v); waiting[v] = 0; remaining--; }
        }
    }
    std::vector<uint8_t> mark(n);
    while (remaining > 0) {
        std::vector<int32_t> up;
        std::fill(mark.begin(), mark.end(), 0);
        for (int c : frontier)
            for (int e = M.in_ptr[c]; e < M.in_ptr[c + 1]; e++) {
                int p = M.tail[M.in_arc[e]];
                if (!mark[p]) { mark[p] = 1; up.push_back(p); }
            }
        if (up.empty()) {
            err = "a V-bar node is not a backward ancestor of any demand point: the reference's shuffleVBarNodes "
                  "(Network.cpp:159-184) never terminates on this instance";
            return SGUFP_ERR_INSTANCE;
        }
        for (int p : up)
            if (waiting[p]) { order.push_back(p); waiting[p] = 0; remaining--; }
        frontier.swap(up);
    }
    vb.swap(order);
    return 0;
}

int Model::build(int n_, int m_, const int32_t *tail_, const int32_t *head_, const int32_t *rew0, const int32_t *vbar_,
                 int nvbar, std::string &err) {
    n = n_; m = m_;
    if (n < 2 || m < 1 || !tail_ || !head_ || !rew0 || (nvbar > 0 && !vbar_)) { err = "bad sizes or null arrays"; return SGUFP_ERR_ARG; }
    if (n > 65535 || m > 32767) { err = "node ids must fit 16 bits (Cut.h:342-344) and arc ids int16 (DD.h:424)"; return SGUFP_ERR_LIMITS; }
    tail.assign(tail_, tail_ + m); head.assign(head_, head_ + m); rew.assign(rew0, rew0 + m);
    in_ptr.assign(n + 1, 0); out_ptr.assign(n + 1, 0);
    for (int a = 0; a < m; a++) {
        if (tail[a] < 0 || tail[a] >= n || head[a] < 0 || head[a] >= n || tail[a] == head[a]) { err = "arc endpoint out of range or self-loop"; return SGUFP_ERR_ARG; }
        in_ptr[head[a] + 1]++; out_ptr[tail[a] + 1]++;
    }
    for (int v = 0; v < n; v++) { in_ptr[v + 1] += in_ptr[v]; out_ptr[v + 1] += out_ptr[v]; }
    in_arc.assign(m, 0); out_arc.assign(m, 0); out_index.assign(m, 0);
    {
        std::vector<int32_t> fi(n, 0), fo(n, 0);
        for (int a = 0; a < m; a++) {  // file order, as incomingArcs/outgoingArcs are filled (Network.cpp:45-48)
            in_arc[in_ptr[head[a]] + fi[head[a]]++] = a;
            out_index[a] = fo[tail[a]];
            out_arc[out_ptr[tail[a]] + fo[tail[a]]++] = a;
        }
    }
    // every dual variable is indexed by NODE pairs (grb.h:44-51): parallel arcs would alias
    for (int v = 0; v < n; v++) {
        std::vector<int32_t> hs;
        for (int e = out_ptr[v]; e < out_ptr[v + 1]; e++) hs.push_back(head[out_arc[e]]);
        std::sort(hs.begin(), hs.end());
        if (std::adjacent_find(hs.begin(), hs.end()) != hs.end()) { err = "parallel arcs: the reference keys beta/gamma/y by node pairs (grb.cpp:148)"; return SGUFP_ERR_INSTANCE; }
    }
    // acyclicity (Kahn)
    {
        std::vector<int32_t> deg(n), st;
        for (int v = 0; v < n; v++) { deg[v] = indeg(v); if (!deg[v]) st.push_back(v); }
        int seen = 0;
        while (!st.empty()) {
            int v = st.back(); st.pop_back(); seen++;
            for (int e = out_ptr[v]; e < out_ptr[v + 1]; e++) if (--deg[head[out_arc[e]]] == 0) st.push_back(head[out_arc[e]]);
        }
        if (seen != n) { err = "network has a directed cycle"; return SGUFP_ERR_CYCLIC; }
    }
    is_vbar.assign(n, 0);
    std::vector<int32_t> vb(vbar_, vbar_ + nvbar);
    for (int v : vb) {
        if (v < 0 || v >= n) { err = "V-bar id out of range"; return SGUFP_ERR_ARG; }
        is_vbar[v] = 1;
    }
    for (int v : vb)
        if (indeg(v) == 0) { err = "V-bar node without in-arcs: stateUpdateMap.insert (Network.cpp:102) would drop the next node's states"; return SGUFP_ERR_INSTANCE; }
    for (int a = 0; a < m; a++)
        if (indeg(tail[a]) == 0 && outdeg(head[a]) == 0) { err = "source->sink arc: the reference clears A4 and gives it no row (Network.cpp:80)"; return SGUFP_ERR_INSTANCE; }
    if (int rc = order_vbar(*this, vb, err)) return rc;
    vbar = vb;
    active.assign(n, 0); is_root.assign(n, 0);
    for (int v = 0; v < n; v++) {
        // no conservation row => no alpha in A1/A2 rows (grb.cpp:54,72); alpha[0] = alpha[n-1] = 0 (grb.cpp:134-135)
        is_root[v] = (indeg(v) == 0 || outdeg(v) == 0 || v == 0 || v == n - 1);
        active[v] = is_vbar[v] && indeg(v) > 0 && outdeg(v) > 0;
        if (active[v] && is_root[v]) { err = "node 0 / n-1 is an interior V-bar node"; return SGUFP_ERR_INSTANCE; }
    }
    // processingOrder (Network.cpp:94-118) and the coefficient keys (grb.h:60-68)
    L = 0;
    for (int q : vbar) L += indeg(q);
    layer_arc.assign(L, 0); arc_layer.assign(m, -1); slot_base.assign(L + 1, 0);
    int ell = 0; T = 0;
    for (int q : vbar)
        for (int e = in_ptr[q]; e < in_ptr[q + 1]; e++) {
            layer_arc[ell] = in_arc[e]; arc_layer[in_arc[e]] = ell; slot_base[ell] = T; T += outdeg(q); ell++;
        }
    slot_base[L] = T;
    slot_in.assign(T, 0); slot_out.assign(T, 0);
    for (ell = 0; ell < L; ell++) {
        int a = layer_arc[ell], q = head[a];
        for (int k = 0; k < outdeg(q); k++) { slot_in[slot_base[ell] + k] = a; slot_out[slot_base[ell] + k] = out_arc[out_ptr[q] + k]; }
    }
    slot_sorted.resize(T);
    std::iota(slot_sorted.begin(), slot_sorted.end(), 0);
    std::sort(slot_sorted.begin(), slot_sorted.end(), [&](int a, int b) {  // std::map<tuple<i,q,j>> order (Cut.h:75)
        int ia = tail[slot_in[a]], ib = tail[slot_in[b]];
        if (ia != ib) return ia < ib;
        int qa = head[slot_in[a]], qb = head[slot_in[b]];
        if (qa != qb) return qa < qb;
        return head[slot_out[a]] < head[slot_out[b]];
    });
    slot_lex_rank.assign(T, 0);
    for (int r = 0; r < T; r++) slot_lex_rank[slot_sorted[r]] = r;
    slot_key_sorted.resize(T);
    for (int r = 0; r < T; r++) {
        const int s = slot_sorted[r];
        const uint64_t i = (uint64_t)tail[slot_in[s]], q = (uint64_t)head[slot_in[s]], j = (uint64_t)head[slot_out[s]];
        slot_key_sorted[r] = q | (i << 16) | (j << 32);
    }
    key_slot.clear();
    key_slot.reserve((size_t)T * 2);
    for (int s = 0; s < T; s++) {   // emplace keeps the FIRST slot of a key: Inavap::Cut::get returns the first match (Cut.h:275-282)
        const uint64_t i = (uint64_t)tail[slot_in[s]], q = (uint64_t)head[slot_in[s]], j = (uint64_t)head[slot_out[s]];
        key_slot.emplace(q | (i << 16) | (j << 32), s);
    }
    cn.assign(n, 0); av_index.assign(n, -1); av_node.clear(); nc = 1; nav = 0;
    for (int v = 0; v < n; v++) {
        if (is_root[v]) cn[v] = 0;
        else if (active[v]) { cn[v] = -1; av_index[v] = nav++; av_node.push_back(v); }
        else cn[v] = nc++;
    }
    long long sum_abs = 0;
    for (int a = 0; a < m; a++) sum_abs += std::abs((long long)rew[a]);
    if (nc > 1000 || sum_abs >= (1 << 18) || nav > 32000) {
        err = "packing limit: contracted nodes <= 1000, sum |reward| < 2^18 (DESIGN.md §5)";
        return SGUFP_ERR_LIMITS;
    }
    return 0;
}

int build_plan(const Model &M, const int16_t *path, int plen, Plan &P, std::string &err, bool lane_tables) {
    const int m = M.m;
    if (!path || plen < 0 || plen > M.L) { err = "path must have at most L entries"; return SGUFP_ERR_ARG; }
    P.match_out.assign(m, -1); P.match_in.assign(m, -1);
    for (int ell = 0; ell < plen; ell++) {
        int b = path[ell];
        if (b == -1) continue;
        if (b < 0 || b >= m) { err = "decision is not an arc id (the reference would read out of bounds, grb.cpp:147)"; return SGUFP_ERR_ARG; }
        int a = M.layer_arc[ell], q = M.head[a];
        if (!M.active[q]) continue;
        // y-bar is keyed by node ids (grb.cpp:145-148): "the out-arc of q whose head is head(path[l])"
        int j = M.head[b], hit = -1;
        for (int e = M.out_ptr[q]; e < M.out_ptr[q + 1]; e++)
            if (M.head[M.out_arc[e]] == j) { hit = M.out_arc[e]; break; }
        if (hit < 0) continue;  // an entry of y-bar no row ever reads
        if (P.match_in[hit] != -1) { err = "two in-arcs matched to one out-arc"; return SGUFP_ERR_MATCHING; }
        P.match_out[a] = hit; P.match_in[hit] = a;
    }
    // chains: maximal runs a1 -> a2 -> ... glued through matched pairs
    struct Ch { int first, sv, ev, qs, qe, r, len; };
    std::vector<Ch> open_ch, closed_ch;
    std::vector<int32_t> nxt(m, -1);
    int covered = 0;
    for (int a0 = 0; a0 < m; a0++) {
        if (M.active[M.tail[a0]] && P.match_in[a0] != -1) continue;
        Ch c{a0, 0, 0, -1, -1, 0, 0};
        if (M.active[M.tail[a0]]) { c.sv = -1; c.qs = M.av_index[M.tail[a0]]; } else c.sv = M.cn[M.tail[a0]];
        int a = a0;
        for (;;) {
            c.r += M.rew[a]; c.len++; covered++;
            int b = M.active[M.head[a]] ? P.match_out[a] : -1;
            if (b < 0) break;
            nxt[a] = b; a = b;
        }
        if (M.active[M.head[a]]) { c.ev = -1; c.qe = M.av_index[M.head[a]]; } else c.ev = M.cn[M.head[a]];
        if (c.len > 1023) { err = "chain longer than 1023 arcs"; return SGUFP_ERR_LIMITS; }
        (c.sv >= 0 && c.ev >= 0 ? open_ch : closed_ch).push_back(c);
    }
    if (covered != m) { err = "matched pairs form a cycle"; return SGUFP_ERR_CYCLIC; }
    {   // open chains in topological order of their tail (Gauss-Seidel label correction converges in few passes)
        std::vector<int32_t> depth(M.nc, 0);
        for (int it = 0, changed = 1; changed && it <= M.nc; it++) {
            changed = 0;
            for (const Ch &c : open_ch)
                if (c.ev != 0 && depth[c.ev] < depth[c.sv] + 1) { depth[c.ev] = depth[c.sv] + 1; changed = 1; }
        }
        std::stable_sort(open_ch.begin(), open_ch.end(), [&](const Ch &a, const Ch &b) { return depth[a.sv] < depth[b.sv]; });
    }
    const int nopen = (int)open_ch.size(), nch = nopen + (int)closed_ch.size();
    P.nch = nch; P.nopen = nopen;
    std::vector<Ch> chains(open_ch);
    chains.insert(chains.end(), closed_ch.begin(), closed_ch.end());

    std::vector<int32_t> arc_cp(m), arc_info(m), arc_pre(m), ch_ends(nch), ch_r(nch), ch_ptr(nch + 1), ch_arcs(m), ch_q(nch);
    std::vector<int32_t> arc_chain(m);
    int fill = 0;
    for (int c = 0; c < nch; c++) {
        const Ch &ch = chains[c];
        ch_ptr[c] = fill;
        ch_ends[c] = (ch.sv + 1) | ((ch.ev + 1) << 16);
        ch_q[c] = (ch.qs + 1) | ((ch.qe + 1) << 16);
        ch_r[c] = ch.r;
        int pos = 0, pre = 0;
        const bool is_open = c < nopen;
        for (int a = ch.first; a >= 0; a = nxt[a], pos++) {
            pre += M.rew[a];
            ch_arcs[fill++] = a; arc_chain[a] = c;
            arc_cp[a] = (c << 10) | pos;
            arc_pre[a] = pre;
            int kind;
            if (is_open) kind = M.active[M.head[a]] ? KIND_SIGMA : (M.active[M.tail[a]] ? KIND_PHI : KIND_GAMMA);
            else {
                // closed chain: the only multipliers are the FREE ones of its dangling end arcs
                const bool last = nxt[a] < 0, first = a == ch.first;
                if (last && ch.ev < 0) kind = KIND_SIGMA;         // unmatched in-arc of a V-bar node
                else if (first && ch.sv < 0) kind = KIND_PHI;     // unmatched out-arc of a V-bar node
                else kind = KIND_GAMMA;                           // interior arc of a closed chain: multiplier is always 0
            }
            arc_info[a] = kind | ((M.arc_layer[a] + 1) << 2);
        }
    }
    ch_ptr[nch] = fill;
    std::vector<int32_t> av_ptr(M.nav + 1, 0), av_arcs, fb_ptr(M.nav + 1, 0), fb_ch;
    for (int i = 0; i < M.nav; i++) {
        int q = M.av_node[i];
        av_ptr[i] = (int)av_arcs.size(); fb_ptr[i] = (int)fb_ch.size();
        for (int e = M.in_ptr[q]; e < M.in_ptr[q + 1]; e++)
            if (P.match_out[M.in_arc[e]] >= 0) av_arcs.push_back(M.in_arc[e]);
        if ((int)av_arcs.size() == av_ptr[i])
            for (int e = M.out_ptr[q]; e < M.out_ptr[q + 1]; e++) {
                int c = arc_chain[M.out_arc[e]];
                if (chains[c].ev >= 0) fb_ch.push_back(c);
            }
    }
    av_ptr[M.nav] = (int)av_arcs.size(); fb_ptr[M.nav] = (int)fb_ch.size();

    // residual slots of the contracted graph, sorted by head label index (lane kernel: backward searches from the sink)
    struct Slot { int tail, head, cost, cd; };
    std::vector<Slot> slots;
    slots.reserve(lane_tables ? 2 * nopen : 0);
    for (int c = 0; lane_tables && c < nopen; c++) {
        const Ch &ch = chains[c];
        slots.push_back({ch.sv, ch.ev == 0 ? M.nc : ch.ev, -ch.r, 2 * c});
        slots.push_back({ch.ev, ch.sv == 0 ? M.nc : ch.sv, ch.r, 2 * c + 1});
    }
    std::sort(slots.begin(), slots.end(), [](const Slot &a, const Slot &b) {
        if (a.head != b.head) return a.head < b.head;
        if (a.tail != b.tail) return a.tail < b.tail;
        return a.cd < b.cd;
    });
    PlanHeader H{};
    H.nch = nch; H.nopen = nopen; H.nc = M.nc; H.nav = M.nav; H.m = m; H.L = M.L;
    std::vector<int32_t> &W = P.words;
    W.clear();
    W.reserve(sizeof(PlanHeader) / 4 + 4 * (size_t)m + 10 * (size_t)nch + 4 * (size_t)M.nav + 3 * (size_t)M.nc + (lane_tables ? 6 * (size_t)nopen + 8 * (size_t)M.nc : 0) + 64);
    W.assign(sizeof(PlanHeader) / 4, 0);
    auto put = [&](const std::vector<int32_t> &v) { int32_t off = (int32_t)W.size(); W.insert(W.end(), v.begin(), v.end()); if (W.size() & 1) W.push_back(0); return off; };
    H.o_arc_cp = put(arc_cp); H.o_arc_info = put(arc_info); H.o_arc_pre = put(arc_pre);
    {
        std::vector<int32_t> arc_av(2 * (size_t)m, -1);
        for (int a = 0; a < m; a++) if (M.active[M.head[a]] && P.match_out[a] >= 0) { arc_av[2 * a] = M.av_index[M.head[a]]; arc_av[2 * a + 1] = P.match_out[a]; }
        H.o_arc_av = put(arc_av);   // W.size() is even before every put: 8-byte aligned
    }
    H.o_ch_ends = put(ch_ends); H.o_ch_r = put(ch_r); H.o_ch_ptr = put(ch_ptr); H.o_ch_arcs = put(ch_arcs); H.o_ch_q = put(ch_q);
    H.o_av_ptr = put(av_ptr); H.o_av_arcs = put(av_arcs); H.o_fb_ptr = put(fb_ptr); H.o_fb_ch = put(fb_ch);
    {
        std::vector<int32_t> st(2 * nopen);
        for (int c = 0; c < nopen; c++) {
            const Ch &ch = chains[c];
            const int hf = ch.ev == 0 ? M.nc : ch.ev, hb = ch.sv == 0 ? M.nc : ch.sv;
            st[2 * c] = ch.sv | (ch.ev << 10) | (hf << 20);
            st[2 * c + 1] = hb | (int32_t)((uint32_t)ch.r << 10);
        }
        H.o_ch_st = put(st);   // W.size() is even here (header even, every array padded to even)
        std::vector<int32_t> pk, in_pd(M.nc + 2, 0);
        size_t s = 0;
        int max_indeg = 0;
        for (int v = 0; v <= M.nc; v++) {
            const int first = (int)(pk.size() / 2);
            for (; s < slots.size() && slots[s].head == v; s++) {
                pk.push_back(slots[s].tail | (slots[s].cd << 16));
                pk.push_back((int32_t)((uint32_t)slots[s].cost << 12) | slots[s].head);
            }
            const int deg = (int)(pk.size() / 2) - first;
            in_pd[v] = first | (deg << 16);
            max_indeg = std::max(max_indeg, deg);
        }
        H.o_slots = put(pk);           // W.size() is even here: 8-byte aligned records
        H.o_in_pd = put(in_pd);
        H.max_indeg = max_indeg;
        P.max_indeg = max_indeg;
    }
    while (W.size() & 3) W.push_back(0);  // keep every plan 16-byte aligned inside a batch
    H.total = (int32_t)W.size();
    std::memcpy(W.data(), &H, sizeof(H));
    return 0;
}

void link_plans(const Plan &prev, const Plan &cur, std::vector<int32_t> &out) {
    out.clear();
    const int32_t *A = prev.words.data(), *B = cur.words.data();
    const PlanHeader *ha = reinterpret_cast<const PlanHeader *>(A), *hb = reinterpret_cast<const PlanHeader *>(B);
    if (ha->m != hb->m || ha->nc != hb->nc) return;
    const int na = ha->nopen, nb = hb->nopen;
    const int32_t *cp_a = A + ha->o_arc_cp, *ptr_a = A + ha->o_ch_ptr, *arcs_a = A + ha->o_ch_arcs;
    const int32_t *ptr_b = B + hb->o_ch_ptr, *arcs_b = B + hb->o_ch_arcs;
    std::vector<int32_t> prev_of(nb, -1);
    std::vector<uint8_t> kept(na, 0);
    int n_new = 0;
    for (int c = 0; c < nb; c++) {
        const int b0 = ptr_b[c], len = ptr_b[c + 1] - b0, cpa = cp_a[arcs_b[b0]], ca = cpa >> 10;
        if ((cpa & 1023) == 0 && ca < na && ptr_a[ca + 1] - ptr_a[ca] == len &&
            std::memcmp(arcs_a + ptr_a[ca], arcs_b + b0, (size_t)len * sizeof(int32_t)) == 0) { prev_of[c] = ca; kept[ca] = 1; }
        else n_new++;
    }
    std::vector<int32_t> removed;
    for (int c = 0; c < na; c++) if (!kept[c]) removed.push_back(c);
    if (4 * (n_new + (int)removed.size()) > std::max(na, nb)) return;
    out.reserve(2 + nb + 2 * removed.size() + 3);
    out.push_back((int32_t)removed.size()); out.push_back(na);
    out.insert(out.end(), prev_of.begin(), prev_of.end());
    out.insert(out.end(), removed.begin(), removed.end());
    const int32_t *st_a = A + ha->o_ch_st;
    for (int c : removed) out.push_back(st_a[2 * c] & 0xfffff);        // sv | ev << 10 of the removed chain (model.hpp: o_ch_st)
    while (out.size() & 3) out.push_back(0);
}

// layers in which two paths differ (the hot loop of order_batch: K^2 / 2 calls per batch, 64 candidates of C2 in ~15 us)
static inline int path_distance(const int16_t *a, const int16_t *b, int L) {
    int t = 0, same = 0;
#if defined(__SSE2__)
    __m128i acc = _mm_setzero_si128();                 // eight 16-bit counters of EQUAL lanes (a compare gives -1 per equal lane)
    for (; t + 8 <= L; t += 8)
        acc = _mm_sub_epi16(acc, _mm_cmpeq_epi16(_mm_loadu_si128(reinterpret_cast<const __m128i *>(a + t)), _mm_loadu_si128(reinterpret_cast<const __m128i *>(b + t))));
    alignas(16) int16_t lanes[8];
    _mm_store_si128(reinterpret_cast<__m128i *>(lanes), acc);
    for (int i = 0; i < 8; i++) same += lanes[i];
#endif
    for (; t < L; t++) same += a[t] == b[t];
    return L - same;
}

void order_batch(const int16_t *paths, int K, int L, const int16_t *start, std::vector<int32_t> &order) {
    order.resize(K);
    for (int k = 0; k < K; k++) order[k] = k;
    if (K > 128 || L <= 0 || K < 2 || (K == 2 && !start)) return;
    std::vector<char> used(K, 0);
    const int16_t *cur = start;
    int j = 0;
    if (!cur) { used[0] = 1; cur = paths; j = 1; }
    for (; j < K; j++) {
        int best = -1, bd = INT_MAX;
        for (int i = 0; i < K; i++) {
            if (used[i]) continue;
            const int d = path_distance(paths + (size_t)i * L, cur, L);
            if (d < bd) { bd = d; best = i; }          // ties: the order given
        }
        order[j] = best; used[best] = 1; cur = paths + (size_t)best * L;
    }
}

// Split-graph arrays of the feasibility-ray kernel (k1_ray).  A dangling end arc books into its
// FREE multiplier, every other arc by its endpoints (rule 6).
void ray_arrays(const Model &M, const Plan &P, std::vector<int32_t> &ts, std::vector<int32_t> &hs, std::vector<int32_t> &info,
                std::vector<int32_t> &pair_layer, std::vector<int32_t> &next, std::vector<int32_t> &aq, std::vector<int32_t> &first_wire, int &nn) {
    const int m = M.m;
    std::vector<int32_t> wire(m, -1);
    nn = M.nc;
    for (int a = 0; a < m; a++) if (P.match_out[a] >= 0) wire[a] = nn++;
    ts.assign(m, 0); hs.assign(m, 0); info.assign(m, 0); pair_layer.assign(m, -1); next.assign(m, -1); aq.assign(m, 0);
    first_wire.assign(std::max(1, M.nav), -1);
    for (int a = 0; a < m; a++) {
        const int t = M.tail[a], h = M.head[a];
        ts[a] = M.active[t] ? (P.match_in[a] >= 0 ? wire[P.match_in[a]] : -1) : M.cn[t];
        hs[a] = M.active[h] ? (P.match_out[a] >= 0 ? wire[a] : -1) : M.cn[h];
        int kind;
        if (hs[a] < 0) kind = KIND_SIGMA; else if (ts[a] < 0) kind = KIND_PHI;
        else kind = M.active[h] ? KIND_SIGMA : (M.active[t] ? KIND_PHI : KIND_GAMMA);
        info[a] = kind | ((M.arc_layer[a] + 1) << 2);
        if (P.match_out[a] >= 0) { pair_layer[a] = M.arc_layer[a]; next[a] = P.match_out[a]; }
        aq[a] = (M.av_index[t] + 1) | ((M.av_index[h] + 1) << 16);
    }
    for (int i = 0; i < M.nav; i++) {
        const int q = M.av_node[i];
        for (int e = M.in_ptr[q]; e < M.in_ptr[q + 1]; e++)
            if (P.match_out[M.in_arc[e]] >= 0) { first_wire[i] = wire[M.in_arc[e]]; break; }
    }
}

uint64_t cut_hash(const uint64_t *keys, const double *vals, int nnz) {
    uint64_t h = 0;  // sum over i of ((key * i) xor hash(value)), Cut.h:247-251
    for (int i = 0; i < nnz; i++) h += ((keys[i] * (uint64_t)i) ^ (uint64_t)std::hash<double>{}(vals[i]));
    return h;
}

}  // namespace sgufp
