// k2_dd.cu — K2: layer-wise longest path of a decision diagram with cut-adjusted arc weights (sm_100a).
//
// The hot loop of RelaxedDDNew::applyOptimalityCut / applyFeasibilityCut
// (/root/reference/DD.cpp:3951-3973, 3860-3882) and of the RestrictedDDNew versions
// (DD.cpp:3454-3470, 3376-3399), batched over B diagrams x C cuts:
//   state(root) = RHS + sum of the coefficients of the fixed prefix decisions   (DD.cpp:3938-3949)
//   state(v)    = max over in-arcs in stored order of state(tail) + w(arc),  w = coef[slot] or 0 for decision -1
// Only fp64 add / compare: results are bit-identical to the reference's (max is exact).
//
// Mapping: one CTA per (cut, diagram); a layer is a barrier-separated step.  The node states never
// leave the SM: two shared-memory buffers ping-pong (a layer only reads the previous one) and the
// cut's dense coefficient vector is staged in shared memory once per CTA, so the only global
// traffic per arc is ONE coalesced 8-byte {tail position, slot} record (4 more for in_ptr where a
// layer is not a tree layer).  A single-node (collapsed) layer is reduced by the whole CTA.
// Diagrams wider than the shared-memory budget fall back to a global-state variant.
#include "k2_dd.cuh"

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cstdlib>

namespace sgufp {

// The dynamic shared-memory limit of a kernel is process-wide state: it is always raised to the same value (the most an SM
// offers), never to "what this launch needs" — host threads with diagrams of different widths would lower it under each other.
template <class F>
static cudaError_t k2_raise_smem_limit(F *kernel) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - (int)fa.sharedSizeBytes);   // static + dynamic <= 227 KB
}

namespace {

constexpr int K2_THREADS = 256;
constexpr int K2_HYBRID = 0;           // global-state variant of the batch path: layers up to this width would keep their states in shared memory.
                                       // Measured with 2048: 2.67e11 -> 2.25e11 arcs/s on the relaxed diagrams (32 KB more per CTA: 18 -> 6 CTAs per SM on a
                                       // latency-bound kernel; the wide last layers hold most of the nodes anyway): off.  SGUFP_K2_HYBRID=<width> turns it on.
constexpr int K2_TERM_SLICES = 64;     // CTAs per diagram in k2_terminal (scratch: B x 64 partial maxima)
constexpr int K2_FIN_THREADS = 1024;   // the one-CTA kernels that sweep a (possibly wide) last layer

__device__ __forceinline__ double block_max(double v, double *red) {
    for (int o = 16; o; o >>= 1) { const double t = __shfl_xor_sync(0xffffffffu, v, o); v = v < t ? t : v; }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        double x = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : -DBL_MAX;
        for (int o = 16; o; o >>= 1) { const double t = __shfl_xor_sync(0xffffffffu, x, o); x = x < t ? t : x; }
        if (threadIdx.x == 0) red[0] = x;
    }
    __syncthreads();
    const double r = red[0];
    __syncthreads();
    return r;
}

// SMEM_STATES: states ping-pong in shared memory (every layer is at most maxw wide); otherwise they live in the
// global scratch block of this (diagram, cut) — `gstate` then has C blocks per diagram — except, in the batch path
// (!DEADS), the layers of at most maxw nodes: those still ping-pong in two shared buffers of maxw states (a relaxed diagram
// is dozens of narrow layers and a few very wide ones at the end: only the wide ones touch global memory).
// DEADS: the diagram carries flags of arcs removed on the device (single-cut path); a flagged arc
// contributes nothing, exactly as if it had been erased from the in-arc list.
template <bool SMEM_STATES, bool DEADS>
__global__ void __launch_bounds__(K2_FIN_THREADS) k2_longest_path(const K2DD *__restrict__ dds, const double *__restrict__ coef,
                                                               const double *__restrict__ rhs, int C, int Tpad, int maxw,
                                                               double *__restrict__ gstate, double *__restrict__ glast, int li_cache) {
    extern __shared__ double sm[];
    __shared__ double red[K2_FIN_THREADS / 32];
    const K2DD d = dds[blockIdx.y];
    const int c = blockIdx.x;
    const int sw = (SMEM_STATES || !DEADS) ? maxw : 0;   // layers up to this width live in the two shared buffers
    double *cf = sm;                       // [Tpad]
    double *buf[2] = {sm + Tpad, sm + Tpad + sw};
    // the per-layer records, staged once: a layer is short, a dependent global load per layer is not
    int4 *li_s = reinterpret_cast<int4 *>(sm + Tpad + 2 * sw);
    const bool li_cached = d.nlayers <= li_cache;
    if (li_cached) for (int i = threadIdx.x; i < d.nlayers; i += blockDim.x) li_s[i] = d.layer_info[i];
    const int4 *layer_info = li_cached ? li_s : d.layer_info;
    for (int i = threadIdx.x; i < Tpad; i += blockDim.x) cf[i] = coef[(size_t)c * Tpad + i];
    const bool keep_all = c == C - 1;      // the host reads every node state of the last cut only
    double *all = gstate + d.state_off + (SMEM_STATES ? 0 : (size_t)c * d.nnodes);
    const bool root_s = SMEM_STATES || sw >= 1;
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = rhs[c];
        for (int k = 0; k < d.nroot; k++) { const int s = d.root_slot[k]; if (s >= 0) v = v + cf[s]; }
        if (root_s) buf[0][0] = v;
        if (!root_s || keep_all) all[0] = v;
        if (d.nlayers == 1) glast[d.last_off + (size_t)c * d.nlast] = v;
    }
    __syncthreads();
    const double *prev = root_s ? buf[0] : all;
    int which = 0;                         // the shared buffer written last
    for (int l = 1; l < d.nlayers; l++) {
        const int4 li = layer_info[l];
        const int v0 = li.x, e0 = li.y, width = li.z;
        const bool cur_s = SMEM_STATES || width <= sw;
        double *cur = cur_s ? buf[which ^ 1] : all + v0;
        const bool is_last = l == d.nlayers - 1;
        double *lastp = glast + d.last_off + (size_t)c * d.nlast;
        // the last layer's states go to `lastp`; nobody reads them from the state block unless the host wants every state
        // (the global-state variant of a relaxed diagram would write its ~10 k-node last layer twice per cut)
        const bool store = cur_s || !is_last || keep_all || DEADS;      // DEADS: the sequence path reads every cut's states afterwards
        const bool mirror = cur_s && keep_all;                           // shared-memory states the host will read
        if (width == 1 && !li.w) {
            // collapsed layer: one node, many in-arcs -> the whole CTA reduces it
            const int e1 = d.in_ptr[v0 + 1];
            double s = -DBL_MAX;
            for (int e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
                if (DEADS && d.arc_dead[e]) continue;
                const int2 ts = d.arc_ts[e];
                const double p = prev[ts.x];
                const double cand = ts.y >= 0 ? p + cf[ts.y] : p;
                s = s < cand ? cand : s;
            }
            s = block_max(s, red);
            if (threadIdx.x == 0) {
                if (store) cur[0] = s;
                if (mirror) all[v0] = s;
                if (is_last) lastp[0] = s;
            }
        } else if (li.w) {
            // tree layer: node i's only in-arc is arc e0 + i
            for (int i = threadIdx.x; i < width; i += blockDim.x) {
                const int2 ts = d.arc_ts[e0 + i];
                const double p = prev[ts.x];
                double s = ts.y >= 0 ? p + cf[ts.y] : p;
                s = -DBL_MAX < s ? s : -DBL_MAX;   // max(DOUBLE_MIN, .) of DD.cpp:3958
                if (DEADS && d.arc_dead[e0 + i]) s = -DBL_MAX;
                if (store) cur[i] = s;
                if (mirror) all[v0 + i] = s;
                if (is_last) lastp[i] = s;
            }
        } else {
            for (int i = threadIdx.x; i < width; i += blockDim.x) {
                double s = -DBL_MAX;
                for (int e = d.in_ptr[v0 + i]; e < d.in_ptr[v0 + i + 1]; e++) {
                    if (DEADS && d.arc_dead[e]) continue;
                    const int2 ts = d.arc_ts[e];
                    const double p = prev[ts.x];
                    const double cand = ts.y >= 0 ? p + cf[ts.y] : p;
                    s = s < cand ? cand : s;
                }
                if (store) cur[i] = s;
                if (mirror) all[v0 + i] = s;
                if (is_last) lastp[i] = s;
            }
        }
        prev = cur;
        if (cur_s) which ^= 1;
        __syncthreads();
    }
}

// term[i] = min(term[i], min_c last[b][c][i]); partial maxima per slice of the last layer (a wide last layer — the
// relaxed diagram ends with ~10 k nodes — is spread over several CTAs), reduced by k2_terminal_bound
__global__ void __launch_bounds__(K2_THREADS) k2_terminal(const K2DD *__restrict__ dds, int C, const double *__restrict__ glast,
                                                           double *__restrict__ partial, int slices, double *__restrict__ bound) {
    __shared__ double red[K2_THREADS / 32];
    const K2DD d = dds[blockIdx.x];
    double best = -DBL_MAX;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < d.nlast; i += slices * blockDim.x) {
        double t = d.term[i];
        for (int c = 0; c < C; c++) {
            const double s = glast[d.last_off + (size_t)c * d.nlast + i];
            t = s < t ? s : t;                 // arc.weight = min(arc.weight, parent.state2) (DD.cpp:3981)
        }
        d.term[i] = t;
        best = best < t ? t : best;            // terminalState = max(terminalState, arc.weight)
    }
    best = block_max(best, red);
    if (threadIdx.x == 0) { if (slices == 1) bound[blockIdx.x] = best; else partial[(size_t)blockIdx.x * slices + blockIdx.y] = best; }
}

__global__ void k2_terminal_bound(const double *__restrict__ partial, int slices, double *__restrict__ bound) {
    double best = -DBL_MAX;
    for (int i = threadIdx.x; i < slices; i += 32) { const double t = partial[(size_t)blockIdx.x * slices + i]; best = best < t ? t : best; }
    for (int o = 16; o; o >>= 1) { const double t = __shfl_xor_sync(0xffffffffu, best, o); best = best < t ? t : best; }
    if (threadIdx.x == 0) bound[blockIdx.x] = best;
}

// ---- single cut on a WIDE diagram: one launch per layer, the layer spread over the whole GPU ------
// (one CTA per diagram would leave 147 SMs idle: the one-cut-at-a-time calls have no batch to fill
// the grid with.)  States live in the diagram's global state block; arcs flagged dead are skipped.
__global__ void __launch_bounds__(K2_THREADS) k2_layer(K2DD d, const double *__restrict__ coef, int l, double *__restrict__ state) {
    const int4 li = d.layer_info[l];
    const int v0 = li.x, e0 = li.y, width = li.z, prev0 = d.layer_info[l - 1].x;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= width) return;
    double s = -DBL_MAX;
    if (li.w) {
        const int2 ts = d.arc_ts[e0 + i];
        if (!d.arc_dead[e0 + i]) {
            const double p = state[prev0 + ts.x];
            s = ts.y >= 0 ? p + coef[ts.y] : p;
            s = -DBL_MAX < s ? s : -DBL_MAX;
        }
    } else {
        for (int e = d.in_ptr[v0 + i]; e < d.in_ptr[v0 + i + 1]; e++) {
            if (d.arc_dead[e]) continue;
            const int2 ts = d.arc_ts[e];
            const double p = state[prev0 + ts.x];
            const double cand = ts.y >= 0 ? p + coef[ts.y] : p;
            s = s < cand ? cand : s;
        }
    }
    state[v0 + i] = s;
}

// a collapsed layer (one node, many in-arcs) of the layered path: one CTA reduces it
__global__ void __launch_bounds__(K2_THREADS) k2_layer_collapsed(K2DD d, const double *__restrict__ coef, int l, double *__restrict__ state) {
    __shared__ double red[K2_THREADS / 32];
    const int4 li = d.layer_info[l];
    const int v0 = li.x, prev0 = d.layer_info[l - 1].x;
    double s = -DBL_MAX;
    for (int e = li.y + threadIdx.x; e < d.in_ptr[v0 + 1]; e += blockDim.x) {
        if (d.arc_dead[e]) continue;
        const int2 ts = d.arc_ts[e];
        const double p = state[prev0 + ts.x];
        const double cand = ts.y >= 0 ? p + coef[ts.y] : p;
        s = s < cand ? cand : s;
    }
    s = block_max(s, red);
    if (threadIdx.x == 0) state[v0] = s;
}

// ---- device-side cut application -----------------------------------------------------------------
__device__ int block_sum(int v, int *red) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        int x = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0;
        for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (threadIdx.x == 0) red[0] = x;
    }
    __syncthreads();
    const int r = red[0];
    __syncthreads();
    return r;
}

// Bound-based pruning of the arcs that enter a collapsed (single-node) layer of a non-exact
// diagram (DD.cpp:3995-4017 with threshold optimal - 0.01, layers 3 .. llayer-2; DD.cpp:3902-3926
// with threshold -0.01, layers 1 .. llayer-1).  Candidates are flagged 2 and committed only if no
// layer loses every arc; returns true if one did (the reference then returns without removing).
constexpr int K2_MAX_COLLAPSED = 64;   // collapsed layers handled per call by the fast path of prune_collapsed

// probe != nullptr: nothing is written to the diagram; *probe = 1 if the call WOULD flag an arc or abort.
__device__ bool prune_collapsed(const K2Apply &a, int first_layer, int end_layer, double max_state, double threshold,
                                int *ired, int *sh, int *probe = nullptr) {
    const K2DD &d = a.d;
    // the collapsed layers (ONE live node) in range, found with one parallel sweep over the live layer sizes
    // — a loop over the layers would pay one dependent global load each.  Their order does not matter: any layer
    // that loses every arc aborts the whole call, otherwise all candidates are committed.
    __shared__ int s_n, s_layer[K2_MAX_COLLAPSED];
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    for (int layer = first_layer + threadIdx.x; layer < end_layer; layer += blockDim.x)
        if (a.layer_alive[layer] == 1) { const int k = atomicAdd(&s_n, 1); if (k < K2_MAX_COLLAPSED) s_layer[k] = layer; }
    __syncthreads();
    const int ncol = s_n;
    const bool listed = ncol <= K2_MAX_COLLAPSED;
    __shared__ int s_abort, s_any;
    if (threadIdx.x == 0) { s_abort = 0; s_any = 0; }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    // one WARP per collapsed layer (they are independent): no block barrier, eight layers in flight
    auto one_layer = [&](int layer, bool commit_only, bool keep) -> int {
        const int4 li = d.layer_info[layer];
        int v = -1;
        for (int i = lane; i < li.z; i += 32) if (!a.node_dead[li.x + i]) v = li.x + i;      // the one node left
        for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
        const int e0 = d.in_ptr[v], e1 = d.in_ptr[v + 1];
        if (commit_only) {
            int changed = 0;
            for (int e = e0 + lane; e < e1; e += 32)
                if (a.arc_dead[e] == 2) { a.arc_dead[e] = keep ? 1 : 0; changed |= keep ? 1 : 0; }
            return changed;
        }
        const double gain = max_state - a.state[v];
        const int tail0 = d.layer_info[layer - 1].x;
        int pruned = 0, total = 0;
        for (int e = e0 + lane; e < e1; e += 32) {
            if (a.arc_dead[e]) continue;
            const int2 ts = d.arc_ts[e];
            const double w = ts.y >= 0 ? a.coef[ts.y] : 0.0;
            total++;
            if ((a.state[tail0 + ts.x] + w + gain) <= threshold) { if (!probe) a.arc_dead[e] = 2; pruned++; }
        }
        for (int o = 16; o; o >>= 1) { pruned += __shfl_xor_sync(0xffffffffu, pruned, o); total += __shfl_xor_sync(0xffffffffu, total, o); }
        if (probe && pruned && lane == 0) atomicOr(&s_any, 1);
        return pruned == total ? 1 : 0;
    };
    if (listed) {
        for (int k = warp; k < ncol; k += nwarp) if (one_layer(s_layer[k], false, false) && lane == 0) atomicOr(&s_abort, 1);
    } else {
        for (int layer = first_layer + warp; layer < end_layer; layer += nwarp) {
            if (a.layer_alive[layer] != 1) continue;
            if (one_layer(layer, false, false) && lane == 0) atomicOr(&s_abort, 1);
        }
    }
    __syncthreads();
    const bool abort = s_abort != 0;
    if (probe) { if (threadIdx.x == 0) *probe = (s_any || s_abort) ? 1 : 0; __syncthreads(); return abort; }
    // commit or discard the candidates: they all sit among the in-arcs of the collapsed nodes
    int changed = 0;
    if (listed) {
        for (int k = warp; k < ncol; k += nwarp) changed |= one_layer(s_layer[k], true, !abort);
    } else {
        const int narcs = d.in_ptr[d.nnodes];
        for (int e = threadIdx.x; e < narcs; e += blockDim.x)
            if (a.arc_dead[e] == 2) { a.arc_dead[e] = abort ? 0 : 1; changed |= abort ? 0 : 1; }
    }
    changed = block_sum(changed, ired);
    if (threadIdx.x == 0 && changed) a.out->changed = 1;
    __syncthreads();
    return abort;
}

// quiet: a probe (k2_prune_probe) found that this cut flags nothing: the pruning pass is skipped
__device__ void finish_body(const K2Apply &a, double *red, int *ired, int *sh, bool quiet = false) {
    const K2DD &d = a.d;
    const int nl = d.nlayers, llayer = nl - 1;
    const int last0 = d.layer_info[llayer].x, nlast = d.nlast;
    if (threadIdx.x == 0) { a.out->changed = 0; a.out->feasible = 1; a.out->bound = -DBL_MAX; }
    __syncthreads();
    if (a.mode == 0) {
        // terminal arcs: weight = min(weight, parent state); bound = max over them (DD.cpp:3975-3984, 3495-3504)
        double best = -DBL_MAX, ms = -DBL_MAX;
        for (int i = threadIdx.x; i < nlast; i += blockDim.x) {
            if (a.node_dead[last0 + i]) continue;
            const double s = a.state[last0 + i];
            double t = d.term[i];
            t = s < t ? s : t;
            d.term[i] = t;
            best = best < t ? t : best;
            ms = ms < s ? s : ms;
        }
        const double terminal = block_max(best, red);
        const double max_state = block_max(ms, red);
        if (a.restricted || terminal <= a.optimal || a.exact || quiet) { if (threadIdx.x == 0) a.out->bound = terminal; return; }
        const bool abort = prune_collapsed(a, 3, llayer - 1, max_state, a.optimal - 0.01, ired, sh);   // layers 3 .. llayer-2
        if (threadIdx.x == 0) a.out->bound = abort ? -DBL_MAX : terminal;
        return;
    }
    // ---- feasibility ----
    const double thr = a.restricted ? -0.5 : -0.01;            // DD.cpp:3398 / 3887
    int alive = 0, drop = 0;
    for (int i = threadIdx.x; i < nlast; i += blockDim.x) {
        if (a.node_dead[last0 + i]) continue;
        alive++;
        if ((nl >= 2 || !a.restricted) && a.state[last0 + i] < thr) drop++;   // the restricted tree guards on size() >= 2 (DD.cpp:3376)
    }
    alive = block_sum(alive, ired);
    drop = block_sum(drop, ired);
    if (a.restricted) {
        if (alive == 0) { if (threadIdx.x == 0) a.out->feasible = 0; return; }
        if (drop) {
            for (int i = threadIdx.x; i < nlast; i += blockDim.x)
                if (!a.node_dead[last0 + i] && a.state[last0 + i] < thr) a.node_dead[last0 + i] = 1;
            if (threadIdx.x == 0) { a.out->changed = 1; a.layer_alive[llayer] = alive - drop; a.out->feasible = alive - drop > 0 ? 1 : 0; }
        }
        return;
    }
    if (drop == alive) { if (threadIdx.x == 0) a.out->feasible = 0; return; }   // nothing is removed then (DD.cpp:3891)
    if (drop) {
        // batchRemoveNodes (DD.cpp:4040-4160): the nodes go, their in-arcs go, and a parent that LOSES its
        // last out-arc this way goes too, layer by layer upwards.  node_dead == 2: dies in this sweep.
        for (int i = threadIdx.x; i < nlast; i += blockDim.x)
            if (!a.node_dead[last0 + i] && a.state[last0 + i] < thr) a.node_dead[last0 + i] = 2;
        if (threadIdx.x == 0) { a.out->changed = 1; a.layer_alive[llayer] = alive - drop; }
        __syncthreads();
        for (int l = llayer; l >= 1; l--) {
            const int4 li = d.layer_info[l], lp = d.layer_info[l - 1];
            // in-arcs of the nodes that die in layer l
            int any = 0;
            for (int i = threadIdx.x; i < li.z; i += blockDim.x) {
                const int v = li.x + i;
                if (a.node_dead[v] != 2) continue;
                any = 1;
                for (int e = d.in_ptr[v]; e < d.in_ptr[v + 1]; e++)
                    if (!a.arc_dead[e]) { a.arc_dead[e] = 1; a.lost[lp.x + d.arc_ts[e].x] = 1; }
                a.node_dead[v] = 1;
            }
            any = block_sum(any, ired);
            if (!any) break;
            // live out-arcs of the parents that lost one
            for (int i = threadIdx.x; i < lp.z; i += blockDim.x) a.cnt[lp.x + i] = 0;
            __syncthreads();
            const int e_end = d.in_ptr[li.x + li.z];
            for (int e = li.y + threadIdx.x; e < e_end; e += blockDim.x)
                if (!a.arc_dead[e]) { const int t = lp.x + d.arc_ts[e].x; if (a.lost[t]) atomicAdd(&a.cnt[t], 1); }
            __syncthreads();
            int died = 0;
            for (int i = threadIdx.x; i < lp.z; i += blockDim.x) {
                const int u = lp.x + i;
                if (a.lost[u]) {
                    a.lost[u] = 0;
                    if (!a.node_dead[u] && a.cnt[u] == 0) { a.node_dead[u] = 2; died++; }
                }
            }
            died = block_sum(died, ired);
            if (threadIdx.x == 0 && died) a.layer_alive[l - 1] -= died;
            __syncthreads();
        }
        // a root that died keeps flag 2 -> settle it
        if (threadIdx.x == 0 && a.node_dead[0] == 2) a.node_dead[0] = 1;
        __syncthreads();
    }
    if (!a.exact) {                                                               // DD.cpp:3895-3928
        double ms = -DBL_MAX;
        for (int i = threadIdx.x; i < nlast; i += blockDim.x)
            if (!a.node_dead[last0 + i]) { const double s = a.state[last0 + i]; ms = ms < s ? s : ms; }
        const double max_state = block_max(ms, red);
        if (prune_collapsed(a, 1, llayer, max_state, -0.01, ired, sh)) { if (threadIdx.x == 0) a.out->feasible = 0; }
    }
}

__global__ void __launch_bounds__(K2_FIN_THREADS) k2_finish(K2Apply a) {
    __shared__ double red[K2_FIN_THREADS / 32];
    __shared__ int ired[K2_FIN_THREADS / 32];
    __shared__ int sh[2];
    finish_body(a, red, ired, sh);
}

// A run of cuts of one kind on one diagram.  Their longest paths were computed side by side on the
// CURRENT structure (k2_longest_path, one CTA per cut); this kernel applies the sequential part cut
// by cut and stops after the first cut that (a) ends the loop of the caller — bound <= optimal
// (NodeExplorer.cpp:942, 982) or not feasible (:936, 977) — or (b) removed arcs or nodes: the states
// of the cuts behind it were computed on a structure that no longer exists and are recomputed.
// One CTA per cut of the window: would the optimality pruning of this cut touch the diagram?  (It cannot
// depend on the cuts before it except through the structure, which is the same for the whole window.)
__global__ void __launch_bounds__(K2_THREADS) k2_prune_probe(K2Apply a, K2Seq q) {
    __shared__ double red[K2_THREADS / 32];
    __shared__ int ired[K2_THREADS / 32];
    __shared__ int sh[2];
    const int k = q.k0 + blockIdx.x;
    K2Apply b = a;
    b.state = q.states + (size_t)blockIdx.x * a.d.nnodes;
    b.coef = q.coef + (size_t)k * q.Tpad;
    const K2DD &d = a.d;
    const int llayer = d.nlayers - 1, last0 = d.layer_info[llayer].x;
    double ms = -DBL_MAX;
    for (int i = threadIdx.x; i < d.nlast; i += blockDim.x)
        if (!a.node_dead[last0 + i]) { const double s = b.state[last0 + i]; ms = ms < s ? s : ms; }
    const double max_state = block_max(ms, red);
    prune_collapsed(b, 3, llayer - 1, max_state, a.optimal - 0.01, ired, sh, q.probe + blockIdx.x);
}

constexpr int K2_SEQ_STAGE = 2048;   // window lengths whose probe flags are staged in shared memory

// Terminal weights of a whole window, speculatively (valid up to the first cut that changes the structure):
// q.last[c][i] becomes min(term[i], min over cuts <= c of the state of last-layer node i) ...
__global__ void __launch_bounds__(K2_THREADS) k2_window_prefix_min(K2Apply a, K2Seq q) {
    const K2DD &d = a.d;
    const int i = blockIdx.x * blockDim.x + threadIdx.x, n = q.k1 - q.k0, nlast = d.nlast;
    if (i >= nlast || a.node_dead[d.layer_info[d.nlayers - 1].x + i]) return;
    double t = d.term[i];
    for (int c = 0; c < n; c++) {
        double *p = q.last + (size_t)c * nlast + i;
        const double s = *p;
        t = s < t ? s : t;                                   // arc.weight = min(arc.weight, parent.state2) (DD.cpp:3981)
        *p = t;
    }
}
// ... and q.bounds[c] = max over the live last-layer nodes: the bound the c-th call returns if nothing was removed before it.
__global__ void __launch_bounds__(K2_THREADS) k2_window_bounds(K2Apply a, K2Seq q) {
    __shared__ double red[K2_THREADS / 32];
    const K2DD &d = a.d;
    const int c = blockIdx.x, nlast = d.nlast, last0 = d.layer_info[d.nlayers - 1].x;
    double m = -DBL_MAX;
    for (int i = threadIdx.x; i < nlast; i += blockDim.x)
        if (!a.node_dead[last0 + i]) { const double t = q.last[(size_t)c * nlast + i]; m = m < t ? t : m; }
    m = block_max(m, red);
    if (threadIdx.x == 0) q.bounds[c] = m;
}

__global__ void __launch_bounds__(K2_FIN_THREADS) k2_finish_seq(K2Apply a, K2Seq q) {
    __shared__ double red[K2_FIN_THREADS / 32];
    __shared__ int ired[K2_FIN_THREADS / 32];
    __shared__ int sh[2];
    __shared__ unsigned char s_quiet[K2_SEQ_STAGE];
    __shared__ int s_stop;
    const K2DD &d = a.d;
    const int n = q.k1 - q.k0, nlast = d.nlast, last0 = d.layer_info[d.nlayers - 1].x;
    // A QUIET optimality cut changes nothing but the terminal weights: the diagram is exact or restricted (no
    // pruning there, DD.cpp:3985-3987, 3493-3504), or the probe found nothing to prune.  For a run of quiet cuts
    // the prefix minima and the bounds are already there (k2_window_prefix_min / k2_window_bounds): the run is
    // scanned for the first bound <= optimal, and the terminal weights are taken from its last cut.
    const bool all_quiet = a.mode == 0 && (a.exact || a.restricted);
    const bool staged = a.mode == 0 && n <= K2_SEQ_STAGE && q.bounds != nullptr && (all_quiet || q.probe != nullptr);
    if (staged) for (int i = threadIdx.x; i < n; i += blockDim.x) s_quiet[i] = all_quiet || q.probe[i] == 0;
    __syncthreads();
    int k = q.k0;
    while (k < q.k1) {
        if (staged && s_quiet[k - q.k0]) {
            int r = k;
            while (r < q.k1 && s_quiet[r - q.k0]) r++;
            if (threadIdx.x == 0) s_stop = INT_MAX;
            __syncthreads();
            for (int c = k + threadIdx.x; c < r; c += blockDim.x) {
                K2Result res; res.bound = q.bounds[c - q.k0]; res.feasible = 1; res.changed = 0; res.path_len = 0;
                q.results[c] = res;
                if (res.bound <= a.optimal) atomicMin(&s_stop, c);   // the caller's loop returns at the first such cut
            }
            __syncthreads();
            const int stop = s_stop, upto = stop == INT_MAX ? r - 1 : stop;
            for (int i = threadIdx.x; i < nlast; i += blockDim.x)
                if (!a.node_dead[last0 + i]) d.term[i] = q.last[(size_t)(upto - q.k0) * nlast + i];
            if (stop != INT_MAX) {
                if (threadIdx.x == 0) { q.ctl[0] = stop + 1; q.ctl[1] = 1; }
                return;
            }
            k = r;
            __syncthreads();
            continue;
        }
        // a cut that prunes (or a feasibility cut): the full sequential step on ITS OWN states; the terminal weights
        // it starts from were just written above (or by the previous step)
        K2Apply b = a;
        b.state = q.states + (size_t)(k - q.k0) * d.nnodes;
        b.coef = q.coef + (size_t)k * q.Tpad;
        b.out = q.results + k;
        finish_body(b, red, ired, sh, q.probe != nullptr && q.probe[k - q.k0] == 0);
        __syncthreads();
        const K2Result r = q.results[k];
        const bool stop = a.mode == 0 ? r.bound <= a.optimal : r.feasible == 0;
        if (r.changed || stop) {
            if (threadIdx.x == 0) { q.ctl[0] = k + 1; q.ctl[1] = stop ? 1 : 0; }
            return;
        }
        __syncthreads();
        // the speculative prefix minima behind this cut did not see it: the window ends here, the rest is recomputed
        if (staged) { if (threadIdx.x == 0) { q.ctl[0] = k + 1; q.ctl[1] = 0; } return; }
        k++;
    }
    if (threadIdx.x == 0) { q.ctl[0] = q.k1; q.ctl[1] = 0; }
}

// getSolution (DD.cpp:3825-3840) / getMaxPath (DD.cpp:3290-3305) + getPathForNode (DD.cpp:3796-3820):
// the first terminal arc of strictly greatest weight, then backwards over the first in-arc whose
// parent state + weight EQUALS the node's state (exact fp compare); if none does, step to the first
// in-arc's parent without recording a decision.  A restricted tree records its only in-arc always.
__global__ void __launch_bounds__(K2_THREADS) k2_extract(K2Apply a) {
    __shared__ double red[K2_THREADS / 32];
    __shared__ int sh[2];
    const K2DD &d = a.d;
    const int llayer = d.nlayers - 1, last0 = d.layer_info[llayer].x;
    // argmax with the lowest index among equals
    double best = -DBL_MAX; int bi = INT_MAX;
    for (int i = threadIdx.x; i < d.nlast; i += blockDim.x)
        if (!a.node_dead[last0 + i] && d.term[i] > best) { best = d.term[i]; bi = i; }
    const double gbest = block_max(best, red);
    if (threadIdx.x == 0) sh[0] = INT_MAX;
    __syncthreads();
    if (bi != INT_MAX && best == gbest && gbest > -DBL_MAX) atomicMin(&sh[0], bi);
    __syncthreads();
    int cur = sh[0] == INT_MAX ? 0 : last0 + sh[0];
    int l = sh[0] == INT_MAX ? 0 : llayer;
    int n = 0;                                   // decisions recorded so far (leaf -> root), kept by every thread
    __syncthreads();
    while (l > 0) {
        const int tail0 = d.layer_info[l - 1].x, e0 = d.in_ptr[cur], e1 = d.in_ptr[cur + 1];
        if (threadIdx.x == 0) { sh[0] = INT_MAX; sh[1] = INT_MAX; }
        __syncthreads();
        const double sc = a.state[cur];
        int first = INT_MAX, match = INT_MAX;
        for (int e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
            if (a.arc_dead[e]) continue;
            if (e < first) first = e;
            const int2 ts = d.arc_ts[e];
            const double w = ts.y >= 0 ? a.coef[ts.y] : 0.0;
            if (e < match && (a.restricted || (a.state[tail0 + ts.x] + w) == sc)) match = e;
        }
        if (first != INT_MAX) atomicMin(&sh[0], first);
        if (match != INT_MAX) atomicMin(&sh[1], match);
        __syncthreads();
        const int f = sh[0], m = sh[1];
        __syncthreads();
        if (f == INT_MAX) break;                 // no in-arc left: the walk ends here
        const int e = m != INT_MAX ? m : f;
        if (m != INT_MAX) { if (threadIdx.x == 0) a.path[n] = (int16_t)a.arc_dec[e]; n++; }
        cur = tail0 + d.arc_ts[e].x;
        l--;
    }
    if (threadIdx.x == 0) a.out->path_len = n;
}

}  // namespace

cudaError_t k2_launch(const K2DD *dds, int B, const double *coef, const double *rhs, int C, int Tpad, double *states, double *last,
                      int max_width_all, int max_layers, int avg_width, cudaStream_t st, int *launches) {
    const int K2_LI_CACHE = max_layers <= 640 ? max_layers : 0;   // shadows the constant: stage exactly what this batch needs
    dim3 grid(C, B);
    if (launches) (*launches)++;
    // narrow diagrams: smaller CTAs, more of them per SM (a layer is one barrier-separated step); wide ones (few CTAs
    // fit an SM: two state buffers of the full width each) get more threads per CTA instead
    // (measured, width 4096: 2.2e11 arcs/s at 256 threads, 3.7e11 at 512, 4.7e11 at 1024; width 1024 and the relaxed
    // diagram — wide only in its last layers — are fastest at 256: the AVERAGE width decides)
    int cap = avg_width >= 2048 ? K2_FIN_THREADS : avg_width >= 1200 ? 512 : K2_THREADS;
    if (const char *e = getenv("SGUFP_K2_THREADS")) { const int t = atoi(e); if (t >= 64 && t <= K2_FIN_THREADS) cap = t; }
    int threads = 64;
    while (threads < cap && threads < max_width_all) threads *= 2;
    const size_t smem = ((size_t)Tpad + 2 * (size_t)max_width_all) * sizeof(double);
    if (k2_states_in_smem(Tpad, max_width_all, avg_width)) {
        const size_t tot = smem + (size_t)K2_LI_CACHE * sizeof(int4);
        cudaError_t e = k2_raise_smem_limit(k2_longest_path<true, false>);
        if (e != cudaSuccess) return e;
        k2_longest_path<true, false><<<grid, threads, tot, st>>>(dds, coef, rhs, C, Tpad, max_width_all, states, last, K2_LI_CACHE);
    } else {
        // global-state variant: layers of at most K2_HYBRID nodes keep their states in two shared buffers all the same
        int hyb_w = K2_HYBRID;
        if (const char *e = getenv("SGUFP_K2_HYBRID")) hyb_w = std::max(0, atoi(e));
        const int hyb = (std::min(hyb_w, max_width_all) + 1) & ~1;      // even: the staged layer records behind the buffers are 16-byte aligned
        const size_t sm2 = ((size_t)Tpad + 2 * (size_t)hyb) * sizeof(double) + (size_t)K2_LI_CACHE * sizeof(int4);
        cudaError_t e = k2_raise_smem_limit(k2_longest_path<false, false>);
        if (e != cudaSuccess) return e;
        k2_longest_path<false, false><<<grid, threads, sm2, st>>>(dds, coef, rhs, C, Tpad, hyb, states, last, K2_LI_CACHE);
    }
    return cudaGetLastError();
}

bool k2_states_in_smem(int Tpad, int max_width_all, int avg_width) {
    const size_t bytes = ((size_t)Tpad + 2 * (size_t)max_width_all) * sizeof(double);
    if (bytes > 200 * 1024) return false;
    if (const char *e = getenv("SGUFP_K2_STATES")) { if (e[0] == 'g') return false; if (e[0] == 's') return true; }
    // A diagram that is narrow almost everywhere but very wide somewhere (the relaxed diagram: dozens of layers of a
    // few hundred nodes, then 10 k in the last ones) would pin one CTA per SM with two full-width buffers: its states
    // go to the global block instead and the SM stays full.
    if (bytes > 64 * 1024 && (long long)avg_width * 8 <= max_width_all) return false;
    return true;
}

cudaError_t k2_terminal_launch(const K2DD *dds, int B, int C, const double *last, double *bound, double *partial, int max_last, cudaStream_t st,
                               int *launches) {
    int slices = (max_last + 4 * K2_THREADS - 1) / (4 * K2_THREADS);      // ~4 nodes per thread
    slices = std::max(1, std::min(slices, K2_TERM_SLICES));
    k2_terminal<<<dim3(B, slices), K2_THREADS, 0, st>>>(dds, C, last, partial, slices, bound);
    if (slices > 1) k2_terminal_bound<<<B, 32, 0, st>>>(partial, slices, bound);
    if (launches) (*launches)++;           // counted as one step of the launch pair (longest path + terminal)
    return cudaGetLastError();
}

cudaError_t k2_single_launch(const K2DD *dd, const double *coef, const double *rhs, int Tpad, double *states, double *last,
                             int max_width, int max_layers, cudaStream_t st, int *launches) {
    const int K2_LI_CACHE = max_layers <= 640 ? max_layers : 0;
    if (launches) (*launches)++;
    int threads = 64;
    while (threads < K2_THREADS && threads < max_width) threads *= 2;
    const size_t smem = ((size_t)Tpad + 2 * (size_t)max_width) * sizeof(double);
    if (smem <= 200 * 1024) {
        const size_t tot = smem + (size_t)K2_LI_CACHE * sizeof(int4);
        cudaError_t e = k2_raise_smem_limit(k2_longest_path<true, true>);
        if (e != cudaSuccess) return e;
        k2_longest_path<true, true><<<dim3(1, 1), threads, tot, st>>>(dd, coef, rhs, 1, Tpad, max_width, states, last, K2_LI_CACHE);
    } else {
        const size_t sm2 = (size_t)Tpad * sizeof(double) + (size_t)K2_LI_CACHE * sizeof(int4);
        cudaError_t e = k2_raise_smem_limit(k2_longest_path<false, true>);
        if (e != cudaSuccess) return e;
        k2_longest_path<false, true><<<dim3(1, 1), threads, sm2, st>>>(dd, coef, rhs, 1, Tpad, 0, states, last, K2_LI_CACHE);
    }
    return cudaGetLastError();
}

cudaError_t k2_layered_launch(const K2DD &d, const int32_t *layer_width_host, const uint8_t *layer_collapsed_host, const double *coef,
                              double *states, cudaStream_t st, int *launches) {
    for (int l = 1; l < d.nlayers; l++) {
        if (layer_collapsed_host[l]) k2_layer_collapsed<<<1, K2_THREADS, 0, st>>>(d, coef, l, states);
        else k2_layer<<<(layer_width_host[l] + K2_THREADS - 1) / K2_THREADS, K2_THREADS, 0, st>>>(d, coef, l, states);
        if (launches) (*launches)++;
    }
    return cudaGetLastError();
}

cudaError_t k2_finish_launch(const K2Apply &a, cudaStream_t st, int *launches) {
    k2_finish<<<1, K2_FIN_THREADS, 0, st>>>(a);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

cudaError_t k2_sequence_launch(const K2DD *dd_device, const K2Apply &a, const K2Seq &q, const double *rhs_device, int max_width,
                               double *last_scratch, cudaStream_t st, int *launches) {
    const int K2_LI_CACHE = a.d.nlayers <= 640 ? a.d.nlayers : 0;
    // longest paths of cuts [k0, k1) side by side: global state blocks, one per cut, dead arcs honoured
    const int n = q.k1 - q.k0;
    int threads = 64;
    while (threads < K2_THREADS && threads < max_width) threads *= 2;
    const size_t sm2 = (size_t)q.Tpad * sizeof(double) + (size_t)K2_LI_CACHE * sizeof(int4);
    cudaError_t e = k2_raise_smem_limit(k2_longest_path<false, true>);
    if (e != cudaSuccess) return e;
    k2_longest_path<false, true><<<dim3(n, 1), threads, sm2, st>>>(dd_device, q.coef + (size_t)q.k0 * q.Tpad, rhs_device + q.k0, n, q.Tpad, 0,
                                                                     q.states, last_scratch, K2_LI_CACHE);
    if (q.probe) { k2_prune_probe<<<n, K2_THREADS, 0, st>>>(a, q); if (launches) (*launches)++; }
    if (q.bounds && a.mode == 0 && a.d.nlast > 0) {
        k2_window_prefix_min<<<(a.d.nlast + K2_THREADS - 1) / K2_THREADS, K2_THREADS, 0, st>>>(a, q);
        k2_window_bounds<<<n, K2_THREADS, 0, st>>>(a, q);
        if (launches) (*launches) += 2;
    }
    k2_finish_seq<<<1, K2_FIN_THREADS, 0, st>>>(a, q);
    if (launches) (*launches) += 2;
    return cudaGetLastError();
}

cudaError_t k2_extract_launch(const K2Apply &a, cudaStream_t st, int *launches) {
    k2_extract<<<1, K2_THREADS, 0, st>>>(a);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

}  // namespace sgufp
