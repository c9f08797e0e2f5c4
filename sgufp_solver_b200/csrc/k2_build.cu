// k2_build.cu — decision-diagram construction on the device, emitting the CSR image directly
// (SURVEY.md §8f-3): RelaxedDDNew::buildTree / buildNextLayer (/root/reference/DD.cpp:3528-3694) and
// RestrictedDDNew::compile / buildNextLayer / buildRestrictedLayer (DD.cpp:3090-3260).
//
// A node's state set is always a subset of {-1} U outgoingArcs(q) of the V-bar node q its layer
// belongs to (stateUpdateMap, Network.cpp:96-102; a child only ever loses the decision that led to
// it), so a state set is ONE 32-bit mask over that sorted base list: bit 0 is -1, iteration in set
// order is ascending bit order, the union of a collapsed layer is an OR, max_element is the highest
// bit.  One CTA builds the whole diagram layer by layer (a prefix sum over the children counts,
// then every parent writes its children); nothing but the sizes goes back to the host.
// The structure — layers, node order, in-arc order — is the reference's, node for node.
#include "k2_dd.cuh"

#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>

namespace cg = cooperative_groups;

namespace sgufp {
namespace {

constexpr int KB_THREADS = 512;

// exclusive prefix sum of cnt(i), i in [0, w), into off[]; returns the total (block-uniform)
template <class F>
__device__ int block_scan(int w, F cnt, int32_t *off, int *sh) {
    int carry = 0;
    for (int base = 0; base < w; base += KB_THREADS) {
        const int i = base + threadIdx.x;
        const int c = i < w ? cnt(i) : 0;
        int x = c;
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += t; }
        if ((threadIdx.x & 31) == 31) sh[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int y = threadIdx.x < KB_THREADS / 32 ? sh[threadIdx.x] : 0;
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, y, o); if (threadIdx.x >= o) y += t; }
            sh[32 + threadIdx.x] = y;                       // inclusive sums of the warp totals
        }
        __syncthreads();
        const int warp_base = (threadIdx.x >> 5) ? sh[32 + (threadIdx.x >> 5) - 1] : 0;
        if (i < w) off[i] = carry + warp_base + x - c;
        carry += sh[32 + KB_THREADS / 32 - 1];
        __syncthreads();
    }
    return carry;
}

__device__ unsigned block_or(unsigned v, unsigned *sh) {
    v = __reduce_or_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned r = 0;
    for (int i = 0; i < KB_THREADS / 32; i++) r |= sh[i];
    __syncthreads();
    return r;
}

__device__ unsigned block_addu(unsigned v, unsigned *sh) {
    v = __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned r = 0;
    for (int i = 0; i < KB_THREADS / 32; i++) r += sh[i];
    __syncthreads();
    return r;
}

// NC = CTAs that build one diagram together: 1 (narrow diagrams: no cluster barrier on the path), or a thread-block cluster of
// NC CTAs for wide layers — every CTA takes a contiguous slice of a layer's parents, the slice totals and the two layer-wide
// reductions go through distributed shared memory, and two cluster barriers per layer order the children written to global
// memory before the next layer reads them.  All CTAs run the same control flow on the same scalars.
template <int NC>
__device__ __forceinline__ void k2_build_body(const K2Build &b) {
    __shared__ int sh[64 + KB_THREADS / 32];
    __shared__ unsigned xch[2][3];                           // [layer parity][slice total, OR of the masks, children's states]
    unsigned *shu = reinterpret_cast<unsigned *>(sh);
    const K2Tables &t = b.t;
    int rank = 0;
    if constexpr (NC > 1) rank = (int)cg::this_cluster().block_rank();
    int v0 = 0, w = 1, n = 1, e_total = 0, nl = 1, exact = 1, exact_layer = 0, max_w = 1;
    unsigned next_size = 0;
    if (rank == 0 && threadIdx.x == 0) { b.mask[0] = b.root_mask; b.layer_info[0] = make_int4(0, 0, 1, 0); b.in_ptr[0] = 0; b.widths[0] = 1; }
    if constexpr (NC > 1) cg::this_cluster().sync(); else __syncthreads();
    for (int g = b.start; g < t.L; g++) {
        const int par = (g - b.start) & 1;
        const int tab = t.lay_tab[g], tp = t.tab_ptr[tab], nst = t.tab_ptr[tab + 1] - tp;
        const unsigned full = nst >= 32 ? 0xffffffffu : (1u << nst) - 1u;
        const int sb = t.slot_base[g];
        const int slice = NC > 1 ? (w + NC - 1) / NC : w, c0 = min(w, rank * slice), c1 = min(w, c0 + slice), wl = c1 - c0;
        if (t.lay_first[g]) {                               // states reset at the first layer of a V-bar node (DD.cpp:3565-3571)
            for (int i = c0 + threadIdx.x; i < c1; i += KB_THREADS) b.mask[v0 + i] = full;
            next_size = (unsigned)w * (unsigned)nst;
            __syncthreads();
        }
        const bool collapse = !b.restricted && next_size >= (unsigned)b.max_width && (unsigned)g < (unsigned)t.L - 5u;   // DD.cpp:3614, threshold 120 in the reference
        const bool one_child = b.restricted && !exact;      // DD.cpp:3204: the greatest state only
        auto cnt = [&](int i) { return one_child ? 1 : __popc(b.mask[v0 + c0 + i]); };
        int sum = block_scan(wl, cnt, b.off + c0, sh), base = 0;
        if constexpr (NC > 1) {                             // slice totals -> this slice's base and the layer's total
            cg::cluster_group cl = cg::this_cluster();
            if (threadIdx.x == 0) xch[par][0] = (unsigned)sum;
            cl.sync();
            sum = 0;
            for (int r = 0; r < NC; r++) { const unsigned tr = cl.map_shared_rank(&xch[par][0], r)[0]; if (r < rank) base += (int)tr; sum += (int)tr; }
        }
        int total = sum;
        if (b.restricted && exact && sum > b.max_width) total = b.max_width;   // the layer stops mid-node at max_width (DD.cpp:3238)
        const int new_nodes = collapse ? 1 : total;
        if (n + new_nodes > b.node_cap || e_total + (collapse ? sum : total) > b.arc_cap) {      // every CTA sees the same sizes and leaves together
            if (rank == 0 && threadIdx.x == 0) b.out->overflow = 1;
            if constexpr (NC > 1) cg::this_cluster().sync();   // nobody exits while a neighbour may still read its shared memory
            return;
        }
        unsigned uni = 0, child_states = 0;
        for (int i = c0 + threadIdx.x; i < c1; i += KB_THREADS) {
            const unsigned m = b.mask[v0 + i];
            const int c = one_child ? 1 : __popc(m), o = base + b.off[i];
            uni |= m;
            for (int r = 0; r < c; r++) {
                const int k = o + r;
                if (k >= total) break;
                // relaxed: states in set order; restricted: in REVERSE order (DD.cpp:3235); non-exact: the greatest
                const int pos = one_child ? 31 - __clz(m) : (b.restricted ? __fns(m, 0, __popc(m) - r) : __fns(m, 0, r + 1));
                const int dec = t.tab_dec[tp + pos], kk = t.tab_k[tp + pos];
                const int e = e_total + k;
                b.arc_ts[e] = make_int2(i, kk >= 0 ? sb + kk : -1);
                b.arc_dec[e] = dec;
                if (!collapse) {
                    const unsigned cm = pos ? m & ~(1u << pos) : m;   // the child loses the decision taken, -1 is never removed
                    b.mask[n + k] = cm;
                    b.in_ptr[n + k] = e;
                    child_states += (unsigned)__popc(cm);
                }
            }
        }
        uni = block_or(uni, shu);
        child_states = block_addu(child_states, shu);
        if constexpr (NC > 1) {                             // layer-wide reductions; the barrier also publishes the children
            cg::cluster_group cl = cg::this_cluster();
            if (threadIdx.x == 0) { xch[par][1] = uni; xch[par][2] = child_states; }
            cl.sync();
            uni = 0; child_states = 0;
            for (int r = 0; r < NC; r++) { const unsigned *x = cl.map_shared_rank(&xch[par][0], r); uni |= x[1]; child_states += x[2]; }
        }
        if (collapse) {
            if (rank == 0 && threadIdx.x == 0) { b.mask[n] = uni; b.in_ptr[n] = e_total; b.layer_info[nl] = make_int4(n, e_total, 1, sum == 1 ? 1 : 0); b.widths[nl] = 1; }
            next_size = (unsigned)__popc(uni);
            exact = 0;
            v0 = n; w = 1; n += 1; e_total += sum;
            if constexpr (NC > 1) cg::this_cluster().sync();   // the one node of the collapsed layer is written by rank 0
        } else {
            next_size = child_states;
            if (rank == 0 && threadIdx.x == 0) { b.layer_info[nl] = make_int4(n, e_total, total, 1); b.widths[nl] = total; }
            if (b.restricted && exact && sum > b.max_width) exact = 0;
            v0 = n; w = total; n += total; e_total += total;
        }
        if (b.restricted && exact) exact_layer++;
        max_w = max(max_w, w);
        nl++;
        __syncthreads();
    }
    if constexpr (NC > 1) cg::this_cluster().sync();       // a CTA's shared memory must outlive its neighbours' last remote reads
    if (rank == 0 && threadIdx.x == 0) {
        b.in_ptr[n] = e_total;
        K2BuildOut o;
        o.nlayers = nl; o.nnodes = n; o.narcs = e_total; o.exact = exact; o.exact_layer = exact_layer; o.overflow = 0; o.max_width = max_w; o.nlast = w;
        *b.out = o;
    }
}

__global__ void __launch_bounds__(KB_THREADS) k2_build(K2Build b) { k2_build_body<1>(b); }
constexpr int KB_CLUSTER = 8;                                // portable cluster size
__global__ void __cluster_dims__(KB_CLUSTER, 1, 1) __launch_bounds__(KB_THREADS) k2_build_cluster(K2Build b) { k2_build_body<KB_CLUSTER>(b); }

// getExactCutSet of a restricted tree (DD.cpp:3279-3288): for every node of the last exact layer its state
// mask and the decisions of its single-parent chain, root first.
__global__ void k2_cutset(const int4 *layer_info, const int32_t *in_ptr, const int2 *arc_ts, const int32_t *arc_dec, const unsigned *mask,
                          int el, unsigned *out_mask, int16_t *out_dec) {
    const int4 li = layer_info[el];
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= li.z) return;
    int v = li.x + j;
    out_mask[j] = mask[v];
    for (int l = el; l >= 1; l--) {
        const int e = in_ptr[v];
        out_dec[(size_t)j * el + (l - 1)] = (int16_t)arc_dec[e];
        v = layer_info[l - 1].x + arc_ts[e].x;
    }
}

// RelaxedDDNew::getCutset (DD.cpp:4179-4218) on the device image: the cut-set is the set of arcs that enter the first
// layer (from layer 3 on) with ONE live node; one thread per such arc writes the state mask of its tail, its decision
// and the tail's path (getPathForNode, DD.cpp:3796-3820: upwards over the first live in-arc whose parent state + weight
// EQUALS the node's state, recorded; otherwise the first live in-arc, not recorded), leaf first.
__global__ void k2_relaxed_cutset(K2Apply a, const unsigned *mask, int layer, int v, int e0, int stride, unsigned *out_mask, int32_t *out_len,
                                  int16_t *out_dec) {
    const K2DD &d = a.d;
    const int j = blockIdx.x * blockDim.x + threadIdx.x, e = e0 + j;
    if (e >= d.in_ptr[v + 1]) return;
    if (a.arc_dead[e]) { out_len[j] = -1; return; }
    int cur = d.layer_info[layer - 1].x + d.arc_ts[e].x, n = 0;
    out_mask[j] = mask[cur];
    out_dec[(size_t)j * stride + n++] = (int16_t)a.arc_dec[e];
    for (int l = layer - 1; l > 0; l--) {
        const int tail0 = d.layer_info[l - 1].x;
        const double sc = a.state[cur];
        int first = -1, match = -1;
        for (int x = d.in_ptr[cur]; x < d.in_ptr[cur + 1] && match < 0; x++) {
            if (a.arc_dead[x]) continue;
            if (first < 0) first = x;
            const int2 ts = d.arc_ts[x];
            const double w = ts.y >= 0 ? a.coef[ts.y] : 0.0;
            if ((a.state[tail0 + ts.x] + w) == sc) match = x;
        }
        if (first < 0) break;
        const int x = match >= 0 ? match : first;
        if (match >= 0) out_dec[(size_t)j * stride + n++] = (int16_t)a.arc_dec[x];
        cur = tail0 + d.arc_ts[x].x;
    }
    out_len[j] = n;
}

// the first layer >= 3 with one live node, and that node (result: {layer, node} or {-1, -1})
__global__ void k2_first_collapsed(K2Apply a, int *out2) {
    const K2DD &d = a.d;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    out2[0] = -1; out2[1] = -1; out2[2] = 0; out2[3] = 0;
    for (int layer = 3; layer < d.nlayers; layer++)
        if (a.layer_alive[layer] == 1) {
            const int4 li = d.layer_info[layer];
            for (int i = 0; i < li.z; i++)
                if (!a.node_dead[li.x + i]) { out2[0] = layer; out2[1] = li.x + i; out2[2] = d.in_ptr[li.x + i]; out2[3] = d.in_ptr[li.x + i + 1]; return; }
            return;
        }
}

__global__ void k2_fill(double *p, double v, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace

// One CTA for diagrams whose layers stay narrow (a cluster barrier per layer would cost more than it saves: the relaxed C2
// diagram builds in 0.34 ms), a cluster of 8 CTAs on 8 SMs for wide layers.  SGUFP_DD_BUILD_CLUSTER=0/1 forces the choice.
cudaError_t k2_build_launch(const K2Build &b, cudaStream_t st, int *launches) {
    bool cluster = b.max_width >= 2048;
    if (const char *e = getenv("SGUFP_DD_BUILD_CLUSTER")) cluster = e[0] == '1';
    if (cluster) k2_build_cluster<<<KB_CLUSTER, KB_THREADS, 0, st>>>(b);
    else k2_build<<<1, KB_THREADS, 0, st>>>(b);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

cudaError_t k2_cutset_launch(const int4 *layer_info, const int32_t *in_ptr, const int2 *arc_ts, const int32_t *arc_dec, const unsigned *mask,
                             int el, int count, unsigned *out_mask, int16_t *out_dec, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    k2_cutset<<<(count + 255) / 256, 256, 0, st>>>(layer_info, in_ptr, arc_ts, arc_dec, mask, el, out_mask, out_dec);
    return cudaGetLastError();
}

cudaError_t k2_first_collapsed_launch(const K2Apply &a, int *out4, cudaStream_t st) {
    k2_first_collapsed<<<1, 32, 0, st>>>(a, out4);
    return cudaGetLastError();
}

cudaError_t k2_relaxed_cutset_launch(const K2Apply &a, const unsigned *mask, int layer, int v, int e0, int count, int stride, unsigned *out_mask,
                                     int32_t *out_len, int16_t *out_dec, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    k2_relaxed_cutset<<<(count + 127) / 128, 128, 0, st>>>(a, mask, layer, v, e0, stride, out_mask, out_len, out_dec);
    return cudaGetLastError();
}

cudaError_t k2_fill_launch(double *p, double v, long long n, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const int blocks = (int)std::min<long long>((n + 255) / 256, 1184);
    k2_fill<<<blocks, 256, 0, st>>>(p, v, n);
    return cudaGetLastError();
}

}  // namespace sgufp
