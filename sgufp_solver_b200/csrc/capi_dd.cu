// capi_dd.cu — C ABI of include/sgufp_b200_dd.h: host diagrams + the K2 device pass.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/sgufp_b200_dd.h"
#include "ctx.hpp"
#include "dd_host.hpp"
#include "k2_dd.cuh"

using namespace sgufp;

struct sgufp_dd {
    sgufp_ctx *ctx = nullptr;
    HostDD *dd = nullptr;
    std::vector<NodeSpec> compile_cutset;
    bool has_cutset = false;
    // device image
    DevBuf<int32_t> d_i32;        // layer_ptr | in_ptr | arc_tail | arc_slot | root_slot
    DevBuf<double> d_term;
    K2DD dev{};
    bool uploaded = false, term_dirty = true;
    unsigned long uploaded_version = 0;
    // one-cut-at-a-time path: the cut semantics run on the device, the host mirror catches up on demand
    DevBuf<double> d_state, d_last, d_coef, d_rhs;
    DevBuf<uint8_t> d_arc_dead, d_node_dead, d_lost;
    DevBuf<int32_t> d_layer_alive, d_cnt, d_arc_dec;
    DevBuf<int16_t> d_path;
    DevBuf<K2DD> d_self;
    DevBuf<K2Result> d_res;
    std::vector<double> last_coef;     // dense coefficients of the cut the device states belong to
    bool dev_ahead = false;            // the device holds states / terminal weights / removals the host has not seen
    bool dev_states_valid = false;     // d_state, d_coef, term describe the last applied cut
    int narcs = 0;
    std::vector<int32_t> layer_width;  // per layer of the uploaded image: nodes, and whether it is a collapsed node
    std::vector<uint8_t> layer_collapsed;
    std::vector<int32_t> root_slot;
    // construction on the device (k2_build.cu): the host mirror is then built only when somebody needs it
    NodeSpec root_spec;
    std::vector<int16_t> root_solution;
    bool host_built = true, device_built = false, exact_flag = true;
    int nlayers_dev = 1, exact_layer_dev = 0;
    std::vector<int16_t> cached_path;  // getSolution of the state left by the last one-cut call (extracted in the same call)
    bool cached_path_valid = false;
    DevBuf<int4> b_layer_info;
    DevBuf<int2> b_arc_ts;
    DevBuf<int32_t> b_in_ptr, b_off, b_widths, b_root_slot;
    DevBuf<unsigned> b_mask;
    DevBuf<K2BuildOut> b_out;
};

namespace {

int upload(sgufp_dd *d) {
    sgufp_ctx *c = d->ctx;
    if (c->device == SGUFP_DEVICE_NONE) return fail(c, SGUFP_ERR_CUDA, "handle was created with SGUFP_DEVICE_NONE: there is no CPU compute path for K2");
    CU(c, cudaSetDevice(c->device));
    // an image built on the device stays as it is until the host mirror (if it exists at all) changes shape
    if (d->device_built && (!d->host_built || d->uploaded_version == d->dd->version())) return 0;
    d->device_built = false;
    const bool was_dirty = !d->uploaded || d->uploaded_version != d->dd->version();
    const DDCsr &C = d->dd->flatten();
    if (was_dirty) {
        // layer_info (int4) | arc_ts (int2) | in_ptr | root_slot, one upload
        std::vector<int32_t> pack((size_t)4 * C.nlayers + (size_t)2 * C.narcs, 0);
        for (int l = 0; l < C.nlayers; l++) {
            const int v0 = C.layer_ptr[l], v1 = C.layer_ptr[l + 1];
            bool uniform = l > 0;
            for (int v = v0; v < v1 && uniform; v++) uniform = C.in_ptr[v + 1] - C.in_ptr[v] == 1;
            pack[4 * l] = v0; pack[4 * l + 1] = C.in_ptr[v0]; pack[4 * l + 2] = v1 - v0; pack[4 * l + 3] = uniform ? 1 : 0;
            if (l > 0)
                for (int e = C.in_ptr[v0]; e < C.in_ptr[v1]; e++) {
                    pack[(size_t)4 * C.nlayers + 2 * e] = C.arc_tail[e] - C.layer_ptr[l - 1];
                    pack[(size_t)4 * C.nlayers + 2 * e + 1] = C.arc_slot[e];
                }
        }
        const size_t o_ts = (size_t)4 * C.nlayers, o_ip = pack.size();
        pack.insert(pack.end(), C.in_ptr.begin(), C.in_ptr.end());
        const size_t o_rs = pack.size();
        pack.insert(pack.end(), C.root_slot.begin(), C.root_slot.end());
        CU(c, d->d_i32.reserve(pack.size() + 4));
        CU(c, cudaMemcpyAsync(d->d_i32.p, pack.data(), pack.size() * 4, cudaMemcpyHostToDevice, c->st));
        d->dev.layer_info = reinterpret_cast<const int4 *>(d->d_i32.p);
        d->dev.arc_ts = reinterpret_cast<const int2 *>(d->d_i32.p + o_ts);
        d->dev.in_ptr = d->d_i32.p + o_ip; d->dev.root_slot = d->d_i32.p + o_rs;
        d->dev.nlayers = C.nlayers; d->dev.nroot = (int)C.root_slot.size(); d->dev.nnodes = C.nnodes; d->dev.nlast = C.nlast;
        d->dev.max_width = C.max_width;
        // device-side cut application: decisions, removal flags, live layer sizes
        std::vector<int32_t> dec(C.narcs), alive(C.nlayers);
        for (int e = 0; e < C.narcs; e++) dec[e] = d->dd->arcs()[C.arc_id[e]].decision;
        for (int l = 0; l < C.nlayers; l++) alive[l] = C.layer_ptr[l + 1] - C.layer_ptr[l];
        d->layer_width = alive;
        d->layer_collapsed.assign(C.nlayers, 0);
        for (int l = 1; l < C.nlayers; l++) d->layer_collapsed[l] = (alive[l] == 1 && pack[4 * l + 3] == 0) ? 1 : 0;
        d->root_slot = C.root_slot;
        CU(c, d->d_arc_dec.reserve(C.narcs)); CU(c, d->d_layer_alive.reserve(C.nlayers));
        CU(c, d->d_arc_dead.reserve(C.narcs)); CU(c, d->d_node_dead.reserve(C.nnodes)); CU(c, d->d_lost.reserve(C.nnodes));
        CU(c, d->d_cnt.reserve(C.nnodes)); CU(c, d->d_state.reserve(C.nnodes)); CU(c, d->d_last.reserve(C.nlast));
        CU(c, d->d_path.reserve(C.nlayers)); CU(c, d->d_self.reserve(1)); CU(c, d->d_res.reserve(1)); CU(c, d->d_rhs.reserve(1));
        if (C.narcs) CU(c, cudaMemcpyAsync(d->d_arc_dec.p, dec.data(), (size_t)C.narcs * 4, cudaMemcpyHostToDevice, c->st));
        CU(c, cudaMemcpyAsync(d->d_layer_alive.p, alive.data(), (size_t)C.nlayers * 4, cudaMemcpyHostToDevice, c->st));
        CU(c, cudaMemsetAsync(d->d_arc_dead.p, 0, std::max(1, C.narcs), c->st));
        CU(c, cudaMemsetAsync(d->d_node_dead.p, 0, C.nnodes, c->st));
        CU(c, cudaMemsetAsync(d->d_lost.p, 0, C.nnodes, c->st));
        CU(c, cudaStreamSynchronize(c->st));   // `pack`, `dec`, `alive` are pageable temporaries
        d->dev.arc_dead = d->d_arc_dead.p;
        d->narcs = C.narcs;
        d->uploaded = true;
        d->uploaded_version = d->dd->version();
        d->term_dirty = true;
        d->dev_ahead = false; d->dev_states_valid = false;
    }
    if (d->term_dirty) {
        const std::vector<double> &t = d->dd->terminal_weights();
        CU(c, d->d_term.reserve(t.size()));
        if (!t.empty()) CU(c, cudaMemcpyAsync(d->d_term.p, t.data(), t.size() * 8, cudaMemcpyHostToDevice, c->st));
        d->dev.term = d->d_term.p;
        d->term_dirty = false;
    }
    return 0;
}

// the host mirror of a diagram that was built on the device: the same deterministic construction, on demand
void ensure_host(sgufp_dd *d) {
    if (d->host_built) return;
    d->dd->build(d->root_spec, &d->compile_cutset);
    d->dd->flatten();
    d->uploaded_version = d->dd->version();   // same structure, same CSR order as the device image: no upload
    d->host_built = true;
}

// bring the host mirror up to date with what the device did since the last upload
int sync_host(sgufp_dd *d) {
    ensure_host(d);
    if (!d->dev_ahead) return 0;
    sgufp_ctx *c = d->ctx;
    CU(c, cudaSetDevice(c->device));
    const int nn = d->dev.nnodes, na = d->narcs, nlast = d->dev.nlast;
    std::vector<double> states(nn), term(nlast);
    std::vector<uint8_t> adead(std::max(1, na)), ndead(nn);
    CU(c, cudaMemcpyAsync(states.data(), d->d_state.p, (size_t)nn * 8, cudaMemcpyDeviceToHost, c->st));
    if (nlast) CU(c, cudaMemcpyAsync(term.data(), d->d_term.p, (size_t)nlast * 8, cudaMemcpyDeviceToHost, c->st));
    if (na) CU(c, cudaMemcpyAsync(adead.data(), d->d_arc_dead.p, (size_t)na, cudaMemcpyDeviceToHost, c->st));
    CU(c, cudaMemcpyAsync(ndead.data(), d->d_node_dead.p, (size_t)nn, cudaMemcpyDeviceToHost, c->st));
    CU(c, cudaStreamSynchronize(c->st));
    d->dd->absorb_device(states, adead, ndead, term, d->last_coef);
    d->dev_ahead = false;
    return 0;
}

// Single-cut longest path as one launch per layer over the whole GPU, or as one CTA walking all layers?  A launch
// costs ~4 us, so the layered shape pays only when the AVERAGE layer is wide (a relaxed diagram has a few wide
// layers at the end and dozens of narrow ones: one CTA wins there).  SGUFP_K2_LAYERED_MIN overrides the threshold
// on the average width (read per call: the tests force either path).
bool use_layered(const sgufp_dd *d) {
    const char *e = getenv("SGUFP_K2_LAYERED_MIN");
    const long long min_avg = e ? std::max(1, atoi(e)) : 1024;
    return (long long)d->dev.nnodes >= min_avg * d->dev.nlayers;
}

K2Apply make_apply(sgufp_dd *d, int mode, double optimal) {
    K2Apply a{};
    a.d = d->dev; a.d.state_off = 0; a.d.last_off = 0;
    a.state = d->d_state.p; a.coef = d->d_coef.p; a.arc_dec = d->d_arc_dec.p;
    a.arc_dead = d->d_arc_dead.p; a.node_dead = d->d_node_dead.p; a.layer_alive = d->d_layer_alive.p;
    a.cnt = d->d_cnt.p; a.lost = d->d_lost.p; a.out = d->d_res.p; a.path = d->d_path.p;
    a.optimal = optimal; a.mode = mode; a.restricted = d->dd->restricted() ? 1 : 0; a.exact = d->exact_flag ? 1 : 0;
    return a;
}

// One cut on one diagram, entirely on the device: longest path, terminal weights, removals.
int apply_on_device(sgufp_dd *d, double rhs, const uint64_t *keys, const double *vals, int nnz, int mode, double optimal, K2Result &res) {
    sgufp_ctx *c = d->ctx;
    if (int rc = upload(d)) return rc;
    const int T = std::max(1, c->M.T), Tpad = (T + 1) & ~1;
    std::vector<double> coef;
    d->dd->densify(keys, vals, nnz, coef);
    d->last_coef = coef;
    coef.resize(Tpad, 0.0);
    CU(c, d->d_coef.reserve(Tpad));
    K2DD self = d->dev; self.state_off = 0; self.last_off = 0;
    CU(c, cudaMemcpyAsync(d->d_coef.p, coef.data(), (size_t)Tpad * 8, cudaMemcpyHostToDevice, c->st));
    CU(c, cudaMemcpyAsync(d->d_rhs.p, &rhs, 8, cudaMemcpyHostToDevice, c->st));
    CU(c, cudaMemcpyAsync(d->d_self.p, &self, sizeof(K2DD), cudaMemcpyHostToDevice, c->st));
    c->dd_launches = 0;
    CU(c, cudaEventRecord(c->evk0, c->st));
    if (use_layered(d)) {
        // wide diagram: the root state on the host (same additions in the same order), then one launch per layer
        double v = rhs;
        for (int s : d->root_slot) if (s >= 0) v = v + coef[s];
        CU(c, cudaMemcpyAsync(d->d_state.p, &v, 8, cudaMemcpyHostToDevice, c->st));
        CU(c, k2_layered_launch(self, d->layer_width.data(), d->layer_collapsed.data(), d->d_coef.p, d->d_state.p, c->st, &c->dd_launches));
    } else
        CU(c, k2_single_launch(d->d_self.p, d->d_coef.p, d->d_rhs.p, Tpad, d->d_state.p, d->d_last.p, d->dev.max_width, d->dev.nlayers, c->st, &c->dd_launches));
    CU(c, k2_finish_launch(make_apply(d, mode, optimal), c->st, &c->dd_launches));
    // the Benders loop asks for the path right after every cut (NodeExplorer.cpp:950): extract it in the same call
    const char *fe = getenv("SGUFP_DD_FUSED_PATH");
    const bool fused = !(fe && fe[0] == '0');
    if (fused) CU(c, k2_extract_launch(make_apply(d, mode, optimal), c->st, &c->dd_launches));
    CU(c, cudaEventRecord(c->evk1, c->st));
    std::vector<int16_t> rev(d->dev.nlayers);
    CU(c, cudaMemcpyAsync(&res, d->d_res.p, sizeof(K2Result), cudaMemcpyDeviceToHost, c->st));
    if (fused) CU(c, cudaMemcpyAsync(rev.data(), d->d_path.p, (size_t)d->dev.nlayers * 2, cudaMemcpyDeviceToHost, c->st));
    CU(c, cudaStreamSynchronize(c->st));   // also covers the pageable `coef`, `rhs`, `self`
    d->cached_path_valid = false;
    if (fused) {
        d->cached_path = d->root_solution;
        for (int i = res.path_len - 1; i >= 0; i--) d->cached_path.push_back(rev[i]);
        d->cached_path_valid = true;
    }
    CU(c, cudaEventElapsedTime(&c->dd_kernel_ms, c->evk0, c->evk1));
    c->dd_arcs = (long long)d->narcs + d->dev.nlast;   // in-arcs + terminal arcs of the uploaded image
    d->dev_ahead = true; d->dev_states_valid = true;
    d->term_dirty = false;                  // the device holds the authoritative terminal weights now
    return 0;
}

struct Scratch {   // per-context scratch for the batch call + the static tables of the device construction
    DevBuf<double> coef, rhs, states, last, bound, wbounds;
    DevBuf<K2DD> dds;
    DevBuf<K2Result> seq_res;
    DevBuf<int> seq_ctl, seq_probe;
    bool tables_built = false, tables_ok = false;
    DevBuf<int32_t> t_lay_tab, t_tab_ptr, t_slot_base;
    DevBuf<uint8_t> t_lay_first;
    DevBuf<int16_t> t_tab_dec, t_tab_k;
    K2Tables tables{};
    std::vector<int32_t> h_lay_tab, h_tab_ptr;
    std::vector<int16_t> h_tab_dec;
    std::vector<uint8_t> h_lay_first;
};
Scratch &scratch_of(sgufp_ctx *c) {
    if (!c->dd_scratch) {
        c->dd_scratch = new Scratch();
        c->dd_scratch_free = [](void *p) {
            Scratch *s = static_cast<Scratch *>(p);
            s->coef.release(); s->rhs.release(); s->states.release(); s->last.release(); s->bound.release(); s->dds.release();
            s->wbounds.release(); s->seq_res.release(); s->seq_ctl.release(); s->seq_probe.release();
            s->t_lay_tab.release(); s->t_tab_ptr.release(); s->t_slot_base.release(); s->t_lay_first.release(); s->t_tab_dec.release(); s->t_tab_k.release();
            delete s;
        };
    }
    return *static_cast<Scratch *>(c->dd_scratch);
}


// static tables of the device construction: per V-bar node the base state list {-1} U outgoingArcs (ascending,
// Network.cpp:96-102) and, per entry, the coefficient slot of that decision relative to slot_base[layer]
int ensure_tables(sgufp_ctx *c) {
    Scratch &S = scratch_of(c);
    if (S.tables_built) return 0;
    const Model &M = c->M;
    std::vector<int32_t> lay_tab(M.L), tab_ptr{0}, tab_of(M.n, -1);
    std::vector<uint8_t> lay_first(M.L, 0);
    std::vector<int16_t> tab_dec, tab_k;
    bool ok = M.m <= 32767;
    for (int g = 0; g < M.L; g++) {
        const int q = M.head[M.layer_arc[g]];
        lay_first[g] = (g == 0 || M.head[M.layer_arc[g - 1]] != q) ? 1 : 0;
        if (tab_of[q] < 0) {
            tab_of[q] = (int)tab_ptr.size() - 1;
            std::vector<int32_t> outs(M.out_arc.begin() + M.out_ptr[q], M.out_arc.begin() + M.out_ptr[q + 1]);
            std::sort(outs.begin(), outs.end());
            tab_dec.push_back(-1); tab_k.push_back(-1);
            for (int a : outs) {
                int k = -1;
                for (int i = 0; i < M.outdeg(q) && k < 0; i++) if (M.head[M.out_arc[M.out_ptr[q] + i]] == M.head[a]) k = i;
                tab_dec.push_back((int16_t)a); tab_k.push_back((int16_t)k);
            }
            if (outs.size() + 1 > 32) ok = false;          // a state set is one 32-bit mask
            tab_ptr.push_back((int32_t)tab_dec.size());
        }
        lay_tab[g] = tab_of[q];
    }
    S.tables_built = true; S.tables_ok = false;
    if (!ok || M.L == 0 || c->device == SGUFP_DEVICE_NONE) return 0;
    CU(c, cudaSetDevice(c->device));
    CU(c, S.t_lay_tab.reserve(M.L)); CU(c, S.t_lay_first.reserve(M.L)); CU(c, S.t_slot_base.reserve(M.L));
    CU(c, S.t_tab_ptr.reserve(tab_ptr.size())); CU(c, S.t_tab_dec.reserve(tab_dec.size())); CU(c, S.t_tab_k.reserve(tab_k.size()));
    CU(c, cudaMemcpy(S.t_lay_tab.p, lay_tab.data(), (size_t)M.L * 4, cudaMemcpyHostToDevice));
    CU(c, cudaMemcpy(S.t_lay_first.p, lay_first.data(), (size_t)M.L, cudaMemcpyHostToDevice));
    CU(c, cudaMemcpy(S.t_slot_base.p, M.slot_base.data(), (size_t)M.L * 4, cudaMemcpyHostToDevice));
    CU(c, cudaMemcpy(S.t_tab_ptr.p, tab_ptr.data(), tab_ptr.size() * 4, cudaMemcpyHostToDevice));
    CU(c, cudaMemcpy(S.t_tab_dec.p, tab_dec.data(), tab_dec.size() * 2, cudaMemcpyHostToDevice));
    CU(c, cudaMemcpy(S.t_tab_k.p, tab_k.data(), tab_k.size() * 2, cudaMemcpyHostToDevice));
    S.tables = K2Tables{S.t_lay_tab.p, S.t_lay_first.p, S.t_tab_ptr.p, S.t_tab_dec.p, S.t_tab_k.p, S.t_slot_base.p, M.L};
    S.h_lay_tab = lay_tab; S.h_tab_ptr = tab_ptr; S.h_tab_dec = tab_dec; S.h_lay_first = lay_first;
    S.tables_ok = true;
    return 0;
}

bool device_build_enabled() { const char *e = getenv("SGUFP_DD_BUILD"); return !(e && e[0] == 'h'); }   // SGUFP_DD_BUILD=host forces the host builder

// Build the diagram on the device (k2_build.cu).  Returns 1 if it did, 0 if this case is left to the
// host builder (no device, state sets wider than 32, a root that is not a cut-set node, capacity), < 0 on error.
int device_build(sgufp_dd *d, const NodeSpec &root, int *cutset_nodes) {
    sgufp_ctx *c = d->ctx;
    const Model &M = c->M;
    if (c->device == SGUFP_DEVICE_NONE || !device_build_enabled()) return 0;
    if (int rc = ensure_tables(c)) return rc;
    Scratch &S = scratch_of(c);
    const int start = root.global_layer;
    if (!S.tables_ok || start >= M.L || (int)root.solution.size() != start) return 0;
    const int tab = S.h_lay_tab[start], tp = S.h_tab_ptr[tab], nst = S.h_tab_ptr[tab + 1] - tp;
    unsigned mask = 0;
    if (!S.h_lay_first[start]) {                             // at the first layer of a V-bar node the states are reset anyway
        for (int16_t s : root.states) {
            int pos = -1;
            for (int p = 0; p < nst && pos < 0; p++) if (S.h_tab_dec[tp + p] == s) pos = p;
            if (pos < 0 || ((mask >> pos) & 1)) return 0;   // not a subset of the layer's base list (or a repeated state)
            mask |= 1u << pos;
        }
        if (!(mask & 1u)) return 0;                          // every reference state set holds -1
        for (size_t i = 1; i < root.states.size(); i++) if (root.states[i - 1] >= root.states[i]) return 0;   // set order
    }
    const bool restricted = d->dd->restricted();
    const long long layers = M.L - start;
    CU(c, cudaSetDevice(c->device));
    // capacity: exact for a restricted tree (a layer never exceeds max_width); a relaxed diagram starts with room for
    // 256 k nodes and is rebuilt with 8x more when it runs out (up to 16 M nodes), so that the usual diagram of a few
    // ten thousand nodes does not pin hundreds of megabytes
    long long node_cap = restricted ? 1 + layers * (long long)d->dd->max_width() : std::max<long long>(1LL << 18, (long long)d->b_mask.cap);
    if (node_cap > (1LL << 26)) return 0;
    K2BuildOut o{};
    for (;;) {
        const long long arc_cap = restricted ? node_cap : 2 * node_cap;
        CU(c, d->b_layer_info.reserve(layers + 1)); CU(c, d->b_in_ptr.reserve(node_cap + 1)); CU(c, d->b_arc_ts.reserve(arc_cap));
        CU(c, d->d_arc_dec.reserve(arc_cap)); CU(c, d->b_mask.reserve(node_cap)); CU(c, d->b_off.reserve(node_cap));
        CU(c, d->b_widths.reserve(layers + 1)); CU(c, d->b_out.reserve(1));
        K2Build b{};
        b.t = S.tables; b.start = start; b.root_mask = mask; b.restricted = restricted ? 1 : 0;
        b.max_width = restricted ? d->dd->max_width() : (d->dd->max_width() > 0 ? d->dd->max_width() : 120);
        b.node_cap = (int)node_cap; b.arc_cap = (int)arc_cap;
        b.layer_info = d->b_layer_info.p; b.in_ptr = d->b_in_ptr.p; b.arc_ts = d->b_arc_ts.p; b.arc_dec = d->d_arc_dec.p; b.mask = d->b_mask.p;
        b.off = d->b_off.p; b.widths = d->b_widths.p; b.out = d->b_out.p;
        o = K2BuildOut{};
        o.overflow = 1;
        CU(c, cudaMemcpyAsync(d->b_out.p, &o, sizeof(o), cudaMemcpyHostToDevice, c->st));
        c->dd_launches = 0;
        CU(c, k2_build_launch(b, c->st, &c->dd_launches));
        CU(c, cudaMemcpyAsync(&o, d->b_out.p, sizeof(o), cudaMemcpyDeviceToHost, c->st));
        CU(c, cudaStreamSynchronize(c->st));
        if (!o.overflow) break;
        if (restricted || node_cap >= (1LL << 24)) return 0;    // leave it to the host builder
        node_cap *= 8;
    }
    std::vector<int4> li(o.nlayers);
    d->layer_width.resize(o.nlayers);
    CU(c, cudaMemcpyAsync(li.data(), d->b_layer_info.p, (size_t)o.nlayers * sizeof(int4), cudaMemcpyDeviceToHost, c->st));
    CU(c, cudaMemcpyAsync(d->layer_width.data(), d->b_widths.p, (size_t)o.nlayers * 4, cudaMemcpyDeviceToHost, c->st));
    // what the single-cut path keeps per diagram
    const int T = std::max(1, M.T), Tpad = (T + 1) & ~1;
    d->root_slot.clear();
    for (int i = 0; i < start; i++) d->root_slot.push_back(d->dd->slot_of(i, root.solution[i]));
    CU(c, d->b_root_slot.reserve(std::max(1, start)));
    if (start) CU(c, cudaMemcpyAsync(d->b_root_slot.p, d->root_slot.data(), (size_t)start * 4, cudaMemcpyHostToDevice, c->st));
    CU(c, d->d_term.reserve(o.nlast)); CU(c, d->d_layer_alive.reserve(o.nlayers));
    CU(c, d->d_arc_dead.reserve(std::max(1, o.narcs))); CU(c, d->d_node_dead.reserve(o.nnodes)); CU(c, d->d_lost.reserve(o.nnodes));
    CU(c, d->d_cnt.reserve(o.nnodes)); CU(c, d->d_state.reserve(o.nnodes)); CU(c, d->d_last.reserve(o.nlast)); CU(c, d->d_coef.reserve(Tpad));
    CU(c, d->d_path.reserve(o.nlayers)); CU(c, d->d_self.reserve(1)); CU(c, d->d_res.reserve(1)); CU(c, d->d_rhs.reserve(1));
    CU(c, cudaMemcpyAsync(d->d_layer_alive.p, d->b_widths.p, (size_t)o.nlayers * 4, cudaMemcpyDeviceToDevice, c->st));
    CU(c, cudaMemsetAsync(d->d_arc_dead.p, 0, std::max(1, o.narcs), c->st));
    CU(c, cudaMemsetAsync(d->d_node_dead.p, 0, o.nnodes, c->st));
    CU(c, cudaMemsetAsync(d->d_lost.p, 0, o.nnodes, c->st));
    CU(c, cudaMemsetAsync(d->d_coef.p, 0, (size_t)Tpad * 8, c->st));
    CU(c, k2_fill_launch(d->d_term.p, DD_MAX, o.nlast, c->st));          // terminal arcs start at DOUBLE_MAX (DD.cpp:3145, 3595)
    CU(c, k2_fill_launch(d->d_state.p, DD_LOWEST, o.nnodes, c->st));     // state2 of a fresh node (DD.h:453)
    CU(c, cudaStreamSynchronize(c->st));
    d->layer_collapsed.assign(o.nlayers, 0);
    for (int l = 1; l < o.nlayers; l++) d->layer_collapsed[l] = (li[l].z == 1 && li[l].w == 0) ? 1 : 0;
    K2DD dev{};
    dev.layer_info = d->b_layer_info.p; dev.in_ptr = d->b_in_ptr.p; dev.arc_ts = d->b_arc_ts.p; dev.root_slot = d->b_root_slot.p;
    dev.term = d->d_term.p; dev.nlayers = o.nlayers; dev.nroot = start; dev.nnodes = o.nnodes; dev.nlast = o.nlast; dev.max_width = o.max_width;
    dev.arc_dead = d->d_arc_dead.p;
    d->dev = dev;
    d->narcs = o.narcs; d->nlayers_dev = o.nlayers; d->exact_flag = o.exact != 0; d->exact_layer_dev = o.exact_layer;
    d->device_built = true; d->host_built = false; d->uploaded = true; d->term_dirty = false;
    d->dev_ahead = false; d->dev_states_valid = true;     // fresh states: every walk takes the first in-arc, as on the host
    d->last_coef.assign(T, 0.0);
    d->has_cutset = restricted && !d->exact_flag;
    d->compile_cutset.clear();
    if (cutset_nodes) *cutset_nodes = d->has_cutset ? d->layer_width[o.exact_layer] : -1;
    return 1;
}

// the exact cut-set of a restricted tree that was built on the device, without building the host mirror
int device_cutset(sgufp_dd *d) {
    sgufp_ctx *c = d->ctx;
    Scratch &S = scratch_of(c);
    const int el = d->exact_layer_dev, count = d->layer_width[el], start = d->root_spec.global_layer;
    DevBuf<unsigned> om; DevBuf<int16_t> od;
    CU(c, cudaSetDevice(c->device));
    CU(c, om.reserve(count)); CU(c, od.reserve((size_t)count * std::max(1, el)));
    std::vector<unsigned> masks(count);
    std::vector<int16_t> decs((size_t)count * std::max(1, el));
    cudaError_t e = k2_cutset_launch(d->b_layer_info.p, d->b_in_ptr.p, d->b_arc_ts.p, d->d_arc_dec.p, d->b_mask.p, el, count, om.p, od.p, c->st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(masks.data(), om.p, (size_t)count * 4, cudaMemcpyDeviceToHost, c->st);
    if (e == cudaSuccess && el) e = cudaMemcpyAsync(decs.data(), od.p, (size_t)count * el * 2, cudaMemcpyDeviceToHost, c->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->st);
    om.release(); od.release();
    CU(c, e);
    const int g = start + el, tp = S.h_tab_ptr[S.h_lay_tab[g]];
    d->compile_cutset.assign(count, NodeSpec());
    for (int j = 0; j < count; j++) {
        NodeSpec &n = d->compile_cutset[j];
        n.global_layer = g;
        for (unsigned m = masks[j]; m; m &= m - 1) n.states.push_back(S.h_tab_dec[tp + __builtin_ctz(m)]);
        n.solution = d->root_spec.solution;
        n.solution.insert(n.solution.end(), decs.begin() + (size_t)j * el, decs.begin() + (size_t)(j + 1) * el);
    }
    return 0;
}

// RelaxedDDNew::getCutset on a diagram that was built on the device and whose states are there: no host mirror.
// Returns 1 and fills `out`, or 0 if this diagram is not in that situation.
int device_relaxed_cutset(sgufp_dd *d, std::vector<NodeSpec> &out) {
    sgufp_ctx *c = d->ctx;
    if (!d->device_built || !d->dev_states_valid || d->dd->restricted()) return 0;
    Scratch &S = scratch_of(c);
    CU(c, cudaSetDevice(c->device));
    const K2Apply a = make_apply(d, 0, 0.0);
    DevBuf<int> d4; DevBuf<unsigned> om; DevBuf<int32_t> ol; DevBuf<int16_t> od;
    CU(c, d4.reserve(4));
    int h4[4] = {-1, -1, 0, 0};
    cudaError_t e = k2_first_collapsed_launch(a, d4.p, c->st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h4, d4.p, 16, cudaMemcpyDeviceToHost, c->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->st);
    d4.release();
    CU(c, e);
    out.clear();
    const int layer = h4[0], count = h4[3] - h4[2];
    if (layer < 0 || count <= 0) return 1;                       // no collapsed layer: the reference returns nothing useful either
    const int stride = layer;                                    // at most one decision per layer 1..layer
    std::vector<unsigned> masks(count); std::vector<int32_t> lens(count); std::vector<int16_t> decs((size_t)count * stride);
    e = om.reserve(count);
    if (e == cudaSuccess) e = ol.reserve(count);
    if (e == cudaSuccess) e = od.reserve((size_t)count * stride);
    if (e == cudaSuccess) e = k2_relaxed_cutset_launch(a, d->b_mask.p, layer, h4[1], h4[2], count, stride, om.p, ol.p, od.p, c->st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(masks.data(), om.p, (size_t)count * 4, cudaMemcpyDeviceToHost, c->st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(lens.data(), ol.p, (size_t)count * 4, cudaMemcpyDeviceToHost, c->st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(decs.data(), od.p, (size_t)count * stride * 2, cudaMemcpyDeviceToHost, c->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->st);
    om.release(); ol.release(); od.release();
    CU(c, e);
    const Model &M = c->M;
    const int start = d->root_spec.global_layer, gl = start + layer;
    const bool fresh = gl < M.L && S.h_lay_first[gl];            // hasStateChanged[gl] (DD.cpp:4189)
    const int tp_tail = S.h_tab_ptr[S.h_lay_tab[gl - 1]];        // the tail's states are over the list of ITS layer
    for (int j = 0; j < count; j++) {
        if (lens[j] < 0) continue;                               // arc removed by a cut
        NodeSpec n;
        n.global_layer = gl;
        const int16_t dec = decs[(size_t)j * stride];
        n.solution = d->root_spec.solution;
        for (int k = lens[j] - 1; k >= 0; k--) n.solution.push_back(decs[(size_t)j * stride + k]);
        if (fresh) {
            const int tp = S.h_tab_ptr[S.h_lay_tab[gl]], nst = S.h_tab_ptr[S.h_lay_tab[gl] + 1] - tp;
            for (int p = 0; p < nst; p++) n.states.push_back(S.h_tab_dec[tp + p]);
        } else {
            for (unsigned m = masks[j]; m; m &= m - 1) {
                const int16_t s = S.h_tab_dec[tp_tail + __builtin_ctz(m)];
                if (dec == -1 || s != dec) n.states.push_back(s);
            }
        }
        out.push_back(std::move(n));
    }
    return 1;
}

// densify C cuts, run K2 over B diagrams, read back what the host semantics need
int run_k2(sgufp_ctx *c, sgufp_dd **dds, int B, const std::vector<std::vector<double>> &coefs, const double *rhs, int C,
           std::vector<std::vector<double>> *states_last_cut /* per diagram, states of cut C-1, or null */, double *bound,
           bool update_terminal) {
    const int T = std::max(1, c->M.T), Tpad = (T + 1) & ~1;
    Scratch &S = scratch_of(c);
    std::vector<K2DD> hd(B);
    long long off = 0, loff = 0, arcs = 0;
    int maxw = 1, maxl = 1, avgw = 1;
    for (int b = 0; b < B; b++) {
        if (dds[b]->ctx != c) return fail(c, SGUFP_ERR_ARG, "all diagrams of a batch must belong to one context");
        if (int rc = upload(dds[b])) return rc;
        maxw = std::max(maxw, dds[b]->dev.max_width);
        maxl = std::max(maxl, dds[b]->dev.nlayers);
        avgw = std::max(avgw, dds[b]->dev.nnodes / std::max(1, dds[b]->dev.nlayers));
    }
    const bool in_smem = k2_states_in_smem(Tpad, maxw, avgw);   // else: one global state block per (diagram, cut)
    for (int b = 0; b < B; b++) {
        hd[b] = dds[b]->dev;
        hd[b].state_off = off; hd[b].last_off = loff;
        off += (long long)hd[b].nnodes * (in_smem ? 1 : C);
        loff += (long long)hd[b].nlast * C;
        arcs += ((long long)dds[b]->narcs + dds[b]->dev.nlast) * C;
    }
    std::vector<double> cf((size_t)C * Tpad, 0.0);
    for (int k = 0; k < C; k++) std::copy(coefs[k].begin(), coefs[k].begin() + T, cf.begin() + (size_t)k * Tpad);
    CU(c, S.coef.reserve(cf.size())); CU(c, S.rhs.reserve(C)); CU(c, S.states.reserve((size_t)off)); CU(c, S.last.reserve((size_t)loff)); CU(c, S.bound.reserve(B)); CU(c, S.dds.reserve(B));
    CU(c, cudaMemcpyAsync(S.coef.p, cf.data(), cf.size() * 8, cudaMemcpyHostToDevice, c->st));
    CU(c, cudaMemcpyAsync(S.rhs.p, rhs, (size_t)C * 8, cudaMemcpyHostToDevice, c->st));
    CU(c, cudaMemcpyAsync(S.dds.p, hd.data(), (size_t)B * sizeof(K2DD), cudaMemcpyHostToDevice, c->st));
    c->dd_launches = 0;
    CU(c, cudaEventRecord(c->evk0, c->st));
    CU(c, k2_launch(S.dds.p, B, S.coef.p, S.rhs.p, C, Tpad, S.states.p, S.last.p, maxw, maxl, avgw, c->st, &c->dd_launches));
    if (update_terminal) {
        int max_last = 1;
        for (int b = 0; b < B; b++) max_last = std::max(max_last, hd[b].nlast);
        CU(c, S.wbounds.reserve((size_t)B * 64));
        CU(c, k2_terminal_launch(S.dds.p, B, C, S.last.p, S.bound.p, S.wbounds.p, max_last, c->st, &c->dd_launches));
    }
    CU(c, cudaEventRecord(c->evk1, c->st));
    if (states_last_cut) {
        states_last_cut->resize(B);
        for (int b = 0; b < B; b++) {
            (*states_last_cut)[b].resize(hd[b].nnodes);
            CU(c, cudaMemcpyAsync((*states_last_cut)[b].data(), S.states.p + hd[b].state_off + (in_smem ? 0 : (size_t)(C - 1) * hd[b].nnodes),
                                  (size_t)hd[b].nnodes * 8, cudaMemcpyDeviceToHost, c->st));
        }
    }
    if (bound && update_terminal) CU(c, cudaMemcpyAsync(bound, S.bound.p, (size_t)B * 8, cudaMemcpyDeviceToHost, c->st));
    CU(c, cudaStreamSynchronize(c->st));
    CU(c, cudaEventElapsedTime(&c->dd_kernel_ms, c->evk0, c->evk1));
    c->dd_arcs = arcs;
    return 0;
}

int write_nodes(sgufp_ctx *c, const std::vector<NodeSpec> &v, int32_t *words, int capacity) {
    int k = 0;
    for (const NodeSpec &n : v) {
        const int need = 3 + (int)n.states.size() + (int)n.solution.size();
        if (k + need > capacity) return fail(c, SGUFP_ERR_ARG, "cut-set buffer too small: " + std::to_string(v.size()) + " nodes do not fit " + std::to_string(capacity) + " words");
        words[k++] = n.global_layer;
        words[k++] = (int32_t)n.states.size();
        for (int16_t s : n.states) words[k++] = s;
        words[k++] = (int32_t)n.solution.size();
        for (int16_t s : n.solution) words[k++] = s;
    }
    return k;
}

}  // namespace

extern "C" {

int sgufp_dd_create(sgufp_ctx *ctx, int kind, int max_width, sgufp_dd **out) {
    if (!ctx || !out) return SGUFP_ERR_ARG;
    if (kind != SGUFP_DD_RELAXED && kind != SGUFP_DD_RESTRICTED) return fail(ctx, SGUFP_ERR_ARG, "kind must be SGUFP_DD_RELAXED or SGUFP_DD_RESTRICTED");
    if (kind == SGUFP_DD_RESTRICTED && max_width < 1) return fail(ctx, SGUFP_ERR_ARG, "restricted DD needs max_width >= 1");
    sgufp_dd *d = new sgufp_dd();
    d->ctx = ctx;
    d->dd = new HostDD(&ctx->M, kind == SGUFP_DD_RESTRICTED, max_width);
    NodeSpec root;
    d->dd->build(root, nullptr);
    *out = d;
    return 0;
}

void sgufp_dd_destroy(sgufp_dd *d) {
    if (!d) return;
    d->d_i32.release(); d->d_term.release();
    d->d_state.release(); d->d_last.release(); d->d_coef.release(); d->d_rhs.release(); d->d_arc_dead.release(); d->d_node_dead.release();
    d->d_lost.release(); d->d_layer_alive.release(); d->d_cnt.release(); d->d_arc_dec.release(); d->d_path.release(); d->d_self.release();
    d->d_res.release();
    d->b_layer_info.release(); d->b_arc_ts.release(); d->b_in_ptr.release(); d->b_off.release(); d->b_widths.release(); d->b_root_slot.release();
    d->b_mask.release(); d->b_out.release();
    delete d->dd;
    delete d;
}

int sgufp_dd_build(sgufp_dd *d, const int16_t *states, int ns, const int16_t *sol, int nsol, int global_layer, int *cutset_nodes) {
    if (!d || ns < 0 || nsol < 0 || (ns && !states) || (nsol && !sol)) return SGUFP_ERR_ARG;
    if (global_layer < 0 || global_layer > d->ctx->M.L) return fail(d->ctx, SGUFP_ERR_ARG, "globalLayer out of range");
    NodeSpec root;
    root.states.assign(states, states + ns);
    root.solution.assign(sol, sol + nsol);
    root.global_layer = global_layer;
    d->root_spec = root;
    d->root_solution = root.solution;
    d->dev_ahead = false; d->dev_states_valid = false;   // a new tree: whatever the device held is void
    d->cached_path_valid = false;
    const int built = device_build(d, root, cutset_nodes);
    if (built < 0) return built;
    if (built) return 0;
    d->dd->build(root, &d->compile_cutset);
    d->host_built = true; d->device_built = false;
    d->exact_flag = d->dd->is_exact();
    d->has_cutset = d->dd->restricted() && !d->dd->is_exact();
    d->uploaded = false; d->term_dirty = true;
    if (cutset_nodes) *cutset_nodes = d->has_cutset ? (int)d->compile_cutset.size() : -1;
    return 0;
}

int sgufp_dd_is_exact(const sgufp_dd *d) { return d ? (d->exact_flag ? 1 : 0) : SGUFP_ERR_ARG; }
int sgufp_dd_num_layers(const sgufp_dd *d) { return d ? (d->host_built ? (int)d->dd->tree().size() : d->nlayers_dev) : SGUFP_ERR_ARG; }
int sgufp_dd_layer_sizes(sgufp_dd *d, int32_t *sizes) {
    if (!d || !sizes) return SGUFP_ERR_ARG;
    if (int rc = sync_host(d)) return rc;
    for (size_t l = 0; l < d->dd->tree().size(); l++) sizes[l] = (int32_t)d->dd->tree()[l].size();
    return 0;
}
int sgufp_dd_counts(sgufp_dd *d, int64_t *nodes, int64_t *arcs) {
    if (!d) return SGUFP_ERR_ARG;
    if (int rc = sync_host(d)) return rc;
    long n = 0;
    for (auto &l : d->dd->tree()) n += (long)l.size();
    if (nodes) *nodes = n;
    if (arcs) *arcs = d->dd->count_arcs();
    return 0;
}

int sgufp_dd_dump(sgufp_dd *d, int32_t *node_layer, double *node_state, int64_t *node_inptr, int32_t *arc_tailpos, int32_t *arc_decision,
                  double *terminal_weight) {
    if (!d) return SGUFP_ERR_ARG;
    if (int rc = sync_host(d)) return rc;
    const DDCsr &C = d->dd->flatten();
    for (int l = 0; l < C.nlayers; l++)
        for (int v = C.layer_ptr[l]; v < C.layer_ptr[l + 1]; v++) {
            if (node_layer) node_layer[v] = l;
            if (node_state) node_state[v] = d->dd->nodes()[C.node_id[v]].state2;
            if (node_inptr) node_inptr[v] = C.in_ptr[v];
            for (int e = C.in_ptr[v]; e < C.in_ptr[v + 1]; e++) {
                if (arc_tailpos) arc_tailpos[e] = C.arc_tail[e] - C.layer_ptr[l - 1];
                if (arc_decision) arc_decision[e] = d->dd->arcs()[C.arc_id[e]].decision;
            }
        }
    if (node_inptr) node_inptr[C.nnodes] = C.narcs;
    if (terminal_weight) std::copy(d->dd->terminal_weights().begin(), d->dd->terminal_weights().end(), terminal_weight);
    return 0;
}

int sgufp_dd_dump_device(sgufp_dd *d, int32_t *layer_sizes, int64_t *node_inptr, int32_t *arc_tailpos, int32_t *arc_decision, int32_t *arc_slot) {
    if (!d) return SGUFP_ERR_ARG;
    sgufp_ctx *c = d->ctx;
    if (int rc = upload(d)) return rc;
    CU(c, cudaSetDevice(c->device));
    const int nl = d->dev.nlayers, nn = d->dev.nnodes, na = d->narcs;
    std::vector<int4> li(nl);
    std::vector<int32_t> ip(nn + 1), dec(std::max(1, na));
    std::vector<int2> ts(std::max(1, na));
    CU(c, cudaMemcpy(li.data(), d->dev.layer_info, (size_t)nl * sizeof(int4), cudaMemcpyDeviceToHost));
    CU(c, cudaMemcpy(ip.data(), d->dev.in_ptr, (size_t)(nn + 1) * 4, cudaMemcpyDeviceToHost));
    if (na) CU(c, cudaMemcpy(ts.data(), d->dev.arc_ts, (size_t)na * sizeof(int2), cudaMemcpyDeviceToHost));
    if (na) CU(c, cudaMemcpy(dec.data(), d->d_arc_dec.p, (size_t)na * 4, cudaMemcpyDeviceToHost));
    for (int l = 0; l < nl; l++) if (layer_sizes) layer_sizes[l] = li[l].z;
    for (int v = 0; v <= nn; v++) if (node_inptr) node_inptr[v] = ip[v];
    for (int e = 0; e < na; e++) {
        if (arc_tailpos) arc_tailpos[e] = ts[e].x;
        if (arc_slot) arc_slot[e] = ts[e].y;
        if (arc_decision) arc_decision[e] = dec[e];
    }
    return d->device_built ? 1 : 0;
}

int sgufp_dd_apply_optimality(sgufp_dd *d, double rhs, const uint64_t *keys, const double *vals, int nnz, double optimal, double ub,
                              double *bound) {
    if (!d || nnz < 0 || (nnz && (!keys || !vals))) return SGUFP_ERR_ARG;
    (void)ub;
    K2Result r{};
    if (int rc = apply_on_device(d, rhs, keys, vals, nnz, 0, optimal, r)) return rc;
    if (bound) *bound = r.bound;
    return 0;
}

int sgufp_dd_apply_feasibility(sgufp_dd *d, double rhs, const uint64_t *keys, const double *vals, int nnz, int *feasible) {
    if (!d || nnz < 0 || (nnz && (!keys || !vals))) return SGUFP_ERR_ARG;
    K2Result r{};
    if (int rc = apply_on_device(d, rhs, keys, vals, nnz, 1, 0.0, r)) return rc;
    if (feasible) *feasible = r.feasible;
    return 0;
}

int sgufp_dd_apply_sequence(sgufp_dd *d, int mode, const double *rhs, const uint64_t *keys, const double *vals, const int32_t *cut_ptr, int C,
                            double optimal, double *bound, int *feasible, int *applied) {
    if (!d || C < 1 || !rhs || !cut_ptr || (mode != 0 && mode != 1)) return SGUFP_ERR_ARG;
    sgufp_ctx *c = d->ctx;
    if (int rc = upload(d)) return rc;
    Scratch &S = scratch_of(c);
    const int T = std::max(1, c->M.T), Tpad = (T + 1) & ~1, nn = d->dev.nnodes, nlast = d->dev.nlast;
    std::vector<double> cf((size_t)C * Tpad, 0.0), one;
    for (int k = 0; k < C; k++) {
        d->dd->densify(keys + cut_ptr[k], vals + cut_ptr[k], cut_ptr[k + 1] - cut_ptr[k], one);
        std::copy(one.begin(), one.begin() + T, cf.begin() + (size_t)k * Tpad);
    }
    // widest window: 256 cuts, and at most 1 GB of node states (the scratch is kept between calls, it only grows)
    const int chunk = (int)std::max<long long>(1, std::min<long long>(std::min(C, 256), (1LL << 27) / std::max(1, nn)));
    DevBuf<K2Result> &res = S.seq_res; DevBuf<int> &ctl = S.seq_ctl, &probe = S.seq_probe; DevBuf<double> &wb = S.wbounds;
    const bool probing = mode == 0 && !d->exact_flag && !d->dd->restricted();   // only there does an optimality cut prune arcs
    CU(c, S.coef.reserve(cf.size())); CU(c, S.rhs.reserve(C)); CU(c, S.states.reserve((size_t)chunk * nn)); CU(c, S.last.reserve((size_t)chunk * std::max(1, nlast)));
    CU(c, res.reserve(C)); CU(c, ctl.reserve(2)); CU(c, probe.reserve(chunk)); CU(c, wb.reserve(chunk));
    CU(c, d->d_coef.reserve(Tpad));   // a host-built diagram (SGUFP_DD_BUILD=host, the fallbacks) has none before its first single-cut apply
    K2DD self = d->dev; self.state_off = 0; self.last_off = 0;
    cudaError_t e = cudaMemcpyAsync(S.coef.p, cf.data(), cf.size() * 8, cudaMemcpyHostToDevice, c->st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(S.rhs.p, rhs, (size_t)C * 8, cudaMemcpyHostToDevice, c->st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d->d_self.p, &self, sizeof(K2DD), cudaMemcpyHostToDevice, c->st);
    c->dd_launches = 0;
    if (e == cudaSuccess) e = cudaEventRecord(c->evk0, c->st);
    // Speculation window: the longest paths of `win` cuts are computed side by side on the current structure; a cut
    // that changes it voids the ones behind it.  The window doubles after a clean run and restarts small after a change.
    int k0 = 0, stopped = 0, last_base = 0, win = std::min(chunk, 8);
    while (e == cudaSuccess && k0 < C && !stopped) {
        K2Seq q{};
        q.coef = S.coef.p; q.states = S.states.p; q.results = res.p; q.ctl = ctl.p; q.k0 = k0; q.k1 = std::min(C, k0 + win); q.Tpad = Tpad; q.probe = probing ? probe.p : nullptr; q.last = S.last.p; q.bounds = mode == 0 ? wb.p : nullptr;
        e = k2_sequence_launch(d->d_self.p, make_apply(d, mode, optimal), q, S.rhs.p, d->dev.max_width, S.last.p, c->st, &c->dd_launches);
        int h[2] = {0, 0};
        if (e == cudaSuccess) e = cudaMemcpyAsync(h, ctl.p, 8, cudaMemcpyDeviceToHost, c->st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->st);
        last_base = k0;
        win = h[0] == q.k1 ? std::min(chunk, win * 2) : std::max(4, std::min(chunk, (h[0] - k0) * 2));
        k0 = h[0]; stopped = h[1];
    }
    const int done = k0;   // cuts applied, the stopping one included
    std::vector<K2Result> hr(std::max(1, done));
    if (e == cudaSuccess && done > 0) {
        // the diagram keeps the states and coefficients of the last cut applied (getSolution, the host mirror)
        e = cudaMemcpyAsync(d->d_state.p, S.states.p + (size_t)(done - 1 - last_base) * nn, (size_t)nn * 8, cudaMemcpyDeviceToDevice, c->st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d->d_coef.p, S.coef.p + (size_t)(done - 1) * Tpad, (size_t)Tpad * 8, cudaMemcpyDeviceToDevice, c->st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(hr.data(), res.p, (size_t)done * sizeof(K2Result), cudaMemcpyDeviceToHost, c->st);
    }
    if (e == cudaSuccess) e = cudaEventRecord(c->evk1, c->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->st);
    CU(c, e);
    CU(c, cudaEventElapsedTime(&c->dd_kernel_ms, c->evk0, c->evk1));
    c->dd_arcs = ((long long)d->narcs + nlast) * done;
    for (int k = 0; k < done; k++) { if (bound) bound[k] = hr[k].bound; if (feasible) feasible[k] = hr[k].feasible; }
    if (applied) *applied = done;
    if (done > 0) {
        d->last_coef.assign(cf.begin() + (size_t)(done - 1) * Tpad, cf.begin() + (size_t)(done - 1) * Tpad + T);
        d->dev_ahead = true; d->dev_states_valid = true; d->term_dirty = false;
        d->cached_path_valid = false;
    }
    return 0;
}

int sgufp_dd_apply_optimality_batch(sgufp_dd **dds, int B, const double *rhs, const uint64_t *keys, const double *vals, const int32_t *cut_ptr,
                                    int C, double *bound) {
    if (!dds || B < 1 || C < 1 || !rhs || !cut_ptr || !dds[0]) return SGUFP_ERR_ARG;
    sgufp_ctx *c = dds[0]->ctx;
    for (int b = 0; b < B; b++) { if (!dds[b]) return SGUFP_ERR_ARG; if (int rc = sync_host(dds[b])) return rc; }
    std::vector<std::vector<double>> coef(C), states;
    for (int k = 0; k < C; k++) dds[0]->dd->densify(keys + cut_ptr[k], vals + cut_ptr[k], cut_ptr[k + 1] - cut_ptr[k], coef[k]);
    std::vector<double> bnd(B);
    if (int rc = run_k2(c, dds, B, coef, rhs, C, &states, bnd.data(), true)) return rc;
    // bring the host mirrors up to date: node states / arc weights of the last cut, terminal weights from the device
    for (int b = 0; b < B; b++) {
        std::vector<double> &t = dds[b]->dd->terminal_weights();
        if (!t.empty()) CU(c, cudaMemcpy(t.data(), dds[b]->dev.term, t.size() * 8, cudaMemcpyDeviceToHost));
        std::vector<double> keep = t;
        dds[b]->dd->finish_optimality(coef[C - 1], states[b], DD_MAX, DD_MAX);   // optimal = +max: no pruning, early return
        t = keep;   // finish_optimality folded cut C-1 once more; min is idempotent, keep the device values anyway
        dds[b]->term_dirty = false;
        dds[b]->dev_states_valid = false;   // the batch keeps its states in the shared scratch: paths come from the host mirror
        dds[b]->cached_path_valid = false;
    }
    if (bound) std::copy(bnd.begin(), bnd.end(), bound);
    return 0;
}

int sgufp_dd_solution(sgufp_dd *d, int16_t *path, int capacity) {
    if (!d || !path) return SGUFP_ERR_ARG;
    std::vector<int16_t> p;
    if (d->cached_path_valid) p = d->cached_path;
    else if (d->dev_states_valid) {
        // getSolution on the device image: only the path comes back
        sgufp_ctx *c = d->ctx;
        CU(c, cudaSetDevice(c->device));
        K2Result r{};
        std::vector<int16_t> rev(d->dev.nlayers);
        CU(c, k2_extract_launch(make_apply(d, 0, 0.0), c->st, &c->dd_launches));
        CU(c, cudaMemcpyAsync(&r, d->d_res.p, sizeof(K2Result), cudaMemcpyDeviceToHost, c->st));
        CU(c, cudaMemcpyAsync(rev.data(), d->d_path.p, (size_t)d->dev.nlayers * 2, cudaMemcpyDeviceToHost, c->st));
        CU(c, cudaStreamSynchronize(c->st));
        p = d->root_solution;
        for (int i = r.path_len - 1; i >= 0; i--) p.push_back(rev[i]);
    } else { ensure_host(d); p = d->dd->solution(); }
    if ((int)p.size() > capacity) return SGUFP_ERR_ARG;
    std::copy(p.begin(), p.end(), path);
    return (int)p.size();
}

int sgufp_dd_cutset(sgufp_dd *d, double ub, int32_t *words, int capacity) {
    if (!d || !words) return SGUFP_ERR_ARG;
    if (d->dd->restricted()) {
        // the cut-set of the last compile: structure only, cuts applied since do not change it
        if (!d->has_cutset) return 0;
        if (d->compile_cutset.empty() && d->device_built && !d->host_built) { if (int rc = device_cutset(d)) return rc; }
        else ensure_host(d);
        return write_nodes(d->ctx, d->compile_cutset, words, capacity);
    }
    {
        std::vector<NodeSpec> cs;
        const int got = device_relaxed_cutset(d, cs);
        if (got < 0) return got;
        if (got) return write_nodes(d->ctx, cs, words, capacity);
    }
    if (int rc = sync_host(d)) return rc;
    return write_nodes(d->ctx, d->dd->cutset(ub), words, capacity);
}

int sgufp_dd_last_stats(const sgufp_dd *d, float *kernel_ms, int64_t *arcs_touched, int *kernel_launches) {
    if (!d) return SGUFP_ERR_ARG;
    if (kernel_ms) *kernel_ms = d->ctx->dd_kernel_ms;
    if (arcs_touched) *arcs_touched = d->ctx->dd_arcs;
    if (kernel_launches) *kernel_launches = d->ctx->dd_launches;
    return 0;
}

}  // extern "C"
