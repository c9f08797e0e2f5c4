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
};

namespace {

int upload(sgufp_dd *d) {
    sgufp_ctx *c = d->ctx;
    if (c->device == SGUFP_DEVICE_NONE) return fail(c, SGUFP_ERR_CUDA, "handle was created with SGUFP_DEVICE_NONE: there is no CPU compute path for K2");
    CU(c, cudaSetDevice(c->device));
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

// bring the host mirror up to date with what the device did since the last upload
int sync_host(sgufp_dd *d) {
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

// diagrams at least this wide run the single-cut longest path as one launch per layer over the whole GPU
int layered_min_width() {
    const char *e = getenv("SGUFP_K2_LAYERED_MIN");   // read per call: the tests force either path
    return e ? std::max(1, atoi(e)) : 2048;
}

K2Apply make_apply(sgufp_dd *d, int mode, double optimal) {
    K2Apply a{};
    a.d = d->dev; a.d.state_off = 0; a.d.last_off = 0;
    a.state = d->d_state.p; a.coef = d->d_coef.p; a.arc_dec = d->d_arc_dec.p;
    a.arc_dead = d->d_arc_dead.p; a.node_dead = d->d_node_dead.p; a.layer_alive = d->d_layer_alive.p;
    a.cnt = d->d_cnt.p; a.lost = d->d_lost.p; a.out = d->d_res.p; a.path = d->d_path.p;
    a.optimal = optimal; a.mode = mode; a.restricted = d->dd->restricted() ? 1 : 0; a.exact = d->dd->is_exact() ? 1 : 0;
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
    if (d->dev.max_width >= layered_min_width()) {
        // wide diagram: the root state on the host (same additions in the same order), then one launch per layer
        double v = rhs;
        for (int s : d->root_slot) if (s >= 0) v = v + coef[s];
        CU(c, cudaMemcpyAsync(d->d_state.p, &v, 8, cudaMemcpyHostToDevice, c->st));
        CU(c, k2_layered_launch(self, d->layer_width.data(), d->layer_collapsed.data(), d->d_coef.p, d->d_state.p, c->st, &c->dd_launches));
    } else
        CU(c, k2_single_launch(d->d_self.p, d->d_coef.p, d->d_rhs.p, Tpad, d->d_state.p, d->d_last.p, d->dev.max_width, c->st, &c->dd_launches));
    CU(c, k2_finish_launch(make_apply(d, mode, optimal), c->st, &c->dd_launches));
    CU(c, cudaEventRecord(c->evk1, c->st));
    CU(c, cudaMemcpyAsync(&res, d->d_res.p, sizeof(K2Result), cudaMemcpyDeviceToHost, c->st));
    CU(c, cudaStreamSynchronize(c->st));   // also covers the pageable `coef`, `rhs`, `self`
    CU(c, cudaEventElapsedTime(&c->dd_kernel_ms, c->evk0, c->evk1));
    c->dd_arcs = (long long)d->narcs + d->dev.nlast;   // in-arcs + terminal arcs of the uploaded image
    d->dev_ahead = true; d->dev_states_valid = true;
    d->term_dirty = false;                  // the device holds the authoritative terminal weights now
    return 0;
}

struct Scratch {   // per-context scratch for the batch call
    DevBuf<double> coef, rhs, states, last, bound;
    DevBuf<K2DD> dds;
};
Scratch &scratch_of(sgufp_ctx *c) {
    if (!c->dd_scratch) {
        c->dd_scratch = new Scratch();
        c->dd_scratch_free = [](void *p) {
            Scratch *s = static_cast<Scratch *>(p);
            s->coef.release(); s->rhs.release(); s->states.release(); s->last.release(); s->bound.release(); s->dds.release();
            delete s;
        };
    }
    return *static_cast<Scratch *>(c->dd_scratch);
}

// densify C cuts, run K2 over B diagrams, read back what the host semantics need
int run_k2(sgufp_ctx *c, sgufp_dd **dds, int B, const std::vector<std::vector<double>> &coefs, const double *rhs, int C,
           std::vector<std::vector<double>> *states_last_cut /* per diagram, states of cut C-1, or null */, double *bound,
           bool update_terminal) {
    const int T = std::max(1, c->M.T), Tpad = (T + 1) & ~1;
    Scratch &S = scratch_of(c);
    std::vector<K2DD> hd(B);
    long long off = 0, loff = 0, arcs = 0;
    int maxw = 1;
    for (int b = 0; b < B; b++) {
        if (dds[b]->ctx != c) return fail(c, SGUFP_ERR_ARG, "all diagrams of a batch must belong to one context");
        if (int rc = upload(dds[b])) return rc;
        maxw = std::max(maxw, dds[b]->dev.max_width);
    }
    const bool in_smem = k2_states_in_smem(Tpad, maxw);   // else: one global state block per (diagram, cut)
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
    CU(c, k2_launch(S.dds.p, B, S.coef.p, S.rhs.p, C, Tpad, S.states.p, S.last.p, maxw, c->st, &c->dd_launches));
    if (update_terminal) CU(c, k2_terminal_launch(S.dds.p, B, C, S.last.p, S.bound.p, c->st, &c->dd_launches));
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

int write_nodes(const std::vector<NodeSpec> &v, int32_t *words, int capacity) {
    int k = 0;
    for (const NodeSpec &n : v) {
        const int need = 3 + (int)n.states.size() + (int)n.solution.size();
        if (k + need > capacity) return SGUFP_ERR_ARG;
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
    d->dd->build(root, &d->compile_cutset);
    d->has_cutset = d->dd->restricted() && !d->dd->is_exact();
    d->uploaded = false; d->term_dirty = true;
    d->dev_ahead = false; d->dev_states_valid = false;   // a new tree: whatever the device held is void
    if (cutset_nodes) *cutset_nodes = d->has_cutset ? (int)d->compile_cutset.size() : -1;
    return 0;
}

int sgufp_dd_is_exact(const sgufp_dd *d) { return d ? (d->dd->is_exact() ? 1 : 0) : SGUFP_ERR_ARG; }
int sgufp_dd_num_layers(const sgufp_dd *d) { return d ? (int)d->dd->tree().size() : SGUFP_ERR_ARG; }
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
    }
    if (bound) std::copy(bnd.begin(), bnd.end(), bound);
    return 0;
}

int sgufp_dd_solution(sgufp_dd *d, int16_t *path, int capacity) {
    if (!d || !path) return SGUFP_ERR_ARG;
    std::vector<int16_t> p;
    if (d->dev_states_valid) {
        // getSolution on the device image: only the path comes back
        sgufp_ctx *c = d->ctx;
        CU(c, cudaSetDevice(c->device));
        K2Result r{};
        std::vector<int16_t> rev(d->dev.nlayers);
        CU(c, k2_extract_launch(make_apply(d, 0, 0.0), c->st, &c->dd_launches));
        CU(c, cudaMemcpyAsync(&r, d->d_res.p, sizeof(K2Result), cudaMemcpyDeviceToHost, c->st));
        CU(c, cudaMemcpyAsync(rev.data(), d->d_path.p, (size_t)d->dev.nlayers * 2, cudaMemcpyDeviceToHost, c->st));
        CU(c, cudaStreamSynchronize(c->st));
        p = d->dd->root_solution();
        for (int i = r.path_len - 1; i >= 0; i--) p.push_back(rev[i]);
    } else p = d->dd->solution();
    if ((int)p.size() > capacity) return SGUFP_ERR_ARG;
    std::copy(p.begin(), p.end(), path);
    return (int)p.size();
}

int sgufp_dd_cutset(sgufp_dd *d, double ub, int32_t *words, int capacity) {
    if (!d || !words) return SGUFP_ERR_ARG;
    if (int rc = sync_host(d)) return rc;
    if (d->dd->restricted()) return d->has_cutset ? write_nodes(d->compile_cutset, words, capacity) : 0;
    return write_nodes(d->dd->cutset(ub), words, capacity);
}

int sgufp_dd_last_stats(const sgufp_dd *d, float *kernel_ms, int64_t *arcs_touched, int *kernel_launches) {
    if (!d) return SGUFP_ERR_ARG;
    if (kernel_ms) *kernel_ms = d->ctx->dd_kernel_ms;
    if (arcs_touched) *arcs_touched = d->ctx->dd_arcs;
    if (kernel_launches) *kernel_launches = d->ctx->dd_launches;
    return 0;
}

}  // extern "C"
