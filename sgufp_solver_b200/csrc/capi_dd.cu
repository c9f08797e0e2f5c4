// capi_dd.cu — C ABI of include/sgufp_b200_dd.h: host diagrams + the K2 device pass.
#include <cuda_runtime.h>

#include <algorithm>
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
        d->uploaded = true;
        d->uploaded_version = d->dd->version();
        d->term_dirty = true;
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
        arcs += (long long)dds[b]->dd->count_arcs() * C;
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
    if (cutset_nodes) *cutset_nodes = d->has_cutset ? (int)d->compile_cutset.size() : -1;
    return 0;
}

int sgufp_dd_is_exact(const sgufp_dd *d) { return d ? (d->dd->is_exact() ? 1 : 0) : SGUFP_ERR_ARG; }
int sgufp_dd_num_layers(const sgufp_dd *d) { return d ? (int)d->dd->tree().size() : SGUFP_ERR_ARG; }
int sgufp_dd_layer_sizes(const sgufp_dd *d, int32_t *sizes) {
    if (!d || !sizes) return SGUFP_ERR_ARG;
    for (size_t l = 0; l < d->dd->tree().size(); l++) sizes[l] = (int32_t)d->dd->tree()[l].size();
    return 0;
}
int sgufp_dd_counts(const sgufp_dd *d, int64_t *nodes, int64_t *arcs) {
    if (!d) return SGUFP_ERR_ARG;
    long n = 0;
    for (auto &l : d->dd->tree()) n += (long)l.size();
    if (nodes) *nodes = n;
    if (arcs) *arcs = d->dd->count_arcs();
    return 0;
}

int sgufp_dd_dump(sgufp_dd *d, int32_t *node_layer, double *node_state, int64_t *node_inptr, int32_t *arc_tailpos, int32_t *arc_decision,
                  double *terminal_weight) {
    if (!d) return SGUFP_ERR_ARG;
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
    std::vector<std::vector<double>> coef(1), states;
    d->dd->densify(keys, vals, nnz, coef[0]);
    if (int rc = run_k2(d->ctx, &d, 1, coef, &rhs, 1, &states, nullptr, false)) return rc;
    const double b = d->dd->finish_optimality(coef[0], states[0], optimal, ub);
    d->term_dirty = true;   // the host holds the authoritative terminal weights after a single-cut call
    if (bound) *bound = b;
    return 0;
}

int sgufp_dd_apply_feasibility(sgufp_dd *d, double rhs, const uint64_t *keys, const double *vals, int nnz, int *feasible) {
    if (!d || nnz < 0 || (nnz && (!keys || !vals))) return SGUFP_ERR_ARG;
    std::vector<std::vector<double>> coef(1), states;
    d->dd->densify(keys, vals, nnz, coef[0]);
    if (int rc = run_k2(d->ctx, &d, 1, coef, &rhs, 1, &states, nullptr, false)) return rc;
    const int f = d->dd->finish_feasibility(coef[0], states[0]);
    d->term_dirty = true;
    if (feasible) *feasible = f;
    return 0;
}

int sgufp_dd_apply_optimality_batch(sgufp_dd **dds, int B, const double *rhs, const uint64_t *keys, const double *vals, const int32_t *cut_ptr,
                                    int C, double *bound) {
    if (!dds || B < 1 || C < 1 || !rhs || !cut_ptr || !dds[0]) return SGUFP_ERR_ARG;
    sgufp_ctx *c = dds[0]->ctx;
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
    }
    if (bound) std::copy(bnd.begin(), bnd.end(), bound);
    return 0;
}

int sgufp_dd_solution(const sgufp_dd *d, int16_t *path, int capacity) {
    if (!d || !path) return SGUFP_ERR_ARG;
    const std::vector<int16_t> p = d->dd->solution();
    if ((int)p.size() > capacity) return SGUFP_ERR_ARG;
    std::copy(p.begin(), p.end(), path);
    return (int)p.size();
}

int sgufp_dd_cutset(const sgufp_dd *d, double ub, int32_t *words, int capacity) {
    if (!d || !words) return SGUFP_ERR_ARG;
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
