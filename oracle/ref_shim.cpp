// ref_shim.cpp — ORACLE A.  TEST INFRASTRUCTURE ONLY.
//
// A thin extern "C" window onto the UNMODIFIED reference sources, which are compiled where they
// lie under /root/reference (oracle/Makefile) into oracle/_ref/libsgufp_ref.so:
//   Network            /root/reference/Network.h:69-117, Network.cpp:10-186
//   Inavap::Cut & co.  /root/reference/Cut.h:185-421
//   RelaxedDDNew       /root/reference/DD.h:734-810,  DD.cpp:3509-4229
//   RestrictedDDNew    /root/reference/DD.h:653-730,  DD.cpp:3090-3505
// No reference source is copied; this file only calls the reference's public members and reads
// its containers.  Only tests/, smoke() and bench.py's CPU-baseline legs may load the library.
#define private public   // RestrictedDDNew keeps its containers private (DD.h:655-681); read-only use
#include "DD.h"
#undef private
#include <cstring>
#include <memory>

using namespace Inavap;

namespace {
Inavap::Cut make_cut(double rhs, const uint64_t *keys, const double *vals, int nnz) {
    std::vector<std::pair<uint64_t, double>> c;
    c.reserve(nnz);
    for (int i = 0; i < nnz; i++) c.emplace_back(keys[i], vals[i]);
    return Inavap::Cut{rhs, c};
}
Inavap::Node make_node(const int16_t *states, int ns, const int16_t *sol, int nsol, int gl) {
    return Inavap::Node{std::vector<int16_t>(states, states + ns), std::vector<int16_t>(sol, sol + nsol),
                        DOUBLE_MIN, DOUBLE_MIN, static_cast<uint16_t>(gl)};
}
struct NetHolder { std::shared_ptr<Network> p; };
}  // namespace

extern "C" {

// ---- Network -----------------------------------------------------------------------------
void *ref_net_load(const char *file) { auto *h = new NetHolder{std::make_shared<Network>(std::string(file))}; return h; }
void ref_net_free(void *h) { delete static_cast<NetHolder *>(h); }
int ref_net_n(void *h) { return static_cast<NetHolder *>(h)->p->n; }
int ref_net_m(void *h) { return static_cast<NetHolder *>(h)->p->edges; }
int ref_net_total_layers(void *h) { return static_cast<NetHolder *>(h)->p->totalLayers; }
int ref_net_nvbar(void *h) { return static_cast<NetHolder *>(h)->p->Vbar.size(); }
void ref_net_vbar(void *h, int *out) { auto &v = static_cast<NetHolder *>(h)->p->Vbar; for (size_t i = 0; i < v.size(); i++) out[i] = v[i]; }
void ref_net_processing_order(void *h, int *layer, int *arc) {
    auto &po = static_cast<NetHolder *>(h)->p->processingOrder;
    for (size_t i = 0; i < po.size(); i++) { layer[i] = po[i].first; arc[i] = po[i].second; }
}
void ref_net_has_state_changed(void *h, int *out) { auto &v = static_cast<NetHolder *>(h)->p->hasStateChanged; for (size_t i = 0; i < v.size(); i++) out[i] = v[i]; }
int ref_net_nhsc(void *h) { return static_cast<NetHolder *>(h)->p->hasStateChanged.size(); }
// stateUpdateMap flattened: for layer key k, states (ascending, as std::set iterates)
int ref_net_state_update(void *h, int key, int *out) {
    auto &mp = static_cast<NetHolder *>(h)->p->stateUpdateMap;
    auto it = mp.find(key);
    if (it == mp.end()) return -1;
    int k = 0;
    for (int s : it->second) out[k++] = s;
    return k;
}
void ref_net_classes(void *h, int *na1, int *na2, int *na3, int *na4) {
    auto &n = *static_cast<NetHolder *>(h)->p; *na1 = n.A1.size(); *na2 = n.A2.size(); *na3 = n.A3.size(); *na4 = n.A4.size();
}

// ---- Inavap::Cut (Cut.h:201-337) and cutToCut (Cut.h:406-421) ------------------------------
// ::Cut given as (i,q,j,val) tuples -> Inavap::Cut pairs + hash.
int ref_cut_to_cut(int type, double rhs, const int *ci, const int *cq, const int *cj, const double *cv, int cnt,
                   uint64_t *keys, double *vals, uint64_t *hash) {
    CutCoefficients cc;
    for (int k = 0; k < cnt; k++) cc[std::make_tuple(ci[k], cq[k], cj[k])] = cv[k];
    ::Cut old{static_cast<CutType>(type), rhs, cc};
    Inavap::Cut c = cutToCut(old, nullptr);
    for (size_t k = 0; k < c.coeff.size(); k++) { keys[k] = c.coeff[k].first; vals[k] = c.coeff[k].second; }
    *hash = c.hash_val;
    return c.coeff.size();
}
uint64_t ref_cut_hash(double rhs, const uint64_t *keys, const double *vals, int nnz) { return make_cut(rhs, keys, vals, nnz).hash_val; }
double ref_cut_get(double rhs, const uint64_t *keys, const double *vals, int nnz, uint64_t key) { return make_cut(rhs, keys, vals, nnz).get(key); }
uint64_t ref_get_key(uint64_t q, uint64_t i, uint64_t j) { return getKey(q, i, j); }

// ---- RelaxedDDNew ----------------------------------------------------------------------
void *ref_rel_new(void *net) { return new RelaxedDDNew(static_cast<NetHolder *>(net)->p.get()); }
void ref_rel_free(void *d) { delete static_cast<RelaxedDDNew *>(d); }
void ref_rel_build(void *d, const int16_t *states, int ns, const int16_t *sol, int nsol, int gl) {
    static_cast<RelaxedDDNew *>(d)->buildTree(make_node(states, ns, sol, nsol, gl));
}
int ref_rel_is_exact(void *d) { return static_cast<RelaxedDDNew *>(d)->isTreeExact(); }
double ref_rel_apply_opt(void *d, double rhs, const uint64_t *keys, const double *vals, int nnz, double optimal, double ub) {
    return static_cast<RelaxedDDNew *>(d)->applyOptimalityCut(make_cut(rhs, keys, vals, nnz), optimal, ub);
}
int ref_rel_apply_feas(void *d, double rhs, const uint64_t *keys, const double *vals, int nnz) {
    return static_cast<RelaxedDDNew *>(d)->applyFeasibilityCut(make_cut(rhs, keys, vals, nnz));
}
int ref_rel_solution(void *d, int16_t *out) {
    auto p = static_cast<RelaxedDDNew *>(d)->getSolution();
    for (size_t i = 0; i < p.size(); i++) out[i] = p[i];
    return p.size();
}
int ref_rel_nlayers(void *d) { return static_cast<RelaxedDDNew *>(d)->tree.size(); }
void ref_rel_layer_sizes(void *d, int *out) { auto &t = static_cast<RelaxedDDNew *>(d)->tree; for (size_t i = 0; i < t.size(); i++) out[i] = t[i].size(); }
// Flatten the live structure layer by layer, nodes in tree order, in-arcs in stored order.
// node_inptr has (#nodes+1) entries; per in-arc: position of the tail node in the PREVIOUS layer and the decision.
// Returns the number of in-arcs written (terminal layer included).
long ref_rel_dump(void *dv, int *node_layer, double *node_state, long *node_inptr, int *arc_tailpos, int *arc_decision, double *arc_weight) {
    auto *d = static_cast<RelaxedDDNew *>(dv);
    long nn = 0, na = 0;
    std::unordered_map<uint, int> pos_prev, pos_cur;
    for (size_t l = 0; l < d->tree.size(); l++) {
        pos_cur.clear();
        int k = 0;
        for (uint id : d->tree[l]) {
            const auto &nd = d->nodes.at(id);
            pos_cur[id] = k++;
            node_layer[nn] = l; node_state[nn] = nd.state2; node_inptr[nn] = na;
            if (l > 0)
                for (uint aid : nd.incomingArcs) {
                    const auto &a = d->arcs.at(aid);
                    arc_tailpos[na] = pos_prev.at(a.tail); arc_decision[na] = a.decision; arc_weight[na] = a.weight; na++;
                }
            nn++;
        }
        pos_prev.swap(pos_cur);
    }
    node_inptr[nn] = na;
    return na;
}
long ref_rel_count_nodes(void *dv) { long c = 0; for (auto &l : static_cast<RelaxedDDNew *>(dv)->tree) c += l.size(); return c; }
long ref_rel_count_arcs(void *dv) {
    auto *d = static_cast<RelaxedDDNew *>(dv); long c = 0;
    for (size_t l = 1; l < d->tree.size(); l++) for (uint id : d->tree[l]) c += d->nodes.at(id).incomingArcs.size();
    return c;
}
// getCutset (DD.cpp:4179-4218): nodes flattened as (globalLayer, #states, states..., #sol, sol...)
int ref_rel_cutset(void *dv, double ub, int *buf, int cap) {
    auto v = static_cast<RelaxedDDNew *>(dv)->getCutset(ub);
    int k = 0;
    for (auto &nd : v) {
        if (k + 3 + (int)nd.states.size() + (int)nd.solutionVector.size() > cap) return -1;
        buf[k++] = nd.globalLayer; buf[k++] = nd.states.size();
        for (auto s : nd.states) buf[k++] = s;
        buf[k++] = nd.solutionVector.size();
        for (auto s : nd.solutionVector) buf[k++] = s;
    }
    return k;
}

// ---- RestrictedDDNew --------------------------------------------------------------------
struct ResHolder { std::shared_ptr<Network> net; RestrictedDDNew dd; std::optional<std::vector<Node>> cutset;
                   ResHolder(std::shared_ptr<Network> n, uint w) : net(n), dd(n, w) {} };
void *ref_res_new(void *net, int width) { return new ResHolder(static_cast<NetHolder *>(net)->p, width); }
void ref_res_free(void *d) { delete static_cast<ResHolder *>(d); }
int ref_res_compile(void *d, const int16_t *states, int ns, const int16_t *sol, int nsol, int gl) {
    auto *h = static_cast<ResHolder *>(d);
    h->cutset = h->dd.compile(make_node(states, ns, sol, nsol, gl));
    return h->cutset ? (int)h->cutset->size() : -1;   // -1: tree exact, no cut-set (DD.cpp:3157)
}
int ref_res_is_exact(void *d) { return static_cast<ResHolder *>(d)->dd.isTreeExact(); }
double ref_res_apply_opt(void *d, double rhs, const uint64_t *keys, const double *vals, int nnz) {
    return static_cast<ResHolder *>(d)->dd.applyOptimalityCut(make_cut(rhs, keys, vals, nnz));
}
int ref_res_apply_feas(void *d, double rhs, const uint64_t *keys, const double *vals, int nnz) {
    return static_cast<ResHolder *>(d)->dd.applyFeasibilityCut(make_cut(rhs, keys, vals, nnz));
}
int ref_res_solution(void *d, int16_t *out) {
    auto p = static_cast<ResHolder *>(d)->dd.getMaxPath();
    for (size_t i = 0; i < p.size(); i++) out[i] = p[i];
    return p.size();
}
int ref_res_nlayers(void *d) { return static_cast<ResHolder *>(d)->dd.tree.size(); }
void ref_res_layer_sizes(void *d, int *out) { auto &t = static_cast<ResHolder *>(d)->dd.tree; for (size_t i = 0; i < t.size(); i++) out[i] = t[i].size(); }
// per node (tree order, layers 1..last-1): parent position in previous layer and decision
long ref_res_dump(void *dv, int *parentpos, int *decision, double *state) {
    auto &d = static_cast<ResHolder *>(dv)->dd;
    long nn = 0;
    std::unordered_map<uint, int> pos_prev, pos_cur;
    for (size_t l = 0; l + 1 < d.tree.size(); l++) {
        pos_cur.clear();
        int k = 0;
        for (uint id : d.tree[l]) {
            const auto &nd = d.nodes.at(id);
            pos_cur[id] = k++;
            if (l > 0) { const auto &a = d.arcs.at(nd.incomingArc); parentpos[nn] = pos_prev.at(a.tail); decision[nn] = a.decision; }
            else { parentpos[nn] = -1; decision[nn] = 0; }
            state[nn] = nd.state2; nn++;
        }
        pos_prev.swap(pos_cur);
    }
    return nn;
}
int ref_res_cutset(void *dv, int *buf, int cap) {
    auto *h = static_cast<ResHolder *>(dv);
    if (!h->cutset) return 0;
    int k = 0;
    for (auto &nd : *h->cutset) {
        if (k + 3 + (int)nd.states.size() + (int)nd.solutionVector.size() > cap) return -1;
        buf[k++] = nd.globalLayer; buf[k++] = nd.states.size();
        for (auto s : nd.states) buf[k++] = s;
        buf[k++] = nd.solutionVector.size();
        for (auto s : nd.solutionVector) buf[k++] = s;
    }
    return k;
}
}  // extern "C"
