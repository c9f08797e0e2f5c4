// dd_host.cpp — see dd_host.hpp.  Host-only.
#include "dd_host.hpp"

#include <algorithm>
#include <set>
#include <unordered_map>

namespace sgufp {

// stateUpdateMap (Network.cpp:96-102): at the first layer of a V-bar node the states become
// {-1} U outgoingArcs(q), iterated as a std::set<int> (ascending, -1 first).
std::vector<int16_t> HostDD::layer_states(int g) const {
    const Model &M = *M_;
    if (g < 0 || g >= M.L) return {};
    const int q = M.head[M.layer_arc[g]];
    if (g > 0 && M.head[M.layer_arc[g - 1]] == q) return {};
    std::vector<int16_t> st{-1};
    std::vector<int32_t> outs(M.out_arc.begin() + M.out_ptr[q], M.out_arc.begin() + M.out_ptr[q + 1]);
    std::sort(outs.begin(), outs.end());
    for (int a : outs) st.push_back((int16_t)a);
    return st;
}

// key (q,i,j) of layer g and decision d (DD.cpp:3952-3965): i,q = endpoints of the layer's arc,
// j = head(d).  The slot is the first out-arc of q whose head is j (keys are node ids).
int HostDD::slot_of(int g, int decision) const {
    const Model &M = *M_;
    if (decision < 0 || g < 0 || g >= M.L || decision >= M.m) return -1;
    const int q = M.head[M.layer_arc[g]], j = M.head[decision];
    for (int k = 0; k < M.outdeg(q); k++)
        if (M.head[M.out_arc[M.out_ptr[q] + k]] == j) return M.slot_base[g] + k;
    return -1;   // cut.get() finds no such key -> weight 0 (Cut.h:282)
}

void HostDD::build(const NodeSpec &root, std::vector<NodeSpec> *cutset) {
    const Model &M = *M_;
    nodes_.clear(); arcs_.clear(); tree_.clear(); term_.clear(); last_coef_.clear();
    exact_ = true; dirty_ = true; version_++;
    start_ = root.global_layer;
    root_solution_ = root.solution;
    const int r = new_node();
    nodes_[r].states = root.states;
    nodes_[r].global_layer = root.global_layer;
    tree_.push_back({r});
    if (!restricted_) {
        unsigned next_size = 0;
        int index = 0;
        for (int g = start_; g < M.L; ++g, ++index) {
            const std::vector<int16_t> st = layer_states(g);
            if (!st.empty()) {
                for (int id : tree_[index]) nodes_[id].states = st;
                next_size = (unsigned)(tree_[index].size() * st.size());   // DD.cpp:3570
            }
            relaxed_next_layer(index, next_size);
        }
    } else {
        bool ex = true;
        int exact_layer = 0;
        std::vector<int32_t> cur = tree_[0];
        for (int g = start_; g < M.L; ++g) {
            const std::vector<int16_t> st = layer_states(g);
            if (!st.empty()) for (int id : cur) nodes_[id].states = st;
            std::vector<int32_t> nxt = restricted_next_layer(cur, ex);
            if (ex) exact_layer++;
            tree_.push_back(nxt);
            cur.swap(nxt);
        }
        exact_ = ex;
        if (cutset) {
            cutset->clear();
            if (!ex)   // getExactCutSet (DD.cpp:3279-3288)
                for (int id : tree_[exact_layer]) cutset->push_back({nodes_[id].states, path_for_node(id), nodes_[id].global_layer});
        }
    }
    term_.assign(tree_.back().size(), DD_MAX);   // terminal arcs start at DOUBLE_MAX (DD.cpp:3145, 3595)
}

// RelaxedDDNew::buildNextLayer (DD.cpp:3603-3694)
void HostDD::relaxed_next_layer(int index, unsigned &next_size) {
    const Model &M = *M_;
    const std::vector<int32_t> cur = tree_[index];
    const unsigned gl = (unsigned)nodes_[cur.front()].global_layer;
    const unsigned threshold = max_width_ > 0 ? (unsigned)max_width_ : 120u;   // RELAXED_MAX_WIDTH = 120 (DD.h:732), a runtime parameter here
    if (next_size >= threshold && gl < (unsigned)M.L - 5u) {
        exact_ = false;
        const int nn = new_node();
        std::set<int16_t> all;
        for (int id : cur) {
            const std::vector<int16_t> st = nodes_[id].states;
            all.insert(st.begin(), st.end());
            for (int16_t s : st) {
                const int a = new_arc(id, nn, s);
                nodes_[id].out.push_back(a);
                nodes_[nn].in.push_back(a);
            }
        }
        next_size = (unsigned)all.size();
        nodes_[nn].node_layer = nodes_[cur[0]].node_layer + 1;
        nodes_[nn].global_layer = nodes_[cur[0]].global_layer + 1;
        nodes_[nn].states.assign(all.begin(), all.end());
        tree_.push_back({nn});
        return;
    }
    std::vector<int32_t> nxt;
    next_size = 0;
    for (int id : cur) {
        const std::vector<int16_t> st = nodes_[id].states;
        for (int16_t s : st) {
            std::vector<int16_t> ns = st;
            if (s != -1) ns.erase(std::remove(ns.begin(), ns.end(), s), ns.end());
            next_size += (unsigned)ns.size();
            const int c = new_node();
            const int a = new_arc(id, c, s);
            nodes_[c].in.push_back(a);
            nodes_[c].node_layer = nodes_[id].node_layer + 1;
            nodes_[c].global_layer = nodes_[id].global_layer + 1;
            nodes_[c].states = std::move(ns);
            nodes_[id].out.push_back(a);
            nxt.push_back(c);
        }
    }
    tree_.push_back(std::move(nxt));
}

// RestrictedDDNew::buildNextLayer / buildRestrictedLayer (DD.cpp:3161-3260)
std::vector<int32_t> HostDD::restricted_next_layer(const std::vector<int32_t> &cur, bool &exact) {
    std::vector<int32_t> nxt;
    auto child = [&](int id, int16_t dec, const std::vector<int16_t> &st) {
        const int c = new_node();
        const int a = new_arc(id, c, dec);
        nodes_[id].out.push_back(a);
        nodes_[c].in_single = a;
        nodes_[c].in.push_back(a);
        nodes_[c].states = st;
        if (dec != -1) {
            auto it = std::find(nodes_[c].states.begin(), nodes_[c].states.end(), dec);
            if (it != nodes_[c].states.end()) nodes_[c].states.erase(it);
        }
        nodes_[c].node_layer = nodes_[id].node_layer + 1;
        nodes_[c].global_layer = nodes_[id].global_layer + 1;
        nxt.push_back(c);
    };
    if (exact) {
        size_t count = 0;
        for (int id : cur) {
            const std::vector<int16_t> st = nodes_[id].states;
            for (auto it = st.rbegin(); it != st.rend(); ++it) {   // states in REVERSE order (DD.cpp:3235)
                if (count >= (size_t)max_width_) { exact = false; return nxt; }
                child(id, *it, st);
                count++;
            }
        }
        return nxt;
    }
    for (int id : cur) {
        const std::vector<int16_t> st = nodes_[id].states;
        child(id, *std::max_element(st.begin(), st.end()), st);   // DD.cpp:3204
    }
    return nxt;
}

long HostDD::count_arcs() const {
    long c = 0;
    for (size_t l = 1; l < tree_.size(); l++)
        for (int id : tree_[l]) c += (long)nodes_[id].in.size();
    return c + (long)tree_.back().size();   // + terminal arcs
}

const DDCsr &HostDD::flatten() {
    if (!dirty_) return csr_;
    DDCsr &C = csr_;
    C = DDCsr();
    C.nlayers = (int)tree_.size();
    std::unordered_map<int, int> pos;
    C.layer_ptr.push_back(0);
    C.in_ptr.push_back(0);
    int width = 0;
    for (size_t l = 0; l < tree_.size(); l++) {
        width = std::max(width, (int)tree_[l].size());
        for (int id : tree_[l]) {
            pos[id] = (int)C.node_id.size();
            C.node_id.push_back(id);
            if (l > 0)
                for (int a : nodes_[id].in) {
                    C.arc_tail.push_back(pos.at(arcs_[a].tail));
                    C.arc_slot.push_back(slot_of((int)root_solution_.size() + (int)l - 1, arcs_[a].decision));   // i advances once per prefix decision (DD.cpp:3938-3952)
                    C.arc_id.push_back(a);
                }
            C.in_ptr.push_back((int32_t)C.arc_tail.size());
        }
        C.layer_ptr.push_back((int32_t)C.node_id.size());
    }
    C.nnodes = (int)C.node_id.size(); C.narcs = (int)C.arc_tail.size(); C.nlast = (int)tree_.back().size(); C.max_width = width;
    // justified RHS of the root: one term per fixed prefix decision, -1 entries advance the layer too (DD.cpp:3938-3949)
    for (size_t i = 0; i < root_solution_.size(); i++) C.root_slot.push_back(slot_of((int)i, root_solution_[i]));
    dirty_ = false;
    return C;
}

void HostDD::densify(const uint64_t *keys, const double *vals, int nnz, std::vector<double> &coef) const {
    const Model &M = *M_;
    coef.assign(std::max(1, M.T), 0.0);
    std::vector<uint8_t> seen(std::max(1, M.T), 0);
    const auto &slot = M.key_slot;            // key (48 bits) -> first slot with that (q,i,j), built once per model
    for (int k = 0; k < nnz; k++) {
        auto it = slot.find(keys[k] & 0xFFFFFFFFFFFFull);   // IQJ_MASK (Cut.h:197)
        if (it == slot.end() || seen[it->second]) continue;   // get() returns the FIRST match (Cut.h:275-282)
        seen[it->second] = 1;
        coef[it->second] = vals[k];
    }
}

double HostDD::arc_weight_now(const Arc &a) const {
    // arc.weight is rewritten by every cut for decisions != -1 and never for -1 (stays 0 from the
    // reset, DD.cpp:3520-3526, 3962-3969)
    return a.weight;
}

double HostDD::finish_optimality(const std::vector<double> &coef, const std::vector<double> &states, double optimal, double ub) {
    (void)ub;
    const DDCsr &C = flatten();
    last_coef_ = coef;
    for (int v = 0; v < C.nnodes; v++) nodes_[C.node_id[v]].state2 = states[v];
    for (int e = 0; e < C.narcs; e++)
        if (C.arc_slot[e] >= 0 || arcs_[C.arc_id[e]].decision != -1) arcs_[C.arc_id[e]].weight = C.arc_slot[e] >= 0 ? coef[C.arc_slot[e]] : 0.0;
    const std::vector<int32_t> &last = tree_.back();
    double terminal = DD_LOWEST;
    for (size_t i = 0; i < last.size(); i++) {
        term_[i] = std::min(term_[i], nodes_[last[i]].state2);   // weights persist across cuts (DD.cpp:3981)
        terminal = std::max(terminal, term_[i]);
    }
    if (restricted_) return terminal;                              // DD.cpp:3493-3504
    if (terminal <= optimal) return terminal;                      // DD.cpp:3985
    if (!exact_) {                                                 // DD.cpp:3987-4021
        const size_t llayer = tree_.size() - 1;
        double max_state = DD_LOWEST;
        for (int id : tree_[llayer]) max_state = std::max(max_state, nodes_[id].state2);
        std::vector<int32_t> drop;
        for (size_t layer = 3; layer + 1 < llayer; layer++) {
            if (tree_[layer].size() != 1) continue;
            size_t pruned = 0, total = 0;
            const double gain = max_state - nodes_[tree_[layer][0]].state2;
            for (int id : tree_[layer - 1])
                for (int a : nodes_[id].out) {
                    if ((nodes_[arcs_[a].tail].state2 + arc_weight_now(arcs_[a]) + gain) <= (optimal - 0.01)) { drop.push_back(a); pruned++; }
                    total++;
                }
            if (pruned == total) return DD_LOWEST;                 // DD.cpp:4015
        }
        if (!drop.empty()) {
            for (int a : drop) {                                   // batchRemoveArcs (DD.cpp:4162-4177)
                auto &hin = nodes_[arcs_[a].head].in; hin.erase(std::remove(hin.begin(), hin.end(), a), hin.end());
                auto &tout = nodes_[arcs_[a].tail].out; tout.erase(std::remove(tout.begin(), tout.end(), a), tout.end());
            }
            dirty_ = true; version_++;
        }
    }
    return terminal;
}

void HostDD::bottom_up_delete(int id, std::vector<uint8_t> &dead) {
    const std::vector<int32_t> in = nodes_[id].in;
    for (int a : in) {
        const int p = arcs_[a].tail;
        auto &pout = nodes_[p].out; pout.erase(std::remove(pout.begin(), pout.end(), a), pout.end());
        auto &nin = nodes_[id].in; nin.erase(std::remove(nin.begin(), nin.end(), a), nin.end());
        if (pout.empty()) bottom_up_delete(p, dead);
    }
    dead[id] = 1;
}

// RelaxedDDNew::batchRemoveNodes / removeNode / updateTree (DD.cpp:4040-4160)
void HostDD::remove_last_layer_nodes(const std::vector<int32_t> &ids) {
    std::vector<uint8_t> dead(nodes_.size(), 0);
    for (int id : ids) {
        nodes_[id].out.clear();   // its arc to the terminal
        const std::vector<int32_t> in = nodes_[id].in;
        for (int a : in) {
            const int p = arcs_[a].tail;
            auto &pout = nodes_[p].out; pout.erase(std::remove(pout.begin(), pout.end(), a), pout.end());
            auto &nin = nodes_[id].in; nin.erase(std::remove(nin.begin(), nin.end(), a), nin.end());
            if (pout.empty()) bottom_up_delete(p, dead);
        }
        dead[id] = 1;
    }
    for (size_t l = 0; l < tree_.size(); l++) {
        const bool is_last = l + 1 == tree_.size();
        std::vector<int32_t> keep;
        std::vector<double> keep_term;
        for (size_t i = 0; i < tree_[l].size(); i++)
            if (!dead[tree_[l][i]]) { keep.push_back(tree_[l][i]); if (is_last) keep_term.push_back(term_[i]); }
        tree_[l].swap(keep);
        if (is_last) term_.swap(keep_term);
    }
    dirty_ = true; version_++;
}

void HostDD::absorb_device(const std::vector<double> &states, const std::vector<uint8_t> &arc_dead, const std::vector<uint8_t> &node_dead,
                           const std::vector<double> &term, const std::vector<double> &coef) {
    const DDCsr &C = csr_;   // the image the device worked on
    last_coef_ = coef;
    for (int v = 0; v < C.nnodes; v++) nodes_[C.node_id[v]].state2 = states[v];
    for (int e = 0; e < C.narcs; e++)
        if (C.arc_slot[e] >= 0 || arcs_[C.arc_id[e]].decision != -1) arcs_[C.arc_id[e]].weight = C.arc_slot[e] >= 0 ? coef[C.arc_slot[e]] : 0.0;
    term_ = term;
    bool changed = false;
    std::vector<uint8_t> gone(arcs_.size(), 0), touched(nodes_.size(), 0);
    for (int e = 0; e < C.narcs; e++)
        if (arc_dead[e]) { const int a = C.arc_id[e]; gone[a] = 1; touched[arcs_[a].head] = 1; touched[arcs_[a].tail] = 1; changed = true; }
    if (changed)
        for (size_t id = 0; id < nodes_.size(); id++)
            if (touched[id]) {
                auto dead = [&](int a) { return gone[a] != 0; };
                auto &in = nodes_[id].in; in.erase(std::remove_if(in.begin(), in.end(), dead), in.end());
                auto &out = nodes_[id].out; out.erase(std::remove_if(out.begin(), out.end(), dead), out.end());
            }
    std::vector<uint8_t> dead_id(nodes_.size(), 0);
    bool nodes_gone = false;
    for (int v = 0; v < C.nnodes; v++) if (node_dead[v]) { dead_id[C.node_id[v]] = 1; nodes_gone = true; }
    if (nodes_gone) {
        for (size_t l = 0; l < tree_.size(); l++) {
            const bool is_last = l + 1 == tree_.size();
            std::vector<int32_t> keep;
            std::vector<double> keep_term;
            for (size_t i = 0; i < tree_[l].size(); i++) {
                const int id = tree_[l][i];
                if (!dead_id[id]) { keep.push_back(id); if (is_last) keep_term.push_back(term_[i]); }
                else if (is_last) nodes_[id].out.clear();   // its arc to the terminal
            }
            tree_[l].swap(keep);
            if (is_last) term_.swap(keep_term);
        }
        changed = true;
    }
    if (changed) { dirty_ = true; version_++; }
}

int HostDD::finish_feasibility(const std::vector<double> &coef, const std::vector<double> &states) {
    const DDCsr &C = flatten();
    last_coef_ = coef;
    for (int v = 0; v < C.nnodes; v++) nodes_[C.node_id[v]].state2 = states[v];
    for (int e = 0; e < C.narcs; e++)
        if (C.arc_slot[e] >= 0 || arcs_[C.arc_id[e]].decision != -1) arcs_[C.arc_id[e]].weight = C.arc_slot[e] >= 0 ? coef[C.arc_slot[e]] : 0.0;
    const size_t llayer = tree_.size() - 1;
    if (restricted_) {                                             // DD.cpp:3374-3422
        if (tree_.size() >= 2) {
            // the arc weight of a -1 decision IS rewritten to 0 here (DD.cpp:3395); already 0
        }
        if (term_.empty()) return 0;
        std::vector<int32_t> drop;
        if (tree_.size() >= 2)
            for (int id : tree_[llayer]) if (nodes_[id].state2 < -0.5) drop.push_back(id);
        if (!drop.empty()) {
            std::vector<int32_t> keep; std::vector<double> keep_term;
            std::set<int32_t> d(drop.begin(), drop.end());
            for (size_t i = 0; i < tree_[llayer].size(); i++)
                if (!d.count(tree_[llayer][i])) { keep.push_back(tree_[llayer][i]); keep_term.push_back(term_[i]); }
                else nodes_[tree_[llayer][i]].out.clear();
            tree_[llayer].swap(keep); term_.swap(keep_term);
            dirty_ = true; version_++;
        }
        return term_.empty() ? 0 : 1;
    }
    std::vector<int32_t> drop;
    for (int id : tree_[llayer]) if (nodes_[id].state2 < -0.01) drop.push_back(id);   // DD.cpp:3887
    if (drop.size() == tree_[llayer].size()) return 0;
    if (!drop.empty()) remove_last_layer_nodes(drop);
    if (!exact_) {                                                 // DD.cpp:3895-3928
        double max_state = DD_LOWEST;
        for (int id : tree_[llayer]) max_state = std::max(max_state, nodes_[id].state2);
        std::vector<int32_t> cut_arcs;
        for (size_t layer = 1; layer < llayer; layer++) {
            if (tree_[layer].size() != 1) continue;
            size_t pruned = 0, total = 0;
            const double gain = max_state - nodes_[tree_[layer][0]].state2;
            for (int id : tree_[layer - 1])
                for (int a : nodes_[id].out) {
                    if ((nodes_[arcs_[a].tail].state2 + arc_weight_now(arcs_[a]) + gain) <= -0.01) { cut_arcs.push_back(a); pruned++; }
                    total++;
                }
            if (total == pruned) return 0;
        }
        if (!cut_arcs.empty()) {
            for (int a : cut_arcs) {
                auto &hin = nodes_[arcs_[a].head].in; hin.erase(std::remove(hin.begin(), hin.end(), a), hin.end());
                auto &tout = nodes_[arcs_[a].tail].out; tout.erase(std::remove(tout.begin(), tout.end(), a), tout.end());
            }
            dirty_ = true; version_++;
        }
    }
    return 1;
}

// RelaxedDDNew::getPathForNode (DD.cpp:3796-3820) / RestrictedDDNew::getPathForNode (DD.cpp:3262-3276)
std::vector<int16_t> HostDD::path_for_node(int id) const {
    std::vector<int16_t> rev;
    int cur = id;
    if (restricted_) {
        while (nodes_[cur].node_layer) {
            const Arc &a = arcs_[nodes_[cur].in_single];
            rev.push_back(a.decision);
            cur = a.tail;
        }
    } else {
        while (nodes_[cur].node_layer) {
            if (nodes_[cur].in.empty()) break;   // the reference would read incomingArcs[0] of an empty vector
            int next = arcs_[nodes_[cur].in[0]].tail;
            for (int a : nodes_[cur].in) {
                const int p = arcs_[a].tail;
                if ((nodes_[p].state2 + arc_weight_now(arcs_[a])) == nodes_[cur].state2) {   // exact fp compare (DD.cpp:3808)
                    rev.push_back(arcs_[a].decision);
                    next = p;
                    break;
                }
            }
            cur = next;   // no match: step to in-arc 0's parent WITHOUT recording a decision (DD.cpp:3803,3814)
        }
    }
    std::vector<int16_t> sol(root_solution_);
    sol.insert(sol.end(), rev.rbegin(), rev.rend());
    return sol;
}

std::vector<int16_t> HostDD::solution() const {   // getSolution (DD.cpp:3825-3840) / getMaxPath (DD.cpp:3290-3305)
    int best = tree_[0][0];                          // maxId = 0: the root
    double w = DD_LOWEST;
    const std::vector<int32_t> &last = tree_.back();
    for (size_t i = 0; i < last.size(); i++)
        if (term_[i] > w) { w = term_[i]; best = last[i]; }
    return path_for_node(best);
}

std::vector<NodeSpec> HostDD::cutset(double ub) const {
    (void)ub;
    const Model &M = *M_;
    std::vector<NodeSpec> out;
    size_t layer = 3;
    while (layer < tree_.size() && tree_[layer].size() != 1) layer++;   // first collapsed layer (DD.cpp:4181-4182)
    if (layer >= tree_.size()) return out;                               // the reference runs off the end here
    const int gl = nodes_[tree_[layer][0]].global_layer;
    const std::vector<int16_t> fresh = layer_states(gl);                 // hasStateChanged[gl] (DD.cpp:4189)
    for (int id : tree_[layer - 1]) {
        const std::vector<int16_t> partial = path_for_node(id);
        for (int a : nodes_[id].out) {
            std::vector<int16_t> sol = partial;
            sol.push_back(arcs_[a].decision);
            if (!fresh.empty() && gl < M.L) out.push_back({fresh, sol, gl});
            else {
                std::vector<int16_t> st = nodes_[id].states;
                if (arcs_[a].decision != -1) st.erase(std::remove(st.begin(), st.end(), arcs_[a].decision), st.end());
                out.push_back({st, sol, gl});
            }
        }
    }
    return out;
}

}  // namespace sgufp
