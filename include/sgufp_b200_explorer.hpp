// sgufp_b200_explorer.hpp — the caller of the hot path in C++: the Benders loop on branch-and-bound nodes over the
// adapters of sgufp_b200.hpp, batch-aware.
//
//   process(node, ...)        Inavap::NodeExplorer::process (/root/reference/NodeExplorer.cpp:915-986) statement by statement;
//                             its loops over the GLOBAL cuts (:935-944, :975-983) are one device call each
//                             (sgufp_dd_apply_sequence), with the results of the one-by-one calls.
//   process_many(nodes, ...)  W nodes side by side, one diagram each: every round hands the current argmax path of every
//                             node that is still in its cut loop to ONE sgufp_solve_paths call (K = nodes still looping),
//                             then each node applies ITS OWN cut to its own diagram and takes its own decision
//                             (NodeExplorer.cpp:949-971 per node, unchanged).  This is what the reference's N_WORKERS
//                             explorers do on N different nodes at the same time (DDSolver.cpp:701-712), in lock step: the
//                             global containers receive every new cut when it is generated, a node reads them when it
//                             starts — the K1 launch sees K candidates instead of 1.
//
// Include AFTER the reference's "Network.h", "Cut.h", "DD.h" (NodeT = Inavap::Node, ContainerT = Inavap::Container,
// CutNodeT = Inavap::cut_node_t): the reference's own NodeExplorer.h cannot be included without Gurobi (it pulls grb.h).
#pragma once
#include <algorithm>
#include <memory>
#include <vector>

#include "sgufp_b200.hpp"

namespace sgufp {

// Inavap::OutObj (NodeExplorer.h:86-103)
template <class NodeT>
struct OutObjectT {
    enum STATUS_OP : uint16_t { SUCCESS = 0x0, PRUNED_BY_FEASIBILITY_CUT = 0x1, PRUNED_BY_OPTIMALITY_CUT = 0x2 };
    double lb = std::numeric_limits<double>::lowest();
    double ub = std::numeric_limits<double>::lowest();
    std::vector<NodeT> nodes;
    uint16_t status = SUCCESS;
};

template <class NodeT, class ContainerT, class CutNodeT>
class NodeExplorerT {
    using Out = OutObjectT<NodeT>;
    using Path = std::vector<int16_t>;

public:
    // `width`: how many nodes process_many takes side by side (one reusable diagram each, NodeExplorer.h:116)
    explicit NodeExplorerT(const std::shared_ptr<Network> &net, int device = 0, int width = 1, int device_count = 1)
        : solver_(net, device, device_count) {
        for (int i = 0; i < std::max(1, width); i++) dds_.emplace_back(new RelaxedDDNewT<NodeT>(solver_));
    }
    GuroSolver &solver() { return solver_; }
    long cuts_generated = 0, k1_calls = 0;

    Out process(NodeT node, double optimalLB, ContainerT &globalFeasCuts, ContainerT &globalOptCuts) {
        std::vector<NodeT> one;
        one.push_back(std::move(node));
        return std::move(process_many(std::move(one), optimalLB, globalFeasCuts, globalOptCuts)[0]);
    }

    std::vector<Out> process_many(std::vector<NodeT> nodes, double optimalLB, ContainerT &globalFeasCuts, ContainerT &globalOptCuts) {
        const double DOUBLE_MIN_ = std::numeric_limits<double>::lowest();
        const int n = (int)nodes.size();
        if (n > (int)dds_.size()) raise("sgufp_b200: process_many takes at most `width` nodes");
        std::vector<Out> out(n);
        std::vector<double> ub(n);
        std::vector<char> looping(n, 0);
        std::vector<std::vector<Path>> seen(n);
        auto pruned = [&](int i, uint16_t why) { out[i].lb = DOUBLE_MIN_; out[i].ub = DOUBLE_MIN_; out[i].nodes.clear(); out[i].status = why; looping[i] = 0; };
        // the cuts every node of this call starts from (a node reads the containers when it starts, :928-929)
        std::vector<const Inavap::Cut *> feas, opt;
        for (const CutNodeT *c = globalFeasCuts.get(); c; c = c->next) feas.push_back(&c->cut);
        for (const CutNodeT *c = globalOptCuts.get(); c; c = c->next) opt.push_back(&c->cut);
        for (int i = 0; i < n; i++) {
            RelaxedDDNewT<NodeT> &dd = *dds_[i];
            ub[i] = nodes[i].ub;
            dd.buildTree(nodes[i]);                                                            // :922
            const bool exact = dd.isTreeExact();                                               // :931
            if (!dd.applyFeasibilityCuts(feas)) { pruned(i, Out::PRUNED_BY_FEASIBILITY_CUT); continue; }   // :935-938 / :975-978
            double b = 0;
            const int applied = dd.applyOptimalityCuts(opt, optimalLB, exact, b);              // :940-944 / :980-983
            if (applied > 0) ub[i] = exact ? b : std::min(ub[i], b);
            if (applied > 0 && ub[i] <= optimalLB) { pruned(i, Out::PRUNED_BY_OPTIMALITY_CUT); continue; }
            if (exact) looping[i] = 1;
            else { out[i].lb = DOUBLE_MIN_; out[i].ub = ub[i]; out[i].nodes = dd.getCutset(ub[i]); out[i].status = Out::SUCCESS; }   // :985
        }
        // the cut loops of the exact nodes, in lock step (:949-971 per node)
        for (;;) {
            std::vector<int> who;
            std::vector<Path> paths;
            for (int i = 0; i < n; i++) {
                if (!looping[i]) continue;
                Path path = dds_[i]->getSolution();                                            // :950
                if (std::find(seen[i].begin(), seen[i].end(), path) != seen[i].end()) {       // :951-953
                    out[i].lb = ub[i]; out[i].ub = ub[i]; out[i].nodes.clear(); out[i].status = Out::SUCCESS; looping[i] = 0;
                    continue;
                }
                seen[i].push_back(path);
                who.push_back(i);
                paths.push_back(std::move(path));
            }
            if (who.empty()) break;
            auto cuts = solver_.solveSubProblems(paths);                                       // :957 for every node at once (K1)
            k1_calls++;
            cuts_generated += (long)cuts.size();
            for (size_t k = 0; k < who.size(); k++) {
                const int i = who[k];
                CutNodeT *new_cut = new CutNodeT{cuts[k].second};                              // :958
                if (cuts[k].first == FEASIBILITY) {
                    globalFeasCuts.add(new_cut);
                    if (!dds_[i]->applyFeasibilityCut(cuts[k].second)) pruned(i, Out::PRUNED_BY_FEASIBILITY_CUT);   // :961-962
                } else {
                    globalOptCuts.add(new_cut);
                    ub[i] = dds_[i]->applyOptimalityCut(cuts[k].second, optimalLB, ub[i]);    // :966
                    if (ub[i] <= optimalLB) pruned(i, Out::PRUNED_BY_OPTIMALITY_CUT);        // :967-968
                }
            }
        }
        return out;
    }

private:
    GuroSolver solver_;
    std::vector<std::unique_ptr<RelaxedDDNewT<NodeT>>> dds_;
};

// Sequential depth-first branch and bound over cut-set nodes (stand-in for the worker loop of DDSolver.cpp:658-776: pop,
// prune on ub <= incumbent, process, raise the incumbent, push the children) taking up to `width` nodes per round.
template <class NodeT, class ContainerT, class CutNodeT>
double solve_frontier(NodeExplorerT<NodeT, ContainerT, CutNodeT> &ex, int width, long max_nodes, long *nodes_processed) {
    ContainerT feas, opt;
    double best = std::numeric_limits<double>::lowest();
    std::vector<NodeT> stack;
    stack.emplace_back();
    stack.back().ub = std::numeric_limits<double>::max();
    long processed = 0;
    while (!stack.empty() && processed < max_nodes) {
        std::vector<NodeT> batch;
        while (!stack.empty() && (int)batch.size() < width) {
            NodeT nd = std::move(stack.back());
            stack.pop_back();
            if (nd.ub <= best) continue;                                                       // DDSolver.cpp:707-711
            batch.push_back(std::move(nd));
        }
        if (batch.empty()) break;
        auto outs = ex.process_many(std::move(batch), best, feas, opt);                        // DDSolver.cpp:712
        processed += (long)outs.size();
        for (auto &o : outs) {
            if (o.status != OutObjectT<NodeT>::SUCCESS) continue;
            if (o.lb > best) best = o.lb;                                                      // DDSolver.cpp:723-731
            for (auto &ch : o.nodes) if (ch.ub > best) stack.push_back(std::move(ch));         // DDSolver.cpp:744-748
        }
    }
    if (nodes_processed) *nodes_processed = processed;
    return best;
}

}  // namespace sgufp
