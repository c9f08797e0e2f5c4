// dd_host.hpp — host-side decision diagrams, flat and device-uploadable.
//
// Behavioural mirrors (re-designed: index arenas + a CSR-by-layer export instead of hash maps):
//   Inavap::RelaxedDDNew     /root/reference/DD.h:734-810,  DD.cpp:3509-4229
//   Inavap::RestrictedDDNew  /root/reference/DD.h:653-730,  DD.cpp:3090-3505
// The structure (layers, node order, in-arc order, states) is built on the host exactly as the
// reference builds it; the layer-wise longest path with cut-adjusted arc weights — the hot loop of
// applyOptimalityCut / applyFeasibilityCut (DD.cpp:3951-3973, 3860-3882, 3454-3470) — runs on the
// GPU (k2_dd.cu).  The sequential cut semantics around it (terminal min over cuts, node removal,
// bound-based arc pruning, path extraction) exist twice: here, for the batch entry point and as the
// statement of the semantics, and on the device (k2_finish / k2_extract) for the one-cut-at-a-time
// calls, whose flags absorb_device() replays into this mirror on demand.  Cut-sets stay on the host.
#pragma once
#include <cstdint>
#include <limits>
#include <string>
#include <vector>

#include "model.hpp"

namespace sgufp {

constexpr double DD_LOWEST = std::numeric_limits<double>::lowest();  // Inavap::DOUBLE_MIN (DD.h:453)
constexpr double DD_MAX = std::numeric_limits<double>::max();        // Inavap::DOUBLE_MAX (DD.h:454)

struct NodeSpec {  // Inavap::Node (DD.h:456-479) without the bounds
    std::vector<int16_t> states, solution;
    int global_layer = 0;
};

// CSR-by-layer image of a DD for the device.  Node indices are positions in tree
// order; the terminal node is implicit (one terminal arc per node of the last layer).
struct DDCsr {
    int nlayers = 0, nnodes = 0, narcs = 0, nlast = 0, max_width = 0;
    std::vector<int32_t> layer_ptr;   // [nlayers+1] node offsets
    std::vector<int32_t> in_ptr;      // [nnodes+1]  in-arc offsets (stored in-arc order)
    std::vector<int32_t> arc_tail;    // [narcs]     node index of the tail
    std::vector<int32_t> arc_slot;    // [narcs]     coefficient slot, -1 for decision -1
    std::vector<int32_t> root_slot;   // [nroot]     slots of the fixed prefix decisions, -1 for decision -1
    std::vector<int32_t> node_id;     // [nnodes]    arena id of each node (host use)
    std::vector<int32_t> arc_id;      // [narcs]     arena id of each arc (host use)
};

class HostDD {
public:
    struct Node {
        std::vector<int32_t> in, out;   // arena arc ids, in creation order
        std::vector<int16_t> states;
        double state2 = DD_LOWEST;
        int node_layer = 0, global_layer = 0;
        int32_t in_single = -1;         // restricted: the one incoming arc
    };
    struct Arc {
        int32_t tail = 0, head = 0;
        int16_t decision = 0;
        double weight = 0.0;
    };

    HostDD(const Model *M, bool restricted, int max_width) : M_(M), restricted_(restricted), max_width_(max_width) {}

    // buildTree (DD.cpp:3528-3600) / compile (DD.cpp:3090-3159).  For the restricted DD, `cutset`
    // receives the exact cut-set when the tree is not exact (DD.cpp:3157-3158).
    void build(const NodeSpec &root, std::vector<NodeSpec> *cutset);
    bool is_exact() const { return exact_; }
    bool restricted() const { return restricted_; }
    int max_width() const { return max_width_; }
    int slot_of(int global_layer, int decision) const;  // coefficient slot of (layer, decision) or -1

    // --- cut application, split around the device pass --------------------------------------
    // 1. flatten() -> upload -> K2 computes every node's state2 for the cut
    // 2. the host finishes with exactly the reference's sequential semantics
    const DDCsr &flatten();
    bool dirty() const { return dirty_; }
    unsigned long version() const { return version_; }   // bumped by every structural change
    // dense coefficient vector of a sparse Inavap::Cut: first matching key wins (Cut.h:275-282)
    void densify(const uint64_t *keys, const double *vals, int nnz, std::vector<double> &coef) const;
    // `states` in CSR node order, `term` = terminal arc weights after min with the last-layer states
    double finish_optimality(const std::vector<double> &coef, const std::vector<double> &states, double optimal, double ub);
    int finish_feasibility(const std::vector<double> &coef, const std::vector<double> &states);
    std::vector<double> &terminal_weights() { return term_; }   // one per node of the last layer, in layer order
    // Replay what the device did to its image of THIS flatten() (k2_finish): node
    // states and terminal weights of the last cut, arcs and nodes it removed (flags in CSR order).
    // The result is the structure finish_optimality / finish_feasibility would have left.
    void absorb_device(const std::vector<double> &states, const std::vector<uint8_t> &arc_dead, const std::vector<uint8_t> &node_dead,
                       const std::vector<double> &term, const std::vector<double> &coef);

    std::vector<int16_t> solution() const;                       // getSolution / getMaxPath
    std::vector<NodeSpec> cutset(double ub) const;                // RelaxedDDNew::getCutset (DD.cpp:4179-4218)

    // introspection for the parity tests
    const std::vector<std::vector<int32_t>> &tree() const { return tree_; }
    const std::vector<Node> &nodes() const { return nodes_; }
    const std::vector<Arc> &arcs() const { return arcs_; }
    long count_arcs() const;
    int start_tree() const { return start_; }
    const std::vector<int16_t> &root_solution() const { return root_solution_; }

private:
    const Model *M_;
    bool restricted_;
    int max_width_;
    std::vector<Node> nodes_;
    std::vector<Arc> arcs_;
    std::vector<std::vector<int32_t>> tree_;   // layers of arena node ids; the terminal is NOT a layer here
    std::vector<double> term_;                 // terminal arc weight per last-layer node (parallel to tree_.back())
    std::vector<int16_t> root_solution_;
    int start_ = 0;
    bool exact_ = true, dirty_ = true;
    unsigned long version_ = 0;
    DDCsr csr_;
    std::vector<double> last_coef_;

    int new_node() { nodes_.emplace_back(); return (int)nodes_.size() - 1; }
    int new_arc(int tail, int head, int16_t dec) { arcs_.push_back({tail, head, dec, 0.0}); return (int)arcs_.size() - 1; }
    std::vector<int16_t> layer_states(int g) const;     // stateUpdateMap[g] (Network.cpp:98-102) or empty
    void relaxed_next_layer(int index, unsigned &next_size);
    std::vector<int32_t> restricted_next_layer(const std::vector<int32_t> &cur, bool &exact);
    std::vector<int16_t> path_for_node(int id) const;
    double arc_weight_now(const Arc &a) const;           // weight under the last applied cut
    void remove_last_layer_nodes(const std::vector<int32_t> &ids);
    void bottom_up_delete(int id, std::vector<uint8_t> &dead);
};

}  // namespace sgufp
