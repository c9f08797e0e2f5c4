// model.hpp — host-side instance model and per-candidate evaluation plan.
//
// Mirrors what the reference derives once per instance:
//   Network::Network / shuffleVBarNodes / processingOrder   /root/reference/Network.cpp:10-186
//   the (i,q,j) coefficient keys GuroSolver pre-seeds       /root/reference/grb.h:60-68
// and what it derives once per candidate:
//   path -> y-bar                                            /root/reference/grb.cpp:141-150
// The plan is the scenario-invariant part of the subproblem: with y-bar fixed the
// second stage is a max-reward flow on a CONTRACTED graph whose arcs are chains of network arcs
// glued through matched V-bar pairs.  Only capacities differ between scenarios.
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

namespace sgufp {

enum ArcKind : int8_t { KIND_GAMMA = 0, KIND_SIGMA = 1, KIND_PHI = 2 };

struct Model {
    int n = 0, m = 0;
    std::vector<int32_t> tail, head, rew;
    std::vector<int32_t> in_ptr, in_arc, out_ptr, out_arc, out_index;
    std::vector<int32_t> vbar;               // after the shuffle
    std::vector<uint8_t> is_vbar, active, is_root;
    int L = 0, T = 0;
    std::vector<int32_t> layer_arc, arc_layer, slot_base, slot_in, slot_out, slot_lex_rank, slot_sorted;
    std::vector<uint64_t> slot_key_sorted;   // getKey(q, i, j) of slot_sorted[r] (Cut.h:342-344)
    std::unordered_map<uint64_t, int32_t> key_slot;   // getKey(q,i,j) (Cut.h:342-344) -> first slot with those node ids
    int nc = 1;                              // contracted nodes, 0 = root
    std::vector<int32_t> cn;                 // node -> contracted id (0 root, -1 active V-bar)
    int nav = 0;                             // active V-bar nodes
    std::vector<int32_t> av_index;           // node -> compact active-V-bar index or -1
    std::vector<int32_t> av_node;

    // returns 0 or a negative SGUFP_ERR_* code, `err` explains
    int build(int n, int m, const int32_t *tail, const int32_t *head, const int32_t *rew0, const int32_t *vbar,
              int nvbar, std::string &err);
    int indeg(int v) const { return in_ptr[v + 1] - in_ptr[v]; }
    int outdeg(int v) const { return out_ptr[v + 1] - out_ptr[v]; }
};

// Flat, device-uploadable description of one candidate.  All arrays are int32 and live in one
// pool so that a batch of K plans is a single host->device copy.
struct PlanHeader {
    int32_t nch, nopen, nc, nav, m, L;
    // offsets (in int32 units, relative to the plan's base) of the arrays below
    int32_t o_arc_cp;      // [m]      (chain << 10) | position-in-chain
    int32_t o_arc_info;    // [m]      kind | (layer+1) << 2     (layer of the arc as a V-bar in-arc, -1 if none)
    int32_t o_arc_pre;     // [m]      prefix reward of the chain up to and including the arc
    int32_t o_ch_ends;     // [nch]    (sv+1) | (ev+1) << 16     (contracted ids, 0 = dangling)
    int32_t o_ch_r;        // [nch]    chain reward
    int32_t o_ch_ptr;      // [nch+1]
    int32_t o_ch_arcs;     // [m]      arcs in chain order
    int32_t o_ch_q;        // [nch]    (qs_av+1) | (qe_av+1) << 16  compact V-bar index of a dangling end, 0 = none
    int32_t o_av_ptr;      // [nav+1]  matched in-arcs of each active V-bar node, in incomingArcs order
    int32_t o_av_arcs;     // [#matched]
    int32_t o_fb_ptr;      // [nav+1]  fallback chains (end-anchored chains starting at the node) for nodes without a matched pair
    int32_t o_fb_ch;       // [#fallback]
    int32_t o_ch_st;       // [nopen]    int2, the static half of an open chain in one 64-bit load (8-byte aligned):
                           //            x = sv | ev << 10 | hf << 20,  y = hb | reward << 10, where hf / hb are the label
                           //            indices of the heads of the forward (sv->ev, cost -r) and backward (ev->sv, cost +r)
                           //            residual arcs: an arc entering the root ends at index nc (the root seen as a path END)
    // in-slots of the lane-per-scenario kernel (k1_lane.cu): the residual arcs of the contracted graph sorted by HEAD label
    // index (1..nc; an arc entering the root ends at index nc, the root seen as a path END)
    int32_t o_slots;       // int2 per slot, 8-byte aligned: x = tail | (2*chain+dir) << 16, y = cost << 12 | head
                           //   (dir 0: the chain's forward arc sv->ev at cost -r; dir 1: its backward arc ev->sv at cost +r)
    int32_t o_in_pd;       // [nc+2]     first slot | in-degree << 16 of every head index (index 0, the root as a source, has none)
    int32_t max_indeg;     // largest in-degree (decides the cursor width of the lane kernel)
    int32_t o_arc_av;      // [2m]     int2 per arc: x = compact index of the active V-bar node the arc enters as a MATCHED in-arc (-1 otherwise),
                           //          y = its matched out-arc (cut kernel: lambda / mu per arc)
    int32_t total;         // int32 words used by this plan, header included (the header is an even number of words)
};

static_assert(sizeof(PlanHeader) % 8 == 0, "64-bit records follow the header");

struct Plan {
    std::vector<int32_t> words;   // PlanHeader followed by the arrays
    std::vector<int32_t> match_out, match_in;
    int nch = 0, nopen = 0, max_indeg = 0;
};

// returns 0 / SGUFP_ERR_ARG / SGUFP_ERR_MATCHING / SGUFP_ERR_CYCLIC
// lane_tables: also emit the head-sorted slot tables that only the lane-per-scenario kernel reads (k1_lane.cu)
int build_plan(const Model &M, const int16_t *path, int plen, Plan &P, std::string &err, bool lane_tables = true);

// Link of two consecutive candidates of a batch (k1_cut.cu: warm start of candidate k from the optimal flow of candidate
// k-1 on the same scenario; consecutive paths of the Benders loop differ in a few layers, NodeExplorer.cpp:949-971).
// out = { n_removed, nopen(prev), prev[nopen(cur)], removed[n_removed], ends[n_removed] }: prev[c] = the open chain of `prev`
// made of the same arcs as open chain c of `cur` (-1: none), removed = the open chains of `prev` that `cur` does not have,
// ends = their contracted end nodes sv | ev << 10.  out stays empty when more than a quarter of the chains differ (a cold
// start is cheaper then).
void link_plans(const Plan &prev, const Plan &cur, std::vector<int32_t> &out);

// The order in which the flow kernel takes the candidates of a batch: a nearest-neighbour chain by the number of layers in which
// two paths differ, from `start` (the path whose state the handle holds; nullptr: from candidate 0).  The cuts do not depend on
// it; a warm start costs what the step changes, and the paths a Benders loop emits at one node stay within a few layers of EACH
// OTHER, not only of their predecessor (the 16 committed C4 emissions: 2 - 8 layers between any two), so the chain's steps are
// shorter than the emission's (profiles/r02_k1_flow_census.md).  Batches of more than 128 candidates keep the order given.
void order_batch(const int16_t *paths, int K, int L, const int16_t *start, std::vector<int32_t> &order);

void ray_arrays(const Model &M, const Plan &P, std::vector<int32_t> &ts, std::vector<int32_t> &hs, std::vector<int32_t> &info,
                std::vector<int32_t> &pair_layer, std::vector<int32_t> &next, std::vector<int32_t> &aq, std::vector<int32_t> &first_wire, int &nn);

// Inavap::Cut hash (Cut.h:243-251)
uint64_t cut_hash(const uint64_t *keys, const double *vals, int nnz);

}  // namespace sgufp
