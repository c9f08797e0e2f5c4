// k2_dd.cuh — launch interface of the DD longest-path kernels (implementation in k2_dd.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace sgufp {

struct K2DD {   // device image of one decision diagram (CSR by layer)
    const int4 *layer_info;     // [nlayers]  {first node, first in-arc, width, 1 if every node has exactly one in-arc}
    const int32_t *in_ptr;      // [nnodes+1] in-arc offsets (only read for layers that are not uniform)
    const int2 *arc_ts;         // [narcs]    {position of the tail in the previous layer, coefficient slot or -1}
    const int32_t *root_slot;   // [nroot]    slots of the fixed prefix decisions
    double *term;               // [nlast]    terminal arc weights: min over every cut applied so far
    long long state_off;        // offset of this DD's full-state block (states of the LAST cut of a batch), in doubles
    long long last_off;         // offset of this DD's [C][nlast] block of last-layer states, in doubles
    int nlayers, nroot, nnodes, nlast, max_width;
    const uint8_t *arc_dead;    // [narcs] or null: arcs removed on the device since the last upload (k2_finish)
};

// One CTA per (cut c, diagram b).  Node states ping-pong between two shared-memory buffers (a layer
// only reads the previous one); the cut's dense coefficient vector is staged in shared memory.
// Only the last layer goes to global memory (last[b][c][i]); the states of every node are written
// for cut C-1 alone, which is what the host semantics read back.
cudaError_t k2_launch(const K2DD *dds_device, int B, const double *coef_device /*[C][Tpad]*/, const double *rhs_device /*[C]*/,
                      int C, int Tpad, double *states_device, double *last_device, int max_width_all, int max_layers, int avg_width, cudaStream_t st,
                      int *launches);   // max_layers: the per-layer records are staged in shared memory when <= 640;
                                        // avg_width (nodes per layer of the widest diagram) picks the threads per CTA
// true if the widest layer fits the shared-memory state buffers (else: global state blocks, one per cut)
bool k2_states_in_smem(int Tpad, int max_width_all, int avg_width);
// term[i] = min(term[i], min_c last[b][c][i]); bound[b] = max_i term[i]
cudaError_t k2_terminal_launch(const K2DD *dds_device, int B, int C, const double *last_device, double *bound_device,
                               double *partial_device /*[B * 64] scratch*/, int max_last, cudaStream_t st, int *launches);

// ---- device-side cut application ----------------------------------------------
// What follows the longest path in RelaxedDDNew / RestrictedDDNew::applyOptimalityCut and
// applyFeasibilityCut (DD.cpp:3975-4021, 3884-3929, 3398-3422, 3493-3504), on the device image:
// terminal weights, removal of last-layer nodes (with the bottom-up cascade of DD.cpp:4040-4160),
// bound-based arc pruning of non-exact diagrams, and getSolution / getMaxPath (DD.cpp:3825-3840,
// 3796-3820, 3290-3305).  Removed arcs and nodes are FLAGGED (the CSR stays as uploaded); the host
// mirror replays the flags when somebody asks for the structure.
struct K2Result {
    double bound;        // optimality: the upper / lower bound returned by the reference
    int feasible;        // feasibility: the uint8_t the reference returns
    int changed;         // 1 if arcs or nodes were removed
    int path_len;        // extract: decisions written
};
struct K2Apply {
    K2DD d;                       // d.term, d.arc_dead: this diagram's persistent buffers
    const double *state;          // [nnodes] states of the cut just applied
    const double *coef;           // [Tpad]   its dense coefficient vector
    const int32_t *arc_dec;       // [narcs]  decision of every arc (out-arc id or -1)
    uint8_t *arc_dead, *node_dead;   // [narcs], [nnodes]
    int32_t *layer_alive;         // [nlayers] nodes still in each layer
    int32_t *cnt;                 // [nnodes] scratch
    uint8_t *lost;                // [nnodes] scratch, zero on entry and on exit
    K2Result *out;
    int16_t *path;                // [nlayers] extract: decisions root -> leaf (after the fixed prefix)
    double optimal;
    int mode;                     // 0 optimality, 1 feasibility
    int restricted, exact;
};
// single-cut longest path of ONE diagram that honours arc_dead; every node state goes to `states`
cudaError_t k2_single_launch(const K2DD *dd_device, const double *coef_device, const double *rhs_device, int Tpad, double *states,
                             double *last, int max_width, int max_layers, cudaStream_t st, int *launches);
// the same for a WIDE diagram: one launch per layer over the whole GPU; states[0] (the root) must already be set.
// layer_width_host / layer_collapsed_host: per layer, the node count and whether it is one node with several in-arcs.
cudaError_t k2_layered_launch(const K2DD &d, const int32_t *layer_width_host, const uint8_t *layer_collapsed_host, const double *coef_device,
                              double *states, cudaStream_t st, int *launches);
cudaError_t k2_finish_launch(const K2Apply &a, cudaStream_t st, int *launches);
// a run of cuts on one diagram (k2_finish_seq)
struct K2Seq {
    const double *coef;     // [C][Tpad] all cuts of the call
    double *states;         // [k1-k0][nnodes] states of cuts k0.., computed side by side on the current structure
    K2Result *results;      // [C]
    int *ctl;               // [2]: cuts consumed so far; 1 if the caller's loop ends there
    double *last;           // [k1-k0][nlast] last-layer states of the window's cuts (written by the longest-path kernel)
    double *bounds;         // [k1-k0] or null: speculative bound of every cut of the window (k2_window_bounds)
    int *probe;             // [k1-k0] or null: per cut, 1 if its optimality pruning would touch the diagram (k2_prune_probe)
    int k0, k1, Tpad;
};
cudaError_t k2_sequence_launch(const K2DD *dd_device, const K2Apply &a, const K2Seq &q, const double *rhs_device, int max_width,
                               double *last_scratch /*[(k1-k0)*nlast]*/, cudaStream_t st, int *launches);
cudaError_t k2_extract_launch(const K2Apply &a, cudaStream_t st, int *launches);

// ---- construction on the device ---------------------------------------
struct K2Tables {               // static per network, uploaded once per context
    const int32_t *lay_tab;     // [L]      state table (V-bar node) of each global layer
    const uint8_t *lay_first;   // [L]      1 at the first layer of a V-bar node: states reset (Network.cpp:96-102)
    const int32_t *tab_ptr;     // [ntab+1]
    const int16_t *tab_dec;     // base list of a table: -1, then the node's out-arcs ascending
    const int16_t *tab_k;       // coefficient slot of that decision relative to slot_base[layer], -1 for none
    const int32_t *slot_base;   // [L]
    int L;
};
struct K2BuildOut { int nlayers, nnodes, narcs, exact, exact_layer, overflow, max_width, nlast; };
struct K2Build {
    K2Tables t;
    int start;                  // global layer of the root (== number of fixed prefix decisions)
    unsigned root_mask;         // the root's states over the base list of layer `start`
    int restricted, max_width;
    int node_cap, arc_cap;
    int4 *layer_info; int32_t *in_ptr; int2 *arc_ts; int32_t *arc_dec; unsigned *mask; int32_t *off /*[node_cap] scratch*/;
    int32_t *widths;            // [L - start + 1]
    K2BuildOut *out;
};
cudaError_t k2_build_launch(const K2Build &b, cudaStream_t st, int *launches);
cudaError_t k2_fill_launch(double *p, double v, long long n, cudaStream_t st);
// relaxed cut-set on the device image: out4 = {layer, node, first in-arc, end in-arc} of the first layer >= 3 with one live node
cudaError_t k2_first_collapsed_launch(const K2Apply &a, int *out4, cudaStream_t st);
// per in-arc j of that node: tail's state mask, number of decisions (-1: arc removed) and the decisions leaf first
cudaError_t k2_relaxed_cutset_launch(const K2Apply &a, const unsigned *mask, int layer, int v, int e0, int count, int stride, unsigned *out_mask,
                                     int32_t *out_len, int16_t *out_dec, cudaStream_t st);
// restricted tree: state mask + decisions (root first, `el` per node) of every node of layer `el`
cudaError_t k2_cutset_launch(const int4 *layer_info, const int32_t *in_ptr, const int2 *arc_ts, const int32_t *arc_dec, const unsigned *mask,
                             int el, int count, unsigned *out_mask, int16_t *out_dec, cudaStream_t st);

}  // namespace sgufp
