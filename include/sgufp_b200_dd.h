/*
 * sgufp_b200_dd.h — C ABI of the decision-diagram half of the hot path (K2).
 *
 * Replaces the members of `Inavap::RelaxedDDNew` (/root/reference/DD.h:797-808, DD.cpp:3528-4229)
 * and `Inavap::RestrictedDDNew` (/root/reference/DD.h:713-728, DD.cpp:3090-3505) that
 * NodeExplorer::process calls (/root/reference/NodeExplorer.cpp:922-981).  The diagram is built on
 * the host exactly as the reference builds it; the layer-wise longest path with cut-adjusted arc
 * weights runs on the GPU.  Same conventions as sgufp_b200.h: host pointers, 0 / negative codes.
 * A diagram belongs to the sgufp_ctx it was created from (network model, device, stream) and, like
 * the reference's, to one host thread.
 */
#ifndef SGUFP_B200_DD_H
#define SGUFP_B200_DD_H
#include "sgufp_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

#define SGUFP_DD_RELAXED 0    /* RelaxedDDNew: collapse threshold = max_width, or RELAXED_MAX_WIDTH = 120 (DD.h:732) if max_width <= 0 */
#define SGUFP_DD_RESTRICTED 1 /* RestrictedDDNew(max_width) (DD.h:710) */

typedef struct sgufp_dd sgufp_dd;

int sgufp_dd_create(sgufp_ctx *ctx, int kind, int max_width, sgufp_dd **out);
void sgufp_dd_destroy(sgufp_dd *dd);

/* buildTree(Node) (DD.cpp:3528) / compile(Node) (DD.cpp:3090).  The node is (states, solutionVector,
 * globalLayer) of `Inavap::Node` (DD.h:456-479); an empty node with globalLayer 0 is the root.
 * cutset_nodes (may be NULL): restricted DD only, the size of the exact cut-set, or -1 if the tree
 * is exact (DD.cpp:3157-3158); fetch it with sgufp_dd_cutset. */
/* buildTree(node) (DD.cpp:3528-3600) / compile(node) (DD.cpp:3090-3159).  The diagram is built ON THE
 * DEVICE (state sets as 32-bit masks, one prefix sum + one expansion per
 * layer, CSR emitted in place) whenever the root is a cut-set node (solution length == globalLayer, states a
 * subset of the layer's {-1} U outgoingArcs) and no V-bar node has more than 31 out-arcs; the host
 * mirror is then constructed lazily, only if dump / cutset / layer_sizes / counts / the batch call need
 * it.  SGUFP_DD_BUILD=host forces the host builder. */
int sgufp_dd_build(sgufp_dd *dd, const int16_t *states, int nstates, const int16_t *solution, int nsolution,
                   int global_layer, int *cutset_nodes);
int sgufp_dd_is_exact(const sgufp_dd *dd); /* isTreeExact(): 1 / 0 */

/* Structure, for inspection and parity tests.  Layers exclude the terminal node; arcs include the
 * terminal arcs.  dump: nodes in tree order; per in-arc the POSITION of its tail in the previous
 * layer and its decision, in stored in-arc order (node_inptr has nodes+1 entries). */
int sgufp_dd_num_layers(const sgufp_dd *dd);
int sgufp_dd_layer_sizes(sgufp_dd *dd, int32_t *sizes /*[num_layers]*/);
int sgufp_dd_counts(sgufp_dd *dd, int64_t *nodes, int64_t *arcs);
int sgufp_dd_dump(sgufp_dd *dd, int32_t *node_layer, double *node_state, int64_t *node_inptr, int32_t *arc_tailpos,
                  int32_t *arc_decision, double *terminal_weight /*[last layer]*/);

/* Test introspection: the CSR image as it sits on the device (built there by k2_build when the root is a
 * cut-set node and state sets fit 32 bits, uploaded from the host mirror otherwise).  Sizes as for
 * sgufp_dd_dump, which answers from the host mirror.  Returns 1 if the image was built on the device. */
int sgufp_dd_dump_device(sgufp_dd *dd, int32_t *layer_sizes, int64_t *node_inptr, int32_t *arc_tailpos, int32_t *arc_decision,
                         int32_t *arc_slot);

/* applyOptimalityCut(cut, optimal, upperbound) -> bound (DD.cpp:3932-4023, 3425-3505) and
 * applyFeasibilityCut(cut) -> feasible (DD.cpp:3842-3930, 3340-3423).  The cut is an Inavap::Cut:
 * RHS + nnz (key,value) pairs.  `optimal`/`upperbound` are ignored by the restricted DD, whose
 * reference signature has none.
 * The whole call runs on the device: longest path, terminal weights, removal of
 * last-layer nodes with its bottom-up cascade, bound-based arc pruning of non-exact diagrams.  Only
 * the bound / the flag come back; removed arcs and nodes are flagged in the device image and the
 * host mirror replays them when layer_sizes / counts / dump / cutset / the batch call ask for it. */
int sgufp_dd_apply_optimality(sgufp_dd *dd, double rhs, const uint64_t *keys, const double *vals, int nnz,
                              double optimal, double upperbound, double *bound);
int sgufp_dd_apply_feasibility(sgufp_dd *dd, double rhs, const uint64_t *keys, const double *vals, int nnz,
                               int *feasible);

/* A RUN of cuts of one kind on one diagram, as the loops of NodeExplorer::process apply the global cuts to a
 * fresh diagram (NodeExplorer.cpp:935-944, 975-983): mode 0 = applyOptimalityCut(cut, optimal, .) for every cut,
 * mode 1 = applyFeasibilityCut(cut).  Stops after the first cut whose bound <= optimal (mode 0) or that is
 * infeasible (mode 1) — where the caller's loop returns — and reports how many cuts were applied; bound[k] /
 * feasible[k] are the values the reference's calls return (whichever array the mode fills; the other may be
 * NULL).  The result is exactly that of the one-by-one calls: the longest paths of the cuts are computed side
 * by side, the sequential part (terminal weights, removals, pruning) in order, and whenever a cut changes the
 * structure the cuts behind it are recomputed on the new one. */
int sgufp_dd_apply_sequence(sgufp_dd *dd, int mode, const double *rhs, const uint64_t *keys, const double *vals,
                            const int32_t *cut_ptr, int C, double optimal, double *bound /*[C]*/, int *feasible /*[C]*/,
                            int *applied);

/* getSolution() / getMaxPath() (DD.cpp:3825-3840, 3290-3305): returns the path length.  After a
 * single-cut apply the path is extracted on the device; only its L int16 decisions come back. */
int sgufp_dd_solution(sgufp_dd *dd, int16_t *path, int capacity);
/* getCutset(ub) (DD.cpp:4179-4218) for the relaxed DD; the cut-set of the last compile for the
 * restricted DD.  Nodes are flattened as (globalLayer, #states, states..., #solution, solution...);
 * returns the number of int32 words written or SGUFP_ERR_ARG if `capacity` is too small. */
int sgufp_dd_cutset(sgufp_dd *dd, double ub, int32_t *words, int capacity);

/* Batched K2: apply the same C optimality cuts to B diagrams in ONE launch pair.
 * cut c is rhs[c] + (keys,vals)[cut_ptr[c] .. cut_ptr[c+1]).  For every diagram the terminal arc
 * weights become min(previous, min over the C cuts) and bound[b] = max over terminal arcs — what C
 * sequential applyOptimalityCut calls leave behind on an exact diagram.  Bound-based arc pruning of
 * non-exact diagrams (DD.cpp:3987-4021) is NOT performed here; node states are those of cut C-1. */
int sgufp_dd_apply_optimality_batch(sgufp_dd **dds, int B, const double *rhs, const uint64_t *keys, const double *vals,
                                    const int32_t *cut_ptr, int C, double *bound /*[B]*/);
/* Device time of the last K2 launch pair on the diagram's context, and arcs touched by it
 * (sum over diagrams and cuts of in-arcs + terminal arcs). */
int sgufp_dd_last_stats(const sgufp_dd *dd, float *kernel_ms, int64_t *arcs_touched, int *kernel_launches);

#ifdef __cplusplus
}
#endif
#endif /* SGUFP_B200_DD_H */
