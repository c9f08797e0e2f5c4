// k1_cut.cuh — launch interface of the scenario-cut kernels (implementation in k1_cut.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace sgufp {

struct K1Launch {
    const double *cap_u, *cap_l;  // [S][m_pad] scenario-major fp64 in HBM
    int S, m, m_pad;
    long long scen_offset;        // global index of local scenario 0
    const int32_t *plans;         // device pool of K plans
    const int32_t *plan_off;      // [K] word offset of each plan
    const int32_t *link_off;      // [K] word offset of the link of candidate k to candidate k-1 (model.hpp: link_plans), -1 = none; or nullptr
    const int32_t *order;         // [K] the order in which a run takes the candidates (link_off[order[j]] is the step from order[j-1]); or nullptr: as given
    int group;                    // candidates per work item: a warp takes `group` consecutive candidates on one scenario and warm-starts
                                  // each from the one before (0 or 1: every candidate from zero flow)
    int K, W, L;
    // State kept between launches on one handle (one-path-at-a-time callers: the reference's Benders loop hands solveSubProblem
    // one path per iteration, NodeExplorer.cpp:949-971): per scenario { valid, pot[nc], x[open chains] } of the LAST candidate
    // of the previous launch, so that the first candidate of this one can be warm-started too.  Only when the whole batch
    // is one run per scenario (group >= K): the warp that reads a scenario's row is the one that rewrites it.
    int32_t *state;               // [S][state_stride] or nullptr: { valid, lab[nc + 1], x[open chains] }
    int state_stride;             // >= 2 + nc + max open chains
    int32_t *xout;                // [K][S][xstride]: the optimal flow per open chain, flow kernel -> cut kernel (x[0] = INT_MIN: infeasible scenario)
    int xstride;                  // >= max(1, max open chains)
    int state_io;                 // bit 0: the link of the FIRST candidate taken (order[0]) is the step from the stored candidate, read the row; bit 1: write it
    unsigned long long *sums;     // [K][W] exact integer accumulators (two's complement)
    long long *first_inf;         // [K] lowest infeasible global scenario (init LLONG_MAX)
    unsigned long long *work;     // work-item queue of the warp kernel (one word, zeroed before the launch) or nullptr: fixed assignment
    double *obj;                  // [K][S] or nullptr
    uint8_t *status;              // [K][S] or nullptr
    int max_nch, max_nopen, nc, nav;
    int max_cap;                  // largest capacity of this handle's scenarios (the lane kernel keeps 8- or 16-bit flow state)
    // what decides whether the lane-per-scenario kernel (k1_lane.cu) takes the batch
    int has_lower;                // some lower bound of this handle's scenarios is positive (then: warp kernel)
    int lane_tables;              // the plans carry the head-sorted in-slot tables
    int sum_abs_r;                // sum |reward| of the instance (bounds every label)
    int max_indeg;                // largest in-degree of a contracted node over the batch
};

// K1 dispatch: the warp-per-scenario kernel (k1_cut.cu) unless SGUFP_K1_MODE=lane asks for the lane-per-scenario kernel
// (k1_lane.cu), which then takes every batch it accepts: no positive lower bounds, capacities < 65536, state fits shared memory.
bool k1_lane_tables_wanted();                               // false when the mode rules the lane kernel out
bool k1_lane_eligible(const K1Launch &p, int sm_count);
cudaError_t k1_lane_launch(const K1Launch &p, cudaStream_t st, int sm_count);   // cudaErrorInvalidConfiguration: state too large

int k1_group(int K, int S, int sm_count);                   // K1Launch::group for a batch of K candidates on S scenarios

// Returns cudaSuccess or the launch error.  *launches is incremented by the kernels launched.
// state_kept: the launch left K1Launch::state describing the last candidate (false for the lane kernel, which ignores it)
cudaError_t k1_launch(const K1Launch &p, cudaStream_t st, int sm_count, int *launches, bool *state_kept = nullptr);

struct RayLaunch {
    const double *cap_u, *cap_l;
    int s_local, m, m_pad, nn;    // nn = split-graph nodes
    const int32_t *arc_ts, *arc_hs;   // [m] split-graph endpoints, -1 = dangling end
    const int32_t *arc_info;          // [m] kind | (layer+1)<<2 with RAY kinds
    const int32_t *arc_pair_layer;    // [m] for a matched in-arc: its layer, else -1
    const int32_t *arc_next;          // [m] matched out-arc of a matched in-arc, else -1
    const int32_t *arc_q;             // [m] (tail_av+1) | (head_av+1)<<16
    const int32_t *av_first_wire;     // [nav] split node of the first matched pair, -1 if none
    int nav, L;
    int32_t *scratch;                 // device scratch, >= 8*(nn+m)+64 words
    unsigned long long *sums;         // [W]
};
cudaError_t ray_launch(const RayLaunch &p, cudaStream_t st, int *launches);

// scenario-major fp64 re-layout of a slab of the reference's arc-major int32 capacity arrays: arcs a0 .. a0+na
cudaError_t relayout_launch(const int32_t *src /*[na][S] device*/, double *dst /*[S][m_pad]*/, int na, int S, int m_pad, int a0,
                            cudaStream_t st, int *launches);

}  // namespace sgufp
