// ctx.hpp — the handle behind sgufp_ctx (shared by capi.cu and capi_dd.cu).
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "../../include/sgufp_b200.h"
#include "host_pool.hpp"
#include "model.hpp"

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, (n ? n : 1) * sizeof(T));
        if (e == cudaSuccess) cap = n ? n : 1;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// The plans of one batch of candidate paths (K1).  The handle keeps the last batch: the sharded sequence
// sgufp_paths_partial -> all-reduce -> sgufp_finalize_paths hands the same paths to both calls, and a plan is 15 us (C2)
// to 65 us (C4) of host work.
struct PlanBatch {
    std::vector<sgufp::Plan> plans;
    std::vector<int32_t> off;
    std::vector<std::vector<int32_t>> links;   // links[k]: candidate k seen from candidate k-1 (model.hpp: link_plans), empty = none
    std::vector<int32_t> link_off;   // [K] word offset of links[k] in the pool behind the plans, -1 = none
    std::vector<int32_t> order;      // [K] the order in which a run takes the candidates (model.hpp: order_batch); links[order[j]] is seen from order[j-1]
    size_t total_words = 0;          // the plans are gathered into the handle's pinned staging buffer at launch
    int max_nch = 0, max_nopen = 0, max_indeg = 0;
    std::vector<int16_t> key_paths;  // what the plans were built from (valid when key_K > 0)
    int key_K = 0, key_L = -1;
    bool key_lane = false;
};

struct Partition;                     // capi_shard.cu: the communicator(s) and peer handles of a scenario partition

// The scenario-major capacity arrays in HBM, shared read-only by every handle cloned from the one that uploaded them
// (sgufp_clone: N_WORKERS + 1 host threads each own a GuroSolver, NodeExplorer.h:115-116 / DDSolver.cpp:583,675).
struct CapStore {
    double *d_u = nullptr, *d_l = nullptr;
    int device = 0;
    int refs = 1;                     // guarded by cap_store_mutex() (capi.cu)
};

struct sgufp_ctx {
    sgufp::Model M;
    Partition *part = nullptr;
    int S = 0, m_pad = 0, device = 0, sm_count = 0, max_cap = 0, max_lower = 0, sum_abs_r = 0;
    long long scen_off = 0, S_total = 0;
    long long S_view = -1;                      // leader of a single-process partition: the scenarios the caller sees (all of them)
    double *d_u = nullptr, *d_l = nullptr;      // aliases of caps->d_u / d_l
    CapStore *caps = nullptr;
    cudaStream_t st = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evk0 = nullptr, evk1 = nullptr;
    DevBuf<int32_t> d_plans, d_plan_off, d_ray_i32;
    DevBuf<unsigned long long> d_sums, d_work;
    DevBuf<long long> d_finf, d_ray_scratch;
    DevBuf<double> d_obj;
    DevBuf<uint8_t> d_status;
    std::string err;
    PlanBatch batch;                            // plans of the last batch (reused when the next call brings the same paths)
    // K1 state kept between launches (k1_cut.cuh: K1Launch::state): per scenario the optimal flow and potentials of the last
    // candidate of the previous launch, and that candidate's plan on the host (the next launch links its first plan to it)
    DevBuf<int32_t> d_state;
    DevBuf<int32_t> d_xout;                     // [K][S][max open chains]: optimal flows, flow kernel -> cut kernel
    sgufp::Plan state_plan;
    std::vector<int16_t> state_path;            // the path state_plan was built from (where the next batch's order starts)
    bool state_valid = false;
    int state_stride = 0;
    sgufp::HostPool *pool = nullptr;            // persistent host threads for the plans of a batch (created on first use)
    int32_t *h_words = nullptr;                 // pinned staging of a batch's plans (cudaHostAlloc), h_words_cap int32 words
    size_t h_words_cap = 0;
    long long *h_out = nullptr;                 // pinned landing buffer of a batch's sums + first-infeasible marks
    size_t h_out_cap = 0;
    cudaEvent_t ev_h2d = nullptr;               // recorded after the uploads out of h_words
    bool h2d_pending = false;
    int last_launches = 0;
    float last_ms = 0.f;
    bool kernel_timed = false;
    // K2 statistics
    float dd_kernel_ms = 0.f;
    long long dd_arcs = 0;
    int dd_launches = 0;
    void *dd_scratch = nullptr;                 // K2 batch scratch (capi_dd.cu), freed through dd_scratch_free
    void (*dd_scratch_free)(void *) = nullptr;
    int W() const { return 1 + M.L + M.m; }
};

#define CU(ctx, call)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                      \
            return SGUFP_ERR_CUDA;                                                                 \
        }                                                                                          \
    } while (0)

static inline int fail(sgufp_ctx *c, int code, const std::string &msg) { c->err = msg; return code; }
