// k1_emul.cpp — TEST INFRASTRUCTURE.  The body of the K1 kernels (sgufp_solver_b200/csrc/k1_cut.cu)
// compiled for the HOST with one lane per scenario (TILE = 1) so that the kernel's algorithm —
// plan decoding, level-wise shortest paths, blocking flow, SPEC-LP potentials, lifting — can be
// checked against Oracle B in the CPU test suite, where there is no GPU.  The warp-level parallel
// execution itself is only exercised by the `-m gpu` tests.  Never part of the product.
#define SGUFP_K1_EMULATE
#define __CUDA_RUNTIME_H__          // skip <cuda_runtime.h>
#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

typedef int cudaError_t;
typedef void *cudaStream_t;
struct double2 { double x, y; };
struct int2 { int x, y; };
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
struct int4 { int x, y, z, w; };
struct EmulIdx { int x = 0, y = 0; };
static EmulIdx threadIdx, blockIdx;
static struct { int x = 1, y = 1; } gridDim, blockDim;
static int *sgufp_emul_smem = nullptr;
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__
using std::max;
using std::min;
static inline int atomicMin(int *p, int v) { int o = *p; if (v < o) *p = v; return o; }
static inline int atomicMax(int *p, int v) { int o = *p; if (v > o) *p = v; return o; }
static inline int atomicAdd(int *p, int v) { int o = *p; *p += v; return o; }
static inline int atomicSub(int *p, int v) { int o = *p; *p -= v; return o; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }
static inline long long atomicMin(long long *p, long long v) { long long o = *p; if (v < o) *p = v; return o; }
static inline unsigned atomicOr(unsigned *p, unsigned v) { unsigned o = *p; *p |= v; return o; }
static inline unsigned atomicAnd(unsigned *p, unsigned v) { unsigned o = *p; *p &= v; return o; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
static inline unsigned __reduce_or_sync(unsigned, unsigned v) { return v; }
static inline int __reduce_min_sync(unsigned, int v) { return v; }
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline int __any_sync(unsigned, int p) { return p; }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int, int = 32) { return v; }
template <class T> static inline T __shfl_sync(unsigned, T v, int, int = 32) { return v; }
static inline double2 __ldg(const double2 *p) { return *p; }
static inline double2 __ldcs(const double2 *p) { return *p; }

#include "../../sgufp_solver_b200/csrc/k1_cut.cu"
#include "../../sgufp_solver_b200/csrc/k1_lane.cu"
#include "../../sgufp_solver_b200/csrc/model.cpp"

using namespace sgufp;

// Same contract as sgufp_paths_partial + sgufp_ray_partial, on the host.  sums: [K][W] (W = 1+L+m),
// first_inf: [K], obj/status: [K][S].  ray_sums[K][W]: filled for candidates with an infeasible scenario.
// state between calls (K1Launch::state), as capi.cu: launch_batch keeps it on a handle: rows of the last candidate of the
// previous call + its plan.  emul_state(1) turns it on (and forgets what was stored), emul_state(0) off.
static bool g_state_on = false, g_state_valid = false;
static std::vector<int32_t> g_state;
static sgufp::Plan g_state_plan;
static std::vector<int16_t> g_state_path;
extern "C" void emul_state(int on) { g_state_on = on != 0; g_state_valid = false; g_state.clear(); g_state_path.clear(); }
static int g_order = 1;
extern "C" void emul_set_order(int on) { g_order = on; }   // 1 (as capi.cu: make_batch): runs take the candidates along a nearest-neighbour chain; 0: as given
static int g_group = 0;
extern "C" void emul_set_group(int g) { g_group = g; }   // candidates per work item (0: the whole batch, 1: no warm starts)
extern "C" void emul_warm_counts(long long *out2) { out2[0] = sgufp::sgufp_emul_warm[0]; out2[1] = sgufp::sgufp_emul_warm[1]; sgufp::sgufp_emul_warm[0] = sgufp::sgufp_emul_warm[1] = 0; }
extern "C" void emul_warm_counts8(long long *out8) { for (int i = 0; i < 8; i++) { out8[i] = sgufp::sgufp_emul_warm[i]; sgufp::sgufp_emul_warm[i] = 0; } }
extern "C" void emul_counts16(long long *out16) { for (int i = 0; i < 16; i++) { out16[i] = sgufp::sgufp_emul_cnt[i]; sgufp::sgufp_emul_cnt[i] = 0; } }   // tools/proto/flow_census.py
extern "C" void emul_order(const int16_t *paths, int K, int L, const int16_t *start, int32_t *order) {   // model.hpp: order_batch, as the product calls it
    std::vector<int32_t> o;
    sgufp::order_batch(paths, K, L, start, o);
    std::memcpy(order, o.data(), (size_t)K * 4);
}
static int g_last_nc = 0;
extern "C" int emul_last_nc() { return g_last_nc; }   // contracted nodes of the last instance: which size class ran

extern "C" int emul_paths(int n, int m, int S, const int32_t *tail, const int32_t *head, const int32_t *upper, const int32_t *lower,
                          const int32_t *rew0, const int32_t *vbar, int nvbar, const int16_t *paths, int K, int L,
                          long long *sums, long long *first_inf, double *obj, uint8_t *status, long long *ray_sums, int lane_variant) {
    Model M;
    std::string err;
    if (int rc = M.build(n, m, tail, head, rew0, vbar, nvbar, err)) return rc;
    const int m_pad = (m + 1) & ~1, W = 1 + M.L + m;
    std::vector<double> cu((size_t)S * m_pad, 0.0), cl((size_t)S * m_pad, 0.0);
    for (int a = 0; a < m; a++)
        for (int s = 0; s < S; s++) { cu[(size_t)s * m_pad + a] = upper[(size_t)a * S + s]; cl[(size_t)s * m_pad + a] = lower[(size_t)a * S + s]; }
    std::vector<Plan> plans(K);
    std::vector<int32_t> words, off(K);
    int max_nch = 0, max_nopen = 0;
    for (int k = 0; k < K; k++) {
        if (int rc = build_plan(M, paths + (size_t)k * L, L, plans[k], err)) return rc;
        off[k] = (int32_t)words.size();
        words.insert(words.end(), plans[k].words.begin(), plans[k].words.end());
        max_nch = std::max(max_nch, plans[k].nch); max_nopen = std::max(max_nopen, plans[k].nopen);
    }
    // the order of the candidates in a run and the links of neighbours in that order, behind the plans (as capi.cu: make_batch lays them out)
    std::vector<int32_t> order(K);
    for (int k = 0; k < K; k++) order[k] = k;
    if (g_order) order_batch(paths, K, L, g_state_on && g_state_valid && (int)g_state_path.size() == L ? g_state_path.data() : nullptr, order);
    std::vector<int32_t> link_off(K, -1);
    for (int j = 1; j < K; j++) {
        std::vector<int32_t> lk;
        link_plans(plans[order[j - 1]], plans[order[j]], lk);
        if (!lk.empty()) { link_off[order[j]] = (int32_t)words.size(); words.insert(words.end(), lk.begin(), lk.end()); }
    }
    const int group = g_group > 0 ? g_group : K;
    int state_io = 0, state_stride = 2 + M.nc + m;
    if (g_state_on && group >= K && !lane_variant) {
        if (g_state.size() != (size_t)S * state_stride) { g_state.assign((size_t)S * state_stride, 0); g_state_valid = false; }
        state_io = 2;
        if (g_state_valid) {
            std::vector<int32_t> lk;
            link_plans(g_state_plan, plans[order[0]], lk);
            if (!lk.empty()) { state_io |= 1; link_off[order[0]] = (int32_t)words.size(); words.insert(words.end(), lk.begin(), lk.end()); }
        }
    }
    std::memset(sums, 0, (size_t)K * W * 8);
    for (int k = 0; k < K; k++) first_inf[k] = LLONG_MAX;
    K1Launch p{};
    p.cap_u = cu.data(); p.cap_l = cl.data(); p.S = S; p.m = m; p.m_pad = m_pad; p.scen_offset = 0;
    p.plans = words.data(); p.plan_off = off.data(); p.K = K; p.W = W; p.L = M.L;
    p.sums = reinterpret_cast<unsigned long long *>(sums); p.first_inf = first_inf; p.obj = obj; p.status = status;
    unsigned long long work_queue = 0;
    p.work = (K & 1) ? &work_queue : nullptr;     // both ways of handing out work items are exercised
    p.link_off = link_off.data(); p.order = order.data(); p.group = group;   // runs of warm-started candidates (default: the whole batch is one run)
    p.state = state_io ? g_state.data() : nullptr; p.state_stride = state_stride; p.state_io = state_io;
    if (state_io) { g_state_plan = plans[order[K - 1]]; g_state_valid = true; g_state_path.assign(paths + (size_t)order[K - 1] * L, paths + (size_t)(order[K - 1] + 1) * L); } else g_state_valid = false;
    g_last_nc = M.nc;
    p.max_nch = max_nch; p.max_nopen = max_nopen; p.nc = M.nc; p.nav = M.nav; p.max_cap = 65535;
    p.xstride = std::max(1, max_nopen);
    std::vector<int32_t> xout((size_t)K * std::max(1, S) * p.xstride, 0);
    p.xout = xout.data();
    const int wpt = k1_words_per_tile(p);
    std::vector<int> smem((size_t)WARPS * 32 * wpt, 0);
    sgufp_emul_smem = smem.data();
    threadIdx.x = 0; blockIdx.x = 0; gridDim.x = 1;
    // one "thread" walks all items: with TILE = 1 the tile stride is WARPS*32, so visit every tile slot of the CTA
    if (lane_variant) {   // the lane-per-scenario kernel (k1_lane.cu) with one lane; lane_variant 1/2/3 = state widths small/mid/wide
        int max_lower = 0, max_cap = 0, max_indeg = 0;
        long long sum_abs = 0;
        for (size_t i = 0; i < (size_t)m * S; i++) { max_lower = std::max(max_lower, lower[i]); max_cap = std::max(max_cap, upper[i]); }
        for (int a = 0; a < m; a++) sum_abs += std::abs((long long)rew0[a]);
        for (int k = 0; k < K; k++) max_indeg = std::max(max_indeg, plans[k].max_indeg);
        if (max_lower > 0) return 99;                 // the lane kernel does not take instances with positive lower bounds
        p.max_cap = max_cap; p.has_lower = 0; p.lane_tables = 1; p.sum_abs_r = (int)sum_abs; p.max_indeg = max_indeg;
        threadIdx.x = 0; blockIdx.x = 0; gridDim.x = 1;
        auto run = [&](auto cfg) -> int {
            using Cfg = decltype(cfg);
            if (!lane_cfg_fits<Cfg>(p)) return 98;
            std::vector<int> lsm(lane_layout<Cfg>(max_nopen, M.nc, M.nav, 1).bytes / 4 + 16, 0);
            sgufp_emul_smem = lsm.data();
            k1_lane_pd<1, Cfg>(p);
            return 0;
        };
        const int rc = lane_variant == 1 ? run(CfgSmall{}) : lane_variant == 2 ? run(CfgMid{}) : run(CfgWide{});
        if (rc) return rc;
    } else
    {
        for (int t = 0; t < WARPS * 32; t++) { threadIdx.x = t; if (p.nc <= SMALL_NC) k1_cut_eval<1, WARPS, false>(p, wpt); else k1_cut_eval<1, WARPS, true>(p, wpt); }
        for (int t = 0; t < WARPS * 32; t++) { threadIdx.x = t; if (p.nc <= SMALL_NC) k1_cut_fold<1, WARPS, false>(p, wpt); else k1_cut_fold<1, WARPS, true>(p, wpt); }
    }
    for (int k = 0; k < K; k++) {
        if (first_inf[k] == LLONG_MAX || first_inf[k] < 0) continue;
        std::vector<int32_t> ts, hs, info, pl, nx, aq, fw;
        int nn = 0;
        ray_arrays(M, plans[k], ts, hs, info, pl, nx, aq, fw, nn);
        std::vector<long long> scratch((size_t)2 * nn + (size_t)(2 * m + 4 * nn) / 2 + 8);
        std::memset(ray_sums + (size_t)k * W, 0, (size_t)W * 8);
        RayLaunch r{};
        r.cap_u = cu.data(); r.cap_l = cl.data(); r.s_local = (int)first_inf[k]; r.m = m; r.m_pad = m_pad; r.nn = nn;
        r.arc_ts = ts.data(); r.arc_hs = hs.data(); r.arc_info = info.data(); r.arc_pair_layer = pl.data(); r.arc_next = nx.data(); r.arc_q = aq.data();
        r.av_first_wire = fw.data(); r.nav = M.nav; r.L = M.L; r.scratch = reinterpret_cast<int32_t *>(scratch.data());
        r.sums = reinterpret_cast<unsigned long long *>(ray_sums + (size_t)k * W);
        threadIdx.x = 0; blockIdx.x = 0;
        k1_ray(r);
    }
    return 0;
}
