// capi.cu — the C ABI of include/sgufp_b200.h: handle, HBM layout, launches, result conversion.
//
// Host-side mirror of GuroSolver (/root/reference/grb.h:17-104).  There is no CPU fallback:
// every compute entry point needs a CUDA device and fails with SGUFP_ERR_CUDA otherwise.
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "capi_internal.hpp"
#include "ctx.hpp"
#include "k1_cut.cuh"

using namespace sgufp;

namespace {
thread_local std::string g_create_error;
}  // namespace

// ---- helpers shared by the one-GPU and the sharded entry points (capi_internal.hpp) -------------

// Host threads that build plans of one batch side by side (a plan is ~15 us of host work at C2, ~65 us at C4:
// serial construction of 64 plans would cost a third of the kernel).  The threads persist in the handle.
int host_threads(int K) {
    int cap = 8;
    const int hw = (int)std::thread::hardware_concurrency();
    if (hw > 0) cap = std::min(cap, hw);
    // several ranks of a partition on one box (torchrun, mpirun export the local world size) share its cores: a rank that
    // oversubscribes them delays the all-reduce of everybody (measured on 8 ranks / 32 cores: 2.80 ms per C2 step with 8 pool
    // threads per rank, 2.58 - 2.69 ms with 1 - 4)
    for (const char *name : {"LOCAL_WORLD_SIZE", "OMPI_COMM_WORLD_LOCAL_SIZE", "MPI_LOCALNRANKS"})
        if (const char *e = getenv(name)) { const int lw = atoi(e); if (lw > 1 && hw > 0) cap = std::min(cap, std::max(1, hw / (2 * lw))); break; }
    if (const char *e = getenv("SGUFP_HOST_THREADS")) cap = std::max(1, atoi(e));
    return std::max(1, std::min(cap, K / 4));
}

// reuse: the caller is the second half of one operation on these paths (sgufp_finalize_paths after sgufp_paths_partial)
// and may take the plans the first half built; every other entry point builds its plans anew, whatever the last call was.
int make_batch(sgufp_ctx *c, const int16_t *paths, int K, int L, Batch &B, bool reuse) {
    if (!paths || K < 1 || L < 0 || L > c->M.L) return fail(c, SGUFP_ERR_ARG, "paths: need K >= 1 and 0 <= L <= totalLayers");
    const bool lane_tables = k1_lane_tables_wanted() && c->max_lower == 0;
    const size_t npath = (size_t)K * L;
    if (reuse && B.key_K == K && B.key_L == L && B.key_lane == lane_tables && B.key_paths.size() == npath &&
        (npath == 0 || memcmp(B.key_paths.data(), paths, npath * sizeof(int16_t)) == 0))
        return 0;                    // the same paths as the first half of the operation: its plans stand
    B.key_K = 0;
    B.max_nch = B.max_nopen = B.max_indeg = 0;
    B.plans.resize(K); B.off.resize(K);
    std::vector<int> rc(K, 0);
    std::vector<std::string> errs(K);
    const std::function<void(int, int)> work = [&](int t, int nt) {
        for (int k = t; k < K; k += nt) rc[k] = build_plan(c->M, paths + (size_t)k * L, L, B.plans[k], errs[k], lane_tables);
    };
    const int nt = host_threads(K);
    if (nt > 1) {
        if (!c->pool || c->pool->size() < nt) { delete c->pool; c->pool = new HostPool(nt - 1); }
        c->pool->run(nt, work);
    } else work(0, 1);
    for (int k = 0; k < K; k++)
        if (rc[k]) return fail(c, rc[k], "path " + std::to_string(k) + ": " + errs[k]);   // the lowest failing index, as a serial loop would report
    // the order in which a run takes the candidates, and the links between neighbours in that order (warm starts inside a work
    // item of the warp kernel, k1_cut.cu): the chain starts at the path whose state the handle holds, if it holds one
    // Laying the chain costs ~0.2 ns x K^2 L on one host thread (64 candidates of C2: ~60 us), the shorter steps save ~6 % of
    // K x S evaluations of 16 - 70 ns each: it pays when K L is small against S (C4, C5: 4 304 against 10 000 - 100 000 scenarios),
    // not for 64 candidates on 1000 scenarios of C2 — those batches keep the order given.  SGUFP_K1_ORDER=0 / 1 decides for all.
    const char *order_env = getenv("SGUFP_K1_ORDER");
    const bool order_on = order_env ? atoi(order_env) != 0 : (long long)K * L <= 4LL * std::max(1, c->S);
    if (!order_on) { B.order.resize(K); for (int k = 0; k < K; k++) B.order[k] = k; }
    else order_batch(paths, K, L, c->state_valid && (int)c->state_path.size() == L ? c->state_path.data() : nullptr, B.order);
    B.links.resize(K); B.link_off.assign(K, -1);
    for (int k = 0; k < K; k++) B.links[k].clear();
    if (K > 1) {
        const std::function<void(int, int)> link = [&](int t, int nt) {
            for (int j = 1 + t; j < K; j += nt) link_plans(B.plans[B.order[j - 1]], B.plans[B.order[j]], B.links[B.order[j]]);
        };
        if (nt > 1) c->pool->run(nt, link); else link(0, 1);
    }
    size_t total = 0;
    for (int k = 0; k < K; k++) {
        B.off[k] = (int32_t)total;
        total += B.plans[k].words.size();
        B.max_nch = std::max(B.max_nch, B.plans[k].nch);
        B.max_nopen = std::max(B.max_nopen, B.plans[k].nopen);
        B.max_indeg = std::max(B.max_indeg, B.plans[k].max_indeg);
    }
    for (int k = 0; k < K; k++)
        if (!B.links[k].empty()) { B.link_off[k] = (int32_t)total; total += B.links[k].size(); }
    B.total_words = total;
    B.key_paths.assign(paths, paths + npath); B.key_K = K; B.key_L = L; B.key_lane = lane_tables;
    return 0;
}

int launch_batch(sgufp_ctx *c, const Batch &B, int K, unsigned long long *d_sums, long long *d_finf, double *d_obj,
                        uint8_t *d_status, cudaStream_t st) {
    // State between launches (one-path-at-a-time callers): when the whole batch is one run per scenario, its first candidate is
    // linked to the last candidate of the previous launch on this handle, whose flow and potentials every scenario's row of
    // d_state holds, and its last candidate leaves its own there.  SGUFP_K1_STATE=0 turns it off.
    const int group = k1_group(K, c->S, c->sm_count);
    const char *state_env = getenv("SGUFP_K1_STATE");
    const bool state_on = !state_env || atoi(state_env) != 0;
    int state_io = 0;
    std::vector<int32_t> link0;
    if (state_on && group >= K && c->S > 0) {
        const int stride = 2 + c->M.nc + c->M.m;
        if (!c->d_state.p) {
            if (c->d_state.reserve((size_t)c->S * stride) == cudaSuccess) {
                c->state_stride = stride; c->state_valid = false;
                CU(c, cudaMemsetAsync(c->d_state.p, 0, (size_t)c->S * stride * 4, st));
            } else cudaGetLastError();               // no room for it: every launch starts from zero flow
        }
        if (c->d_state.p) {
            state_io = 2;
            if (c->state_valid) { link_plans(c->state_plan, B.plans[B.order[0]], link0); if (!link0.empty()) state_io |= 1; }
        }
    }
    const size_t pool_words = B.total_words + link0.size();
    CU(c, c->d_plans.reserve(pool_words));
    CU(c, c->d_plan_off.reserve(3 * (size_t)K));
    // gather the plans (+ offsets + the initial "no infeasible scenario" marks) into pinned memory: the uploads are then
    // true asynchronous copies.  The staging buffer is reused by the next call: an event tells when the copies left it.
    const size_t o_off = pool_words, o_inf = (o_off + 3 * (size_t)K + 1) & ~(size_t)1, need = o_inf + 2 * (size_t)K;
    if (c->h2d_pending) { CU(c, cudaEventSynchronize(c->ev_h2d)); c->h2d_pending = false; }
    if (c->h_words_cap < need) {
        if (c->h_words) cudaFreeHost(c->h_words);
        c->h_words = nullptr; c->h_words_cap = 0;
        CU(c, cudaHostAlloc(reinterpret_cast<void **>(&c->h_words), need * 2 * sizeof(int32_t), cudaHostAllocDefault));
        c->h_words_cap = need * 2;
    }
    for (int k = 0; k < K; k++) std::memcpy(c->h_words + B.off[k], B.plans[k].words.data(), B.plans[k].words.size() * 4);
    for (int k = 0; k < K; k++) if (B.link_off[k] >= 0) std::memcpy(c->h_words + B.link_off[k], B.links[k].data(), B.links[k].size() * 4);
    std::memcpy(c->h_words + o_off, B.off.data(), (size_t)K * 4);
    std::memcpy(c->h_words + o_off + K, B.link_off.data(), (size_t)K * 4);
    std::memcpy(c->h_words + o_off + 2 * (size_t)K, B.order.data(), (size_t)K * 4);
    if (state_io & 1) { std::memcpy(c->h_words + B.total_words, link0.data(), link0.size() * 4); c->h_words[o_off + K + B.order[0]] = (int32_t)B.total_words; }
    long long *inf = reinterpret_cast<long long *>(c->h_words + o_inf);
    for (int k = 0; k < K; k++) inf[k] = LLONG_MAX;
    CU(c, cudaMemcpyAsync(c->d_plans.p, c->h_words, pool_words * 4, cudaMemcpyHostToDevice, st));
    CU(c, cudaMemcpyAsync(c->d_plan_off.p, c->h_words + o_off, 3 * (size_t)K * 4, cudaMemcpyHostToDevice, st));
    CU(c, cudaMemcpyAsync(d_finf, inf, (size_t)K * 8, cudaMemcpyHostToDevice, st));
    CU(c, cudaEventRecord(c->ev_h2d, st));
    c->h2d_pending = true;
    CU(c, cudaMemsetAsync(d_sums, 0, (size_t)K * c->W() * 8, st));
    CU(c, c->d_work.reserve(1));
    CU(c, cudaMemsetAsync(c->d_work.p, 0, 8, st));
    K1Launch p{};
    p.cap_u = c->d_u; p.cap_l = c->d_l; p.S = c->S; p.m = c->M.m; p.m_pad = c->m_pad; p.scen_offset = c->scen_off;
    p.plans = c->d_plans.p; p.plan_off = c->d_plan_off.p; p.link_off = c->d_plan_off.p + K; p.order = c->d_plan_off.p + 2 * (size_t)K; p.group = group; p.K = K;
    p.state = state_io ? c->d_state.p : nullptr; p.state_stride = c->state_stride; p.state_io = state_io;
    p.xstride = std::max(1, B.max_nopen);
    CU(c, c->d_xout.reserve((size_t)K * std::max(1, c->S) * p.xstride));
    p.xout = c->d_xout.p; p.W = c->W(); p.L = c->M.L;
    p.sums = d_sums; p.first_inf = d_finf; p.obj = d_obj; p.status = d_status; p.work = c->d_work.p;
    p.max_nch = B.max_nch; p.max_nopen = B.max_nopen; p.nc = c->M.nc; p.nav = c->M.nav; p.max_cap = c->max_cap;
    p.has_lower = c->max_lower > 0; p.lane_tables = B.key_lane; p.sum_abs_r = c->sum_abs_r; p.max_indeg = B.max_indeg;
    c->kernel_timed = false;
    bool state_kept = false;
    if (c->S > 0) {
        CU(c, cudaEventRecord(c->evk0, st));
        {
            const cudaError_t ek = k1_launch(p, st, c->sm_count, &c->last_launches, &state_kept);
            if (ek == cudaErrorInvalidConfiguration) { cudaGetLastError(); return fail(c, SGUFP_ERR_LIMITS, "the per-scenario state of this network does not fit one SM's shared memory (227 KB) even with one warp per CTA"); }
            CU(c, ek);
        }
        CU(c, cudaEventRecord(c->evk1, st));
        c->kernel_timed = true;
    }
    // what the rows of d_state describe after this launch
    // (the lane kernel, SGUFP_K1_MODE=lane, does not touch them: a launch through it leaves no state to go on from)
    if ((state_io & 2) && state_kept) {
        const int last = B.order[K - 1];
        c->state_plan = B.plans[last]; c->state_valid = true;
        c->state_path.assign(B.key_paths.begin() + (size_t)last * B.key_L, B.key_paths.begin() + (size_t)(last + 1) * B.key_L);
    }
    else c->state_valid = false;
    return 0;
}

int run_ray(sgufp_ctx *c, const Plan &P, long long global_s, unsigned long long *d_sums, cudaStream_t st) {
    const long long sl = global_s - c->scen_off;
    if (sl < 0 || sl >= c->S) return fail(c, SGUFP_ERR_ARG, "ray: scenario is not in this rank's block");
    std::vector<int32_t> ts, hs, info, pl, nx, aq, fw;
    int nn = 0;
    ray_arrays(c->M, P, ts, hs, info, pl, nx, aq, fw, nn);
    const int m = c->M.m, nav = std::max(1, c->M.nav);
    std::vector<int32_t> pack;
    pack.reserve((size_t)6 * m + nav);
    for (auto *v : {&ts, &hs, &info, &pl, &nx, &aq}) pack.insert(pack.end(), v->begin(), v->end());
    pack.insert(pack.end(), fw.begin(), fw.end());
    CU(c, c->d_ray_i32.reserve(pack.size()));
    CU(c, c->d_ray_scratch.reserve((size_t)2 * nn + (size_t)(2 * m + 4 * nn) / 2 + 8));
    CU(c, cudaMemcpyAsync(c->d_ray_i32.p, pack.data(), pack.size() * 4, cudaMemcpyHostToDevice, st));
    CU(c, cudaMemsetAsync(d_sums, 0, (size_t)c->W() * 8, st));
    CU(c, cudaStreamSynchronize(st));
    RayLaunch r{};
    r.cap_u = c->d_u; r.cap_l = c->d_l; r.s_local = (int)sl; r.m = m; r.m_pad = c->m_pad; r.nn = nn;
    const int32_t *b = c->d_ray_i32.p;
    r.arc_ts = b; r.arc_hs = b + m; r.arc_info = b + 2 * m; r.arc_pair_layer = b + 3 * m; r.arc_next = b + 4 * m; r.arc_q = b + 5 * m;
    r.av_first_wire = b + 6 * m; r.nav = c->M.nav; r.L = c->M.L;
    r.scratch = reinterpret_cast<int32_t *>(c->d_ray_scratch.p);
    r.sums = d_sums;
    CU(c, ray_launch(r, st, &c->last_launches));
    return 0;
}

// exact integer sums -> Inavap::Cut (cutToCut, Cut.h:406-421)
void finalize_one(const sgufp_ctx *c, const Plan &P, const long long *sums, bool feas, int *cut_type, double *rhs,
                         uint64_t *keys, double *vals, int *nnz, double *coef_dense) {
    const Model &M = c->M;
    const int L = M.L, T = M.T;
    const double div = feas ? 1.0 : (double)c->S_total;   // `double scenarios` (grb.cpp:169); no 1/S on a ray (grb.cpp:301-347)
    const PlanHeader *H = reinterpret_cast<const PlanHeader *>(P.words.data());
    const int32_t *plan_info = P.words.data() + H->o_arc_info;
    auto kind = [&](int a) -> int {
        if (!feas) return plan_info[a] & 3;
        const int t = M.tail[a], h = M.head[a];
        if (M.active[h] && P.match_out[a] < 0) return KIND_SIGMA;
        if (M.active[t] && P.match_in[a] < 0) return KIND_PHI;
        return M.active[h] ? KIND_SIGMA : (M.active[t] ? KIND_PHI : KIND_GAMMA);
    };
    if (cut_type) *cut_type = feas ? SGUFP_CUT_FEASIBILITY : SGUFP_CUT_OPTIMALITY;
    if (rhs) *rhs = (double)sums[0] / div;
    static thread_local std::vector<double> scratch;
    if (!coef_dense && (int)scratch.size() < T) scratch.resize(T);
    double *coef = coef_dense ? coef_dense : scratch.data();      // slot order (the dense form the caller may ask for)
    const int32_t *mo = P.match_out.data();
    for (int ell = 0; ell < L; ell++) {
        const int a = M.layer_arc[ell], ma = mo[a];
        const long long sig = kind(a) == KIND_SIGMA ? sums[1 + L + a] : 0, lam = sums[1 + ell];
        for (int s = M.slot_base[ell]; s < M.slot_base[ell + 1]; s++) {
            const int b = M.slot_out[s];
            long long v = sig;
            if (ma == b) v -= lam;
            if (kind(b) == KIND_PHI) v += sums[1 + L + b];
            coef[s] = (double)v / div;
        }
    }
    int k = 0;
    if (keys && vals) {
        for (int r = 0; r < T; r++) {
            const double v = coef[M.slot_sorted[r]];
            if (v == 0) continue;     // `if (v == 0) continue;` (Cut.h:412)
            keys[k] = M.slot_key_sorted[r];   // getKey (Cut.h:342-344), (i,q,j)-lexicographic order
            vals[k] = v;
            k++;
        }
    } else
        for (int s = 0; s < T; s++) k += coef[s] != 0;
    if (nnz) *nnz = k;
}

std::mutex &cap_store_mutex() { static std::mutex m; return m; }

// stream, events, device properties of a handle (shared by sgufp_create, sgufp_create_from_cache and sgufp_clone)
int init_device(sgufp_ctx *c, int device, std::string &err) {
#define CUI(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e__); return SGUFP_ERR_CUDA; } } while (0)
    CUI(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUI(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) { err = std::string("device ") + prop.name + " is not sm_100-class; the kernels are built for sm_100a only"; return SGUFP_ERR_CUDA; }
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    CUI(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
    CUI(cudaEventCreate(&c->ev0));
    CUI(cudaEventCreate(&c->ev1));
    CUI(cudaEventCreate(&c->evk0));
    CUI(cudaEventCreate(&c->evk1));
    CUI(cudaEventCreateWithFlags(&c->ev_h2d, cudaEventDisableTiming));
#undef CUI
    return 0;
}

// rows are padded to an even number of arcs: the pad column of both arrays is zero
cudaError_t zero_pad_column(sgufp_ctx *c) {
    if (c->m_pad <= c->M.m || c->S == 0) return cudaSuccess;
    cudaError_t e = cudaMemset2DAsync(c->d_u + c->M.m, (size_t)c->m_pad * 8, 0, 8, (size_t)c->S, c->st);
    if (e == cudaSuccess) e = cudaMemset2DAsync(c->d_l + c->M.m, (size_t)c->m_pad * 8, 0, 8, (size_t)c->S, c->st);
    return e;
}

extern "C" {

const char *sgufp_last_error(const sgufp_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int sgufp_create(sgufp_ctx **out, int n, int m, int S, const int32_t *tail, const int32_t *head, const int32_t *upper,
                 const int32_t *lower, const int32_t *reward0, const int32_t *vbar, int nvbar, int device,
                 int64_t scenario_offset, int64_t S_total) {
    if (!out) { g_create_error = "out is null"; return SGUFP_ERR_ARG; }
    *out = nullptr;
    if (S < 0 || S_total < 1 || scenario_offset < 0 || scenario_offset + S > S_total || (S > 0 && device != SGUFP_DEVICE_NONE && (!upper || !lower))) {
        g_create_error = "scenario block must lie inside [0, S_total)"; return SGUFP_ERR_ARG;
    }
    sgufp_ctx *c = new sgufp_ctx();
    auto bail = [&](int code, const std::string &msg) { g_create_error = msg; sgufp_destroy(c); return code; };
    std::string e;
    if (int rc = c->M.build(n, m, tail, head, reward0, vbar, nvbar, e)) return bail(rc, e);
    for (size_t i = 0; upper && lower && i < (size_t)m * S; i++) {
        c->max_cap = std::max(c->max_cap, std::max(upper[i], lower[i]));
        c->max_lower = std::max(c->max_lower, lower[i]);
        if (upper[i] < 0 || lower[i] < 0 || upper[i] >= (1 << 20) || lower[i] >= (1 << 20))
            return bail(SGUFP_ERR_LIMITS, "capacities must lie in [0, 2^20) (DESIGN.md §5)");
    }
    { long long sa = 0; for (int a = 0; a < m; a++) sa += std::abs((long long)reward0[a]); c->sum_abs_r = (int)sa; }   // < 2^18 (Model::build)
    c->S = S; c->scen_off = scenario_offset; c->S_total = S_total; c->device = device;
    c->m_pad = (m + 1) & ~1;   // rows 16-byte aligned for 128-bit loads
#define CUC(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return bail(SGUFP_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); } while (0)
    if (device == SGUFP_DEVICE_NONE) { *out = c; return SGUFP_OK; }   // model only: no compute entry point will run
    { std::string de; if (int rc = init_device(c, device, de)) return bail(rc, de); }
    if (S > 0) {
        const size_t cells = (size_t)S * c->m_pad;
        c->caps = new CapStore();
        c->caps->device = device;
        CUC(cudaMalloc(&c->caps->d_u, cells * sizeof(double)));
        CUC(cudaMalloc(&c->caps->d_l, cells * sizeof(double)));
        c->d_u = c->caps->d_u; c->d_l = c->caps->d_l;
        // the reference's arc-major int32 vectors go up in slabs of arcs (at most 64 MiB of staging), each re-laid out on the device
        const int slab = (int)std::max<size_t>(32, std::min<size_t>((size_t)m, ((size_t)64 << 20) / ((size_t)S * 4)) / 32 * 32);
        int32_t *tmp = nullptr;
        CUC(cudaMalloc(&tmp, (size_t)std::min(slab, m) * S * 4));
        int launches = 0;
        cudaError_t e1 = cudaSuccess;
        for (int pass = 0; pass < 2 && e1 == cudaSuccess; pass++) {
            const int32_t *src = pass == 0 ? upper : lower;
            double *dst = pass == 0 ? c->d_u : c->d_l;
            for (int a0 = 0; a0 < m && e1 == cudaSuccess; a0 += slab) {
                const int na = std::min(slab, m - a0);
                e1 = cudaMemcpyAsync(tmp, src + (size_t)a0 * S, (size_t)na * S * 4, cudaMemcpyHostToDevice, c->st);
                if (e1 == cudaSuccess) e1 = relayout_launch(tmp, dst, na, S, c->m_pad, a0, c->st, &launches);
                if (e1 == cudaSuccess) e1 = cudaStreamSynchronize(c->st);      // the staging slab is reused
            }
        }
        if (e1 == cudaSuccess && c->m_pad > m) e1 = zero_pad_column(c);
        if (e1 == cudaSuccess) e1 = cudaStreamSynchronize(c->st);
        cudaFree(tmp);
        CUC(e1);
    }
#undef CUC
    *out = c;
    return SGUFP_OK;
}

void sgufp_destroy(sgufp_ctx *c) {
    if (!c) return;
    partition_destroy(c);          // communicators and the peer handles of a single-process partition (capi_shard.cu)
    if (c->dd_scratch && c->dd_scratch_free) c->dd_scratch_free(c->dd_scratch);
    delete c->pool;
    if (c->h_words) cudaFreeHost(c->h_words);
    if (c->h_out) cudaFreeHost(c->h_out);
    if (c->ev_h2d) cudaEventDestroy(c->ev_h2d);
    if (c->caps) {      // the capacity arrays go with their last handle
        bool last;
        { std::lock_guard<std::mutex> g(cap_store_mutex()); last = --c->caps->refs == 0; }
        if (last) { cudaSetDevice(c->caps->device); if (c->caps->d_u) cudaFree(c->caps->d_u); if (c->caps->d_l) cudaFree(c->caps->d_l); delete c->caps; }
        c->caps = nullptr; c->d_u = c->d_l = nullptr;
    }
    c->d_plans.release(); c->d_plan_off.release(); c->d_ray_i32.release(); c->d_sums.release(); c->d_work.release(); c->d_finf.release();
    c->d_ray_scratch.release(); c->d_obj.release(); c->d_status.release(); c->d_state.release(); c->d_xout.release();
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->evk0) cudaEventDestroy(c->evk0);
    if (c->evk1) cudaEventDestroy(c->evk1);
    if (c->st) cudaStreamDestroy(c->st);
    delete c;
}

int sgufp_dims(const sgufp_ctx *c, int *L, int *T, int *nvbar) {
    if (!c) return SGUFP_ERR_ARG;
    if (L) *L = c->M.L;
    if (T) *T = c->M.T;
    if (nvbar) *nvbar = (int)c->M.vbar.size();
    return 0;
}
int sgufp_network(const sgufp_ctx *c, int *n, int *m, int *S_local, int64_t *scenario_offset, int64_t *S_total, int32_t *tail, int32_t *head) {
    if (!c) return SGUFP_ERR_ARG;
    if (n) *n = c->M.n;
    if (m) *m = c->M.m;
    if (S_local) *S_local = c->S_view >= 0 ? (int)c->S_view : c->S;
    if (scenario_offset) *scenario_offset = c->S_view >= 0 ? 0 : c->scen_off;
    if (S_total) *S_total = c->S_total;
    if (tail) std::memcpy(tail, c->M.tail.data(), (size_t)c->M.m * 4);
    if (head) std::memcpy(head, c->M.head.data(), (size_t)c->M.m * 4);
    return 0;
}
int sgufp_vbar_order(const sgufp_ctx *c, int32_t *vbar) {
    if (!c || !vbar) return SGUFP_ERR_ARG;
    std::memcpy(vbar, c->M.vbar.data(), c->M.vbar.size() * 4);
    return 0;
}
int sgufp_processing_order(const sgufp_ctx *c, int32_t *layer_arc) {
    if (!c || !layer_arc) return SGUFP_ERR_ARG;
    std::memcpy(layer_arc, c->M.layer_arc.data(), (size_t)c->M.L * 4);
    return 0;
}
int sgufp_slots(const sgufp_ctx *c, int32_t *si, int32_t *sq, int32_t *sj, int32_t *lex_rank) {
    if (!c) return SGUFP_ERR_ARG;
    for (int s = 0; s < c->M.T; s++) {
        if (si) si[s] = c->M.tail[c->M.slot_in[s]];
        if (sq) sq[s] = c->M.head[c->M.slot_in[s]];
        if (sj) sj[s] = c->M.head[c->M.slot_out[s]];
        if (lex_rank) lex_rank[s] = c->M.slot_lex_rank[s];
    }
    return 0;
}
int sgufp_partial_width(const sgufp_ctx *c) { return c ? c->W() : SGUFP_ERR_ARG; }

int sgufp_paths_partial(sgufp_ctx *c, const int16_t *paths, int K, int L, int64_t *sums_device, int64_t *first_inf_device,
                        double *obj_device, uint8_t *status_device, void *cuda_stream) {
    if (!c || !sums_device || !first_inf_device) return SGUFP_ERR_ARG;
    if (c->device == SGUFP_DEVICE_NONE) return fail(c, SGUFP_ERR_CUDA, "handle was created with SGUFP_DEVICE_NONE: it holds the model only, there is no CPU compute path");
    CU(c, cudaSetDevice(c->device));
    c->last_launches = 0;
    Batch &B = c->batch;
    if (int rc = make_batch(c, paths, K, L, B)) return rc;
    cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : c->st;
    return launch_batch(c, B, K, reinterpret_cast<unsigned long long *>(sums_device), reinterpret_cast<long long *>(first_inf_device),
                        obj_device, status_device, st);
}

int sgufp_ray_partial(sgufp_ctx *c, const int16_t *path, int L, int64_t global_scenario, int64_t *sums_device, void *cuda_stream) {
    if (!c || !sums_device) return SGUFP_ERR_ARG;
    if (c->device == SGUFP_DEVICE_NONE) return fail(c, SGUFP_ERR_CUDA, "handle was created with SGUFP_DEVICE_NONE: it holds the model only, there is no CPU compute path");
    CU(c, cudaSetDevice(c->device));
    Batch B;
    if (int rc = make_batch(c, path, 1, L, B)) return rc;
    cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : c->st;
    return run_ray(c, B.plans[0], global_scenario, reinterpret_cast<unsigned long long *>(sums_device), st);
}

int sgufp_finalize_paths(sgufp_ctx *c, const int16_t *paths, int K, int L, const int64_t *sums_host, const int64_t *first_inf_host,
                         int *cut_type, double *rhs, uint64_t *keys, double *vals, int *nnz, double *coef_dense) {
    if (!c || !sums_host || !first_inf_host) return SGUFP_ERR_ARG;
    for (int k = 0; k < K; k++)
        if (first_inf_host[k] < 0) return fail(c, SGUFP_ERR_LIMITS, "path " + std::to_string(k) + ": the iteration guard of the subproblem solver was hit on some rank (k1_cut.cu: fuel)");
    Batch &B = c->batch;             // after sgufp_paths_partial with the same paths: no plan is built again
    if (int rc = make_batch(c, paths, K, L, B, true)) return rc;
    const int T = c->M.T, W = c->W();
    const std::function<void(int, int)> fin = [&](int t, int nt) {
        for (int k = t; k < K; k += nt)
            finalize_one(c, B.plans[k], reinterpret_cast<const long long *>(sums_host) + (size_t)k * W, first_inf_host[k] != LLONG_MAX,
                         cut_type ? cut_type + k : nullptr, rhs ? rhs + k : nullptr, keys ? keys + (size_t)k * T : nullptr,
                         vals ? vals + (size_t)k * T : nullptr, nnz ? nnz + k : nullptr, coef_dense ? coef_dense + (size_t)k * T : nullptr);
    };
    const int nt = host_threads(K);
    if (nt > 1 && c->pool) c->pool->run(nt, fin); else fin(0, 1);   // the pool exists if the plans of this batch were built on it
    return 0;
}

int sgufp_solve_paths(sgufp_ctx *c, const int16_t *paths, int K, int L, int *cut_type, double *rhs, uint64_t *keys, double *vals,
                      int *nnz, double *coef_dense, double *obj, uint8_t *status, int64_t *first_infeasible) {
    if (!c) return SGUFP_ERR_ARG;
    if (c->part)     // a scenario partition (sgufp_create_sharded / sgufp_comm_init): partial sums, ONE exchange, cuts (capi_shard.cu)
        return solve_paths_partitioned(c, paths, K, L, cut_type, rhs, keys, vals, nnz, coef_dense, obj, status, first_infeasible);
    if (c->scen_off != 0 || c->S != c->S_total)
        return fail(c, SGUFP_ERR_ARG, "handle holds a scenario shard without a communicator: call sgufp_comm_init first, or use sgufp_paths_partial + your own all-reduce + sgufp_finalize_paths");
    if (c->device == SGUFP_DEVICE_NONE) return fail(c, SGUFP_ERR_CUDA, "handle was created with SGUFP_DEVICE_NONE: it holds the model only, there is no CPU compute path");
    CU(c, cudaSetDevice(c->device));
    c->last_launches = 0;
    Batch &B = c->batch;
    if (int rc = make_batch(c, paths, K, L, B)) return rc;
    const int W = c->W(), T = c->M.T;
    const size_t KS = (size_t)K * c->S;
    CU(c, c->d_sums.reserve((size_t)K * W));
    CU(c, c->d_finf.reserve(K));
    if (obj) CU(c, c->d_obj.reserve(KS));
    if (status) CU(c, c->d_status.reserve(KS));
    CU(c, cudaEventRecord(c->ev0, c->st));
    if (int rc = launch_batch(c, B, K, c->d_sums.p, c->d_finf.p, obj ? c->d_obj.p : nullptr, status ? c->d_status.p : nullptr, c->st)) return rc;
    const size_t n_out = (size_t)K * W + (size_t)K;
    if (c->h_out_cap < n_out) {
        if (c->h_out) cudaFreeHost(c->h_out);
        c->h_out = nullptr; c->h_out_cap = 0;
        CU(c, cudaHostAlloc(reinterpret_cast<void **>(&c->h_out), n_out * 2 * sizeof(long long), cudaHostAllocDefault));
        c->h_out_cap = n_out * 2;
    }
    long long *sums = c->h_out, *finf = c->h_out + (size_t)K * W;     // pinned: the copies run at link speed, no staging
    CU(c, cudaMemcpyAsync(sums, c->d_sums.p, (size_t)K * W * 8, cudaMemcpyDeviceToHost, c->st));
    CU(c, cudaMemcpyAsync(finf, c->d_finf.p, (size_t)K * 8, cudaMemcpyDeviceToHost, c->st));
    if (obj) CU(c, cudaMemcpyAsync(obj, c->d_obj.p, KS * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    if (status) CU(c, cudaMemcpyAsync(status, c->d_status.p, KS, cudaMemcpyDeviceToHost, c->st));
    CU(c, cudaStreamSynchronize(c->st));
    for (int k = 0; k < K; k++)
        if (finf[k] < 0) return fail(c, SGUFP_ERR_LIMITS, "path " + std::to_string(k) + ": the iteration guard of the subproblem solver was hit (k1_cut.cu: fuel)");
    for (int k = 0; k < K; k++) {
        if (finf[k] == LLONG_MAX) continue;
        // the lowest-index infeasible scenario alone defines the cut (grb.cpp:284-351)
        if (int rc = run_ray(c, B.plans[k], finf[k], c->d_sums.p + (size_t)k * W, c->st)) return rc;
        CU(c, cudaMemcpyAsync(sums + (size_t)k * W, c->d_sums.p + (size_t)k * W, (size_t)W * 8, cudaMemcpyDeviceToHost, c->st));
        CU(c, cudaStreamSynchronize(c->st));
    }
    CU(c, cudaEventRecord(c->ev1, c->st));
    CU(c, cudaEventSynchronize(c->ev1));
    CU(c, cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
    const std::function<void(int, int)> fin = [&](int t, int nt) {
        for (int k = t; k < K; k += nt) {
            const bool feas = finf[k] != LLONG_MAX;
            finalize_one(c, B.plans[k], sums + (size_t)k * W, feas, cut_type ? cut_type + k : nullptr, rhs ? rhs + k : nullptr,
                         keys ? keys + (size_t)k * T : nullptr, vals ? vals + (size_t)k * T : nullptr, nnz ? nnz + k : nullptr,
                         coef_dense ? coef_dense + (size_t)k * T : nullptr);
            if (first_infeasible) first_infeasible[k] = feas ? finf[k] : -1;
        }
    };
    const int nt = host_threads(K);
    if (nt > 1 && c->pool) c->pool->run(nt, fin); else fin(0, 1);
    return SGUFP_OK;
}

int sgufp_solve_path(sgufp_ctx *c, const int16_t *path, int L, int *cut_type, double *rhs, uint64_t *keys, double *vals, int *nnz,
                     double *coef_dense, double *obj, uint8_t *status, int64_t *first_infeasible) {
    return sgufp_solve_paths(c, path, 1, L, cut_type, rhs, keys, vals, nnz, coef_dense, obj, status, first_infeasible);
}

uint64_t sgufp_cut_hash(const uint64_t *keys, const double *vals, int nnz) { return cut_hash(keys, vals, nnz); }

int sgufp_run_length(const sgufp_ctx *c, int K) {
    if (!c || K < 1) return SGUFP_ERR_ARG;
    return k1_group(K, c->S, c->sm_count);
}

int sgufp_last_stats(const sgufp_ctx *c, int *kernel_launches, float *device_ms) {
    if (!c) return SGUFP_ERR_ARG;
    if (kernel_launches) *kernel_launches = c->last_launches;
    if (device_ms) *device_ms = c->last_ms;
    return 0;
}

int sgufp_last_kernel_ms(sgufp_ctx *c, float *kernel_ms) {
    if (!c || !kernel_ms) return SGUFP_ERR_ARG;
    if (!c->kernel_timed) return fail(c, SGUFP_ERR_ARG, "no K1 launch has been recorded on this handle");
    CU(c, cudaEventSynchronize(c->evk1));
    CU(c, cudaEventElapsedTime(kernel_ms, c->evk0, c->evk1));
    return 0;
}

}  // extern "C"
