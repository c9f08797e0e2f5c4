// capi_shard.cu — the scenario partition behind the C ABI (include/sgufp_b200.h, "scenario partition across GPUs").
//
// What is sharded is the reference's scenario loop (`for scenario in 0..S-1`, /root/reference/grb.cpp:174): scenarios are
// independent given y-bar, the only coupling is the sum of grb.cpp:241-278 and the "lowest-index infeasible scenario
// defines the cut" rule of grb.cpp:284-351.  Rank g owns a contiguous block of scenarios; per batch of K candidates there
// is ONE exchange: an all-reduce (SUM, int64) over K*W exact partial sums plus K flag words ("this rank met an infeasible
// scenario"), in place in the buffer K1 accumulated into.  Only when a flag comes back non-zero do the cold steps run:
// all-reduce (MIN) of the first infeasible index, the ray on the owning rank, and an all-reduce of that one row (the other
// ranks contribute zeros: a broadcast without a root to agree on).  Integer sums: 1 GPU and N GPUs give bit-identical cuts.
//
// Two ways to form a partition:
//   * sgufp_create_sharded: ONE process drives N devices (ncclCommInitAll; one host thread enqueues on every device's stream,
//     the collectives go in one ncclGroup).  A C++ host that holds one GuroSolver per thread (grb.h:36, NodeExplorer.h:115)
//     gets N GPUs behind an unchanged solveSubProblem.  Blocks that share a device (tests on a one-GPU box) are reduced by a
//     small kernel instead of NCCL.
//   * sgufp_comm_init: one process per GPU (torchrun / MPI): the host passes the ncclUniqueId of rank 0 around.
// NCCL is loaded at run time (dlopen "libnccl.so.2"): a process that already carries an NCCL (PyTorch) shares it, a plain C++
// host picks up the system library, and a one-GPU user needs none.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <climits>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "capi_internal.hpp"
#include "ctx.hpp"
#include "k1_cut.cuh"

using namespace sgufp;

namespace {

struct NcclApi {
    void *lib = nullptr;
    std::string err;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok() const { return lib != nullptr; }
};

// loaded once per process, by whichever thread asks first (a function-local static: its initialisation is thread-safe)
NcclApi load_nccl() {
    NcclApi api;
    void *h = nullptr;
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
        h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
        if (h) break;
    }
    if (!h) { const char *de = dlerror(); api.err = std::string("NCCL not found (dlopen libnccl.so.2): ") + (de ? de : ""); return api; }
    bool all = true;
    auto sym = [&](const char *n) { void *p = dlsym(h, n); if (!p) { all = false; api.err = std::string("NCCL symbol missing: ") + n; } return p; };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    if (all) api.lib = h;
    return api;
}

NcclApi &nccl() {
    static NcclApi api = load_nccl();
    return api;
}

// flags[k] = 1 if this rank met an infeasible scenario for candidate k (or hit its iteration guard): summed with the partial sums
__global__ void k1_pack_flags(const long long *__restrict__ first_inf, unsigned long long *__restrict__ flags, int K) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < K) flags[k] = first_inf[k] != LLONG_MAX ? 1ull : 0ull;
}
// blocks that share one device: dst += src / dst = min(dst, src)
__global__ void k1_add_rows(unsigned long long *__restrict__ dst, const unsigned long long *__restrict__ src, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] += src[i];
}
__global__ void k1_min_rows(long long *__restrict__ dst, const long long *__restrict__ src, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = min(dst[i], src[i]);
}

}  // namespace

struct Partition {
    int world = 1, rank0 = 0;               // this process holds ranks [rank0, rank0 + local.size())
    std::vector<sgufp_ctx *> local;         // local[0] is the handle the caller holds; the others belong to the partition
    std::vector<ncclComm_t> comms;          // one per local rank; empty when the blocks share one device
    bool same_device = false;
    std::vector<cudaEvent_t> ev;            // same_device: "rank r's partial sums are complete"
    int exchanges = 0;                      // collectives (or reduction kernels) of the last call: 1 on the feasible path
};

#define NC(c, call)                                                                                               \
    do {                                                                                                          \
        ncclResult_t r__ = (call);                                                                                \
        if (r__ != ncclSuccess) return fail(c, SGUFP_ERR_CUDA, std::string(#call) + ": " + nccl().GetErrorString(r__)); \
    } while (0)

// in-place all-reduce of `count` int64 words at `offset` of every local rank's buffer (sums: d_sums, finf: d_finf)
static int exchange(sgufp_ctx *c, bool finf, size_t offset, size_t count, ncclRedOp_t op) {
    Partition *P = c->part;
    P->exchanges++;
    const size_t nl = P->local.size();
    if (!P->comms.empty()) {
        NC(c, nccl().GroupStart());
        for (size_t i = 0; i < nl; i++) {
            sgufp_ctx *r = P->local[i];
            void *buf = finf ? static_cast<void *>(r->d_finf.p + offset) : static_cast<void *>(r->d_sums.p + offset);
            NC(c, nccl().AllReduce(buf, buf, count, ncclInt64, op, P->comms[i], r->st));
        }
        NC(c, nccl().GroupEnd());
        return 0;
    }
    if (P->world == 1) return 0;
    // blocks on one device: every stream's work must be complete before the leader folds it in, and the peers get the result back
    sgufp_ctx *lead = P->local[0];
    for (size_t i = 1; i < nl; i++) {
        sgufp_ctx *r = P->local[i];
        CU(c, cudaEventRecord(P->ev[i], r->st));
        CU(c, cudaStreamWaitEvent(lead->st, P->ev[i], 0));
        if (finf) k1_min_rows<<<(unsigned)((count + 127) / 128), 128, 0, lead->st>>>(lead->d_finf.p + offset, r->d_finf.p + offset, (int)count);
        else k1_add_rows<<<(unsigned)std::min<size_t>(1024, (count + 255) / 256), 256, 0, lead->st>>>(lead->d_sums.p + offset, r->d_sums.p + offset, count);
        CU(c, cudaGetLastError());
        lead->last_launches++;
    }
    // the peers' next writes into their buffers wait for the fold (only the leader's copy is read back: one process owns all blocks)
    CU(c, cudaEventRecord(P->ev[0], lead->st));
    for (size_t i = 1; i < nl; i++) CU(c, cudaStreamWaitEvent(P->local[i]->st, P->ev[0], 0));
    return 0;
}

// partial sums of every local block + flags + the ONE exchange; asynchronous on the handles' streams
static int partial_and_exchange(sgufp_ctx *c, const Batch &B, int K, bool want_obj, bool want_status) {
    Partition *P = c->part;
    const int W = c->W();
    P->exchanges = 0;
    for (sgufp_ctx *r : P->local) {
        CU(c, cudaSetDevice(r->device));
        r->last_launches = 0;
        const size_t KS = (size_t)K * r->S;
        CU(c, r->d_sums.reserve((size_t)K * W + K));
        CU(c, r->d_finf.reserve(K));
        if (want_obj) CU(c, r->d_obj.reserve(KS));
        if (want_status) CU(c, r->d_status.reserve(KS));
        if (r == c) CU(c, cudaEventRecord(c->ev0, c->st));
        if (int rc = launch_batch(r, B, K, r->d_sums.p, r->d_finf.p, want_obj ? r->d_obj.p : nullptr, want_status ? r->d_status.p : nullptr, r->st)) {
            if (r != c) c->err = r->err;
            return rc;
        }
        k1_pack_flags<<<(K + 127) / 128, 128, 0, r->st>>>(r->d_finf.p, r->d_sums.p + (size_t)K * W, K);
        CU(c, cudaGetLastError());
        r->last_launches++;
    }
    if (int rc = exchange(c, false, 0, (size_t)K * W + K, ncclSum)) return rc;
    CU(c, cudaSetDevice(c->device));
    return 0;
}

int solve_paths_partitioned(sgufp_ctx *c, const int16_t *paths, int K, int L, int *cut_type, double *rhs, uint64_t *keys, double *vals,
                            int *nnz, double *coef_dense, double *obj, uint8_t *status, int64_t *first_infeasible) {
    Partition *P = c->part;
    if (c->device == SGUFP_DEVICE_NONE) return fail(c, SGUFP_ERR_CUDA, "handle was created with SGUFP_DEVICE_NONE: there is no CPU compute path");
    Batch &B = c->batch;
    if (int rc = make_batch(c, paths, K, L, B)) return rc;
    const int W = c->W(), T = c->M.T;
    if (int rc = partial_and_exchange(c, B, K, obj != nullptr, status != nullptr)) return rc;
    const size_t n_out = (size_t)K * W + 2 * (size_t)K;
    if (c->h_out_cap < n_out) {
        if (c->h_out) cudaFreeHost(c->h_out);
        c->h_out = nullptr; c->h_out_cap = 0;
        CU(c, cudaHostAlloc(reinterpret_cast<void **>(&c->h_out), n_out * 2 * sizeof(long long), cudaHostAllocDefault));
        c->h_out_cap = n_out * 2;
    }
    long long *sums = c->h_out, *flags = sums + (size_t)K * W, *finf = flags + K;
    CU(c, cudaMemcpyAsync(sums, c->d_sums.p, ((size_t)K * W + K) * 8, cudaMemcpyDeviceToHost, c->st));   // sums + flags: one copy, no stop in between
    // per-scenario outputs: a single-process partition fills the caller's [K][S_total] arrays block by block;
    // a multi-process rank returns its own block [K][S_local]
    const bool whole = P->local.size() > 1 || P->world == 1;
    const size_t pitch = whole ? (size_t)c->S_total : (size_t)c->S;
    for (sgufp_ctx *r : P->local) {
        if (r->S == 0) continue;
        const size_t off = whole ? (size_t)r->scen_off : 0;
        if (obj) CU(c, cudaMemcpy2DAsync(obj + off, pitch * 8, r->d_obj.p, (size_t)r->S * 8, (size_t)r->S * 8, K, cudaMemcpyDeviceToHost, r->st));
        if (status) CU(c, cudaMemcpy2DAsync(status + off, pitch, r->d_status.p, (size_t)r->S, (size_t)r->S, K, cudaMemcpyDeviceToHost, r->st));
    }
    for (sgufp_ctx *r : P->local) { CU(c, cudaSetDevice(r->device)); CU(c, cudaStreamSynchronize(r->st)); }
    CU(c, cudaSetDevice(c->device));
    bool any_inf = false;
    for (int k = 0; k < K; k++) { finf[k] = LLONG_MAX; any_inf |= flags[k] != 0; }
    if (any_inf) {   // cold path: which scenario, the ray on its owner, the row to everybody (grb.cpp:284-351)
        if (int rc = exchange(c, true, 0, (size_t)K, ncclMin)) return rc;
        CU(c, cudaMemcpyAsync(finf, c->d_finf.p, (size_t)K * 8, cudaMemcpyDeviceToHost, c->st));
        CU(c, cudaStreamSynchronize(c->st));
        for (int k = 0; k < K; k++)
            if (finf[k] < 0) return fail(c, SGUFP_ERR_LIMITS, "path " + std::to_string(k) + ": the iteration guard of the subproblem solver was hit on some rank (k1_cut.cu: fuel)");
        for (int k = 0; k < K; k++) {
            if (finf[k] == LLONG_MAX) continue;
            for (sgufp_ctx *r : P->local) {
                CU(c, cudaSetDevice(r->device));
                unsigned long long *row = r->d_sums.p + (size_t)k * W;
                if (finf[k] >= r->scen_off && finf[k] < r->scen_off + r->S) {
                    if (int rc = run_ray(r, B.plans[k], finf[k], row, r->st)) { if (r != c) c->err = r->err; return rc; }
                } else CU(c, cudaMemsetAsync(row, 0, (size_t)W * 8, r->st));
            }
            if (int rc = exchange(c, false, (size_t)k * W, (size_t)W, ncclSum)) return rc;
            CU(c, cudaSetDevice(c->device));
            CU(c, cudaMemcpyAsync(sums + (size_t)k * W, c->d_sums.p + (size_t)k * W, (size_t)W * 8, cudaMemcpyDeviceToHost, c->st));
        }
        for (sgufp_ctx *r : P->local) { CU(c, cudaSetDevice(r->device)); CU(c, cudaStreamSynchronize(r->st)); }
        CU(c, cudaSetDevice(c->device));
    }
    CU(c, cudaEventRecord(c->ev1, c->st));
    CU(c, cudaEventSynchronize(c->ev1));
    CU(c, cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
    for (size_t i = 1; i < P->local.size(); i++) c->last_launches += P->local[i]->last_launches;
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

void partition_destroy(sgufp_ctx *c) {
    Partition *P = c->part;
    if (!P) return;
    c->part = nullptr;
    for (ncclComm_t cm : P->comms) if (cm && nccl().ok()) nccl().CommDestroy(cm);
    for (cudaEvent_t e : P->ev) if (e) cudaEventDestroy(e);
    for (size_t i = 1; i < P->local.size(); i++) { P->local[i]->part = nullptr; sgufp_destroy(P->local[i]); }
    delete P;
}

extern "C" {

int sgufp_comm_unique_id(void *id128) {
    if (!id128) return SGUFP_ERR_ARG;
    if (!nccl().ok()) return SGUFP_ERR_CUDA;
    ncclUniqueId id;
    if (nccl().GetUniqueId(&id) != ncclSuccess) return SGUFP_ERR_CUDA;
    std::memcpy(id128, &id, sizeof(id));
    return SGUFP_OK;
}

int sgufp_comm_init(sgufp_ctx *c, const void *id128, int rank, int world) {
    if (!c || !id128 || world < 1 || rank < 0 || rank >= world) return SGUFP_ERR_ARG;
    if (c->part) return fail(c, SGUFP_ERR_ARG, "handle already belongs to a partition");
    if (c->device == SGUFP_DEVICE_NONE) return fail(c, SGUFP_ERR_CUDA, "handle was created with SGUFP_DEVICE_NONE");
    if (!nccl().ok()) return fail(c, SGUFP_ERR_CUDA, nccl().err);
    CU(c, cudaSetDevice(c->device));
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = nullptr;
    NC(c, nccl().CommInitRank(&comm, world, id, rank));
    Partition *P = new Partition();
    P->world = world; P->rank0 = rank; P->local = {c}; P->comms = {comm};
    c->part = P;
    return SGUFP_OK;
}

int sgufp_comm_info(const sgufp_ctx *c, int *world, int *local_ranks, int *uses_nccl, int *exchanges_last_call) {
    if (!c) return SGUFP_ERR_ARG;
    const Partition *P = c->part;
    if (world) *world = P ? P->world : 1;
    if (local_ranks) *local_ranks = P ? (int)P->local.size() : 1;
    if (uses_nccl) *uses_nccl = P && !P->comms.empty();
    if (exchanges_last_call) *exchanges_last_call = P ? P->exchanges : 0;
    return SGUFP_OK;
}

int sgufp_create_sharded(sgufp_ctx **out, int n, int m, int S, const int32_t *tail, const int32_t *head, const int32_t *upper,
                         const int32_t *lower, const int32_t *reward0, const int32_t *vbar, int nvbar, const int *devices, int device_count) {
    if (!out) return SGUFP_ERR_ARG;
    *out = nullptr;
    if (!devices || device_count < 1 || S < 0 || !upper || !lower) return SGUFP_ERR_ARG;
    bool all_same = true, distinct = true;
    for (int i = 0; i < device_count; i++) {
        all_same &= devices[i] == devices[0];
        for (int j = 0; j < i; j++) distinct &= devices[i] != devices[j];
    }
    if (device_count > 1 && !all_same && !distinct) return SGUFP_ERR_ARG;   // NCCL takes one rank per device
    Partition *P = new Partition();
    P->world = device_count; P->rank0 = 0;
    P->same_device = device_count > 1 && all_same;
    std::vector<int32_t> bu, bl;
    int rc = SGUFP_OK;
    for (int r = 0; r < device_count && rc == SGUFP_OK; r++) {
        const long long base = S / device_count, rem = S % device_count;
        const long long lo = r * base + std::min<long long>(r, rem), Sr = base + (r < rem ? 1 : 0);   // contiguous blocks: "lowest infeasible index" stays a MIN
        bu.resize((size_t)m * Sr + 1); bl.resize((size_t)m * Sr + 1);
        for (int a = 0; a < m; a++) {
            std::memcpy(bu.data() + (size_t)a * Sr, upper + (size_t)a * S + lo, (size_t)Sr * 4);
            std::memcpy(bl.data() + (size_t)a * Sr, lower + (size_t)a * S + lo, (size_t)Sr * 4);
        }
        sgufp_ctx *ctx = nullptr;
        rc = sgufp_create(&ctx, n, m, (int)Sr, tail, head, bu.data(), bl.data(), reward0, vbar, nvbar, devices[r], lo, std::max(1, S));
        if (rc == SGUFP_OK) P->local.push_back(ctx);
    }
    auto bail = [&](int code) {
        for (sgufp_ctx *x : P->local) sgufp_destroy(x);
        for (cudaEvent_t e : P->ev) if (e) cudaEventDestroy(e);
        delete P;
        return code;
    };
    if (rc != SGUFP_OK) return bail(rc);     // sgufp_last_error(NULL) holds the message of the failing sgufp_create
    sgufp_ctx *lead = P->local[0];
    if (device_count > 1 && !P->same_device) {
        if (!nccl().ok()) return bail(SGUFP_ERR_CUDA);
        P->comms.assign(device_count, nullptr);
        if (nccl().CommInitAll(P->comms.data(), device_count, devices) != ncclSuccess) return bail(SGUFP_ERR_CUDA);
    }
    if (P->same_device) {
        cudaSetDevice(devices[0]);
        P->ev.assign(device_count, nullptr);
        for (auto &e : P->ev) if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return bail(SGUFP_ERR_CUDA);
    }
    lead->part = P;       // the peers only serve their block: the partition is driven through the leader
    lead->S_view = S;
    *out = lead;
    return SGUFP_OK;
}

int sgufp_paths_reduced(sgufp_ctx *c, const int16_t *paths, int K, int L, int64_t **sums_device, int64_t **first_inf_device) {
    if (!c) return SGUFP_ERR_ARG;
    if (!c->part) return fail(c, SGUFP_ERR_ARG, "handle is not part of a scenario partition (sgufp_create_sharded / sgufp_comm_init)");
    Batch &B = c->batch;
    if (int rc = make_batch(c, paths, K, L, B)) return rc;
    if (int rc = partial_and_exchange(c, B, K, false, false)) return rc;
    if (sums_device) *sums_device = reinterpret_cast<int64_t *>(c->d_sums.p);
    if (first_inf_device) *first_inf_device = reinterpret_cast<int64_t *>(c->d_finf.p);
    return SGUFP_OK;
}

void *sgufp_stream(const sgufp_ctx *c) { return c ? static_cast<void *>(c->st) : nullptr; }

}  // extern "C"
