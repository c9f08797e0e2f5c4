/*
 * sgufp_b200.h — C ABI of the B200-native scenario-cut / DD-longest-path hot path.
 *
 * The reference has no FFI: its seam is the C++ class `GuroSolver` (/root/reference/grb.h:17-104)
 * and the cut-application members of `RelaxedDDNew` / `RestrictedDDNew`
 * (/root/reference/DD.h:713-728,797-808).  Every entry point below names the reference interface
 * it replaces.  Plain pointers and sizes only; all buffers are HOST memory owned by the caller
 * unless a parameter says "device".  Every function returns 0 on success or a negative
 * SGUFP_ERR_* code (the reference has no error convention: Release is -fno-exceptions,
 * CMakeLists.txt:27); `sgufp_last_error` returns a human-readable reason.
 *
 * There is no CPU fallback: every compute entry point fails with SGUFP_ERR_CUDA when no
 * sm_100-class device is usable.
 */
#ifndef SGUFP_B200_H
#define SGUFP_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SGUFP_OK 0
#define SGUFP_ERR_ARG (-1)      /* null pointer, size mismatch, arc id out of range                 */
#define SGUFP_ERR_MATCHING (-2) /* path assigns one out-arc to two in-arcs (never done by the DD)    */
#define SGUFP_ERR_CYCLIC (-3)   /* network is not a DAG                                              */
#define SGUFP_ERR_INSTANCE (-4) /* instance the reference itself mishandles (see sgufp_last_error)   */
#define SGUFP_ERR_LIMITS (-5)   /* exceeds a packing limit of the kernels         */
#define SGUFP_ERR_CUDA (-6)     /* CUDA runtime error / no device                                    */

#define SGUFP_DEVICE_NONE (-1)  /* sgufp_create: build the host-side model only (dims, orders, slots,
                                   sgufp_finalize_paths); every compute entry point then fails */

#define SGUFP_CUT_OPTIMALITY 0 /* CutType, /root/reference/Cut.h:22-25 */
#define SGUFP_CUT_FEASIBILITY 1

typedef struct sgufp_ctx sgufp_ctx;

/* ---- instance + solver handle ----------------------------------------------------------------
 * Replaces `Network::Network(file)` (Network.cpp:10-129) + `GuroSolver::GuroSolver` (grb.h:36-69).
 * Arrays are exactly what `NetworkArc` holds (Network.h:26-41): upper/lower are [m][S_local]
 * arc-major int32; reward0[a] = rewards[0] of arc a (the only column the production path reads,
 * grb.cpp:53,71,89); vbar is the V-bar list in FILE order (shuffleVBarNodes is applied inside).
 * The capacities are re-laid-out scenario-major fp64 in HBM.
 * Multi-GPU: each rank passes only its contiguous scenario block [scenario_offset,
 * scenario_offset + S_local) of S_total scenarios; S_total is the `scenarios` divisor of
 * grb.cpp:169.  Single GPU: scenario_offset = 0, S_total = S_local.
 * `device` is the CUDA ordinal.  One handle is re-entrant per thread like one GuroSolver
 * (NodeExplorer.h:115); use one handle per host thread. */
int sgufp_create(sgufp_ctx **out, int n, int m, int S_local, const int32_t *tail, const int32_t *head,
                 const int32_t *upper, const int32_t *lower, const int32_t *reward0, const int32_t *vbar,
                 int nvbar, int device, int64_t scenario_offset, int64_t S_total);
void sgufp_destroy(sgufp_ctx *ctx);
const char *sgufp_last_error(const sgufp_ctx *ctx); /* ctx may be NULL: error of the last failed create */

/* Derived model, identical to the reference's (Network.cpp:94-121, grb.h:60-68):
 * L = totalLayers, T = number of (i,q,j) coefficient keys, nvbar after the shuffle. */
int sgufp_dims(const sgufp_ctx *ctx, int *L, int *T, int *nvbar);
int sgufp_vbar_order(const sgufp_ctx *ctx, int32_t *vbar /*[nvbar]*/);
int sgufp_processing_order(const sgufp_ctx *ctx, int32_t *layer_arc /*[L]*/); /* processingOrder[l].second */
/* sizes and arcs of the network behind a handle (Network.h:69-117): what a handle made from a cache file knows */
int sgufp_network(const sgufp_ctx *ctx, int *n, int *m, int *S_local, int64_t *scenario_offset, int64_t *S_total,
                  int32_t *tail /*[m] or NULL*/, int32_t *head /*[m] or NULL*/);
/* slot s in processingOrder-major / outgoingArcs-minor order -> node ids (i,q,j) and its
 * rank in the (i,q,j)-lexicographic order of std::map (Cut.h:75) */
int sgufp_slots(const sgufp_ctx *ctx, int32_t *si, int32_t *sq, int32_t *sj, int32_t *lex_rank /*[T] each*/);

/* ---- K1: scenario-cut evaluation ---------------------------------------------------------------
 * Replaces `GuroSolver::solveSubProblem(const vector<int16_t>& path)` (grb.h:75, grb.cpp:139-360).
 * path[L] is the DD encoding (DD.h:424): id of the matched out-arc, or -1.
 * Outputs the `Inavap::Cut` the reference would build (Cut.h:406-421): RHS and nnz (key,value)
 * pairs, key = q | i<<16 | j<<32, (i,q,j)-lexicographic, exact zeros dropped; keys/vals need room
 * for T entries.  Optional (may be NULL): coef_dense[T] in slot order, obj[S_local] per-scenario
 * subproblem optimum, status[S_local] (0 optimal, 1 infeasible), first_infeasible (global
 * scenario index or -1).  A feasibility cut is built from the lowest-index infeasible scenario
 * alone, unscaled (grb.cpp:284-351). */
int sgufp_solve_path(sgufp_ctx *ctx, const int16_t *path, int L, int *cut_type, double *rhs, uint64_t *keys,
                     double *vals, int *nnz, double *coef_dense, double *obj, uint8_t *status,
                     int64_t *first_infeasible);

/* K candidate paths per call (paths[K][L]); outputs are arrays of K (keys/vals/coef_dense: K*T,
 * obj/status: K*S_local).  Same semantics as K calls of sgufp_solve_path, one kernel launch. */
int sgufp_solve_paths(sgufp_ctx *ctx, const int16_t *paths, int K, int L, int *cut_type, double *rhs,
                      uint64_t *keys, double *vals, int *nnz, double *coef_dense, double *obj,
                      uint8_t *status, int64_t *first_infeasible);

/* Scenario-sharded form.  Step 1, per rank: evaluate the local scenarios and leave
 * the UNSCALED exact integer partial sums in a caller-owned DEVICE buffer
 *   sums[k][0] = RHS numerator, sums[k][1 .. W) = internal accumulators, W = sgufp_partial_width(ctx),
 *   first_inf[k] = lowest local infeasible GLOBAL scenario index or INT64_MAX
 * on `cuda_stream` (a cudaStream_t; NULL = the stream the handle owns, so pass a real non-default stream to order the call with your own work).  Step 2, caller: all-reduce `sums` (SUM, int64)
 * and `first_inf` (MIN, int64) across ranks — integers, so every rank count gives bit-identical
 * cuts.  Step 3: sgufp_finalize_paths on the reduced HOST copies (given the paths of step 1 on the same
 * handle it takes the plans step 1 built instead of building them again).  If first_inf[k] is finite the
 * owning rank (the one whose block contains it) must call sgufp_ray_partial for that path and
 * broadcast its sums before step 3. */
int sgufp_partial_width(const sgufp_ctx *ctx);
int sgufp_paths_partial(sgufp_ctx *ctx, const int16_t *paths, int K, int L, int64_t *sums_device,
                        int64_t *first_inf_device, double *obj_device /*[K][S_local] or NULL*/,
                        uint8_t *status_device /*[K][S_local] or NULL*/, void *cuda_stream);
int sgufp_ray_partial(sgufp_ctx *ctx, const int16_t *path, int L, int64_t global_scenario,
                      int64_t *sums_device /*[W]*/, void *cuda_stream);
int sgufp_finalize_paths(sgufp_ctx *ctx, const int16_t *paths, int K, int L, const int64_t *sums_host,
                         const int64_t *first_inf_host, int *cut_type, double *rhs, uint64_t *keys,
                         double *vals, int *nnz, double *coef_dense);

/* ---- scenario partition across GPUs, exchange inside the library -------------------
 * What is sharded is the scenario loop of GuroSolver::solveSubProblem (`for scenario`, grb.cpp:174); the
 * coupling is the sum of grb.cpp:241-278 and the lowest-index-infeasible rule of grb.cpp:284-351.
 * On a handle that belongs to a partition sgufp_solve_path(s) runs: K1 on every block -> ONE all-reduce
 * (SUM, int64, K*W partial sums + K "met an infeasible scenario" flags, in place in the buffer K1 wrote)
 * -> one device-to-host copy -> cuts.  Only a non-zero flag triggers the cold steps (MIN of the first
 * infeasible index, the ray on the owning rank, all-reduce of that row).  Integer sums: the cuts are
 * bit-identical for every number of GPUs.
 *
 * (1) One process, N devices (a C++ host with one GuroSolver per thread, grb.h:36 / NodeExplorer.h:115):
 *     sgufp_create_sharded takes the whole instance as sgufp_create does, cuts the scenarios into N contiguous
 *     blocks, puts block r on devices[r] and opens the communicator itself (ncclCommInitAll).  The returned
 *     handle is used exactly like a one-GPU handle (obj/status: [K][S]).  Equal device ids (a one-GPU box) are
 *     allowed: the blocks are then reduced by a kernel, no NCCL involved.
 * (2) One process per GPU (torchrun, MPI): every rank creates its block with sgufp_create(..., scenario_offset,
 *     S_total); rank 0 calls sgufp_comm_unique_id and the HOST hands the 128 bytes to the others (its own
 *     launcher's channel); every rank calls sgufp_comm_init.  sgufp_solve_path(s) is then a collective call:
 *     same paths on every rank, same cuts back on every rank (obj/status: this rank's block, [K][S_local]).
 * NCCL is loaded at run time (dlopen "libnccl.so.2"); a one-GPU user needs none. */
#define SGUFP_COMM_ID_BYTES 128
int sgufp_create_sharded(sgufp_ctx **out, int n, int m, int S, const int32_t *tail, const int32_t *head,
                         const int32_t *upper, const int32_t *lower, const int32_t *reward0, const int32_t *vbar,
                         int nvbar, const int *devices, int device_count);
int sgufp_comm_unique_id(void *id128);
int sgufp_comm_init(sgufp_ctx *ctx, const void *id128, int rank, int world);
int sgufp_comm_info(const sgufp_ctx *ctx, int *world, int *local_ranks, int *uses_nccl, int *exchanges_last_call);
/* The device half of a partitioned call alone, asynchronous on the handle's stream (sgufp_stream): K1 on
 * every local block + the one all-reduce.  *sums_device: [K*W] reduced sums followed by [K] flag counts;
 * *first_inf_device: this rank's (not reduced) first infeasible indices.  For device-side timing. */
int sgufp_paths_reduced(sgufp_ctx *ctx, const int16_t *paths, int K, int L, int64_t **sums_device,
                        int64_t **first_inf_device);
/* the cudaStream_t every call on this handle is enqueued on */
void *sgufp_stream(const sgufp_ctx *ctx);

/* ---- on-disk cache either side of the path and handles that share the capacities --------
 * The reference parses `n m S`, then per arc `tail head (lb ub reward) x S` from text (Network::Network,
 * Network.cpp:18-51): O(m*S) tokens.  The cache file holds the scenario-major fp64 arrays exactly as they lie in
 * HBM (header + raw arrays, mmap-able); sgufp_create_from_cache streams the rows of one scenario block
 * [scenario_offset, scenario_offset + S_local) (S_local < 0: to the end) to the device through pinned staging
 * chunks.  The handle behaves like one made by sgufp_create(..., scenario_offset, S_total = S of the file).
 * Errors of these three calls: sgufp_cache_last_error(). */
int sgufp_cache_write(const char *path, int n, int m, int S, const int32_t *tail, const int32_t *head,
                      const int32_t *upper /*[m][S]*/, const int32_t *lower, const int32_t *reward0,
                      const int32_t *vbar, int nvbar);
int sgufp_cache_dims(const char *path, int *n, int *m, int *S, int *nvbar, int64_t *file_bytes);
int sgufp_create_from_cache(sgufp_ctx **out, const char *path, int device, int64_t scenario_offset, int64_t S_local);
const char *sgufp_cache_last_error(void);
/* Another handle on the SAME device-resident capacities (read-only, reference-counted): own stream, own plan and
 * result buffers.  The reference runs N_WORKERS + 1 host threads, each with its own GuroSolver
 * (NodeExplorer.h:115-116, DDSolver.cpp:583,675, main.cpp:20): one upload, one clone per thread. */
int sgufp_clone(const sgufp_ctx *src, sgufp_ctx **out);

/* `Inavap::Cut::Cut` hash (Cut.h:243-251) of a (key,value) list, so a host wrapper can rebuild the
 * reference object bit for bit. */
uint64_t sgufp_cut_hash(const uint64_t *keys, const double *vals, int nnz);

/* Counters of the last compute call on this handle: kernels launched, device milliseconds
 * (CUDA events on the handle's stream). */
int sgufp_last_stats(const sgufp_ctx *ctx, int *kernel_launches, float *device_ms);
/* Device time of the last K1 launch alone (CUDA events recorded on the launching stream right
 * around the kernel); blocks until that launch has finished. */
int sgufp_last_kernel_ms(sgufp_ctx *ctx, float *kernel_ms);
/* How sgufp_solve_paths takes a batch of K candidates on this handle's scenarios: in runs of this many
 * candidates per scenario, neighbours in a nearest-neighbour chain the library lays through the batch's paths when that is
 * cheap against the batch (K <= 128 and K * L <= 4 * S; SGUFP_K1_ORDER=0 / 1: never / whenever K <= 128), else in the order given.  Results always come back in the order given.  The first candidate of a run is solved from zero flow, each of the others from the optimal flow
 * and potentials of the one before it (the paths NodeExplorer::process emits one after the other, NodeExplorer.cpp:949-971,
 * differ in a few layers).  The cuts do not depend on it.  1 = no warm starts; SGUFP_K1_GROUP=n overrides the choice. */
int sgufp_run_length(const sgufp_ctx *ctx, int K);

#ifdef __cplusplus
}
#endif
#endif /* SGUFP_B200_H */
