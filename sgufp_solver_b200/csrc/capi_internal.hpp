// capi_internal.hpp — helpers of capi.cu that the partitioned entry points (capi_shard.cu) share.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "ctx.hpp"

using Batch = PlanBatch;   // ctx.hpp

int host_threads(int K);
// reuse: the caller is the second half of one operation on these paths and may take the plans the first half built
int make_batch(sgufp_ctx *c, const int16_t *paths, int K, int L, Batch &B, bool reuse = false);
// uploads the plans of B to c's device and launches K1 on c's scenario block (asynchronous on st)
int launch_batch(sgufp_ctx *c, const Batch &B, int K, unsigned long long *d_sums, long long *d_finf, double *d_obj, uint8_t *d_status,
                 cudaStream_t st);
int run_ray(sgufp_ctx *c, const sgufp::Plan &P, long long global_s, unsigned long long *d_sums, cudaStream_t st);
void finalize_one(const sgufp_ctx *c, const sgufp::Plan &P, const long long *sums, bool feas, int *cut_type, double *rhs, uint64_t *keys,
                  double *vals, int *nnz, double *coef_dense);

#include <mutex>
#include <string>
std::mutex &cap_store_mutex();
int init_device(sgufp_ctx *c, int device, std::string &err);
cudaError_t zero_pad_column(sgufp_ctx *c);

// capi_shard.cu
int solve_paths_partitioned(sgufp_ctx *c, const int16_t *paths, int K, int L, int *cut_type, double *rhs, uint64_t *keys, double *vals,
                            int *nnz, double *coef_dense, double *obj, uint8_t *status, int64_t *first_infeasible);
void partition_destroy(sgufp_ctx *c);
