// k2_dd.cu — K2: layer-wise longest path of a decision diagram with cut-adjusted arc weights (sm_100a).
//
// The hot loop of RelaxedDDNew::applyOptimalityCut / applyFeasibilityCut
// (/root/reference/DD.cpp:3951-3973, 3860-3882) and of the RestrictedDDNew versions
// (DD.cpp:3454-3470, 3376-3399), batched over B diagrams x C cuts:
//   state(root) = RHS + sum of the coefficients of the fixed prefix decisions   (DD.cpp:3938-3949)
//   state(v)    = max over in-arcs in stored order of state(tail) + w(arc),  w = coef[slot] or 0 for decision -1
// Only fp64 add / compare: results are bit-identical to the reference's (max is exact).
//
// Mapping: one CTA per (cut, diagram); a layer is a barrier-separated step.  The node states never
// leave the SM: two shared-memory buffers ping-pong (a layer only reads the previous one) and the
// cut's dense coefficient vector is staged in shared memory once per CTA, so the only global
// traffic per arc is ONE coalesced 8-byte {tail position, slot} record (4 more for in_ptr where a
// layer is not a tree layer).  A single-node (collapsed) layer is reduced by the whole CTA.
// Diagrams wider than the shared-memory budget fall back to a global-state variant.
#include "k2_dd.cuh"

#include <cfloat>

namespace sgufp {
namespace {

constexpr int K2_THREADS = 256;

__device__ __forceinline__ double block_max(double v, double *red) {
    for (int o = 16; o; o >>= 1) { const double t = __shfl_xor_sync(0xffffffffu, v, o); v = v < t ? t : v; }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        double x = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : -DBL_MAX;
        for (int o = 16; o; o >>= 1) { const double t = __shfl_xor_sync(0xffffffffu, x, o); x = x < t ? t : x; }
        if (threadIdx.x == 0) red[0] = x;
    }
    __syncthreads();
    const double r = red[0];
    __syncthreads();
    return r;
}

// SMEM_STATES: states ping-pong in shared memory (width <= maxw); otherwise they live in the
// global scratch block of this (diagram, cut) — `gstate` then has C blocks per diagram.
template <bool SMEM_STATES>
__global__ void __launch_bounds__(K2_THREADS) k2_longest_path(const K2DD *__restrict__ dds, const double *__restrict__ coef,
                                                               const double *__restrict__ rhs, int C, int Tpad, int maxw,
                                                               double *__restrict__ gstate, double *__restrict__ glast) {
    extern __shared__ double sm[];
    __shared__ double red[K2_THREADS / 32];
    const K2DD d = dds[blockIdx.y];
    const int c = blockIdx.x;
    double *cf = sm;                       // [Tpad]
    double *buf0 = sm + Tpad, *buf1 = buf0 + (SMEM_STATES ? maxw : 0);
    for (int i = threadIdx.x; i < Tpad; i += blockDim.x) cf[i] = coef[(size_t)c * Tpad + i];
    const bool keep_all = c == C - 1;      // the host reads every node state of the last cut only
    double *all = gstate + d.state_off + (SMEM_STATES ? 0 : (size_t)c * d.nnodes);
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = rhs[c];
        for (int k = 0; k < d.nroot; k++) { const int s = d.root_slot[k]; if (s >= 0) v = v + cf[s]; }
        if (SMEM_STATES) buf0[0] = v;
        if (!SMEM_STATES || keep_all) all[0] = v;
        if (d.nlayers == 1) glast[d.last_off + (size_t)c * d.nlast] = v;
    }
    __syncthreads();
    for (int l = 1; l < d.nlayers; l++) {
        const int4 li = d.layer_info[l];
        const int v0 = li.x, e0 = li.y, width = li.z;
        const double *prev = SMEM_STATES ? ((l & 1) ? buf0 : buf1) : all + d.layer_info[l - 1].x;
        double *cur = SMEM_STATES ? ((l & 1) ? buf1 : buf0) : all + v0;
        const bool is_last = l == d.nlayers - 1;
        double *lastp = glast + d.last_off + (size_t)c * d.nlast;
        if (width == 1 && !li.w) {
            // collapsed layer: one node, many in-arcs -> the whole CTA reduces it
            const int e1 = d.in_ptr[v0 + 1];
            double s = -DBL_MAX;
            for (int e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
                const int2 ts = d.arc_ts[e];
                const double p = prev[ts.x];
                const double cand = ts.y >= 0 ? p + cf[ts.y] : p;
                s = s < cand ? cand : s;
            }
            s = block_max(s, red);
            if (threadIdx.x == 0) {
                cur[0] = s;
                if (SMEM_STATES && keep_all) all[v0] = s;
                if (is_last) lastp[0] = s;
            }
        } else if (li.w) {
            // tree layer: node i's only in-arc is arc e0 + i
            for (int i = threadIdx.x; i < width; i += blockDim.x) {
                const int2 ts = d.arc_ts[e0 + i];
                const double p = prev[ts.x];
                double s = ts.y >= 0 ? p + cf[ts.y] : p;
                s = -DBL_MAX < s ? s : -DBL_MAX;   // max(DOUBLE_MIN, .) of DD.cpp:3958
                cur[i] = s;
                if (SMEM_STATES && keep_all) all[v0 + i] = s;
                if (is_last) lastp[i] = s;
            }
        } else {
            for (int i = threadIdx.x; i < width; i += blockDim.x) {
                double s = -DBL_MAX;
                for (int e = d.in_ptr[v0 + i]; e < d.in_ptr[v0 + i + 1]; e++) {
                    const int2 ts = d.arc_ts[e];
                    const double p = prev[ts.x];
                    const double cand = ts.y >= 0 ? p + cf[ts.y] : p;
                    s = s < cand ? cand : s;
                }
                cur[i] = s;
                if (SMEM_STATES && keep_all) all[v0 + i] = s;
                if (is_last) lastp[i] = s;
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(K2_THREADS) k2_terminal(const K2DD *__restrict__ dds, int C, const double *__restrict__ glast,
                                                           double *__restrict__ bound) {
    __shared__ double red[K2_THREADS / 32];
    const K2DD d = dds[blockIdx.x];
    double best = -DBL_MAX;
    for (int i = threadIdx.x; i < d.nlast; i += blockDim.x) {
        double t = d.term[i];
        for (int c = 0; c < C; c++) {
            const double s = glast[d.last_off + (size_t)c * d.nlast + i];
            t = s < t ? s : t;                 // arc.weight = min(arc.weight, parent.state2) (DD.cpp:3981)
        }
        d.term[i] = t;
        best = best < t ? t : best;            // terminalState = max(terminalState, arc.weight)
    }
    best = block_max(best, red);
    if (threadIdx.x == 0) bound[blockIdx.x] = best;
}

}  // namespace

cudaError_t k2_launch(const K2DD *dds, int B, const double *coef, const double *rhs, int C, int Tpad, double *states, double *last,
                      int max_width_all, cudaStream_t st, int *launches) {
    dim3 grid(C, B);
    if (launches) (*launches)++;
    // narrow diagrams: smaller CTAs, more of them per SM (a layer is one barrier-separated step)
    int threads = 64;
    while (threads < K2_THREADS && threads < max_width_all) threads *= 2;
    const size_t smem = ((size_t)Tpad + 2 * (size_t)max_width_all) * sizeof(double);
    if (smem <= 200 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k2_longest_path<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k2_longest_path<true><<<grid, threads, smem, st>>>(dds, coef, rhs, C, Tpad, max_width_all, states, last);
    } else {
        const size_t sm2 = (size_t)Tpad * sizeof(double);
        cudaError_t e = cudaFuncSetAttribute(k2_longest_path<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
        if (e != cudaSuccess) return e;
        k2_longest_path<false><<<grid, threads, sm2, st>>>(dds, coef, rhs, C, Tpad, 0, states, last);
    }
    return cudaGetLastError();
}

bool k2_states_in_smem(int Tpad, int max_width_all) { return ((size_t)Tpad + 2 * (size_t)max_width_all) * sizeof(double) <= 200 * 1024; }

cudaError_t k2_terminal_launch(const K2DD *dds, int B, int C, const double *last, double *bound, cudaStream_t st, int *launches) {
    k2_terminal<<<B, K2_THREADS, 0, st>>>(dds, C, last, bound);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

}  // namespace sgufp
