// k2_dd.cu — K2: layer-wise longest path of a decision diagram with cut-adjusted arc weights (sm_100a).
//
// The hot loop of RelaxedDDNew::applyOptimalityCut / applyFeasibilityCut
// (/root/reference/DD.cpp:3951-3973, 3860-3882) and of the RestrictedDDNew versions
// (DD.cpp:3454-3470, 3376-3399), batched over B diagrams x C cuts:
//   state(root) = RHS + sum of the coefficients of the fixed prefix decisions   (DD.cpp:3938-3949)
//   state(v)    = max over in-arcs in stored order of state(tail) + w(arc),  w = coef[slot] or 0 for decision -1
// Only fp64 add / compare: results are bit-identical to the reference's (max is exact).
// Mapping: one CTA per (cut, diagram); a layer is a barrier-separated step; nodes of a layer are
// spread over the threads, a single-node (collapsed) layer is reduced by the whole CTA.  The CSR
// arrays stream once per CTA with coalesced loads (arcs of a layer are contiguous); the
// coefficient vector of the cut is gathered from L1/L2.  HBM-bound by design: 8 B (tail + slot)
// per arc + 8 B per parent state read + 8 B per node state written.
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
        double x = threadIdx.x < (K2_THREADS >> 5) ? red[threadIdx.x] : -DBL_MAX;
        for (int o = 16; o; o >>= 1) { const double t = __shfl_xor_sync(0xffffffffu, x, o); x = x < t ? t : x; }
        if (threadIdx.x == 0) red[0] = x;
    }
    __syncthreads();
    const double r = red[0];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(K2_THREADS) k2_longest_path(const K2DD *__restrict__ dds, const double *__restrict__ coef,
                                                               const double *__restrict__ rhs, int C, int Tpad, double *__restrict__ states) {
    __shared__ double red[K2_THREADS / 32];
    const K2DD d = dds[blockIdx.y];
    const int c = blockIdx.x;
    const double *__restrict__ cf = coef + (size_t)c * Tpad;
    double *__restrict__ st = states + d.state_off + (size_t)c * d.nnodes;
    if (threadIdx.x == 0) {
        double v = rhs[c];
        for (int k = 0; k < d.nroot; k++) { const int s = d.root_slot[k]; if (s >= 0) v = v + cf[s]; }
        st[0] = v;
    }
    __syncthreads();
    for (int l = 1; l < d.nlayers; l++) {
        const int v0 = d.layer_ptr[l], v1 = d.layer_ptr[l + 1];
        if (v1 - v0 == 1 && d.in_ptr[v0 + 1] - d.in_ptr[v0] > 64) {
            // collapsed layer: one node, many in-arcs -> the whole CTA reduces it
            double s = -DBL_MAX;
            for (int e = d.in_ptr[v0] + threadIdx.x; e < d.in_ptr[v0 + 1]; e += K2_THREADS) {
                const double p = st[d.arc_tail[e]];
                const int sl = d.arc_slot[e];
                const double cand = sl >= 0 ? p + cf[sl] : p;
                s = s < cand ? cand : s;
            }
            s = block_max(s, red);
            if (threadIdx.x == 0) st[v0] = s;
        } else {
            for (int v = v0 + threadIdx.x; v < v1; v += K2_THREADS) {
                double s = -DBL_MAX;   // DOUBLE_MIN = numeric_limits<double>::lowest() (DD.h:453)
                for (int e = d.in_ptr[v]; e < d.in_ptr[v + 1]; e++) {
                    const double p = st[d.arc_tail[e]];
                    const int sl = d.arc_slot[e];
                    const double cand = sl >= 0 ? p + cf[sl] : p;
                    s = s < cand ? cand : s;
                }
                st[v] = s;
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(K2_THREADS) k2_terminal(const K2DD *__restrict__ dds, int C, const double *__restrict__ states,
                                                           double *__restrict__ bound) {
    __shared__ double red[K2_THREADS / 32];
    const K2DD d = dds[blockIdx.x];
    const int last0 = d.layer_ptr[d.nlayers - 1];
    double best = -DBL_MAX;
    for (int i = threadIdx.x; i < d.nlast; i += K2_THREADS) {
        double t = d.term[i];
        for (int c = 0; c < C; c++) {
            const double s = states[d.state_off + (size_t)c * d.nnodes + last0 + i];
            t = s < t ? s : t;                 // arc.weight = min(arc.weight, parent.state2) (DD.cpp:3981)
        }
        d.term[i] = t;
        best = best < t ? t : best;            // terminalState = max(terminalState, arc.weight)
    }
    best = block_max(best, red);
    if (threadIdx.x == 0) bound[blockIdx.x] = best;
}

}  // namespace

cudaError_t k2_launch(const K2DD *dds, int B, const double *coef, const double *rhs, int C, int Tpad, double *states, int max_width_all,
                      cudaStream_t st, int *launches) {
    (void)max_width_all;
    dim3 grid(C, B);
    k2_longest_path<<<grid, K2_THREADS, 0, st>>>(dds, coef, rhs, C, Tpad, states);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

cudaError_t k2_terminal_launch(const K2DD *dds, int B, int C, const double *states, double *bound, cudaStream_t st, int *launches) {
    k2_terminal<<<B, K2_THREADS, 0, st>>>(dds, C, states, bound);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

}  // namespace sgufp
