// k2_dd.cuh — launch interface of the DD longest-path kernels (implementation in k2_dd.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace sgufp {

struct K2DD {   // device image of one decision diagram (CSR by layer)
    const int32_t *layer_ptr, *in_ptr, *arc_tail, *arc_slot, *root_slot;
    double *term;          // [nlast] terminal arc weights: min over every cut applied so far
    long long state_off;   // offset of this DD's state block in the batch state buffer (in doubles)
    int nlayers, nroot, nnodes, nlast, max_width;
};

// One CTA per (cut c, diagram b): layer-wise longest path with weights coef[c][slot].
// states: per diagram C consecutive blocks of nnodes doubles starting at state_off.
cudaError_t k2_launch(const K2DD *dds_device, int B, const double *coef_device /*[C][Tpad]*/, const double *rhs_device /*[C]*/,
                      int C, int Tpad, double *states_device, int max_width_all, cudaStream_t st, int *launches);
// term[i] = min(term[i], min_c state_c[last layer][i]); bound[b] = max_i term[i]
cudaError_t k2_terminal_launch(const K2DD *dds_device, int B, int C, const double *states_device, double *bound_device,
                               cudaStream_t st, int *launches);

}  // namespace sgufp
