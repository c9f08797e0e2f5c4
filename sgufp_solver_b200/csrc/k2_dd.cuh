// k2_dd.cuh — launch interface of the DD longest-path kernels (implementation in k2_dd.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace sgufp {

struct K2DD {   // device image of one decision diagram (CSR by layer)
    const int4 *layer_info;     // [nlayers]  {first node, first in-arc, width, 1 if every node has exactly one in-arc}
    const int32_t *in_ptr;      // [nnodes+1] in-arc offsets (only read for layers that are not uniform)
    const int2 *arc_ts;         // [narcs]    {position of the tail in the previous layer, coefficient slot or -1}
    const int32_t *root_slot;   // [nroot]    slots of the fixed prefix decisions
    double *term;               // [nlast]    terminal arc weights: min over every cut applied so far
    long long state_off;        // offset of this DD's full-state block (states of the LAST cut of a batch), in doubles
    long long last_off;         // offset of this DD's [C][nlast] block of last-layer states, in doubles
    int nlayers, nroot, nnodes, nlast, max_width;
};

// One CTA per (cut c, diagram b).  Node states ping-pong between two shared-memory buffers (a layer
// only reads the previous one); the cut's dense coefficient vector is staged in shared memory.
// Only the last layer goes to global memory (last[b][c][i]); the states of every node are written
// for cut C-1 alone, which is what the host semantics read back.
cudaError_t k2_launch(const K2DD *dds_device, int B, const double *coef_device /*[C][Tpad]*/, const double *rhs_device /*[C]*/,
                      int C, int Tpad, double *states_device, double *last_device, int max_width_all, cudaStream_t st, int *launches);
// true if the widest layer fits the shared-memory state buffers (else: global state blocks, one per cut)
bool k2_states_in_smem(int Tpad, int max_width_all);
// term[i] = min(term[i], min_c last[b][c][i]); bound[b] = max_i term[i]
cudaError_t k2_terminal_launch(const K2DD *dds_device, int B, int C, const double *last_device, double *bound_device,
                               cudaStream_t st, int *launches);

}  // namespace sgufp
