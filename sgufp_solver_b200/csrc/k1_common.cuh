// k1_common.cuh — what the two K1 kernels (k1_cut.cu: a warp per scenario; k1_lane.cu: a lane per scenario) share:
// the device view of a candidate's plan (model.hpp) and a few constants.  Included inside namespace-less scope of a
// translation unit that has already included model.hpp.
#pragma once
#include <climits>
#include <cstdint>

#include "model.hpp"

namespace sgufp {
namespace {

constexpr int HB = 10;                 // position bits packed under a chain capacity (warp kernel)
constexpr int LAB_INF = 0x3fffffff;    // "unreached"; label + increment never overflows an int
constexpr int NEG_INF = INT_MIN / 4;
constexpr int WARPS = 8;               // warps per CTA (warp kernel)
#ifndef SGUFP_K1_MINBLOCKS
#define SGUFP_K1_MINBLOCKS 4            // resident CTAs per SM the register allocation aims at (64 registers per thread)
#endif

struct PlanView {
    const PlanHeader *h;
    const int2 *arc_av;            // per arc: the V-bar node it enters as a matched in-arc (-1: none), its matched out-arc
    const int32_t *arc_cp, *arc_info, *arc_pre, *ch_ends, *ch_r, *ch_ptr, *ch_arcs, *ch_q, *av_ptr, *av_arcs, *fb_ptr, *fb_ch;
    const int2 *ch_st;             // static half of an open chain (model.hpp)
    const int2 *slots;             // lane kernel: in-slots sorted by head
    const int32_t *in_pd;          // lane kernel: first slot | in-degree << 16 per head index
    __device__ explicit PlanView(const int32_t *base) {
        h = reinterpret_cast<const PlanHeader *>(base);
        arc_av = reinterpret_cast<const int2 *>(base + h->o_arc_av); arc_cp = base + h->o_arc_cp; arc_info = base + h->o_arc_info; arc_pre = base + h->o_arc_pre;
        ch_ends = base + h->o_ch_ends; ch_r = base + h->o_ch_r; ch_ptr = base + h->o_ch_ptr; ch_arcs = base + h->o_ch_arcs;
        ch_q = base + h->o_ch_q; av_ptr = base + h->o_av_ptr; av_arcs = base + h->o_av_arcs; fb_ptr = base + h->o_fb_ptr; fb_ch = base + h->o_fb_ch;
        ch_st = reinterpret_cast<const int2 *>(base + h->o_ch_st);
        slots = reinterpret_cast<const int2 *>(base + h->o_slots);
        in_pd = base + h->o_in_pd;
    }
};

// the static half of an open chain, unpacked (model.hpp)
struct ChainEnds {
    int sv, ev, hf, hb, r;
    __device__ __forceinline__ explicit ChainEnds(const int2 st) {
        sv = st.x & 1023; ev = (st.x >> 10) & 1023; hf = (st.x >> 20) & 1023; hb = st.y & 1023; r = st.y >> 10;
    }
};

}  // namespace
}  // namespace sgufp
