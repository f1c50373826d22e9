// "Quad" mapping shared by the sample-parallel stages: an 8-lane group owns one sample and lane i
// of the group owns SH channels 4i..4i+3 of every corner row (one 128-bit load / vector reduction
// per corner row when rows are 16-byte aligned; a scalar path of the same code otherwise).
#pragma once
#include "common.cuh"

namespace svox2b {
namespace quad {

struct LaneConst {
    float Yc[3][4];  // SH basis value of channel e if it belongs to colour c, else 0
    int qi;          // quad index (lane & 7)
    bool quad_on;    // this lane owns at least one real channel
};

__device__ __forceinline__ void lane_const_init(LaneConst& lc, int lane, int B, int D,
                                                const float* dirw) {
    lc.qi = lane & 7;
    lc.quad_on = lc.qi * 4 < D;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int j = lc.qi * 4 + e;
        const bool on = j < D;
        const int c = on ? j / B : -1;
        const float y = on ? sh_basis_k(B, j % B, dirw) : 0.f;
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) lc.Yc[cc][e] = (c == cc) ? y : 0.f;
    }
}

template <bool VEC>
__device__ __forceinline__ float4 load_quad(const float* __restrict__ p, int qi, int D) {
    if (VEC) return __ldg(reinterpret_cast<const float4*>(p));
    float4 v;
    const int j = qi * 4;
    v.x = (j + 0 < D) ? __ldg(p + 0) : 0.f;
    v.y = (j + 1 < D) ? __ldg(p + 1) : 0.f;
    v.z = (j + 2 < D) ? __ldg(p + 2) : 0.f;
    v.w = (j + 3 < D) ? __ldg(p + 3) : 0.f;
    return v;
}

template <bool VEC>
__device__ __forceinline__ void red_quad(float* __restrict__ p, int qi, int D, float a, float b,
                                         float c, float d) {
    if (VEC) {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b),
                     "f"(c), "f"(d)
                     : "memory");
    } else {
        const int j = qi * 4;
        if (j + 0 < D) atomicAdd(p + 0, a);
        if (j + 1 < D) atomicAdd(p + 1, b);
        if (j + 2 < D) atomicAdd(p + 2, c);
        if (j + 3 < D) atomicAdd(p + 3, d);
    }
}

// trilerp_backward_cuvol_one weights (render_util.cuh:94-121), corner q = 4*dx + 2*dy + dz
__device__ __forceinline__ void corner_weights(float px, float py, float pz, float* wq) {
    const float ax = 1.f - px, ay = 1.f - py, az = 1.f - pz;
    const float yz00 = ay * az, yz01 = ay * pz, yz10 = py * az, yz11 = py * pz;
    wq[0] = yz00 * ax; wq[1] = yz01 * ax; wq[2] = yz10 * ax; wq[3] = yz11 * ax;
    wq[4] = yz00 * px; wq[5] = yz01 * px; wq[6] = yz10 * px; wq[7] = yz11 * px;
}

// Colours (+0.5, unclamped) of the sample an 8-lane group holds: 8 unconditional 128-bit row
// loads in flight, weights applied as FMAs, masked-basis dot product, 3-step butterfly.
template <bool VEC>
__device__ __forceinline__ void gather_colour(const GridView& g, const LaneConst& lc,
                                              const int32_t* L, const float* wq, bool ld,
                                              float* col) {
    const int D = g.sh_dim;
    float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* shq = g.sh + (lc.quad_on ? 4 * lc.qi : 0);
    float4 v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const unsigned row = (ld && L[q] >= 0) ? (unsigned)L[q] : 0u;
        v[q] = load_quad<VEC>(shq + (size_t)row * (unsigned)g.sh_stride, lc.qi, D);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        if (ld && L[q] >= 0) {
            u.x = fmaf(wq[q], v[q].x, u.x);
            u.y = fmaf(wq[q], v[q].y, u.y);
            u.z = fmaf(wq[q], v[q].z, u.z);
            u.w = fmaf(wq[q], v[q].w, u.w);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float x = lc.Yc[c][0] * u.x;
        x = fmaf(lc.Yc[c][1], u.y, x);
        x = fmaf(lc.Yc[c][2], u.z, x);
        x = fmaf(lc.Yc[c][3], u.w, x);
        x += __shfl_xor_sync(FULL, x, 1);
        x += __shfl_xor_sync(FULL, x, 2);
        x += __shfl_xor_sync(FULL, x, 4);
        col[c] = x + 0.5f;
    }
}

}  // namespace quad
}  // namespace svox2b
