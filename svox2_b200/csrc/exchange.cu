// Multi-GPU gradient exchange helpers (no reference counterpart).
//
// Ranks march disjoint ray slices against a replicated grid.  The touched rows of all ranks
// (union of the sparse masks) are PACKED into one dense [rows, 28] buffer -- 27 SH gradients and
// the density gradient of a row side by side -- so that a single NCCL all-reduce moves exactly
// the bytes that matter, and the masked RMSProp then reads its gradients straight from the packed
// buffer (no unpack pass, density and SH updated by one kernel).
//
// Both kernels use lane = one float4 of a packed row (7 lanes per row for SH degree 2), 4 rows
// per warp pass, grid-stride; they require the 16-byte-aligned padded row layout SparseGrid keeps.
#include "common.cuh"

namespace svox2b {
namespace {

constexpr int EX_THREADS = 128;
#ifndef EX_ZERO_IN_PACK
#define EX_ZERO_IN_PACK 0  // 1: the pack kernel zeroes what it gathers; 0: the packed optimiser does
#endif

struct PackArgs {
    float* __restrict__ grad_sh;
    int64_t gsh_stride;
    int C;  // SH channels; the density value sits at packed[:, C]
    float* __restrict__ grad_density;
    int64_t gd_stride;
    const int64_t* __restrict__ rows;
    int64_t n_sel;
    float* __restrict__ packed;
    int64_t pack_stride;
};

__global__ void __launch_bounds__(EX_THREADS)
pack_rows_kernel(PackArgs a) {
    const int lane = threadIdx.x & 31;
    const int qpr = (int)(a.pack_stride >> 2);
    const int rpp = WARP / qpr;
    const int grp = lane / qpr;
    const int ch0 = (lane - grp * qpr) * 4;
    if (grp >= rpp) return;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp * rpp + grp; i < a.n_sel; i += nwarps * rpp) {
        const int64_t row = __ldg(a.rows + i);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ch0 < a.C) {  // quads past the SH channels hold only the density slot / padding
            float4* src = reinterpret_cast<float4*>(a.grad_sh + row * a.gsh_stride + ch0);
            v = *src;
#if EX_ZERO_IN_PACK
            *src = make_float4(0.f, 0.f, 0.f, 0.f);
#endif
        }
        float* ve = reinterpret_cast<float*>(&v);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (ch0 + e == a.C) {
                float* d = a.grad_density + row * a.gd_stride;
                ve[e] = *d;
#if EX_ZERO_IN_PACK
                *d = 0.f;
#endif
            } else if (ch0 + e > a.C) {
                ve[e] = 0.f;
            }
        }
        *reinterpret_cast<float4*>(a.packed + i * a.pack_stride + ch0) = v;
    }
}

struct PackedOptArgs {
    float* __restrict__ sh_data;
    float* __restrict__ sh_rms;
    int64_t sh_stride;
    int C;
    float* __restrict__ density_data;
    float* __restrict__ density_rms;
    int64_t density_stride;
    const int64_t* __restrict__ rows;
    int64_t n_sel;
    const float* __restrict__ packed;
    int64_t pack_stride;
    const uint8_t* __restrict__ mask_sh;
    const uint8_t* __restrict__ mask_density;
    float* __restrict__ grad_sh;       // zeroed for every listed row (may be NULL)
    int64_t gsh_stride;
    float* __restrict__ grad_density;
    int64_t gd_stride;
    float beta, lr_sh, lr_density, eps, minval_sh, minval_density;
};

// rmsprop_once, optim_kernel.cu:15-25 (same arithmetic as optim.cu)
__device__ __forceinline__ void rmsprop_update(float& d, float& rms, float g, float beta, float lr,
                                               float eps, float minval) {
    const float g2 = __fmul_rn(g, g);
    rms = (rms == 0.f) ? g2 : __fmaf_rn(beta, __fsub_rn(rms, g2), g2);
    d = fmaxf(__fsub_rn(d, __fdiv_rn(__fmul_rn(lr, g), __fadd_rn(sqrtf(rms), eps))), minval);
}

__global__ void __launch_bounds__(EX_THREADS)
rmsprop_packed_kernel(PackedOptArgs a) {
    const int lane = threadIdx.x & 31;
    const int qpr = (int)(a.pack_stride >> 2);
    const int rpp = WARP / qpr;
    const int grp = lane / qpr;
    const int ch0 = (lane - grp * qpr) * 4;
    if (grp >= rpp) return;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp * rpp + grp; i < a.n_sel; i += nwarps * rpp) {
        const int64_t row = __ldg(a.rows + i);
        const bool on_sh = __ldg(a.mask_sh + row) != 0 && ch0 < a.C;
        const bool on_den = __ldg(a.mask_density + row) != 0 && ch0 <= a.C && a.C < ch0 + 4;
        if (a.grad_sh) {  // the local gradients of this row are consumed: zero them for the next step
            if (ch0 < a.C)
                *reinterpret_cast<float4*>(a.grad_sh + row * a.gsh_stride + ch0) = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ch0 <= a.C && a.C < ch0 + 4) a.grad_density[row * a.gd_stride] = 0.f;
        }
        if (!on_sh && !on_den) continue;
        const float4 g = __ldg(reinterpret_cast<const float4*>(a.packed + i * a.pack_stride + ch0));
        const float* ge = reinterpret_cast<const float*>(&g);
        if (on_sh) {
            float4* dp = reinterpret_cast<float4*>(a.sh_data + row * a.sh_stride + ch0);
            float4* rp = reinterpret_cast<float4*>(a.sh_rms + row * a.sh_stride + ch0);
            float4 d = *dp, r = *rp;
            float* de = reinterpret_cast<float*>(&d);
            float* re = reinterpret_cast<float*>(&r);
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (ch0 + e < a.C)
                    rmsprop_update(de[e], re[e], ge[e], a.beta, a.lr_sh, a.eps, a.minval_sh);
            *dp = d;
            *rp = r;
        }
        if (on_den) {
            float* dp = a.density_data + row * a.density_stride;
            float* rp = a.density_rms + row * a.density_stride;
            float d = *dp, r = *rp;
            rmsprop_update(d, r, ge[a.C - ch0], a.beta, a.lr_density, a.eps, a.minval_density);
            *dp = d;
            *rp = r;
        }
    }
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

unsigned ex_blocks(int64_t n_sel, int rpp) {
    const int64_t warps = (n_sel + rpp - 1) / rpp;
    const int64_t want = (warps * WARP + EX_THREADS - 1) / EX_THREADS;
    const int64_t cap = 148 * 32;
    return (unsigned)(want < cap ? want : cap);
}

}  // namespace
}  // namespace svox2b

using namespace svox2b;

extern "C" SVOX2_API int svox2_pack_grad_rows(float* grad_sh, int64_t grad_sh_stride,
                                              int32_t sh_channels, float* grad_density,
                                              int64_t grad_density_stride, const int64_t* rows,
                                              int64_t n_sel, float* packed, int64_t pack_stride,
                                              svox2_stream_t stream) {
    SVOX2_REQUIRE(n_sel >= 0, "n_sel must be >= 0");
    if (n_sel == 0) return SVOX2_OK;
    SVOX2_REQUIRE(grad_sh && grad_density && rows && packed, "null argument");
    SVOX2_REQUIRE(sh_channels >= 1 && pack_stride % 4 == 0 && pack_stride > sh_channels &&
                      pack_stride <= 128,
                  "pack_stride must be a multiple of 4 in (sh_channels, 128]");
    SVOX2_REQUIRE(grad_sh_stride % 4 == 0 && grad_sh_stride >= ((sh_channels + 3) / 4) * 4 &&
                      aligned16(grad_sh) && aligned16(packed),
                  "grad_sh/packed need 16-byte aligned rows padded to a multiple of 4 floats");
    SVOX2_REQUIRE(grad_density_stride >= 1, "bad density stride");
    PackArgs a{grad_sh, grad_sh_stride, sh_channels, grad_density, grad_density_stride,
               rows,    n_sel,          packed,      pack_stride};
    const int rpp = WARP / (int)(pack_stride / 4);
    pack_rows_kernel<<<ex_blocks(n_sel, rpp), EX_THREADS, 0, (cudaStream_t)stream>>>(a);
    return check_launch(__func__);
}

extern "C" SVOX2_API int svox2_rmsprop_step_packed(
    float* sh_data, float* sh_rms, int64_t sh_stride, int32_t sh_channels, float* density_data,
    float* density_rms, int64_t density_stride, const int64_t* rows, int64_t n_sel,
    const float* packed, int64_t pack_stride, const uint8_t* mask_sh, const uint8_t* mask_density,
    float* grad_sh, int64_t grad_sh_stride, float* grad_density, int64_t grad_density_stride,
    float beta, float lr_sh, float lr_density, float epsilon, float minval_sh, float minval_density,
    svox2_stream_t stream) {
    SVOX2_REQUIRE(n_sel >= 0, "n_sel must be >= 0");
    if (n_sel == 0) return SVOX2_OK;
    SVOX2_REQUIRE(sh_data && sh_rms && density_data && density_rms && rows && packed && mask_sh &&
                      mask_density,
                  "null argument");
    SVOX2_REQUIRE(sh_channels >= 1 && pack_stride % 4 == 0 && pack_stride > sh_channels &&
                      pack_stride <= 128,
                  "pack_stride must be a multiple of 4 in (sh_channels, 128]");
    SVOX2_REQUIRE(sh_stride % 4 == 0 && sh_stride >= ((sh_channels + 3) / 4) * 4 &&
                      aligned16(sh_data) && aligned16(sh_rms) && aligned16(packed),
                  "sh_data/sh_rms/packed need 16-byte aligned rows padded to a multiple of 4 floats");
    SVOX2_REQUIRE(density_stride >= 1, "bad density stride");
    SVOX2_REQUIRE(!grad_sh || (grad_density && aligned16(grad_sh) && grad_sh_stride % 4 == 0 &&
                               grad_sh_stride >= ((sh_channels + 3) / 4) * 4 &&
                               grad_density_stride >= 1),
                  "grad_sh/grad_density (optional, zeroed) need the same padded layout");
    PackedOptArgs a{sh_data,     sh_rms,  sh_stride, sh_channels, density_data,
                    density_rms, density_stride, rows, n_sel, packed, pack_stride, mask_sh,
                    mask_density, grad_sh, grad_sh_stride, grad_density, grad_density_stride, beta, lr_sh, lr_density, epsilon, minval_sh, minval_density};
    const int rpp = WARP / (int)(pack_stride / 4);
    rmsprop_packed_kernel<<<ex_blocks(n_sel, rpp), EX_THREADS, 0, (cudaStream_t)stream>>>(a);
    return check_launch(__func__);
}
