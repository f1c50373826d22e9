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
    const uint8_t* __restrict__ mask_a;  // optional: packed[:, C+1] = mask_a[row] != 0
    const uint8_t* __restrict__ mask_b;  // optional: packed[:, C+2] = mask_b[row] != 0
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
                if (a.mask_a && ch0 + e == a.C + 1) ve[e] = __ldg(a.mask_a + row) ? 1.f : 0.f;
                if (a.mask_b && ch0 + e == a.C + 2) ve[e] = __ldg(a.mask_b + row) ? 1.f : 0.f;
            }
        }
        *reinterpret_cast<float4*>(a.packed + i * a.pack_stride + ch0) = v;
    }
}

// packed rows -> array rows: dst = src (assign) or dst += src (add); mask bytes are OR-ed with the
// packed flags.  Rows are unique, so no atomics.
struct UnpackArgs {
    const float* __restrict__ packed;
    int64_t pack_stride;
    const int64_t* __restrict__ rows;
    int64_t n_sel;
    float* __restrict__ dst_sh;
    int64_t sh_stride;
    int C;
    float* __restrict__ dst_density;
    int64_t d_stride;
    int add;
    uint8_t* __restrict__ mask_a;
    uint8_t* __restrict__ mask_b;
};

__global__ void __launch_bounds__(EX_THREADS)
unpack_rows_kernel(UnpackArgs a) {
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
        const float4 v = __ldg(reinterpret_cast<const float4*>(a.packed + i * a.pack_stride + ch0));
        const float* ve = reinterpret_cast<const float*>(&v);
        if (ch0 < a.C) {
            float4* dp = reinterpret_cast<float4*>(a.dst_sh + row * a.sh_stride + ch0);
            float4 d = *dp;
            float* de = reinterpret_cast<float*>(&d);
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (ch0 + e < a.C) de[e] = a.add ? __fadd_rn(de[e], ve[e]) : ve[e];
            *dp = d;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (ch0 + e == a.C) {
                float* dp = a.dst_density + row * a.d_stride;
                *dp = a.add ? __fadd_rn(*dp, ve[e]) : ve[e];
            }
            if (a.mask_a && ch0 + e == a.C + 1 && ve[e] != 0.f) a.mask_a[row] = 1;
            if (a.mask_b && ch0 + e == a.C + 2 && ve[e] != 0.f) a.mask_b[row] = 1;
        }
    }
}

// ---- slab mode: per-ray combination of the all-gathered partial results -----------------------
// lt_in[q] = sum of logT_all[j][q] over the slabs ray q crosses before slab `rank`
__global__ void slab_prefix_kernel(const float* __restrict__ logT_all, const float* __restrict__ dirs,
                                   int64_t Q, int G, int rank, float* __restrict__ lt_in) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    const bool fwd = __ldg(dirs + q * 3) >= 0.f;
    float s = 0.f;
    if (fwd) for (int j = 0; j < rank; ++j) s = __fadd_rn(s, __ldg(logT_all + (int64_t)j * Q + q));
    else for (int j = G - 1; j > rank; --j) s = __fadd_rn(s, __ldg(logT_all + (int64_t)j * Q + q));
    lt_in[q] = s;
}

// total colour, dL/d(rgb) (cuvol.cu:735-747) and the reference's accum at slab entry (cuvol.cu:248-264)
__global__ void slab_combine_kernel(const float* __restrict__ part_all /*[G,Q,4]*/,
                                    const float* __restrict__ logT_all /*[G,Q]*/,
                                    const float* __restrict__ dirs, const float* __restrict__ rgb_gt,
                                    int64_t Q, int G, int rank, float background_brightness,
                                    float norm_factor, float beta_loss, float* __restrict__ rgb,
                                    float* __restrict__ gout, float* __restrict__ accum_start,
                                    int logT_is_total) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    const bool fwd = __ldg(dirs + q * 3) >= 0.f;
    float c[3] = {0.f, 0.f, 0.f}, lt = 0.f;
    bool stopped = false;
    // slabs in the order the ray crosses them
    for (int i = 0; i < G; ++i) {
        const int j = fwd ? i : G - 1 - i;
        const float4 p = __ldg(reinterpret_cast<const float4*>(part_all) + (int64_t)j * Q + q);
        c[0] += p.x; c[1] += p.y; c[2] += p.z;
        if (!logT_is_total) {
            stopped = stopped || p.w != 0.f;
            lt = __fadd_rn(lt, __ldg(logT_all + (int64_t)j * Q + q));
        }
    }
    if (stopped) lt = -1e3f;
    // ship mode: the rank that traced the ray composited it in one piece (-1e3 if it stopped)
    if (logT_is_total) lt = __ldcg(logT_all + q);
    const float T = __expf(lt);
    float g[3], accum = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        c[k] += T * background_brightness;
        g[k] = (c[k] - __ldg(rgb_gt + q * 3 + k)) * norm_factor;
        rgb[q * 3 + k] = c[k];
        gout[q * 3 + k] = g[k];
    }
    accum = fmaf(c[0], g[0], fmaf(c[1], g[1], c[2] * g[2]));
    if (beta_loss > 0.f)
        accum += (float)((double)beta_loss * (1 - (double)T / (1 - (double)T + 1e-3)));
    // minus what the slabs before this rank's already consumed: sum_i w_i * total_i = g . rgb_j
    for (int i = 0; i < G; ++i) {
        const int j = fwd ? i : G - 1 - i;
        if (j == rank) break;
        const float4 p = __ldg(reinterpret_cast<const float4*>(part_all) + (int64_t)j * Q + q);
        accum -= fmaf(p.x, g[0], fmaf(p.y, g[1], p.z * g[2]));
    }
    accum_start[q] = accum;
}

// ---- peer stores over NVLink (symmetric memory): all-gather without a collective library ------
// Every rank owns a buffer of identical layout; the base pointers of all ranks' buffers are mapped
// into every process (torch symmetric memory).  One launch copies this rank's slice into the same
// slot of EVERY rank's buffer with plain 128-bit stores (NVLink P2P writes); a device-side barrier
// (the caller's) then makes the gathered array readable everywhere.  20 B per ray per rank: the
// cost is two kernel launches, not a collective.
struct PeerPtrs {
    float* base[SVOX2_MAX_PEERS];
    int n;
};

__global__ void __launch_bounds__(256)
peer_store_kernel(PeerPtrs peers, int64_t dst_off4, const float4* __restrict__ src, int64_t n4) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = __ldg(src + i);
        for (int p = 0; p < peers.n; ++p) reinterpret_cast<float4*>(peers.base[p])[dst_off4 + i] = v;
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
                                              const uint8_t* mask_a, const uint8_t* mask_b,
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
    SVOX2_REQUIRE(!(mask_a || mask_b) || pack_stride >= sh_channels + 3,
                  "pack_stride must leave two slots behind the density value for the mask flags");
    PackArgs a{grad_sh, grad_sh_stride, sh_channels, grad_density, grad_density_stride,
               rows,    n_sel,          packed,      pack_stride, mask_a, mask_b};
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

extern "C" SVOX2_API int svox2_unpack_rows(const float* packed, int64_t pack_stride, const int64_t* rows,
                                           int64_t n_sel, float* dst_sh, int64_t dst_sh_stride,
                                           int32_t sh_channels, float* dst_density,
                                           int64_t dst_density_stride, int32_t add, uint8_t* mask_a,
                                           uint8_t* mask_b, svox2_stream_t stream) {
    SVOX2_REQUIRE(n_sel >= 0, "n_sel must be >= 0");
    if (n_sel == 0) return SVOX2_OK;
    SVOX2_REQUIRE(packed && rows && dst_sh && dst_density, "null argument");
    SVOX2_REQUIRE(sh_channels >= 1 && pack_stride % 4 == 0 && pack_stride > sh_channels &&
                      pack_stride <= 128,
                  "pack_stride must be a multiple of 4 in (sh_channels, 128]");
    SVOX2_REQUIRE(!(mask_a || mask_b) || pack_stride >= sh_channels + 3,
                  "pack_stride must leave two slots behind the density value for the mask flags");
    SVOX2_REQUIRE(dst_sh_stride % 4 == 0 && dst_sh_stride >= ((sh_channels + 3) / 4) * 4 &&
                      aligned16(dst_sh) && aligned16(packed),
                  "dst_sh/packed need 16-byte aligned rows padded to a multiple of 4 floats");
    SVOX2_REQUIRE(dst_density_stride >= 1, "bad density stride");
    UnpackArgs a{packed, pack_stride, rows, n_sel, dst_sh, dst_sh_stride, sh_channels, dst_density,
                 dst_density_stride, add, mask_a, mask_b};
    const int rpp = WARP / (int)(pack_stride / 4);
    unpack_rows_kernel<<<ex_blocks(n_sel, rpp), EX_THREADS, 0, (cudaStream_t)stream>>>(a);
    return check_launch(__func__);
}

extern "C" SVOX2_API int svox2_slab_prefix(const float* logT_all, const float* dirs, int64_t n_rays,
                                           int32_t n_slabs, int32_t rank, float* lt_in,
                                           svox2_stream_t stream) {
    SVOX2_REQUIRE(n_rays >= 0 && n_slabs >= 1 && rank >= 0 && rank < n_slabs, "bad slab arguments");
    if (n_rays == 0) return SVOX2_OK;
    SVOX2_REQUIRE(logT_all && dirs && lt_in, "null argument");
    slab_prefix_kernel<<<div_up(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(logT_all, dirs, n_rays,
                                                                               n_slabs, rank, lt_in);
    return check_launch(__func__);
}

extern "C" SVOX2_API int svox2_slab_combine(const float* part_all, const float* logT_all,
                                            const float* dirs, const float* rgb_gt, int64_t n_rays,
                                            int32_t n_slabs, int32_t rank,
                                            float background_brightness, float beta_loss,
                                            float* rgb, float* gout, float* accum_start,
                                            svox2_stream_t stream) {
    SVOX2_REQUIRE(n_rays >= 0 && n_slabs >= 1 && rank >= 0 && rank < n_slabs, "bad slab arguments");
    if (n_rays == 0) return SVOX2_OK;
    SVOX2_REQUIRE(part_all && logT_all && dirs && rgb_gt && rgb && gout && accum_start, "null argument");
    SVOX2_REQUIRE(aligned16(part_all), "part_all must be 16-byte aligned");
    slab_combine_kernel<<<div_up(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(
        part_all, logT_all, dirs, rgb_gt, n_rays, n_slabs, rank, background_brightness,
        2.f / (float)(3 * (int)n_rays), beta_loss / (float)n_rays, rgb, gout, accum_start, 0);
    return check_launch(__func__);
}

extern "C" SVOX2_API int svox2_ship_combine(const float* part_all, const float* ray_logT,
                                            const float* dirs, const float* rgb_gt, int64_t n_rays,
                                            int32_t n_slabs, int32_t rank,
                                            float background_brightness, float beta_loss,
                                            float* rgb, float* gout, float* accum_start,
                                            svox2_stream_t stream) {
    SVOX2_REQUIRE(n_rays >= 0 && n_slabs >= 1 && rank >= 0 && rank < n_slabs, "bad slab arguments");
    if (n_rays == 0) return SVOX2_OK;
    SVOX2_REQUIRE(part_all && ray_logT && dirs && rgb_gt && rgb && gout && accum_start, "null argument");
    SVOX2_REQUIRE(aligned16(part_all), "part_all must be 16-byte aligned");
    slab_combine_kernel<<<div_up(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(
        part_all, ray_logT, dirs, rgb_gt, n_rays, n_slabs, rank, background_brightness,
        2.f / (float)(3 * (int)n_rays), beta_loss / (float)n_rays, rgb, gout, accum_start, 1);
    return check_launch(__func__);
}

extern "C" SVOX2_API int svox2_peer_store(const uint64_t* peer_base_ptrs, int32_t n_peers,
                                          int64_t dst_offset_bytes, const void* src, int64_t n_bytes,
                                          svox2_stream_t stream) {
    SVOX2_REQUIRE(n_bytes >= 0 && n_peers >= 0 && n_peers <= SVOX2_MAX_PEERS, "bad peer count / size");
    if (n_bytes == 0 || n_peers == 0) return SVOX2_OK;
    SVOX2_REQUIRE(peer_base_ptrs && src, "null argument");
    SVOX2_REQUIRE(n_bytes % 16 == 0 && dst_offset_bytes % 16 == 0 && dst_offset_bytes >= 0 &&
                      aligned16(src),
                  "sizes, offsets and src must be 16-byte aligned");
    PeerPtrs peers;
    peers.n = n_peers;
    for (int p = 0; p < n_peers; ++p) {
        SVOX2_REQUIRE(peer_base_ptrs[p] != 0 && peer_base_ptrs[p] % 16 == 0, "bad peer pointer");
        peers.base[p] = reinterpret_cast<float*>(peer_base_ptrs[p]);
    }
    const int64_t n4 = n_bytes / 16;
    const int64_t want = (n4 + 255) / 256;
    peer_store_kernel<<<(unsigned)(want < 148 * 4 ? want : 148 * 4), 256, 0, (cudaStream_t)stream>>>(
        peers, dst_offset_bytes / 16, static_cast<const float4*>(src), n4);
    return check_launch(__func__);
}
