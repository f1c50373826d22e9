// Ray-march kernels: forward, backward, fused seed, whole-image, MSI background, depth.
//
// Warp-per-ray march.  Lanes 0..7 own the eight corner links / sigma values of the current
// cell (one LDG each instead of eight redundant broadcasts per lane), lane j < 3*basis_dim
// owns SH channel j for the colour gather and the gradient scatter.  The traversal
// (t lattice, skipping, sigma, log-transmittance) is bit-exact with the reference:
//   forward   render_lerp_kernel_cuvol.cu:30-120, 611-690
//   backward  render_lerp_kernel_cuvol.cu:227-383, 692-773
//   MSI       render_lerp_kernel_cuvol.cu:386-607, 775-860
//   depth     render_lerp_kernel_cuvol.cu:122-225, 862-902
#include "common.cuh"
#include "march.cuh"

namespace svox2b {

namespace v1 {

constexpr int MARCH_THREADS = 128;
constexpr int BG_THREADS = 128;

// Sum over each group of B consecutive lanes, valid in the group's first lane.
__device__ __forceinline__ float seg_sum(float x, int k, int B) {
#pragma unroll
    for (int off = 1; off < 16; off <<= 1) {
        const float o = __shfl_down_sync(FULL, x, off);
        if (k + off < B) x += o;
    }
    return x;
}

__device__ __forceinline__ float corner_w(const float* pos, int c, float g) {
    // (ay*az) * ((1-px)*g), the association used by render_util.cuh:102-119
    const float wx = (c & 4) ? pos[0] : 1.f - pos[0];
    const float wy = (c & 2) ? pos[1] : 1.f - pos[1];
    const float wz = (c & 1) ? pos[2] : 1.f - pos[2];
    return (wy * wz) * (wx * g);
}

struct RaySource {
    const float* __restrict__ origins;
    const float* __restrict__ dirs;
    CamView cam;
    int row_begin;
};

template <bool FROM_CAM>
__device__ __forceinline__ void load_ray(const RaySource& src, int64_t ray_id, float* o, float* d) {
    if (FROM_CAM) {
        const int ix = (int)(ray_id % src.cam.width);
        const int iy = (int)(ray_id / src.cam.width) + src.row_begin;
        cam2world_ray<false>(ix, iy, src.cam, d, o);
    } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            o[i] = __ldg(src.origins + ray_id * 3 + i);
            d[i] = __ldg(src.dirs + ray_id * 3 + i);
        }
    }
}

// ---------------------------------------------------------------------------------------
// forward march
// ---------------------------------------------------------------------------------------
template <bool FROM_CAM, bool COUNT>
__global__ void __launch_bounds__(MARCH_THREADS)
march_fwd_kernel(GridView g, Opts opt, RaySource src, int64_t n_rays, float* __restrict__ out,
                 float* __restrict__ logT_out, unsigned long long* __restrict__ counters) {
    const int64_t ray_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (ray_id >= n_rays) return;
    const int lane = threadIdx.x & 31;
    const int B = g.basis_dim, D = g.sh_dim;
    const int k = lane % B, cgrp = lane / B;
    const bool chan_lane = lane < D;

    float ow[3], dw[3];
    load_ray<FROM_CAM>(src, ray_id, ow, dw);
    const float Yk = sh_basis_k(B, k, dw);  // raw input direction (cuvol.cu:632-635)
    Ray ray;
    ray_setup(ray, ow, dw, g, opt);

    const bool has_bg = g.bg_nlayers > 0;
    float outv = 0.f, logT = 0.f;
    unsigned n_skip = 0, n_vis = 0, n_act = 0;

    if (ray.tmin <= ray.tmax) {
        float t = ray.tmin;
        while (t <= ray.tmax) {
            int l[3];
            float pos[3];
            locate(ray, g, t, l, pos);
            const int64_t base = (int64_t)l[0] * g.offx + (int64_t)l[1] * g.offy + l[2];
            int32_t mylink = -1;
            if (lane < 8) mylink = __ldg(g.links + base + corner_off(g, lane));
            const int32_t link0 = __shfl_sync(FULL, mylink, 0);
            const float skip = skip_dist(ray, link0, l, pos);
            if (skip >= opt.step_size) {
                t = skip_advance(t, skip, opt.step_size);
                if (COUNT) ++n_skip;
                continue;
            }
            float myv = 0.f;
            if (lane < 8 && mylink >= 0) myv = __ldg(g.density + (int64_t)mylink * g.density_stride);
            const float sigma = trilerp_lanes(myv, pos);
            if (COUNT) ++n_vis;
            if (opt.last_sample_opaque && __fadd_rn(t, opt.step_size) > ray.tmax)
                ray.world_step = 1e9f;
            if (sigma > opt.sigma_thresh) {
                if (COUNT) ++n_act;
                float v[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int32_t lk = __shfl_sync(FULL, mylink, c);
                    v[c] = (chan_lane && lk >= 0) ? __ldg(g.sh + (int64_t)lk * g.sh_stride + lane) : 0.f;
                }
                const float total = seg_sum(trilerp8(v, pos) * Yk, k, B);
                const float pcnt = __fmul_rn(ray.world_step, sigma);
                const float weight = __fmul_rn(fexp(logT), __fsub_rn(1.f, fexp(-pcnt)));
                logT = __fsub_rn(logT, pcnt);
                outv = fmaf(weight, fmaxf(total + 0.5f, 0.f), outv);
                if (fexp(logT) < opt.stop_thresh) {
                    logT = -1e3f;
                    break;
                }
            }
            t = __fadd_rn(t, opt.step_size);
        }
        if (!has_bg) outv = fmaf(fexp(logT), opt.background_brightness, outv);
    } else {
        outv = has_bg ? 0.f : opt.background_brightness;
        logT = 0.f;
    }
    if (chan_lane && k == 0) out[ray_id * 3 + cgrp] = outv;
    if (lane == 0) {
        if (logT_out) logT_out[ray_id] = logT;
        if (COUNT) {
            atomicAdd(counters + 0, (unsigned long long)n_skip);
            atomicAdd(counters + 1, (unsigned long long)n_vis);
            atomicAdd(counters + 2, (unsigned long long)n_act);
        }
    }
}

// ---------------------------------------------------------------------------------------
// backward march (re-materialises the samples)
// ---------------------------------------------------------------------------------------
struct BwdArgs {
    const float* __restrict__ grad_in;      // grad_out [Q,3], or rgb_gt when fused
    const float* __restrict__ color_cache;  // [Q,3]
    const float* __restrict__ logT_in;      // forward log-transmittance (beta loss) or NULL
    float* __restrict__ accum_out;          // [Q] or NULL (MSI hand-over)
    float* __restrict__ logT_out;           // [Q] or NULL
    float beta_loss;                        // already divided by norm_rays
    float sparsity_loss;
    float norm_factor;                      // 2/(3*norm_rays)
    int fused;
};

__global__ void __launch_bounds__(MARCH_THREADS)
march_bwd_kernel(GridView g, Opts opt, const float* __restrict__ origins,
                 const float* __restrict__ dirs, int64_t n_rays, BwdArgs a, GradView gr) {
    const int64_t ray_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (ray_id >= n_rays) return;
    const int lane = threadIdx.x & 31;
    const int B = g.basis_dim, D = g.sh_dim;
    const int k = lane % B, cgrp = lane / B;
    const bool chan_lane = lane < D;

    float ow[3], dw[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        ow[i] = __ldg(origins + ray_id * 3 + i);
        dw[i] = __ldg(dirs + ray_id * 3 + i);
    }
    const float Yk = sh_basis_k(B, k, dw);
    Ray ray;
    ray_setup(ray, ow, dw, g, opt);

    float gout[3], cc[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        cc[i] = __ldg(a.color_cache + ray_id * 3 + i);
        const float gi = __ldg(a.grad_in + ray_id * 3 + i);
        gout[i] = a.fused ? (cc[i] - gi) * a.norm_factor : gi;  // cuvol.cu:735-747
    }
    float accum = fmaf(cc[0], gout[0], fmaf(cc[1], gout[1], cc[2] * gout[2]));
    float beta_loss = a.beta_loss;
    float logT = 0.f;

    if (ray.tmin <= ray.tmax) {
        if (beta_loss > 0.f) {  // cuvol.cu:259-264
            const float transmit_in = fexp(a.logT_in ? a.logT_in[ray_id] : 0.f);
            beta_loss = (float)((double)beta_loss *
                                (1 - (double)transmit_in / (1 - (double)transmit_in + 1e-3)));
            accum += beta_loss;
        }
        const float my_gout = cgrp == 0 ? gout[0] : (cgrp == 1 ? gout[1] : gout[2]);
        float t = ray.tmin;
        while (t <= ray.tmax) {
            int l[3];
            float pos[3];
            locate(ray, g, t, l, pos);
            const int64_t base = (int64_t)l[0] * g.offx + (int64_t)l[1] * g.offy + l[2];
            int32_t mylink = -1;
            if (lane < 8) mylink = __ldg(g.links + base + corner_off(g, lane));
            const int32_t link0 = __shfl_sync(FULL, mylink, 0);
            const float skip = skip_dist(ray, link0, l, pos);
            if (skip >= opt.step_size) {
                t = skip_advance(t, skip, opt.step_size);
                continue;
            }
            float myv = 0.f;
            if (lane < 8 && mylink >= 0) myv = __ldg(g.density + (int64_t)mylink * g.density_stride);
            const float sigma = trilerp_lanes(myv, pos);
            if (opt.last_sample_opaque && __fadd_rn(t, opt.step_size) > ray.tmax)
                ray.world_step = 1e9f;
            if (sigma > opt.sigma_thresh) {
                int32_t lk[8];
                float v[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    lk[c] = __shfl_sync(FULL, mylink, c);
                    v[c] = (chan_lane && lk[c] >= 0) ? __ldg(g.sh + (int64_t)lk[c] * g.sh_stride + lane) : 0.f;
                }
                const float lct = seg_sum(trilerp8(v, pos) * Yk, k, B) + 0.5f;
                const float pcnt = __fmul_rn(ray.world_step, sigma);
                const float weight = __fmul_rn(fexp(logT), __fsub_rn(1.f, fexp(-pcnt)));
                logT = __fsub_rn(logT, pcnt);

                const float c0 = __shfl_sync(FULL, lct, 0);
                const float c1 = __shfl_sync(FULL, lct, B);
                const float c2 = __shfl_sync(FULL, lct, 2 * B);
                float total_color = fmaxf(c0, 0.f) * gout[0];
                total_color += fmaxf(c2, 0.f) * gout[2];
                total_color += fmaxf(c1, 0.f) * gout[1];
                const float my_c = cgrp == 0 ? c0 : (cgrp == 1 ? c1 : c2);
                const float in01 = (fmaxf(my_c, 0.f) == my_c) ? 1.f : 0.f;
                const float grad_common = weight * in01 * my_gout;
                const float curr_grad_color = Yk * grad_common;

                accum -= weight * total_color;
                float curr_grad_sigma = ray.world_step * (total_color * fexp(logT) - accum);
                if (a.sparsity_loss > 0.f)
                    curr_grad_sigma += a.sparsity_loss * (4 * sigma / (1 + 2 * (sigma * sigma)));

                if (chan_lane) {
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        if (lk[c] >= 0)
                            atomicAdd(gr.grad_sh + (int64_t)lk[c] * g.sh_stride + lane,
                                      corner_w(pos, c, curr_grad_color));
                }
                if (lane < 8 && mylink >= 0) {
                    atomicAdd(gr.grad_density + (int64_t)mylink * g.density_stride,
                              corner_w(pos, lane, curr_grad_sigma));
                    if (gr.mask) gr.mask[mylink] = 1;
                }
                if (fexp(logT) < opt.stop_thresh) break;
            }
            t = __fadd_rn(t, opt.step_size);
        }
        accum -= beta_loss;  // "cancel beta loss out in case of background", cuvol.cu:375-377
    }
    if (lane == 0) {
        if (a.accum_out) a.accum_out[ray_id] = accum;
        if (a.logT_out) a.logT_out[ray_id] = logT;
    }
}

// ---------------------------------------------------------------------------------------
// MSI background, one thread per ray
// ---------------------------------------------------------------------------------------
struct BgRay {
    float o[3], d[3];
    float world_step;
};

// ray_find_bounds_bg, render_util.cuh:630-676
__device__ __forceinline__ void bg_ray_setup(BgRay& r, const float* ow, const float* dw,
                                             const GridView& g) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        r.o[i] = __fmaf_rn(ow[i], g.scaling[i], g.offset[i]);
        r.d[i] = __fmul_rn(dw[i], g.scaling[i]);
    }
    const float delta_scale = rnorm3df(r.d[0], r.d[1], r.d[2]);
    r.world_step = delta_scale;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float ss = __fdiv_rn(2.f, (float)g.size[i]);
        r.o[i] = __fmaf_rn(__fadd_rn(r.o[i], 0.5f), ss, -1.f);
        r.d[i] = __fmul_rn(__fmul_rn(r.d[i], delta_scale), ss);
    }
    const float inorm = rnorm3df(r.d[0], r.d[1], r.d[2]);
    r.world_step = __fmul_rn(r.world_step, inorm);
#pragma unroll
    for (int i = 0; i < 3; ++i) r.d[i] = __fmul_rn(r.d[i], inorm);
}

struct BgSample {
    int l[3];
    float pos[3];
    int32_t lk[4];
    float invr_mid;
};

// position on shell -> equirect cell + layer (cuvol.cu:408-431, render_util.cuh:482-490)
__device__ __forceinline__ void bg_locate(const BgRay& r, const GridView& g, float t, BgSample& s) {
    const int R = g.bg_reso, L = g.bg_nlayers;
    float p[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) p[j] = fmaf(t, r.d[j], r.o[j]);
    s.invr_mid = rnorm3df(p[0], p[1], p[2]);
#pragma unroll
    for (int j = 0; j < 3; ++j) p[j] *= s.invr_mid;
    const float lat = asinf(p[1]);
    const float lon = atan2f(p[0], p[2]);
    s.pos[0] = R * 2 * (0.5 + lon * 0.5 * M_1_PI);
    s.pos[1] = R * (0.5 - lat * M_1_PI);
    s.pos[2] = fminf(fmaxf((1.f - s.invr_mid) * L - 0.5f, 0.f), L - 1);
#pragma unroll
    for (int j = 0; j < 3; ++j) s.l[j] = (int)s.pos[j];
    s.l[0] = min(s.l[0], R * 2 - 1);
    s.l[1] = min(s.l[1], R - 1);
    s.l[2] = min(s.l[2], L - 2);
#pragma unroll
    for (int j = 0; j < 3; ++j) s.pos[j] -= s.l[j];
    const int ny = s.l[1] < (R - 1) ? (s.l[1] + 1) : 0;
    const int nx = s.l[0] < (2 * R - 1) ? (s.l[0] + 1) : 0;
    s.lk[0] = __ldg(g.bg_links + R * s.l[0] + s.l[1]);
    s.lk[1] = __ldg(g.bg_links + R * s.l[0] + ny);
    s.lk[2] = __ldg(g.bg_links + R * nx + s.l[1]);
    s.lk[3] = __ldg(g.bg_links + R * nx + ny);
}

// All four channels of the MSI texel pair at once (one 16-byte load per layer instead of the
// reference's four scalar trilerps): trilerp_bg_one, render_util.cuh:157-188.
__device__ __forceinline__ float4 bg_trilerp4(const GridView& g, const BgSample& s) {
    float4 v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (s.lk[q] >= 0) {
            const float4* d = reinterpret_cast<const float4*>(
                g.bg_data + ((int64_t)s.lk[q] * g.bg_nlayers + s.l[2]) * 4);
            const float4 a = __ldg(d), b = __ldg(d + 1);
            v[q] = make_float4(lerp(a.x, b.x, s.pos[2]), lerp(a.y, b.y, s.pos[2]),
                               lerp(a.z, b.z, s.pos[2]), lerp(a.w, b.w, s.pos[2]));
        } else {
            v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    float4 r;
#define BG_BILERP(m) lerp(lerp(v[0].m, v[1].m, s.pos[1]), lerp(v[2].m, v[3].m, s.pos[1]), s.pos[0])
    r.x = BG_BILERP(x);
    r.y = BG_BILERP(y);
    r.z = BG_BILERP(z);
    r.w = BG_BILERP(w);
#undef BG_BILERP
    return r;
}

// trilerp_backward_bg_one, render_util.cuh:190-232, for all four channels of one sample.
__device__ __forceinline__ void bg_scatter(const GridView& g, const GradView& gr, const BgSample& s,
                                           const float* grad4, const bool* on4) {
    const float ay = 1.f - s.pos[1], az = 1.f - s.pos[2];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (s.lk[q] < 0) continue;
        const int64_t row = (int64_t)s.lk[q] * g.bg_nlayers + s.l[2];
        float* gp = gr.grad_bg + row * 4;
        const float wx = (q & 2) ? s.pos[0] : 1.0f - s.pos[0];
        const float wy = (q & 1) ? s.pos[1] : ay;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            if (!on4[ch]) continue;
            const float val = wy * (wx * grad4[ch]);
            atomicAdd(gp + ch, val * az);
            atomicAdd(gp + 4 + ch, val * s.pos[2]);
        }
        if (gr.mask_bg) {
            gr.mask_bg[row] = 1;
            gr.mask_bg[row + 1] = 1;
        }
    }
}

template <bool FROM_CAM>
__global__ void __launch_bounds__(BG_THREADS)
bg_fwd_kernel(GridView g, Opts opt, RaySource src, int64_t n_rays,
              const float* __restrict__ logT_in, float* __restrict__ out,
              unsigned long long* __restrict__ counters) {
    const int64_t ray_id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ray_id >= n_rays) return;
    float logT = logT_in[ray_id];
    if (logT < -25.f) return;
    float ow[3], dw[3];
    load_ray<FROM_CAM>(src, ray_id, ow, dw);
    BgRay r;
    bg_ray_setup(r, ow, dw, g);

    SphereIsect csi(r.o, r.d);
    const float cx = r.o[1] * r.d[2] - r.o[2] * r.d[1];
    const float cy = r.o[2] * r.d[0] - r.o[0] * r.d[2];
    const float cz = r.o[0] * r.d[1] - r.o[1] * r.d[0];
    const float inner_radius = fmaxf(norm3df(cx, cy, cz) + 1e-3f, 1.f);
    float t, invr_last = 1.f / inner_radius;
    const int n_steps = int(g.bg_nlayers / opt.step_size) + 2;
    float outv[3] = {0.f, 0.f, 0.f};
    unsigned n_act = 0;
    for (int i = 0; i < n_steps; ++i) {
        const float rad = n_steps / (n_steps - i - 0.5);
        if (rad < inner_radius || !csi.intersect(rad, &t)) continue;
        BgSample s;
        bg_locate(r, g, t, s);
        const float4 rgba = bg_trilerp4(g, s);
        const float sigma = rgba.w;
        if (sigma > 0.f) {
            ++n_act;
            const float pcnt = (invr_last - s.invr_mid) * r.world_step * sigma;
            const float weight = fexp(logT) * (1.f - fexp(-pcnt));
            logT -= pcnt;
            outv[0] += weight * fmaxf(rgba.x * kC0 + 0.5f, 0.f);
            outv[1] += weight * fmaxf(rgba.y * kC0 + 0.5f, 0.f);
            outv[2] += weight * fmaxf(rgba.z * kC0 + 0.5f, 0.f);
            if (fexp(logT) < opt.stop_thresh) break;
        }
        invr_last = s.invr_mid;
    }
    const float bgc = fexp(logT) * opt.background_brightness;
#pragma unroll
    for (int c = 0; c < 3; ++c) out[ray_id * 3 + c] += outv[c] + bgc;
    if (counters && n_act) atomicAdd(counters + 3, (unsigned long long)n_act);
}

__global__ void __launch_bounds__(BG_THREADS)
bg_bwd_kernel(GridView g, Opts opt, const float* __restrict__ origins,
              const float* __restrict__ dirs, int64_t n_rays, const float* __restrict__ grad_in,
              const float* __restrict__ color_cache, const float* __restrict__ logT_in,
              const float* __restrict__ accum_in, int fused, float norm_factor,
              float sparsity_loss, GradView gr) {
    const int64_t ray_id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ray_id >= n_rays) return;
    float logT = logT_in[ray_id];
    if (logT < -25.f) return;
    float ow[3], dw[3], gout[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        ow[i] = __ldg(origins + ray_id * 3 + i);
        dw[i] = __ldg(dirs + ray_id * 3 + i);
        const float gi = __ldg(grad_in + ray_id * 3 + i);
        gout[i] = fused ? (__ldg(color_cache + ray_id * 3 + i) - gi) * norm_factor : gi;
    }
    BgRay r;
    bg_ray_setup(r, ow, dw, g);
    float accum = accum_in[ray_id];

    SphereIsect csi(r.o, r.d);
    const int n_steps = int(g.bg_nlayers / opt.step_size) + 2;
    const float cx = r.o[1] * r.d[2] - r.o[2] * r.d[1];
    const float cy = r.o[2] * r.d[0] - r.o[0] * r.d[2];
    const float cz = r.o[0] * r.d[1] - r.o[1] * r.d[0];
    const float inner_radius = fmaxf(norm3df(cx, cy, cz) + 1e-3f, 1.f);
    float t, invr_last = 1.f / inner_radius;
    for (int i = 0; i < n_steps; ++i) {
        const float rad = n_steps / (n_steps - i - 0.5);
        if (rad < inner_radius || !csi.intersect(rad, &t)) continue;
        BgSample s;
        bg_locate(r, g, t, s);
        const float4 rgba = bg_trilerp4(g, s);
        const float sigma = rgba.w;
        if (sigma > 0.f) {
            const float pcnt = r.world_step * (invr_last - s.invr_mid) * sigma;
            const float weight = fexp(logT) * (1.f - fexp(-pcnt));
            logT -= pcnt;
            const float col[3] = {rgba.x * kC0 + 0.5f, rgba.y * kC0 + 0.5f, rgba.z * kC0 + 0.5f};
            float total_color = 0.f;
            float grad4[4];
            bool on4[4];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                total_color += fmaxf(col[c], 0.f) * gout[c];
                on4[c] = col[c] > 0.f;
                grad4[c] = kC0 * weight * gout[c];
            }
            accum -= weight * total_color;
            float curr_grad_sigma = r.world_step * (invr_last - s.invr_mid) *
                                    (total_color * fexp(logT) - accum);
            if (sparsity_loss > 0.f)
                curr_grad_sigma += sparsity_loss * (4 * sigma / (1 + 2 * (sigma * sigma)));
            grad4[3] = curr_grad_sigma;
            on4[3] = true;
            bg_scatter(g, gr, s, grad4, on4);
            if (fexp(logT) < opt.stop_thresh) break;
        }
        invr_last = s.invr_mid;
    }
}

// ---------------------------------------------------------------------------------------
// depth, one thread per ray (cuvol.cu:122-225)
// ---------------------------------------------------------------------------------------
template <bool SIGMA_THRESH>
__global__ void __launch_bounds__(MARCH_THREADS)
depth_kernel(GridView g, Opts opt, const float* __restrict__ origins,
             const float* __restrict__ dirs, int64_t n_rays, float thresh,
             float* __restrict__ out) {
    const int64_t ray_id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ray_id >= n_rays) return;
    float ow[3], dw[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        ow[i] = __ldg(origins + ray_id * 3 + i);
        dw[i] = __ldg(dirs + ray_id * 3 + i);
    }
    Ray ray;
    ray_setup(ray, ow, dw, g, opt);
    float outv = 0.f;
    if (ray.tmin <= ray.tmax) {
        float t = ray.tmin, logT = 0.f;
        while (t <= ray.tmax) {
            int l[3];
            float pos[3];
            locate(ray, g, t, l, pos);
            const int32_t* lp = g.links + ((int64_t)l[0] * g.offx + (int64_t)l[1] * g.offy + l[2]);
            const float skip = skip_dist(ray, __ldg(lp), l, pos);
            if (skip >= opt.step_size) {
                t = skip_advance(t, skip, opt.step_size);
                continue;
            }
            float v[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int32_t lk = __ldg(lp + corner_off(g, c));
                v[c] = lk >= 0 ? __ldg(g.density + (int64_t)lk * g.density_stride) : 0.f;
            }
            const float sigma = trilerp8(v, pos);
            if (SIGMA_THRESH) {
                if (sigma > thresh) {
                    outv = (t / opt.step_size) * ray.world_step;
                    break;
                }
            } else if (sigma > opt.sigma_thresh) {
                const float pcnt = __fmul_rn(ray.world_step, sigma);
                const float weight = fexp(logT) * (1.f - fexp(-pcnt));
                logT = __fsub_rn(logT, pcnt);
                outv += weight * (t / opt.step_size) * ray.world_step;
                if (fexp(logT) < opt.stop_thresh) break;
            }
            t = __fadd_rn(t, opt.step_size);
        }
    }
    out[ray_id] = outv;
}

}  // namespace v1

// ---------------------------------------------------------------------------------------
// host launchers used by render_abi.cu
// ---------------------------------------------------------------------------------------
using namespace v1;

static RaySource make_source(const MarchArgs& a) {
    return RaySource{a.origins, a.dirs, a.cam, a.row_begin};
}

// v1 (warp-per-ray, lane-per-channel) march kernels: kept as a debugging / A-B reference for the
// lattice-batch kernel; selected with SVOX2_B200_MARCH=v1.  Gradient strides must equal the
// data strides on this path.
void launch_v1_fwd(bool from_cam, const GridView& g, const Opts& opt, const MarchArgs& a,
                   cudaStream_t st) {
    const unsigned blocks = div_up(a.n_rays * WARP, MARCH_THREADS);
    const RaySource src = make_source(a);
    if (from_cam)
        march_fwd_kernel<true, false><<<blocks, MARCH_THREADS, 0, st>>>(g, opt, src, a.n_rays, a.rgb_out,
                                                                       a.logT_out, nullptr);
    else if (a.counters)
        march_fwd_kernel<false, true><<<blocks, MARCH_THREADS, 0, st>>>(g, opt, src, a.n_rays, a.rgb_out,
                                                                       a.logT_out, a.counters);
    else
        march_fwd_kernel<false, false><<<blocks, MARCH_THREADS, 0, st>>>(g, opt, src, a.n_rays, a.rgb_out,
                                                                        a.logT_out, nullptr);
}

void launch_v1_bwd(const GridView& g, const Opts& opt, const MarchArgs& a, const GradView& gr,
                   cudaStream_t st) {
    BwdArgs b{};
    b.grad_in = a.grad_in;
    b.color_cache = a.color_cache;
    b.logT_in = a.logT_in;
    b.accum_out = a.accum_out;
    b.logT_out = a.logT_out;
    b.beta_loss = a.beta_loss;
    b.sparsity_loss = a.sparsity_loss;
    b.norm_factor = a.norm_factor;
    b.fused = a.fused_seed;
    march_bwd_kernel<<<div_up(a.n_rays * WARP, MARCH_THREADS), MARCH_THREADS, 0, st>>>(
        g, opt, a.origins, a.dirs, a.n_rays, b, gr);
}

void launch_bg_fwd(bool from_cam, const GridView& g, const Opts& opt, const MarchArgs& a,
                   const float* logT, cudaStream_t st) {
    const RaySource src = make_source(a);
    const unsigned blocks = div_up(a.n_rays, BG_THREADS);
    if (from_cam)
        bg_fwd_kernel<true><<<blocks, BG_THREADS, 0, st>>>(g, opt, src, a.n_rays, logT, a.rgb_out, a.counters);
    else
        bg_fwd_kernel<false><<<blocks, BG_THREADS, 0, st>>>(g, opt, src, a.n_rays, logT, a.rgb_out, a.counters);
}

void launch_bg_bwd(const GridView& g, const Opts& opt, const MarchArgs& a, const float* color_cache,
                   const float* logT, const float* accum, const GradView& gr, cudaStream_t st) {
    bg_bwd_kernel<<<div_up(a.n_rays, BG_THREADS), BG_THREADS, 0, st>>>(
        g, opt, a.origins, a.dirs, a.n_rays, a.grad_in, color_cache, logT, accum, a.fused_seed,
        a.norm_factor, a.sparsity_loss, gr);
}

void launch_depth(bool sigma_thresh_mode, const GridView& g, const Opts& opt, const float* origins,
                  const float* dirs, int64_t n_rays, float thresh, float* out, cudaStream_t st) {
    const unsigned blocks = div_up(n_rays, MARCH_THREADS);
    if (sigma_thresh_mode)
        depth_kernel<true><<<blocks, MARCH_THREADS, 0, st>>>(g, opt, origins, dirs, n_rays, thresh, out);
    else
        depth_kernel<false><<<blocks, MARCH_THREADS, 0, st>>>(g, opt, origins, dirs, n_rays, thresh, out);
}

}  // namespace svox2b
