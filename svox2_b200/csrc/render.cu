// MSI background (warp per ray, lane = shell) and depth kernels (thread per ray); the foreground march
// lives in march.cu / queue_march.cu.  Traversal bit-exact with the reference:
//   MSI       render_lerp_kernel_cuvol.cu:386-607, 775-860
//   depth     render_lerp_kernel_cuvol.cu:122-225, 862-902
#include "common.cuh"
#include "march.cuh"

namespace svox2b {

namespace aux {

constexpr int MARCH_THREADS = 128;
constexpr int BG_THREADS = 128;

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
// MSI background: one warp per ray, lane = shell.  The reference walks the ~130 shells of a ray in
// one thread (render_background_kernel, cuvol.cu:386-476); a 5000-ray batch is then 40 CTAs of
// dependent asinf/atan2f -> links -> texel chains.  Here the 32 shells of a round locate, gather
// and interpolate in parallel; only the running log-transmittance is replayed in shell order.
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

// One ROUND of the MSI march: lane k looks at shell i0 + k (cuvol.cu:404-431).  The reference's
// `invr_last` -- 1/r of the previous shell the ray actually crossed -- is the only state carried
// from shell to shell besides the transmittance, and it does not depend on the samples' values, so
// it is resolved with one ballot: nearest lower crossing lane of the round, else the carry.
struct BgRound {
    BgSample s;
    float4 rgba;
    float dinv;     // invr_last - invr_mid
    bool crosses;   // the ray intersects this shell outside the inner radius
};

__device__ __forceinline__ void bg_round(const BgRay& r, const GridView& g, const SphereIsect& csi,
                                         float inner_radius, int n_steps, int i0, int lane,
                                         float& invr_carry, BgRound& b) {
    const int i = i0 + lane;
    float t = 0.f;
    const float rad = n_steps / (n_steps - i - 0.5);
    b.crosses = i < n_steps && !(rad < inner_radius) && csi.intersect(rad, &t);
    b.rgba = make_float4(0.f, 0.f, 0.f, 0.f);
    b.s.invr_mid = 0.f;
    if (b.crosses) {
        bg_locate(r, g, t, b.s);
        b.rgba = bg_trilerp4(g, b.s);
    }
    const unsigned cm = __ballot_sync(FULL, b.crosses);
    const unsigned lower = cm & ((1u << lane) - 1u);
    const int prev = lower ? 31 - __clz(lower) : 0;
    const float prev_invr = __shfl_sync(FULL, b.s.invr_mid, prev);
    b.dinv = __fsub_rn(lower ? prev_invr : invr_carry, b.s.invr_mid);
    if (cm) invr_carry = __shfl_sync(FULL, b.s.invr_mid, 31 - __clz(cm));
}

// Running log-transmittance over a round's active samples, in shell order (bit-exact running
// difference; the exp / stop test of all samples runs in parallel afterwards, cuvol.cu:440-452).
// Returns the active mask truncated behind the sample that fired the stop threshold.
__device__ __forceinline__ unsigned bg_chain(bool active, float pcnt, float stop_thresh, int lane,
                                             float& logT, bool& stop, float& my_before) {
    unsigned act_m = __ballot_sync(FULL, active);
    my_before = 0.f;
    unsigned it = act_m;
    while (it) {
        const int i = __ffs(it) - 1;
        it &= it - 1;
        const float p = __shfl_sync(FULL, pcnt, i);
        my_before = (lane == i) ? logT : my_before;
        logT = __fsub_rn(logT, p);
    }
    const float aft = __fsub_rn(my_before, pcnt);
    const unsigned tm = __ballot_sync(FULL, ((act_m >> lane) & 1u) && fexp(aft) < stop_thresh);
    if (tm) {
        const int sidx = __ffs(tm) - 1;
        act_m &= (2u << sidx) - 1u;
        stop = true;
        logT = __shfl_sync(FULL, aft, sidx);
    }
    return act_m;
}

// `gate`: optional device flag; the kernel runs only while (*gate != 0) == gate_want (queue path
// vs. its overflow fallback, queue_march.cu).
template <bool FROM_CAM>
__global__ void __launch_bounds__(BG_THREADS)
bg_fwd_kernel(GridView g, Opts opt, RaySource src, int64_t n_rays,
              const float* __restrict__ logT_in, float* __restrict__ out,
              unsigned long long* __restrict__ counters, const int* __restrict__ gate, int gate_want) {
    const int64_t ray_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (ray_id >= n_rays) return;
    if (gate != nullptr && (*gate != 0) != (gate_want != 0)) return;
    const int lane = threadIdx.x & 31;
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
    float invr_carry = 1.f / inner_radius;
    const int n_steps = int(g.bg_nlayers / opt.step_size) + 2;
    float outv[3] = {0.f, 0.f, 0.f};
    unsigned n_act = 0;
    bool stop = false;
    for (int i0 = 0; i0 < n_steps && !stop; i0 += WARP) {
        BgRound b;
        bg_round(r, g, csi, inner_radius, n_steps, i0, lane, invr_carry, b);
        const float sigma = b.rgba.w;
        const bool active = b.crosses && sigma > 0.f;
        const float pcnt = active ? __fmul_rn(__fmul_rn(b.dinv, r.world_step), sigma) : 0.f;
        float before;
        const unsigned act_m = bg_chain(active, pcnt, opt.stop_thresh, lane, logT, stop, before);
        n_act += __popc(act_m);
        if ((act_m >> lane) & 1u) {
            const float weight = fexp(before) * (1.f - fexp(-pcnt));
            outv[0] += weight * fmaxf(b.rgba.x * kC0 + 0.5f, 0.f);
            outv[1] += weight * fmaxf(b.rgba.y * kC0 + 0.5f, 0.f);
            outv[2] += weight * fmaxf(b.rgba.z * kC0 + 0.5f, 0.f);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) outv[c] += __shfl_xor_sync(FULL, outv[c], off);
    const float bgc = fexp(logT) * opt.background_brightness;
    if (lane < 3) {
        const float v = lane == 0 ? outv[0] : (lane == 1 ? outv[1] : outv[2]);
        out[ray_id * 3 + lane] += v + bgc;
    }
    if (counters && n_act && lane == 0) atomicAdd(counters + 3, (unsigned long long)n_act);
}

// trilerp_backward_bg_one (render_util.cuh:190-232) for the RGBA texel pair of each corner: one
// 16-byte vector reduction per texel instead of four scalar atomics (channels whose colour is
// clamped contribute an exact 0).
template <bool VEC>
__device__ __forceinline__ void bg_scatter4(const GridView& g, const GradView& gr, const BgSample& s,
                                            const float* grad4, const bool* on4) {
    if (!VEC) {
        bg_scatter(g, gr, s, grad4, on4);
        return;
    }
    const float ay = 1.f - s.pos[1], az = 1.f - s.pos[2];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (s.lk[q] < 0) continue;
        const int64_t row = (int64_t)s.lk[q] * g.bg_nlayers + s.l[2];
        float* gp = gr.grad_bg + row * 4;
        const float wx = (q & 2) ? s.pos[0] : 1.0f - s.pos[0];
        const float wy = (q & 1) ? s.pos[1] : ay;
        float val[4];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) val[ch] = on4[ch] ? wy * (wx * grad4[ch]) : 0.f;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gp), "f"(val[0] * az),
                     "f"(val[1] * az), "f"(val[2] * az), "f"(val[3] * az)
                     : "memory");
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gp + 4), "f"(val[0] * s.pos[2]),
                     "f"(val[1] * s.pos[2]), "f"(val[2] * s.pos[2]), "f"(val[3] * s.pos[2])
                     : "memory");
        if (gr.mask_bg) {
            gr.mask_bg[row] = 1;
            gr.mask_bg[row + 1] = 1;
        }
    }
}

template <bool VEC>
__global__ void __launch_bounds__(BG_THREADS)
bg_bwd_kernel(GridView g, Opts opt, const float* __restrict__ origins,
              const float* __restrict__ dirs, int64_t n_rays, const float* __restrict__ grad_in,
              const float* __restrict__ color_cache, const float* __restrict__ logT_in,
              const float* __restrict__ accum_in, int fused, float norm_factor,
              float sparsity_loss, GradView gr, const int* __restrict__ gate, int gate_want) {
    const int64_t ray_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (ray_id >= n_rays) return;
    if (gate != nullptr && (*gate != 0) != (gate_want != 0)) return;
    const int lane = threadIdx.x & 31;
    float logT = logT_in[ray_id];
    if (logT < -25.f) return;
    float ow[3], dw[3], gout[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        ow[i] = __ldg(origins + ray_id * 3 + i);
        dw[i] = __ldg(dirs + ray_id * 3 + i);
        const float gi = __ldg(grad_in + ray_id * 3 + i);
        gout[i] = fused ? (__ldcg(color_cache + ray_id * 3 + i) - gi) * norm_factor : gi;
    }
    BgRay r;
    bg_ray_setup(r, ow, dw, g);
    float accum = __ldcg(accum_in + ray_id);

    SphereIsect csi(r.o, r.d);
    const int n_steps = int(g.bg_nlayers / opt.step_size) + 2;
    const float cx = r.o[1] * r.d[2] - r.o[2] * r.d[1];
    const float cy = r.o[2] * r.d[0] - r.o[0] * r.d[2];
    const float cz = r.o[0] * r.d[1] - r.o[1] * r.d[0];
    const float inner_radius = fmaxf(norm3df(cx, cy, cz) + 1e-3f, 1.f);
    float invr_carry = 1.f / inner_radius;
    bool stop = false;
    for (int i0 = 0; i0 < n_steps && !stop; i0 += WARP) {
        BgRound b;
        bg_round(r, g, csi, inner_radius, n_steps, i0, lane, invr_carry, b);
        const float sigma = b.rgba.w;
        const bool active = b.crosses && sigma > 0.f;
        const float pcnt = active ? __fmul_rn(__fmul_rn(r.world_step, b.dinv), sigma) : 0.f;
        float before;
        const unsigned act_m = bg_chain(active, pcnt, opt.stop_thresh, lane, logT, stop, before);
        const bool on = (act_m >> lane) & 1u;
        const float weight = on ? fexp(before) * (1.f - fexp(-pcnt)) : 0.f;
        const float col[3] = {b.rgba.x * kC0 + 0.5f, b.rgba.y * kC0 + 0.5f, b.rgba.z * kC0 + 0.5f};
        float total_color = 0.f;
        float grad4[4];
        bool on4[4];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            total_color += fmaxf(col[c], 0.f) * gout[c];
            on4[c] = col[c] > 0.f;
            grad4[c] = kC0 * weight * gout[c];
        }
        // the reference's running `accum -= weight * total_color` as a warp scan in shell order
        float pre = weight * total_color;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const float o = __shfl_up_sync(FULL, pre, off);
            if (lane >= off) pre += o;
        }
        const float my_accum = accum - pre;
        accum -= __shfl_sync(FULL, pre, 31);
        if (on) {
            float curr_grad_sigma = r.world_step * b.dinv * (total_color * fexp(__fsub_rn(before, pcnt)) - my_accum);
            if (sparsity_loss > 0.f)
                curr_grad_sigma += sparsity_loss * (4 * sigma / (1 + 2 * (sigma * sigma)));
            grad4[3] = curr_grad_sigma;
            on4[3] = true;
            bg_scatter4<VEC>(g, gr, b.s, grad4, on4);
        }
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

}  // namespace aux

// ---------------------------------------------------------------------------------------
// host launchers used by render_abi.cu
// ---------------------------------------------------------------------------------------
using namespace aux;

static RaySource make_source(const MarchArgs& a) {
    return RaySource{a.origins, a.dirs, a.cam, a.row_begin};
}

void launch_bg_fwd(bool from_cam, const GridView& g, const Opts& opt, const MarchArgs& a,
                   const float* logT, cudaStream_t st, const int* gate, int gate_want) {
    const RaySource src = make_source(a);
    const unsigned blocks = div_up(a.n_rays * WARP, BG_THREADS);   // one warp per ray
    if (from_cam)
        bg_fwd_kernel<true><<<blocks, BG_THREADS, 0, st>>>(g, opt, src, a.n_rays, logT, a.rgb_out, a.counters,
                                                           gate, gate_want);
    else
        bg_fwd_kernel<false><<<blocks, BG_THREADS, 0, st>>>(g, opt, src, a.n_rays, logT, a.rgb_out, a.counters,
                                                            gate, gate_want);
}

void launch_bg_bwd(const GridView& g, const Opts& opt, const MarchArgs& a, const float* color_cache,
                   const float* logT, const float* accum, const GradView& gr, cudaStream_t st,
                   const int* gate, int gate_want) {
    const unsigned blocks = div_up(a.n_rays * WARP, BG_THREADS);
    const bool vec = (reinterpret_cast<uintptr_t>(gr.grad_bg) & 15u) == 0;
    if (vec)
        bg_bwd_kernel<true><<<blocks, BG_THREADS, 0, st>>>(g, opt, a.origins, a.dirs, a.n_rays, a.grad_in,
                                                           color_cache, logT, accum, a.fused_seed, a.norm_factor,
                                                           a.sparsity_loss, gr, gate, gate_want);
    else
        bg_bwd_kernel<false><<<blocks, BG_THREADS, 0, st>>>(g, opt, a.origins, a.dirs, a.n_rays, a.grad_in,
                                                            color_cache, logT, accum, a.fused_seed, a.norm_factor,
                                                            a.sparsity_loss, gr, gate, gate_want);
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
