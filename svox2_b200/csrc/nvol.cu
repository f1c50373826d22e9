// `nvol` backend: the Neural-Volumes compositing rule (absolute instead of relative transmittance)
// behind volume_render_nvol / _backward / _fused.  Reference: render_lerp_kernel_nvol.cu:30-296
// (device code), :298-401 (kernels), :406-502 (launchers).  It is the backend the reference's own
// parity script exercises (test/test_render_gradcheck.py:20).
//
// The reference runs the cuvol lane-per-channel warp; here the ray is walked with the same
// warp-cooperative lattice as the cuvol kernels (traverse.cuh: lane = lattice candidate, so the
// sample set, the saturation decision and therefore the mask are bit-exact) and active samples
// are shaded four at a time by 8-lane groups with 128-bit row loads / vector reductions
// (quad.cuh).  Only the compositing differs from cuvol:
//
//   alpha_total' = min(alpha_total + 1 - exp(-delta*sigma), 1);  weight = alpha_total' - alpha_total
//   the march ends at the first sample where alpha_total' reaches 1 (that sample still shades)
//
// Backward = the reference's two passes: (1) over active samples: SH gradients, and for all but
// the saturating sample dL/dsigma = delta*exp(-delta*sigma)*<colour, gout> (+ Cauchy sparsity);
// (2) when the ray saturated (or the background term is on): every VISITED sample before the
// saturation point -- sub-threshold ones included, as upstream -- receives
// -delta*exp(-delta*sigma)*<last colour | background, gout>.  MSI backgrounds are not supported
// by this backend (upstream: "Background is NOT supported").
#include "common.cuh"
#include "march.cuh"
#include "quad.cuh"
#include "traverse.cuh"

namespace svox2b {
namespace {

constexpr int NV_THREADS = 64;

using quad::LaneConst;

struct RayCtx {
    Ray ray;
    LaneConst lc;
    int64_t ray_id;
    int lane;
};

__device__ __forceinline__ bool ray_begin(const GridView& g, const Opts& opt, const MarchArgs& a,
                                          RayCtx& rc) {
    rc.ray_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    rc.lane = threadIdx.x & 31;
    if (rc.ray_id >= a.n_rays) return false;
    float ow[3], dw[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        ow[i] = __ldg(a.origins + rc.ray_id * 3 + i);
        dw[i] = __ldg(a.dirs + rc.ray_id * 3 + i);
    }
    quad::lane_const_init(rc.lc, rc.lane, g.basis_dim, g.sh_dim, dw);  // raw input direction
    ray_setup(rc.ray, ow, dw, g, opt);
    return true;
}

// One round of 32 lattice candidates: which are samples, their sigma, and exp(-delta*sigma).
struct RoundOut {
    Candidates c;
    unsigned members;
    float sigma, ct, t_next;
};

__device__ __forceinline__ void next_round(const GridView& g, const Opts& opt, const Ray& ray,
                                           float t0, int lane, bool& done, RoundOut& r) {
    load_candidates<true>(ray, g, t0, opt.step_size, lane, r.c);
    unsigned n_skips = 0;
    r.t_next = 0.f;
    r.members = resolve_lattice(ray, r.c, opt.step_size, r.t_next, done, n_skips);
    r.sigma = member_sigma(g, r.c, (r.members >> lane) & 1u);
    r.ct = fexp(__fmul_rn(-ray.world_step, r.sigma));  // nvol.cu:89-90, :174
}

// Saturating alpha chain over the lanes of `mask` in sample order (nvol.cu:89-92, :175-178, :266).
// my_w = weight of this lane's sample; sat_lane = lane whose sample brought alpha_total to 1 (-1).
__device__ __forceinline__ void alpha_chain(unsigned mask, float ct, int lane, float& ta, float& my_w,
                                            int& sat_lane) {
    my_w = 0.f;
    sat_lane = -1;
    unsigned it = mask;
    while (it) {
        const int i = __ffs(it) - 1;
        it &= it - 1;
        const float p = __shfl_sync(FULL, ct, i);
        const float na = fminf(__fsub_rn(__fadd_rn(ta, 1.f), p), 1.f);
        my_w = (lane == i) ? __fsub_rn(na, ta) : my_w;
        ta = na;
        if (na >= 1.f) {
            sat_lane = i;
            break;
        }
    }
}

// Hands the round's active samples to the four 8-lane groups, four per pass.
struct GroupSample {
    bool has;
    int src;
    int32_t L[8];
    float wq[8];
    float w;
};

__device__ __forceinline__ void pick_samples(unsigned& todo, const Candidates& c, float my_w, int lane,
                                             GroupSample& s) {
    const int grp = lane >> 3;
    int src = -1;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int b = todo ? __ffs(todo) - 1 : -1;
        src = (grp == k) ? b : src;
        todo &= todo - 1;
    }
    s.has = src >= 0;
    s.src = src;
    const int from = s.has ? src : 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) s.L[q] = __shfl_sync(FULL, c.lk[q], from);
    const float px = __shfl_sync(FULL, c.pos[0], from);
    const float py = __shfl_sync(FULL, c.pos[1], from);
    const float pz = __shfl_sync(FULL, c.pos[2], from);
    s.w = __shfl_sync(FULL, my_w, from);
    quad::corner_weights(px, py, pz, s.wq);
}

// lane qi of a group owns corner qi: density gradient + mask (trilerp_backward_cuvol_one_density)
__device__ __forceinline__ void scatter_density_group(const GradView& gr, int64_t gd_stride,
                                                      const GroupSample& s, int qi, float cgs) {
    int32_t Lown = s.L[0];
    float wown = s.wq[0];
#pragma unroll
    for (int q = 1; q < 8; ++q) {
        Lown = (qi == q) ? s.L[q] : Lown;
        wown = (qi == q) ? s.wq[q] : wown;
    }
    if (Lown >= 0) {
        atomicAdd(gr.grad_density + (int64_t)Lown * gd_stride, wown * cgs);
        if (gr.mask) gr.mask[Lown] = 1;
    }
}

// render_ray_kernel + trace_ray_nvol, nvol.cu:30-111, :298-331
template <bool VEC>
__global__ void __launch_bounds__(NV_THREADS) nvol_forward_kernel(GridView g, Opts opt, MarchArgs a) {
    RayCtx rc;
    if (!ray_begin(g, opt, a, rc)) return;
    const int lane = rc.lane;
    const float bg_brightness = g.bg_nlayers == 0 ? opt.background_brightness : 0.f;
    if (rc.ray.tmin > rc.ray.tmax) {
        if (lane < 3) a.rgb_out[rc.ray_id * 3 + lane] = bg_brightness;
        return;
    }
    float out3[3] = {0.f, 0.f, 0.f};
    float ta = 0.f, t0 = rc.ray.tmin;
    bool done = false;
    while (!done) {
        RoundOut r;
        next_round(g, opt, rc.ray, t0, lane, done, r);
        const bool active = ((r.members >> lane) & 1u) && r.sigma > opt.sigma_thresh;
        unsigned act_m = __ballot_sync(FULL, active);
        float my_w;
        int sat;
        alpha_chain(act_m, r.ct, lane, ta, my_w, sat);
        if (sat >= 0) {
            act_m &= (2u << sat) - 1u;
            done = true;
        }
        unsigned todo = act_m;
        while (todo) {
            GroupSample s;
            pick_samples(todo, r.c, my_w, lane, s);
            float col[3];
            quad::gather_colour<VEC>(g, rc.lc, s.L, s.wq, s.has, col);
            if (s.has) {
#pragma unroll
                for (int c = 0; c < 3; ++c) out3[c] = fmaf(s.w, fmaxf(col[c], 0.f), out3[c]);
            }
        }
        t0 = r.t_next;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float x = out3[c];
        x += __shfl_xor_sync(FULL, x, 8);
        x += __shfl_xor_sync(FULL, x, 16);
        out3[c] = fmaf(1.f - ta, bg_brightness, x);  // nvol.cu:104-106
    }
    if (lane < 3) a.rgb_out[rc.ray_id * 3 + lane] = lane == 0 ? out3[0] : (lane == 1 ? out3[1] : out3[2]);
}

// render_ray_backward_kernel + trace_ray_nvol_backward, nvol.cu:113-296, :333-401
template <bool VEC>
__global__ void __launch_bounds__(NV_THREADS)
nvol_backward_kernel(GridView g, Opts opt, MarchArgs a, GradView gr) {
    RayCtx rc;
    if (!ray_begin(g, opt, a, rc)) return;
    if (rc.ray.tmin > rc.ray.tmax) return;
    const int lane = rc.lane;
    const int D = g.sh_dim;
    float gout[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float gi = __ldg(a.grad_in + rc.ray_id * 3 + i);
        gout[i] = a.fused_seed ? (__ldg(a.color_cache + rc.ray_id * 3 + i) - gi) * a.norm_factor : gi;
    }
    const float ws = rc.ray.world_step;

    // ---- pass 1: active samples ----
    float ta = 0.f, t0 = rc.ray.tmin;
    float last_total = 0.f;
    bool done = false;
    while (!done) {
        RoundOut r;
        next_round(g, opt, rc.ray, t0, lane, done, r);
        const bool active = ((r.members >> lane) & 1u) && r.sigma > opt.sigma_thresh;
        unsigned act_m = __ballot_sync(FULL, active);
        float my_w;
        int sat;
        alpha_chain(act_m, r.ct, lane, ta, my_w, sat);
        if (sat >= 0) {
            act_m &= (2u << sat) - 1u;
            done = true;
            rc.ray.tmax = __shfl_sync(FULL, r.c.tk, sat);  // nvol.cu:238: the remat pass ends here
        }
        unsigned todo = act_m;
        while (todo) {
            GroupSample s;
            pick_samples(todo, r.c, my_w, lane, s);
            const int from = s.has ? s.src : 0;
            const float sg = __shfl_sync(FULL, r.sigma, from);
            const float ct = __shfl_sync(FULL, r.ct, from);
            float col[3];
            quad::gather_colour<VEC>(g, rc.lc, s.L, s.wq, s.has, col);
            float gc[3], tc[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float cl = fmaxf(col[c], 0.f);
                tc[c] = cl * gout[c];
                gc[c] = (cl == col[c]) ? s.w * gout[c] : 0.f;
            }
            float total = tc[0];  // leader lanes: c0 += c2; c0 += c1 (nvol.cu:188-190)
            total += tc[2];
            total += tc[1];
            const bool is_last = s.has && s.src == sat;
            if (s.has) {
                if (rc.lc.quad_on) {
                    float ge[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        ge[e] = fmaf(rc.lc.Yc[2][e], gc[2], fmaf(rc.lc.Yc[1][e], gc[1], rc.lc.Yc[0][e] * gc[0]));
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        if (s.L[q] >= 0)
                            quad::red_quad<VEC>(gr.grad_sh + (int64_t)s.L[q] * a.gsh_stride + 4 * rc.lc.qi,
                                                rc.lc.qi, D, s.wq[q] * ge[0], s.wq[q] * ge[1],
                                                s.wq[q] * ge[2], s.wq[q] * ge[3]);
                }
                if (!is_last) {
                    float cgs = ws * ct * total;
                    if (a.sparsity_loss > 0.f) cgs += a.sparsity_loss * (4 * sg / (1 + 2 * (sg * sg)));
                    scatter_density_group(gr, a.gd_stride, s, rc.lc.qi, cgs);
                }
            }
            const unsigned lm = __ballot_sync(FULL, is_last);
            if (lm) last_total = __shfl_sync(FULL, total, __ffs(lm) - 1);
        }
        t0 = r.t_next;
    }
    if (ta < 1.f)  // never saturated: the background term's gradient (nvol.cu:243-247)
        last_total = opt.background_brightness * (gout[0] + gout[1] + gout[2]);
    if (last_total == 0.f) return;

    // ---- pass 2: every visited sample before saturation, one lane per sample ----
    ta = 0.f;
    t0 = rc.ray.tmin;
    done = false;
    while (!done) {
        RoundOut r;
        next_round(g, opt, rc.ray, t0, lane, done, r);
        float my_w;
        int sat;
        alpha_chain(r.members, r.ct, lane, ta, my_w, sat);
        unsigned mem = r.members;
        if (sat >= 0) {
            mem &= (1u << sat) - 1u;  // the saturating sample itself gets nothing (nvol.cu:270)
            done = true;
        }
        if ((mem >> lane) & 1u) {
            const float cgs = -ws * r.ct * last_total;
            float wq[8];
            quad::corner_weights(r.c.pos[0], r.c.pos[1], r.c.pos[2], wq);
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (r.c.lk[q] >= 0) {
                    atomicAdd(gr.grad_density + (int64_t)r.c.lk[q] * a.gd_stride, wq[q] * cgs);
                    if (gr.mask) gr.mask[r.c.lk[q]] = 1;
                }
        }
        t0 = r.t_next;
    }
}

int validate_common(const svox2_grid_t* grid, const float* origins, const float* dirs, int64_t n_rays,
                    const svox2_render_opts_t* opt, const char* who) {
    if (int e = validate_grid(grid, who)) return e;
    if (n_rays < 0 || (n_rays > 0 && (!origins || !dirs))) {
        set_error("%s: rays must be non-null [Q,3] float32 device arrays", who);
        return SVOX2_ERR_INVALID;
    }
    if (!opt || !(opt->step_size > 0.f)) {
        set_error("%s: opt must be non-null with a positive step_size", who);
        return SVOX2_ERR_INVALID;
    }
    return SVOX2_OK;
}

MarchArgs nvol_args(const float* origins, const float* dirs, int64_t n_rays) {
    MarchArgs a{};
    a.origins = origins;
    a.dirs = dirs;
    a.n_rays = n_rays;
    return a;
}

void launch_fwd(const GridView& g, const Opts& opt, const MarchArgs& a, cudaStream_t st) {
    const unsigned blocks = div_up(a.n_rays * WARP, NV_THREADS);
    if (march_vec_ok(g, nullptr, 0)) nvol_forward_kernel<true><<<blocks, NV_THREADS, 0, st>>>(g, opt, a);
    else nvol_forward_kernel<false><<<blocks, NV_THREADS, 0, st>>>(g, opt, a);
}

void launch_bwd(const GridView& g, const Opts& opt, const MarchArgs& a, const GradView& gr,
                cudaStream_t st) {
    const unsigned blocks = div_up(a.n_rays * WARP, NV_THREADS);
    if (march_vec_ok(g, &gr, a.gsh_stride))
        nvol_backward_kernel<true><<<blocks, NV_THREADS, 0, st>>>(g, opt, a, gr);
    else
        nvol_backward_kernel<false><<<blocks, NV_THREADS, 0, st>>>(g, opt, a, gr);
}

GradView grad_view(const svox2_grads_t* gr) {
    GradView v;
    v.grad_density = gr->grad_density_out;
    v.grad_sh = gr->grad_sh_out;
    v.grad_bg = nullptr;
    v.mask = gr->mask_out;
    v.mask_bg = nullptr;
    return v;
}

}  // namespace
}  // namespace svox2b

using namespace svox2b;

extern "C" SVOX2_API int svox2_volume_render_nvol(const svox2_grid_t* grid, const float* origins,
                                                  const float* dirs, int64_t n_rays,
                                                  const svox2_render_opts_t* opt, float* out,
                                                  svox2_stream_t stream) {
    if (int e = validate_common(grid, origins, dirs, n_rays, opt, __func__)) return e;
    SVOX2_REQUIRE(out || n_rays == 0, "out must not be null");
    if (n_rays == 0) return SVOX2_OK;
    MarchArgs a = nvol_args(origins, dirs, n_rays);
    a.rgb_out = out;
    launch_fwd(make_grid_view(grid), *opt, a, (cudaStream_t)stream);
    return check_launch(__func__);
}

extern "C" SVOX2_API int svox2_volume_render_nvol_backward(
    const svox2_grid_t* grid, const float* origins, const float* dirs, int64_t n_rays,
    const svox2_render_opts_t* opt, const float* grad_out, const float* color_cache,
    const svox2_grads_t* grads, svox2_stream_t stream) {
    if (int e = validate_common(grid, origins, dirs, n_rays, opt, __func__)) return e;
    SVOX2_REQUIRE(grads && grads->grad_density_out && grads->grad_sh_out,
                  "grad_density_out and grad_sh_out are required");
    SVOX2_REQUIRE((grad_out && color_cache) || n_rays == 0, "grad_out/color_cache must not be null");
    if (n_rays == 0) return SVOX2_OK;
    MarchArgs a = nvol_args(origins, dirs, n_rays);
    a.grad_in = grad_out;
    a.color_cache = color_cache;
    a.gsh_stride = grads->grad_sh_stride > 0 ? grads->grad_sh_stride : grid->sh_stride;
    a.gd_stride = grads->grad_density_stride > 0 ? grads->grad_density_stride : grid->density_stride;
    launch_bwd(make_grid_view(grid), *opt, a, grad_view(grads), (cudaStream_t)stream);
    return check_launch(__func__);
}

extern "C" SVOX2_API int svox2_volume_render_nvol_fused(
    const svox2_grid_t* grid, const float* origins, const float* dirs, int64_t n_rays,
    int64_t norm_rays, const svox2_render_opts_t* opt, const float* rgb_gt, float sparsity_loss,
    float* rgb_out, const svox2_grads_t* grads, svox2_stream_t stream) {
    if (int e = validate_common(grid, origins, dirs, n_rays, opt, __func__)) return e;
    SVOX2_REQUIRE(grads && grads->grad_density_out && grads->grad_sh_out,
                  "grad_density_out and grad_sh_out are required");
    SVOX2_REQUIRE((rgb_gt && rgb_out) || n_rays == 0, "rgb_gt/rgb_out must not be null");
    SVOX2_REQUIRE(norm_rays > 0 || n_rays == 0, "norm_rays must be positive");
    if (n_rays == 0) return SVOX2_OK;
    const GridView g = make_grid_view(grid);
    MarchArgs a = nvol_args(origins, dirs, n_rays);
    a.rgb_out = rgb_out;
    launch_fwd(g, *opt, a, (cudaStream_t)stream);
    a.grad_in = rgb_gt;          // nvol.cu:487-497: residual = rgb_out - rgb_gt
    a.color_cache = rgb_out;
    a.fused_seed = 1;
    a.norm_factor = 2.f / (float)(3 * (int)norm_rays);  // nvol.cu:365
    a.sparsity_loss = sparsity_loss;
    a.gsh_stride = grads->grad_sh_stride > 0 ? grads->grad_sh_stride : grid->sh_stride;
    a.gd_stride = grads->grad_density_stride > 0 ? grads->grad_density_stride : grid->density_stride;
    launch_bwd(g, *opt, a, grad_view(grads), (cudaStream_t)stream);
    return check_launch(__func__);
}
