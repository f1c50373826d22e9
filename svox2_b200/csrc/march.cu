// B200 ray-march kernel ("lattice batch" design): forward, backward and single-kernel fused
// forward+backward of the cuvol renderer.
//
// One warp per ray, but -- unlike the reference, where 27 lanes redo the same scalar traversal
// and only the SH gather is parallel -- the 32 lanes are used twice over:
//
//  (1) TRAVERSAL: lane k evaluates lattice candidate t_k = t_0 + k*step (k sequentially rounded
//      adds, exactly the reference's t += step chain), does its own cell lookup, link probe and
//      skip test.  Ballots resolve which candidates are really on the ray's sample lattice
//      (a skip at lane j jumps to t_j + ceil(skip/step)*step; if that value is bit-equal to a
//      later candidate the batch continues there, otherwise a new batch starts at it).  Member
//      lanes then fetch their 8 links + 8 sigmas (16 independent loads per lane in flight) and
//      do the z,y,x fmaf trilerp.  The log-transmittance chain and the stop test are replayed
//      sequentially over the active lanes with shuffles, so termination is bit-exact too.
//  (2) SAMPLES: active samples are processed four at a time; each 8-lane group owns one sample
//      and lane i of the group owns SH channels 4i..4i+3: one 128-bit load per corner row
//      (row stride padded to 28 floats), trilinear weights applied as 32 FMAs, colours by a
//      masked-basis dot product and a 3-step butterfly.  The backward scatters with one 128-bit
//      vector reduction (red.global.add.v4.f32) per corner row instead of 27 scalar atomics.
//
// Fused mode runs the forward pass, seeds the MSE gradient from the ray's own colour and then
// replays the ray for the backward pass inside the same warp: one launch, no rgb/logT round
// trip through HBM, rows re-gathered while still cache-hot.
//
// Parity: the lattice, cell indices, sigma, log-transmittance chain and therefore the sparse
// mask are bit-exact with render_lerp_kernel_cuvol.cu:30-120,227-383 (see common.cuh); colours
// and gradients are reassociated (weights-then-dot instead of lerp-per-channel) and agree to
// ~1e-6 relative.
#include "common.cuh"
#include "march.cuh"
#include "traverse.cuh"

namespace svox2b {
namespace {

constexpr int LB_THREADS = 64;  // 2 rays per CTA: fine-grained tail balance at 5000 rays / launch

enum { MODE_FWD = 0, MODE_BWD = 1, MODE_FUSED = 2 };


// Per-lane constants of the quad mapping: lane i of each 8-lane group owns channels 4i..4i+3.
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

// Tunables (A/B-tested on B200 with tools/build_variants.sh; see profiles/):
#ifndef LB_PREFETCH_SH
#define LB_PREFETCH_SH 2      // 0 off, 1 prefetch SH rows to L1, 2 to L2
#endif
#ifndef LB_PREFETCH_LINKS
#define LB_PREFETCH_LINKS 1   // prefetch the next batch's link lines during the current batch
#endif
#ifndef LB_LINKS_TWO_PHASE
#define LB_LINKS_TWO_PHASE 1
#endif
#ifndef LB_MIN_BLOCKS_FWD
#define LB_MIN_BLOCKS_FWD 16  // forward-only instantiations (image rendering is issue-bound)
#endif
#ifndef LB_IMAGE_TILES
#define LB_IMAGE_TILES 1   // image rendering: warps are mapped to pixels in 2-D tiles
#endif
#ifndef LB_MIN_BLOCKS
#define LB_MIN_BLOCKS 16      // __launch_bounds__ min CTAs/SM: 64-register cap, 32 warps/SM (A/B: profiles/r1_ab_variants.txt)
#endif

__device__ __forceinline__ void prefetch_line(const void* p) {
#if LB_PREFETCH_SH == 1
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
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
        // sm_90+ 128-bit vector reduction: one L2 atomic transaction for four channels
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

struct PassOut {
    float out3[3];  // forward: this group's partial colour
    float logT;     // running log transmittance at the end of the pass
    bool stopped;   // forward: terminated by stop_thresh
};

// One traversal of the ray.  BWD=false: accumulate colour.  BWD=true: scatter gradients.
template <bool BWD, bool VEC, bool COUNT>
__device__ __forceinline__ void march_pass(const GridView& g, const Opts& opt, const Ray& ray,
                                           const LaneConst& lc, int lane, PassOut& po,
                                           const float* gout, float& accum, float sparsity_loss,
                                           const GradView& gr, int64_t gsh_stride,
                                           int64_t gd_stride, unsigned* cnt,
                                           float4* __restrict__ ccache, int ccap) {
    const int grp = lane >> 3;
    int n_seen = 0;  // active samples of this ray processed so far (warp-uniform)
    const float step = opt.step_size;
    const int D = g.sh_dim;
    float t0 = ray.tmin;
    float lt = 0.f;
    bool done = false;
    po.out3[0] = po.out3[1] = po.out3[2] = 0.f;
    po.stopped = false;

    while (!done) {
        // ---- (1)+(2) lattice candidates and which of them are samples (traverse.cuh) ----
        Candidates c;
        load_candidates<LB_LINKS_TWO_PHASE != 0>(ray, g, t0, step, lane, c);
        float t_next = 0.f;
        unsigned n_skips = 0;
        unsigned members = resolve_lattice(ray, c, step, t_next, done, n_skips);
        if (COUNT) cnt[0] += n_skips;
#if LB_PREFETCH_LINKS
        if (!done) prefetch_next_links(ray, g, t_next, step, lane);
#endif

        // ---- (3) sigma at member candidates ----
        const bool member = (members >> lane) & 1u;
        const float sigma = member_sigma(g, c, member);
        const float tk = c.tk;
        const int32_t (&lk)[8] = c.lk;
        const float (&pos)[3] = c.pos;

        // ---- (4) log-transmittance chain + stop test, replayed in sample order ----
        const float ws = (opt.last_sample_opaque && __fadd_rn(tk, step) > ray.tmax) ? 1e9f
                                                                                   : ray.world_step;
        const bool active = member && (sigma > opt.sigma_thresh);
#if LB_PREFETCH_SH
        // Start pulling this lane's 8 SH rows towards the SM now: the rounds below consume them
        // one dependent round trip later, after the (serial) transmittance chain.  A 112-byte row
        // touches at most two 128-byte lines.
        if (active && (!BWD || ccache == nullptr)) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (c.lk[q] >= 0) {
                    const float* row = g.sh + (int64_t)c.lk[q] * g.sh_stride;
                    prefetch_line(row);
                    prefetch_line(row + (D - 1));
                }
        }
#endif
        const float pcnt = __fmul_rn(ws, sigma);
        float my_before;
        unsigned act_m = transmittance_chain(active, pcnt, opt.stop_thresh, lane, lt, members, done,
                                             po.stopped, my_before);
        if (COUNT) {
            cnt[1] += __popc(members);
            cnt[2] += __popc(act_m);
        }
        const float weight = __fmul_rn(fexp(my_before), __fsub_rn(1.f, fexp(-pcnt)));
        const float my_after = __fsub_rn(my_before, pcnt);

        // ---- (5) active samples, four per round (one per 8-lane group) ----
        unsigned todo = act_m;
        while (todo) {
            int src = -1;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int b = todo ? __ffs(todo) - 1 : -1;
                src = (grp == k) ? b : src;
                todo &= todo - 1;  // 0 stays 0
            }
            const bool has = src >= 0;
            const int s = has ? src : 0;
            int32_t L[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) L[q] = __shfl_sync(FULL, lk[q], s);
            const float px = __shfl_sync(FULL, pos[0], s);
            const float py = __shfl_sync(FULL, pos[1], s);
            const float pz = __shfl_sync(FULL, pos[2], s);
            const float w = __shfl_sync(FULL, weight, s);
            float wq[8];
            {
                const float ax = 1.f - px, ay = 1.f - py, az = 1.f - pz;
                const float yz00 = ay * az, yz01 = ay * pz, yz10 = py * az, yz11 = py * pz;
                wq[0] = yz00 * ax; wq[1] = yz01 * ax; wq[2] = yz10 * ax; wq[3] = yz11 * ax;
                wq[4] = yz00 * px; wq[5] = yz01 * px; wq[6] = yz10 * px; wq[7] = yz11 * px;
            }
            // ordinal of this group's sample among the ray's active samples: slot in the
            // forward->backward colour cache of the fused kernel
            const int ord = n_seen + __popc(act_m & ((1u << s) - 1u));
            const bool cached = BWD && has && ccache != nullptr && ord < ccap;
            float col[3] = {0.f, 0.f, 0.f};
            if (!BWD || __any_sync(FULL, has && !cached)) {
                // Eight UNCONDITIONAL, independent 128-bit loads (all in flight together): an empty
                // corner (or a group without a sample) reads row 0 and gets weight 0 instead of
                // being predicated off, which would serialise the loads through a shared temp.
                float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
                const bool ld = has && !cached;
                const float* shq = g.sh + (lc.quad_on ? 4 * lc.qi : 0);
                float4 v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const unsigned row = (ld && L[q] >= 0) ? (unsigned)L[q] : 0u;
                    v[q] = load_quad<VEC>(shq + (size_t)row * (unsigned)g.sh_stride, lc.qi, D);
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (ld && L[q] >= 0) {  // predicated FMAs: a NaN in row 0 can never leak in
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
            if (!BWD) {
                if (has) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) po.out3[c] = fmaf(w, fmaxf(col[c], 0.f), po.out3[c]);
                    if (ccache != nullptr && lc.qi == 0 && ord < ccap)
                        ccache[ord] = make_float4(col[0], col[1], col[2], 0.f);
                }
            } else {
                if (cached) {  // written by this warp's forward pass (ordered by __syncwarp)
                    const float4 cc = __ldcg(ccache + ord);
                    col[0] = cc.x;
                    col[1] = cc.y;
                    col[2] = cc.z;
                }
                const float la = __shfl_sync(FULL, my_after, s);
                const float sg = __shfl_sync(FULL, sigma, s);
                const float wss = __shfl_sync(FULL, ws, s);
                float gc[3];
                float tc[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float cl = fmaxf(col[c], 0.f);
                    tc[c] = cl * gout[c];
                    gc[c] = (cl == col[c]) ? w * gout[c] : 0.f;  // colour not clamped: cuvol.cu:318-327
                }
                float total = tc[0];
                total += tc[2];
                total += tc[1];
                const float wt = has ? w * total : 0.f;
                float my_accum = 0.f;
#pragma unroll
                for (int k = 0; k < 4; ++k) {  // accum -= weight * total_color, in sample order
                    accum -= __shfl_sync(FULL, wt, 8 * k);
                    my_accum = (grp == k) ? accum : my_accum;
                }
                if (has) {
                    float cgs = wss * (total * fexp(la) - my_accum);
                    if (sparsity_loss > 0.f) cgs += sparsity_loss * (4 * sg / (1 + 2 * (sg * sg)));
                    if (lc.quad_on) {
                        float ge[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            ge[e] = fmaf(lc.Yc[2][e], gc[2], fmaf(lc.Yc[1][e], gc[1], lc.Yc[0][e] * gc[0]));
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            if (L[q] >= 0)
                                red_quad<VEC>(gr.grad_sh + (int64_t)L[q] * gsh_stride + 4 * lc.qi, lc.qi, D,
                                              wq[q] * ge[0], wq[q] * ge[1], wq[q] * ge[2], wq[q] * ge[3]);
                    }
                    // density + mask: lane i of the group owns corner i
                    int32_t Lown = L[0];
                    float wown = wq[0];
#pragma unroll
                    for (int q = 1; q < 8; ++q) {
                        Lown = (lc.qi == q) ? L[q] : Lown;
                        wown = (lc.qi == q) ? wq[q] : wown;
                    }
                    if (Lown >= 0) {
                        atomicAdd(gr.grad_density + (int64_t)Lown * gd_stride, wown * cgs);
                        if (gr.mask) gr.mask[Lown] = 1;
                    }
                }
            }
        }
        n_seen += __popc(act_m);
        t0 = t_next;
    }
    po.logT = lt;
}

template <int MODE, bool VEC, bool FROM_CAM, bool COUNT>
__global__ void __launch_bounds__(LB_THREADS, MODE == 0 ? LB_MIN_BLOCKS_FWD : LB_MIN_BLOCKS)
lattice_march_kernel(GridView g, Opts opt, MarchArgs a, GradView gr) {
    int64_t ray_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;

    float ow[3], dw[3];
    if (FROM_CAM) {
#if LB_IMAGE_TILES
        // warp -> pixel in 64x64 super-tiles of 8x8 tiles: the ~4700 warps resident on the GPU at
        // any time cover a compact patch of the image (their rays cross the same voxels -> L2/L1
        // reuse) instead of a 6-pixel-high strip
        const int W = a.cam.width, rows = (int)(a.n_rays / W);
        const int supers_x = (W + 63) >> 6;
        const int64_t s = ray_id >> 12;
        const int r = (int)(ray_id & 4095);
        const int ix = (int)(s % supers_x) * 64 + ((r >> 6) & 7) * 8 + (r & 7);
        const int iyl = (int)(s / supers_x) * 64 + (r >> 9) * 8 + ((r >> 3) & 7);
        if (ix >= W || iyl >= rows) return;
        ray_id = (int64_t)iyl * W + ix;
        cam2world_ray<false>(ix, iyl + a.row_begin, a.cam, dw, ow);
#else
        if (ray_id >= a.n_rays) return;
        const int ix = (int)(ray_id % a.cam.width);
        const int iy = (int)(ray_id / a.cam.width) + a.row_begin;
        cam2world_ray<false>(ix, iy, a.cam, dw, ow);
#endif
    } else {
        if (ray_id >= a.n_rays) return;
        if (a.run_if != nullptr && *a.run_if == 0) return;  // fallback launch that is not needed
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            ow[i] = __ldg(a.origins + ray_id * 3 + i);
            dw[i] = __ldg(a.dirs + ray_id * 3 + i);
        }
    }
    LaneConst lc;
    lane_const_init(lc, lane, g.basis_dim, g.sh_dim, dw);  // raw input direction, cuvol.cu:632-635
    Ray ray;
    ray_setup(ray, ow, dw, g, opt);
    const bool hit = ray.tmin <= ray.tmax;
    const bool has_bg = g.bg_nlayers > 0;
    unsigned cnt[3] = {0u, 0u, 0u};
    float dummy_accum = 0.f;
    float rgb[3] = {0.f, 0.f, 0.f};
    float logT_fwd = 0.f;

    if (MODE != MODE_BWD) {
        PassOut po;
        po.out3[0] = po.out3[1] = po.out3[2] = 0.f;
        po.logT = 0.f;
        po.stopped = false;
        float4* ccache = nullptr;
        if (MODE == MODE_FUSED && a.ccache != nullptr && a.ccap > 0)
            ccache = a.ccache + ray_id * (int64_t)a.ccap;
        if (hit)
            march_pass<false, VEC, COUNT>(g, opt, ray, lc, lane, po, nullptr, dummy_accum, 0.f, gr, 0, 0, cnt,
                                          ccache, a.ccap);
#pragma unroll
        for (int c = 0; c < 3; ++c) {  // the four groups' partial colours
            float x = po.out3[c];
            x += __shfl_xor_sync(FULL, x, 8);
            x += __shfl_xor_sync(FULL, x, 16);
            rgb[c] = x;
        }
        logT_fwd = po.stopped ? -1e3f : po.logT;  // cuvol.cu:103-106
        if (hit) {
            if (!has_bg) {
                const float bgc = fexp(logT_fwd) * opt.background_brightness;
#pragma unroll
                for (int c = 0; c < 3; ++c) rgb[c] += bgc;
            }
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) rgb[c] = has_bg ? 0.f : opt.background_brightness;
            logT_fwd = 0.f;
        }
        if (lane < 3) a.rgb_out[ray_id * 3 + lane] = lane == 0 ? rgb[0] : (lane == 1 ? rgb[1] : rgb[2]);
        if (lane == 0) {
            if (a.logT_out && MODE == MODE_FWD) a.logT_out[ray_id] = logT_fwd;
            if (COUNT) {
                atomicAdd(a.counters + 0, (unsigned long long)cnt[0]);
                atomicAdd(a.counters + 1, (unsigned long long)cnt[1]);
                atomicAdd(a.counters + 2, (unsigned long long)cnt[2]);
            }
        }
    }

    if (MODE != MODE_FWD) {
        float gout[3], cc[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            cc[i] = (MODE == MODE_FUSED) ? rgb[i] : __ldg(a.color_cache + ray_id * 3 + i);
            const float gi = __ldg(a.grad_in + ray_id * 3 + i);
            gout[i] = a.fused_seed ? (cc[i] - gi) * a.norm_factor : gi;  // cuvol.cu:735-747
        }
        float accum = fmaf(cc[0], gout[0], fmaf(cc[1], gout[1], cc[2] * gout[2]));
        float beta_loss = a.beta_loss;
        float logT_bwd = 0.f;
        if (hit) {
            if (beta_loss > 0.f) {  // cuvol.cu:259-264
                const float lin = (MODE == MODE_FUSED) ? logT_fwd : (a.logT_in ? a.logT_in[ray_id] : 0.f);
                const float transmit_in = fexp(lin);
                beta_loss = (float)((double)beta_loss *
                                    (1 - (double)transmit_in / (1 - (double)transmit_in + 1e-3)));
                accum += beta_loss;
            }
            PassOut po;
            float4* ccache = nullptr;
            if (MODE == MODE_FUSED && a.ccache != nullptr && a.ccap > 0) {
                ccache = a.ccache + ray_id * (int64_t)a.ccap;
                __syncwarp();  // forward-pass colour cache stores -> backward-pass loads
            }
            march_pass<true, VEC, false>(g, opt, ray, lc, lane, po, gout, accum, a.sparsity_loss, gr,
                                         a.gsh_stride, a.gd_stride, cnt, ccache, a.ccap);
            logT_bwd = po.logT;
            accum -= beta_loss;  // cuvol.cu:375-377
        }
        if (lane == 0) {
            if (a.accum_out) a.accum_out[ray_id] = accum;
            if (a.logT_out && MODE == MODE_BWD) a.logT_out[ray_id] = logT_bwd;
        }
    }
}

template <int MODE, bool VEC, bool FROM_CAM>
void launch_mode(const GridView& g, const Opts& opt, const MarchArgs& a, const GradView& gr,
                 cudaStream_t st) {
    int64_t warps = a.n_rays;
#if LB_IMAGE_TILES
    if (FROM_CAM) {  // padded to whole 64x64 super-tiles; out-of-image warps exit at once
        const int64_t W = a.cam.width, rows = a.n_rays / W;
        warps = ((W + 63) / 64) * ((rows + 63) / 64) * 4096;
    }
#endif
    const unsigned blocks = div_up(warps * WARP, LB_THREADS);
    if (a.counters)
        lattice_march_kernel<MODE, VEC, FROM_CAM, true><<<blocks, LB_THREADS, 0, st>>>(g, opt, a, gr);
    else
        lattice_march_kernel<MODE, VEC, FROM_CAM, false><<<blocks, LB_THREADS, 0, st>>>(g, opt, a, gr);
}

}  // namespace

// 128-bit path needs 16-byte aligned rows for data and (when scattering) gradients.
bool march_vec_ok(const GridView& g, const GradView* gr, int64_t gsh_stride) {
    const int need = ((g.sh_dim + 3) / 4) * 4;
    bool ok = (g.sh_stride % 4 == 0) && g.sh_stride >= need &&
              (reinterpret_cast<uintptr_t>(g.sh) % 16 == 0);
    if (gr && gr->grad_sh)
        ok = ok && (gsh_stride % 4 == 0) && gsh_stride >= need &&
             (reinterpret_cast<uintptr_t>(gr->grad_sh) % 16 == 0);
    return ok;
}

void launch_lattice_march(int mode, bool from_cam, const GridView& g, const Opts& opt,
                          const MarchArgs& a, const GradView& gr, cudaStream_t st) {
    const bool vec = march_vec_ok(g, mode == MODE_FWD ? nullptr : &gr, a.gsh_stride);
#define LB_DISPATCH(M)                                                                 \
    do {                                                                               \
        if (vec) {                                                                     \
            if (from_cam) launch_mode<M, true, true>(g, opt, a, gr, st);               \
            else launch_mode<M, true, false>(g, opt, a, gr, st);                       \
        } else {                                                                       \
            if (from_cam) launch_mode<M, false, true>(g, opt, a, gr, st);              \
            else launch_mode<M, false, false>(g, opt, a, gr, st);                      \
        }                                                                              \
    } while (0)
#define LB_DISPATCH_RAYS(M)                                                            \
    do {                                                                               \
        if (vec) launch_mode<M, true, false>(g, opt, a, gr, st);                       \
        else launch_mode<M, false, false>(g, opt, a, gr, st);                          \
    } while (0)
    if (mode == MODE_FWD) LB_DISPATCH(MODE_FWD);
    else if (mode == MODE_BWD) LB_DISPATCH_RAYS(MODE_BWD);
    else LB_DISPATCH_RAYS(MODE_FUSED);
#undef LB_DISPATCH
#undef LB_DISPATCH_RAYS
}

}  // namespace svox2b
