// Shared device-side math for the B200-native Plenoxels kernels.
//
// Everything that decides WHERE a sample falls (the t lattice, cell index, skip distance,
// sigma trilerp, log-transmittance chain) is written with explicitly rounded intrinsics
// (__fmul_rn/__fadd_rn/__fmaf_rn/__fdiv_rn) so that nvcc can neither contract nor reorder it:
// links indexing and the sparse gradient mask must be bit-exact with the reference
// (SURVEY.md hard part 1).  Behaviour follows svox2/csrc/include/render_util.cuh and
// render_lerp_kernel_cuvol.cu of the reference; see the per-function citations.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/svox2_b200.h"

namespace svox2b {

constexpr int WARP = 32;
constexpr unsigned FULL = 0xffffffffu;

// ---------------------------------------------------------------------------------------
// error plumbing (host)
// ---------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define SVOX2_REQUIRE(cond, msg)                                     \
    do {                                                             \
        if (!(cond)) {                                               \
            svox2b::set_error("%s: %s", __func__, msg);              \
            return SVOX2_ERR_INVALID;                                \
        }                                                            \
    } while (0)

inline unsigned div_up(int64_t a, int64_t b) { return (unsigned)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------
// POD views passed to kernels by value
// ---------------------------------------------------------------------------------------
struct GridView {
    const float* __restrict__ density;
    const float* __restrict__ sh;
    const int32_t* __restrict__ links;
    int size[3];
    int basis_dim;
    int sh_dim;  // 3 * basis_dim
    int sh_stride;      // floats per row (32-bit so row*stride is one IMAD.WIDE)
    int density_stride;
    int64_t offx;  // Y*Z
    int offy;      // Z
    float offset[3];
    float scaling[3];
    const int32_t* __restrict__ bg_links;
    const float* __restrict__ bg_data;
    int bg_reso;
    int bg_nlayers;
};

struct GradView {
    float* __restrict__ grad_density;
    float* __restrict__ grad_sh;
    float* __restrict__ grad_bg;
    uint8_t* __restrict__ mask;
    uint8_t* __restrict__ mask_bg;
};

struct CamView {
    const float* __restrict__ c2w;
    float fx, fy, cx, cy;
    int width, height;
    float ndc_coeffx, ndc_coeffy;
};

typedef svox2_render_opts_t Opts;

inline GridView make_grid_view(const svox2_grid_t* g) {
    GridView v;
    v.density = g->density_data;
    v.sh = g->sh_data;
    v.links = g->links;
    for (int i = 0; i < 3; ++i) {
        v.size[i] = g->size[i];
        v.offset[i] = g->offset[i];
        v.scaling[i] = g->scaling[i];
    }
    v.basis_dim = g->basis_dim;
    v.sh_dim = 3 * g->basis_dim;
    v.sh_stride = (int)g->sh_stride;
    v.density_stride = (int)g->density_stride;
    v.offx = (int64_t)g->size[1] * g->size[2];
    v.offy = g->size[2];
    const bool bg = g->background_links && g->background_data && g->background_nlayers > 0 &&
                    g->background_rows > 0;
    v.bg_links = bg ? g->background_links : nullptr;
    v.bg_data = bg ? g->background_data : nullptr;
    v.bg_reso = bg ? g->background_reso : 0;
    v.bg_nlayers = bg ? g->background_nlayers : 0;
    return v;
}

int validate_grid(const svox2_grid_t* g, const char* who);

// ---------------------------------------------------------------------------------------
// device math
// ---------------------------------------------------------------------------------------
#ifdef __CUDACC__

// cuda_util.cuh:72-75
__device__ __forceinline__ float lerp(float a, float b, float w) {
    return __fmaf_rn(w, __fsub_rn(b, a), a);
}

// util.hpp:14-17: _EXP == __expf (ex2.approx of x*log2e, non-ftz fix-up emitted by nvcc)
__device__ __forceinline__ float fexp(float x) { return __expf(x); }

// Zero that depends on a loaded value: an INDEPENDENT store to an address whose load is still in
// flight stalls the memory pipeline badly (measured in round 1; again in round 2 on
// the MSI optimiser: 0.39 -> 1.3 ms when the compiler hoisted the constant-zero gradient stores).
__device__ __forceinline__ float dependent_zero(float x) {
    int z;
    asm volatile("and.b32 %0, %1, 0;" : "=r"(z) : "r"(__float_as_int(x)));
    return __int_as_float(z);
}

__constant__ const float kC0 = 0.28209479177387814f;

// calc_sh, render_util.cuh:332-363.  out[k] for k < basis_dim.
__device__ __forceinline__ void calc_sh(int basis_dim, const float* dir, float* out) {
    const float C1 = 0.4886025119029199f;
    const float C2_0 = 1.0925484305920792f, C2_1 = -1.0925484305920792f,
                C2_2 = 0.31539156525252005f, C2_3 = -1.0925484305920792f,
                C2_4 = 0.5462742152960396f;
    out[0] = 0.28209479177387814f;
    const float x = dir[0], y = dir[1], z = dir[2];
    const float xx = x * x, yy = y * y, zz = z * z;
    const float xy = x * y, yz = y * z, xz = x * z;
    if (basis_dim >= 9) {
        out[4] = C2_0 * xy;
        out[5] = C2_1 * yz;
        out[6] = (float)((double)C2_2 * (2.0 * (double)zz - (double)xx - (double)yy));
        out[7] = C2_3 * xz;
        out[8] = C2_4 * (xx - yy);
    }
    if (basis_dim >= 4) {
        out[1] = -C1 * y;
        out[2] = C1 * z;
        out[3] = -C1 * x;
    }
}

// One SH basis value for index k (lane-private register instead of a shared table).
__device__ __forceinline__ float sh_basis_k(int basis_dim, int k, const float* dir) {
    float tmp[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) tmp[i] = 0.f;
    calc_sh(basis_dim, dir, tmp);
    float r = tmp[0];
#pragma unroll
    for (int i = 1; i < 9; ++i) r = (k == i) ? tmp[i] : r;
    return r;
}

struct Ray {
    float o[3];     // origin, grid coords
    float d[3];     // unit direction, grid coords
    float invd[3];  // IEEE 1/d (the reference's (float)(1.0/d), render_util.cuh:258,612)
    float tmin, tmax, world_step;
};

// ConcentricSpheresIntersector, render_util.cuh:546-576
struct SphereIsect {
    float q2a, qb, f;
    __device__ __forceinline__ SphereIsect(const float* o, const float* d) {
        q2a = 2 * (d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        qb = 2 * (o[0] * d[0] + o[1] * d[1] + o[2] * d[2]);
        f = qb * qb - 2 * q2a * (o[0] * o[0] + o[1] * o[1] + o[2] * o[2]);
    }
    __device__ __forceinline__ bool intersect(float r, float* out, bool near = false) const {
        const float det = f + 2 * q2a * r * r;
        if (det < 0) return false;
        *out = near ? (-qb - sqrtf(det)) / q2a : (-qb + sqrtf(det)) / q2a;
        return true;
    }
};

// ray_find_bounds, render_util.cuh:578-628
__device__ __forceinline__ void ray_setup(Ray& r, const float* ow, const float* dw,
                                          const GridView& g, const Opts& opt) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        r.o[i] = __fmaf_rn(ow[i], g.scaling[i], g.offset[i]);
        r.d[i] = __fmul_rn(dw[i], g.scaling[i]);
    }
    const float delta_scale = rnorm3df(r.d[0], r.d[1], r.d[2]);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        r.d[i] = __fmul_rn(r.d[i], delta_scale);
        r.invd[i] = __frcp_rn(r.d[i]);
    }
    r.world_step = __fmul_rn(delta_scale, opt.step_size);

    if (opt.use_spheric_clip) {
        float so[3], sd[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float ss = __fdiv_rn(2.f, (float)g.size[i]);
            so[i] = __fmaf_rn(__fadd_rn(r.o[i], 0.5f), ss, -1.f);
            sd[i] = __fmul_rn(r.d[i], ss);
        }
        SphereIsect csi(so, sd);
        if (!csi.intersect(1.f, &r.tmax) || !csi.intersect(1.f - opt.near_clip, &r.tmin, true)) {
            r.tmin = 1e-9f;
            r.tmax = 0.f;
        }
    } else {
        r.tmin = __fmul_rn(__fdiv_rn(opt.near_clip, r.world_step), opt.step_size);
        r.tmax = 2e3f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float t1 = __fmul_rn(__fsub_rn(-0.5f, r.o[i]), r.invd[i]);
            const float t2 =
                __fmul_rn(__fsub_rn(__fsub_rn((float)g.size[i], 0.5f), r.o[i]), r.invd[i]);
            if (r.d[i] != 0.f) {
                r.tmin = fmaxf(r.tmin, fminf(t1, t2));
                r.tmax = fminf(r.tmax, fmaxf(t1, t2));
            }
        }
    }
}

// cuvol.cu:57-63: sample position -> cell index l and fractional position
__device__ __forceinline__ void locate(const Ray& r, const GridView& g, float t, int* l,
                                       float* pos) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float p = __fmaf_rn(t, r.d[j], r.o[j]);
        p = fminf(fmaxf(p, 0.f), (float)g.size[j] - 1.f);
        l[j] = min((int)p, g.size[j] - 2);
        pos[j] = __fsub_rn(p, (float)l[j]);
    }
}

// compute_skip_dist, render_util.cuh:235-295, given the already loaded link of cell l.
__device__ __forceinline__ float skip_dist(const Ray& r, int32_t link_val, const int* l,
                                           const float* pos) {
    if (link_val >= -1) return 0.f;
    const uint32_t shift = (uint32_t)(-link_val) - 1u;
    const uint32_t side = (uint32_t)((float)(1 << shift) - 1.f);
    float tmin = 0.f, tmax = 1e9f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        int ul = ((l[i] >> shift) << shift) - l[i];
        const float t1 = __fmul_rn(__fsub_rn((float)ul, pos[i]), r.invd[i]);
        const float t2 =
            __fmul_rn(__fsub_rn((float)((uint32_t)ul + side), pos[i]), r.invd[i]);
        if (r.d[i] != 0.f) {
            tmin = fmaxf(tmin, fminf(t1, t2));
            tmax = fminf(tmax, fmaxf(t1, t2));
        }
    }
    if (tmin > 0.f) return 0.f;
    return tmax;
}

// t advance on a skip: t += ceilf(skip/step)*step  (cuvol.cu:69-72; FMUL + FADD in the
// reference's SASS, no contraction)
__device__ __forceinline__ float skip_advance(float t, float skip, float step) {
    return __fadd_rn(t, __fmul_rn(ceilf(__fdiv_rn(skip, step)), step));
}

// Offset (in links elements) of corner c = 4*dx + 2*dy + dz from the cell base.
__device__ __forceinline__ int64_t corner_off(const GridView& g, int c) {
    return (int64_t)(c >> 2) * g.offx + (int64_t)((c >> 1) & 1) * g.offy + (c & 1);
}

// trilerp weights of trilerp_backward_cuvol_one (render_util.cuh:94-121) for corner c.
__device__ __forceinline__ float corner_weight(const float* pos, int c) {
    const float wx = (c & 4) ? pos[0] : 1.f - pos[0];
    const float wy = (c & 2) ? pos[1] : 1.f - pos[1];
    const float wz = (c & 1) ? pos[2] : 1.f - pos[2];
    return wy * wz * wx;
}

// z, y, x lerp of 8 corner values held in registers (render_util.cuh:82-89)
__device__ __forceinline__ float trilerp8(const float* v, const float* pos) {
    const float ix0y0 = lerp(v[0], v[1], pos[2]);
    const float ix0y1 = lerp(v[2], v[3], pos[2]);
    const float ix0 = lerp(ix0y0, ix0y1, pos[1]);
    const float ix1y0 = lerp(v[4], v[5], pos[2]);
    const float ix1y1 = lerp(v[6], v[7], pos[2]);
    const float ix1 = lerp(ix1y0, ix1y1, pos[1]);
    return lerp(ix0, ix1, pos[0]);
}

// Same trilerp when corner c's value lives in lane c (c < 8) of the warp: three xor-shuffle
// rounds (z, y, x) reproduce the reference's lerp order bit for bit; the result is valid in
// lane 0 and broadcast.
__device__ __forceinline__ float trilerp_lanes(float v, const float* pos) {
    float o = __shfl_xor_sync(FULL, v, 1);
    v = lerp(v, o, pos[2]);  // valid in lanes with dz == 0
    o = __shfl_xor_sync(FULL, v, 2);
    v = lerp(v, o, pos[1]);  // valid in lanes with dz == dy == 0
    o = __shfl_xor_sync(FULL, v, 4);
    v = lerp(v, o, pos[0]);  // valid in lane 0
    return __shfl_sync(FULL, v, 0);
}

// cam2world_ray + world2ndc, render_util.cuh:504-544.
// Written with explicitly rounded operations that reproduce the reference's sm_100 SASS, so
// that generated rays are bit-identical (grid_weight_render thresholds decide links):
//   dir_i = fma(c2w[i][2], z, fma(c2w[i][0], x, c2w[i][1]*y))  in every reference kernel;
//   x*x + y*y is contracted as fma(y,y,x*x) in grid_weight_render_kernel (XX_FIRST = true)
//   and as fma(x,x,y*y) in render_ray_image_kernel / render_background_image_kernel.
template <bool XX_FIRST>
__device__ __forceinline__ void cam2world_ray(int ix, int iy, const CamView& cam, float* dir,
                                              float* origin) {
    float x = __fdiv_rn(__fsub_rn(__fadd_rn((float)ix, 0.5f), cam.cx), cam.fx);
    float y = __fdiv_rn(__fsub_rn(__fadd_rn((float)iy, 0.5f), cam.cy), cam.fy);
    const float s = XX_FIRST ? __fmaf_rn(y, y, __fmul_rn(x, x)) : __fmaf_rn(x, x, __fmul_rn(y, y));
    float z = __fsqrt_rn(__fadd_rn(s, 1.f));
    x = __fdiv_rn(x, z);
    y = __fdiv_rn(y, z);
    z = __frcp_rn(z);
    const float* m = cam.c2w;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        dir[i] = __fmaf_rn(__ldg(m + 4 * i + 2), z,
                           __fmaf_rn(__ldg(m + 4 * i), x, __fmul_rn(__ldg(m + 4 * i + 1), y)));
        origin[i] = __ldg(m + 4 * i + 3);
    }
    if (cam.ndc_coeffx > 0.f) {
        const float t = __fdiv_rn(__fsub_rn(1.f, origin[2]), dir[2]);  // near = 1
#pragma unroll
        for (int i = 0; i < 3; ++i) origin[i] = __fmaf_rn(t, dir[i], origin[i]);
        const float ox = __fdiv_rn(origin[0], origin[2]);
        const float oy = __fdiv_rn(origin[1], origin[2]);
        const float two_over_z = __fdiv_rn(2.f, origin[2]);
        dir[0] = __fmul_rn(cam.ndc_coeffx, __fsub_rn(__fdiv_rn(dir[0], dir[2]), ox));
        dir[1] = __fmul_rn(cam.ndc_coeffy, __fsub_rn(__fdiv_rn(dir[1], dir[2]), oy));
        dir[2] = two_over_z;
        origin[0] = __fmul_rn(cam.ndc_coeffx, ox);
        origin[1] = __fmul_rn(cam.ndc_coeffy, oy);
        origin[2] = __fsub_rn(1.f, two_over_z);
        const float rn = rnorm3df(dir[0], dir[1], dir[2]);
        dir[0] = __fmul_rn(dir[0], rn);
        dir[1] = __fmul_rn(dir[1], rn);
        dir[2] = __fmul_rn(dir[2], rn);
    }
}

#endif  // __CUDACC__

}  // namespace svox2b
