// Grid maintenance and point sampling: dilate, accel_dist_prop, grid_weight_render,
// sample_grid(+backward).  Behaviour follows the reference's misc_kernel.cu and
// svox2_kernel.cu (citations per function); all of it is off the per-iteration path
// (resample / sparsify / load).
#include "common.cuh"

namespace svox2b {
namespace {

constexpr int MISC_THREADS = 256;

// ---------------------------------------------------------------------------------------
// dilate: 3x3x3 OR with clamped borders (misc_kernel.cu:20-48)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MISC_THREADS)
dilate_kernel(const uint8_t* __restrict__ grid, int X, int Y, int Z, uint8_t* __restrict__ out) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (int64_t)X * Y * Z) return;
    const int z = (int)(tid % Z);
    const int64_t xy = tid / Z;
    const int y = (int)(xy % Y);
    const int x = (int)(xy / Y);
    const int xs[3] = {max(x - 1, 0), x, min(x + 1, X - 1)};
    const int ys[3] = {max(y - 1, 0), y, min(y + 1, Y - 1)};
    const int zs[3] = {max(z - 1, 0), z, min(z + 1, Z - 1)};
    uint8_t v = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const uint8_t* row = grid + ((int64_t)xs[i] * Y + ys[j]) * Z;
            v |= __ldg(row + zs[0]) | __ldg(row + zs[1]) | __ldg(row + zs[2]);
        }
    out[tid] = v ? 1 : 0;
}

// ---------------------------------------------------------------------------------------
// accel_dist_prop (misc_kernel.cu:109-181): occupancy pyramid, then
// links[v] = -1 - (number of consecutive empty ancestors) for empty voxels.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int div2_ceil(int x) { return ((x - 1) >> 1) + 1; }

__global__ void __launch_bounds__(MISC_THREADS)
accel_set_kernel(const int32_t* __restrict__ links, int X, int Y, int Z, uint8_t* __restrict__ tmp) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (int64_t)X * Y * Z) return;
    if (__ldg(links + tid) < 0) return;
    int z = (int)(tid % Z);
    const int64_t xy = tid / Z;
    int y = (int)(xy % Y);
    int x = (int)(xy / Y);
    int sx = X, sy = Y, sz = Z;
    uint8_t* base = tmp;
    while (sx > 1 && sy > 1 && sz > 1) {
        x >>= 1; y >>= 1; z >>= 1;
        sx = div2_ceil(sx); sy = div2_ceil(sy); sz = div2_ceil(sz);
        base[(int64_t)x * sy * sz + (int64_t)y * sz + z] = 1;  // benign same-value race
        base += (int64_t)sx * sy * sz;
    }
}

__global__ void __launch_bounds__(MISC_THREADS)
accel_prop_kernel(int32_t* __restrict__ links, int X, int Y, int Z, const uint8_t* __restrict__ tmp) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (int64_t)X * Y * Z) return;
    if (links[tid] >= 0) return;
    int z = (int)(tid % Z);
    const int64_t xy = tid / Z;
    int y = (int)(xy % Y);
    int x = (int)(xy / Y);
    int sx = X, sy = Y, sz = Z;
    const uint8_t* base = tmp;
    int result = -1;
    while (sx > 1 && sy > 1 && sz > 1) {
        x >>= 1; y >>= 1; z >>= 1;
        sx = div2_ceil(sx); sy = div2_ceil(sy); sz = div2_ceil(sz);
        if (__ldg(base + (int64_t)x * sy * sz + (int64_t)y * sz + z)) break;
        result -= 1;
        base += (int64_t)sx * sy * sz;
    }
    links[tid] = result;
}

// ---------------------------------------------------------------------------------------
// grid_weight_render (misc_kernel.cu:184-334): march a dense sigma grid per camera pixel and
// max the sample weight into the 8 corner cells.  Weights are >= 0, so the max is taken with
// one native integer atomicMax on the float's bit pattern (the reference spins on a CAS loop);
// max is order independent, so the result is bit-identical.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_max_nonneg(float* p, float v) {
    if (v >= 0.f) {
        atomicMax(reinterpret_cast<int*>(p), __float_as_int(v));
    } else {  // not reachable for sigma > 0; kept for exactness with the reference's fmaxf
        unsigned* pu = reinterpret_cast<unsigned*>(p);
        unsigned old = *pu, assumed;
        do {
            assumed = old;
            old = atomicCAS(pu, assumed, __float_as_uint(fmaxf(v, __uint_as_float(assumed))));
        } while (old != assumed);
    }
}

__global__ void __launch_bounds__(MISC_THREADS)
grid_weight_render_kernel(const float* __restrict__ data, int X, int Y, int Z, CamView cam,
                          float step_size, float stop_thresh, int last_sample_opaque,
                          const float* __restrict__ offset, const float* __restrict__ scaling,
                          float* __restrict__ grid_weight) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (int64_t)cam.width * cam.height) return;
    const int iy = (int)(tid / cam.width), ix = (int)(tid % cam.width);
    float d[3], o[3];
    cam2world_ray<true>(ix, iy, cam, d, o);
    const int size[3] = {X, Y, Z};
    float invd[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float sc = __ldg(scaling + i);
        o[i] = __fmaf_rn(o[i], sc, __ldg(offset + i));
        d[i] = __fmul_rn(d[i], sc);
    }
    const float delta_scale = rnorm3df(d[0], d[1], d[2]);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        d[i] = __fmul_rn(d[i], delta_scale);
        invd[i] = __frcp_rn(d[i]);
    }
    const float world_step = __fmul_rn(delta_scale, step_size);
    float t = 0.f, tmax = 2e3f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float t1 = __fmul_rn(__fsub_rn(-0.5f, o[i]), invd[i]);
        const float t2 = __fmul_rn(__fsub_rn(__fsub_rn((float)size[i], 0.5f), o[i]), invd[i]);
        if (d[i] != 0.f) {
            t = fmaxf(t, fminf(t1, t2));
            tmax = fminf(tmax, fmaxf(t1, t2));
        }
    }
    if (t > tmax) return;
    const int64_t stride0 = (int64_t)Y * Z;
    const int stride1 = Z;
    float logI = 0.f;
    while (t <= tmax) {
        float pos[3];
        int l[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float p = __fmaf_rn(t, d[j], o[j]);  // o + t*d is contracted to FMA in the reference
            p = fminf(fmaxf(p, 0.f), (float)size[j] - 1.f);
            l[j] = min((int)p, size[j] - 2);
            pos[j] = __fsub_rn(p, (float)l[j]);
        }
        const int64_t idx = l[0] * stride0 + (int64_t)l[1] * stride1 + l[2];
        const float* s000 = data + idx;
        const float* s100 = s000 + stride0;
        float v[8] = {__ldg(s000), __ldg(s000 + 1), __ldg(s000 + stride1), __ldg(s000 + stride1 + 1),
                      __ldg(s100), __ldg(s100 + 1), __ldg(s100 + stride1), __ldg(s100 + stride1 + 1)};
        // misc_kernel.cu:236-244: z lerps, then y, then x
        const float ix0y0 = lerp(v[0], v[1], pos[2]);
        const float ix0y1 = lerp(v[2], v[3], pos[2]);
        const float ix1y0 = lerp(v[4], v[5], pos[2]);
        const float ix1y1 = lerp(v[6], v[7], pos[2]);
        const float ix0 = lerp(ix0y0, ix0y1, pos[1]);
        const float ix1 = lerp(ix1y0, ix1y1, pos[1]);
        float sigma = lerp(ix0, ix1, pos[0]);
        if (last_sample_opaque && __fadd_rn(t, step_size) > tmax) {
            sigma += 1e9f;
            logI = 0.f;
        }
        if (sigma > 1e-8f) {
            const float log_att = __fmul_rn(-world_step, sigma);
            const float weight = __fmul_rn(fexp(logI), __fsub_rn(1.f, fexp(log_att)));
            logI = __fadd_rn(logI, log_att);
            float* w000 = grid_weight + idx;
            float* w100 = w000 + stride0;
            atomic_max_nonneg(w000, weight);
            atomic_max_nonneg(w000 + 1, weight);
            atomic_max_nonneg(w000 + stride1, weight);
            atomic_max_nonneg(w000 + stride1 + 1, weight);
            atomic_max_nonneg(w100, weight);
            atomic_max_nonneg(w100 + 1, weight);
            atomic_max_nonneg(w100 + stride1, weight);
            atomic_max_nonneg(w100 + stride1 + 1, weight);
            if (fexp(logI) < stop_thresh) break;
        }
        t = __fadd_rn(t, step_size);
    }
}

// ---------------------------------------------------------------------------------------
// sample_grid (svox2_kernel.cu:10-78) and its backward (:81-176)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void sample_locate(const GridView& g, const float* __restrict__ points,
                                              int64_t pid, int* l, float* pos) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float p = __fmaf_rn(__ldg(points + pid * 3 + i), g.scaling[i], g.offset[i]);
        p = fminf(fmaxf(p, 0.f), (float)g.size[i] - 1.f);
        l[i] = min((int)p, g.size[i] - 2);
        pos[i] = __fsub_rn(p, (float)l[i]);
    }
}

__global__ void __launch_bounds__(MISC_THREADS)
sample_density_kernel(GridView g, const float* __restrict__ points, int64_t P,
                      float* __restrict__ out) {
    const int64_t pid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pid >= P) return;
    int l[3];
    float pos[3];
    sample_locate(g, points, pid, l, pos);
    const int32_t* lp = g.links + ((int64_t)l[0] * g.offx + (int64_t)l[1] * g.offy + l[2]);
    float v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int32_t lk = __ldg(lp + corner_off(g, c));
        v[c] = lk >= 0 ? __ldg(g.density + (int64_t)lk * g.density_stride) : 0.f;
    }
    out[pid] = trilerp8(v, pos);
}

__global__ void __launch_bounds__(MISC_THREADS)
sample_sh_kernel(GridView g, const float* __restrict__ points, int64_t P, float* __restrict__ out) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= P * g.sh_dim) return;
    const int idx = (int)(tid % g.sh_dim);
    const int64_t pid = tid / g.sh_dim;
    int l[3];
    float pos[3];
    sample_locate(g, points, pid, l, pos);
    const int32_t* lp = g.links + ((int64_t)l[0] * g.offx + (int64_t)l[1] * g.offy + l[2]);
    float v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int32_t lk = __ldg(lp + corner_off(g, c));
        v[c] = lk >= 0 ? __ldg(g.sh + (int64_t)lk * g.sh_stride + idx) : 0.f;
    }
    out[tid] = trilerp8(v, pos);
}

__global__ void __launch_bounds__(MISC_THREADS)
sample_backward_kernel(GridView g, const float* __restrict__ points, int64_t P,
                       const float* __restrict__ grad_out, int channels, int64_t out_stride,
                       float* __restrict__ grad_data) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= P * channels) return;
    const int idx = (int)(tid % channels);
    const int64_t pid = tid / channels;
    int l[3];
    float pos[3];
    sample_locate(g, points, pid, l, pos);
    const int32_t* lp = g.links + ((int64_t)l[0] * g.offx + (int64_t)l[1] * g.offy + l[2]);
    const float go = __ldg(grad_out + tid);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int32_t lk = __ldg(lp + corner_off(g, c));
        if (lk >= 0) {
            const float wx = (c & 4) ? pos[0] : 1.f - pos[0];
            const float wy = (c & 2) ? pos[1] : 1.f - pos[1];
            const float wz = (c & 1) ? pos[2] : 1.f - pos[2];
            atomicAdd(grad_data + (int64_t)lk * out_stride + idx, (wy * wz) * (wx * go));
        }
    }
}

}  // namespace
}  // namespace svox2b

using namespace svox2b;

extern "C" SVOX2_API int svox2_dilate(const uint8_t* grid, int32_t X, int32_t Y, int32_t Z, uint8_t* out,
                            svox2_stream_t stream) {
    SVOX2_REQUIRE(grid && out, "grid/out must not be null");
    SVOX2_REQUIRE(X > 0 && Y > 0 && Z > 0, "grid must be 3-D and non-empty");
    const int64_t n = (int64_t)X * Y * Z;
    dilate_kernel<<<div_up(n, MISC_THREADS), MISC_THREADS, 0, (cudaStream_t)stream>>>(grid, X, Y, Z, out);
    return check_launch(__func__);
}

extern "C" SVOX2_API size_t svox2_accel_workspace_bytes(int32_t X, int32_t Y, int32_t Z) {
    int64_t sx = X, sy = Y, sz = Z, req = 0;
    while (sx > 1 && sy > 1 && sz > 1) {
        sx = ((sx - 1) >> 1) + 1;
        sy = ((sy - 1) >> 1) + 1;
        sz = ((sz - 1) >> 1) + 1;
        req += sx * sy * sz;
    }
    return (size_t)(req > 0 ? req : 1);
}

extern "C" SVOX2_API int svox2_accel_dist_prop(int32_t* links, int32_t X, int32_t Y, int32_t Z,
                                     void* workspace, size_t workspace_bytes,
                                     svox2_stream_t stream) {
    SVOX2_REQUIRE(links, "links must not be null");
    SVOX2_REQUIRE(X > 0 && Y > 0 && Z > 0, "links must be 3-D and non-empty");
    const size_t need = svox2_accel_workspace_bytes(X, Y, Z);
    if (!workspace || workspace_bytes < need) {
        set_error("%s: workspace too small (%zu < %zu bytes)", __func__, workspace_bytes, need);
        return SVOX2_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(workspace, 0, need, st) != cudaSuccess) return check_launch(__func__);
    const int64_t n = (int64_t)X * Y * Z;
    accel_set_kernel<<<div_up(n, MISC_THREADS), MISC_THREADS, 0, st>>>(
        links, X, Y, Z, reinterpret_cast<uint8_t*>(workspace));
    accel_prop_kernel<<<div_up(n, MISC_THREADS), MISC_THREADS, 0, st>>>(
        links, X, Y, Z, reinterpret_cast<const uint8_t*>(workspace));
    return check_launch(__func__);
}

extern "C" SVOX2_API int svox2_grid_weight_render(const float* data, int32_t X, int32_t Y, int32_t Z,
                                        const svox2_camera_t* cam, float step_size,
                                        float stop_thresh, int32_t last_sample_opaque,
                                        const float* offset, const float* scaling,
                                        float* grid_weight_out, svox2_stream_t stream) {
    SVOX2_REQUIRE(data && cam && cam->c2w && offset && scaling && grid_weight_out,
                  "null argument");
    SVOX2_REQUIRE(X > 1 && Y > 1 && Z > 1, "data must be 3-D with every dimension >= 2");
    SVOX2_REQUIRE(cam->width > 0 && cam->height > 0, "camera size must be positive");
    CamView cv{cam->c2w, cam->fx, cam->fy, cam->cx, cam->cy, cam->width, cam->height,
               cam->ndc_coeffx, cam->ndc_coeffy};
    const int64_t n = (int64_t)cam->width * cam->height;
    grid_weight_render_kernel<<<div_up(n, MISC_THREADS), MISC_THREADS, 0, (cudaStream_t)stream>>>(
        data, X, Y, Z, cv, step_size, stop_thresh, last_sample_opaque, offset, scaling,
        grid_weight_out);
    return check_launch(__func__);
}

extern "C" SVOX2_API int svox2_sample_grid(const svox2_grid_t* grid, const float* points, int64_t n_points,
                                 float* out_density, float* out_sh, svox2_stream_t stream) {
    if (int e = validate_grid(grid, __func__)) return e;
    SVOX2_REQUIRE(n_points >= 0 && (n_points == 0 || (points && out_density)), "null argument");
    if (n_points == 0) return SVOX2_OK;
    const GridView g = make_grid_view(grid);
    cudaStream_t st = (cudaStream_t)stream;
    sample_density_kernel<<<div_up(n_points, MISC_THREADS), MISC_THREADS, 0, st>>>(g, points, n_points,
                                                                                 out_density);
    if (out_sh)
        sample_sh_kernel<<<div_up(n_points * g.sh_dim, MISC_THREADS), MISC_THREADS, 0, st>>>(
            g, points, n_points, out_sh);
    return check_launch(__func__);
}

extern "C" SVOX2_API int svox2_sample_grid_backward(const svox2_grid_t* grid, const float* points,
                                          int64_t n_points, const float* grad_out_density,
                                          const float* grad_out_sh, float* grad_density_out,
                                          int64_t grad_density_stride, float* grad_sh_out,
                                          int64_t grad_sh_stride, svox2_stream_t stream) {
    if (int e = validate_grid(grid, __func__)) return e;
    SVOX2_REQUIRE(n_points >= 0 && (n_points == 0 || (points && grad_out_density && grad_density_out)),
                  "null argument");
    if (n_points == 0) return SVOX2_OK;
    const GridView g = make_grid_view(grid);
    cudaStream_t st = (cudaStream_t)stream;
    sample_backward_kernel<<<div_up(n_points, MISC_THREADS), MISC_THREADS, 0, st>>>(
        g, points, n_points, grad_out_density, 1,
        grad_density_stride > 0 ? grad_density_stride : g.density_stride, grad_density_out);
    if (grad_out_sh && grad_sh_out)
        sample_backward_kernel<<<div_up(n_points * g.sh_dim, MISC_THREADS), MISC_THREADS, 0, st>>>(
            g, points, n_points, grad_out_sh, g.sh_dim,
            grad_sh_stride > 0 ? grad_sh_stride : g.sh_stride, grad_sh_out);
    return check_launch(__func__);
}
