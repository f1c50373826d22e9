// C-ABI entry points of the renderer (include/svox2_b200.h): argument checks, workspace carving
// and kernel sequencing.  The march itself is the lattice-batch kernel of march.cu; the MSI
// background and depth kernels live in render.cu.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "march.cuh"

using namespace svox2b;

namespace {

// SVOX2_B200_MARCH=lattice forces the single-kernel fused path (A/B and debugging)
bool use_queue() {
    static const int v = [] {
        const char* e = std::getenv("SVOX2_B200_MARCH");
        return (e && std::strcmp(e, "lattice") == 0) ? 0 : 1;
    }();
    return v == 1;
}

int validate_rays(const float* origins, const float* dirs, int64_t n, const char* who) {
    if (n < 0 || (n > 0 && (!origins || !dirs))) {
        set_error("%s: rays must be non-null [Q,3] float32 device arrays", who);
        return SVOX2_ERR_INVALID;
    }
    if (n > ((int64_t)1 << 40)) {
        set_error("%s: too many rays", who);
        return SVOX2_ERR_INVALID;
    }
    return SVOX2_OK;
}

int validate_opt(const svox2_render_opts_t* opt, const char* who) {
    if (!opt) {
        set_error("%s: opt must not be null", who);
        return SVOX2_ERR_INVALID;
    }
    if (!(opt->step_size > 0.f)) {
        set_error("%s: step_size must be positive", who);
        return SVOX2_ERR_INVALID;
    }
    return SVOX2_OK;
}

int validate_grads(const svox2_grads_t* grads, const char* who) {
    if (!grads || !grads->grad_density_out || !grads->grad_sh_out) {
        set_error("%s: grad_density_out and grad_sh_out are required", who);
        return SVOX2_ERR_INVALID;
    }
    return SVOX2_OK;
}

GradView make_grad_view(const svox2_grads_t* gr) {
    GradView v;
    v.grad_density = gr->grad_density_out;
    v.grad_sh = gr->grad_sh_out;
    v.grad_bg = gr->grad_background_out;
    v.mask = gr->mask_out;
    v.mask_bg = gr->mask_background_out;
    return v;
}

struct Workspace {
    float* logT;
    float* accum;
};

int carve_workspace(void* ws, size_t bytes, int64_t n_rays, Workspace& w, const char* who) {
    const size_t need = svox2_render_workspace_bytes(n_rays);
    if (!ws || bytes < need) {
        set_error("%s: workspace too small (%zu < %zu bytes)", who, bytes, need);
        return SVOX2_ERR_WORKSPACE;
    }
    w.logT = reinterpret_cast<float*>(ws);
    w.accum = w.logT + ((n_rays + 63) / 64) * 64;
    return SVOX2_OK;
}

MarchArgs base_args(const float* origins, const float* dirs, int64_t n_rays) {
    MarchArgs a;
    std::memset(&a, 0, sizeof(a));
    a.origins = origins;
    a.dirs = dirs;
    a.n_rays = n_rays;
    return a;
}

void set_grad_strides(MarchArgs& a, const svox2_grid_t* grid, const svox2_grads_t* grads) {
    a.gsh_stride = grads->grad_sh_stride > 0 ? grads->grad_sh_stride : grid->sh_stride;
    a.gd_stride = grads->grad_density_stride > 0 ? grads->grad_density_stride : grid->density_stride;
}

void run_fwd(bool from_cam, const GridView& g, const Opts& opt, const MarchArgs& a, cudaStream_t st) {
    launch_lattice_march(LB_MODE_FWD, from_cam, g, opt, a, GradView{}, st);
}

int run_bwd(const GridView& g, const Opts& opt, const MarchArgs& a, const GradView& gr,
            cudaStream_t st, const char* who) {
    launch_lattice_march(LB_MODE_BWD, false, g, opt, a, gr, st);
    return SVOX2_OK;
}

}  // namespace

extern "C" SVOX2_API size_t svox2_render_workspace_bytes(int64_t n_rays) {
    const int64_t padded = ((n_rays + 63) / 64) * 64;
    return (size_t)padded * 2 * sizeof(float) + 256;
}

extern "C" SVOX2_API size_t svox2_fused_workspace_bytes(int64_t n_rays, int64_t slots_per_ray) {
    if (n_rays < 0) n_rays = 0;
    if (slots_per_ray < 0) slots_per_ray = 0;
    const size_t lattice = svox2_render_workspace_bytes(n_rays) + 16 * (size_t)slots_per_ray * (size_t)n_rays;
    const size_t queue = queue_workspace_bytes(n_rays, slots_per_ray * n_rays);
    return lattice > queue ? lattice : queue;
}

extern "C" SVOX2_API int svox2_volume_render_cuvol(const svox2_grid_t* grid, const float* origins,
                                                   const float* dirs, int64_t n_rays,
                                                   const svox2_render_opts_t* opt, float* out,
                                                   void* workspace, size_t workspace_bytes,
                                                   int64_t* counters, svox2_stream_t stream) {
    if (int e = validate_grid(grid, __func__)) return e;
    if (int e = validate_rays(origins, dirs, n_rays, __func__)) return e;
    if (int e = validate_opt(opt, __func__)) return e;
    SVOX2_REQUIRE(out || n_rays == 0, "out must not be null");
    if (n_rays == 0) return SVOX2_OK;
    const GridView g = make_grid_view(grid);
    const bool has_bg = g.bg_nlayers > 0;
    Workspace w{nullptr, nullptr};
    if (has_bg)
        if (int e = carve_workspace(workspace, workspace_bytes, n_rays, w, __func__)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    MarchArgs a = base_args(origins, dirs, n_rays);
    a.rgb_out = out;
    a.logT_out = w.logT;
    a.counters = reinterpret_cast<unsigned long long*>(counters);
    run_fwd(false, g, *opt, a, st);
    if (has_bg) launch_bg_fwd(false, g, *opt, a, w.logT, st);
    return check_launch(__func__);
}

extern "C" SVOX2_API int svox2_volume_render_cuvol_image(const svox2_grid_t* grid,
                                                         const svox2_camera_t* cam,
                                                         const svox2_render_opts_t* opt,
                                                         int32_t row_begin, int32_t row_end,
                                                         float* out, void* workspace,
                                                         size_t workspace_bytes,
                                                         svox2_stream_t stream) {
    if (int e = validate_grid(grid, __func__)) return e;
    if (int e = validate_opt(opt, __func__)) return e;
    SVOX2_REQUIRE(cam && cam->c2w && out, "cam/out must not be null");
    SVOX2_REQUIRE(cam->width > 0 && cam->height > 0, "camera size must be positive");
    SVOX2_REQUIRE(row_begin >= 0 && row_end <= cam->height && row_begin <= row_end,
                  "row range must lie inside the image");
    const int64_t n_rays = (int64_t)(row_end - row_begin) * cam->width;
    if (n_rays == 0) return SVOX2_OK;
    const GridView g = make_grid_view(grid);
    const bool has_bg = g.bg_nlayers > 0;
    Workspace w{nullptr, nullptr};
    if (has_bg)
        if (int e = carve_workspace(workspace, workspace_bytes, n_rays, w, __func__)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    MarchArgs a = base_args(nullptr, nullptr, n_rays);
    a.cam = CamView{cam->c2w, cam->fx, cam->fy, cam->cx, cam->cy, cam->width, cam->height,
                    cam->ndc_coeffx, cam->ndc_coeffy};
    a.row_begin = row_begin;
    a.rgb_out = out + (int64_t)row_begin * cam->width * 3;
    a.logT_out = w.logT;
    run_fwd(true, g, *opt, a, st);
    if (has_bg) launch_bg_fwd(true, g, *opt, a, w.logT, st);
    return check_launch(__func__);
}

extern "C" SVOX2_API int svox2_volume_render_cuvol_backward(
    const svox2_grid_t* grid, const float* origins, const float* dirs, int64_t n_rays,
    const svox2_render_opts_t* opt, const float* grad_out, const float* color_cache,
    const svox2_grads_t* grads, void* workspace, size_t workspace_bytes, svox2_stream_t stream) {
    if (int e = validate_grid(grid, __func__)) return e;
    if (int e = validate_rays(origins, dirs, n_rays, __func__)) return e;
    if (int e = validate_opt(opt, __func__)) return e;
    if (int e = validate_grads(grads, __func__)) return e;
    SVOX2_REQUIRE((grad_out && color_cache) || n_rays == 0, "grad_out/color_cache must not be null");
    if (n_rays == 0) return SVOX2_OK;
    const GridView g = make_grid_view(grid);
    const GradView gr = make_grad_view(grads);
    const bool use_bg = g.bg_nlayers > 0;
    Workspace w{nullptr, nullptr};
    if (use_bg)
        if (int e = carve_workspace(workspace, workspace_bytes, n_rays, w, __func__)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    MarchArgs a = base_args(origins, dirs, n_rays);
    set_grad_strides(a, grid, grads);
    a.grad_in = grad_out;
    a.color_cache = color_cache;
    a.accum_out = w.accum;   // MSI hand-over (cuvol.cu:1047-1048): backward-derived logT, quirk kept
    a.logT_out = w.logT;
    if (int e = run_bwd(g, *opt, a, gr, st, __func__)) return e;
    if (use_bg && gr.grad_bg) launch_bg_bwd(g, *opt, a, color_cache, w.logT, w.accum, gr, st);
    return check_launch(__func__);
}

extern "C" SVOX2_API int svox2_volume_render_cuvol_fused(
    const svox2_grid_t* grid, const float* origins, const float* dirs, int64_t n_rays,
    int64_t norm_rays, const svox2_render_opts_t* opt, const float* rgb_gt, float beta_loss,
    float sparsity_loss, float* rgb_out, const svox2_grads_t* grads, void* workspace,
    size_t workspace_bytes, int64_t* counters, svox2_stream_t stream) {
    if (int e = validate_grid(grid, __func__)) return e;
    if (int e = validate_rays(origins, dirs, n_rays, __func__)) return e;
    if (int e = validate_opt(opt, __func__)) return e;
    if (int e = validate_grads(grads, __func__)) return e;
    SVOX2_REQUIRE((rgb_gt && rgb_out) || n_rays == 0, "rgb_gt/rgb_out must not be null");
    SVOX2_REQUIRE(norm_rays > 0 || n_rays == 0, "norm_rays must be positive");
    if (n_rays == 0) return SVOX2_OK;
    const GridView g = make_grid_view(grid);
    const GradView gr = make_grad_view(grads);
    const bool use_bg = g.bg_nlayers > 0;
    cudaStream_t st = (cudaStream_t)stream;

    MarchArgs a = base_args(origins, dirs, n_rays);
    set_grad_strides(a, grid, grads);
    a.rgb_out = rgb_out;
    a.grad_in = rgb_gt;
    a.beta_loss = beta_loss / (float)norm_rays;            // cuvol.cu:1127
    a.sparsity_loss = sparsity_loss;
    a.norm_factor = 2.f / (float)(3 * (int)norm_rays);     // cuvol.cu:736
    a.fused_seed = 1;
    a.counters = reinterpret_cast<unsigned long long*>(counters);

    const int* overflow = nullptr;   // set: the queue path ran; what follows is its conditional fallback
    float *q_logT = nullptr, *q_accum = nullptr;
    if (use_queue()) {
        // sample-queue path (queue_march.cu): trace -> shade -> [MSI fwd] -> seed -> scatter ->
        // [MSI bwd]; the lattice kernels are enqueued behind it as a conditional fallback for a
        // queue overflow.
        if (launch_queue_fused(g, *opt, a, gr, workspace, workspace_bytes, &overflow, &q_logT, &q_accum, st)) {
            if (!use_bg) {
                MarchArgs fb = a;
                fb.counters = nullptr;
                fb.run_if = overflow;
                launch_lattice_march(LB_MODE_FUSED, false, g, *opt, fb, gr, st);
                return check_launch(__func__);
            }
            a.counters = nullptr;
            a.run_if = overflow;
        }
    }
    if (!use_bg) {
        // single launch: forward pass, gradient seed and backward pass inside one warp per ray.
        // Whatever workspace lies beyond the per-ray scalars is used as the forward->backward
        // colour cache (16 B per active sample); samples beyond a ray's share are re-gathered.
        const size_t base = svox2_render_workspace_bytes(n_rays);
        if (workspace && workspace_bytes > base &&
            (reinterpret_cast<uintptr_t>(workspace) + base) % 16 == 0) {
            const size_t slots = (workspace_bytes - base) / (sizeof(float4) * (size_t)n_rays);
            a.ccap = (int)(slots > 4096 ? 4096 : slots);
            a.ccache = a.ccap > 0 ? reinterpret_cast<float4*>(static_cast<char*>(workspace) + base)
                                  : nullptr;
        }
        launch_lattice_march(LB_MODE_FUSED, false, g, *opt, a, gr, st);
        return check_launch(__func__);
    }
    // MSI present: the background composites between the two foreground passes, so the sequence
    // of the reference (cuvol.cu:1099-1149) is kept: fwd, bg fwd, bwd, bg bwd.
    Workspace w{q_logT, q_accum};   // the fallback must not carve over the queue header it is gated by
    if (!overflow)
        if (int e = carve_workspace(workspace, workspace_bytes, n_rays, w, __func__)) return e;
    MarchArgs f = a;
    f.logT_out = w.logT;
    run_fwd(false, g, *opt, f, st);
    if (use_bg) launch_bg_fwd(false, g, *opt, f, w.logT, st, overflow, 1);
    MarchArgs b = a;
    b.counters = nullptr;
    b.color_cache = rgb_out;
    b.logT_in = beta_loss > 0.f ? w.logT : nullptr;
    b.accum_out = use_bg ? w.accum : nullptr;
    b.logT_out = nullptr;
    if (int e = run_bwd(g, *opt, b, gr, st, __func__)) return e;
    if (use_bg && gr.grad_bg) launch_bg_bwd(g, *opt, b, rgb_out, w.logT, w.accum, gr, st, overflow, 1);
    return check_launch(__func__);
}

extern "C" SVOX2_API int svox2_volume_render_expected_term(const svox2_grid_t* grid,
                                                           const float* origins, const float* dirs,
                                                           int64_t n_rays,
                                                           const svox2_render_opts_t* opt,
                                                           float* out, svox2_stream_t stream) {
    if (int e = validate_grid(grid, __func__)) return e;
    if (int e = validate_rays(origins, dirs, n_rays, __func__)) return e;
    if (int e = validate_opt(opt, __func__)) return e;
    SVOX2_REQUIRE(out || n_rays == 0, "out must not be null");
    if (n_rays == 0) return SVOX2_OK;
    launch_depth(false, make_grid_view(grid), *opt, origins, dirs, n_rays, 0.f, out,
                 (cudaStream_t)stream);
    return check_launch(__func__);
}

extern "C" SVOX2_API int svox2_volume_render_sigma_thresh(const svox2_grid_t* grid,
                                                          const float* origins, const float* dirs,
                                                          int64_t n_rays,
                                                          const svox2_render_opts_t* opt,
                                                          float sigma_thresh, float* out,
                                                          svox2_stream_t stream) {
    if (int e = validate_grid(grid, __func__)) return e;
    if (int e = validate_rays(origins, dirs, n_rays, __func__)) return e;
    if (int e = validate_opt(opt, __func__)) return e;
    SVOX2_REQUIRE(out || n_rays == 0, "out must not be null");
    if (n_rays == 0) return SVOX2_OK;
    launch_depth(true, make_grid_view(grid), *opt, origins, dirs, n_rays, sigma_thresh, out,
                 (cudaStream_t)stream);
    return check_launch(__func__);
}

// ---- slab mode (grid sharded over GPUs by x-range; host side: svox2_b200/slab.py) -------------
extern "C" SVOX2_API int svox2_slab_trace(const svox2_grid_t* grid, const float* origins,
                                          const float* dirs, int64_t n_rays,
                                          const svox2_render_opts_t* opt, int32_t x_lo, int32_t x_hi,
                                          float* logT_local, void* workspace,
                                          size_t workspace_bytes, svox2_stream_t stream) {
    if (int e = validate_grid(grid, __func__)) return e;
    if (int e = validate_rays(origins, dirs, n_rays, __func__)) return e;
    if (int e = validate_opt(opt, __func__)) return e;
    SVOX2_REQUIRE(0 <= x_lo && x_lo <= x_hi && x_hi <= grid->size[0], "bad slab range");
    SVOX2_REQUIRE(grid->background_nlayers == 0, "slab mode does not support an MSI background");
    SVOX2_REQUIRE(logT_local || n_rays == 0, "logT_local must not be null");
    if (n_rays == 0) return SVOX2_OK;
    const GridView g = make_grid_view(grid);
    MarchArgs a = base_args(origins, dirs, n_rays);
    SVOX2_REQUIRE(launch_slab_trace(g, *opt, a, workspace, workspace_bytes, x_lo, x_hi, logT_local,
                                    (cudaStream_t)stream),
                  "workspace too small or not 256-byte aligned (svox2_fused_workspace_bytes)");
    return check_launch(__func__);
}

extern "C" SVOX2_API int svox2_slab_shade(const svox2_grid_t* grid, int64_t n_rays,
                                          const svox2_render_opts_t* opt, const float* lt_in,
                                          float* part, void* workspace, size_t workspace_bytes,
                                          svox2_stream_t stream) {
    if (int e = validate_grid(grid, __func__)) return e;
    if (int e = validate_opt(opt, __func__)) return e;
    SVOX2_REQUIRE((lt_in && part) || n_rays == 0, "null argument");
    if (n_rays == 0) return SVOX2_OK;
    const GridView g = make_grid_view(grid);
    SVOX2_REQUIRE(launch_slab_shade(g, *opt, n_rays, workspace, workspace_bytes, lt_in, part,
                                    (cudaStream_t)stream),
                  "workspace too small or not 256-byte aligned (svox2_fused_workspace_bytes)");
    return check_launch(__func__);
}

extern "C" SVOX2_API int svox2_slab_backward(const svox2_grid_t* grid, int64_t n_rays,
                                             const svox2_render_opts_t* opt, const float* lt_in,
                                             const float* gout, const float* accum_start,
                                             float sparsity_loss, const svox2_grads_t* grads,
                                             void* workspace, size_t workspace_bytes,
                                             svox2_stream_t stream) {
    if (int e = validate_grid(grid, __func__)) return e;
    if (int e = validate_opt(opt, __func__)) return e;
    if (int e = validate_grads(grads, __func__)) return e;
    SVOX2_REQUIRE((lt_in && gout && accum_start) || n_rays == 0, "null argument");
    if (n_rays == 0) return SVOX2_OK;
    const GridView g = make_grid_view(grid);
    const GradView gr = make_grad_view(grads);
    MarchArgs a = base_args(nullptr, nullptr, n_rays);
    set_grad_strides(a, grid, grads);
    a.sparsity_loss = sparsity_loss;
    SVOX2_REQUIRE(launch_slab_backward(g, *opt, a, gr, workspace, workspace_bytes, lt_in, gout,
                                       accum_start, (cudaStream_t)stream),
                  "workspace too small or not 256-byte aligned (svox2_fused_workspace_bytes)");
    return check_launch(__func__);
}

// ---- ship mode (slab.py): rays sharded for the traversal, the grid for everything else ----------
extern "C" SVOX2_API size_t svox2_ship_workspace_bytes(int64_t n_rays_global, int32_t n_slabs,
                                                       int64_t region_cap) {
    if (n_rays_global < 0 || n_slabs < 1 || region_cap < 0) return 0;
    return ship_workspace_bytes(n_rays_global, n_slabs, region_cap);
}

extern "C" SVOX2_API int svox2_ship_trace(const svox2_grid_t* grid, const float* origins, const float* dirs,
                                          int64_t n_rays_global, int64_t ray_begin, int64_t ray_count,
                                          const svox2_render_opts_t* opt, const int32_t* slab_bounds,
                                          int32_t n_slabs, int32_t rank, const uint64_t* peer_workspaces,
                                          size_t workspace_bytes, int64_t region_cap, int64_t* counters,
                                          svox2_stream_t stream) {
    if (int e = validate_grid(grid, __func__)) return e;
    if (int e = validate_rays(origins, dirs, n_rays_global, __func__)) return e;
    if (int e = validate_opt(opt, __func__)) return e;
    SVOX2_REQUIRE(n_slabs >= 1 && n_slabs <= SVOX2_MAX_PEERS && rank >= 0 && rank < n_slabs && slab_bounds &&
                      peer_workspaces,
                  "bad slab arguments");
    SVOX2_REQUIRE(ray_begin >= 0 && ray_count >= 0 && ray_begin + ray_count <= n_rays_global, "bad ray range");
    SVOX2_REQUIRE(grid->background_nlayers == 0, "slab mode does not support an MSI background");
    for (int d = 0; d < n_slabs; ++d)
        SVOX2_REQUIRE(slab_bounds[d] <= slab_bounds[d + 1] && peer_workspaces[d] != 0, "bad slab bounds / peer pointer");
    if (n_rays_global == 0) return SVOX2_OK;
    const GridView g = make_grid_view(grid);
    MarchArgs a = base_args(origins, dirs, ray_count);
    a.counters = reinterpret_cast<unsigned long long*>(counters);
    SVOX2_REQUIRE(launch_ship_trace(g, *opt, a, ray_begin, n_rays_global, slab_bounds, n_slabs, rank,
                                    peer_workspaces, workspace_bytes, region_cap, (cudaStream_t)stream),
                  "workspace too small or not 256-byte aligned (svox2_ship_workspace_bytes)");
    return check_launch(__func__);
}

extern "C" SVOX2_API int svox2_ship_shade(const svox2_grid_t* grid, int64_t n_rays_global, int32_t n_slabs,
                                          int64_t region_cap, float* part, void* workspace,
                                          size_t workspace_bytes, svox2_stream_t stream) {
    if (int e = validate_grid(grid, __func__)) return e;
    SVOX2_REQUIRE(part || n_rays_global == 0, "null argument");
    if (n_rays_global == 0) return SVOX2_OK;
    const GridView g = make_grid_view(grid);
    SVOX2_REQUIRE(launch_ship_shade(g, n_rays_global, workspace, workspace_bytes, n_slabs, region_cap, part,
                                    (cudaStream_t)stream),
                  "workspace too small or not 256-byte aligned (svox2_ship_workspace_bytes)");
    return check_launch(__func__);
}

extern "C" SVOX2_API const float* svox2_ship_ray_logT(void* workspace, size_t workspace_bytes,
                                                      int64_t n_rays_global, int32_t n_slabs,
                                                      int64_t region_cap) {
    return ship_ray_logT(workspace, workspace_bytes, n_rays_global, n_slabs, region_cap);
}

extern "C" SVOX2_API int svox2_ship_backward(const svox2_grid_t* grid, int64_t n_rays_global,
                                             const svox2_render_opts_t* opt, int32_t n_slabs,
                                             int64_t region_cap, const float* gout, const float* accum_start,
                                             float sparsity_loss, const svox2_grads_t* grads, void* workspace,
                                             size_t workspace_bytes, svox2_stream_t stream) {
    if (int e = validate_grid(grid, __func__)) return e;
    if (int e = validate_opt(opt, __func__)) return e;
    if (int e = validate_grads(grads, __func__)) return e;
    SVOX2_REQUIRE((gout && accum_start) || n_rays_global == 0, "null argument");
    if (n_rays_global == 0) return SVOX2_OK;
    const GridView g = make_grid_view(grid);
    const GradView gr = make_grad_view(grads);
    MarchArgs a = base_args(nullptr, nullptr, n_rays_global);
    set_grad_strides(a, grid, grads);
    a.sparsity_loss = sparsity_loss;
    SVOX2_REQUIRE(launch_ship_backward(g, *opt, a, gr, workspace, workspace_bytes, n_slabs, region_cap, gout,
                                       accum_start, (cudaStream_t)stream),
                  "workspace too small or not 256-byte aligned (svox2_ship_workspace_bytes)");
    return check_launch(__func__);
}
