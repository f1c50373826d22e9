/*
 * svox2_b200.h -- C-ABI of the B200-native Plenoxels hot path (libsvox2_b200.so).
 *
 * One entry point per operator the reference exports from its pybind module `svox2.csrc`
 * (svox2/csrc/svox2.cpp:74-172).  Signatures take raw DEVICE pointers, sizes, POD option
 * structs and a cudaStream_t; they never allocate through a framework, never synchronise
 * the device and return an int status (0 = ok).  On failure svox2_last_error() holds a
 * message.  The reference prints CUDA errors and carries on (cuda_util.cuh:18-21); here
 * they are returned so the host shim can raise.
 *
 * Conventions
 *   - All pointers are device pointers on the CURRENT device unless marked HOST.
 *   - float32 data, int32 links, uint8 (bool) masks, as in the reference
 *     (data_spec.hpp:18-135).  Tensors are dense row-major; the two data arrays carry an
 *     explicit row stride so that 16-byte padded rows (stride 28/32) can be used:
 *     sh_stride == 3*basis_dim and density_stride == 1 reproduce the reference layout.
 *   - Gradient outputs are ACCUMULATED into (atomic add), masks are OR-ed into
 *     (render_util.cuh:137-141); callers zero them as the reference's callers do.
 *   - Kernels are enqueued on `stream` (the reference uses the legacy default stream).
 *   - Limits: Q*32, X*Y*Z and N*sh_stride must fit in int64 (the reference: int32).
 */
#ifndef SVOX2_B200_H
#define SVOX2_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SVOX2_API __attribute__((visibility("default")))
#else
#define SVOX2_API
#endif

typedef struct CUstream_st* svox2_stream_t; /* == cudaStream_t */

enum {
    SVOX2_OK = 0,
    SVOX2_ERR_INVALID = 1, /* bad argument (the reference: TORCH_CHECK / spec.check())      */
    SVOX2_ERR_CUDA = 2,    /* CUDA runtime / launch error                                   */
    SVOX2_ERR_WORKSPACE = 3 /* workspace too small                                          */
};

/* SparseGridSpec (data_spec.hpp:18-67) + PackedSparseGridSpec (data_spec_packed.cuh:11-62) */
typedef struct {
    const float* density_data;      /* [N, density_stride], channel 0 is sigma               */
    const float* sh_data;           /* [N, sh_stride], first 3*basis_dim floats are used     */
    const int32_t* links;           /* [X, Y, Z] contiguous                                  */
    int32_t size[3];                /* X, Y, Z                                               */
    int32_t basis_dim;              /* 1, 4 or 9 (SH only: basis_type == 1)                  */
    int64_t n_rows;                 /* N                                                     */
    int64_t sh_stride;              /* floats between consecutive sh rows                    */
    int64_t density_stride;         /* floats between consecutive density rows               */
    float offset[3];                /* SparseGridSpec._offset  (HOST values, by value)       */
    float scaling[3];               /* SparseGridSpec._scaling                               */
    const int32_t* background_links; /* [2R, R] or NULL                                      */
    const float* background_data;   /* [Nbg, L, 4] or NULL                                   */
    int32_t background_reso;        /* R                                                     */
    int32_t background_nlayers;     /* L; 0 disables the MSI                                 */
    int64_t background_rows;        /* Nbg                                                   */
} svox2_grid_t;

/* RenderOptions (data_spec.hpp:115-135) */
typedef struct {
    float background_brightness;
    float step_size;
    float sigma_thresh;
    float stop_thresh;
    float near_clip;
    int32_t use_spheric_clip;
    int32_t last_sample_opaque;
} svox2_render_opts_t;

/* CameraSpec (data_spec.hpp:82-103): c2w stays on the device, as in the reference */
typedef struct {
    const float* c2w;               /* device, [3 or 4, 4] row-major                         */
    float fx, fy, cx, cy;
    int32_t width, height;
    float ndc_coeffx, ndc_coeffy;   /* <= 0 disables NDC                                     */
} svox2_camera_t;

/* GridOutputGrads (data_spec.hpp:69-80); strides are those of the matching data arrays */
typedef struct {
    float* grad_density_out;        /* [N, density_stride] or NULL                           */
    float* grad_sh_out;             /* [N, sh_stride] or NULL                                */
    float* grad_background_out;     /* [Nbg, L, 4] or NULL                                   */
    uint8_t* mask_out;              /* [N] or NULL                                           */
    uint8_t* mask_background_out;   /* [Nbg, L] or NULL                                      */
    int64_t grad_sh_stride;         /* floats between grad_sh rows; 0 = grid->sh_stride      */
    int64_t grad_density_stride;    /* floats between grad_density rows; 0 = density_stride  */
} svox2_grads_t;

/* Extension (not in the reference): per-batch march-loop iteration counts, the unit of the
 * roofline model.  Device array of 4 int64: skip, visited, active, bg. */
#define SVOX2_N_COUNTERS 4

SVOX2_API const char* svox2_last_error(void);
SVOX2_API const char* svox2_version(void);
/* 1 if a CUDA device is usable from this process (lets host code fail loudly). */
SVOX2_API int svox2_cuda_available(void);

/* Scratch bytes the renderer entry points need for Q rays (log_transmit + accum). */
SVOX2_API size_t svox2_render_workspace_bytes(int64_t n_rays);

/* Scratch for svox2_volume_render_cuvol_fused sized for an average of `slots_per_ray` active
 * samples per ray (record queue of the sample-queue kernels / colour cache of the single-kernel
 * path).  Any size >= svox2_render_workspace_bytes() is legal: overflow falls back, never fails.
 * The workspace must be 256-byte aligned. */
SVOX2_API size_t svox2_fused_workspace_bytes(int64_t n_rays, int64_t slots_per_ray);

/* volume_render_cuvol  (svox2.cpp:18, render_lerp_kernel_cuvol.cu:918-958)
 * out [Q,3].  counters may be NULL. */
SVOX2_API int svox2_volume_render_cuvol(const svox2_grid_t* grid, const float* origins, const float* dirs,
                              int64_t n_rays, const svox2_render_opts_t* opt, float* out,
                              void* workspace, size_t workspace_bytes, int64_t* counters,
                              svox2_stream_t stream);

/* volume_render_cuvol_image  (svox2.cpp:19-20, cuvol.cu:960-1008); out [H,W,3].
 * row_begin/row_end restrict the render to image rows [row_begin,row_end) (multi-GPU row
 * sharding); pass 0,height for the reference behaviour.  workspace for (row_end-row_begin)*W rays. */
SVOX2_API int svox2_volume_render_cuvol_image(const svox2_grid_t* grid, const svox2_camera_t* cam,
                                    const svox2_render_opts_t* opt, int32_t row_begin,
                                    int32_t row_end, float* out, void* workspace,
                                    size_t workspace_bytes, svox2_stream_t stream);

/* volume_render_cuvol_backward  (svox2.cpp:21-23, cuvol.cu:1010-1068) */
SVOX2_API int svox2_volume_render_cuvol_backward(const svox2_grid_t* grid, const float* origins,
                                       const float* dirs, int64_t n_rays,
                                       const svox2_render_opts_t* opt, const float* grad_out,
                                       const float* color_cache, const svox2_grads_t* grads,
                                       void* workspace, size_t workspace_bytes,
                                       svox2_stream_t stream);

/* volume_render_cuvol_fused  (svox2.cpp:24-26, cuvol.cu:1070-1152)
 * norm_rays: ray count used in the MSE normalisation 2/(3*norm_rays) and beta_loss/norm_rays.
 * The reference uses the local batch size (cuvol.cu:736,1127): pass n_rays for drop-in
 * behaviour, the GLOBAL batch size when the batch is sharded over GPUs. */
SVOX2_API int svox2_volume_render_cuvol_fused(const svox2_grid_t* grid, const float* origins,
                                    const float* dirs, int64_t n_rays, int64_t norm_rays,
                                    const svox2_render_opts_t* opt, const float* rgb_gt,
                                    float beta_loss, float sparsity_loss, float* rgb_out,
                                    const svox2_grads_t* grads, void* workspace,
                                    size_t workspace_bytes, int64_t* counters,
                                    svox2_stream_t stream);

/* volume_render_expected_term / volume_render_sigma_thresh (svox2.cpp:27-30, cuvol.cu:1154-1195) */
SVOX2_API int svox2_volume_render_expected_term(const svox2_grid_t* grid, const float* origins,
                                      const float* dirs, int64_t n_rays,
                                      const svox2_render_opts_t* opt, float* out,
                                      svox2_stream_t stream);
SVOX2_API int svox2_volume_render_sigma_thresh(const svox2_grid_t* grid, const float* origins,
                                     const float* dirs, int64_t n_rays,
                                     const svox2_render_opts_t* opt, float sigma_thresh,
                                     float* out, svox2_stream_t stream);

/* volume_render_nvol / _backward / _fused (svox2.cpp:80-85; render_lerp_kernel_nvol.cu:406-502):
 * the Neural-Volumes compositing ablation (absolute transmittance) -- the backend the reference's
 * own parity script exercises (test/test_render_gradcheck.py:20).  No MSI background, no beta
 * loss (the reference ignores that argument too).  No workspace is needed. */
SVOX2_API int svox2_volume_render_nvol(const svox2_grid_t* grid, const float* origins,
                             const float* dirs, int64_t n_rays, const svox2_render_opts_t* opt,
                             float* out, svox2_stream_t stream);
SVOX2_API int svox2_volume_render_nvol_backward(const svox2_grid_t* grid, const float* origins,
                                      const float* dirs, int64_t n_rays,
                                      const svox2_render_opts_t* opt, const float* grad_out,
                                      const float* color_cache, const svox2_grads_t* grads,
                                      svox2_stream_t stream);
SVOX2_API int svox2_volume_render_nvol_fused(const svox2_grid_t* grid, const float* origins,
                                   const float* dirs, int64_t n_rays, int64_t norm_rays,
                                   const svox2_render_opts_t* opt, const float* rgb_gt,
                                   float sparsity_loss, float* rgb_out,
                                   const svox2_grads_t* grads, svox2_stream_t stream);

/* sample_grid / sample_grid_backward  (svox2.cpp:11-17, svox2_kernel.cu:178-267)
 * out_density [P,1]; out_sh [P,3*basis_dim] or NULL (want_colors=false).
 * Unlike the reference no extra streams are created and nothing is synchronised. */
SVOX2_API int svox2_sample_grid(const svox2_grid_t* grid, const float* points, int64_t n_points,
                      float* out_density, float* out_sh, svox2_stream_t stream);
SVOX2_API int svox2_sample_grid_backward(const svox2_grid_t* grid, const float* points, int64_t n_points,
                               const float* grad_out_density, const float* grad_out_sh,
                               float* grad_density_out, int64_t grad_density_stride,
                               float* grad_sh_out, int64_t grad_sh_stride,
                               svox2_stream_t stream);

/* dilate  (svox2.cpp:60, misc_kernel.cu:20-48,339-354): 3x3x3 OR, clamped borders. */
SVOX2_API int svox2_dilate(const uint8_t* grid, int32_t X, int32_t Y, int32_t Z, uint8_t* out,
                 svox2_stream_t stream);

/* accel_dist_prop  (svox2.cpp:61, misc_kernel.cu:109-181,356-416): in place on links.
 * workspace >= svox2_accel_workspace_bytes(X,Y,Z). */
SVOX2_API size_t svox2_accel_workspace_bytes(int32_t X, int32_t Y, int32_t Z);
SVOX2_API int svox2_accel_dist_prop(int32_t* links, int32_t X, int32_t Y, int32_t Z, void* workspace,
                          size_t workspace_bytes, svox2_stream_t stream);

/* grid_weight_render  (svox2.cpp:62-65, misc_kernel.cu:184-334,418-445)
 * data / grid_weight_out dense [X,Y,Z]; offset/scaling DEVICE float[3] as in the reference. */
SVOX2_API int svox2_grid_weight_render(const float* data, int32_t X, int32_t Y, int32_t Z,
                             const svox2_camera_t* cam, float step_size, float stop_thresh,
                             int32_t last_sample_opaque, const float* offset,
                             const float* scaling, float* grid_weight_out,
                             svox2_stream_t stream);

/* tv / tv_grad / tv_grad_sparse / msi_tv_grad_sparse  (svox2.cpp:32-49, loss_kernel.cu:482-641)
 * data [N, data_stride] with ddim logical channels; grad_data [N, grad_stride] (0 = data_stride).
 * tv: *out (device scalar) is overwritten. */
SVOX2_API int svox2_tv(const int32_t* links, int32_t X, int32_t Y, int32_t Z, const float* data,
             int64_t data_stride, int32_t start_dim, int32_t end_dim, int32_t ignore_edge,
             float ndc_coeffx, float ndc_coeffy, float* out, svox2_stream_t stream);
SVOX2_API int svox2_tv_grad(const int32_t* links, int32_t X, int32_t Y, int32_t Z, const float* data,
                  int64_t data_stride, int32_t start_dim, int32_t end_dim, float scale,
                  int32_t ignore_edge, float ndc_coeffx, float ndc_coeffy, float* grad_data,
                  int64_t grad_stride, svox2_stream_t stream);
/* scale_cells: the cell count the scale is divided by (the reference: n_cells,
 * loss_kernel.cu:584-595; the global count when cells are sharded over GPUs). */
SVOX2_API int svox2_tv_grad_sparse(const int32_t* links, int32_t X, int32_t Y, int32_t Z,
                         const float* data, int64_t data_stride, const int32_t* rand_cells,
                         int64_t n_cells, int64_t scale_cells, uint8_t* mask_out,
                         int32_t start_dim, int32_t end_dim, float scale, int32_t ignore_edge,
                         int32_t ignore_last_z, float ndc_coeffx, float ndc_coeffy,
                         float* grad_data, int64_t grad_stride, svox2_stream_t stream);
/* Same as svox2_tv_grad_sparse for the contiguous wrapped cell run [cell_start, cell_start+n_cells)
 * mod X*Y*Z that SparseGrid._get_rand_cells(contiguous=True) builds (svox2.py:2230-2237), without
 * materialising the index tensor. */
SVOX2_API int svox2_tv_grad_sparse_run(const int32_t* links, int32_t X, int32_t Y, int32_t Z,
                             const float* data, int64_t data_stride, int64_t cell_start,
                             int64_t n_cells, int64_t scale_cells, uint8_t* mask_out,
                             int32_t start_dim, int32_t end_dim, float scale, int32_t ignore_edge,
                             int32_t ignore_last_z, float* grad_data, int64_t grad_stride,
                             svox2_stream_t stream);
SVOX2_API int svox2_msi_tv_grad_sparse(const int32_t* background_links, int32_t reso, const float* msi,
                             int32_t nlayers, int32_t nchannels, const int32_t* rand_cells,
                             int64_t n_cells, int64_t scale_cells, uint8_t* mask_out, float scale,
                             float scale_last, float* grad_msi, svox2_stream_t stream);

/* ---- Extension: the whole train step of opt/opt.py:474-603 as one call ------------------------
 * volume_render_cuvol_fused + tv_grad_sparse (density, contiguous run) + tv_grad_sparse (SH) +
 * rmsprop_step (density, SH), with the step's gradients kept in a compact [touched rows, W] buffer
 * inside the workspace instead of dense .grad arrays.  Same kernels / arithmetic as
 * the separate entry points; density_data, sh_data (through `grid`, written despite the const) and
 * the two RMS arrays are updated in place.  No MSI background; sh rows must be 16-byte aligned.
 * The workspace is created once (svox2_train_step_workspace_bytes + svox2_train_step_init) and
 * must be passed unchanged to every step on this grid / batch size.  slots_per_ray sizes the
 * sample-record queue (records per ray on average); if it overflows the step's ray gradients are
 * dropped and the sticky counter reported by svox2_train_step_masks is incremented.
 * mse_out (device float, may be NULL) receives mean((rgb - rgb_gt)^2) of the batch, the loss opt.py
 * reads back every step (opt.py:505-508).  The TV work runs on a library-owned side stream between
 * event pairs recorded on `stream` (capturable; SVOX2_B200_STEP_OVERLAP=0 disables it). */
typedef struct {
    int64_t tv_density_start;   /* first cell of the wrapped contiguous run (svox2.py:2230-2237)  */
    int64_t tv_density_count;   /* 0 disables the density TV                                      */
    float tv_density_scale;     /* lambda_tv                                                      */
    int64_t tv_sh_start;
    int64_t tv_sh_count;
    float tv_sh_scale;          /* lambda_tv_sh                                                   */
    int64_t tv_scale_cells;     /* cells the TV is normalised by; 0 = the run's count             */
    float beta_loss;            /* lambda_beta (divided by norm_rays inside, cuvol.cu:1127)       */
    float sparsity_loss;
    int64_t norm_rays;          /* batch size of the MSE normalisation (global batch if sharded)  */
    float lr_density, lr_sh, rms_beta, epsilon;
} svox2_step_params_t;
SVOX2_API size_t svox2_train_step_workspace_bytes(int64_t n_rows, int32_t basis_dim, int64_t n_rays,
                                                  int64_t slots_per_ray);
SVOX2_API int svox2_train_step_init(void* workspace, size_t workspace_bytes, int64_t n_rows,
                                    int32_t basis_dim, int64_t n_rays, int64_t slots_per_ray,
                                    svox2_stream_t stream);
SVOX2_API int svox2_train_step(const svox2_grid_t* grid, float* density_rms, float* sh_rms,
                               const float* origins, const float* dirs, int64_t n_rays,
                               const svox2_render_opts_t* opt, const float* rgb_gt, float* rgb_out,
                               const svox2_step_params_t* params, void* workspace,
                               size_t workspace_bytes, int64_t slots_per_ray, float* mse_out,
                               int64_t* counters, svox2_stream_t stream);
/* The two sparse masks of the LAST step as uint8[N] (OR-ed into the arrays; either may be NULL) and
 * device int32[3] = {touched rows, queue overflow flag of the last step, overflowed steps before it}. */
SVOX2_API int svox2_train_step_masks(void* workspace, size_t workspace_bytes, int64_t n_rows,
                                     int32_t basis_dim, int64_t n_rays, int64_t slots_per_ray,
                                     uint8_t* mask_density, uint8_t* mask_sh, int32_t* status3,
                                     svox2_stream_t stream);

/* lumisphere_tv_grad_sparse (svox2.cpp:50-51, loss_kernel.cu:335-434, 643-678): TV of the colour
 * seen from one view direction; basis_fn / basis_fn_u = SH basis values [basis_dim] of that
 * direction and of a perturbed one (weight dir_factor).  scale is divided by n_cells inside.
 * Cells whose decoded x lies outside the grid (the reference reads past `links` there) are skipped. */
SVOX2_API int svox2_lumisphere_tv_grad_sparse(const int32_t* links, int32_t X, int32_t Y, int32_t Z,
                                    const float* sh_data, int64_t sh_stride, int32_t basis_dim,
                                    const int32_t* rand_cells, int64_t n_cells,
                                    const float* basis_fn, const float* basis_fn_u, float scale,
                                    float ndc_coeffx, float ndc_coeffy, float dir_factor,
                                    uint8_t* mask_out, float* grad_sh, int64_t grad_stride,
                                    svox2_stream_t stream);

/* rmsprop_step / sgd_step  (svox2.cpp:52-58, optim_kernel.cu:154-267)
 * data/rms/grad [rows, row_stride] with n_channels logical channels; the last channel uses
 * lr_last (lr_last < 0 => lr).  indexer_kind: 0 dense, 1 uint8 mask[rows], 2 int64 index[n_index].
 * The grad of every updated element is zeroed. */
enum { SVOX2_INDEXER_DENSE = 0, SVOX2_INDEXER_MASK = 1, SVOX2_INDEXER_INDEX = 2 };
SVOX2_API int svox2_rmsprop_step(float* data, float* rms, float* grad, int64_t rows, int32_t n_channels,
                       int64_t row_stride, int32_t indexer_kind, const uint8_t* mask,
                       const int64_t* index, int64_t n_index, float beta, float lr, float epsilon,
                       float minval, float lr_last, svox2_stream_t stream);
SVOX2_API int svox2_sgd_step(float* data, float* grad, int64_t rows, int32_t n_channels,
                   int64_t row_stride, int32_t indexer_kind, const uint8_t* mask,
                   const int64_t* index, int64_t n_index, float lr, float lr_last,
                   svox2_stream_t stream);

/* Multi-GPU gradient exchange (no reference counterpart).
 * svox2_pack_grad_rows: packed[i, 0:C] = grad_sh[rows[i], 0:C], packed[i, C] = grad_density[rows[i]],
 *   then (optional) mask_a[rows[i]], mask_b[rows[i]] as 0/1 floats, remaining floats 0.
 *   rows is an int64 list identical on every rank that takes part.
 * svox2_rmsprop_step_packed: the masked RMSProp of svox2_rmsprop_step (optim_kernel.cu:15-25) for
 *   density (if mask_density[row]) and SH (if mask_sh[row]) of every listed row, gradients read
 *   from the all-reduced packed buffer; if grad_sh/grad_density are given, those rows of the
 *   local gradient arrays are zeroed (as svox2_rmsprop_step zeroes the gradients it consumed).
 * Both require 16-byte aligned SH rows padded to a multiple of 4 floats. */
SVOX2_API int svox2_pack_grad_rows(float* grad_sh, int64_t grad_sh_stride, int32_t sh_channels,
                         float* grad_density, int64_t grad_density_stride, const int64_t* rows,
                         int64_t n_sel, float* packed, int64_t pack_stride,
                         const uint8_t* mask_a, const uint8_t* mask_b, svox2_stream_t stream);
/* svox2_unpack_rows: the inverse of svox2_pack_grad_rows for any pair of [rows, stride] arrays:
 *   dst[rows[i]] = packed[i] (add == 0) or += packed[i] (add != 0); mask_a/mask_b (optional) are
 *   OR-ed with the flags packed behind the density slot (mask_a/mask_b of svox2_pack_grad_rows). */
SVOX2_API int svox2_unpack_rows(const float* packed, int64_t pack_stride, const int64_t* rows,
                         int64_t n_sel, float* dst_sh, int64_t dst_sh_stride, int32_t sh_channels,
                         float* dst_density, int64_t dst_density_stride, int32_t add,
                         uint8_t* mask_a, uint8_t* mask_b, svox2_stream_t stream);
SVOX2_API int svox2_rmsprop_step_packed(float* sh_data, float* sh_rms, int64_t sh_stride,
                         int32_t sh_channels, float* density_data, float* density_rms,
                         int64_t density_stride, const int64_t* rows, int64_t n_sel,
                         const float* packed, int64_t pack_stride, const uint8_t* mask_sh,
                         const uint8_t* mask_density, float* grad_sh, int64_t grad_sh_stride,
                         float* grad_density, int64_t grad_density_stride, float beta,
                         float lr_sh, float lr_density, float epsilon, float minval_sh, float minval_density,
                         svox2_stream_t stream);

/* Slab mode: the grid is sharded over GPUs by x-range of cells [x_lo, x_hi) (no reference
 * counterpart).  Every rank traverses ALL rays of the global batch, so the sample
 * lattice is the reference's (render_lerp_kernel_cuvol.cu:56-109), but gathers / records only the
 * samples whose cell lies in its slab.  Between the three calls the ranks exchange per-ray scalars:
 *   svox2_slab_trace    -> logT_local[Q]   (this slab's sum of -sigma*delta)
 *   ... all-gather, lt_in[Q] = log-transmittance at slab entry along each ray ...
 *   svox2_slab_shade    -> part[Q,4] (caller zeroes): .xyz += this slab's colour, .w = 1 where the
 *                          stop threshold fired
 *   ... all-gather, total colour -> gout[Q,3] = dL/d(rgb), accum_start[Q] = the reference's
 *       `accum` (cuvol.cu:248-264) at slab entry ...
 *   svox2_slab_backward -> gradients / mask of this slab's samples.
 * The workspace (svox2_fused_workspace_bytes) carries the sample records between the calls. */
SVOX2_API int svox2_slab_trace(const svox2_grid_t* grid, const float* origins, const float* dirs,
                     int64_t n_rays, const svox2_render_opts_t* opt, int32_t x_lo, int32_t x_hi,
                     float* logT_local, void* workspace, size_t workspace_bytes,
                     svox2_stream_t stream);
SVOX2_API int svox2_slab_shade(const svox2_grid_t* grid, int64_t n_rays, const svox2_render_opts_t* opt,
                     const float* lt_in, float* part, void* workspace, size_t workspace_bytes,
                     svox2_stream_t stream);
SVOX2_API int svox2_slab_backward(const svox2_grid_t* grid, int64_t n_rays,
                     const svox2_render_opts_t* opt, const float* lt_in, const float* gout,
                     const float* accum_start, float sparsity_loss, const svox2_grads_t* grads,
                     void* workspace, size_t workspace_bytes, svox2_stream_t stream);

/* svox2_peer_store: copies [src, src+n_bytes) to base_p + dst_offset_bytes for every p, where
 * base_p are n_peers device pointers mapped into this process (symmetric / peer memory over
 * NVLink).  One launch = this rank's contribution to an all-gather; the caller provides the
 * cross-GPU barrier before the gathered buffer is read. */
#define SVOX2_MAX_PEERS 16
SVOX2_API int svox2_peer_store(const uint64_t* peer_base_ptrs, int32_t n_peers,
                     int64_t dst_offset_bytes, const void* src, int64_t n_bytes,
                     svox2_stream_t stream);

/* per-ray glue between the slab calls (one thread per ray; all arrays of the all-gathers are
 * [n_slabs, Q(, 4)]): lt_in = prefix of logT over the slabs before `rank` in ray order;
 * combine = total colour (+ background term unless stopped), gout = (rgb - rgb_gt) * 2/(3Q),
 * accum_start = rgb.gout (+ beta term) - sum over earlier slabs of gout.rgb_partial. */
SVOX2_API int svox2_slab_prefix(const float* logT_all, const float* dirs, int64_t n_rays,
                     int32_t n_slabs, int32_t rank, float* lt_in, svox2_stream_t stream);
SVOX2_API int svox2_slab_combine(const float* part_all, const float* logT_all, const float* dirs,
                     const float* rgb_gt, int64_t n_rays, int32_t n_slabs, int32_t rank,
                     float background_brightness, float beta_loss, float* rgb, float* gout,
                     float* accum_start, svox2_stream_t stream);

/* Ship mode (default multi-GPU step, svox2_b200/slab.py): the RAYS are sharded for the traversal,
 * the GRID (x-slabs) for everything else.  The rank that traces a ray composites it exactly as the
 * single-GPU kernel does (same lattice, same sequential transmittance: masks stay bit-exact) and
 * ships every active sample's 96-byte record to the rank that owns the sample's cell, straight
 * into that rank's queue over NVLink: `peer_workspaces[d]` is rank d's workspace mapped into this
 * process (symmetric memory), all of identical size / layout; region `rank` of every queue belongs
 * to the records shipped by `rank` (region_cap records).  After a cross-GPU barrier each rank
 * shades the records it received (part[Qg,4] partial colours), the partials are all-gathered,
 * svox2_ship_combine makes total colour / dL/drgb / the reference's accum at slab entry, and
 * svox2_ship_backward scatters the gradients of the received records and resets the queue.
 * origins / dirs / rgb_gt hold the GLOBAL batch on every rank; rank r traces
 * [ray_begin, ray_begin + ray_count). */
SVOX2_API size_t svox2_ship_workspace_bytes(int64_t n_rays_global, int32_t n_slabs, int64_t region_cap);
SVOX2_API int svox2_ship_trace(const svox2_grid_t* grid, const float* origins, const float* dirs,
                     int64_t n_rays_global, int64_t ray_begin, int64_t ray_count,
                     const svox2_render_opts_t* opt, const int32_t* slab_bounds, int32_t n_slabs,
                     int32_t rank, const uint64_t* peer_workspaces, size_t workspace_bytes,
                     int64_t region_cap, int64_t* counters, svox2_stream_t stream);
SVOX2_API int svox2_ship_shade(const svox2_grid_t* grid, int64_t n_rays_global, int32_t n_slabs,
                     int64_t region_cap, float* part, void* workspace, size_t workspace_bytes,
                     svox2_stream_t stream);
SVOX2_API const float* svox2_ship_ray_logT(void* workspace, size_t workspace_bytes, int64_t n_rays_global,
                     int32_t n_slabs, int64_t region_cap);
SVOX2_API int svox2_ship_combine(const float* part_all, const float* ray_logT, const float* dirs,
                     const float* rgb_gt, int64_t n_rays, int32_t n_slabs, int32_t rank,
                     float background_brightness, float beta_loss, float* rgb, float* gout,
                     float* accum_start, svox2_stream_t stream);
SVOX2_API int svox2_ship_backward(const svox2_grid_t* grid, int64_t n_rays_global,
                     const svox2_render_opts_t* opt, int32_t n_slabs, int64_t region_cap,
                     const float* gout, const float* accum_start, float sparsity_loss,
                     const svox2_grads_t* grads, void* workspace, size_t workspace_bytes,
                     svox2_stream_t stream);

/* Masked RMSProp (optim_kernel.cu:15-25) on a single-channel array that ship mode keeps replicated
 * (density), with the publication fused in: send[row] = 0 keeps the new value local, 1 stores it at
 * the same offset of every peer's array, 2 of every peer but `skip_peer` (a co-owner of the row that
 * applies the identical update itself). */
SVOX2_API int svox2_rmsprop_step_publish(float* data, float* rms, float* grad, int64_t rows,
                     int64_t row_stride, const uint8_t* mask, const uint8_t* send,
                     const uint64_t* peer_base_ptrs, int32_t n_peers, int32_t self_rank,
                     int32_t skip_peer, float beta, float lr, float epsilon, float minval,
                     svox2_stream_t stream);

/* Multi-GPU train step on the compact gradient buffer (step.cu; host side svox2_b200/slab.py): the
 * single-call step above cut at the points where the x-slab ranks exchange per-ray scalars and
 * boundary planes.  `step_ws` is a svox2_train_step workspace sized for the GLOBAL batch; in ship
 * mode the sample records live in the symmetric ship workspaces instead of its queue part.
 *   svox2_mg_step_begin     clear, TV marking of this rank's pieces of the step's two TV runs,
 *                           boundary rows (co-owned planes, row 0) entered into the touched set, traversal
 *   svox2_mg_step_shade     part[Qg,4] += colours of this slab's records; touched-row bits
 *   svox2_mg_step_backward  slot numbering, seed, scatter -> slot rows, TV -> slot rows
 *   svox2_mg_slot_pack / svox2_mg_slot_unpack_add   boundary rows <-> 32-float packets
 *                           [slot row | density flag | SH flag] (the layout of svox2_pack_grad_rows)
 *   svox2_mg_step_finish    RMSProp over the slot list; with send / peer_density (ship mode) the new
 *                           density of a row is also stored into the peers' copies (see
 *                           svox2_rmsprop_step_publish for the send codes; skip_peer = rank - 1) */
typedef struct svox2_mg_step_t {
    int32_t n_slabs, rank, x_lo, x_hi;
    int32_t ship;                                      /* 0: every rank traverses all rays; 1: ship mode */
    int64_t tv_density_start[2], tv_density_count[2];  /* this rank's pieces of the density TV run        */
    int64_t tv_sh_start[2], tv_sh_count[2];
    float tv_density_scale, tv_sh_scale;               /* lambda; 0 disables                              */
    int64_t tv_density_scale_cells, tv_sh_scale_cells; /* cells of the WHOLE run (the normaliser)         */
    float sparsity_loss;
    float lr_density, lr_sh, rms_beta, epsilon;
} svox2_mg_step_t;
SVOX2_API int svox2_mg_step_begin(const svox2_grid_t* grid, const float* origins, const float* dirs,
                     int64_t n_rays_global, const svox2_render_opts_t* opt, const svox2_mg_step_t* params,
                     const int64_t* boundary_rows, int64_t n_boundary, void* step_ws, size_t step_ws_bytes,
                     int64_t slots_per_ray, float* logT_local, const int32_t* slab_bounds,
                     const uint64_t* peer_ship_workspaces, size_t ship_ws_bytes, int64_t region_cap,
                     svox2_stream_t stream);
SVOX2_API int svox2_mg_step_shade(const svox2_grid_t* grid, int64_t n_rays_global,
                     const svox2_render_opts_t* opt, const svox2_mg_step_t* params, const float* lt_in,
                     float* part, void* step_ws, size_t step_ws_bytes, int64_t slots_per_ray, void* ship_ws,
                     size_t ship_ws_bytes, int64_t region_cap, svox2_stream_t stream);
SVOX2_API int svox2_mg_step_backward(const svox2_grid_t* grid, int64_t n_rays_global,
                     const svox2_render_opts_t* opt, const svox2_mg_step_t* params, const float* lt_in,
                     const float* gout, const float* accum_start, void* step_ws, size_t step_ws_bytes,
                     int64_t slots_per_ray, void* ship_ws, size_t ship_ws_bytes, int64_t region_cap,
                     svox2_stream_t stream);
SVOX2_API int svox2_mg_slot_pack(const int64_t* rows, int64_t n_sel, float* packed, int64_t pack_stride,
                     int32_t clear, void* step_ws, size_t step_ws_bytes, int64_t n_rows, int32_t basis_dim,
                     int64_t n_rays, int64_t slots_per_ray, svox2_stream_t stream);
SVOX2_API int svox2_mg_slot_unpack_add(const float* packed, int64_t pack_stride, const int64_t* rows,
                     int64_t n_sel, void* step_ws, size_t step_ws_bytes, int64_t n_rows, int32_t basis_dim,
                     int64_t n_rays, int64_t slots_per_ray, svox2_stream_t stream);
SVOX2_API int svox2_mg_step_finish(const svox2_grid_t* grid, float* density_rms, float* sh_rms,
                     int64_t n_rays_global, const svox2_mg_step_t* params, const uint8_t* send,
                     const uint64_t* peer_density, void* step_ws, size_t step_ws_bytes,
                     int64_t slots_per_ray, svox2_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SVOX2_B200_H */
