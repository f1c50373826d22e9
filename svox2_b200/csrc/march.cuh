// Interface of the lattice-batch march kernel (march.cu) used by the C-ABI entry points.
#pragma once
#include "common.cuh"

namespace svox2b {

struct MarchArgs {
    const float* __restrict__ origins;      // [Q,3] (ignored when rays come from the camera)
    const float* __restrict__ dirs;
    CamView cam;
    int row_begin;
    int64_t n_rays;
    float* __restrict__ rgb_out;            // forward / fused: [Q,3]
    float* __restrict__ logT_out;           // forward: final log-transmittance (-1e3 if stopped);
                                            // backward-only: the backward march's value; or NULL
    const float* __restrict__ grad_in;      // backward: grad_out [Q,3]; fused seed: rgb_gt
    const float* __restrict__ color_cache;  // backward-only: forward colours [Q,3]
    const float* __restrict__ logT_in;      // backward-only: forward log-transmittance (beta loss)
    float* __restrict__ accum_out;          // backward: MSI hand-over, or NULL
    float beta_loss;                        // already divided by norm_rays
    float sparsity_loss;
    float norm_factor;                      // 2/(3*norm_rays)
    int fused_seed;                         // 1: gradient seed is (colour - grad_in)*norm_factor
    unsigned long long* __restrict__ counters;
    int64_t gsh_stride, gd_stride;          // row strides of the gradient arrays
    float4* __restrict__ ccache;            // fused: [Q, ccap] forward->backward colour cache, or NULL
    int ccap;                               // cache slots per ray (overflow samples are re-gathered)
    const int* __restrict__ run_if;         // if set: the kernel is a no-op unless *run_if != 0
    float* __restrict__ mse_out;            // fused train step: += mean((rgb - rgb_gt)^2) of the batch, or NULL
};

enum { LB_MODE_FWD = 0, LB_MODE_BWD = 1, LB_MODE_FUSED = 2 };

bool march_vec_ok(const GridView& g, const GradView* gr, int64_t gsh_stride);
// sample-queue fused path (queue_march.cu)
size_t queue_workspace_bytes(int64_t n_rays, int64_t records);
bool launch_queue_fused(const GridView& g, const Opts& opt, const MarchArgs& a, const GradView& gr,
                        void* workspace, size_t workspace_bytes, const int** overflow_flag,
                        float** ray_logT, float** ray_accum, cudaStream_t st);
// fused train step with the compact gradient buffer (step.cu, slots.cuh)
struct SlotMap;
bool launch_step_trace_shade(const GridView& g, const Opts& opt, const MarchArgs& a, void* workspace,
                             size_t workspace_bytes, const SlotMap& sm, cudaStream_t st);
bool step_queue_header(void* workspace, size_t workspace_bytes, int64_t n_rays, int** counter,
                       int** overflow, int** sticky);
bool launch_step_seed(const GridView& g, const Opts& opt, const MarchArgs& a, void* workspace,
                      size_t workspace_bytes, cudaStream_t st);
bool launch_step_scatter(const GridView& g, const MarchArgs& a, void* workspace, size_t workspace_bytes,
                         const SlotMap& sm, cudaStream_t st);
// slab mode (grid sharded by x-range over GPUs): trace | exchange | shade | exchange | backward
bool launch_slab_trace(const GridView& g, const Opts& opt, const MarchArgs& a, void* workspace,
                       size_t workspace_bytes, int x_lo, int x_hi, float* logT_local,
                       cudaStream_t st);
bool launch_slab_shade(const GridView& g, const Opts& opt, int64_t n_rays, void* workspace,
                       size_t workspace_bytes, const float* lt_in, float* part, cudaStream_t st);
bool launch_slab_backward(const GridView& g, const Opts& opt, const MarchArgs& a, const GradView& gr,
                          void* workspace, size_t workspace_bytes, const float* lt_in,
                          const float* gout, const float* accum_start, cudaStream_t st);
// ship mode (rays sharded for the traversal, records shipped to the owning slab over NVLink)
size_t ship_workspace_bytes(int64_t n_rays_global, int n_slabs, int64_t region_cap);
bool launch_ship_trace(const GridView& g, const Opts& opt, const MarchArgs& a, int64_t ray_offset,
                       int64_t n_rays_global, const int* bounds, int n_slabs, int rank,
                       const uint64_t* peer_ws, size_t workspace_bytes, int64_t region_cap, cudaStream_t st);
bool launch_ship_shade(const GridView& g, int64_t n_rays_global, void* workspace, size_t workspace_bytes,
                       int n_slabs, int64_t region_cap, float* part, cudaStream_t st);
const float* ship_ray_logT(void* workspace, size_t workspace_bytes, int64_t n_rays_global, int n_slabs,
                           int64_t region_cap);
bool launch_ship_backward(const GridView& g, const Opts& opt, const MarchArgs& a, const GradView& gr,
                          void* workspace, size_t workspace_bytes, int n_slabs, int64_t region_cap,
                          const float* gout, const float* accum_start, cudaStream_t st);
// multi-GPU step on the compact gradient buffer (step.cu): where the sample records live -- a plain
// queue workspace (n_slabs == 0: slab mode, every rank traverses all rays) or a ship workspace
struct QueueRef {
    void* ws;
    size_t bytes;
    int64_t n_rays;        // rays the per-ray arrays of the workspace are sized for (the global batch)
    int n_slabs;
    int64_t region_cap;
};
bool launch_mg_shade(const GridView& g, const Opts& opt, const QueueRef& q, const float* lt_in, float* part,
                     const SlotMap& sm, cudaStream_t st);
bool launch_mg_seed(const Opts& opt, const MarchArgs& a, const QueueRef& q, const float* lt_in,
                    const float* gout, const float* accum_start, cudaStream_t st);
bool launch_mg_scatter(const GridView& g, const QueueRef& q, const SlotMap& sm, cudaStream_t st);
void launch_lattice_march(int mode, bool from_cam, const GridView& g, const Opts& opt,
                          const MarchArgs& a, const GradView& gr, cudaStream_t st);

}  // namespace svox2b

namespace svox2b {
// launchers defined in render.cu (MSI background, depth)
void launch_bg_fwd(bool from_cam, const GridView& g, const Opts& opt, const MarchArgs& a,
                   const float* logT, cudaStream_t st, const int* gate = nullptr, int gate_want = 0);
void launch_bg_bwd(const GridView& g, const Opts& opt, const MarchArgs& a, const float* color_cache,
                   const float* logT, const float* accum, const GradView& gr, cudaStream_t st,
                   const int* gate = nullptr, int gate_want = 0);
void launch_depth(bool sigma_thresh_mode, const GridView& g, const Opts& opt, const float* origins,
                  const float* dirs, int64_t n_rays, float thresh, float* out, cudaStream_t st);
}  // namespace svox2b
