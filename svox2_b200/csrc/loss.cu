// Total-variation kernels: tv (value), tv_grad (dense), tv_grad_sparse, msi_tv_grad_sparse.
// Behaviour follows the reference's loss_kernel.cu (citations per kernel).  The exact-zero
// tests that decide mask bits use the same separately-rounded operations as the reference's
// sm_100 SASS (FMUL by size/256, FADD sums, FFMA chain inside the rsqrt).
#include "common.cuh"
#include "slots.cuh"
#include "step.cuh"

namespace svox2b {
namespace {

constexpr int TV_THREADS = 256;

// Destination of the fused train step (step.cu): gradients go to the compact slot rows, mask bits
// to the slot flags (slots.cuh).  ch_off = column of channel 0 inside a slot row, which = flag byte.
struct SlotDest {
    SlotMap m;
    int ch_off;
    int which;
};

// calculate_ray_scale, loss_kernel.cu:22-52 (the NDC branch is commented out upstream)
__device__ __forceinline__ void ray_scale(int X, int Y, int Z, float* s) {
    s[0] = __fmul_rn((float)X, 1.f / 256.f);
    s[1] = __fmul_rn((float)Y, 1.f / 256.f);
    s[2] = __fmul_rn((float)Z, 1.f / 256.f);
}

__device__ __forceinline__ float inv_delta(float dx, float dy, float dz) {
    return rsqrtf(__fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmaf_rn(dx, dx, 1e-9f))));
}

// tv_kernel, loss_kernel.cu:64-109.  Block-level sum, one atomic per block.
__global__ void __launch_bounds__(TV_THREADS)
tv_kernel(const int32_t* __restrict__ links, int X, int Y, int Z, const float* __restrict__ data,
          int64_t stride, int start_dim, int end_dim, float scale, int64_t Q, int ignore_edge,
          float* __restrict__ out) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float tresult = 0.f;
    if (tid < Q) {
        const int nd = end_dim - start_dim;
        const int idx = (int)(tid % nd) + start_dim;
        const int64_t xyz = tid / nd;
        const int z = (int)(xyz % (Z - 1));
        const int64_t xy = xyz / (Z - 1);
        const int y = (int)(xy % (Y - 1));
        const int x = (int)(xy / (Y - 1));
        const int64_t offx = (int64_t)Y * Z;
        const int32_t* lp = links + (x * offx + (int64_t)y * Z + z);
        const int32_t l000 = __ldg(lp);
        if (!(ignore_edge && l000 == 0)) {
            float sc[3];
            ray_scale(X, Y, Z, sc);
            const int32_t l100 = __ldg(lp + offx), l010 = __ldg(lp + Z), l001 = __ldg(lp + 1);
            const float v000 = l000 >= 0 ? __ldg(data + (int64_t)l000 * stride + idx) : 0.f;
            const float nullv = ignore_edge ? v000 : 0.f;
            const float v100 = l100 >= 0 ? __ldg(data + (int64_t)l100 * stride + idx) : nullv;
            const float v010 = l010 >= 0 ? __ldg(data + (int64_t)l010 * stride + idx) : nullv;
            const float v001 = l001 >= 0 ? __ldg(data + (int64_t)l001 * stride + idx) : nullv;
            const float dx = (v100 - v000) * sc[0];
            const float dy = (v010 - v000) * sc[1];
            const float dz = (v001 - v000) * sc[2];
            tresult = sqrtf(1e-5f + dx * dx + dy * dy + dz * dz);
        }
    }
    // warp shuffle reduce, then one shared slot per warp
    __shared__ float warp_sums[TV_THREADS / 32];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) tresult += __shfl_xor_sync(FULL, tresult, off);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = tresult;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < TV_THREADS / 32 ? warp_sums[threadIdx.x] : 0.f;
#pragma unroll
        for (int off = 4; off > 0; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
        if (threadIdx.x == 0) atomicAdd(out, v * scale);
    }
}

// tv_grad_kernel, loss_kernel.cu:111-176
__global__ void __launch_bounds__(TV_THREADS)
tv_grad_kernel(const int32_t* __restrict__ links, int X, int Y, int Z,
               const float* __restrict__ data, int64_t stride, int start_dim, int end_dim,
               float scale, int64_t Q, int ignore_edge, float* __restrict__ grad, int64_t gstride) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= Q) return;
    const int nd = end_dim - start_dim;
    const int idx = (int)(tid % nd) + start_dim;
    const int64_t xyz = tid / nd;
    const int z = (int)(xyz % (Z - 1));
    const int64_t xy = xyz / (Z - 1);
    const int y = (int)(xy % (Y - 1));
    const int x = (int)(xy / (Y - 1));
    const int64_t offx = (int64_t)Y * Z;
    const int32_t* lp = links + (x * offx + (int64_t)y * Z + z);
    const int32_t l000 = __ldg(lp);
    if (ignore_edge && l000 == 0) return;
    float sc[3];
    ray_scale(X, Y, Z, sc);
    const int32_t l100 = __ldg(lp + offx), l010 = __ldg(lp + Z), l001 = __ldg(lp + 1);
    float v000 = 0.f, v100 = 0.f, v010 = 0.f, v001 = 0.f;
    if (l000 >= 0) v000 = __ldg(data + (int64_t)l000 * stride + idx);
    if (l100 >= 0) v100 = __ldg(data + (int64_t)l100 * stride + idx); else if (ignore_edge) v100 = v000;
    if (l010 >= 0) v010 = __ldg(data + (int64_t)l010 * stride + idx); else if (ignore_edge) v010 = v000;
    if (l001 >= 0) v001 = __ldg(data + (int64_t)l001 * stride + idx); else if (ignore_edge) v001 = v000;
    float dx = __fsub_rn(v100, v000), dy = __fsub_rn(v010, v000), dz = __fsub_rn(v001, v000);
    const float idelta = __fmul_rn(scale, inv_delta(dx, dy, dz));
    dx = __fmul_rn(dx, sc[0]);
    dy = __fmul_rn(dy, sc[1]);
    dz = __fmul_rn(dz, sc[2]);
    if (dx != 0.f && l100 >= 0) atomicAdd(grad + (int64_t)l100 * gstride + idx, __fmul_rn(dx, idelta));
    if (dy != 0.f && l010 >= 0) atomicAdd(grad + (int64_t)l010 * gstride + idx, __fmul_rn(dy, idelta));
    if (dz != 0.f && l001 >= 0) atomicAdd(grad + (int64_t)l001 * gstride + idx, __fmul_rn(dz, idelta));
    if (l000 >= 0)
        atomicAdd(grad + (int64_t)l000 * gstride + idx,
                  __fmul_rn(-__fadd_rn(__fadd_rn(dx, dy), dz), idelta));
}

// tv_grad_sparse_kernel, loss_kernel.cu:178-246
//
// A warp owns 32 consecutive entries of rand_cells.  PROBE: lane j decodes cell j and loads its
// four links (coalesced for the contiguous cell runs opt.py uses); cells that cannot contribute
// (all four links empty, or the reference's early-outs) drop out by ballot, which is ~90% of a
// run through a 10%-occupied grid.  WORK: COOP=true  - the warp sweeps the surviving cells one
// at a time with lane = channel (coalesced 108 B rows; SH TV, 27 channels);
// COOP=false - each lane finishes its own cell (density TV, 1 channel).
struct TvCell {
    int32_t l000, l001, l010, l100;
    bool on;
};

__device__ __forceinline__ TvCell tv_probe(const int32_t* __restrict__ links, int X, int Y, int Z,
                                           const int32_t* __restrict__ rand_cells,
                                           int64_t cell_start, int64_t cell, int64_t n_cells,
                                           int ignore_edge, int ignore_last_z) {
    TvCell c{-1, -1, -1, -1, false};
    if (cell >= n_cells) return c;
    // rand_cells == NULL: the contiguous wrapped run [cell_start, cell_start + n_cells) that
    // SparseGrid._get_rand_cells(contiguous=True) would materialise (svox2.py:2230-2237)
    int xyz;
    if (rand_cells) {
        xyz = __ldg(rand_cells + cell);
    } else {  // wrapped run: one conditional subtraction instead of a 64-bit modulo (X*Y*Z < 2^31)
        const int64_t total = (int64_t)X * Y * Z;
        int64_t lin = cell_start + cell;
        if (lin >= total) lin -= total;
        if (lin >= total) lin %= total;
        xyz = (int)lin;
    }
    const unsigned uxyz = (unsigned)xyz;    // 32-bit unsigned divisions (cell indices are >= 0)
    const int z = (int)(uxyz % (unsigned)Z);
    const unsigned xy = uxyz / (unsigned)Z;
    const int y = (int)(xy % (unsigned)Y);
    const int x = (int)(xy / (unsigned)Y);
    const int64_t offx = (int64_t)Y * Z;
    const int32_t* lp = links + (x * offx + (int64_t)y * Z + z);
    c.l000 = __ldg(lp);
    // out-of-range neighbours read row 0, not "empty" (reference quirk)
    c.l001 = ((z + 1 < Z) && (!ignore_last_z || z != Z - 2)) ? __ldg(lp + 1) : 0;
    c.l010 = y + 1 < Y ? __ldg(lp + Z) : 0;
    c.l100 = x + 1 < X ? __ldg(lp + offx) : 0;
    c.on = !(ignore_edge && c.l000 == 0) && !(ignore_last_z && z == Z - 2) &&
           !(c.l000 < 0 && c.l001 < 0 && c.l010 < 0 && c.l100 < 0);
    return c;
}

template <bool SLOTS>
__device__ __forceinline__ void tv_channel(const TvCell& c, int idx, const float* __restrict__ data,
                                           int64_t stride, const float* sc, float scale,
                                           int ignore_edge, uint8_t* __restrict__ mask_out,
                                           float* __restrict__ grad, int64_t gstride,
                                           const SlotDest& sd) {
    const float v000 = c.l000 >= 0 ? __ldg(data + (int64_t)c.l000 * stride + idx) : 0.f;
    const float nullv = ignore_edge ? v000 : 0.f;
    const float v001 = c.l001 >= 0 ? __ldg(data + (int64_t)c.l001 * stride + idx) : nullv;
    const float v010 = c.l010 >= 0 ? __ldg(data + (int64_t)c.l010 * stride + idx) : nullv;
    const float v100 = c.l100 >= 0 ? __ldg(data + (int64_t)c.l100 * stride + idx) : nullv;
    float dx = __fsub_rn(v100, v000), dy = __fsub_rn(v010, v000), dz = __fsub_rn(v001, v000);
    const float idelta = __fmul_rn(scale, inv_delta(dx, dy, dz));
    dx = __fmul_rn(dx, sc[0]);
    dy = __fmul_rn(dy, sc[1]);
    dz = __fmul_rn(dz, sc[2]);
    const float sm = -__fadd_rn(__fadd_rn(dx, dy), dz);
#define MAYBE_ADD_SET(lnk, val)                                                      \
    if ((lnk) >= 0 && (val) != 0.f) {                                                \
        if (SLOTS) {                                                                 \
            const int64_t slot = slot_of(sd.m, (lnk));                               \
            atomicAdd(sd.m.grad + slot * sd.m.width + sd.ch_off + idx, __fmul_rn((val), idelta)); \
            sd.m.flags[2 * slot + sd.which] = 1;                                     \
        } else {                                                                     \
            atomicAdd(grad + (int64_t)(lnk) * gstride + idx, __fmul_rn((val), idelta)); \
            if (mask_out) mask_out[(lnk)] = 1;                                       \
        }                                                                            \
    }
    MAYBE_ADD_SET(c.l000, sm);
    MAYBE_ADD_SET(c.l001, dz);
    MAYBE_ADD_SET(c.l010, dy);
    MAYBE_ADD_SET(c.l100, dx);
#undef MAYBE_ADD_SET
}

template <bool COOP, bool SLOTS>
__global__ void __launch_bounds__(TV_THREADS)
tv_grad_sparse_kernel(const int32_t* __restrict__ links, int X, int Y, int Z,
                      const float* __restrict__ data, int64_t stride,
                      const int32_t* __restrict__ rand_cells, int64_t cell_start, int start_dim,
                      int end_dim, float scale, int64_t n_cells, int ignore_edge, int ignore_last_z,
                      uint8_t* __restrict__ mask_out, float* __restrict__ grad, int64_t gstride,
                      SlotDest sd) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    if (tid - lane >= n_cells) return;  // whole warp past the end
    const TvCell mine = tv_probe(links, X, Y, Z, rand_cells, cell_start, tid, n_cells, ignore_edge, ignore_last_z);
    float sc[3];
    ray_scale(X, Y, Z, sc);
    if (!COOP) {
        if (mine.on)
            for (int idx = start_dim; idx < end_dim; ++idx)
                tv_channel<SLOTS>(mine, idx, data, stride, sc, scale, ignore_edge, mask_out, grad, gstride, sd);
        return;
    }
    unsigned m = __ballot_sync(FULL, mine.on);
    while (m) {
        const int j = __ffs(m) - 1;
        m &= m - 1;
        TvCell c;
        c.l000 = __shfl_sync(FULL, mine.l000, j);
        c.l001 = __shfl_sync(FULL, mine.l001, j);
        c.l010 = __shfl_sync(FULL, mine.l010, j);
        c.l100 = __shfl_sync(FULL, mine.l100, j);
        c.on = true;
        for (int idx = start_dim + lane; idx < end_dim; idx += WARP)
            tv_channel<SLOTS>(c, idx, data, stride, sc, scale, ignore_edge, mask_out, grad, gstride, sd);
    }
}

// Vector form of the COOP sweep for 16-byte-aligned padded rows: lane = one float4 of a row, so a
// 27-channel cell takes 7 lanes, a warp finishes 4 surviving cells per pass and keeps TV_VEC_PASSES
// passes (32 independent 128-bit loads) in flight.  The scalar sweep walked a warp's surviving
// cells one by one -- up to 32 dependent load->atomic round trips in the warps that cross a blob.
// Per-element arithmetic is unchanged; channels whose value is exactly 0 add 0.0f (the reference
// skips them) and a link's mask byte is set iff one of its channels is non-zero.
#ifndef TV_USE_VEC
#define TV_USE_VEC 1
#endif
#ifndef TV_VEC_PASSES
#define TV_VEC_PASSES 2
#endif
#ifdef TV_VEC_MINBLOCKS      // A/B: register cap through resident CTAs per SM
template <bool SLOTS>
__global__ void __launch_bounds__(TV_THREADS, TV_VEC_MINBLOCKS)
#else
template <bool SLOTS>
__global__ void __launch_bounds__(TV_THREADS)
#endif
tv_grad_sparse_vec_kernel(const int32_t* __restrict__ links, int X, int Y, int Z,
                          const float* __restrict__ data, int64_t stride,
                          const int32_t* __restrict__ rand_cells, int64_t cell_start, int start_dim,
                          int end_dim, float scale, int64_t n_cells, int ignore_edge,
                          int ignore_last_z, uint8_t* __restrict__ mask_out,
                          float* __restrict__ grad, int64_t gstride, int qpr, SlotDest sd) {
    constexpr int U = TV_VEC_PASSES;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    // PROBE, then the block pools its surviving cells in shared memory: blobs are contiguous in
    // the cell run, so without pooling one warp of a block would sweep 32 cells while its
    // neighbours idle
    __shared__ int4 s_cell[TV_THREADS];
    __shared__ int s_n;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    const TvCell mine = tv_probe(links, X, Y, Z, rand_cells, cell_start, tid, n_cells, ignore_edge, ignore_last_z);
    {
        const unsigned m = __ballot_sync(FULL, mine.on);
        int base = 0;
        if (lane == 0 && m) base = atomicAdd(&s_n, __popc(m));
        base = __shfl_sync(FULL, base, 0);
        if (mine.on)
            s_cell[base + __popc(m & ((1u << lane) - 1u))] = make_int4(mine.l000, mine.l001, mine.l010, mine.l100);
    }
    __syncthreads();
    const int n_live = s_n;
    float sc[3];
    ray_scale(X, Y, Z, sc);
    const int cpp = WARP / qpr;  // cells per pass
    const int grp = lane / qpr;
    const int ch0 = start_dim + (lane - grp * qpr) * 4;
    const bool lane_on = grp < cpp;
    bool ok[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) ok[e] = ch0 + e < end_dim;
    const int per_round = U * cpp;
    for (int i0 = (threadIdx.x >> 5) * per_round; i0 < n_live; i0 += (TV_THREADS / WARP) * per_round) {
        int32_t l[U][4];
        bool on[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + u * cpp + grp;
            on[u] = lane_on && i < n_live;
            const int4 c = s_cell[on[u] ? i : 0];
            l[u][0] = c.x; l[u][1] = c.y; l[u][2] = c.z; l[u][3] = c.w;
        }
        float4 v[U][4];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (on[u] && l[u][k] >= 0)
                    v[u][k] = __ldg(reinterpret_cast<const float4*>(data + (int64_t)l[u][k] * stride + ch0));
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!on[u]) continue;
            float out[4][4];  // [link][element]
            bool any[4] = {false, false, false, false};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float v000 = l[u][0] >= 0 ? reinterpret_cast<const float*>(&v[u][0])[e] : 0.f;
                const float nullv = ignore_edge ? v000 : 0.f;
                const float v001 = l[u][1] >= 0 ? reinterpret_cast<const float*>(&v[u][1])[e] : nullv;
                const float v010 = l[u][2] >= 0 ? reinterpret_cast<const float*>(&v[u][2])[e] : nullv;
                const float v100 = l[u][3] >= 0 ? reinterpret_cast<const float*>(&v[u][3])[e] : nullv;
                float dx = __fsub_rn(v100, v000), dy = __fsub_rn(v010, v000), dz = __fsub_rn(v001, v000);
                const float idelta = __fmul_rn(scale, inv_delta(dx, dy, dz));
                dx = __fmul_rn(dx, sc[0]);
                dy = __fmul_rn(dy, sc[1]);
                dz = __fmul_rn(dz, sc[2]);
                const float val[4] = {-__fadd_rn(__fadd_rn(dx, dy), dz), dz, dy, dx};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const bool nz = ok[e] && val[k] != 0.f;
                    out[k][e] = nz ? __fmul_rn(val[k], idelta) : 0.f;
                    any[k] = any[k] || nz;
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (l[u][k] >= 0 && any[k]) {
                    float* gp;
                    if (SLOTS) {
                        const int64_t slot = slot_of(sd.m, l[u][k]);
                        gp = sd.m.grad + slot * sd.m.width + sd.ch_off + ch0;
                        sd.m.flags[2 * slot + sd.which] = 1;
                    } else {
                        gp = grad + (int64_t)l[u][k] * gstride + ch0;
                        if (mask_out) mask_out[l[u][k]] = 1;
                    }
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gp),
                                 "f"(out[k][0]), "f"(out[k][1]), "f"(out[k][2]), "f"(out[k][3])
                                 : "memory");
                }
        }
    }
}

// msi_tv_grad_sparse_kernel, loss_kernel.cu:248-333
__global__ void __launch_bounds__(TV_THREADS)
msi_tv_grad_sparse_kernel(const int32_t* __restrict__ links, int reso,
                          const float* __restrict__ msi, int nlayers, int nchan,
                          const int32_t* __restrict__ rand_cells, float scale, float scale_last,
                          int64_t Q, uint8_t* __restrict__ mask, float* __restrict__ grad) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= Q) return;
    const int LX = 2 * reso, LY = reso;
    const int ch = (int)(tid % nchan);
    const int msi_idx = __ldg(rand_cells + tid / nchan);
    const int z = msi_idx % nlayers;
    const int tmp = msi_idx / nlayers;
    const int y = tmp % LY;
    const int x = tmp / LY;
    const int nx = (x == LX - 1) ? 0 : x + 1;
    const int ny = (y == LY - 1) ? 0 : y + 1;
    const int lnk00 = __ldg(links + x * LY + y);
    const int lnk01 = __ldg(links + x * LY + ny);
    const int lnk10 = __ldg(links + nx * LY + y);
#define MSI_AT(l, zz) (((int64_t)(l) * nlayers + (zz)) * nchan + ch)
    const float v00 = lnk00 >= 0 ? __ldg(msi + MSI_AT(lnk00, z)) : 0.f;
    const float v_nxl = (lnk00 >= 0 && z + 1 < nlayers) ? __ldg(msi + MSI_AT(lnk00, z + 1))
                                                        : ((ch == nchan - 1) ? 0.f : v00);
    const float v01 = lnk01 >= 0 ? __ldg(msi + MSI_AT(lnk01, z)) : 0.f;
    const float v10 = lnk10 >= 0 ? __ldg(msi + MSI_AT(lnk10, z)) : 0.f;
    if (ch == nchan - 1) scale = scale_last;
    float dx = __fsub_rn(v10, v00), dy = __fsub_rn(v01, v00), dz = __fsub_rn(v_nxl, v00);
    const float idelta = __fmul_rn(scale, inv_delta(dx, dy, dz));
    dx = __fmul_rn(dx, __fmul_rn((float)LX, 1.f / 256.f));
    dy = __fmul_rn(dy, __fmul_rn((float)LY, 1.f / 256.f));
    dz = __fmul_rn(dz, __fmul_rn((float)nlayers, 1.f / 256.f));
    const float sm = -__fadd_rn(__fadd_rn(dx, dy), dz);
#define MAYBE_ADD_SET(link, zz, val)                                       \
    if ((link) >= 0 && (val) != 0.f) {                                     \
        atomicAdd(grad + MSI_AT(link, zz), __fmul_rn((val), idelta));      \
        if (mask) mask[(int64_t)(link) * nlayers + (zz)] = 1;              \
    }
    MAYBE_ADD_SET(lnk00, z, sm);
    if (z + 1 < nlayers) { MAYBE_ADD_SET(lnk00, z + 1, dz); }
    MAYBE_ADD_SET(lnk01, z, dy);
    MAYBE_ADD_SET(lnk10, z, dx);
#undef MAYBE_ADD_SET
#undef MSI_AT
}

// lumisphere_tv_grad_sparse_kernel, loss_kernel.cu:335-434: TV of the COLOUR seen from one random
// view direction (+ its difference to a slightly rotated direction, weighted by dir_factor).
// Warp per cell, lane = SH channel; per colour the basis-weighted channel sums are formed with a
// group-wise shuffle loop (the same order for every corner, so equal corners give an exact 0 and
// the mask decisions `val != 0` match the reference's segmented sums).  The reference decodes the
// cell index with (Y-1, Z-1) although the caller draws it from [0, X*Y*Z) (svox2.py:1846,
// 2224-2241) and reads links out of bounds for x >= X-1; such cells are skipped here.
__global__ void __launch_bounds__(TV_THREADS)
lumisphere_tv_grad_sparse_kernel(const int32_t* __restrict__ links, int X, int Y, int Z,
                                 const float* __restrict__ sh, int64_t stride, int B,
                                 const int32_t* __restrict__ rand_cells, int64_t n_cells,
                                 const float* __restrict__ basis_fn,
                                 const float* __restrict__ basis_fn_u, float scale, float dir_factor,
                                 uint8_t* __restrict__ mask_out, float* __restrict__ grad,
                                 int64_t gstride) {
    const int64_t cell = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (cell >= n_cells) return;
    const int lane = threadIdx.x & 31;
    const int D = 3 * B;
    const int xyz = __ldg(rand_cells + cell);
    const int z = xyz % (Z - 1);
    const int xy = xyz / (Z - 1);
    const int y = xy % (Y - 1);
    const int x = xy / (Y - 1);
    if (xyz < 0 || x >= X - 1) return;
    const int64_t offx = (int64_t)Y * Z;
    const int32_t* lp = links + (x * offx + (int64_t)y * Z + z);
    const int32_t l000 = __ldg(lp);
    if (l000 == 0) return;  // sic (loss_kernel.cu:374)
    const int32_t l001 = __ldg(lp + 1), l010 = __ldg(lp + Z), l100 = __ldg(lp + offx);
    float sc[3];
    ray_scale(X, Y, Z, sc);
    const bool on = lane < D;
    const int idx = on ? lane : 0;
    const int cg = idx % B, grp0 = (idx / B) * B;
    const float v000 = (on && l000 >= 0) ? __ldg(sh + (int64_t)l000 * stride + idx) : 0.f;
    const float v001 = on ? (l001 >= 0 ? __ldg(sh + (int64_t)l001 * stride + idx) : v000) : 0.f;
    const float v010 = on ? (l010 >= 0 ? __ldg(sh + (int64_t)l010 * stride + idx) : v000) : 0.f;
    const float v100 = on ? (l100 >= 0 ? __ldg(sh + (int64_t)l100 * stride + idx) : v000) : 0.f;
    const float sv = __ldg(basis_fn + cg), svu = __ldg(basis_fn_u + cg);
    const float a000 = v000 * sv, a001 = v001 * sv, a010 = v010 * sv, a100 = v100 * sv, u000 = v000 * svu;
    float s000 = 0.f, s001 = 0.f, s010 = 0.f, s100 = 0.f, su = 0.f;
    for (int k = 0; k < B; ++k) {  // colour-group sums, identical order for all five
        const int src = grp0 + k;
        s000 += __shfl_sync(FULL, a000, src);
        s001 += __shfl_sync(FULL, a001, src);
        s010 += __shfl_sync(FULL, a010, src);
        s100 += __shfl_sync(FULL, a100, src);
        su += __shfl_sync(FULL, u000, src);
    }
    if (!on) return;
    float dx = __fmul_rn(__fsub_rn(s100, s000), sc[0]);
    float dy = __fmul_rn(__fsub_rn(s010, s000), sc[1]);
    float dz = __fmul_rn(__fsub_rn(s001, s000), sc[2]);
    float du = __fmul_rn(__fsub_rn(su, s000), dir_factor);
    const float idelta = scale * rsqrtf(1e-9f + dx * dx + dy * dy + dz * dz + du * du);
    dx = __fmul_rn(dx, sc[0]);
    dy = __fmul_rn(dy, sc[1]);
    dz = __fmul_rn(dz, sc[2]);
    du = __fmul_rn(du, dir_factor);
    const float sm = -dx * sv - dy * sv - dz * sv + du * (svu - sv);
    const int32_t L[4] = {l000, l001, l010, l100};
    const float val[4] = {sm, __fmul_rn(dz, sv), __fmul_rn(dy, sv), __fmul_rn(dx, sv)};
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (L[q] >= 0 && val[q] != 0.f) {
            atomicAdd(grad + (int64_t)L[q] * gstride + idx, val[q] * idelta);
            if (mask_out) mask_out[L[q]] = 1;
        }
}

// Fused train step: every stored voxel that a TV cell of this step can add to is entered into the
// touched-row set BEFORE the slots are numbered (a superset of the rows the TV kernels will flag:
// those depend on the data, loss_kernel.cu:232-237).  Cells [0, n_a) belong to run A (density TV:
// ignore_edge = 0, ignore_last_z as given), cells [n_a, n_a + n_b) to run B (SH TV: ignore_edge = 1).
__global__ void __launch_bounds__(TV_THREADS)
tv_mark_kernel(const int32_t* __restrict__ links, int X, int Y, int Z, int64_t start_a, int64_t n_a,
               int ignore_last_z_a, int64_t start_b, int64_t n_b, SlotMap sm) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n_a + n_b) return;
    const bool a = tid < n_a;
    const TvCell c = tv_probe(links, X, Y, Z, nullptr, a ? start_a : start_b, a ? tid : tid - n_a,
                              a ? n_a : n_b, a ? 0 : 1, a ? ignore_last_z_a : 0);
    if (!c.on) return;
    if (c.l000 >= 0) slot_mark(sm, c.l000);
    if (c.l001 >= 0) slot_mark(sm, c.l001);
    if (c.l010 >= 0) slot_mark(sm, c.l010);
    if (c.l100 >= 0) slot_mark(sm, c.l100);
}

int check_tv_args(const int32_t* links, int X, int Y, int Z, const float* data, int64_t stride,
                  int start_dim, int end_dim, const char* who) {
    if (!links || !data) {
        set_error("%s: links/data must not be null", who);
        return SVOX2_ERR_INVALID;
    }
    if (X < 2 || Y < 2 || Z < 2) {
        set_error("%s: links must be 3-D with every dimension >= 2", who);
        return SVOX2_ERR_INVALID;
    }
    if (start_dim < 0 || end_dim <= start_dim || end_dim > stride) {
        set_error("%s: bad channel range [%d,%d) for row stride %lld", who, start_dim, end_dim,
                  (long long)stride);
        return SVOX2_ERR_INVALID;
    }
    return SVOX2_OK;
}

}  // namespace

void launch_tv_mark(const int32_t* links, int X, int Y, int Z, int64_t start_a, int64_t n_a,
                    int ignore_last_z_a, int64_t start_b, int64_t n_b, const SlotMap& sm,
                    cudaStream_t st) {
    if (n_a + n_b <= 0) return;
    tv_mark_kernel<<<div_up(n_a + n_b, TV_THREADS), TV_THREADS, 0, st>>>(links, X, Y, Z, start_a, n_a,
                                                                       ignore_last_z_a, start_b, n_b, sm);
}

// tv_grad_sparse over the contiguous wrapped run [cell_start, cell_start + n_cells) into slot rows;
// the same kernels and arithmetic as svox2_tv_grad_sparse_run.
void launch_tv_slots(const int32_t* links, int X, int Y, int Z, const float* data, int64_t data_stride,
                     int64_t cell_start, int64_t n_cells, int64_t scale_cells, int start_dim, int end_dim,
                     float scale, int ignore_edge, int ignore_last_z, const SlotMap& sm, int ch_off,
                     int which, cudaStream_t st) {
    if (n_cells <= 0) return;
    if (scale_cells <= 0) scale_cells = n_cells;
    const float s = scale / (float)(int)scale_cells;
    const int nd = end_dim - start_dim;
    const int qpr = (nd + 3) / 4;
    const SlotDest sd{sm, ch_off, which};
    const unsigned blocks = div_up(n_cells, TV_THREADS);
    if (nd <= 2)
        tv_grad_sparse_kernel<false, true><<<blocks, TV_THREADS, 0, st>>>(
            links, X, Y, Z, data, data_stride, nullptr, cell_start, start_dim, end_dim, s, n_cells,
            ignore_edge, ignore_last_z, nullptr, nullptr, 0, sd);
    else if (qpr <= WARP && start_dim % 4 == 0 && data_stride % 4 == 0 && ch_off % 4 == 0 &&
             start_dim + qpr * 4 <= data_stride && ch_off + start_dim + qpr * 4 <= sm.width &&
             (reinterpret_cast<uintptr_t>(data) & 15u) == 0)
        tv_grad_sparse_vec_kernel<true><<<blocks, TV_THREADS, 0, st>>>(
            links, X, Y, Z, data, data_stride, nullptr, cell_start, start_dim, end_dim, s, n_cells,
            ignore_edge, ignore_last_z, nullptr, nullptr, 0, qpr, sd);
    else
        tv_grad_sparse_kernel<true, true><<<blocks, TV_THREADS, 0, st>>>(
            links, X, Y, Z, data, data_stride, nullptr, cell_start, start_dim, end_dim, s, n_cells,
            ignore_edge, ignore_last_z, nullptr, nullptr, 0, sd);
}

}  // namespace svox2b

using namespace svox2b;

extern "C" SVOX2_API int svox2_tv(const int32_t* links, int32_t X, int32_t Y, int32_t Z, const float* data,
                        int64_t data_stride, int32_t start_dim, int32_t end_dim,
                        int32_t ignore_edge, float ndc_coeffx, float ndc_coeffy, float* out,
                        svox2_stream_t stream) {
    (void)ndc_coeffx; (void)ndc_coeffy;
    if (int e = check_tv_args(links, X, Y, Z, data, data_stride, start_dim, end_dim, __func__)) return e;
    SVOX2_REQUIRE(out, "out must not be null");
    const int nl = (X - 1) * (Y - 1) * (Z - 1);
    const int64_t Q = (int64_t)nl * (end_dim - start_dim);
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(out, 0, sizeof(float), st);
    tv_kernel<<<div_up(Q, TV_THREADS), TV_THREADS, 0, st>>>(links, X, Y, Z, data, data_stride,
                                                            start_dim, end_dim, 1.f / nl, Q,
                                                            ignore_edge, out);
    return check_launch(__func__);
}

extern "C" SVOX2_API int svox2_tv_grad(const int32_t* links, int32_t X, int32_t Y, int32_t Z,
                             const float* data, int64_t data_stride, int32_t start_dim,
                             int32_t end_dim, float scale, int32_t ignore_edge, float ndc_coeffx,
                             float ndc_coeffy, float* grad_data, int64_t grad_stride,
                             svox2_stream_t stream) {
    (void)ndc_coeffx; (void)ndc_coeffy;
    if (int e = check_tv_args(links, X, Y, Z, data, data_stride, start_dim, end_dim, __func__)) return e;
    SVOX2_REQUIRE(grad_data, "grad_data must not be null");
    const int nl = (X - 1) * (Y - 1) * (Z - 1);
    const int64_t Q = (int64_t)nl * (end_dim - start_dim);
    tv_grad_kernel<<<div_up(Q, TV_THREADS), TV_THREADS, 0, (cudaStream_t)stream>>>(
        links, X, Y, Z, data, data_stride, start_dim, end_dim, scale / nl, Q, ignore_edge, grad_data,
        grad_stride > 0 ? grad_stride : data_stride);
    return check_launch(__func__);
}

static int tv_grad_sparse_impl(const int32_t* links, int32_t X, int32_t Y, int32_t Z,
                               const float* data, int64_t data_stride, const int32_t* rand_cells,
                               int64_t cell_start, int64_t n_cells, int64_t scale_cells,
                               uint8_t* mask_out, int32_t start_dim, int32_t end_dim, float scale,
                               int32_t ignore_edge, int32_t ignore_last_z, float* grad_data,
                               int64_t grad_stride, svox2_stream_t stream, const char* who) {
    if (int e = check_tv_args(links, X, Y, Z, data, data_stride, start_dim, end_dim, who)) return e;
    if (!grad_data) {
        set_error("%s: grad_data must not be null", who);
        return SVOX2_ERR_INVALID;
    }
    if (n_cells == 0) return SVOX2_OK;
    if (scale_cells <= 0) scale_cells = n_cells;
    const int nd = end_dim - start_dim;
    const float s = scale / (float)(int)scale_cells;
    const int64_t gs = grad_stride > 0 ? grad_stride : data_stride;
    cudaStream_t st = (cudaStream_t)stream;
#define TV_LAUNCH(COOP)                                                                         \
    tv_grad_sparse_kernel<COOP, false><<<div_up(n_cells, TV_THREADS), TV_THREADS, 0, st>>>(     \
        links, X, Y, Z, data, data_stride, rand_cells, cell_start, start_dim, end_dim, s,       \
        n_cells, ignore_edge, ignore_last_z, mask_out, grad_data, gs, SlotDest{})
    const int qpr = (nd + 3) / 4;
    const uintptr_t al = reinterpret_cast<uintptr_t>(data) | reinterpret_cast<uintptr_t>(grad_data);
    if (nd <= 2) TV_LAUNCH(false);
    else if (TV_USE_VEC && qpr <= WARP && start_dim % 4 == 0 && data_stride % 4 == 0 && gs % 4 == 0 &&
             start_dim + qpr * 4 <= data_stride && start_dim + qpr * 4 <= gs && (al & 15u) == 0)
        tv_grad_sparse_vec_kernel<false><<<div_up(n_cells, TV_THREADS), TV_THREADS, 0, st>>>(
            links, X, Y, Z, data, data_stride, rand_cells, cell_start, start_dim, end_dim, s,
            n_cells, ignore_edge, ignore_last_z, mask_out, grad_data, gs, qpr, SlotDest{});
    else TV_LAUNCH(true);
#undef TV_LAUNCH
    return check_launch(who);
}

extern "C" SVOX2_API int svox2_tv_grad_sparse(const int32_t* links, int32_t X, int32_t Y, int32_t Z,
                                    const float* data, int64_t data_stride,
                                    const int32_t* rand_cells, int64_t n_cells,
                                    int64_t scale_cells, uint8_t* mask_out, int32_t start_dim,
                                    int32_t end_dim, float scale, int32_t ignore_edge,
                                    int32_t ignore_last_z, float ndc_coeffx, float ndc_coeffy,
                                    float* grad_data, int64_t grad_stride, svox2_stream_t stream) {
    (void)ndc_coeffx; (void)ndc_coeffy;
    SVOX2_REQUIRE(n_cells >= 0 && (n_cells == 0 || rand_cells), "rand_cells must not be null");
    return tv_grad_sparse_impl(links, X, Y, Z, data, data_stride, rand_cells, 0, n_cells, scale_cells,
                               mask_out, start_dim, end_dim, scale, ignore_edge, ignore_last_z,
                               grad_data, grad_stride, stream, __func__);
}

extern "C" SVOX2_API int svox2_tv_grad_sparse_run(const int32_t* links, int32_t X, int32_t Y, int32_t Z,
                                        const float* data, int64_t data_stride,
                                        int64_t cell_start, int64_t n_cells, int64_t scale_cells,
                                        uint8_t* mask_out, int32_t start_dim, int32_t end_dim,
                                        float scale, int32_t ignore_edge, int32_t ignore_last_z,
                                        float* grad_data, int64_t grad_stride,
                                        svox2_stream_t stream) {
    SVOX2_REQUIRE(n_cells >= 0 && cell_start >= 0, "bad cell run");
    return tv_grad_sparse_impl(links, X, Y, Z, data, data_stride, nullptr, cell_start, n_cells,
                               scale_cells, mask_out, start_dim, end_dim, scale, ignore_edge,
                               ignore_last_z, grad_data, grad_stride, stream, __func__);
}

extern "C" SVOX2_API int svox2_msi_tv_grad_sparse(const int32_t* background_links, int32_t reso,
                                        const float* msi, int32_t nlayers, int32_t nchannels,
                                        const int32_t* rand_cells, int64_t n_cells,
                                        int64_t scale_cells, uint8_t* mask_out, float scale,
                                        float scale_last, float* grad_msi, svox2_stream_t stream) {
    SVOX2_REQUIRE(background_links && msi && grad_msi, "null argument");
    SVOX2_REQUIRE(reso > 0 && nlayers > 0 && nchannels > 0, "bad MSI shape");
    SVOX2_REQUIRE(n_cells >= 0 && (n_cells == 0 || rand_cells), "rand_cells must not be null");
    if (n_cells == 0) return SVOX2_OK;
    if (scale_cells <= 0) scale_cells = n_cells;
    const int64_t Q = n_cells * nchannels;
    msi_tv_grad_sparse_kernel<<<div_up(Q, TV_THREADS), TV_THREADS, 0, (cudaStream_t)stream>>>(
        background_links, reso, msi, nlayers, nchannels, rand_cells,
        scale / (float)(int)scale_cells, scale_last / (float)(int)scale_cells, Q, mask_out, grad_msi);
    return check_launch(__func__);
}

extern "C" SVOX2_API int svox2_lumisphere_tv_grad_sparse(
    const int32_t* links, int32_t X, int32_t Y, int32_t Z, const float* sh_data, int64_t sh_stride,
    int32_t basis_dim, const int32_t* rand_cells, int64_t n_cells, const float* basis_fn,
    const float* basis_fn_u, float scale, float ndc_coeffx, float ndc_coeffy, float dir_factor,
    uint8_t* mask_out, float* grad_sh, int64_t grad_stride, svox2_stream_t stream) {
    (void)ndc_coeffx; (void)ndc_coeffy;  // the NDC branch of calculate_ray_scale is commented out upstream
    SVOX2_REQUIRE(links && sh_data && grad_sh && basis_fn && basis_fn_u, "null argument");
    SVOX2_REQUIRE(X > 1 && Y > 1 && Z > 1, "grid must be at least 2 voxels along every axis");
    SVOX2_REQUIRE(basis_dim >= 1 && 3 * basis_dim <= WARP, "basis_dim must be in 1..10");
    SVOX2_REQUIRE(sh_stride >= 3 * basis_dim, "bad sh row stride");
    SVOX2_REQUIRE(n_cells >= 0 && (n_cells == 0 || rand_cells), "rand_cells must not be null");
    if (n_cells == 0) return SVOX2_OK;
    lumisphere_tv_grad_sparse_kernel<<<div_up(n_cells * WARP, TV_THREADS), TV_THREADS, 0,
                                       (cudaStream_t)stream>>>(
        links, X, Y, Z, sh_data, sh_stride, basis_dim, rand_cells, n_cells, basis_fn, basis_fn_u,
        scale / (float)n_cells, dir_factor, mask_out, grad_sh,
        grad_stride > 0 ? grad_stride : sh_stride);
    return check_launch(__func__);
}
