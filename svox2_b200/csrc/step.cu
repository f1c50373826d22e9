// Fused train step: opt/opt.py's inner loop (opt.py:474-603) as ONE C-ABI call on a compact
// per-step gradient buffer (slots.cuh).
//
//   reference (per batch)                                  here (one call, 11 launches, no host sync)
//   ----------------------------------------------------   -------------------------------------------
//   volume_render_fused   (svox2.py:1072-1142)              clear | trace | shade (+ touched-row bits)
//   inplace_tv_grad       (svox2.py:1731-1766)              tv_mark | rank scan (slot numbering)
//   inplace_tv_color_grad (svox2.py:1768-1820)              seed | scatter -> slot rows
//   optim_density_step / optim_sh_step (svox2.py:1972-2034) TV sigma -> slots | TV SH -> slots
//   (mask -> index list with .item(), svox2.py:2215-2221)   RMSProp over the slot list
//
// What changes against the operator-by-operator path (grid.py) is only WHERE a gradient lives
// during the step: a [touched, W] buffer in row order instead of dense [N,1] / [N,27] `.grad`
// arrays.  Sample set, per-sample arithmetic, TV arithmetic, the two masks (slot flags) and the
// RMSProp formula are the same code; `svox2_train_step_masks` exports the masks as bool[N] for the
// parity tests.  The dense `.grad` tensors are not touched by this entry point.
//
// Cost model per touched row: the old optimiser pair moved 6 sparse 112-byte rows (grad, data, rms
// read + written) + 6 sparse floats; here 4 sparse rows (data, rms) + 2 dense ones (slot read, slot
// zeroing) and the scatter's vector reductions land in a buffer 10x smaller than `.grad`.
#include <cstdlib>

#include "common.cuh"
#include "march.cuh"
#include "slots.cuh"
#include "step.cuh"

namespace svox2b {
namespace {

constexpr int RANK_THREADS = 256;   // words per block of the rank scan
#ifndef SLOT_OPT_THREADS
#define SLOT_OPT_THREADS 128
#endif
constexpr int OPT_THREADS = SLOT_OPT_THREADS;

struct StepWs {
    void* queue;            // sample-record queue of the march kernels
    size_t queue_bytes;
    uint32_t* bits;         // [n_words]
    uint32_t* prefix;       // [n_words]
    uint32_t* block_sums;   // [n_blocks]
    int32_t* n_slots;       // [0] slots of the current step, [1] of the previous step (flags to clear)
    int32_t* slot_row;      // [cap]
    uint8_t* flags;         // [cap, 2]
    float* grad;            // [cap, width]
    int64_t n_words, n_blocks, cap;
    int width;
};

inline size_t up256(size_t x) { return (x + 255) / 256 * 256; }

size_t step_ws_bytes(int64_t n_rows, int basis_dim, int64_t n_rays, int64_t slots_per_ray) {
    const int64_t n_words = (n_rows + 31) / 32, n_blocks = (n_words + RANK_THREADS - 1) / RANK_THREADS;
    const int W = slot_width(basis_dim);
    return up256(queue_workspace_bytes(n_rays, slots_per_ray * n_rays)) + up256((size_t)n_words * 4) * 2 +
           up256((size_t)n_blocks * 4) + 256 + up256((size_t)n_rows * 4) + up256((size_t)n_rows * 2) +
           up256((size_t)n_rows * W * 4);
}

bool carve_step_ws(void* workspace, size_t bytes, int64_t n_rows, int basis_dim, int64_t n_rays,
                   int64_t slots_per_ray, StepWs& w) {
    if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u) ||
        bytes < step_ws_bytes(n_rows, basis_dim, n_rays, slots_per_ray))
        return false;
    w.n_words = (n_rows + 31) / 32;
    w.n_blocks = (w.n_words + RANK_THREADS - 1) / RANK_THREADS;
    w.cap = n_rows;
    w.width = slot_width(basis_dim);
    char* p = static_cast<char*>(workspace);
    w.queue = p;
    w.queue_bytes = up256(queue_workspace_bytes(n_rays, slots_per_ray * n_rays));
    p += w.queue_bytes;
    w.bits = reinterpret_cast<uint32_t*>(p);
    p += up256((size_t)w.n_words * 4);
    w.prefix = reinterpret_cast<uint32_t*>(p);
    p += up256((size_t)w.n_words * 4);
    w.block_sums = reinterpret_cast<uint32_t*>(p);
    p += up256((size_t)w.n_blocks * 4);
    w.n_slots = reinterpret_cast<int32_t*>(p);
    p += 256;
    w.slot_row = reinterpret_cast<int32_t*>(p);
    p += up256((size_t)n_rows * 4);
    w.flags = reinterpret_cast<uint8_t*>(p);
    p += up256((size_t)n_rows * 2);
    w.grad = reinterpret_cast<float*>(p);
    return true;
}

SlotMap slot_map(const StepWs& w, int basis_dim) {
    return SlotMap{w.bits, w.prefix, w.grad, w.flags, w.width, 3 * basis_dim};
}

// ---- start of a step: forget the previous step's touched set --------------------------------
__global__ void __launch_bounds__(256)
step_clear_kernel(uint32_t* __restrict__ bits, int64_t n_words, uint16_t* __restrict__ flags,
                  int32_t* __restrict__ n_slots, int* __restrict__ q_counter, int* __restrict__ q_overflow,
                  int* __restrict__ q_sticky, float* __restrict__ mse_out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n_prev = n_slots[0];
    for (int64_t i = t0; i < n_words; i += stride) bits[i] = 0u;
    for (int64_t i = t0; i < n_prev; i += stride) flags[i] = 0;
    if (t0 == 0) {  // sample-record queue header: fold the last overflow into the sticky count
        if (*q_overflow) *q_sticky = (*q_sticky < 0 ? 0 : *q_sticky) + 1;
        *q_counter = 0;
        *q_overflow = 0;
        if (mse_out) *mse_out = 0.f;
    }
}

// Slot rows consumed by the previous step's optimiser are zeroed HERE, on the side stream while the
// latency-bound trace kernel runs, instead of by the optimiser behind its own reads (a dense
// 112-byte store per slot taken off the critical path).
#ifndef SLOT_ZERO_IN_OPT
#define SLOT_ZERO_IN_OPT 1   // A/B (profiles/r2_ab_variants.txt): zeroing on the side stream gained nothing
#endif
__global__ void __launch_bounds__(256)
zero_slots_kernel(float4* __restrict__ grad, int width4, const int32_t* __restrict__ n_slots) {
    const int64_t n4 = (int64_t)n_slots[0] * width4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride)
        grad[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ---- rank scan: slot numbering in row order ---------------------------------------------------
__global__ void __launch_bounds__(RANK_THREADS)
rank_count_kernel(const uint32_t* __restrict__ bits, int64_t n_words, uint32_t* __restrict__ block_sums) {
    const int64_t w = (int64_t)blockIdx.x * RANK_THREADS + threadIdx.x;
    unsigned c = w < n_words ? __popc(__ldcg(bits + w)) : 0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(FULL, c, off);
    __shared__ unsigned s[RANK_THREADS / 32];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
#pragma unroll
        for (int i = 0; i < RANK_THREADS / 32; ++i) t += s[i];
        block_sums[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(RANK_THREADS)
rank_scan_kernel(const uint32_t* __restrict__ bits, int64_t n_words,
                 const uint32_t* __restrict__ block_sums, uint32_t* __restrict__ prefix,
                 int32_t* __restrict__ slot_row, int32_t* __restrict__ n_slots) {
    __shared__ unsigned s_warp[RANK_THREADS / 32];
    __shared__ unsigned s_base;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // offset of this block = sum of the counts of all earlier blocks (a few hundred values)
    unsigned part = 0;
    for (int i = threadIdx.x; i < (int)blockIdx.x; i += RANK_THREADS) part += __ldcg(block_sums + i);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(FULL, part, off);
    if (lane == 0) s_warp[wid] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
#pragma unroll
        for (int i = 0; i < RANK_THREADS / 32; ++i) t += s_warp[i];
        s_base = t;
    }
    __syncthreads();
    const unsigned base = s_base;
    __syncthreads();
    const int64_t w = (int64_t)blockIdx.x * RANK_THREADS + threadIdx.x;
    const uint32_t word = w < n_words ? __ldcg(bits + w) : 0u;
    const unsigned c = __popc(word);
    unsigned incl = c;  // inclusive scan inside the warp, then across the block's warps
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned o = __shfl_up_sync(FULL, incl, off);
        if (lane >= off) incl += o;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    unsigned before = 0;
#pragma unroll
    for (int i = 0; i < RANK_THREADS / 32; ++i) before += (i < wid) ? s_warp[i] : 0u;
    const unsigned excl = base + before + incl - c;
    if (w < n_words) {
        prefix[w] = excl;
        uint32_t m = word;
        unsigned k = excl;
        while (m) {  // the list of touched rows, ascending
            slot_row[k++] = (int32_t)(w * 32 + (__ffs(m) - 1));
            m &= m - 1;
        }
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == RANK_THREADS - 1) {
        n_slots[1] = n_slots[0];
        n_slots[0] = (int32_t)(excl + c);
    }
}

// ---- RMSProp over the slot list (optim_kernel.cu:15-25; same arithmetic as optim.cu) -----------
struct SlotOptArgs {
    float* __restrict__ sh_data;
    float* __restrict__ sh_rms;
    int64_t sh_stride;
    int C;                       // SH channels
    float* __restrict__ density;
    float* __restrict__ density_rms;
    int64_t density_stride;
    const int32_t* __restrict__ slot_row;
    const uint8_t* __restrict__ flags;
    float* __restrict__ grad;    // [n_slots, width]; zeroed behind the read
    const int32_t* __restrict__ n_slots;
    int width;
    float beta, lr_sh, lr_density, eps, minval;
};

// Ship mode keeps density_data replicated: after the RMSProp, the new density of every updated row
// goes into the peers' copies.  send[row] = 0 local only, 1 every peer, 2 every peer but
// `skip_peer` (a co-owner applying the identical update).  thread = slot: slots are in row order,
// so runs of updated rows leave the SM as coalesced NVLink writes (publishing from inside the
// optimiser -- one lane in seven, 4 bytes each -- doubled its run time on 8 GPUs).
struct SlotPublishArgs {
    const float* __restrict__ density;
    int64_t density_stride;
    const int32_t* __restrict__ slot_row;
    const uint8_t* __restrict__ flags;
    const int32_t* __restrict__ n_slots;
    const uint8_t* __restrict__ send;
    float* peer_density[SVOX2_MAX_PEERS];
    int n_peers, self, skip_peer;
};
__global__ void __launch_bounds__(256)
slot_publish_kernel(SlotPublishArgs a) {
    const int64_t n = *a.n_slots;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (!__ldcg(a.flags + 2 * i)) continue;           // density of this row was not updated
        const int64_t row = __ldg(a.slot_row + i);
        const int mode = __ldg(a.send + row);
        if (mode == 0) continue;
        const float v = a.density[row * a.density_stride];
        for (int p = 0; p < a.n_peers; ++p)
            if (p != a.self && !(mode == 2 && p == a.skip_peer)) a.peer_density[p][row * a.density_stride] = v;
    }
}



__device__ __forceinline__ void rmsprop_update(float& d, float& rms, float g, float beta, float lr,
                                               float eps, float minval) {
    const float g2 = __fmul_rn(g, g);
    rms = (rms == 0.f) ? g2 : __fmaf_rn(beta, __fsub_rn(rms, g2), g2);
    d = fmaxf(__fsub_rn(d, __fdiv_rn(__fmul_rn(lr, g), __fadd_rn(sqrtf(rms), eps))), minval);
}

// A/B on B200 (profiles/r2_ab_variants.txt): 2 passes in flight at ptxas' 64 registers (8 CTAs of 128
// threads = 32 warps per SM) against 1 pass capped at 40 registers (12 CTAs = 48 warps per SM, no
// spills): the latter is 3 us per step faster in two separate runs; 48 / 32 registers spill and lose.
#ifndef SLOT_OPT_PASSES
#define SLOT_OPT_PASSES 1
#endif
#ifndef SLOT_OPT_BLOCKS
#define SLOT_OPT_BLOCKS 12
#endif
#ifndef SLOT_OPT_MINBLOCKS
#define SLOT_OPT_MINBLOCKS 12
#endif
__global__ void __launch_bounds__(OPT_THREADS, SLOT_OPT_MINBLOCKS)
slot_rmsprop_kernel(SlotOptArgs a) {
    constexpr int U = SLOT_OPT_PASSES;
    const int lane = threadIdx.x & 31;
    const int qpr = a.width >> 2;          // lanes per slot row
    const int rpp = WARP / qpr;            // slot rows per warp pass
    const int grp = lane / qpr;
    const int ch0 = (lane - grp * qpr) * 4;
    if (grp >= rpp) return;
    const int64_t n = *a.n_slots;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const bool has_sh = ch0 < a.C;
    const bool has_den = ch0 <= a.C && a.C < ch0 + 4;
    for (int64_t i0 = warp * rpp * U; i0 < n; i0 += nwarps * rpp * U) {
        int64_t slot[U], row[U];
        float4 g[U], d[U], r[U];
        float dd[U], dr[U];
        bool on_sh[U], on_den[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            slot[u] = i0 + u * rpp + grp;
            const bool in = slot[u] < n;
            row[u] = in ? __ldg(a.slot_row + slot[u]) : 0;
            const uint16_t f = in ? __ldcg(reinterpret_cast<const uint16_t*>(a.flags) + slot[u]) : 0;
            on_den[u] = in && has_den && (f & 0x00ffu);
            on_sh[u] = in && has_sh && (f & 0xff00u);
            if (!in) slot[u] = -1;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (slot[u] >= 0) g[u] = __ldcg(reinterpret_cast<const float4*>(a.grad + slot[u] * a.width + ch0));
            if (on_sh[u]) {
                d[u] = *reinterpret_cast<const float4*>(a.sh_data + row[u] * a.sh_stride + ch0);
                r[u] = *reinterpret_cast<const float4*>(a.sh_rms + row[u] * a.sh_stride + ch0);
            }
            if (on_den[u]) {
                dd[u] = a.density[row[u] * a.density_stride];
                dr[u] = a.density_rms[row[u] * a.density_stride];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (slot[u] < 0) continue;
            const float* ge = reinterpret_cast<const float*>(&g[u]);
            if (on_sh[u]) {
                float* de = reinterpret_cast<float*>(&d[u]);
                float* re = reinterpret_cast<float*>(&r[u]);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (ch0 + e < a.C) rmsprop_update(de[e], re[e], ge[e], a.beta, a.lr_sh, a.eps, a.minval);
                *reinterpret_cast<float4*>(a.sh_data + row[u] * a.sh_stride + ch0) = d[u];
                *reinterpret_cast<float4*>(a.sh_rms + row[u] * a.sh_stride + ch0) = r[u];
            }
            if (on_den[u]) {
                const int k = a.C - ch0;  // position of the density gradient inside this lane's quad
                const float gden = k == 0 ? ge[0] : (k == 1 ? ge[1] : (k == 2 ? ge[2] : ge[3]));
                rmsprop_update(dd[u], dr[u], gden, a.beta, a.lr_density, a.eps, a.minval);
                a.density[row[u] * a.density_stride] = dd[u];
                a.density_rms[row[u] * a.density_stride] = dr[u];
            }
#if SLOT_ZERO_IN_OPT
            const float z = dependent_zero(ge[0]);
            *reinterpret_cast<float4*>(a.grad + slot[u] * a.width + ch0) = make_float4(z, z, z, z);
#endif
        }
    }
}

// ---- multi-GPU step: boundary rows in slot space -------------------------------------------------
// rows whose gradient another rank contributes to (co-owned planes, row 0) get a slot whether or
// not this rank touches them
__global__ void __launch_bounds__(256)
mark_rows_kernel(const int64_t* __restrict__ rows, int64_t n, SlotMap sm) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) slot_mark(sm, (int32_t)__ldg(rows + i));
}

// packet row = [W slot floats (SH gradients, density gradient at 3*basis_dim) | density flag | SH flag | 0..],
// the layout of exchange.cu's pack kernel; thread = one float4 of a packet row
__global__ void __launch_bounds__(256)
slot_pack_kernel(const int64_t* __restrict__ rows, int64_t n, SlotMap sm, float* __restrict__ packed,
                 int64_t pack_stride, int clear) {
    const int qpp = (int)(pack_stride >> 2);
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = t / qpp;
    const int q = (int)(t - i * qpp);
    if (i >= n) return;
    const int32_t row = (int32_t)__ldg(rows + i);
    const bool has = (__ldcg(sm.bits + (row >> 5)) >> (row & 31)) & 1u;
    const int slot = has ? slot_of(sm, row) : -1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (slot >= 0) {
        if (q * 4 < sm.width) {
            float4* src = reinterpret_cast<float4*>(sm.grad + (size_t)slot * sm.width + q * 4);
            v = *src;
            if (clear) *src = make_float4(dependent_zero(v.x), dependent_zero(v.x), dependent_zero(v.x), dependent_zero(v.x));
        } else if (q * 4 == sm.width) {
            uint8_t* f = sm.flags + 2 * (size_t)slot;
            v.x = f[0] ? 1.f : 0.f;
            v.y = f[1] ? 1.f : 0.f;
            if (clear) *reinterpret_cast<uint16_t*>(f) = 0;
        }
    }
    *reinterpret_cast<float4*>(packed + i * pack_stride + q * 4) = v;
}

// slot row += packet row; flags |= packet flags.  Every listed row has a slot (mark_rows_kernel ran
// before the rank scan); rows are unique, so plain adds (a + b == b + a: both co-owners of a plane
// end up with bit-identical gradients).
__global__ void __launch_bounds__(256)
slot_unpack_add_kernel(const float* __restrict__ packed, int64_t pack_stride, const int64_t* __restrict__ rows,
                       int64_t n, SlotMap sm) {
    const int qpp = (int)(pack_stride >> 2);
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = t / qpp;
    const int q = (int)(t - i * qpp);
    if (i >= n) return;
    const int32_t row = (int32_t)__ldg(rows + i);
    if (!((__ldcg(sm.bits + (row >> 5)) >> (row & 31)) & 1u)) return;   // cannot happen (marked)
    const int slot = slot_of(sm, row);
    const float4 v = __ldcg(reinterpret_cast<const float4*>(packed + i * pack_stride + q * 4));
    if (q * 4 < sm.width) {
        float4* dst = reinterpret_cast<float4*>(sm.grad + (size_t)slot * sm.width + q * 4);
        float4 d = *dst;
        d.x = __fadd_rn(d.x, v.x); d.y = __fadd_rn(d.y, v.y); d.z = __fadd_rn(d.z, v.z); d.w = __fadd_rn(d.w, v.w);
        *dst = d;
    } else if (q * 4 == sm.width) {
        uint8_t* f = sm.flags + 2 * (size_t)slot;
        if (v.x != 0.f) f[0] = 1;
        if (v.y != 0.f) f[1] = 1;
    }
}

// ---- masks of the last step as bool[N] (tests, bench bookkeeping) -------------------------------
__global__ void __launch_bounds__(256)
export_masks_kernel(const int32_t* __restrict__ slot_row, const uint8_t* __restrict__ flags,
                    const int32_t* __restrict__ n_slots, uint8_t* __restrict__ mask_density,
                    uint8_t* __restrict__ mask_sh) {
    const int64_t n = n_slots[0];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int32_t row = slot_row[i];
        if (mask_density && flags[2 * i]) mask_density[row] = 1;
        if (mask_sh && flags[2 * i + 1]) mask_sh[row] = 1;
    }
}

// ---- fork / join onto a side stream --------------------------------------------------------------
// The TV work of a step is independent of the ray work until the optimiser: the marking pass only
// needs the cleared bit set, the two TV kernels only the slot numbering.  They run on a per-device
// side stream between event pairs (capturable), next to the latency-bound trace / scatter kernels
// that leave most issue slots idle.  SVOX2_B200_STEP_OVERLAP=0 keeps everything on one stream.
struct SideStream {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork[2] = {nullptr, nullptr}, join[2] = {nullptr, nullptr};
    bool ok = false;
};

SideStream* side_stream() {
    static SideStream per_dev[64];
    static const bool enabled = [] {
        const char* e = std::getenv("SVOX2_B200_STEP_OVERLAP");
        return !(e && e[0] == '0');
    }();
    if (!enabled) return nullptr;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    SideStream& s = per_dev[dev];
    if (!s.stream) {
        s.ok = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) == cudaSuccess;
        for (int i = 0; i < 2 && s.ok; ++i)
            s.ok = cudaEventCreateWithFlags(&s.fork[i], cudaEventDisableTiming) == cudaSuccess &&
                   cudaEventCreateWithFlags(&s.join[i], cudaEventDisableTiming) == cudaSuccess;
        if (!s.ok) cudaGetLastError();
    }
    return s.ok ? &s : nullptr;
}

}  // namespace
}  // namespace svox2b

using namespace svox2b;

extern "C" SVOX2_API size_t svox2_train_step_workspace_bytes(int64_t n_rows, int32_t basis_dim,
                                                             int64_t n_rays, int64_t slots_per_ray) {
    if (n_rows < 0 || n_rays < 0 || slots_per_ray < 0) return 0;
    return step_ws_bytes(n_rows, basis_dim, n_rays, slots_per_ray);
}

extern "C" SVOX2_API int svox2_train_step_init(void* workspace, size_t workspace_bytes, int64_t n_rows,
                                               int32_t basis_dim, int64_t n_rays, int64_t slots_per_ray,
                                               svox2_stream_t stream) {
    StepWs w;
    SVOX2_REQUIRE(carve_step_ws(workspace, workspace_bytes, n_rows, basis_dim, n_rays, slots_per_ray, w),
                  "workspace too small or not 256-byte aligned (svox2_train_step_workspace_bytes)");
    cudaStream_t st = (cudaStream_t)stream;
    // queue header (counter / overflow / sticky), bit set, prefix, counters, flags and slot rows: the
    // step relies on `slot rows all zero` and `flags all zero` between steps
    cudaMemsetAsync(w.queue, 0, 256, st);
    const size_t tail = static_cast<char*>(workspace) + step_ws_bytes(n_rows, basis_dim, n_rays, slots_per_ray) -
                        reinterpret_cast<char*>(w.bits);
    cudaMemsetAsync(w.bits, 0, tail, st);
    return check_launch(__func__);
}

extern "C" SVOX2_API int svox2_train_step(
    const svox2_grid_t* grid, float* density_rms, float* sh_rms, const float* origins, const float* dirs,
    int64_t n_rays, const svox2_render_opts_t* opt, const float* rgb_gt, float* rgb_out,
    const svox2_step_params_t* p, void* workspace, size_t workspace_bytes, int64_t slots_per_ray,
    float* mse_out, int64_t* counters, svox2_stream_t stream) {
    if (int e = validate_grid(grid, __func__)) return e;
    SVOX2_REQUIRE(opt && opt->step_size > 0.f, "opt must be non-null with a positive step_size");
    SVOX2_REQUIRE(p, "params must not be null");
    SVOX2_REQUIRE(n_rays > 0 && origins && dirs && rgb_gt && rgb_out, "rays / rgb_gt / rgb_out must be non-null");
    SVOX2_REQUIRE(density_rms && sh_rms, "RMSProp state must not be null");
    SVOX2_REQUIRE(grid->background_nlayers == 0, "the fused step has no MSI background path");
    SVOX2_REQUIRE(p->norm_rays > 0, "norm_rays must be positive");
    const int need = ((3 * grid->basis_dim + 3) / 4) * 4;
    SVOX2_REQUIRE(grid->sh_stride % 4 == 0 && grid->sh_stride >= need &&
                      (reinterpret_cast<uintptr_t>(grid->sh_data) & 15u) == 0 &&
                      (reinterpret_cast<uintptr_t>(sh_rms) & 15u) == 0,
                  "sh_data / sh_rms need 16-byte aligned rows padded to a multiple of 4 floats");
    StepWs w;
    if (!carve_step_ws(workspace, workspace_bytes, grid->n_rows, grid->basis_dim, n_rays, slots_per_ray, w)) {
        set_error("%s: workspace too small or not 256-byte aligned (svox2_train_step_workspace_bytes)", __func__);
        return SVOX2_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const GridView g = make_grid_view(grid);
    const SlotMap sm = slot_map(w, grid->basis_dim);
    const int X = grid->size[0], Y = grid->size[1], Z = grid->size[2];

    SideStream* side = side_stream();
    const bool tv_d = p->tv_density_count > 0 && p->tv_density_scale != 0.f;
    const bool tv_s = p->tv_sh_count > 0 && p->tv_sh_scale != 0.f;
    cudaStream_t sd = side ? side->stream : st;     // side stream (== st: everything in order)
    const bool forked = sd != st;

    // main: clear | trace | shade ........... | seed | ........ scatter | ............ RMSProp
    // side:       | tv_mark ... (wait shade) rank count + scan | TV sigma | TV SH |
    // 1. forget the previous step's touched set, reset the record queue and the loss accumulator
    int *q_counter = nullptr, *q_overflow = nullptr, *q_sticky = nullptr;
    SVOX2_REQUIRE(step_queue_header(w.queue, w.queue_bytes, n_rays, &q_counter, &q_overflow, &q_sticky),
                  "sample-record queue could not be carved");
    step_clear_kernel<<<148 * 4, 256, 0, st>>>(w.bits, w.n_words, reinterpret_cast<uint16_t*>(w.flags), w.n_slots,
                                               q_counter, q_overflow, q_sticky, mse_out);
    if (forked) {
        cudaEventRecord(side->fork[0], st);
        cudaStreamWaitEvent(sd, side->fork[0], 0);
    }
#if !SLOT_ZERO_IN_OPT
    zero_slots_kernel<<<148 * 8, 256, 0, sd>>>(reinterpret_cast<float4*>(w.grad), w.width / 4, w.n_slots);
#endif
    // 2. (side) rows the TV cells of this step can reach
    launch_tv_mark(grid->links, X, Y, Z, p->tv_density_start, tv_d ? p->tv_density_count : 0,
                   opt->last_sample_opaque, p->tv_sh_start, tv_s ? p->tv_sh_count : 0, sm, sd);

    // 3. (main) traversal + colours; the shade kernel enters every touched row into the bit set
    MarchArgs a{};
    a.origins = origins;
    a.dirs = dirs;
    a.n_rays = n_rays;
    a.rgb_out = rgb_out;
    a.grad_in = rgb_gt;
    a.beta_loss = p->beta_loss / (float)p->norm_rays;        // cuvol.cu:1127
    a.sparsity_loss = p->sparsity_loss;
    a.norm_factor = 2.f / (float)(3 * (int)p->norm_rays);    // cuvol.cu:736
    a.fused_seed = 1;
    a.counters = reinterpret_cast<unsigned long long*>(counters);
    a.mse_out = mse_out;
    SVOX2_REQUIRE(launch_step_trace_shade(g, *opt, a, w.queue, w.queue_bytes, sm, st),
                  "sample-record queue could not be carved");

    // 4. (side, after the shade kernel) slot numbering; then the sparse TV of the step's contiguous
    //    cell runs (svox2.py:1751-1759, 1803-1812) into the slot rows
    if (forked) {
        cudaEventRecord(side->fork[1], st);
        cudaStreamWaitEvent(sd, side->fork[1], 0);
    }
    rank_count_kernel<<<(unsigned)w.n_blocks, RANK_THREADS, 0, sd>>>(w.bits, w.n_words, w.block_sums);
    rank_scan_kernel<<<(unsigned)w.n_blocks, RANK_THREADS, 0, sd>>>(w.bits, w.n_words, w.block_sums, w.prefix,
                                                                 w.slot_row, w.n_slots);
    if (forked) cudaEventRecord(side->join[0], sd);
    if (tv_d)
        launch_tv_slots(grid->links, X, Y, Z, grid->density_data, grid->density_stride, p->tv_density_start,
                        p->tv_density_count, p->tv_scale_cells, 0, 1, p->tv_density_scale, 0,
                        opt->last_sample_opaque, sm, sm.sigma_at, 0, sd);
    if (tv_s)
        launch_tv_slots(grid->links, X, Y, Z, grid->sh_data, grid->sh_stride, p->tv_sh_start, p->tv_sh_count,
                        p->tv_scale_cells, 0, 3 * grid->basis_dim, p->tv_sh_scale, 1, 0, sm, 0, 1, sd);
    if (forked) cudaEventRecord(side->join[1], sd);

    // 5. (main) gradient seed per ray (needs the colours only), then -- once the slots are numbered --
    //    the scatter into the slot rows
    SVOX2_REQUIRE(launch_step_seed(g, *opt, a, w.queue, w.queue_bytes, st),
                  "sample-record queue could not be carved");
    if (forked) cudaStreamWaitEvent(st, side->join[0], 0);
    SVOX2_REQUIRE(launch_step_scatter(g, a, w.queue, w.queue_bytes, sm, st),
                  "sample-record queue could not be carved");
    if (forked) cudaStreamWaitEvent(st, side->join[1], 0);

    // 6. RMSProp over the slot list; slot rows are zeroed behind the read
    SlotOptArgs o{};
    o.sh_data = const_cast<float*>(grid->sh_data);
    o.sh_rms = sh_rms;
    o.sh_stride = grid->sh_stride;
    o.C = 3 * grid->basis_dim;
    o.density = const_cast<float*>(grid->density_data);
    o.density_rms = density_rms;
    o.density_stride = grid->density_stride;
    o.slot_row = w.slot_row;
    o.flags = w.flags;
    o.grad = w.grad;
    o.n_slots = w.n_slots;
    o.width = w.width;
    o.beta = p->rms_beta;
    o.lr_sh = p->lr_sh;
    o.lr_density = p->lr_density;
    o.eps = p->epsilon;
    o.minval = -1e9f;
    slot_rmsprop_kernel<<<148 * SLOT_OPT_BLOCKS, OPT_THREADS, 0, st>>>(o);
    return check_launch(__func__);
}

extern "C" SVOX2_API int svox2_train_step_masks(void* workspace, size_t workspace_bytes, int64_t n_rows,
                                                int32_t basis_dim, int64_t n_rays, int64_t slots_per_ray,
                                                uint8_t* mask_density, uint8_t* mask_sh,
                                                int32_t* n_touched_and_overflow, svox2_stream_t stream) {
    StepWs w;
    SVOX2_REQUIRE(carve_step_ws(workspace, workspace_bytes, n_rows, basis_dim, n_rays, slots_per_ray, w),
                  "workspace too small or not 256-byte aligned (svox2_train_step_workspace_bytes)");
    cudaStream_t st = (cudaStream_t)stream;
    if (mask_density || mask_sh)
        export_masks_kernel<<<148 * 4, 256, 0, st>>>(w.slot_row, w.flags, w.n_slots, mask_density, mask_sh);
    if (n_touched_and_overflow) {  // device int32[3]: slots of the last step, queue overflow flag, sticky count
        cudaMemcpyAsync(n_touched_and_overflow, w.n_slots, 4, cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(n_touched_and_overflow + 1, static_cast<char*>(w.queue) + 128, 4, cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(n_touched_and_overflow + 2, static_cast<char*>(w.queue) + 192, 4, cudaMemcpyDeviceToDevice, st);
    }
    return check_launch(__func__);
}

// =====================================================================================================
// Multi-GPU step on the compact gradient buffer (svox2_b200/slab.py).  The grid is sharded into
// x-slabs; the phases below are the single-GPU step above cut at the points where the ranks exchange
// per-ray scalars / boundary planes (the exchanges themselves are peer stores + barriers driven by
// the host side):
//   begin    clear | TV marking + boundary-row marking | traversal (slab: all rays, own slab's
//            records; ship: own rays, records shipped to the owning slabs)
//   shade    colours of this slab's records (+ touched-row bits)
//   backward slot numbering | seed | scatter -> slots | TV -> slots
//   (slot_pack / slot_unpack_add: boundary planes and the row-0 packet in slot space)
//   finish   RMSProp over the slot list (ship: updated density rows stored into every peer's copy)
// =====================================================================================================
namespace {

struct MgCtx {
    StepWs w;
    SlotMap sm;
    QueueRef q;
};

bool mg_ctx(const svox2_grid_t* grid, int64_t n_rays_global, void* step_ws, size_t step_ws_bytes,
            int64_t slots_per_ray, void* ship_ws, size_t ship_ws_bytes, int32_t n_slabs, int64_t region_cap,
            MgCtx& c) {
    if (!carve_step_ws(step_ws, step_ws_bytes, grid->n_rows, grid->basis_dim, n_rays_global, slots_per_ray, c.w))
        return false;
    c.sm = slot_map(c.w, grid->basis_dim);
    if (ship_ws) c.q = QueueRef{ship_ws, ship_ws_bytes, n_rays_global, n_slabs, region_cap};
    else c.q = QueueRef{c.w.queue, c.w.queue_bytes, n_rays_global, 0, 0};
    return true;
}

}  // namespace

extern "C" SVOX2_API int svox2_mg_step_begin(
    const svox2_grid_t* grid, const float* origins, const float* dirs, int64_t n_rays_global,
    const svox2_render_opts_t* opt, const svox2_mg_step_t* p, const int64_t* boundary_rows, int64_t n_boundary,
    void* step_ws, size_t step_ws_bytes, int64_t slots_per_ray, float* logT_local,
    const int32_t* slab_bounds, const uint64_t* peer_ship_ws, size_t ship_ws_bytes, int64_t region_cap,
    svox2_stream_t stream) {
    if (int e = validate_grid(grid, __func__)) return e;
    SVOX2_REQUIRE(opt && opt->step_size > 0.f && p, "opt / params must not be null");
    SVOX2_REQUIRE(n_rays_global > 0 && origins && dirs, "rays must be non-null");
    SVOX2_REQUIRE(grid->background_nlayers == 0, "the slab step has no MSI background path");
    SVOX2_REQUIRE(p->n_slabs >= 1 && p->n_slabs <= SVOX2_MAX_PEERS && p->rank >= 0 && p->rank < p->n_slabs,
                  "bad slab arguments");
    SVOX2_REQUIRE(p->ship ? (peer_ship_ws && slab_bounds && n_rays_global % p->n_slabs == 0) : logT_local != nullptr,
                  "ship mode needs the peers' workspaces and a batch divisible by the ranks; slab mode needs logT_local");
    MgCtx c;
    void* my_ship = p->ship ? reinterpret_cast<void*>(peer_ship_ws[p->rank]) : nullptr;
    SVOX2_REQUIRE(mg_ctx(grid, n_rays_global, step_ws, step_ws_bytes, slots_per_ray, my_ship, ship_ws_bytes,
                         p->n_slabs, region_cap, c),
                  "workspace too small or not 256-byte aligned (svox2_train_step_workspace_bytes)");
    cudaStream_t st = (cudaStream_t)stream;
    const GridView g = make_grid_view(grid);
    const int X = grid->size[0], Y = grid->size[1], Z = grid->size[2];
    int *q_counter = nullptr, *q_overflow = nullptr, *q_sticky = nullptr;
    SVOX2_REQUIRE(step_queue_header(c.w.queue, c.w.queue_bytes, n_rays_global, &q_counter, &q_overflow, &q_sticky),
                  "sample-record queue could not be carved");
    step_clear_kernel<<<148 * 4, 256, 0, st>>>(c.w.bits, c.w.n_words, reinterpret_cast<uint16_t*>(c.w.flags),
                                               c.w.n_slots, q_counter, q_overflow, q_sticky, nullptr);
    // rows the TV cells of this rank's pieces of the step's runs can reach, and the boundary rows:
    // on the side stream, next to the latency-bound traversal
    SideStream* side = side_stream();
    cudaStream_t sd = side ? side->stream : st;
    if (sd != st) {
        cudaEventRecord(side->fork[0], st);
        cudaStreamWaitEvent(sd, side->fork[0], 0);
    }
    for (int k = 0; k < 2; ++k)
        launch_tv_mark(grid->links, X, Y, Z, p->tv_density_start[k],
                       p->tv_density_scale != 0.f ? p->tv_density_count[k] : 0, opt->last_sample_opaque,
                       p->tv_sh_start[k], p->tv_sh_scale != 0.f ? p->tv_sh_count[k] : 0, c.sm, sd);
    if (n_boundary > 0) {
        SVOX2_REQUIRE(boundary_rows, "boundary_rows must not be null");
        mark_rows_kernel<<<div_up(n_boundary, 256), 256, 0, sd>>>(boundary_rows, n_boundary, c.sm);
    }
    if (sd != st) cudaEventRecord(side->join[0], sd);
    MarchArgs a{};
    a.origins = origins;
    a.dirs = dirs;
    if (p->ship) {
        const int64_t per = n_rays_global / p->n_slabs;
        a.n_rays = per;
        SVOX2_REQUIRE(launch_ship_trace(g, *opt, a, per * p->rank, n_rays_global, slab_bounds, p->n_slabs, p->rank,
                                        peer_ship_ws, ship_ws_bytes, region_cap, st),
                      "ship workspace too small or not 256-byte aligned (svox2_ship_workspace_bytes)");
    } else {
        a.n_rays = n_rays_global;
        SVOX2_REQUIRE(launch_slab_trace(g, *opt, a, c.w.queue, c.w.queue_bytes, p->x_lo, p->x_hi, logT_local, st),
                      "sample-record queue could not be carved");
    }
    if (sd != st) cudaStreamWaitEvent(st, side->join[0], 0);
    return check_launch(__func__);
}

extern "C" SVOX2_API int svox2_mg_step_shade(
    const svox2_grid_t* grid, int64_t n_rays_global, const svox2_render_opts_t* opt, const svox2_mg_step_t* p,
    const float* lt_in, float* part, void* step_ws, size_t step_ws_bytes, int64_t slots_per_ray, void* ship_ws,
    size_t ship_ws_bytes, int64_t region_cap, svox2_stream_t stream) {
    if (int e = validate_grid(grid, __func__)) return e;
    SVOX2_REQUIRE(opt && p && part && (p->ship || lt_in), "null argument");
    MgCtx c;
    SVOX2_REQUIRE(mg_ctx(grid, n_rays_global, step_ws, step_ws_bytes, slots_per_ray, p->ship ? ship_ws : nullptr,
                         ship_ws_bytes, p->n_slabs, region_cap, c),
                  "workspace too small or not 256-byte aligned");
    SVOX2_REQUIRE(launch_mg_shade(make_grid_view(grid), *opt, c.q, lt_in, part, c.sm, (cudaStream_t)stream),
                  "sample-record queue could not be carved / SH rows not 16-byte aligned");
    return check_launch(__func__);
}

extern "C" SVOX2_API int svox2_mg_step_backward(
    const svox2_grid_t* grid, int64_t n_rays_global, const svox2_render_opts_t* opt, const svox2_mg_step_t* p,
    const float* lt_in, const float* gout, const float* accum_start, void* step_ws, size_t step_ws_bytes,
    int64_t slots_per_ray, void* ship_ws, size_t ship_ws_bytes, int64_t region_cap, svox2_stream_t stream) {
    if (int e = validate_grid(grid, __func__)) return e;
    SVOX2_REQUIRE(opt && p && gout && accum_start && (p->ship || lt_in), "null argument");
    MgCtx c;
    SVOX2_REQUIRE(mg_ctx(grid, n_rays_global, step_ws, step_ws_bytes, slots_per_ray, p->ship ? ship_ws : nullptr,
                         ship_ws_bytes, p->n_slabs, region_cap, c),
                  "workspace too small or not 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const GridView g = make_grid_view(grid);
    const int X = grid->size[0], Y = grid->size[1], Z = grid->size[2];
    const StepWs& w = c.w;
    rank_count_kernel<<<(unsigned)w.n_blocks, RANK_THREADS, 0, st>>>(w.bits, w.n_words, w.block_sums);
    rank_scan_kernel<<<(unsigned)w.n_blocks, RANK_THREADS, 0, st>>>(w.bits, w.n_words, w.block_sums, w.prefix,
                                                                 w.slot_row, w.n_slots);
    // the TV kernels only need the slot numbering: they run on the side stream next to seed / scatter
    SideStream* side = side_stream();
    cudaStream_t sd = side ? side->stream : st;
    if (sd != st) {
        cudaEventRecord(side->fork[0], st);
        cudaStreamWaitEvent(sd, side->fork[0], 0);
    }
    for (int k = 0; k < 2; ++k) {
        if (p->tv_density_scale != 0.f && p->tv_density_count[k] > 0)
            launch_tv_slots(grid->links, X, Y, Z, grid->density_data, grid->density_stride, p->tv_density_start[k],
                            p->tv_density_count[k], p->tv_density_scale_cells, 0, 1, p->tv_density_scale, 0,
                            opt->last_sample_opaque, c.sm, c.sm.sigma_at, 0, sd);
        if (p->tv_sh_scale != 0.f && p->tv_sh_count[k] > 0)
            launch_tv_slots(grid->links, X, Y, Z, grid->sh_data, grid->sh_stride, p->tv_sh_start[k], p->tv_sh_count[k],
                            p->tv_sh_scale_cells, 0, 3 * grid->basis_dim, p->tv_sh_scale, 1, 0, c.sm, 0, 1, sd);
    }
    if (sd != st) cudaEventRecord(side->join[0], sd);
    MarchArgs a{};
    a.n_rays = n_rays_global;
    a.sparsity_loss = p->sparsity_loss;
    SVOX2_REQUIRE(launch_mg_seed(*opt, a, c.q, lt_in, gout, accum_start, st), "sample-record queue could not be carved");
    SVOX2_REQUIRE(launch_mg_scatter(g, c.q, c.sm, st), "sample-record queue could not be carved");
    if (sd != st) cudaStreamWaitEvent(st, side->join[0], 0);
    return check_launch(__func__);
}

extern "C" SVOX2_API int svox2_mg_slot_pack(const int64_t* rows, int64_t n_sel, float* packed, int64_t pack_stride,
                                            int32_t clear, void* step_ws, size_t step_ws_bytes, int64_t n_rows,
                                            int32_t basis_dim, int64_t n_rays, int64_t slots_per_ray,
                                            svox2_stream_t stream) {
    SVOX2_REQUIRE(n_sel >= 0, "n_sel must be >= 0");
    if (n_sel == 0) return SVOX2_OK;
    StepWs w;
    SVOX2_REQUIRE(carve_step_ws(step_ws, step_ws_bytes, n_rows, basis_dim, n_rays, slots_per_ray, w),
                  "workspace too small or not 256-byte aligned");
    SVOX2_REQUIRE(rows && packed && pack_stride % 4 == 0 && pack_stride >= w.width + 4 &&
                      (reinterpret_cast<uintptr_t>(packed) & 15u) == 0,
                  "packed rows must be 16-byte aligned and hold the slot row + 4 floats");
    const int64_t threads = n_sel * (pack_stride / 4);
    slot_pack_kernel<<<div_up(threads, 256), 256, 0, (cudaStream_t)stream>>>(rows, n_sel, slot_map(w, basis_dim), packed,
                                                                             pack_stride, clear);
    return check_launch(__func__);
}

extern "C" SVOX2_API int svox2_mg_slot_unpack_add(const float* packed, int64_t pack_stride, const int64_t* rows,
                                                  int64_t n_sel, void* step_ws, size_t step_ws_bytes, int64_t n_rows,
                                                  int32_t basis_dim, int64_t n_rays, int64_t slots_per_ray,
                                                  svox2_stream_t stream) {
    SVOX2_REQUIRE(n_sel >= 0, "n_sel must be >= 0");
    if (n_sel == 0) return SVOX2_OK;
    StepWs w;
    SVOX2_REQUIRE(carve_step_ws(step_ws, step_ws_bytes, n_rows, basis_dim, n_rays, slots_per_ray, w),
                  "workspace too small or not 256-byte aligned");
    SVOX2_REQUIRE(rows && packed && pack_stride % 4 == 0 && pack_stride >= w.width + 4 &&
                      (reinterpret_cast<uintptr_t>(packed) & 15u) == 0,
                  "packed rows must be 16-byte aligned and hold the slot row + 4 floats");
    const int64_t threads = n_sel * (pack_stride / 4);
    slot_unpack_add_kernel<<<div_up(threads, 256), 256, 0, (cudaStream_t)stream>>>(packed, pack_stride, rows, n_sel,
                                                                                   slot_map(w, basis_dim));
    return check_launch(__func__);
}

extern "C" SVOX2_API int svox2_mg_step_finish(
    const svox2_grid_t* grid, float* density_rms, float* sh_rms, int64_t n_rays_global, const svox2_mg_step_t* p,
    const uint8_t* send, const uint64_t* peer_density, void* step_ws, size_t step_ws_bytes, int64_t slots_per_ray,
    svox2_stream_t stream) {
    if (int e = validate_grid(grid, __func__)) return e;
    SVOX2_REQUIRE(p && density_rms && sh_rms, "null argument");
    SVOX2_REQUIRE((send == nullptr) == (peer_density == nullptr), "send and peer_density go together");
    StepWs w;
    SVOX2_REQUIRE(carve_step_ws(step_ws, step_ws_bytes, grid->n_rows, grid->basis_dim, n_rays_global, slots_per_ray, w),
                  "workspace too small or not 256-byte aligned");
    SlotOptArgs o{};
    o.sh_data = const_cast<float*>(grid->sh_data);
    o.sh_rms = sh_rms;
    o.sh_stride = grid->sh_stride;
    o.C = 3 * grid->basis_dim;
    o.density = const_cast<float*>(grid->density_data);
    o.density_rms = density_rms;
    o.density_stride = grid->density_stride;
    o.slot_row = w.slot_row;
    o.flags = w.flags;
    o.grad = w.grad;
    o.n_slots = w.n_slots;
    o.width = w.width;
    o.beta = p->rms_beta;
    o.lr_sh = p->lr_sh;
    o.lr_density = p->lr_density;
    o.eps = p->epsilon;
    o.minval = -1e9f;
    slot_rmsprop_kernel<<<148 * SLOT_OPT_BLOCKS, OPT_THREADS, 0, (cudaStream_t)stream>>>(o);
    if (send) {
        SlotPublishArgs pa{};
        pa.density = grid->density_data;
        pa.density_stride = grid->density_stride;
        pa.slot_row = w.slot_row;
        pa.flags = w.flags;
        pa.n_slots = w.n_slots;
        pa.send = send;
        pa.n_peers = p->n_slabs;
        pa.self = p->rank;
        pa.skip_peer = p->rank - 1;
        for (int k = 0; k < p->n_slabs; ++k) {
            SVOX2_REQUIRE(peer_density[k] != 0, "bad peer pointer");
            pa.peer_density[k] = reinterpret_cast<float*>(peer_density[k]);
        }
        slot_publish_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(pa);
    }
    return check_launch(__func__);
}
