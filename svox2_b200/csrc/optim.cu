// RMSProp / SGD restricted to the sparse gradient mask (reference: optim_kernel.cu).
//
// Pure streaming work: per updated element read data/rms/grad, write data/rms/grad (24 B).
// The reference launches one thread per (row, channel) of the WHOLE array for the bool-mask
// variant (rows*C threads that almost all exit) and therefore converts the mask to an index
// list on the host first (one device sync per step, svox2.py:2215-2221).  Here a warp scans 32
// mask bytes with one coalesced load, ballots, and then sweeps only the touched rows with
// lane = channel (108-byte coalesced rows): cost ~ N/32 warp-steps + touched rows, no sync.
#include "common.cuh"

namespace svox2b {
namespace {

#ifndef OPT_ROWS_IN_FLIGHT
#define OPT_ROWS_IN_FLIGHT 4
#endif
#ifndef OPT_THREADS_N
#define OPT_THREADS_N 64
#endif
constexpr int OPT_THREADS = OPT_THREADS_N;
constexpr int RIF = OPT_ROWS_IN_FLIGHT;

// rmsprop_once, optim_kernel.cu:15-25
__device__ __forceinline__ void rmsprop_once(float* __restrict__ d, float* __restrict__ r,
                                             float* __restrict__ gp, float beta, float lr,
                                             float eps, float minval) {
    const float g = *gp;
    float rms = *r;
    const float g2 = __fmul_rn(g, g);
    rms = (rms == 0.f) ? g2 : __fmaf_rn(beta, __fsub_rn(rms, g2), g2);
    *r = rms;
    *d = fmaxf(__fsub_rn(*d, __fdiv_rn(__fmul_rn(lr, g), __fadd_rn(sqrtf(rms), eps))), minval);
    *gp = 0.f;
}

// sgd_once, optim_kernel.cu:97-103 (data -= lr*grad is contracted to one FMA upstream)
__device__ __forceinline__ void sgd_once(float* __restrict__ d, float* __restrict__ gp, float lr) {
    *d = __fmaf_rn(-lr, *gp, *d);
    *gp = 0.f;
}

struct OptArgs {
    float* __restrict__ data;
    float* __restrict__ rms;   // NULL for SGD
    float* __restrict__ grad;
    int64_t rows;              // rows of the arrays (dense / mask) or number of indices
    int C;
    int64_t stride;
    const uint8_t* __restrict__ mask;
    const int64_t* __restrict__ index;
    float beta, lr, eps, minval, lr_last;
};

template <bool RMS>
__device__ __forceinline__ void update_row(const OptArgs& a, int64_t row, int lane) {
    for (int ch = lane; ch < a.C; ch += WARP) {
        const int64_t off = row * a.stride + ch;
        const float lr = (ch == a.C - 1) ? a.lr_last : a.lr;
        if (RMS) rmsprop_once(a.data + off, a.rms + off, a.grad + off, a.beta, lr, a.eps, a.minval);
        else sgd_once(a.data + off, a.grad + off, lr);
    }
}

// bool-mask variant, C > 1: a warp owns 32 consecutive rows
template <bool RMS>
__global__ void __launch_bounds__(OPT_THREADS)
masked_rows_kernel(OptArgs a) {
    const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t row0 = warp_id * WARP;
    if (row0 >= a.rows) return;
    const int64_t my_row = row0 + lane;
    const bool on = my_row < a.rows && __ldg(a.mask + my_row) != 0;
    unsigned m = __ballot_sync(FULL, on);
    if (a.C > WARP) {
        while (m) {
            const int r = __ffs(m) - 1;
            m &= m - 1;
            update_row<RMS>(a, row0 + r, lane);
        }
        return;
    }
    // C <= 32: four touched rows in flight per iteration (12 independent loads per lane)
    const bool chan = lane < a.C;
    const float lr = (lane == a.C - 1) ? a.lr_last : a.lr;
    while (m) {
        int64_t off[RIF];
        float d[RIF], r[RIF], g[RIF];
#pragma unroll
        for (int k = 0; k < RIF; ++k) {
            off[k] = -1;
            if (m) {
                off[k] = (row0 + (__ffs(m) - 1)) * a.stride + lane;
                m &= m - 1;
            }
        }
#pragma unroll
        for (int k = 0; k < RIF; ++k)
            if (chan && off[k] >= 0) {
                g[k] = a.grad[off[k]];
                d[k] = a.data[off[k]];
                if (RMS) r[k] = a.rms[off[k]];
            }
#pragma unroll
        for (int k = 0; k < RIF; ++k)
            if (chan && off[k] >= 0) {
                if (RMS) {
                    rmsprop_once(&d[k], &r[k], &g[k], a.beta, lr, a.eps, a.minval);
                    a.rms[off[k]] = r[k];
                } else {
                    sgd_once(&d[k], &g[k], lr);
                }
                a.data[off[k]] = d[k];
                a.grad[off[k]] = 0.f;
            }
    }
}

// bool-mask variant on 16-byte-aligned padded rows (the layout SparseGrid keeps sh_data / sh_rms /
// grad in): lane = one float4 of a row, so a 27-channel row takes 7 lanes and a warp updates
// 4 touched rows per pass, OPT_VEC_PASSES passes in flight.  Persistent warps stride over the
// 32-row chunks, which keeps every resident warp busy (the one-chunk-per-warp kernel spent most
// of its 52 k blocks on chunks with nothing to do) and cuts the instruction count ~4x (the
// scalar kernel was issue-bound: IEEE sqrt + divide per element).
#ifndef OPT_VEC_PASSES
#define OPT_VEC_PASSES 2
#endif
#ifndef OPT_USE_VEC
#define OPT_USE_VEC 1
#endif
#ifndef OPT_VEC_BLOCKS
#define OPT_VEC_BLOCKS 16
#endif
#ifndef OPT_VEC_GRID_CAP
#define OPT_VEC_GRID_CAP 0   // persistent CTAs per SM (0: one CTA per 8 super-chunks, not capped)
#endif
constexpr int OPT_SUPER = 512;  // rows scanned per warp step
template <bool RMS>
__global__ void __launch_bounds__(OPT_THREADS, OPT_VEC_BLOCKS)
masked_rows_vec_kernel(OptArgs a, int qpr) {
    constexpr int U = OPT_VEC_PASSES;
    const int lane = threadIdx.x & 31;
    const int rpp = WARP / qpr;            // rows per pass
    const int grp = lane / qpr;
    const int ch0 = (lane - grp * qpr) * 4;
    const bool lane_on = grp < rpp;
    float lr4[4];
    bool ok[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        ok[e] = ch0 + e < a.C;
        lr4[e] = (ch0 + e == a.C - 1) ? a.lr_last : a.lr;
    }
    // a warp scans 512 mask bytes per step (one 128-bit load per lane): the scan is a chain of
    // dependent global loads per warp, so it must be short
    const int64_t nsuper = (a.rows + OPT_SUPER - 1) / OPT_SUPER;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const bool mask_al = (reinterpret_cast<uintptr_t>(a.mask) & 15u) == 0;
    for (int64_t sc = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; sc < nsuper; sc += nwarps) {
        const int64_t base = sc * OPT_SUPER + lane * 16;
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        if (mask_al && base + 16 <= a.rows) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(a.mask + base));
            w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
        } else {
            for (int k = 0; k < 16 && base + k < a.rows; ++k)
                w[k >> 2] |= (uint32_t)(__ldg(a.mask + base + k) != 0) << (8 * (k & 3));
        }
        if (!__any_sync(FULL, (w[0] | w[1] | w[2] | w[3]) != 0u)) continue;
        // 16 bytes -> 16 bits, then two lanes' halves make the ballot of one 32-row chunk
        uint32_t h = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t nz = __vcmpne4(w[i], 0u) & 0x01010101u;
            h |= (((nz * 0x01020408u) >> 24) & 0xfu) << (4 * i);
        }
#pragma unroll 1
        for (int c = 0; c < OPT_SUPER / WARP; ++c) {
        unsigned m = __shfl_sync(FULL, h, 2 * c) | (__shfl_sync(FULL, h, 2 * c + 1) << 16);
        const int64_t row0 = sc * OPT_SUPER + c * WARP;
        if (qpr == 1) {
            // 4-channel rows (MSI texels): lane = row, adjacent lanes touch adjacent 16-byte rows; the
            // touched rows of a ray are runs of consecutive layers, so most chunks are empty or dense
            if ((m >> lane) & 1u) {
                const int64_t o = (row0 + lane) * a.stride;
                float4 g4 = *reinterpret_cast<const float4*>(a.grad + o);
                float4 d4 = *reinterpret_cast<const float4*>(a.data + o);
                float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (RMS) r4 = *reinterpret_cast<const float4*>(a.rms + o);
                float* gv = reinterpret_cast<float*>(&g4);
                float* dv = reinterpret_cast<float*>(&d4);
                float* rv = reinterpret_cast<float*>(&r4);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (ok[e]) {
                        if (RMS) rmsprop_once(dv + e, rv + e, gv + e, a.beta, lr4[e], a.eps, a.minval);
                        else sgd_once(dv + e, gv + e, lr4[e]);
                    }
                if (RMS) *reinterpret_cast<float4*>(a.rms + o) = r4;
                *reinterpret_cast<float4*>(a.data + o) = d4;
                *reinterpret_cast<float4*>(a.grad + o) = g4;
            }
            continue;
        }
        while (m) {
            // pop the next U*rpp touched rows in address order; group g of pass u keeps row g + u*rpp
            int sel[U];
#pragma unroll
            for (int u = 0; u < U; ++u) sel[u] = -1;
            for (int i = 0; i < U * rpp && m; ++i) {
                const int bit = __ffs(m) - 1;
                m &= m - 1;
#pragma unroll
                for (int u = 0; u < U; ++u) sel[u] = (i == grp + u * rpp) ? bit : sel[u];
            }
            int64_t off[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                off[u] = (lane_on && sel[u] >= 0) ? (row0 + sel[u]) * a.stride + ch0 : -1;
            float4 g[U], d[U], r[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (off[u] >= 0) {
                    g[u] = *reinterpret_cast<const float4*>(a.grad + off[u]);
                    d[u] = *reinterpret_cast<const float4*>(a.data + off[u]);
                    if (RMS) r[u] = *reinterpret_cast<const float4*>(a.rms + off[u]);
                }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (off[u] >= 0) {
                    float* gv = reinterpret_cast<float*>(&g[u]);
                    float* dv = reinterpret_cast<float*>(&d[u]);
                    float* rv = reinterpret_cast<float*>(&r[u]);
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (ok[e]) {  // padding floats keep their values
                            if (RMS) rmsprop_once(dv + e, rv + e, gv + e, a.beta, lr4[e], a.eps, a.minval);
                            else sgd_once(dv + e, gv + e, lr4[e]);
                        }
                    if (RMS) *reinterpret_cast<float4*>(a.rms + off[u]) = r[u];
                    *reinterpret_cast<float4*>(a.data + off[u]) = d[u];
                    *reinterpret_cast<float4*>(a.grad + off[u]) = g[u];
                }
        }
        }
    }
}

// bool-mask variant, C == 1 (density): four rows per thread, one 32-bit mask load
template <bool RMS>
__global__ void __launch_bounds__(OPT_THREADS)
masked_scalar_kernel(OptArgs a) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t row0 = t * 4;
    if (row0 >= a.rows) return;
    uint32_t bits = 0;
    if (row0 + 4 <= a.rows && (reinterpret_cast<uintptr_t>(a.mask) & 3u) == 0) {
        bits = __ldg(reinterpret_cast<const uint32_t*>(a.mask + row0));
    } else {
        for (int k = 0; k < 4 && row0 + k < a.rows; ++k)
            bits |= (uint32_t)(__ldg(a.mask + row0 + k) != 0) << (8 * k);
    }
    if (bits == 0) return;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if ((bits >> (8 * k)) & 0xffu) {
            const int64_t off = (row0 + k) * a.stride;
            if (RMS) rmsprop_once(a.data + off, a.rms + off, a.grad + off, a.beta, a.lr_last, a.eps, a.minval);
            else sgd_once(a.data + off, a.grad + off, a.lr_last);
        }
}

// Ship mode (slab.py): masked RMSProp on a single-channel array that is kept REPLICATED over the
// ranks (density): the owner of a row updates it and stores the new value into every peer's copy
// (NVLink stores, lane = row).  send[row]: 0 = not mine to publish, 1 = to every peer, 2 = to every
// peer but `skip_peer` (a co-owner that applies the identical update to its own copy).
struct PublishPeers {
    float* base[SVOX2_MAX_PEERS];
    int n, self, skip_peer;
};
__global__ void __launch_bounds__(256)
masked_scalar_publish_kernel(OptArgs a, const uint8_t* __restrict__ send, PublishPeers peers) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= a.rows || __ldg(a.mask + row) == 0) return;
    const int64_t off = row * a.stride;
    rmsprop_once(a.data + off, a.rms + off, a.grad + off, a.beta, a.lr_last, a.eps, a.minval);
    const int mode = __ldg(send + row);
    if (mode == 0) return;
    const float v = a.data[off];
    for (int p = 0; p < peers.n; ++p)
        if (p != peers.self && !(mode == 2 && p == peers.skip_peer)) peers.base[p][off] = v;
}

// dense / index variants and single-channel arrays: one thread per (row, channel) element
template <bool RMS, int KIND>
__global__ void __launch_bounds__(OPT_THREADS)
elementwise_kernel(OptArgs a, int64_t Q) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= Q) return;
    const int64_t r = tid / a.C;
    const int ch = (int)(tid - r * a.C);
    int64_t row = r;
    if (KIND == SVOX2_INDEXER_INDEX) row = __ldg(a.index + r);
    if (KIND == SVOX2_INDEXER_MASK && __ldg(a.mask + r) == 0) return;
    const int64_t off = row * a.stride + ch;
    const float lr = (ch == a.C - 1) ? a.lr_last : a.lr;
    if (RMS) rmsprop_once(a.data + off, a.rms + off, a.grad + off, a.beta, lr, a.eps, a.minval);
    else sgd_once(a.data + off, a.grad + off, lr);
}

template <bool RMS>
void launch(const OptArgs& a, int kind, int64_t n_index, cudaStream_t st) {
    if (kind == SVOX2_INDEXER_MASK && a.C > 1) {
        const int qpr = (a.C + 3) / 4;
        const uintptr_t al = reinterpret_cast<uintptr_t>(a.data) | reinterpret_cast<uintptr_t>(a.grad) |
                             (RMS ? reinterpret_cast<uintptr_t>(a.rms) : 0);
        if (OPT_USE_VEC && qpr <= WARP && a.stride % 4 == 0 && a.stride >= qpr * 4 && (al & 15u) == 0) {
            const int64_t supers = (a.rows + OPT_SUPER - 1) / OPT_SUPER;
            const int64_t want = div_up(supers * WARP, OPT_THREADS);
            const int64_t cap = OPT_VEC_GRID_CAP > 0 ? 148 * OPT_VEC_GRID_CAP : want;
            const unsigned blocks = (unsigned)(want < cap ? want : cap);
            masked_rows_vec_kernel<RMS><<<blocks, OPT_THREADS, 0, st>>>(a, qpr);
            return;
        }
        const int64_t warps = (a.rows + WARP - 1) / WARP;
        masked_rows_kernel<RMS><<<div_up(warps * WARP, OPT_THREADS), OPT_THREADS, 0, st>>>(a);
        return;
    }
    if (kind == SVOX2_INDEXER_MASK && a.C == 1) {
        masked_scalar_kernel<RMS><<<div_up((a.rows + 3) / 4, OPT_THREADS), OPT_THREADS, 0, st>>>(a);
        return;
    }
    const int64_t nrow = kind == SVOX2_INDEXER_INDEX ? n_index : a.rows;
    const int64_t Q = nrow * a.C;
    const unsigned blocks = div_up(Q, OPT_THREADS);
    if (kind == SVOX2_INDEXER_DENSE)
        elementwise_kernel<RMS, SVOX2_INDEXER_DENSE><<<blocks, OPT_THREADS, 0, st>>>(a, Q);
    else if (kind == SVOX2_INDEXER_MASK)
        elementwise_kernel<RMS, SVOX2_INDEXER_MASK><<<blocks, OPT_THREADS, 0, st>>>(a, Q);
    else
        elementwise_kernel<RMS, SVOX2_INDEXER_INDEX><<<blocks, OPT_THREADS, 0, st>>>(a, Q);
}

int check_opt_args(const float* data, const float* grad, int64_t rows, int C, int64_t stride,
                   int kind, const uint8_t* mask, const int64_t* index, int64_t n_index,
                   const char* who) {
    if (!data || !grad) {
        set_error("%s: data/grad must not be null", who);
        return SVOX2_ERR_INVALID;
    }
    if (rows < 0 || C < 1 || stride < C) {
        set_error("%s: bad shape rows=%lld channels=%d stride=%lld", who, (long long)rows, C,
                  (long long)stride);
        return SVOX2_ERR_INVALID;
    }
    if (kind == SVOX2_INDEXER_MASK && !mask) {
        set_error("%s: mask indexer is null", who);
        return SVOX2_ERR_INVALID;
    }
    if (kind == SVOX2_INDEXER_INDEX && n_index > 0 && !index) {
        set_error("%s: index indexer is null", who);
        return SVOX2_ERR_INVALID;
    }
    if (kind < 0 || kind > 2) {
        set_error("%s: unknown indexer kind %d", who, kind);
        return SVOX2_ERR_INVALID;
    }
    return SVOX2_OK;
}

}  // namespace
}  // namespace svox2b

using namespace svox2b;

extern "C" SVOX2_API int svox2_rmsprop_step(float* data, float* rms, float* grad, int64_t rows,
                                            int32_t n_channels, int64_t row_stride,
                                            int32_t indexer_kind, const uint8_t* mask,
                                            const int64_t* index, int64_t n_index, float beta,
                                            float lr, float epsilon, float minval, float lr_last,
                                            svox2_stream_t stream) {
    if (int e = check_opt_args(data, grad, rows, n_channels, row_stride, indexer_kind, mask, index,
                               n_index, __func__))
        return e;
    SVOX2_REQUIRE(rms, "rms must not be null");
    if (lr_last < 0.f) lr_last = lr;  // optim_kernel.cu:171
    const int64_t nrow = indexer_kind == SVOX2_INDEXER_INDEX ? n_index : rows;
    if (nrow * n_channels == 0) return SVOX2_OK;
    OptArgs a{data, rms, grad, rows, n_channels, row_stride, mask, index,
              beta, lr, epsilon, minval, lr_last};
    launch<true>(a, indexer_kind, n_index, (cudaStream_t)stream);
    return check_launch(__func__);
}

extern "C" SVOX2_API int svox2_rmsprop_step_publish(float* data, float* rms, float* grad, int64_t rows,
                                                    int64_t row_stride, const uint8_t* mask,
                                                    const uint8_t* send, const uint64_t* peer_base_ptrs,
                                                    int32_t n_peers, int32_t self_rank, int32_t skip_peer,
                                                    float beta, float lr, float epsilon, float minval,
                                                    svox2_stream_t stream) {
    SVOX2_REQUIRE(rows >= 0 && row_stride >= 1, "bad shape");
    if (rows == 0) return SVOX2_OK;
    SVOX2_REQUIRE(data && rms && grad && mask && send && peer_base_ptrs, "null argument");
    SVOX2_REQUIRE(n_peers >= 1 && n_peers <= SVOX2_MAX_PEERS && self_rank >= 0 && self_rank < n_peers,
                  "bad peer arguments");
    PublishPeers peers;
    peers.n = n_peers;
    peers.self = self_rank;
    peers.skip_peer = skip_peer;
    for (int p = 0; p < n_peers; ++p) {
        SVOX2_REQUIRE(peer_base_ptrs[p] != 0, "bad peer pointer");
        peers.base[p] = reinterpret_cast<float*>(peer_base_ptrs[p]);
    }
    OptArgs a{data, rms, grad, rows, 1, row_stride, mask, nullptr, beta, lr, epsilon, minval, lr};
    masked_scalar_publish_kernel<<<div_up(rows, 256), 256, 0, (cudaStream_t)stream>>>(a, send, peers);
    return check_launch(__func__);
}

extern "C" SVOX2_API int svox2_sgd_step(float* data, float* grad, int64_t rows, int32_t n_channels,
                                        int64_t row_stride, int32_t indexer_kind,
                                        const uint8_t* mask, const int64_t* index, int64_t n_index,
                                        float lr, float lr_last, svox2_stream_t stream) {
    if (int e = check_opt_args(data, grad, rows, n_channels, row_stride, indexer_kind, mask, index,
                               n_index, __func__))
        return e;
    if (lr_last < 0.f) lr_last = lr;
    const int64_t nrow = indexer_kind == SVOX2_INDEXER_INDEX ? n_index : rows;
    if (nrow * n_channels == 0) return SVOX2_OK;
    OptArgs a{data, nullptr, grad, rows, n_channels, row_stride, mask, index,
              0.f, lr, 0.f, 0.f, lr_last};
    launch<false>(a, indexer_kind, n_index, (cudaStream_t)stream);
    return check_launch(__func__);
}
