// RMSProp / SGD restricted to the sparse gradient mask (reference: optim_kernel.cu).
// Pure streaming kernels: per updated element read data/rms/grad, write data/rms/grad (24 B).
// One kernel per optimiser; the row selector (dense / bool mask / int64 index list,
// optim_kernel.cu:175-215) is a warp-uniform-per-row predicate, channels are the fast axis.
#include "common.cuh"

namespace svox2b {
namespace {

constexpr int OPT_THREADS = 256;

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

struct RowSel {
    int kind;                       // SVOX2_INDEXER_*
    const uint8_t* __restrict__ mask;
    const int64_t* __restrict__ index;
};

// Element tid -> (row, channel); returns false when the element is not to be updated.
__device__ __forceinline__ bool select(const RowSel& s, int64_t tid, int C, int64_t& row, int& ch) {
    const int64_t r = tid / C;
    ch = (int)(tid - r * C);
    if (s.kind == SVOX2_INDEXER_INDEX) {
        row = __ldg(s.index + r);
        return true;
    }
    row = r;
    if (s.kind == SVOX2_INDEXER_MASK) return __ldg(s.mask + r) != 0;
    return true;
}

__global__ void __launch_bounds__(OPT_THREADS)
rmsprop_kernel(float* __restrict__ data, float* __restrict__ rms, float* __restrict__ grad,
               int64_t Q, int C, int64_t stride, RowSel sel, float beta, float lr, float eps,
               float minval, float lr_last) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= Q) return;
    int64_t row;
    int ch;
    if (!select(sel, tid, C, row, ch)) return;
    const int64_t off = row * stride + ch;
    rmsprop_once(data + off, rms + off, grad + off, beta, ch == C - 1 ? lr_last : lr, eps, minval);
}

__global__ void __launch_bounds__(OPT_THREADS)
sgd_kernel(float* __restrict__ data, float* __restrict__ grad, int64_t Q, int C, int64_t stride,
           RowSel sel, float lr, float lr_last) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= Q) return;
    int64_t row;
    int ch;
    if (!select(sel, tid, C, row, ch)) return;
    const int64_t off = row * stride + ch;
    sgd_once(data + off, grad + off, ch == C - 1 ? lr_last : lr);
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
                                  int32_t n_channels, int64_t row_stride, int32_t indexer_kind,
                                  const uint8_t* mask, const int64_t* index, int64_t n_index,
                                  float beta, float lr, float epsilon, float minval, float lr_last,
                                  svox2_stream_t stream) {
    if (int e = check_opt_args(data, grad, rows, n_channels, row_stride, indexer_kind, mask, index,
                               n_index, __func__))
        return e;
    SVOX2_REQUIRE(rms, "rms must not be null");
    if (lr_last < 0.f) lr_last = lr;  // optim_kernel.cu:171
    const int64_t nrow = indexer_kind == SVOX2_INDEXER_INDEX ? n_index : rows;
    const int64_t Q = nrow * n_channels;
    if (Q == 0) return SVOX2_OK;
    RowSel sel{indexer_kind, mask, index};
    rmsprop_kernel<<<div_up(Q, OPT_THREADS), OPT_THREADS, 0, (cudaStream_t)stream>>>(
        data, rms, grad, Q, n_channels, row_stride, sel, beta, lr, epsilon, minval, lr_last);
    return check_launch(__func__);
}

extern "C" SVOX2_API int svox2_sgd_step(float* data, float* grad, int64_t rows, int32_t n_channels,
                              int64_t row_stride, int32_t indexer_kind, const uint8_t* mask,
                              const int64_t* index, int64_t n_index, float lr, float lr_last,
                              svox2_stream_t stream) {
    if (int e = check_opt_args(data, grad, rows, n_channels, row_stride, indexer_kind, mask, index,
                               n_index, __func__))
        return e;
    if (lr_last < 0.f) lr_last = lr;
    const int64_t nrow = indexer_kind == SVOX2_INDEXER_INDEX ? n_index : rows;
    const int64_t Q = nrow * n_channels;
    if (Q == 0) return SVOX2_OK;
    RowSel sel{indexer_kind, mask, index};
    sgd_kernel<<<div_up(Q, OPT_THREADS), OPT_THREADS, 0, (cudaStream_t)stream>>>(
        data, grad, Q, n_channels, row_stride, sel, lr, lr_last);
    return check_launch(__func__);
}
