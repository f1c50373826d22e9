// Compact per-step gradient buffer ("slots") of the fused train step (step.cu).
//
// Why: the masked RMSProp was the largest kernel of the step -- per touched row it read and
// re-zeroed a 112-byte gradient row scattered in a [N,28] array on top of the data / rms
// read-modify-writes, all isolated rows (~3.4 TB/s).  A training step touches ~10 % of the rows, so
// the gradients of one step are kept in a COMPACT [touched, W] buffer instead:
//
//   bits[N/32]     touched-row bit set, filled by the shade kernel (8 corners of every active
//                  sample) and by a TV marking pass (every stored voxel a TV cell of the step can
//                  reach); 1.7 MB at 13.4 M rows -> L2 resident
//   prefix[N/32]   exclusive popcount scan of bits: slot(row) = prefix[row>>5] + popc(bits below)
//                  (rank query: two 4-byte loads from ~3 MB of tables, no hash, no CAS);
//                  slots come out in ROW ORDER, so the optimiser walks data / rms monotonically
//   grad[slot][W]  W = round4(3*basis_dim + 1): SH gradients + the density gradient in one
//                  16-byte-aligned row (density at index 3*basis_dim) -> one vector reduction
//                  target per row, streamed densely by the optimiser, never re-zeroed sparsely
//   flags[slot][2] the reference's two masks (svox2.py:1109-1113, loss_kernel.cu:232-237): byte 0
//                  = row is in sparse_grad_indexer (density), byte 1 = in sparse_sh_grad_indexer
//
// Invariant between steps: grad is all zero (the optimiser zeroes what it consumed); flags / bits
// of the previous step are cleared by the next step's first kernel.
#pragma once
#include "common.cuh"

namespace svox2b {

struct SlotMap {
    uint32_t* __restrict__ bits;
    const uint32_t* __restrict__ prefix;
    float* __restrict__ grad;
    uint8_t* __restrict__ flags;
    int width;    // floats per slot row (multiple of 4)
    int sigma_at; // index of the density gradient inside a slot row (= 3*basis_dim)
};

#ifdef __CUDACC__
__device__ __forceinline__ void slot_mark(const SlotMap& m, int32_t row) {
    uint32_t* w = m.bits + (row >> 5);
    const uint32_t bit = 1u << (row & 31);
    if (!(__ldcg(w) & bit)) atomicOr(w, bit);
}

// valid after the rank scan, for rows whose bit is set
__device__ __forceinline__ int slot_of(const SlotMap& m, int32_t row) {
    const int w = row >> 5;
    return (int)(__ldg(m.prefix + w) + __popc(__ldcg(m.bits + w) & ((1u << (row & 31)) - 1u)));
}
#endif

inline int slot_width(int basis_dim) { return ((3 * basis_dim + 1 + 3) / 4) * 4; }

}  // namespace svox2b
