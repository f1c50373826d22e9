// Launchers shared by the fused train step (step.cu): TV kernels writing into the compact slot
// rows (loss.cu) -- see slots.cuh for the data structure.
#pragma once
#include "common.cuh"
#include "slots.cuh"

namespace svox2b {

void launch_tv_mark(const int32_t* links, int X, int Y, int Z, int64_t start_a, int64_t n_a,
                    int ignore_last_z_a, int64_t start_b, int64_t n_b, const SlotMap& sm,
                    cudaStream_t st);
void launch_tv_slots(const int32_t* links, int X, int Y, int Z, const float* data, int64_t data_stride,
                     int64_t cell_start, int64_t n_cells, int64_t scale_cells, int start_dim, int end_dim,
                     float scale, int ignore_edge, int ignore_last_z, const SlotMap& sm, int ch_off,
                     int which, cudaStream_t st);

}  // namespace svox2b
