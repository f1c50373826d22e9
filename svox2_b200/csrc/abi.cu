// Error plumbing, argument validation and library-level entry points of the C-ABI.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace svox2b {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

// The reference only printf()s launch errors (cuda_util.cuh:18-21); here they are returned.
int check_launch(const char* what) {
    const cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        set_error("%s: CUDA error: %s", what, cudaGetErrorString(err));
        return SVOX2_ERR_CUDA;
    }
    return SVOX2_OK;
}

// SparseGridSpec::check(), data_spec.hpp:32-51 (device/contiguity are the caller's contract
// at this level; shapes and strides are checked here).
int validate_grid(const svox2_grid_t* g, const char* who) {
    if (!g) {
        set_error("%s: grid is null", who);
        return SVOX2_ERR_INVALID;
    }
    if (!g->density_data || !g->sh_data || !g->links) {
        set_error("%s: density_data, sh_data and links must be non-null device pointers", who);
        return SVOX2_ERR_INVALID;
    }
    if (g->size[0] < 2 || g->size[1] < 2 || g->size[2] < 2) {
        set_error("%s: links must be 3-D with every dimension >= 2", who);
        return SVOX2_ERR_INVALID;
    }
    if (!(g->basis_dim == 1 || g->basis_dim == 4 || g->basis_dim == 9)) {
        set_error("%s: basis_dim must be 1, 4 or 9 (SH basis; 3*basis_dim <= 32 lanes)", who);
        return SVOX2_ERR_INVALID;
    }
    if (g->sh_stride > (1 << 20) || g->density_stride > (1 << 20)) {
        set_error("%s: row strides above 2^20 floats are not supported", who);
        return SVOX2_ERR_INVALID;
    }
    if (g->n_rows < 1) {
        // the march kernel loads row 0 unconditionally for empty corners
        set_error("%s: the grid must store at least one voxel row", who);
        return SVOX2_ERR_INVALID;
    }
    if (g->sh_stride < 3 * g->basis_dim || g->density_stride < 1 || g->n_rows < 0) {
        set_error("%s: bad row strides (sh_stride %lld, density_stride %lld)", who,
                  (long long)g->sh_stride, (long long)g->density_stride);
        return SVOX2_ERR_INVALID;
    }
    if (g->background_nlayers != 0) {
        if (g->background_nlayers < 2 || g->background_reso < 1) {
            set_error("%s: MSI needs >= 2 layers and a positive resolution", who);
            return SVOX2_ERR_INVALID;
        }
        if (g->background_rows > 0 && (!g->background_links || !g->background_data)) {
            set_error("%s: background_links/background_data missing", who);
            return SVOX2_ERR_INVALID;
        }
        if (g->background_data && (reinterpret_cast<uintptr_t>(g->background_data) & 15u)) {
            set_error("%s: background_data must be 16-byte aligned", who);
            return SVOX2_ERR_INVALID;
        }
    }
    return SVOX2_OK;
}

}  // namespace svox2b

extern "C" SVOX2_API const char* svox2_last_error(void) { return svox2b::g_error; }

extern "C" SVOX2_API const char* svox2_version(void) { return "svox2_b200 0.1.0 (sm_100a)"; }

extern "C" SVOX2_API int svox2_cuda_available(void) {
    int n = 0;
    const cudaError_t err = cudaGetDeviceCount(&n);
    if (err != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n > 0 ? 1 : 0;
}
