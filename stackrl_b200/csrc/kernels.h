// Host-side dispatchers implemented by the .cu files, called from capi.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace srl {

int maxplus_f32(const float* walls, const float* rocks, const float* level,
                float* out, int E, int R, int H, int W, int h, float threshold,
                int variant, cudaStream_t stream);

int microbench_addmax(int variant, int iters, double* host_cells_per_s);

}  // namespace srl
