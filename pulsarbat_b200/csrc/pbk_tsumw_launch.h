// pbk_tsumw_launch.h -- host interface to the warp-private time-summing last pass (pbk_tsumw.cuh).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "pbk_fft.cuh"

namespace pbk {

// the tile shape the kernel is instantiated for (2^8 points x 16 lane pairs)
bool tsumw_supported(int log2L, int log2pw);
// `tm`: the pass input as two half-tile boxes of 256 rows x 128 B, CU_TENSOR_MAP_SWIZZLE_128B
cudaError_t tsumw_launch(const PassArgs& a, const CUtensorMap& tm, const float2* d_tables,
                         long long ntiles, int num_sms, cudaStream_t st);

}  // namespace pbk
