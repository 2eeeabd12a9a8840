// pbk_tma_launch.h -- host interface to the TMA-pipelined pass kernels (pbk_tma.cuh); the
// instantiations live in pbk_tma_l6.cu .. pbk_tma_l10.cu.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "pbk_fft.cuh"

namespace pbk {

struct TmaInfo {
  int log2pw, log2L;   // lane pairs per tile, tile length
  int box_rows;        // rows per TMA box (<= 256); a tile is 2^log2L / box_rows boxes
  int groups, buffers, threads;
  size_t smem;
  bool tsum_ok;
  bool mid_only;       // only the MID kernel is instantiated for this tile length
};

// false when there is no TMA instantiation for this tile length
bool tma_info(int log2L, TmaInfo* info);
// MID passes that are the whole plan (user input, final epilogue) and raw inputs stay on the
// LDG kernels: the dispatcher only knows scratch-to-scratch MID passes
cudaError_t tma_launch(int log2L, int mode, const PassArgs& a, const CUtensorMap& tm,
                       const float2* d_tables, long long ntiles, int num_sms, cudaStream_t st);

}  // namespace pbk
