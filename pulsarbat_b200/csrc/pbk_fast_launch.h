// pbk_fast_launch.h -- host interface to the compile-time-shaped pass kernels (pbk_fast.cuh).
// The instantiations live in pbk_fast_r8.cu / pbk_fast_r16.cu so they build in parallel.
#pragma once
#include <cuda_runtime.h>

#include "pbk_fft.cuh"

namespace pbk {

enum { FAMILY_R8 = 0, FAMILY_R16 = 1,
       FAMILY_R16N = 2 /* half-width 2^8-point tiles, detecting channelizer pass only */ };

struct FastInfo {
  int log2pw;        // lane pairs per tile (log2)
  int tw_count;      // float2 entries of the stage tables
  int threads, minb; // launch shape
  size_t smem;
  bool tsum_ok;      // the final INV pass has a time-summing variant (whole tasks per thread)
};

// false when this family has no instantiation for the tile length
bool fast_info(int family, int log2L, FastInfo* info);
// fills tw_count float2 entries
void fast_tables(int family, int log2L, float2* dst);
cudaError_t fast_launch(int family, int log2L, int mode, const PassArgs& a, const float2* d_tables,
                        long long ntiles, int num_sms, cudaStream_t st);

}  // namespace pbk
