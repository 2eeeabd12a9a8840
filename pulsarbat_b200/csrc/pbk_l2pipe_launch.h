// pbk_l2pipe_launch.h -- host interface to the L2-resident pipeline kernel (pbk_l2pipe.cuh).
#pragma once
#include <cuda_runtime.h>

#include "pbk_fft.cuh"

namespace pbk {

struct L2PipeArgs {
  int nblocks;            // 2^l1
  long long tiles_a;      // tiles per block of passes A and C
  long long tiles_b;      // tiles per block of pass B
  unsigned* ticket;       // 1 counter
  unsigned* done;         // [nblocks][2]: completed tiles of pass A / pass B per block
  unsigned* err;          // set when a dependency poll gives up
};

cudaError_t l2pipe_launch_l8_l6(const PassArgs& pa, const PassArgs& pb, const PassArgs& pc,
                                const float2* tab_a, const float2* tab_b, const L2PipeArgs& q,
                                int num_sms, cudaStream_t st);

}  // namespace pbk
