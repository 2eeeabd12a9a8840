// Fast pass kernels for tiles of 2^7 points (128 pts x 32 lanes = 32 KiB); one translation unit per tile length so
// that the units build in parallel.
#include "pbk_fast_inst.cuh"

namespace pbk {

using Cfg = FastCfg<8, 16, 1, 1, 4, 128, 4>;

void fast_info_l7(FastInfo* info) { cfg_info<Cfg>(info); }
void fast_tables_l7(float2* dst) { fast_build_tables<Cfg>(dst); }
cudaError_t fast_launch_l7(int mode, const PassArgs& a, const float2* d_tables, long long ntiles,
                            int num_sms, cudaStream_t st) {
  return cfg_launch<Cfg>(mode, a, d_tables, ntiles, num_sms, st);
}

}  // namespace pbk
