// Fast pass kernels for tiles of 2^6 points (64 pts x 64 lanes = 32 KiB, four CTAs per SM); one translation unit per tile length so
// that the units build in parallel.
#include "pbk_fast_inst.cuh"

namespace pbk {

using Cfg = FastCfg<4, 16, 1, 1, 5, 128, 4>;

void fast_info_l6(FastInfo* info) { cfg_info<Cfg>(info); }
void fast_tables_l6(float2* dst) { fast_build_tables<Cfg>(dst); }
cudaError_t fast_launch_l6(int mode, const PassArgs& a, const float2* d_tables, long long ntiles,
                            int num_sms, cudaStream_t st) {
  return cfg_launch<Cfg>(mode, a, d_tables, ntiles, num_sms, st);
}

}  // namespace pbk
