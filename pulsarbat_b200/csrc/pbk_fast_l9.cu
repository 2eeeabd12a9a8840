// Fast pass kernels for tiles of 2^9 points (512 pts x 16 lanes = 64 KiB); one translation unit per tile length so
// that the units build in parallel.
#include "pbk_fast_inst.cuh"

namespace pbk {

using Cfg = FastCfg<8, 8, 8, 1, 3, 256, 2>;

void fast_info_l9(FastInfo* info) { cfg_info<Cfg>(info); }
void fast_tables_l9(float2* dst) { fast_build_tables<Cfg>(dst); }
cudaError_t fast_launch_l9(int mode, const PassArgs& a, const float2* d_tables, long long ntiles,
                            int num_sms, cudaStream_t st) {
  return cfg_launch<Cfg>(mode, a, d_tables, ntiles, num_sms, st);
}

}  // namespace pbk
