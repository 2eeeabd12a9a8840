// TMA-pipelined pass kernels for tiles of 2^9 points (same tile shape as pbk_fast_l9.cu); one
// translation unit per tile length so that the units build in parallel.
#include "pbk_tma_inst.cuh"

namespace pbk {

using Cfg = FastCfg<8, 8, 8, 1, 3, 256, 2>;

void tma_info_l9(TmaInfo* info) { tma_cfg_info<Cfg>(info); }
cudaError_t tma_launch_l9(int mode, const PassArgs& a, const CUtensorMap& tm,
                           const float2* d_tables, long long ntiles, int num_sms,
                           cudaStream_t st) {
  return tma_cfg_launch<Cfg>(mode, a, tm, d_tables, ntiles, num_sms, st);
}

}  // namespace pbk
