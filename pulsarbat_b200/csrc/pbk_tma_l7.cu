// TMA-pipelined pass kernels for tiles of 2^7 points (same tile shape as pbk_fast_l7.cu); one
// translation unit per tile length so that the units build in parallel.
#include "pbk_tma_inst.cuh"

namespace pbk {

using Cfg = FastCfg<8, 16, 1, 1, 4, 128, 4>;

void tma_info_l7(TmaInfo* info) { tma_cfg_info<Cfg>(info); }
cudaError_t tma_launch_l7(int mode, const PassArgs& a, const CUtensorMap& tm,
                           const float2* d_tables, long long ntiles, int num_sms,
                           cudaStream_t st) {
  return tma_cfg_launch<Cfg>(mode, a, tm, d_tables, ntiles, num_sms, st);
}

}  // namespace pbk
