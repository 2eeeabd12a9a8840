// TMA-pipelined MID pass for tiles of 2^10 points (same tile shape as pbk_fast_l10.cu: 1024 points
// x 8 lanes, 64-byte rows); the forward / inverse passes of this length stay on the LDG kernels.
#include "pbk_tma_inst.cuh"

namespace pbk {

using Cfg = FastCfg<4, 16, 16, 1, 2, 256, 2>;

void tma_info_l10(TmaInfo* info) {
  tma_cfg_info<Cfg>(info);
  info->mid_only = true;
}
cudaError_t tma_launch_l10(int mode, const PassArgs& a, const CUtensorMap& tm,
                           const float2* d_tables, long long ntiles, int num_sms,
                           cudaStream_t st) {
  return tma_cfg_launch_mid<Cfg>(mode, a, tm, d_tables, ntiles, num_sms, st);
}

}  // namespace pbk
