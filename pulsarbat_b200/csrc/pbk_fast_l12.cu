// Fast pass kernels for tiles of 2^12 points (4096 pts x 4 lanes); one translation unit per tile length so
// that the units build in parallel.
#include "pbk_fast_inst.cuh"

namespace pbk {

using Cfg = FastCfg<16, 16, 16, 1, 1, 512, 1>;

void fast_info_l12(FastInfo* info) { cfg_info<Cfg>(info); }
void fast_tables_l12(float2* dst) { fast_build_tables<Cfg>(dst); }
cudaError_t fast_launch_l12(int mode, const PassArgs& a, const float2* d_tables, long long ntiles,
                            int num_sms, cudaStream_t st) {
  return cfg_launch<Cfg>(mode, a, d_tables, ntiles, num_sms, st);
}

}  // namespace pbk
