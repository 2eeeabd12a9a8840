// The L2-resident pipeline of the three middle passes (pbk_l2pipe.cuh), instantiated for the level
// shapes of BASELINE configs 2 and 3: 2^8-point second level, 2^6-point third level.
#include "pbk_l2pipe.cuh"

namespace pbk {

cudaError_t l2pipe_launch_l8_l6(const PassArgs& pa, const PassArgs& pb, const PassArgs& pc,
                                const float2* tab_a, const float2* tab_b, const L2PipeArgs& q,
                                int num_sms, cudaStream_t st) {
  using CA = FastCfg<16, 16, 1, 1, 4, 256, 2>;
  using CB = FastCfg<4, 16, 1, 1, 5, 128, 4>;
  return l2pipe_launch<CA, CB>(pa, pb, pc, tab_a, tab_b, q, num_sms, st);
}

}  // namespace pbk
