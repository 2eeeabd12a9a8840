// Radix-16 family: <= 128 registers per thread.  Short tiles (L <= 256) are 32-64 KiB with
// 32-64 lanes (256-512 byte chunks); two to four CTAs per SM.
#include "pbk_fast_inst.cuh"

namespace pbk {

using R16_L6 = FastCfg<4, 16, 1, 1, 5, 128, 4>;    // 64 pts x 64 lanes = 32 KiB
using R16_L7 = FastCfg<8, 16, 1, 1, 4, 128, 4>;    // 128 pts x 32 lanes = 32 KiB
using R16_L8 = FastCfg<16, 16, 1, 1, 4, 256, 2>;   // 256 pts x 32 lanes = 64 KiB
using R16_L9 = FastCfg<8, 8, 8, 1, 3, 256, 2>;     // 512 pts x 16 lanes = 64 KiB
using R16_L10 = FastCfg<4, 16, 16, 1, 2, 256, 2>;
using R16_L11 = FastCfg<8, 16, 16, 1, 1, 256, 2>;
using R16_L12 = FastCfg<16, 16, 16, 1, 1, 512, 1>;

#define PBK_R16_CASES(X) X(6, R16_L6) X(7, R16_L7) X(8, R16_L8) X(9, R16_L9) X(10, R16_L10) X(11, R16_L11) X(12, R16_L12)

bool fast_info_r16(int log2L, FastInfo* info) {
  switch (log2L) {
#define X(l, C) case l: cfg_info<C>(info); return true;
    PBK_R16_CASES(X)
#undef X
  }
  return false;
}

void fast_tables_r16(int log2L, float2* dst) {
  switch (log2L) {
#define X(l, C) case l: fast_build_tables<C>(dst); break;
    PBK_R16_CASES(X)
#undef X
  }
}

cudaError_t fast_launch_r16(int log2L, int mode, const PassArgs& a, const float2* d_tables,
                            long long ntiles, int num_sms, cudaStream_t st) {
  switch (log2L) {
#define X(l, C) case l: return cfg_launch<C>(mode, a, d_tables, ntiles, num_sms, st);
    PBK_R16_CASES(X)
#undef X
  }
  return cudaErrorInvalidValue;
}

}  // namespace pbk
