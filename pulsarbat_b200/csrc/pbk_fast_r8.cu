// Radix-8 family: <= 64 registers per thread, 512 threads, two CTAs (32 warps) per SM.
#include "pbk_fast_inst.cuh"

namespace pbk {

using R8_L6 = FastCfg<8, 8, 1, 1, 6, 512, 2>;      // 64 pts x 128 lanes = 64 KiB
using R8_L7 = FastCfg<4, 4, 8, 1, 5, 512, 2>;      // 128 pts x 64 lanes
using R8_L8 = FastCfg<4, 8, 8, 1, 4, 512, 2>;      // 256 pts x 32 lanes
using R8_L9 = FastCfg<8, 8, 8, 1, 3, 512, 2>;
using R8_L10 = FastCfg<4, 4, 8, 8, 2, 512, 2>;
using R8_L11 = FastCfg<4, 8, 8, 8, 1, 512, 2>;
using R8_L12 = FastCfg<8, 8, 8, 8, 1, 1024, 1>;

#define PBK_R8_CASES(X) X(6, R8_L6) X(7, R8_L7) X(8, R8_L8) X(9, R8_L9) X(10, R8_L10) X(11, R8_L11) X(12, R8_L12)

bool fast_info_r8(int log2L, FastInfo* info) {
  switch (log2L) {
#define X(l, C) case l: cfg_info<C>(info); return true;
    PBK_R8_CASES(X)
#undef X
  }
  return false;
}

void fast_tables_r8(int log2L, float2* dst) {
  switch (log2L) {
#define X(l, C) case l: fast_build_tables<C>(dst); break;
    PBK_R8_CASES(X)
#undef X
  }
}

cudaError_t fast_launch_r8(int log2L, int mode, const PassArgs& a, const float2* d_tables,
                           long long ntiles, int num_sms, cudaStream_t st) {
  switch (log2L) {
#define X(l, C) case l: return cfg_launch<C>(mode, a, d_tables, ntiles, num_sms, st);
    PBK_R8_CASES(X)
#undef X
  }
  return cudaErrorInvalidValue;
}

}  // namespace pbk
