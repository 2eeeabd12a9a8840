// Dispatcher of the TMA-pipelined pass kernels by tile length.
#include "pbk_tma_launch.h"

namespace pbk {

#define PBK_TMA_LENGTHS(X) X(6) X(7) X(8) X(9) X(10)

#define X(l)                                                                                  \
  void tma_info_l##l(TmaInfo* info);                                                          \
  cudaError_t tma_launch_l##l(int mode, const PassArgs& a, const CUtensorMap& tm,             \
                              const float2* d_tables, long long ntiles, int num_sms,          \
                              cudaStream_t st);
PBK_TMA_LENGTHS(X)
#undef X

bool tma_info(int log2L, TmaInfo* info) {
  switch (log2L) {
#define X(l) case l: tma_info_l##l(info); return true;
    PBK_TMA_LENGTHS(X)
#undef X
  }
  return false;
}

cudaError_t tma_launch(int log2L, int mode, const PassArgs& a, const CUtensorMap& tm,
                       const float2* d_tables, long long ntiles, int num_sms, cudaStream_t st) {
  switch (log2L) {
#define X(l) case l: return tma_launch_l##l(mode, a, tm, d_tables, ntiles, num_sms, st);
    PBK_TMA_LENGTHS(X)
#undef X
  }
  return cudaErrorInvalidValue;
}

}  // namespace pbk
