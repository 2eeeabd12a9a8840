// Dispatcher of the fast pass kernels: radix-16 stages, <= 128 registers per thread; the
// instantiations live in pbk_fast_l6.cu .. pbk_fast_l12.cu (one unit per tile length).
#include "pbk_fast_launch.h"

namespace pbk {

#define PBK_FAST_LENGTHS(X) X(6) X(7) X(8) X(9) X(10) X(11) X(12)

#define X(l)                                                                                   \
  void fast_info_l##l(FastInfo* info);                                                         \
  void fast_tables_l##l(float2* dst);                                                          \
  cudaError_t fast_launch_l##l(int mode, const PassArgs& a, const float2* d_tables,            \
                               long long ntiles, int num_sms, cudaStream_t st);
PBK_FAST_LENGTHS(X)
#undef X

bool fast_info_r16(int log2L, FastInfo* info) {
  switch (log2L) {
#define X(l) case l: fast_info_l##l(info); return true;
    PBK_FAST_LENGTHS(X)
#undef X
  }
  return false;
}

void fast_tables_r16(int log2L, float2* dst) {
  switch (log2L) {
#define X(l) case l: fast_tables_l##l(dst); break;
    PBK_FAST_LENGTHS(X)
#undef X
  }
}

cudaError_t fast_launch_r16(int log2L, int mode, const PassArgs& a, const float2* d_tables,
                            long long ntiles, int num_sms, cudaStream_t st) {
  switch (log2L) {
#define X(l) case l: return fast_launch_l##l(mode, a, d_tables, ntiles, num_sms, st);
    PBK_FAST_LENGTHS(X)
#undef X
  }
  return cudaErrorInvalidValue;
}

}  // namespace pbk
