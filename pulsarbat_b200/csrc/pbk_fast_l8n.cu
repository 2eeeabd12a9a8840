// 2^8-point tiles HALF as wide as pbk_fast_l8.cu's (256 points x 16 lanes, 32 KiB, 128 threads,
// four CTAs per SM) for the detecting last pass of a channelizer plan (FSUM, pbk_fast.cuh): that
// pass only reads, so four small CTAs per SM in different phases overlap its load and compute
// phases better than two large ones (0.33 -> 0.30 ms).  Only the two kernels that pass uses are
// instantiated.  (The time-summing last pass of a dedispersion plan was tried on these tiles too:
// 0.96 -> 1.04 ms -- its tile rows are 16 MB apart, and 128-byte chunks of such rows run into the
// chunk-rate limit of DESIGN.md 5.1.)
#include "pbk_fast_inst.cuh"

namespace pbk {

using Cfg = FastCfg<16, 16, 1, 1, 3, 128, 4>;

void fast_info_l8n(FastInfo* info) { cfg_info<Cfg>(info); }
void fast_tables_l8n(float2* dst) { fast_build_tables<Cfg>(dst); }
cudaError_t fast_launch_l8n(int mode, const PassArgs& a, const float2* d_tables, long long ntiles,
                            int num_sms, cudaStream_t st) {
  if (mode != MODE_FWD || !a.final_epi || a.fsum_log2 <= 0 || a.sign > 0)
    return cudaErrorInvalidValue;
  if (a.load_kind == LOAD_TRANSP_PLANAR)
    return cfg_launch_variant<MODE_FWD, Cfg, LK_TRANSP_PLANAR, EPI_INTENSITY, false, true>(
        a, d_tables, ntiles, num_sms, st);
  return cfg_launch_variant<MODE_FWD, Cfg, LK_PLANAR, EPI_INTENSITY, false, true>(
      a, d_tables, ntiles, num_sms, st);
}

}  // namespace pbk
