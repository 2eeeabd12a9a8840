// Microbenchmark 3 (B200): (A) L2-resident vs DRAM copy bandwidth; (B) strided-tile copies done
// with TMA (cp.async.bulk.tensor) instead of LDG/STG, for the narrow chunks that the LSU path
// cannot feed (profiles/r01_chunk_bw*_microbench.log).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_bw tma_bw.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

// ---------------------------------------------------------------- (A) plain copies
__global__ void copy_rep(const float4* __restrict__ in, float4* __restrict__ out, long long n16,
                         int repeat) {
  for (int r = 0; r < repeat; ++r)
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n16;
         i += (long long)gridDim.x * blockDim.x)
      out[i] = in[i];
}
__global__ void read_rep(const float4* __restrict__ in, float4* __restrict__ out, long long n16,
                         int repeat) {
  float4 acc = make_float4(0, 0, 0, 0);
  for (int r = 0; r < repeat; ++r)
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n16;
         i += (long long)gridDim.x * blockDim.x) {
      float4 v = __ldcg(in + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  if (acc.x == 1234.5f) out[0] = acc;
}

// ---------------------------------------------------------------- (B) TMA tile copies
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        " selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
  }
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1,
                                            int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
      "{%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, int c0, int c1, int c2,
                                             const void* src) {
  asm volatile(
      "cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map),
      "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(src))
      : "memory");
}

// tile = rows x chunk; mode 0: rows run along dim2 (stride 2 MiB), mode 1: along dim1 (1 KiB).
// NBUF tiles in flight per CTA (ring of buffers, one mbarrier each).
template <int NBUF>
__global__ void __launch_bounds__(32) tma_copy(const __grid_constant__ CUtensorMap tin,
                                               const __grid_constant__ CUtensorMap tout,
                                               int chunk_elems, int rows, int mode,
                                               long long ntiles, int chunks_per_row, int do_store) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar[NBUF];
  const int tile_bytes = rows * chunk_elems * 4;
  const int box_bytes = 256 * chunk_elems * 4;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NBUF; ++i) mbar_init(&bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (threadIdx.x != 0) return;
  auto issue = [&](long long t, int buf) {
    const int c = (int)(t % chunks_per_row), r = (int)(t / chunks_per_row);
    mbar_expect_tx(&bar[buf], (uint32_t)tile_bytes);
    for (int b = 0; b < rows / 256; ++b) {
      unsigned char* dst = smem + (size_t)buf * tile_bytes + (size_t)b * box_bytes;
      if (mode == 0) tma_load_3d(dst, &tin, c * chunk_elems, r, b * 256, &bar[buf]);
      else tma_load_3d(dst, &tin, c * chunk_elems, b * 256, r, &bar[buf]);
    }
  };
  long long t = blockIdx.x;
  // prologue
  long long tq[NBUF];
  for (int i = 0; i < NBUF; ++i) {
    tq[i] = t + (long long)i * gridDim.x;
    if (tq[i] < ntiles) issue(tq[i], i);
  }
  uint32_t phase[NBUF];
  for (int i = 0; i < NBUF; ++i) phase[i] = 0;
  int buf = 0;
  for (; t < ntiles; t += gridDim.x) {
    mbar_wait(&bar[buf], phase[buf]);
    phase[buf] ^= 1;
    if (do_store) {
      const int c = (int)(t % chunks_per_row), r = (int)(t / chunks_per_row);
      for (int b = 0; b < rows / 256; ++b) {
        const unsigned char* src = smem + (size_t)buf * tile_bytes + (size_t)b * box_bytes;
        if (mode == 0) tma_store_3d(&tout, c * chunk_elems, r, b * 256, src);
        else tma_store_3d(&tout, c * chunk_elems, b * 256, r, src);
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    const long long tn = t + (long long)NBUF * gridDim.x;
    if (tn < ntiles) issue(tn, buf);
    buf = (buf + 1) % NBUF;
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                             CUtensorMapFloatOOBfill);

int main() {
  const size_t bytes = 4ull << 30;
  float4 *a, *b;
  CK(cudaMalloc(&a, bytes));
  CK(cudaMalloc(&b, bytes));
  CK(cudaMemset(a, 1, bytes));
  CK(cudaMemset(b, 0, bytes));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);

  printf("(A) plain float4 copy / read, footprint F, repeated so that 8 GiB move in total\n");
  for (size_t mb : {8, 16, 32, 48, 64, 96, 128, 256, 4096}) {
    const size_t f = mb << 20;
    const int rep = (int)((8ull << 30) / f);
    for (int kind = 0; kind < 2; ++kind) {
      auto go = [&]() {
        if (kind == 0) copy_rep<<<148 * 8, 512>>>(a, a + (f / 2) / 16, (f / 2) / 16, rep);
        else read_rep<<<148 * 8, 512>>>(a, b, f / 16, rep);
      };
      go();
      CK(cudaDeviceSynchronize());
      cudaEventRecord(e0);
      go();
      cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1));
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double moved = (double)f * rep;   // copy: f/2 read + f/2 written per repeat
      printf("  F=%5zu MiB %s: %8.1f GB/s\n", mb, kind == 0 ? "copy (r+w)" : "read      ",
             moved / ms / 1e6);
    }
  }

  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
  if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }

  printf("(B) TMA tile copy over 4 GiB: tensor (256 f32 | 2048 rows @1KiB | 2048 @2MiB)\n");
  printf("  mode rows chunkB nbuf ctas/SM store   GB/s\n");
  for (int mode = 0; mode < 2; ++mode)
    for (int chunk : {32, 64, 128, 256})
      for (int rows : {2048, 1024, 256})
        for (int cfg = 0; cfg < 3; ++cfg) {
          const int nbuf = cfg == 0 ? 1 : 2;
          const int tile_bytes = rows * chunk;
          const int smem = tile_bytes * nbuf;
          if (smem > 200 * 1024) continue;
          int cps = (220 * 1024) / (smem + 1024);
          if (cps > 8) cps = 8;
          if (cfg == 2) { if (cps < 2) continue; cps = (cps + 1) / 2; }
          const int ce = chunk / 4;
          cuuint64_t dims[3] = {256, 2048, 2048};
          cuuint64_t strides[2] = {1024, 2ull << 20};
          cuuint32_t box[3] = {(cuuint32_t)ce, 1, 1};
          if (mode == 0) box[2] = 256; else box[1] = 256;
          cuuint32_t es[3] = {1, 1, 1};
          CUtensorMap tin, tout;
          CUresult r1 = encode(&tin, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, a, dims, strides, box, es,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          CUresult r2 = encode(&tout, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, b, dims, strides, box, es,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) { printf("encode failed %d %d\n", r1, r2); continue; }
          const int cpr = 1024 / chunk;
          // a tile spans `rows` of the 2048 along its axis: (2048/rows) tiles per column group
          const long long ntiles = 2048ll * cpr * (2048 / rows);
          // tile index -> (chunk c, other-axis index r); rows < 2048 handled by folding the extra
          // factor into r via a smaller tensor: keep it simple and only move the first `rows`
          const long long nt = 2048ll * cpr;
          (void)ntiles;
          for (int st = 1; st >= 0; --st) {
            auto go = [&]() {
              if (nbuf == 1)
                tma_copy<1><<<148 * cps, 32, smem>>>(tin, tout, ce, rows, mode, nt, cpr, st);
              else
                tma_copy<2><<<148 * cps, 32, smem>>>(tin, tout, ce, rows, mode, nt, cpr, st);
            };
            CK(cudaFuncSetAttribute(tma_copy<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            CK(cudaFuncSetAttribute(tma_copy<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            go();
            CK(cudaDeviceSynchronize());
            cudaEventRecord(e0);
            go();
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            const double moved = (double)nt * tile_bytes * (st ? 2 : 1);
            printf("  %4d %4d %6d %4d %7d %5d %8.1f\n", mode, rows, chunk, nbuf, cps, st,
                   moved / ms / 1e6);
          }
        }
  return 0;
}
