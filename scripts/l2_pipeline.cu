// Microbenchmark 4 (B200): can three dependent in-place passes over the same block be kept in L2
// by one persistent kernel that hands out tiles in a software-pipelined order
//   A(c0) A(c1) B(c0) A(c2) B(c1) C(c0) A(c3) B(c2) C(c1) ...
// with per-(block, phase) completion counters instead of kernel boundaries?
// Tile = 32 KiB: phases A and C touch 128 rows x 256 B at stride block/128 (the strided FFT
// levels), phase B touches 32 KiB contiguous (the fused middle level).  Every phase reads its tile
// and writes it back in place.  Compared with three plain full-array passes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_pipeline l2_pipeline.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int kThreads = 128;
constexpr int kTileBytes = 32 * 1024;

__device__ __forceinline__ void touch_tile(float4* base, long long block_bytes, int phase,
                                           int tile_in_block, float mul, int work) {
  // 2048 float4 per tile, 16 per thread
  float4 v[16];
  if (phase == 1) {
    float4* p = base + (long long)tile_in_block * (kTileBytes / 16);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __ldcg(p + threadIdx.x + i * kThreads);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      for (int w = 0; w < work; ++w) v[i].x = fmaf(v[i].x, mul, v[i].y);
      v[i].x *= mul;
      p[threadIdx.x + i * kThreads] = v[i];
    }
  } else {
    // 128 rows x 256 B (16 float4): row stride = block_bytes/128; column chunk = tile_in_block
    const long long rs = block_bytes / 128 / 16;   // in float4
    float4* p = base + (long long)tile_in_block * 16;
    const int j = threadIdx.x & 15, r0 = threadIdx.x >> 4;   // 8 rows per sweep
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __ldcg(p + (long long)(r0 + 8 * i) * rs + j);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      for (int w = 0; w < work; ++w) v[i].x = fmaf(v[i].x, mul, v[i].y);
      v[i].x *= mul;
      p[(long long)(r0 + 8 * i) * rs + j] = v[i];
    }
  }
}

// plain pass: all blocks, one phase
__global__ void __launch_bounds__(kThreads, 4) plain_pass(float4* a, long long block_bytes,
                                                          int nblocks, int phase, int work) {
  const int tpb = (int)(block_bytes / kTileBytes);
  const long long ntiles = (long long)nblocks * tpb;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int blk = (int)(t / tpb), til = (int)(t % tpb);
    touch_tile(a + (long long)blk * (block_bytes / 16), block_bytes, phase, til, 1.0001f, work);
  }
}

// pipelined persistent kernel: ticket -> (slot, tile); slot s runs phase A of block s, phase B of
// block s-1 and phase C of block s-2, interleaved tile by tile so all three make progress together
__global__ void __launch_bounds__(kThreads, 4) fused_pipeline(float4* a, long long block_bytes,
                                                              int nblocks, unsigned* ticket,
                                                              unsigned* done, int work) {
  const int tpb = (int)(block_bytes / kTileBytes);
  const long long total = (long long)(nblocks + 2) * 3 * tpb;
  __shared__ long long s_t;
  for (;;) {
    if (threadIdx.x == 0) s_t = atomicAdd(ticket, 1u);
    __syncthreads();
    const long long t = s_t;
    __syncthreads();
    if (t >= total) break;
    const int slot = (int)(t / (3 * tpb));
    const int r = (int)(t % (3 * tpb));
    const int phase = r % 3, til = r / 3;
    const int blk = slot - phase;
    if (blk < 0 || blk >= nblocks) continue;
    if (phase > 0) {
      // wait until every tile of the previous phase of this block has been written
      if (threadIdx.x == 0) {
        const volatile unsigned* d = done + blk * 3 + (phase - 1);
        while (*d < (unsigned)tpb) __nanosleep(64);
        __threadfence();
      }
      __syncthreads();
    }
    touch_tile(a + (long long)blk * (block_bytes / 16), block_bytes, phase, til, 1.0001f, work);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(done + blk * 3 + phase, 1u);
  }
}

int main() {
  const size_t bytes = 4ull << 30;
  float4* a;
  CK(cudaMalloc(&a, bytes));
  CK(cudaMemset(a, 0, bytes));
  unsigned *ticket, *done;
  CK(cudaMalloc(&ticket, 4));
  CK(cudaMalloc(&done, 3 * 4096 * 4));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int grid = 148 * 4;
  for (int work : {0, 32}) {
    for (long long mb : {4, 8, 16, 32}) {
      const long long bb = mb << 20;
      const int nblocks = (int)(bytes / bb);
      float ms_plain, ms_fused;
      auto plain = [&]() {
        for (int ph = 0; ph < 3; ++ph) plain_pass<<<grid, kThreads>>>(a, bb, nblocks, ph, work);
      };
      plain();
      CK(cudaDeviceSynchronize());
      cudaEventRecord(e0);
      plain();
      cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1));
      cudaEventElapsedTime(&ms_plain, e0, e1);
      auto fused = [&]() {
        cudaMemsetAsync(ticket, 0, 4);
        cudaMemsetAsync(done, 0, 3 * 4096 * 4);
        fused_pipeline<<<grid, kThreads>>>(a, bb, nblocks, ticket, done, work);
      };
      fused();
      CK(cudaDeviceSynchronize());
      cudaEventRecord(e0);
      fused();
      cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1));
      cudaEventElapsedTime(&ms_fused, e0, e1);
      printf("work %2d block %3lld MiB: 3 plain passes %.3f ms (%.0f GB/s r+w per pass) | fused "
             "pipeline %.3f ms (%.2fx)\n",
             work, mb, ms_plain, 3 * 2.0 * bytes / ms_plain / 1e6, ms_fused, ms_plain / ms_fused);
    }
  }
  return 0;
}
