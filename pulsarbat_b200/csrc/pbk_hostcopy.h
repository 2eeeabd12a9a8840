// pbk_hostcopy.h -- copies between PAGEABLE host memory and the device for the *_host entry points.
//
// A numpy block handed to the reference-facing API lives in ordinary pageable memory.  A plain
// cudaMemcpy from or to such memory is staged by the driver through one thread (about 10-20 GB/s
// in, and about 5 GB/s out into a freshly allocated result, whose every 4 KiB page faults on its
// first write).  Here large copies are cut into chunks that several host threads move through
// page-locked bounce buffers, each thread on its own CUDA stream: the host memcpy (and the page
// faults of a fresh result) of one chunk overlap the DMA of the others.  Page-locked or registered
// host pointers, small copies and PBK_BOUNCE=0 take the plain cudaMemcpyAsync path.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace pbk {

constexpr size_t kBounceChunk = 4u << 20;        // bytes per chunk
constexpr size_t kBounceMin = 32u << 20;         // smaller copies go straight to cudaMemcpyAsync
constexpr int kBounceMaxLanes = 16;

struct BounceLane {
  int device = -1;
  cudaStream_t stream = nullptr;
  void* buf[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  cudaEvent_t fin = nullptr;
  bool busy = false;
};

// process-wide pool of lanes (a lane = stream + two page-locked chunks), created on first use
class BouncePool {
 public:
  static BouncePool& get() {
    static BouncePool p;
    return p;
  }
  // up to `want` idle lanes of `device`; fewer (possibly none) when others are in use
  std::vector<BounceLane*> acquire(int device, int want) {
    std::lock_guard<std::mutex> lock(mu_);
    std::vector<BounceLane*> got;
    for (auto* l : lanes_)
      if ((int)got.size() < want && !l->busy && l->device == device) { l->busy = true; got.push_back(l); }
    while ((int)got.size() < want && count(device) < kBounceMaxLanes) {
      BounceLane* l = create(device);
      if (!l) break;
      l->busy = true;
      lanes_.push_back(l);
      got.push_back(l);
    }
    return got;
  }
  void release(const std::vector<BounceLane*>& ls) {
    std::lock_guard<std::mutex> lock(mu_);
    for (auto* l : ls) l->busy = false;
  }

 private:
  int count(int device) const {
    int n = 0;
    for (auto* l : lanes_) n += l->device == device;
    return n;
  }
  static BounceLane* create(int device) {
    BounceLane* l = new BounceLane();
    l->device = device;
    bool ok = cudaStreamCreateWithFlags(&l->stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; ok && i < 2; ++i) {
      ok = cudaHostAlloc(&l->buf[i], kBounceChunk, cudaHostAllocDefault) == cudaSuccess &&
           cudaEventCreateWithFlags(&l->ev[i], cudaEventDisableTiming) == cudaSuccess;
    }
    ok = ok && cudaEventCreateWithFlags(&l->fin, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {   // leave nothing half-made behind
      for (int i = 0; i < 2; ++i) {
        if (l->buf[i]) cudaFreeHost(l->buf[i]);
        if (l->ev[i]) cudaEventDestroy(l->ev[i]);
      }
      if (l->fin) cudaEventDestroy(l->fin);
      if (l->stream) cudaStreamDestroy(l->stream);
      delete l;
      cudaGetLastError();
      return nullptr;
    }
    return l;
  }
  std::mutex mu_;
  std::vector<BounceLane*> lanes_;
};

inline bool bounce_wanted(const void* host_ptr, size_t bytes) {
  if (bytes < kBounceMin) return false;
  const char* e = getenv("PBK_BOUNCE");
  if (e && !strcmp(e, "0")) return false;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, host_ptr) != cudaSuccess) {
    cudaGetLastError();
    return true;    // older drivers report unregistered memory as an error
  }
  return at.type == cudaMemoryTypeUnregistered;
}

inline int bounce_threads() {
  if (const char* e = getenv("PBK_BOUNCE_THREADS")) return std::max(1, std::min(kBounceMaxLanes, atoi(e)));
  const unsigned hw = std::thread::hardware_concurrency();
  return (int)std::max(1u, std::min<unsigned>(8, hw / 2));   // half the cores, at most 8, by default
}

// host -> device; all later work on `st` is ordered after the copy.  A large PAGEABLE source takes
// the bounce pipeline: on return every byte of h_src has been read (the caller may reuse it) while
// the DMA may still be in flight.  A page-locked source, or any copy under 32 MiB, is one plain
// cudaMemcpyAsync: h_src must then stay valid until the stream has reached the copy.
inline cudaError_t host_to_device(void* d_dst, const void* h_src, size_t bytes, cudaStream_t st) {
  if (!bounce_wanted(h_src, bytes))
    return cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st);
  int device = 0;
  cudaGetDevice(&device);
  auto lanes = BouncePool::get().acquire(device, bounce_threads());
  if (lanes.empty()) return cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st);
  const int T = (int)lanes.size();
  const size_t nchunks = (bytes + kBounceChunk - 1) / kBounceChunk;
  cudaEvent_t start = lanes[0]->fin;   // (re-recorded below once the lane is done with it)
  cudaError_t first = cudaEventRecord(start, st);   // the destination may be in use by earlier work
  for (int w = 0; w < T && first == cudaSuccess; ++w) first = cudaStreamWaitEvent(lanes[w]->stream, start, 0);
  std::vector<cudaError_t> err(T, cudaSuccess);
  std::vector<std::thread> th;
  for (int w = 0; w < T && first == cudaSuccess; ++w)
    th.emplace_back([&, w] {
      BounceLane& L = *lanes[w];
      cudaSetDevice(device);
      size_t k = 0;
      for (size_t i = w; i < nchunks && err[w] == cudaSuccess; i += T, ++k) {
        const int b = (int)(k & 1);
        const size_t off = i * kBounceChunk, len = std::min(kBounceChunk, bytes - off);
        if (k >= 2 && (err[w] = cudaEventSynchronize(L.ev[b])) != cudaSuccess) break;
        memcpy(L.buf[b], static_cast<const char*>(h_src) + off, len);
        err[w] = cudaMemcpyAsync(static_cast<char*>(d_dst) + off, L.buf[b], len,
                                 cudaMemcpyHostToDevice, L.stream);
        if (err[w] == cudaSuccess) err[w] = cudaEventRecord(L.ev[b], L.stream);
      }
    });
  for (auto& t : th) t.join();
  for (int w = 0; w < T; ++w) {
    if (first == cudaSuccess && err[w] != cudaSuccess) first = err[w];
    // the bounce buffers must not be refilled before their DMA has finished, and `st` must see the data
    cudaError_t e = cudaEventRecord(lanes[w]->fin, lanes[w]->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(st, lanes[w]->fin, 0);
    if (first == cudaSuccess && e != cudaSuccess) first = e;
  }
  // lanes go back to the pool only when their DMA is done (another caller would overwrite them)
  for (int w = 0; w < T; ++w) cudaEventSynchronize(lanes[w]->fin);
  BouncePool::get().release(lanes);
  return first;
}

// device -> host, after all earlier work on `st`.  Synchronous: the data is in h_dst on return.
inline cudaError_t device_to_host(void* h_dst, const void* d_src, size_t bytes, cudaStream_t st) {
  if (!bounce_wanted(h_dst, bytes)) {
    cudaError_t e = cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, st);
    return e == cudaSuccess ? cudaStreamSynchronize(st) : e;
  }
  int device = 0;
  cudaGetDevice(&device);
  auto lanes = BouncePool::get().acquire(device, bounce_threads());
  if (lanes.empty()) {
    cudaError_t e = cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, st);
    return e == cudaSuccess ? cudaStreamSynchronize(st) : e;
  }
  const int T = (int)lanes.size();
  const size_t nchunks = (bytes + kBounceChunk - 1) / kBounceChunk;
  cudaEvent_t start = lanes[0]->fin;
  cudaError_t first = cudaEventRecord(start, st);    // the kernels that produce d_src
  for (int w = 0; w < T && first == cudaSuccess; ++w) first = cudaStreamWaitEvent(lanes[w]->stream, start, 0);
  std::vector<cudaError_t> err(T, cudaSuccess);
  std::vector<std::thread> th;
  for (int w = 0; w < T && first == cudaSuccess; ++w)
    th.emplace_back([&, w] {
      BounceLane& L = *lanes[w];
      cudaSetDevice(device);
      auto issue = [&](size_t i, int b) {
        const size_t off = i * kBounceChunk, len = std::min(kBounceChunk, bytes - off);
        cudaError_t e = cudaMemcpyAsync(L.buf[b], static_cast<const char*>(d_src) + off, len,
                                        cudaMemcpyDeviceToHost, L.stream);
        return e == cudaSuccess ? cudaEventRecord(L.ev[b], L.stream) : e;
      };
      // two chunks in flight per lane: while one is copied out to the user's pages the next lands
      size_t k = 0;
      for (size_t i = w; i < nchunks && k < 2 && err[w] == cudaSuccess; i += T, ++k) err[w] = issue(i, (int)k);
      k = 0;
      for (size_t i = w; i < nchunks && err[w] == cudaSuccess; i += T, ++k) {
        const int b = (int)(k & 1);
        const size_t off = i * kBounceChunk, len = std::min(kBounceChunk, bytes - off);
        if ((err[w] = cudaEventSynchronize(L.ev[b])) != cudaSuccess) break;
        memcpy(static_cast<char*>(h_dst) + off, L.buf[b], len);
        const size_t nxt = i + 2 * (size_t)T;
        if (nxt < nchunks) err[w] = issue(nxt, b);
      }
      if (err[w] != cudaSuccess) cudaStreamSynchronize(L.stream);
    });
  for (auto& t : th) t.join();
  for (int w = 0; w < T; ++w)
    if (first == cudaSuccess && err[w] != cudaSuccess) first = err[w];
  BouncePool::get().release(lanes);
  if (first == cudaSuccess) first = cudaStreamSynchronize(st);
  return first;
}

}  // namespace pbk
