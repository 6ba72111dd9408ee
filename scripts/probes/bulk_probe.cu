// Micro-benchmark for round 2 (wgrad): how many bytes per second can ONE SM pull from DRAM / L2
// through the bulk-copy (TMA) engine, as a function of ring depth, stage size and copy size, when all SMs stream at once?
// The wgrad kernel sits at ~40 GB/s per SM with a 3 x 64 KB ring of 8-32 KB copies; this probe has the same ring without
// any tensor-core work behind it.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../intrinsic-neural-fields_b200/csrc -o bulk_probe bulk_probe.cu
//   ./bulk_probe [MB per CTA, default 8]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"

using namespace inf;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// warp 0 lane 0: loader; warp 1 lane 0: consumer (waits FULL, releases the stage at once)
__global__ void __launch_bounds__(64) bulk_ring_kernel(const uint8_t* __restrict__ src, size_t bytes_per_cta, int nstage,
                                                       uint32_t stage_bytes, uint32_t piece_bytes, unsigned int* err) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  const uint32_t bars = base + static_cast<uint32_t>(nstage) * stage_bytes;
  auto full = [&](int i) { return bars + 8u * static_cast<uint32_t>(i); };
  auto empty = [&](int i) { return bars + 8u * static_cast<uint32_t>(nstage + i); };
  if (threadIdx.x == 0) {
    for (int i = 0; i < nstage; ++i) { mbar_init(full(i), 1); mbar_init(empty(i), 1); }
    mbar_fence_init();
  }
  __syncthreads();
  const uint8_t* p = src + static_cast<size_t>(blockIdx.x) * bytes_per_cta;
  const int nunits = static_cast<int>(bytes_per_cta / stage_bytes);
  if (threadIdx.x == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int u = 0; u < nunits; ++u) {
      mbar_wait(empty(stage), phase ^ 1u, err, 0x1);
      mbar_arrive_expect_tx(full(stage), stage_bytes);
      for (uint32_t o = 0; o < stage_bytes; o += piece_bytes)
        bulk_g2s(base + static_cast<uint32_t>(stage) * stage_bytes + o, p + static_cast<size_t>(u) * stage_bytes + o, piece_bytes, full(stage));
      if (++stage == nstage) { stage = 0; phase ^= 1u; }
    }
  } else if (threadIdx.x == 32) {
    int stage = 0; uint32_t phase = 0;
    for (int u = 0; u < nunits; ++u) {
      mbar_wait(full(stage), phase, err, 0x2);
      mbar_arrive(empty(stage));
      if (++stage == nstage) { stage = 0; phase ^= 1u; }
    }
  }
}

int main(int argc, char** argv) {
  const size_t mb = argc > 1 ? static_cast<size_t>(atoll(argv[1])) : 8;
  int dev = 0, nsm = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  const size_t per_cta = mb << 20, total = per_cta * static_cast<size_t>(nsm);
  uint8_t* src; unsigned int* err;
  CK(cudaMalloc(&src, total)); CK(cudaMemset(src, 1, total));
  CK(cudaMalloc(&err, 4)); CK(cudaMemset(err, 0, 4));
  CK(cudaFuncSetAttribute(bulk_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  printf("%d SMs x %zu MB streamed through a bulk-copy ring (no consumer work)\n", nsm, mb);
  struct Cfg { int nstage; uint32_t stage, piece; };
  const Cfg cfgs[] = {{3, 65536, 8192}, {3, 65536, 32768}, {3, 65536, 65536}, {6, 32768, 8192}, {6, 32768, 32768}, {7, 32768, 32768},
                      {12, 16384, 16384}, {14, 16384, 8192}, {2, 65536, 32768}, {1, 65536, 32768}, {3, 32768, 32768}};
  for (const Cfg& c : cfgs) {
    const int smem = c.nstage * static_cast<int>(c.stage) + 2 * c.nstage * 8 + 64;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    bulk_ring_kernel<<<nsm, 64, smem>>>(src, per_cta, c.nstage, c.stage, c.piece, err);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < 3; ++i) bulk_ring_kernel<<<nsm, 64, smem>>>(src, per_cta, c.nstage, c.stage, c.piece, err);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); ms /= 3;
    printf("  %2d stages x %3u KB, copies of %2u KB (%3d KB in flight): %7.3f ms  %6.2f TB/s  %5.1f GB/s per SM\n", c.nstage,
           c.stage >> 10, c.piece >> 10, c.nstage * static_cast<int>(c.stage >> 10), ms, total / ms / 1e9, per_cta / ms / 1e6);
  }
  unsigned int h = 0; CK(cudaMemcpy(&h, err, 4, cudaMemcpyDeviceToHost));
  if (h) printf("error word %08x\n", h);
  return 0;
}
