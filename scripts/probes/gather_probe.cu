// Micro-benchmark behind DESIGN.md section 4: what random-row gather bandwidth does a B200 deliver, as a function of
// the contiguous piece a warp pulls from a row before it moves to the next row?  The fused kernel walks a tile's 384
// rows one 256-byte k-chunk at a time (piece = 256 B); this probe reads the same rows in pieces of 256 B ... 4 KB.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_probe gather_probe.cu && ./gather_probe [rows]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int ROW_BYTES = 4096;

// Each warp owns groups of G rows.  For every piece p of the row (PIECE bytes), it visits all G rows and reads
// that piece (PIECE/512 coalesced float4 loads per lane), so consecutive pieces of one row are G visits apart.
template <int PIECE, int G>
__global__ void __launch_bounds__(512) gather_kernel(const float4* __restrict__ table, const uint32_t* __restrict__ idx,
                                                     int64_t n_idx, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  float acc = 0.f;
  constexpr int LPP = PIECE / 512;   // float4 loads per lane per piece
  for (int64_t g0 = warp * G; g0 + G <= n_idx; g0 += nwarps * G) {
    uint32_t rows[G];
#pragma unroll
    for (int i = 0; i < G; ++i) rows[i] = idx[g0 + i];
    for (int p = 0; p < ROW_BYTES / PIECE; ++p) {
#pragma unroll
      for (int i = 0; i < G; i += 4) {
        float4 v[4][LPP];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int l = 0; l < LPP; ++l)
            v[u][l] = __ldg(table + static_cast<int64_t>(rows[i + u]) * (ROW_BYTES / 16) + p * (PIECE / 16) + l * 32 + lane);
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int l = 0; l < LPP; ++l) acc += v[u][l].x + v[u][l].y + v[u][l].z + v[u][l].w;
      }
    }
  }
  if (acc == 12345.678f) out[0] = acc;
}

// 256-byte pieces read the way the fused kernel does: 8 lanes x 16 B... here 16 lanes x 16 B per row, two rows per load.
template <int G>
__global__ void __launch_bounds__(512) gather256_kernel(const float4* __restrict__ table, const uint32_t* __restrict__ idx,
                                                        int64_t n_idx, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int half = lane >> 4, l16 = lane & 15;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  float acc = 0.f;
  for (int64_t g0 = warp * G; g0 + G <= n_idx; g0 += nwarps * G) {
    uint32_t rows[G / 2];
#pragma unroll
    for (int i = 0; i < G / 2; ++i) rows[i] = idx[g0 + 2 * i + half];
    for (int p = 0; p < ROW_BYTES / 256; ++p) {
#pragma unroll
      for (int i = 0; i < G / 2; i += 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(table + static_cast<int64_t>(rows[i + u]) * (ROW_BYTES / 16) + p * 16 + l16);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
      }
    }
  }
  if (acc == 12345.678f) out[0] = acc;
}


// The fused kernel's producer loop on its own (no tensor-core consumer, no ring hand-off): 16 warps, a warp-level load
// touches 4 rays x 128 B, DEPTH passes of 6 LDG.128 per thread in flight, interpolation + fp16 pack + swizzled st.shared.
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  uint32_t r; asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r;
}
template <int FLAVOR>
__device__ __forceinline__ float4 ldg_flavor(const float* p) {
  float4 v;
  if (FLAVOR == 0) return __ldg(reinterpret_cast<const float4*>(p));
  if (FLAVOR == 1) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  if (FLAVOR == 3) asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  if (FLAVOR == 4) asm volatile("ld.global.cv.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  if (FLAVOR == 2) asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
template <int DEPTH, bool STORE, int FENCE = 0, int FLAVOR = 0>
__global__ void __launch_bounds__(512) producer_kernel(const float* __restrict__ table, const uint32_t* __restrict__ idx,
                                                       int ntiles, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint32_t* s_vid = reinterpret_cast<uint32_t*>(smem);          // 384 per tile slot, 2 slots
  uint8_t* panels = smem + 4096;                                 // 4 x 16 KB
  const int tid = threadIdx.x, lane = tid & 31, pw = tid >> 5;
  const int s = lane & 7, gq = lane >> 3;
  float acc = 0.f;
  float4 a[DEPTH][3], b[DEPTH][3];
  auto quad_row = [&](int j) { return 8 * pw + 2 * j + (gq >> 1) + 4 * (gq & 1); };
  int it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    __syncthreads();
    for (int i = tid; i < 384; i += 512) s_vid[i] = idx[static_cast<int64_t>(tile) * 384 + i];
    __syncthreads();
    // passes: 16 chunks x 2 quads; software pipeline of DEPTH passes
    constexpr int NP = 32;
    auto issue = [&](int n, int slot) {
      const int c = n >> 1, q = n & 1;
      const int e = 3 * quad_row(q);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float* p = table + static_cast<int64_t>(s_vid[e + j]) * 1024 + c * 64 + 4 * s;
        a[slot][j] = ldg_flavor<FLAVOR>(p);
        b[slot][j] = ldg_flavor<FLAVOR>(p + 32);
      }
    };
    auto consume = [&](int n, int slot) {
      const int c = n >> 1, q = n & 1;
      const int r = quad_row(q);
      const float w0 = 0.25f, w1 = 0.5f, w2 = 0.25f;
      float x0 = w0 * a[slot][0].x + w1 * a[slot][1].x + w2 * a[slot][2].x, x1 = w0 * a[slot][0].y + w1 * a[slot][1].y + w2 * a[slot][2].y;
      float x2 = w0 * a[slot][0].z + w1 * a[slot][1].z + w2 * a[slot][2].z, x3 = w0 * a[slot][0].w + w1 * a[slot][1].w + w2 * a[slot][2].w;
      float y0 = w0 * b[slot][0].x + w1 * b[slot][1].x + w2 * b[slot][2].x, y1 = w0 * b[slot][0].y + w1 * b[slot][1].y + w2 * b[slot][2].y;
      float y2 = w0 * b[slot][0].z + w1 * b[slot][1].z + w2 * b[slot][2].z, y3 = w0 * b[slot][0].w + w1 * b[slot][1].w + w2 * b[slot][2].w;
      if (STORE) {
        uint8_t* row = panels + (c & 3) * 16384 + r * 128;
        const uint32_t sw = r & 7, half = (s & 1) << 3;
        *reinterpret_cast<uint2*>(row + ((((s >> 1)) ^ sw) << 4) + half) = make_uint2(pack_h2(x0, x1), pack_h2(x2, x3));
        *reinterpret_cast<uint2*>(row + (((4 + (s >> 1)) ^ sw) << 4) + half) = make_uint2(pack_h2(y0, y1), pack_h2(y2, y3));
      } else {
        acc += x0 + x1 + x2 + x3 + y0 + y1 + y2 + y3;
      }
    };
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) issue(d, d);
#pragma unroll 1
    for (int n = 0; n < NP; n += DEPTH) {
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) {
        consume(n + d, d);
        if (n + d + DEPTH < NP) issue(n + d + DEPTH, d);
        // what the fused kernel does after every k-chunk (2 passes): publish the panel to the async proxy
        if (FENCE == 1 && ((n + d) & 1)) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); __syncwarp(); }
        if (FENCE == 2 && ((n + d) & 1)) { __threadfence_block(); __syncwarp(); }
      }
    }
  }
  if (STORE) acc = static_cast<float>(panels[tid]);
  if (acc == 12345.678f) out[0] = acc;
}

// The same loop with cp.async (LDGSTS, 16 B, L1 bypass) into thread-private staging slots instead of register loads.
template <int DEPTH>
__global__ void __launch_bounds__(512) producer_cpasync_kernel(const float* __restrict__ table, const uint32_t* __restrict__ idx,
                                                               int ntiles, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint32_t* s_vid = reinterpret_cast<uint32_t*>(smem);
  uint8_t* panels = smem + 4096;                                 // 4 x 16 KB
  uint8_t* stage = smem + 4096 + 65536;                          // DEPTH x 6 x 512 x 16 B
  const int tid = threadIdx.x, lane = tid & 31, pw = tid >> 5;
  const int s = lane & 7, gq = lane >> 3;
  auto quad_row = [&](int j) { return 8 * pw + 2 * j + (gq >> 1) + 4 * (gq & 1); };
  const uint32_t stage_u32 = static_cast<uint32_t>(__cvta_generic_to_shared(stage));
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    __syncthreads();
    for (int i = tid; i < 384; i += 512) s_vid[i] = idx[static_cast<int64_t>(tile) * 384 + i];
    __syncthreads();
    constexpr int NP = 32;
    auto issue = [&](int n, int slot) {
      const int c = n >> 1, q = n & 1;
      const int e = 3 * quad_row(q);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float* p = table + static_cast<int64_t>(s_vid[e + j]) * 1024 + c * 64 + 4 * s;
        const uint32_t d0 = stage_u32 + (((slot * 6 + 2 * j) * 512 + tid) << 4), d1 = d0 + (512 << 4);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0), "l"(p) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d1), "l"(p + 32) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto consume = [&](int n, int slot) {
      const int c = n >> 1, q = n & 1;
      const int r = quad_row(q);
      float4 a[3], b[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        a[j] = *reinterpret_cast<const float4*>(stage + (((slot * 6 + 2 * j) * 512 + tid) << 4));
        b[j] = *reinterpret_cast<const float4*>(stage + (((slot * 6 + 2 * j + 1) * 512 + tid) << 4));
      }
      const float w0 = 0.25f, w1 = 0.5f, w2 = 0.25f;
      float x0 = w0 * a[0].x + w1 * a[1].x + w2 * a[2].x, x1 = w0 * a[0].y + w1 * a[1].y + w2 * a[2].y;
      float x2 = w0 * a[0].z + w1 * a[1].z + w2 * a[2].z, x3 = w0 * a[0].w + w1 * a[1].w + w2 * a[2].w;
      float y0 = w0 * b[0].x + w1 * b[1].x + w2 * b[2].x, y1 = w0 * b[0].y + w1 * b[1].y + w2 * b[2].y;
      float y2 = w0 * b[0].z + w1 * b[1].z + w2 * b[2].z, y3 = w0 * b[0].w + w1 * b[1].w + w2 * b[2].w;
      uint8_t* row = panels + (c & 3) * 16384 + r * 128;
      const uint32_t sw = r & 7, half = (s & 1) << 3;
      *reinterpret_cast<uint2*>(row + ((((s >> 1)) ^ sw) << 4) + half) = make_uint2(pack_h2(x0, x1), pack_h2(x2, x3));
      *reinterpret_cast<uint2*>(row + (((4 + (s >> 1)) ^ sw) << 4) + half) = make_uint2(pack_h2(y0, y1), pack_h2(y2, y3));
    };
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) issue(d, d);
#pragma unroll 1
    for (int n = 0; n < NP; n += DEPTH) {
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) {
        asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
        consume(n + d, d);
        if (n + d + DEPTH < NP) issue(n + d + DEPTH, d);
        else asm volatile("cp.async.commit_group;" ::: "memory");
      }
    }
  }
  if (panels[tid] == 123) out[0] = 1.f;
}

template <typename F>
float time_ms(F f, int reps) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f(); f();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; ++i) f();
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  float ms;
  CK(cudaEventElapsedTime(&ms, a, b));
  return ms / reps;
}

int main(int argc, char** argv) {
  const int64_t V = argc > 1 ? atoll(argv[1]) : 400000;
  const int64_t n_idx = 3 * 131072 * 4;   // 4 "steps" worth of rows per launch, no repetition pattern
  float4* table; uint32_t* idx; float* out;
  CK(cudaMalloc(&table, V * ROW_BYTES)); CK(cudaMemset(table, 0, V * ROW_BYTES));
  CK(cudaMalloc(&idx, n_idx * 4)); CK(cudaMalloc(&out, 4));
  std::vector<uint32_t> h(n_idx);
  uint64_t s = 88172645463325252ull;
  for (auto& x : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; x = static_cast<uint32_t>(s % V); }
  CK(cudaMemcpy(idx, h.data(), n_idx * 4, cudaMemcpyHostToDevice));
  const double bytes = static_cast<double>(n_idx) * ROW_BYTES;
  printf("table %.2f GB (%lld rows of 4 KB), %lld random rows per launch\n", V * 4096.0 / 1e9, (long long)V, (long long)n_idx);
  for (int ctas : {148, 296, 592}) {
    auto report = [&](const char* name, float ms) { printf("  grid %4d  %-28s %7.3f ms  %6.2f TB/s\n", ctas, name, ms, bytes / ms / 1e9); };
    report("piece 256 B (16 lanes/row), G=48", time_ms([&] { gather256_kernel<48><<<ctas, 512>>>(table, idx, n_idx, out); }, 5));
    report("piece 512 B, G=24", time_ms([&] { gather_kernel<512, 24><<<ctas, 512>>>(table, idx, n_idx, out); }, 5));
    report("piece 1 KB, G=24", time_ms([&] { gather_kernel<1024, 24><<<ctas, 512>>>(table, idx, n_idx, out); }, 5));
    report("piece 2 KB, G=12", time_ms([&] { gather_kernel<2048, 12><<<ctas, 512>>>(table, idx, n_idx, out); }, 5));
    report("piece 4 KB (whole row), G=8", time_ms([&] { gather_kernel<4096, 8><<<ctas, 512>>>(table, idx, n_idx, out); }, 5));
    CK(cudaGetLastError());
  }
  // the fused kernel's producer loop alone: 4 x 1024 tiles of 128 rays (the index list above is 4 steps of 131072 rays)
  {
    const int ntiles = static_cast<int>(n_idx / 384);
    const int smem_bytes = 4096 + 4 * 16384;
    auto run = [&](const char* name, auto kern) {
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
      float ms = time_ms([&] { kern<<<148, 512, smem_bytes>>>(reinterpret_cast<const float*>(table), idx, ntiles, out); }, 5);
      CK(cudaGetLastError());
      printf("  producer loop alone, %-26s %7.3f ms  %6.2f TB/s\n", name, ms, bytes / ms / 1e9);
    };
    run("2 passes in flight, store", producer_kernel<2, true>);
    run("2 passes in flight, no store", producer_kernel<2, false>);
    run("4 passes in flight, store", producer_kernel<4, true>);
    run("4 passes in flight, no store", producer_kernel<4, false>);
    run("1 pass in flight, store", producer_kernel<1, true>);
    // L1 capacity: the fused kernel leaves ~25 KB of the 256 KB L1/shared array to L1; pad the probe's shared memory
    for (int kb : {68, 132, 164, 196, 227}) {
      auto run_kb = [&](const char* name, auto kern) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kb * 1024));
        float ms = time_ms([&] { kern<<<148, 512, kb * 1024>>>(reinterpret_cast<const float*>(table), idx, ntiles, out); }, 5);
        CK(cudaGetLastError());
        printf("  producer loop, smem %3d KB, %-30s %7.3f ms  %6.2f TB/s\n", kb, name, ms, bytes / ms / 1e9);
      };
      run_kb("ld.nc, 2 passes", producer_kernel<2, true, 0, 0>);
      run_kb("ld.nc.no_allocate, 2 passes", producer_kernel<2, true, 0, 1>);
      run_kb("ld.nc.no_allocate.L2::256B, 2 p.", producer_kernel<2, true, 0, 2>);
      run_kb("ld.cg, 2 passes", producer_kernel<2, true, 0, 3>);
      run_kb("ld.cv, 2 passes", producer_kernel<2, true, 0, 4>);
      if (kb >= 164) {
        run_kb("cp.async.cg, 1 pass (48 KB)", producer_cpasync_kernel<1>);
        run_kb("cp.async.cg, 2 passes (96 KB)", producer_cpasync_kernel<2>);
      }
    }
    run("2 passes, store, proxy fence", producer_kernel<2, true, 1>);
    run("4 passes, store, proxy fence", producer_kernel<4, true, 1>);
    run("2 passes, store, membar.cta", producer_kernel<2, true, 2>);
  }
  return 0;
}
