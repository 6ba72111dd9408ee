// Backward half of the vertex-projected path, sm_100a: the per-vertex gradient
//     G[v, :] = sum over the (ray, corner) pairs incident to vertex v of  bary[ray, corner] * [dz_0 | dz_skip][ray, :]
// whose product with the table, [dW0; dWy] = G^T E, is the first-layer weight gradient of trainer.py:71-84
// (dW0 = sum_r dz_0[r]^T x[r] with x[r] = sum_j bary_rj E[v_rj], regrouped by vertex).  The GEMM itself runs in
// field_wgrad.cu over the image written here and the table's operand image.
//
// G is built as a GATHER over a per-batch CSR (vertex -> incident pairs), not as a scatter of atomics:
//   csr_count / csr_scan_* / csr_fill   integer atomics only -- the SET of pairs of a vertex is deterministic, their
//                                        order inside the list is not;
//   vertex_gather_kernel                one warp per vertex, lane = 16-byte chunk of the two stashed dz rows of a ray
//                                        (coalesced 128-byte lines of the swizzled images field_rays.cu stored).  The sum
//                                        is made independent of the list order -- lists of <= 32 pairs are sorted in
//                                        registers first, longer ones are accumulated in 64-bit fixed point (integer
//                                        addition is associative) -- so every gradient is bit-identical from run to run.
#include "common.cuh"
#include "ptx.cuh"

namespace inf {

namespace {

__device__ __forceinline__ int64_t src_row(const FieldParams& p, int64_t ray) {
  if (p.ridx == nullptr) return ray;
  return p.ridx64 ? static_cast<int64_t>(reinterpret_cast<const long long*>(p.ridx)[ray])
                  : static_cast<int64_t>(reinterpret_cast<const int*>(p.ridx)[ray]);
}
__device__ __forceinline__ uint32_t vertex_of(const FieldParams& p, int64_t e) {
  return p.idx64 ? static_cast<uint32_t>(reinterpret_cast<const long long*>(p.vids)[e])
                 : static_cast<uint32_t>(reinterpret_cast<const int*>(p.vids)[e]);
}

}  // namespace

// one thread per (ray, corner) pair
__global__ void csr_count_kernel(const FieldParams p) {
  const int64_t total = p.n_rays * 3;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t ray = i / 3;
    const int j = static_cast<int>(i - ray * 3);
    const uint32_t v = vertex_of(p, src_row(p, ray) * 3 + j);
    if (static_cast<int64_t>(v) < p.n_rows) atomicAdd(&p.vcursor[v], 1);
  }
}

// exclusive scan of the counts, three passes over SCAN_BLOCK-element blocks
__global__ void __launch_bounds__(256) csr_scan_sums_kernel(const int32_t* __restrict__ cnt, int64_t n, int32_t* __restrict__ block_sums) {
  __shared__ int32_t s_part[8];
  const int64_t base = static_cast<int64_t>(blockIdx.x) * SCAN_BLOCK;
  int32_t acc = 0;
#pragma unroll
  for (int k = 0; k < SCAN_BLOCK / 256; ++k) {
    const int64_t i = base + k * 256 + threadIdx.x;
    if (i < n) acc += cnt[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    int32_t t = 0;
    for (int w = 0; w < 8; ++w) t += s_part[w];
    block_sums[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(1024) csr_scan_blocks_kernel(int32_t* __restrict__ block_sums, int nblocks) {
  // single block: exclusive scan in place (nblocks is V / 1024: a few hundred)
  __shared__ int32_t s_warp[32];
  __shared__ int32_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += 1024) {
    const int i = base + threadIdx.x;
    const int32_t x = i < nblocks ? block_sums[i] : 0;
    int32_t incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t y = __shfl_up_sync(0xffffffffu, incl, o);
      if ((threadIdx.x & 31) >= o) incl += y;
    }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
      int32_t w = s_warp[threadIdx.x];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int32_t y = __shfl_up_sync(0xffffffffu, w, o);
        if (threadIdx.x >= o) w += y;
      }
      s_warp[threadIdx.x] = w;   // inclusive over warps
    }
    __syncthreads();
    const int32_t warp_off = (threadIdx.x >> 5) ? s_warp[(threadIdx.x >> 5) - 1] : 0;
    const int32_t carry = s_carry;
    if (i < nblocks) block_sums[i] = carry + warp_off + incl - x;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + warp_off + incl;
    __syncthreads();
  }
}
// vstart[i] = exclusive prefix; vcursor[i] = vstart[i] (the fill pass bumps it); vstart[n] = total
__global__ void __launch_bounds__(256) csr_scan_apply_kernel(int32_t* __restrict__ cnt_cursor, int64_t n, const int32_t* __restrict__ block_sums,
                                                            int32_t* __restrict__ vstart) {
  __shared__ int32_t s_warp[8];
  const int64_t base = static_cast<int64_t>(blockIdx.x) * SCAN_BLOCK;
  // thread t owns the 4 consecutive elements base + 4t .. base + 4t + 3
  int32_t x[4];
  int32_t tsum = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t i = base + 4 * threadIdx.x + k;
    x[k] = i < n ? cnt_cursor[i] : 0;
    tsum += x[k];
  }
  int32_t incl = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int32_t y = __shfl_up_sync(0xffffffffu, incl, o);
    if ((threadIdx.x & 31) >= o) incl += y;
  }
  if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
  __syncthreads();
  int32_t off = block_sums[blockIdx.x] + incl - tsum;
  for (int w = 0; w < (threadIdx.x >> 5); ++w) off += s_warp[w];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t i = base + 4 * threadIdx.x + k;
    if (i <= n) {   // i == n: the total (x = 0 there)
      vstart[i] = off;
      if (i < n) cnt_cursor[i] = off;
    }
    off += x[k];
  }
}

// entry of the CSR: (ray of this call, barycentric weight of the corner)
__global__ void csr_fill_kernel(const FieldParams p) {
  const int64_t total = p.n_rays * 3;
  int2* pairs = reinterpret_cast<int2*>(p.pairs);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t ray = i / 3;
    const int j = static_cast<int>(i - ray * 3);
    const int64_t e = src_row(p, ray) * 3 + j;
    const uint32_t v = vertex_of(p, e);
    if (static_cast<int64_t>(v) < p.n_rows) {
      const int pos = atomicAdd(&p.vcursor[v], 1);
      pairs[pos] = make_int2(static_cast<int>(ray), __float_as_int(p.bary[e]));
    }
  }
}

// Lists of up to 32 pairs (the common case: a vertex is hit ~3 B / V times) are first SORTED by (ray, weight bits) with a
// warp-wide bitonic network and then summed in that order with plain fp32 FMAs; longer lists are summed in 64-bit fixed
// point, where the order does not matter.  Either way the result depends on the set of pairs only.
// measured variants (scripts/variant_sweep.sh; default 53 us): 8 rows in flight per warp 62-64 us, 6 CTAs per SM 62 us,
// non-coherent no-allocate loads 59 us; one half-warp per vertex (two vertices per warp) 72 us
#ifndef VG_BATCH
#define VG_BATCH 4
#endif
#ifndef VG_MINB
#define VG_MINB 4
#endif
#ifndef VG_NC
#define VG_NC 0
#endif
__global__ void __launch_bounds__(256, VG_MINB) vertex_gather_kernel(const FieldParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const Geo g = p.geo;
  if (p.debug & DEBUG_PDL_EARLY) pdl_launch_dependents();
  pdl_wait();   // the stash (field_rays); the incidence lists are ordered before this launch by an event
  const int2* pairs = reinterpret_cast<const int2*>(p.pairs);
  // lanes 0..15: dz_0 row of the ray, lanes 16..31: dz_skip row; 16-byte chunk (lane & 15) of the 128 features
  const size_t blob_off = lane < 16 ? stash_off_dz(g, 0) : stash_off_dz(g, g.skip);
  const uint32_t panel = static_cast<uint32_t>(lane & 15) >> 3;
  const uint32_t chunk = static_cast<uint32_t>(lane) & 7u;
  const int64_t v_end = static_cast<int64_t>(p.ntiles_v) * TILE_M;
  auto row_ptr = [&](int ray) -> const uint4* {
    const uint32_t row = static_cast<uint32_t>(ray) & 127u;
    return reinterpret_cast<const uint4*>(p.stash + static_cast<size_t>(ray >> 7) * p.stash_stride + blob_off + panel * PANEL_BYTES + row * 128u +
                                          ((chunk ^ (row & 7u)) << 4));
  };
  // A vertex is a chain of dependent reads (list bounds -> pairs -> stashed rows), so the bounds of the vertex after next and
  // the pairs of the next vertex are requested while this vertex's rows are summed.
  auto load_range = [&](int64_t vv, int& b, int& e) {
    b = 0; e = 0;
    if (vv < p.n_rows) { b = p.vstart[vv]; e = p.vstart[vv + 1]; }
  };
  auto load_pair = [&](int b, int e) -> int2 {
    const int n_ = e - b;
    return (n_ > 0 && n_ <= 32 && lane < n_) ? pairs[b + lane] : make_int2(0x7FFFFFFF, 0);
  };
  int beg, end, nbeg, nend;
  load_range(warp0, beg, end);
  load_range(warp0 + nwarps, nbeg, nend);
  int2 cur = load_pair(beg, end);
  for (int64_t v = warp0; v < v_end; v += nwarps) {
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = 0.f;
    int nnbeg, nnend;
    load_range(v + 2 * nwarps, nnbeg, nnend);
    const int2 nxt = load_pair(nbeg, nend);
    const int n = end - beg;
    if (n > 0 && n <= 32) {
      // ---- sort the list (one pair per lane; empty lanes sort to the end), then sum in order ----------------
      const int2 mine = cur;
      unsigned long long key = (static_cast<unsigned long long>(static_cast<uint32_t>(mine.x)) << 32) | static_cast<uint32_t>(mine.y);
#pragma unroll
      for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
          const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, j);
          const bool up = (lane & k) == 0;              // ascending block
          const bool lower = (lane & j) == 0;           // this lane keeps the smaller of the pair when ascending
          const bool take_min = (up == lower);
          key = take_min ? (key < other ? key : other) : (key > other ? key : other);
        }
      }
      for (int u0 = 0; u0 < n; u0 += VG_BATCH) {
        uint4 q[VG_BATCH];
        float w[VG_BATCH];
#pragma unroll
        for (int u = 0; u < VG_BATCH; ++u) {
          const unsigned long long ku = __shfl_sync(0xffffffffu, key, (u0 + u) & 31);
          w[u] = (u0 + u < n) ? __uint_as_float(static_cast<uint32_t>(ku)) : 0.f;
          q[u] = (u0 + u < n) ? (VG_NC ? ldg_nc_u4(row_ptr(static_cast<int>(ku >> 32))) : *row_ptr(static_cast<int>(ku >> 32)))
                              : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < VG_BATCH; ++u) {
          const uint32_t vv[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&vv[i]));
            f[2 * i] = fmaf(w[u], h.x, f[2 * i]);
            f[2 * i + 1] = fmaf(w[u], h.y, f[2 * i + 1]);
          }
        }
      }
    } else if (n > 32) {
      // ---- long list: 64-bit fixed point (2^-32 quantum, far below the fp16 resolution of the result) ---------
      long long acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0;
      for (int e0 = beg; e0 < end; e0 += 32) {
        int2 mine = make_int2(0, 0);
        if (e0 + lane < end) mine = pairs[e0 + lane];
        const int m = min(32, end - e0);
        for (int u0 = 0; u0 < m; u0 += 4) {
          uint4 q[4];
          float w[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int ray = __shfl_sync(0xffffffffu, mine.x, (u0 + u) & 31);
            w[u] = (u0 + u < m) ? __int_as_float(__shfl_sync(0xffffffffu, mine.y, (u0 + u) & 31)) : 0.f;
            q[u] = (u0 + u < m) ? *row_ptr(ray) : make_uint4(0u, 0u, 0u, 0u);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint32_t vv[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&vv[i]));
              acc[2 * i] += __float2ll_rn(w[u] * h.x * 4294967296.f);
              acc[2 * i + 1] += __float2ll_rn(w[u] * h.y * 4294967296.f);
            }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = static_cast<float>(acc[i]) * 2.3283064365386963e-10f;   // 2^-32
    }
    const uint32_t row = static_cast<uint32_t>(v) & 127u;
    uint8_t* dst = p.g_image + static_cast<size_t>(v >> 7) * GIMG_TILE_BYTES + (lane < 16 ? 0u : static_cast<uint32_t>(BLOB_BYTES)) +
                   panel * PANEL_BYTES + row * 128u + ((chunk ^ (row & 7u)) << 4);
    *reinterpret_cast<uint4*>(dst) = make_uint4(pack_half2_sat(f[0], f[1]), pack_half2_sat(f[2], f[3]), pack_half2_sat(f[4], f[5]),
                                                pack_half2_sat(f[6], f[7]));
    beg = nbeg; end = nend; cur = nxt;
    nbeg = nnbeg; nend = nnend;
  }
}

// CSR of this call's rays (depends on the inputs only: vids / ray_idxs / bary)
cudaError_t launch_csr_build(const FieldParams& p, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(p.vcursor, 0, static_cast<size_t>(p.n_rows + 1) * 4, stream);
  if (e != cudaSuccess) return e;
  const int64_t total = p.n_rays * 3;
  const int threads = 256;
  const unsigned blocks = static_cast<unsigned>(std::min<int64_t>((total + threads - 1) / threads, 148 * 8));
  csr_count_kernel<<<blocks, threads, 0, stream>>>(p);
  const int nscan = static_cast<int>((p.n_rows + SCAN_BLOCK) / SCAN_BLOCK);   // covers index n_rows (the total) as well
  csr_scan_sums_kernel<<<static_cast<unsigned>(nscan), 256, 0, stream>>>(p.vcursor, p.n_rows, p.scan_tmp);
  csr_scan_blocks_kernel<<<1, 1024, 0, stream>>>(p.scan_tmp, nscan);
  csr_scan_apply_kernel<<<static_cast<unsigned>(nscan), 256, 0, stream>>>(p.vcursor, p.n_rows, p.scan_tmp, p.vstart);
  csr_fill_kernel<<<blocks, threads, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_vertex_gather(const FieldParams& p, int nsm, cudaStream_t stream) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(nsm * VG_MINB));
  cfg.blockDim = dim3(256);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  add_pdl_attr(cfg, attr);
  return cudaLaunchKernelEx(&cfg, vertex_gather_kernel, p);
}

}  // namespace inf
