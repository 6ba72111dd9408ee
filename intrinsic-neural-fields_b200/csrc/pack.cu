// Operand preparation kernels: fp32 master weights -> swizzled fp16 operand images, and the
// eigenfunction-table re-layout (pad k -> k_pad, optional fp16 storage).
#include "common.cuh"
#include "ptx.cuh"

namespace inf {


// Rows of MMA layer l that really exist (the rest of the 128-row image is zero padding).
__device__ __forceinline__ int layer_rows(const Geo& g, int l) {
  if (g.view && l == g.nhid) return g.bneck;
  if (g.view && l == g.nb) return g.dh;
  return HID;
}

// One thread per 16-bit element of the blob section + one per f32 of the tail section.
__global__ void pack_weights_kernel(const PackParams pp) {
  const Geo g = pp.geo;
  const int64_t nhalf = static_cast<int64_t>(g.nblobs) * (BLOB_BYTES / 2);
  const int nbias = (g.nb + 1) * HID;
  const int nf32 = packed_f32_count(g);
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < nhalf) {
    const int blob = static_cast<int>(i / (BLOB_BYTES / 2));
    const int e = static_cast<int>(i % (BLOB_BYTES / 2));
    float v = 0.f;
    uint32_t off;
    if (blob < g.nchunks) {
      // [256 rows][64 k]: row < 128 -> W0, else Wy
      const int row = e >> 6, c = e & 63;
      const int kcol = blob * KC + c;
      if (kcol < g.k) v = (row < HID) ? pp.w[0][static_cast<size_t>(row) * g.k + kcol]
                                      : pp.wy[static_cast<size_t>(row - HID) * g.k + kcol];
      if (pp.col_exp != nullptr) v = ldexpf(v, -static_cast<int>(pp.col_exp[kcol]));   // the table holds E[:,j] * 2^e_j
      off = sw128_off(row, c);
    } else {
      // [128 rows n][128 cols j] as two 64-column panels
      const int row = e >> 7, col = e & 127;
      const int hb = blob - g.nchunks;
      if (hb < g.nhid) {
        const int l = hb + 1;                                   // W_l[n=row][j=col]
        if (row < layer_rows(g, l)) v = pp.w[l][row * HID + col];
      } else if (hb == g.nhid) {                                 // Wd1[:, :bneck]
        if (row < g.dh && col < g.bneck) v = pp.w[g.nb][static_cast<size_t>(row) * (g.bneck + g.edim) + col];
      } else {                                                   // Wd1[:, bneck:bneck+E] in the first panel
        if (row < g.dh && col < g.edim) v = pp.w[g.nb][static_cast<size_t>(row) * (g.bneck + g.edim) + g.bneck + col];
      }
      off = static_cast<uint32_t>(col >> 6) * PANEL_BYTES + sw128_off(row, col & 63);
    }
    *reinterpret_cast<__half*>(pp.packed + static_cast<size_t>(blob) * BLOB_BYTES + off) = __float2half_rn(v);
  } else if (i < nhalf + nf32) {
    const int e = static_cast<int>(i - nhalf);
    float v = 0.f;
    if (e < nbias) {
      const int l = e / HID, n = e % HID;
      if (n < layer_rows(g, l)) {
        v = pp.b[l][n];
        if (l == g.skip) v += pp.by[n];
      }
    } else if (e < nbias + OUT_DIM * HID) {
      const int c = (e - nbias) / HID, j = (e - nbias) % HID;
      if (j < g.wout) v = pp.wlast[c * g.wout + j];
    } else {
      const int c = e - nbias - OUT_DIM * HID;
      v = (c < OUT_DIM) ? pp.blast[c] : 0.f;
    }
    reinterpret_cast<float*>(pp.packed + packed_f32_offset(g))[e] = v;
  }
}

void launch_pack_weights(const PackParams& pp, cudaStream_t stream) {
  const Geo g = pp.geo;
  const int64_t n = static_cast<int64_t>(g.nblobs) * (BLOB_BYTES / 2) + packed_f32_count(g);
  const int threads = 256;
  pack_weights_kernel<<<static_cast<unsigned>((n + threads - 1) / threads), threads, 0, stream>>>(pp);
}

template <typename OutT>
__global__ void relayout_table_kernel(const float* __restrict__ src, int64_t n_rows, int k, int64_t src_stride,
                                      OutT* __restrict__ dst, int k_pad, const int8_t* __restrict__ col_exp) {
  const int64_t total = n_rows * k_pad;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / k_pad;
    const int c = static_cast<int>(i % k_pad);
    float v = (c < k) ? src[r * src_stride + c] : 0.f;
    if (col_exp != nullptr) v = ldexpf(v, static_cast<int>(col_exp[c]));
    if constexpr (sizeof(OutT) == 2) dst[i] = __float2half_rn(v);
    else dst[i] = v;
  }
}

// per-column max |x| as float bits (non-negative floats order like their bit patterns), then the exponent
__global__ void __launch_bounds__(256) column_maxabs_kernel(const float* __restrict__ src, int64_t n_rows, int k, int64_t src_stride,
                                                            unsigned int* __restrict__ maxbits) {
  // block = 8 row lanes x 32 columns; grid.x walks column groups, grid.y row stripes
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rl = threadIdx.x >> 5;
  unsigned int m = 0u;
  if (c < k)
    for (int64_t r = static_cast<int64_t>(blockIdx.y) * 8 + rl; r < n_rows; r += static_cast<int64_t>(gridDim.y) * 8)
      m = max(m, __float_as_uint(fabsf(src[r * src_stride + c])));
  if (c < k && m != 0u) atomicMax(maxbits + c, m);
}
__global__ void column_exponent_kernel(const unsigned int* __restrict__ maxbits, int k, int k_pad, int8_t* __restrict__ col_exp) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= k_pad) return;
  int e = 0;
  if (c < k) {
    const float m = __uint_as_float(maxbits[c]);
    if (m > 0.f && isfinite(m) && (m < 0.015625f || m > 64.f)) {   // outside [2^-6, 2^6]: bring max|x| into [0.5, 1)
      int p;
      frexpf(m, &p);
      e = -p;
      e = e < -100 ? -100 : (e > 100 ? 100 : e);
    }
  }
  col_exp[c] = static_cast<int8_t>(e);
}
void launch_column_exponents(const float* src, int64_t n_rows, int k, int64_t src_stride, int k_pad, int8_t* col_exp,
                             unsigned int* scratch, cudaStream_t stream) {
  cudaMemsetAsync(scratch, 0, static_cast<size_t>(k_pad) * 4, stream);
  if (n_rows > 0) {
    const dim3 grid(static_cast<unsigned>((k + 31) / 32), static_cast<unsigned>(n_rows >= 8 * 64 ? 64 : (n_rows + 7) / 8));
    column_maxabs_kernel<<<grid, 256, 0, stream>>>(src, n_rows, k, src_stride, scratch);
  }
  column_exponent_kernel<<<(k_pad + 255) / 256, 256, 0, stream>>>(scratch, k, k_pad, col_exp);
}

void launch_relayout(const float* src, int64_t n_rows, int k, int64_t src_stride, void* dst, int k_pad, int dst_dtype,
                     const int8_t* col_exp, cudaStream_t stream) {
  const int64_t total = n_rows * k_pad;
  if (total == 0) return;
  const int threads = 256;
  const int64_t want = (total + threads - 1) / threads;
  const unsigned blocks = static_cast<unsigned>(want < 148 * 16 ? want : 148 * 16);
  if (dst_dtype == INF_TABLE_F16)
    relayout_table_kernel<__half><<<blocks, threads, 0, stream>>>(src, n_rows, k, src_stride, reinterpret_cast<__half*>(dst), k_pad, col_exp);
  else
    relayout_table_kernel<float><<<blocks, threads, 0, stream>>>(src, n_rows, k, src_stride, reinterpret_cast<float*>(dst), k_pad, col_exp);
}

// ---------------------------------------------------------------------------------------------
// UMMA self-test (inf_debug_umma): one CTA, D(128,n) = A(128,k) * B(n,k)^T for the three operand
// layout variants the field kernels rely on.  k multiple of 64 (<= 128), n multiple of 64 or 16.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(int mode, int n, int k, const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ d) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  // layout: A image 32 KB | B image 64 KB | bar | tmem ptr
  const uint32_t SA = 0, SB = 32768, SBAR = 32768 + 65536, SPTR = SBAR + 8;
  const int tid = threadIdx.x, warp = tid >> 5;
  const bool a_mn = mode == 1 || mode == 2, b_mn = mode >= 2;   // 3: A K-major, B MN-major
  __half* sa = reinterpret_cast<__half*>(smem + SA);
  __half* sb = reinterpret_cast<__half*>(smem + SB);
  // K-major image: element (row m, col kk) -> panel kk/64 (stride rows*128 B), sw128(row, kk%64)
  // MN-major image: element (m, kk) stored at "row" kk, "col" m: panel m/64 (stride k*128 B), sw128(kk, m%64)
  for (int i = tid; i < 128 * k; i += 128) {
    const int m = a_mn ? i % 128 : i / k, kk = a_mn ? i / 128 : i % k;
    const float v = a[i];  // K-major source (128,k); MN-major source (k,128)
    uint32_t off = a_mn ? static_cast<uint32_t>(m >> 6) * (k * 128u) + sw128_off(kk, m & 63)
                        : static_cast<uint32_t>(kk >> 6) * (128u * 128u) + sw128_off(m, kk & 63);
    *reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(sa) + off) = __float2half_rn(v);
  }
  for (int i = tid; i < n * k; i += 128) {
    const int m = b_mn ? i % n : i / k, kk = b_mn ? i / n : i % k;
    const float v = b[i];
    uint32_t off = b_mn ? static_cast<uint32_t>(m >> 6) * (k * 128u) + sw128_off(kk, m & 63)
                        : static_cast<uint32_t>(kk >> 6) * (n * 128u) + sw128_off(m, kk & 63);
    *reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(sb) + off) = __float2half_rn(v);
  }
  if (tid == 0) { mbar_init(smem_base + SBAR, 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc(smem_base + SPTR, 256); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + SPTR);
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_f16(128, n, a_mn ? 1 : 0, b_mn ? 1 : 0);
    for (int kk = 0; kk < k / 16; ++kk) {
      uint64_t ad, bd;
      if (a_mn) ad = umma_desc_sw128(smem_base + SA + kk * 2048u, k * 128u, 1024);
      else ad = umma_desc_sw128(smem_base + SA + (kk >> 2) * (128u * 128u) + (kk & 3) * 32u, 0, 1024);
      if (b_mn) bd = umma_desc_sw128(smem_base + SB + kk * 2048u, k * 128u, 1024);
      else bd = umma_desc_sw128(smem_base + SB + (kk >> 2) * (n * 128u) + (kk & 3) * 32u, 0, 1024);
      umma_f16(tmem_base, ad, bd, idesc, kk > 0 ? 1u : 0u);
    }
    umma_commit(smem_base + SBAR);
  }
  mbar_wait(smem_base + SBAR, 0, nullptr, 0x90);
  tc_fence_after();
  const uint32_t tlane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  for (int c = 0; c < n; ++c) {
    uint32_t v;
    tmem_ld1(tlane + c, v);
    tmem_ld_wait();
    d[tid * n + c] = __uint_as_float(v);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

void launch_umma_probe(int mode, int n, int k, const float* a, const float* b, float* d, cudaStream_t stream) {
  const size_t smem = 32768 + 65536 + 64 + 1024;
  cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  umma_probe_kernel<<<1, 128, smem, stream>>>(mode, n, k, a, b, d);
}

}  // namespace inf

namespace inf {

// ---------------------------------------------------------------------------------------------
// Materialising gather (mesh.py:313-324 as one kernel): x[b,:] = sum_j bary[b,j] * E[vids[b,j],:].
// Kept for callers that need the dense (B,k) embedding (baking, notebooks); the hot path never
// builds it.  One warp per ray, 16-byte loads along the row.
// ---------------------------------------------------------------------------------------------
template <typename TableT>
__global__ void gather_interpolate_kernel(const TableT* __restrict__ table, int k, int k_pad, const void* __restrict__ vids,
                                          int idx64, const float* __restrict__ bary, int64_t n_rays, float* __restrict__ out,
                                          const int8_t* __restrict__ col_exp) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t b = warp; b < n_rays; b += nwarps) {
    int64_t v[3];
    float w[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      v[j] = idx64 ? reinterpret_cast<const long long*>(vids)[b * 3 + j] : reinterpret_cast<const int*>(vids)[b * 3 + j];
      w[j] = bary[b * 3 + j];
    }
    for (int c0 = lane * 4; c0 < k_pad; c0 += 128) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const TableT* row = table + v[j] * k_pad + c0;
        float e[4];
        if constexpr (sizeof(TableT) == 4) {
          const float4 q = *reinterpret_cast<const float4*>(row);
          e[0] = q.x; e[1] = q.y; e[2] = q.z; e[3] = q.w;
        } else {
          const uint2 q = *reinterpret_cast<const uint2*>(row);
          const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&q.x));
          const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&q.y));
          e[0] = lo.x; e[1] = lo.y; e[2] = hi.x; e[3] = hi.y;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = (j == 0) ? w[0] * e[i] : fmaf(w[j], e[i], acc[i]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (c0 + i < k) out[b * k + c0 + i] = col_exp != nullptr ? ldexpf(acc[i], -static_cast<int>(col_exp[c0 + i])) : acc[i];
    }
  }
}

void launch_gather_interpolate(const void* table, int table_f16, int k, int k_pad, const void* vids, int idx64,
                               const float* bary, int64_t n_rays, float* out, const int8_t* col_exp, cudaStream_t stream) {
  if (n_rays == 0) return;
  const int threads = 256;
  const int64_t want = (n_rays * 32 + threads - 1) / threads;
  const unsigned blocks = static_cast<unsigned>(want < 148 * 8 ? want : 148 * 8);
  if (table_f16)
    gather_interpolate_kernel<__half><<<blocks, threads, 0, stream>>>(reinterpret_cast<const __half*>(table), k, k_pad, vids,
                                                                     idx64, bary, n_rays, out, col_exp);
  else
    gather_interpolate_kernel<float><<<blocks, threads, 0, stream>>>(reinterpret_cast<const float*>(table), k, k_pad, vids, idx64,
                                                                    bary, n_rays, out, col_exp);
}

}  // namespace inf
