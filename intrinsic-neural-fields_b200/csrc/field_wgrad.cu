// Weight-gradient kernel + deterministic partial reduction, sm_100a.
//
// dW_l = sum_rays dz_l[ray,:]^T * input_l[ray,:] is a GEMM whose contraction dimension is the ray
// index.  Each CTA owns a column slice of the gradient for the whole launch and keeps it in TMEM
// (512 fp32 columns x 128 lanes) while it walks its share of the ray tiles:
//   "big" slice s < nbig : 256 columns j of dW0 (TMEM [0,256)) and of dWy (TMEM [256,512));
//                          the x operand is re-gathered + interpolated by the producer warps
//                          exactly as in the forward kernel (it is never stored in HBM);
//   "hidden" slice       : up to four 128x128 gradients dW_l = dz_l^T h_{l-1} (l = 1..L-2).
// Both operands are read MN-major (UMMA transposes on the fly) from the same swizzled images the
// forward/backward kernel stashed, so no transposed copy is ever materialised.
// Ray tiles are dealt round-robin to `groups` CTAs per slice; every CTA finally dumps its
// accumulators to a partial buffer and field_reduce_kernel sums the partials in a fixed order
// (deterministic -> the N-GPU == 1-GPU check can be bit-stable per rank).
#include "common.cuh"
#include "ptx.cuh"

namespace inf {

namespace {

constexpr int NSTAGE = 3;
constexpr uint32_t STAGE_BYTES = 2 * BLOB_BYTES;  // [A 32 KB][B 32 KB]
constexpr uint32_t SM_STAGES = 0;
constexpr uint32_t SM_VID = SM_STAGES + NSTAGE * STAGE_BYTES;  // u32 [4][192]
constexpr uint32_t SM_BARY = SM_VID + 4 * 192 * 4;
constexpr uint32_t SM_BARS = SM_BARY + 4 * 192 * 4;
constexpr uint32_t SM_TMEMPTR = SM_BARS + 16 * 8;
constexpr uint32_t SM_TOTAL = SM_TMEMPTR + 16;
constexpr uint32_t SM_ALLOC = SM_TOTAL + 1024;

constexpr int BAR_FULL = 0;
constexpr int BAR_EMPTY = NSTAGE;
constexpr int BAR_DONE = 2 * NSTAGE;

constexpr uint32_t HALF_PANEL = 64 * 128;  // [64 rays][64 fp16] = 8 KB

// Same lane mapping as the forward producers; rows are the 64 rays of a half tile.
template <typename TableT>
struct SliceLoader;

template <>
struct SliceLoader<float> {
  float4 a[3], b[3];
  __device__ __forceinline__ void load(const float* base, const int64_t (&roff)[3], int koff, int s) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float* p = base + roff[j] + koff + 4 * s;
      a[j] = ldg_nc_f4(p);
      b[j] = ldg_nc_f4(p + 32);
    }
  }
  __device__ __forceinline__ void store(uint32_t panel, int r, int s, const float (&w)[3]) const {
    float x0 = w[0] * a[0].x, x1 = w[0] * a[0].y, x2 = w[0] * a[0].z, x3 = w[0] * a[0].w;
    float y0 = w[0] * b[0].x, y1 = w[0] * b[0].y, y2 = w[0] * b[0].z, y3 = w[0] * b[0].w;
#pragma unroll
    for (int j = 1; j < 3; ++j) {
      x0 = fmaf(w[j], a[j].x, x0); x1 = fmaf(w[j], a[j].y, x1);
      x2 = fmaf(w[j], a[j].z, x2); x3 = fmaf(w[j], a[j].w, x3);
      y0 = fmaf(w[j], b[j].x, y0); y1 = fmaf(w[j], b[j].y, y1);
      y2 = fmaf(w[j], b[j].z, y2); y3 = fmaf(w[j], b[j].w, y3);
    }
    const uint32_t row = panel + static_cast<uint32_t>(r) * 128u;
    const uint32_t sw = static_cast<uint32_t>(r) & 7u;
    const uint32_t half = static_cast<uint32_t>(s & 1) << 3;
    st_shared_v2(row + ((static_cast<uint32_t>(s >> 1) ^ sw) << 4) + half, pack_half2(x0, x1), pack_half2(x2, x3));
    st_shared_v2(row + ((static_cast<uint32_t>(4 + (s >> 1)) ^ sw) << 4) + half, pack_half2(y0, y1),
                 pack_half2(y2, y3));
  }
};

template <>
struct SliceLoader<__half> {
  uint4 a[3];
  __device__ __forceinline__ void load(const __half* base, const int64_t (&roff)[3], int koff, int s) {
#pragma unroll
    for (int j = 0; j < 3; ++j) a[j] = ldg_nc_u4(base + roff[j] + koff + 8 * s);
  }
  __device__ __forceinline__ void store(uint32_t panel, int r, int s, const float (&w)[3]) const {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const uint32_t u[4] = {a[j].x, a[j].y, a[j].z, a[j].w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&u[i]));
        if (j == 0) { acc[2 * i] = w[0] * f.x; acc[2 * i + 1] = w[0] * f.y; }
        else { acc[2 * i] = fmaf(w[j], f.x, acc[2 * i]); acc[2 * i + 1] = fmaf(w[j], f.y, acc[2 * i + 1]); }
      }
    }
    const uint32_t addr = panel + static_cast<uint32_t>(r) * 128u +
                          ((static_cast<uint32_t>(s) ^ (static_cast<uint32_t>(r) & 7u)) << 4);
    st_shared_v4(addr, pack_half2(acc[0], acc[1]), pack_half2(acc[2], acc[3]), pack_half2(acc[4], acc[5]),
                 pack_half2(acc[6], acc[7]));
  }
};

}  // namespace

template <typename TableT>
__global__ void __launch_bounds__(NUM_THREADS, 1) field_wgrad_kernel(const FieldParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const Geo g = p.geo;
  const int L = g.L;
  // CTA -> (slice, ray group): big slices first (groups_big CTAs each), then hidden slices
  const int nbig_ctas = g.nbig * p.groups_big;
  const bool big = static_cast<int>(blockIdx.x) < nbig_ctas;
  const int groups = big ? p.groups_big : p.groups_hid;
  const int rel = big ? blockIdx.x : blockIdx.x - nbig_ctas;
  const int slice = (big ? 0 : g.nbig) + rel / groups;
  const int grp = rel % groups;
  const int ncols = big ? min(256, g.k_pad - 256 * slice) : 0;        // big: columns of this slice
  const int hl0 = big ? 0 : 1 + 4 * (slice - g.nbig);                  // hidden: first layer
  const int hn = big ? 0 : min(4, (L - 2) - 4 * (slice - g.nbig));    // hidden: layer count

  auto bar = [&](int i) -> uint32_t { return smem_base + SM_BARS + 8u * static_cast<uint32_t>(i); };

  if (tid == 0) {
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(bar(BAR_FULL + i), big ? 1 + PROD_WARPS : 1);
      mbar_init(bar(BAR_EMPTY + i), 1);
    }
    mbar_init(bar(BAR_DONE), 1);
    mbar_fence_init();
  }
  if (warp == MMA_WARP) {
    tmem_alloc(smem_base + SM_TMEMPTR, WG_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + SM_TMEMPTR);

  const int ntiles = p.ntiles;
  const int my_tiles = (ntiles > grp) ? (ntiles - grp + groups - 1) / groups : 0;
  const int units_per_tile = big ? 2 : hn;
  const int nunits = my_tiles * units_per_tile;

  if (warp >= EPI_WARPS && warp < EPI_WARPS + PROD_WARPS) {
    // ======================= gather producers (big slices only) ================================
    if (big) {
      const int pt = tid - EPI_THREADS;
      const int pw = pt >> 5;
      const int s = lane & 7, gq = lane >> 3;
      const TableT* table = reinterpret_cast<const TableT*>(p.table);
      uint32_t* s_vid = reinterpret_cast<uint32_t*>(smem + SM_VID);
      float* s_bary = reinterpret_cast<float*>(smem + SM_BARY);
      const int npanels = ncols / KC;
      int stage = 0;
      uint32_t phase = 0;
      // unit metadata ring (4 deep), filled two units ahead of the consumer
      auto load_meta = [&](int u) {
        if (u < nunits && pt < 192) {
          const int tile = grp + (u >> 1) * groups;
          const int64_t ray0 = static_cast<int64_t>(tile) * TILE_M + 64 * (u & 1);
          const int64_t e = ray0 * 3 + pt;
          uint32_t v = 0;
          float w = 0.f;
          if (e < p.n_rays * 3) {
            if (p.dense) { v = static_cast<uint32_t>(ray0 + pt / 3); w = (pt % 3 == 0) ? 1.f : 0.f; }
            else {
              v = p.idx64 ? static_cast<uint32_t>(reinterpret_cast<const long long*>(p.vids)[e])
                          : static_cast<uint32_t>(reinterpret_cast<const int*>(p.vids)[e]);
              w = p.bary[e];
            }
          }
          s_vid[(u & 3) * 192 + pt] = v;
          s_bary[(u & 3) * 192 + pt] = w;
        }
      };
      // software-pipelined gather (see field_main.cu): pass = 4 rays per warp of one 64-column panel;
      // the loads of pass n+2 are issued as soon as pass n is consumed, across panels and units.
      struct Cursor { int u, pn, q; };
      auto advance = [&](Cursor& k) {
        if (++k.q == 2) { k.q = 0; if (++k.pn == npanels) { k.pn = 0; ++k.u; } }
      };
      auto pass_row = [&](int q) { return 8 * pw + 2 * q + (gq >> 1) + 4 * (gq & 1); };
      auto issue = [&](const Cursor& k, SliceLoader<TableT>& ld) {
        if (k.u >= nunits) return;
        const int e = (k.u & 3) * 192 + 3 * pass_row(k.q);
        int64_t roff[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) roff[j] = static_cast<int64_t>(s_vid[e + j]) * p.row_stride;
        ld.load(table, roff, 256 * slice + KC * k.pn, s);
      };
      auto consume = [&](int u, int q, const SliceLoader<TableT>& ld, uint32_t panel) {
        const int r = pass_row(q);
        const int e = (u & 3) * 192 + 3 * r;
        const float w[3] = {s_bary[e], s_bary[e + 1], s_bary[e + 2]};
        ld.store(panel, r, s, w);
      };
      SliceLoader<TableT> ld0, ld1;
      Cursor Lc{0, 0, 0};
      load_meta(0);
      load_meta(1);
      named_bar_sync(2, PROD_THREADS);
      issue(Lc, ld0); advance(Lc);
      issue(Lc, ld1); advance(Lc);
      for (int u = 0; u < nunits; ++u) {
        load_meta(u + 2);
        named_bar_sync(2, PROD_THREADS);
        mbar_wait(bar(BAR_EMPTY + stage), phase ^ 1u, p.err, 0x50);
        const uint32_t bpart = smem_base + SM_STAGES + static_cast<uint32_t>(stage) * STAGE_BYTES + BLOB_BYTES;
        for (int pn = 0; pn < npanels; ++pn) {
          const uint32_t panel = bpart + static_cast<uint32_t>(pn) * HALF_PANEL;
          consume(u, 0, ld0, panel); issue(Lc, ld0); advance(Lc);
          consume(u, 1, ld1, panel); issue(Lc, ld1); advance(Lc);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(BAR_FULL + stage));
        if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == LOAD_WARP) {
    // ======================= stash loader ======================================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = 0; u < nunits; ++u) {
        const int tile = grp + (u / units_per_tile) * groups;
        const uint8_t* st = p.stash + static_cast<size_t>(tile) * g.nslots * BLOB_BYTES;
        mbar_wait(bar(BAR_EMPTY + stage), phase ^ 1u, p.err, 0x60);
        const uint32_t sbase = smem_base + SM_STAGES + static_cast<uint32_t>(stage) * STAGE_BYTES;
        const uint32_t fb = bar(BAR_FULL + stage);
        if (big) {
          const int hh = u & 1;
          mbar_arrive_expect_tx(fb, BLOB_BYTES);
          const uint8_t* dz0 = st + static_cast<size_t>(stash_slot_dz(g, 0)) * BLOB_BYTES + hh * HALF_PANEL;
          const uint8_t* dzs = st + static_cast<size_t>(stash_slot_dz(g, g.skip)) * BLOB_BYTES + hh * HALF_PANEL;
          bulk_g2s(sbase + 0 * HALF_PANEL, dz0, HALF_PANEL, fb);
          bulk_g2s(sbase + 1 * HALF_PANEL, dz0 + PANEL_BYTES, HALF_PANEL, fb);
          bulk_g2s(sbase + 2 * HALF_PANEL, dzs, HALF_PANEL, fb);
          bulk_g2s(sbase + 3 * HALF_PANEL, dzs + PANEL_BYTES, HALF_PANEL, fb);
        } else {
          const int l = hl0 + (u % units_per_tile);
          mbar_arrive_expect_tx(fb, 2 * BLOB_BYTES);
          bulk_g2s(sbase, st + static_cast<size_t>(stash_slot_dz(g, l)) * BLOB_BYTES, BLOB_BYTES, fb);
          bulk_g2s(sbase + BLOB_BYTES, st + static_cast<size_t>(stash_slot_h(g, l - 1)) * BLOB_BYTES, BLOB_BYTES, fb);
        }
        if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == MMA_WARP) {
    // ======================= MMA issuer ===========================================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t idesc_big = umma_idesc_f16(128, ncols > 0 ? ncols : 16, 1, 1);
      const uint32_t idesc_hid = umma_idesc_f16(128, 128, 1, 1);
      for (int u = 0; u < nunits; ++u) {
        mbar_wait(bar(BAR_FULL + stage), phase, p.err, 0x70);
        tc_fence_after();
        const uint32_t abase = smem_base + SM_STAGES + static_cast<uint32_t>(stage) * STAGE_BYTES;
        const uint32_t bbase = abase + BLOB_BYTES;
        if (big) {
          // 64 rays: D0[n, j] += dz0[ray, n] x[ray, j]; Dy[n, j] += dz_skip[ray, n] x[ray, j]
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint32_t ko = static_cast<uint32_t>(kk) * 2048u;
            const uint64_t bd = umma_desc_sw128(bbase + ko, HALF_PANEL, 1024);
            const uint32_t acc = (u > 0 || kk > 0) ? 1u : 0u;
            umma_f16(tmem_base + 0, umma_desc_sw128(abase + ko, HALF_PANEL, 1024), bd, idesc_big, acc);
            umma_f16(tmem_base + 256, umma_desc_sw128(abase + 2 * HALF_PANEL + ko, HALF_PANEL, 1024), bd, idesc_big, acc);
          }
        } else {
          const int li = u % units_per_tile;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint32_t ko = static_cast<uint32_t>(kk) * 2048u;
            umma_f16(tmem_base + 128u * static_cast<uint32_t>(li), umma_desc_sw128(abase + ko, PANEL_BYTES, 1024),
                     umma_desc_sw128(bbase + ko, PANEL_BYTES, 1024), idesc_hid, (u >= units_per_tile || kk > 0) ? 1u : 0u);
          }
        }
        umma_commit(bar(BAR_EMPTY + stage));
        if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
      }
      umma_commit(bar(BAR_DONE));
    }
  } else {
    // ======================= accumulator dump ====================================================
    mbar_wait(bar(BAR_DONE), 0, p.err, 0x80);
    tc_fence_after();
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    float* dst = p.partw + (static_cast<size_t>(blockIdx.x) * HID + tid) * WG_COLS;
    const int used = big ? 512 : 128 * hn;
    for (int c0 = 0; c0 < used; c0 += 32) {
      if (big && (c0 & 255) >= ncols) continue;
      uint32_t v[32];
      if (nunits > 0) {
        tmem_ld32(tlane + static_cast<uint32_t>(c0), v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<uint4*>(dst + c0 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, WG_COLS);
  }
}

size_t field_wgrad_smem_bytes() { return SM_ALLOC; }

template <typename TableT>
static cudaError_t launch_wgrad_one(const FieldParams& p, int grid, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(field_wgrad_kernel<TableT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(SM_ALLOC));
  if (e != cudaSuccess) return e;
  field_wgrad_kernel<TableT><<<grid, NUM_THREADS, SM_ALLOC, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_field_wgrad(const FieldParams& p, bool table_f16, int grid, cudaStream_t stream) {
  return table_f16 ? launch_wgrad_one<__half>(p, grid, stream) : launch_wgrad_one<float>(p, grid, stream);
}

// ---------------------------------------------------------------------------------------------
// Partial reduction: one thread per element of the flat gradient vector, fixed summation order.
// ---------------------------------------------------------------------------------------------
__global__ void field_reduce_kernel(const ReduceParams rp) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i == 0 && rp.loss_sums != nullptr) {
    float a = 0.f, b = 0.f;
    for (int c = 0; c < rp.ncta1; ++c) {
      const float* pl = rp.part1 + static_cast<size_t>(c) * K1_PART_FLOATS + (INF_MAX_LAYERS - 1) * HID + 3 * HID;
      a += pl[4];
      b += pl[5];
    }
    rp.loss_sums[0] = a;
    rp.loss_sums[1] = b;
  }
  if (i >= rp.total) return;
  int sidx = 0;
  for (int s = 1; s < rp.nseg; ++s)
    if (i >= rp.seg[s].off) sidx = s;
  const ReduceSeg sg = rp.seg[sidx];
  const int64_t e = i - sg.off;
  const int row = static_cast<int>(e / sg.cols);
  const int col = static_cast<int>(e % sg.cols);
  float acc = 0.f;
  if (sg.kind == 0) {  // W0 / Wy: row n, column j
    const int slice = col >> 8;
    const int c = (col & 255) + (sg.a ? 256 : 0);
    for (int gI = 0; gI < rp.groups_big; ++gI)
      acc += rp.partw[(static_cast<size_t>(slice * rp.groups_big + gI) * HID + row) * WG_COLS + c];
  } else if (sg.kind == 1) {  // hidden W_l
    const int hs = (sg.a - 1) >> 2;
    const int c = 128 * ((sg.a - 1) & 3) + col;
    for (int gI = 0; gI < rp.groups_hid; ++gI)
      acc += rp.partw[(static_cast<size_t>(rp.geo.nbig * rp.groups_big + hs * rp.groups_hid + gI) * HID + row) * WG_COLS + c];
  } else if (sg.kind == 2) {  // bias of layer a (rows == HID entries, cols == 1)
    for (int c = 0; c < rp.ncta1; ++c) acc += rp.part1[static_cast<size_t>(c) * K1_PART_FLOATS + sg.a * HID + row];
  } else if (sg.kind == 3) {  // output-layer weight [3][128]
    for (int c = 0; c < rp.ncta1; ++c)
      acc += rp.part1[static_cast<size_t>(c) * K1_PART_FLOATS + (INF_MAX_LAYERS - 1) * HID + row * HID + col];
  } else {  // output-layer bias [3]
    for (int c = 0; c < rp.ncta1; ++c)
      acc += rp.part1[static_cast<size_t>(c) * K1_PART_FLOATS + (INF_MAX_LAYERS - 1) * HID + 3 * HID + row];
  }
  rp.flat[i] = acc * rp.scale;
}

cudaError_t launch_field_reduce(const ReduceParams& rp, cudaStream_t stream) {
  const int threads = 256;
  field_reduce_kernel<<<static_cast<unsigned>((rp.total + threads - 1) / threads), threads, 0, stream>>>(rp);
  return cudaGetLastError();
}

}  // namespace inf
