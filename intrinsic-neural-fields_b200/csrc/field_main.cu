// Forward (+ loss + dgrad chain) kernel of the intrinsic-neural-field path, sm_100a.
//
// Per 128-ray tile (persistent CTAs, one per SM):
//   producers (8 warps)  gather 3 table rows per ray with 16-byte loads, interpolate in fp32 with the
//                        barycentric weights, round to fp16 and store the [128 x 64] A-panel of the
//                        current k-chunk straight into the SWIZZLE_128B smem layout the tensor core
//                        reads -- the interpolated embedding (mesh.py:313-324) never exists in HBM;
//   loader (1 thread)    streams the weight operand images (32 KB blobs) with 1-D bulk copies (TMA);
//   MMA (1 thread)       tcgen05.mma kind::f16: GEMM1 = x * [W0;Wy]^T (N=256, K=k_pad) into TMEM
//                        columns [0,256), then the 128-wide hidden chain (model.py:98-112) and, when
//                        training, the dgrad chain and the bias-gradient "ones" MMAs;
//   epilogue (4 warps)   TMEM -> registers, bias + ReLU -> fp16 activation panel in smem (A operand of
//                        the next layer), final 128->3 layer + sigmoid (+ L1 loss gradient) on CUDA
//                        cores; when training, stashes h_l / dz_l tiles for the wgrad kernel.
#include "common.cuh"
#include "ptx.cuh"

namespace inf {

namespace {

constexpr int NSTAGE_A = 4;
constexpr int NSTAGE_B = 2;

// shared memory carve-up (bytes, relative to a 1024-aligned base)
constexpr uint32_t SM_A = 0;
constexpr uint32_t SM_B = SM_A + NSTAGE_A * PANEL_BYTES;          //  65536
constexpr uint32_t SM_H = SM_B + NSTAGE_B * BLOB_BYTES;           // 131072
constexpr uint32_t SM_ONES = SM_H + 2 * BLOB_BYTES;               // 196608
constexpr uint32_t SM_BIAS = SM_ONES + 2048;                      // f32 [L-1][128]
constexpr uint32_t SM_WLAST = SM_BIAS + (INF_MAX_LAYERS - 1) * HID * 4;
constexpr uint32_t SM_BLAST = SM_WLAST + OUT_DIM * HID * 4;
constexpr uint32_t SM_VID = SM_BLAST + 16;                        // u32 [3][384]
constexpr uint32_t SM_BARY = SM_VID + 3 * 384 * 4;                // f32 [3][384]
constexpr uint32_t SM_DZL = SM_BARY + 3 * 384 * 4;                // f32 [128][4]
constexpr uint32_t SM_RED = SM_DZL + TILE_M * 16;                 // f32 [8]
constexpr uint32_t SM_MASK = SM_RED + 32;                         // u32x4 [L-1][128]
constexpr uint32_t SM_BARS = SM_MASK + (INF_MAX_LAYERS - 1) * TILE_M * 16;
constexpr uint32_t SM_TMEMPTR = SM_BARS + 16 * 8;
constexpr uint32_t SM_TOTAL = SM_TMEMPTR + 16;
constexpr uint32_t SM_ALLOC = SM_TOTAL + 1024;  // slack for manual 1024-byte alignment

// barrier indices
constexpr int BAR_A_FULL = 0;                 // [NSTAGE_A]
constexpr int BAR_A_EMPTY = BAR_A_FULL + NSTAGE_A;
constexpr int BAR_B_FULL = BAR_A_EMPTY + NSTAGE_A;
constexpr int BAR_B_EMPTY = BAR_B_FULL + NSTAGE_B;
constexpr int BAR_MMA_DONE = BAR_B_EMPTY + NSTAGE_B;
constexpr int BAR_ACT_READY = BAR_MMA_DONE + 1;
static_assert(BAR_ACT_READY < 16, "barrier table overflow");

constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t TM_Z0 = 0;      // z0 / hidden accumulators / dh
constexpr uint32_t TM_ZY = 128;    // Ly(x) partial, then z_skip
constexpr uint32_t TM_DB = 256;    // bias-gradient accumulators, 16 columns per layer

constexpr uint32_t IDESC_N256 = umma_idesc_f16(128, 256, 0, 0);
constexpr uint32_t IDESC_N128 = umma_idesc_f16(128, 128, 0, 0);
constexpr uint32_t IDESC_DB = umma_idesc_f16(128, 16, 1, 0);  // A = dz^T (MN-major), B = ones

__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + __expf(-v)); }

// 32 fp32 values (columns 32*cb .. 32*cb+31 of row r) -> fp16, into the 2-panel activation image.
__device__ __forceinline__ void store_act_chunk(uint32_t hbuf, int r, int cb, const float (&h)[32]) {
  const uint32_t panel = hbuf + static_cast<uint32_t>(cb >> 1) * PANEL_BYTES + static_cast<uint32_t>(r) * 128u;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t chunk = static_cast<uint32_t>((cb & 1) * 4 + i);
    const uint32_t addr = panel + ((chunk ^ (static_cast<uint32_t>(r) & 7u)) << 4);
    st_shared_v4(addr, pack_half2(h[8 * i + 0], h[8 * i + 1]), pack_half2(h[8 * i + 2], h[8 * i + 3]),
                 pack_half2(h[8 * i + 4], h[8 * i + 5]), pack_half2(h[8 * i + 6], h[8 * i + 7]));
  }
}

template <typename TableT>
struct RowLoader;

// fp32 table: per k-chunk a ray needs 3 x 256 B; 8 lanes x 2 float4 per row.
template <>
struct RowLoader<float> {
  float4 a[3], b[3];
  __device__ __forceinline__ void load(const float* base, const int64_t (&roff)[3], int koff, int s) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float* p = base + roff[j] + koff + 4 * s;
      a[j] = ldg_nc_f4(p);
      b[j] = ldg_nc_f4(p + 32);
    }
  }
  // writes fp16 k-offsets [4s,4s+4) and [32+4s, 32+4s+4) of panel row r
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
    st_shared_v2(row + (((static_cast<uint32_t>(s >> 1)) ^ sw) << 4) + half, pack_half2(x0, x1), pack_half2(x2, x3));
    st_shared_v2(row + (((static_cast<uint32_t>(4 + (s >> 1))) ^ sw) << 4) + half, pack_half2(y0, y1),
                 pack_half2(y2, y3));
  }
};

// fp16 table: per k-chunk a ray needs 3 x 128 B; 8 lanes x one 16-byte load per row.
template <>
struct RowLoader<__half> {
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

template <bool kTrain, typename TableT>
__global__ void __launch_bounds__(NUM_THREADS, 1) field_main_kernel(const FieldParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const Geo g = p.geo;
  const int L = g.L;

  auto bar = [&](int i) -> uint32_t { return smem_base + SM_BARS + 8u * static_cast<uint32_t>(i); };

  // ---- one-time setup ---------------------------------------------------------------------
  if (tid == 0) {
    for (int i = 0; i < NSTAGE_A; ++i) { mbar_init(bar(BAR_A_FULL + i), PROD_WARPS); mbar_init(bar(BAR_A_EMPTY + i), 1); }
    for (int i = 0; i < NSTAGE_B; ++i) { mbar_init(bar(BAR_B_FULL + i), 1); mbar_init(bar(BAR_B_EMPTY + i), 1); }
    mbar_init(bar(BAR_MMA_DONE), 1);
    mbar_init(bar(BAR_ACT_READY), 1);
    mbar_fence_init();
  }
  if (warp == MMA_WARP) {
    tmem_alloc(smem_base + SM_TMEMPTR, TMEM_COLS);
    tmem_relinquish();
  }
  {
    // f32 section of the packed buffer -> smem; ones tile for the bias-gradient MMAs
    const float* f32sec = reinterpret_cast<const float*>(p.packed + packed_f32_offset(g));
    float* s_bias = reinterpret_cast<float*>(smem + SM_BIAS);
    const int nf = (L - 1) * HID + OUT_DIM * HID + 4;
    for (int i = tid; i < nf; i += NUM_THREADS) {
      const float v = f32sec[i];
      if (i < (L - 1) * HID) s_bias[i] = v;
      else reinterpret_cast<float*>(smem + SM_WLAST)[i - (L - 1) * HID] = v;  // wlast then blast (contiguous)
    }
    uint32_t* ones = reinterpret_cast<uint32_t*>(smem + SM_ONES);
    for (int i = tid; i < 512; i += NUM_THREADS) ones[i] = 0x3C003C00u;  // fp16 1.0 x2
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + SM_TMEMPTR);

  const int ntiles = p.ntiles;
  const int nchunks = g.nchunks;

  if (warp >= EPI_WARPS && warp < EPI_WARPS + PROD_WARPS) {
    // =========================== gather producers ===========================================
    const int pt = tid - EPI_THREADS;
    const int pw = pt >> 5;
    const int s = lane & 7, gq = lane >> 3;
    const TableT* table = reinterpret_cast<const TableT*>(p.table);
    uint32_t* s_vid = reinterpret_cast<uint32_t*>(smem + SM_VID);
    float* s_bary = reinterpret_cast<float*>(smem + SM_BARY);
    int stage = 0;
    uint32_t phase = 0;
    // tile metadata ring (3 deep): entry i = (ray i/3, vertex i%3) -> vertex id, barycentric weight
    auto load_meta = [&](int tile, int slot) {
      const int64_t ray0 = static_cast<int64_t>(tile) * TILE_M;
      for (int i = pt; i < 384; i += PROD_THREADS) {
        const int64_t e = ray0 * 3 + i;
        uint32_t v = 0;
        float w = 0.f;
        if (e < p.n_rays * 3) {
          if (p.dense) { v = static_cast<uint32_t>(ray0 + i / 3); w = (i % 3 == 0) ? 1.f : 0.f; }
          else {
            v = p.idx64 ? static_cast<uint32_t>(reinterpret_cast<const long long*>(p.vids)[e])
                        : static_cast<uint32_t>(reinterpret_cast<const int*>(p.vids)[e]);
            w = p.bary[e];
          }
        }
        s_vid[slot * 384 + i] = v;
        s_bary[slot * 384 + i] = w;
      }
    };
    // Software-pipelined gather.  A "pass" = 4 rays per warp (8 lanes per ray) of one k-chunk.  Two passes
    // are always in flight: the loads of pass n+2 are issued right after pass n has been consumed -- also
    // across k-chunk and tile boundaries, because a load needs registers only, not a free ring slot.
    struct Cursor { int tile, slot, c, q; };
    auto advance = [&](Cursor& k) {
      if (++k.q == 4) {
        k.q = 0;
        if (++k.c == nchunks) { k.c = 0; k.tile += gridDim.x; k.slot = (k.slot + 1) % 3; }
      }
    };
    // rows of a pass: lanes 8g..8g+7 take row base + {0,4,1,5}[g] so that 64-bit stores are bank-conflict free
    auto pass_row = [&](int q) { return 16 * pw + 8 * (q >> 1) + 2 * (q & 1) + (gq >> 1) + 4 * (gq & 1); };
    auto issue = [&](const Cursor& k, RowLoader<TableT>& ld) {
      if (k.tile >= ntiles) return;
      const int e = k.slot * 384 + 3 * pass_row(k.q);
      int64_t roff[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) roff[j] = static_cast<int64_t>(s_vid[e + j]) * p.row_stride;
      ld.load(table, roff, k.c * KC, s);
    };
    auto consume = [&](int slot, int q, const RowLoader<TableT>& ld, uint32_t panel) {
      const int r = pass_row(q);
      const float w[3] = {s_bary[slot * 384 + 3 * r], s_bary[slot * 384 + 3 * r + 1], s_bary[slot * 384 + 3 * r + 2]};
      ld.store(panel, r, s, w);
    };
    RowLoader<TableT> ld0, ld1;
    Cursor L{static_cast<int>(blockIdx.x), 0, 0, 0};
    if (static_cast<int>(blockIdx.x) < ntiles) {
      load_meta(blockIdx.x, 0);
      named_bar_sync(2, PROD_THREADS);
      issue(L, ld0); advance(L);
      issue(L, ld1); advance(L);
    }
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int slot = it % 3;
      const int next_tile = tile + gridDim.x;
      if (next_tile < ntiles) load_meta(next_tile, (it + 1) % 3);   // the load cursor enters it 2 passes early
      named_bar_sync(2, PROD_THREADS);
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(bar(BAR_A_EMPTY + stage), phase ^ 1u, p.err, 0x10);
        const uint32_t panel = smem_base + SM_A + static_cast<uint32_t>(stage) * PANEL_BYTES;
        consume(slot, 0, ld0, panel); issue(L, ld0); advance(L);
        consume(slot, 1, ld1, panel); issue(L, ld1); advance(L);
        consume(slot, 2, ld0, panel); issue(L, ld0); advance(L);
        consume(slot, 3, ld1, panel); issue(L, ld1); advance(L);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(BAR_A_FULL + stage));
        if (++stage == NSTAGE_A) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == LOAD_WARP) {
    // =========================== weight loader ================================================
    if (lane == 0) {
      const int nblobs = kTrain ? g.nblobs_train : g.nblobs_fwd;
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int b = 0; b < nblobs; ++b) {
          mbar_wait(bar(BAR_B_EMPTY + stage), phase ^ 1u, p.err, 0x20);
          mbar_arrive_expect_tx(bar(BAR_B_FULL + stage), BLOB_BYTES);
          bulk_g2s(smem_base + SM_B + static_cast<uint32_t>(stage) * BLOB_BYTES,
                   p.packed + static_cast<size_t>(b) * BLOB_BYTES, BLOB_BYTES, bar(BAR_B_FULL + stage));
          if (++stage == NSTAGE_B) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // =========================== MMA issuer ====================================================
    if (lane == 0) {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0, pact = 0;
      int it = 0;
      auto wait_act = [&]() { mbar_wait(bar(BAR_ACT_READY), pact, p.err, 0x31); pact ^= 1u; tc_fence_after(); };
      auto wait_b = [&]() -> uint32_t {
        mbar_wait(bar(BAR_B_FULL + sb), pb, p.err, 0x32);
        tc_fence_after();
        return smem_base + SM_B + static_cast<uint32_t>(sb) * BLOB_BYTES;
      };
      auto release_b = [&]() { umma_commit(bar(BAR_B_EMPTY + sb)); if (++sb == NSTAGE_B) { sb = 0; pb ^= 1u; } };
      // 128x128x128 product of two 2-panel K-major images
      auto mma_128 = [&](uint32_t dcol, uint32_t abase, uint32_t bbase, bool acc_first) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t o = static_cast<uint32_t>(kk >> 2) * PANEL_BYTES + static_cast<uint32_t>(kk & 3) * 32u;
          umma_f16(tmem_base + dcol, umma_desc_sw128(abase + o, 0, 1024), umma_desc_sw128(bbase + o, 0, 1024),
                   IDESC_N128, (acc_first || kk > 0) ? 1u : 0u);
        }
      };
      // bias gradient: D[n, 0..15] += sum_rays dz[ray, n] * 1
      auto mma_db = [&](int l, uint32_t abase, bool acc_first) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          umma_f16(tmem_base + TM_DB + 16u * static_cast<uint32_t>(l),
                   umma_desc_sw128(abase + static_cast<uint32_t>(kk) * 2048u, PANEL_BYTES, 1024),
                   umma_desc_sw128(smem_base + SM_ONES, 0, 1024), IDESC_DB, (acc_first || kk > 0) ? 1u : 0u);
        }
      };
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        // ---- GEMM1: [z0 | zy] = x * [W0;Wy]^T --------------------------------------------
        for (int c = 0; c < nchunks; ++c) {
          mbar_wait(bar(BAR_A_FULL + sa), pa, p.err, 0x30);
          const uint32_t bbase = wait_b();
          const uint32_t abase = smem_base + SM_A + static_cast<uint32_t>(sa) * PANEL_BYTES;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            umma_f16(tmem_base + TM_Z0, umma_desc_sw128(abase + kk * 32u, 0, 1024),
                     umma_desc_sw128(bbase + kk * 32u, 0, 1024), IDESC_N256, (c > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(bar(BAR_A_EMPTY + sa));
          if (++sa == NSTAGE_A) { sa = 0; pa ^= 1u; }
          release_b();
        }
        umma_commit(bar(BAR_MMA_DONE));
        // ---- hidden chain, layers 1..L-2 -----------------------------------------------------
        int wr = 0;  // index of the activation-buffer write being consumed (buffer = wr & 1)
        for (int l = 1; l <= L - 2; ++l, ++wr) {
          wait_act();
          const uint32_t bbase = wait_b();
          const uint32_t abase = smem_base + SM_H + static_cast<uint32_t>(wr & 1) * BLOB_BYTES;
          const bool is_skip = (l == g.skip);
          mma_128(is_skip ? TM_ZY : TM_Z0, abase, bbase, is_skip);
          release_b();
          umma_commit(bar(BAR_MMA_DONE));
        }
        wait_act();  // epilogue of layer L-2 done (inference: tile finished; training: dz_{L-2} ready)
        if (kTrain) {
          wr = L - 1;  // write L-2 was h_{L-2} (CUDA-core consumer), write L-1 is dz_{L-2}
          for (int l = L - 2; l >= 1; --l, ++wr) {
            const uint32_t abase = smem_base + SM_H + static_cast<uint32_t>(wr & 1) * BLOB_BYTES;
            mma_db(l, abase, it > 0);
            const uint32_t bbase = wait_b();
            mma_128(TM_Z0, abase, bbase, false);  // dh_{l-1} = dz_l * W_l
            release_b();
            umma_commit(bar(BAR_MMA_DONE));
            wait_act();  // dz_{l-1} written
          }
          mma_db(0, smem_base + SM_H + static_cast<uint32_t>(wr & 1) * BLOB_BYTES, it > 0);
        }
      }
      if (kTrain) umma_commit(bar(BAR_MMA_DONE));  // bias-gradient accumulators final
    }
  } else {
    // =========================== epilogue warps ================================================
    const int r = tid;  // tile row == TMEM lane
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const float* s_bias = reinterpret_cast<const float*>(smem + SM_BIAS);
    const float* s_wlast = reinterpret_cast<const float*>(smem + SM_WLAST);
    const float* s_blast = reinterpret_cast<const float*>(smem + SM_BLAST);
    float* s_dzl = reinterpret_cast<float*>(smem + SM_DZL);
    uint4* s_mask = reinterpret_cast<uint4*>(smem + SM_MASK);
    uint32_t pdone = 0;
    float loss_abs = 0.f, loss_sq = 0.f;
    float gw[3] = {0.f, 0.f, 0.f}, gb[3] = {0.f, 0.f, 0.f};

    auto wait_done = [&]() { mbar_wait(bar(BAR_MMA_DONE), pdone, p.err, 0x40); pdone ^= 1u; tc_fence_after(); };
    // end of an epilogue step: publish smem writes to the async proxy, stash, release the MMA thread
    auto finish_step = [&](uint32_t hbuf, uint8_t* stash_dst) {
      fence_proxy_async_smem();
      tc_fence_before();
      named_bar_sync(1, EPI_THREADS);
      if (tid == 0) {
        if (kTrain) {
          if (stash_dst != nullptr) bulk_s2g(stash_dst, hbuf, BLOB_BYTES);
          bulk_commit();
          bulk_wait_read<1>();
        }
        mbar_arrive(bar(BAR_ACT_READY));
      }
    };

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int64_t ray = static_cast<int64_t>(tile) * TILE_M + r;
      const bool valid = ray < p.n_rays;
      uint8_t* stash_tile = kTrain ? p.stash + static_cast<size_t>(tile) * g.nslots * BLOB_BYTES : nullptr;
      int wr = 0;
      // ---- forward layers 0..L-3: h_l = relu(z_l + b_l) -> activation panel ------------------
      for (int l = 0; l <= L - 3; ++l, ++wr) {
        wait_done();
        const uint32_t hbuf = smem_base + SM_H + static_cast<uint32_t>(wr & 1) * BLOB_BYTES;
        const uint32_t col0 = (l == g.skip) ? TM_ZY : TM_Z0;
        uint32_t mbits[4];
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
          uint32_t v[32];
          tmem_ld32(tlane + col0 + 32u * cb, v);
          tmem_ld_wait();
          float h[32];
          uint32_t m = 0;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float z = __uint_as_float(v[j]) + s_bias[l * HID + 32 * cb + j];
            m |= (z > 0.f ? 1u : 0u) << j;
            h[j] = fminf(fmaxf(z, 0.f), 60000.f);
          }
          mbits[cb] = m;
          store_act_chunk(hbuf, r, cb, h);
        }
        if (kTrain) s_mask[l * TILE_M + r] = make_uint4(mbits[0], mbits[1], mbits[2], mbits[3]);
        finish_step(hbuf, kTrain ? stash_tile + static_cast<size_t>(stash_slot_h(g, l)) * BLOB_BYTES : nullptr);
      }
      // ---- layer L-2 epilogue + output layer (128 -> 3, sigmoid) on CUDA cores -----------------
      {
        const int l = L - 2;
        wait_done();
        const uint32_t hbuf = smem_base + SM_H + static_cast<uint32_t>(wr & 1) * BLOB_BYTES;
        const uint32_t col0 = (l == g.skip) ? TM_ZY : TM_Z0;
        float y[3] = {s_blast[0], s_blast[1], s_blast[2]};
        uint32_t mbits[4];
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
          uint32_t v[32];
          tmem_ld32(tlane + col0 + 32u * cb, v);
          tmem_ld_wait();
          float h[32];
          uint32_t m = 0;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float z = __uint_as_float(v[j]) + s_bias[l * HID + 32 * cb + j];
            m |= (z > 0.f ? 1u : 0u) << j;
            h[j] = fmaxf(z, 0.f);
            y[0] = fmaf(s_wlast[0 * HID + 32 * cb + j], h[j], y[0]);
            y[1] = fmaf(s_wlast[1 * HID + 32 * cb + j], h[j], y[1]);
            y[2] = fmaf(s_wlast[2 * HID + 32 * cb + j], h[j], y[2]);
            h[j] = fminf(h[j], 60000.f);
          }
          mbits[cb] = m;
          if (kTrain) store_act_chunk(hbuf, r, cb, h);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) y[c] = sigmoidf_(y[c]);
        if (valid) {
          p.pred[ray * 3 + 0] = y[0];
          p.pred[ray * 3 + 1] = y[1];
          p.pred[ray * 3 + 2] = y[2];
        }
        if (!kTrain) {
          finish_step(hbuf, nullptr);
        } else {
          ++wr;
          float dzl[3] = {0.f, 0.f, 0.f};
          if (valid) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              float gy;
              if (p.target != nullptr) {
                const float d = y[c] - p.target[ray * 3 + c];
                loss_abs += fabsf(d);
                loss_sq += d * d;
                gy = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);  // d|d|/dd, sign(0) = 0 as F.l1_loss
              } else {
                gy = p.grad_pred[ray * 3 + c] * p.ext_scale;
              }
              dzl[c] = gy * y[c] * (1.f - y[c]) * DZ_SCALE;
            }
          }
          reinterpret_cast<float4*>(s_dzl)[r] = make_float4(dzl[0], dzl[1], dzl[2], 0.f);
          if (tid == 0) { bulk_commit(); bulk_wait_read<1>(); }  // keeps the 2-buffer store ring in step
          named_bar_sync(1, EPI_THREADS);
          // output-layer weight gradient: thread j reduces over the tile's rays
          {
            const int j = tid;
            const uint8_t* hp = smem + SM_H + (hbuf - (smem_base + SM_H)) + static_cast<uint32_t>(j >> 6) * PANEL_BYTES;
            const uint32_t cj = static_cast<uint32_t>(j & 63);
#pragma unroll 4
            for (int m = 0; m < TILE_M; ++m) {
              const float hv = __half2float(*reinterpret_cast<const __half*>(hp + sw128_off(m, cj)));
              const float4 d = reinterpret_cast<const float4*>(s_dzl)[m];
              gw[0] = fmaf(d.x, hv, gw[0]); gw[1] = fmaf(d.y, hv, gw[1]); gw[2] = fmaf(d.z, hv, gw[2]);
              gb[0] += d.x; gb[1] += d.y; gb[2] += d.z;
            }
          }
          // dz_{L-2} = (dz_last * W_last) masked by relu'
          const uint32_t hbuf2 = smem_base + SM_H + static_cast<uint32_t>(wr & 1) * BLOB_BYTES;
#pragma unroll
          for (int cb = 0; cb < 4; ++cb) {
            float h[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = 32 * cb + j;
              const float dh = dzl[0] * s_wlast[col] + dzl[1] * s_wlast[HID + col] + dzl[2] * s_wlast[2 * HID + col];
              h[j] = ((mbits[cb] >> j) & 1u) ? dh : 0.f;
            }
            store_act_chunk(hbuf2, r, cb, h);
          }
          finish_step(hbuf2, stash_tile + static_cast<size_t>(stash_slot_dz(g, l)) * BLOB_BYTES);
          ++wr;
          // ---- dgrad chain: dz_l = dh_l * relu'(z_l), l = L-3..0 --------------------------------
          for (int lb = L - 3; lb >= 0; --lb, ++wr) {
            wait_done();
            const uint32_t hb = smem_base + SM_H + static_cast<uint32_t>(wr & 1) * BLOB_BYTES;
            const uint4 mk = s_mask[lb * TILE_M + r];
            const uint32_t mm[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
            for (int cb = 0; cb < 4; ++cb) {
              uint32_t v[32];
              tmem_ld32(tlane + TM_Z0 + 32u * cb, v);
              tmem_ld_wait();
              float h[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) h[j] = ((mm[cb] >> j) & 1u) ? __uint_as_float(v[j]) : 0.f;
              store_act_chunk(hb, r, cb, h);
            }
            finish_step(hb, stash_tile + static_cast<size_t>(stash_slot_dz(g, lb)) * BLOB_BYTES);
          }
        }
      }
    }
    if (kTrain) {
      // ---- per-CTA partials: bias gradients (TMEM), output layer, loss sums ---------------------
      wait_done();
      float* part = p.part1 + static_cast<size_t>(blockIdx.x) * K1_PART_FLOATS;
      for (int l = 0; l <= L - 2; ++l) {
        uint32_t v;
        tmem_ld1(tlane + TM_DB + 16u * static_cast<uint32_t>(l), v);
        tmem_ld_wait();
        part[l * HID + tid] = __uint_as_float(v);
      }
      float* plast = part + (INF_MAX_LAYERS - 1) * HID;
      plast[0 * HID + tid] = gw[0];
      plast[1 * HID + tid] = gw[1];
      plast[2 * HID + tid] = gw[2];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        loss_abs += __shfl_xor_sync(0xffffffffu, loss_abs, o);
        loss_sq += __shfl_xor_sync(0xffffffffu, loss_sq, o);
      }
      float* s_red = reinterpret_cast<float*>(smem + SM_RED);
      if (lane == 0) { s_red[warp] = loss_abs; s_red[4 + warp] = loss_sq; }
      named_bar_sync(1, EPI_THREADS);
      if (tid == 0) {
        plast[3 * HID + 0] = gb[0];
        plast[3 * HID + 1] = gb[1];
        plast[3 * HID + 2] = gb[2];
        plast[3 * HID + 4] = s_red[0] + s_red[1] + s_red[2] + s_red[3];
        plast[3 * HID + 5] = s_red[4] + s_red[5] + s_red[6] + s_red[7];
        bulk_wait<0>();  // stash stores fully written before the grid completes
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---- host launcher (kernels stay private to this translation unit) -----------------------------
size_t field_main_smem_bytes() { return SM_ALLOC; }

template <bool kTrain, typename TableT>
static cudaError_t launch_one(const FieldParams& p, int grid, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(field_main_kernel<kTrain, TableT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(SM_ALLOC));
  if (e != cudaSuccess) return e;
  field_main_kernel<kTrain, TableT><<<grid, NUM_THREADS, SM_ALLOC, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_field_main(const FieldParams& p, bool train, bool table_f16, int grid, cudaStream_t stream) {
  if (train) return table_f16 ? launch_one<true, __half>(p, grid, stream) : launch_one<true, float>(p, grid, stream);
  return table_f16 ? launch_one<false, __half>(p, grid, stream) : launch_one<false, float>(p, grid, stream);
}

}  // namespace inf
