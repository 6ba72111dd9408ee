// Forward (+ loss + dgrad chain) kernel of the intrinsic-neural-field path, sm_100a.
//
// Persistent CTAs (one per SM), 128-ray tiles, six warp roles connected by mbarrier rings:
//   producers (16 warps)   gather 3 table rows per ray with 16-byte loads, interpolate in fp32 with the
//                          barycentric weights, round to fp16 and store the [128 x 64] A-panel of the
//                          current k-chunk straight into the SWIZZLE_128B smem layout the tensor core
//                          reads -- the interpolated embedding (mesh.py:313-324) is never read back from HBM
//                          by the forward pass;
//   B1 loader (1 thread)   streams the [W0;Wy] operand chunks as 16 KB half-stages (W0 rows, Wy rows) with 1-D bulk
//                          copies (TMA), each fetched once per 2-CTA cluster and multicast;
//   GEMM1 issuer (1 thr)   tcgen05.mma kind::f16: [z0|zy] = x * [W0;Wy]^T (N=256, K=k_pad) into one of TWO
//                          256-column TMEM buffers, so the gather/GEMM1 of tile t+1 overlaps the layer chain of
//                          tile t; when training it also bulk-stores each x panel to the stash (operand of
//                          the first-layer weight gradient in field_wgrad.cu);
//   B2 loader (1 thread)   streams the 128x128 hidden / view-branch weight images;
//   chain issuer (1 thr)   the 128-wide hidden chain (model.py:98-112), the view-direction layer
//                          (model.py:168-173) and, when training, the dgrad chain -- the dgrad GEMMs read the
//                          SAME weight images MN-major, so no transposed weights exist;
//   epilogue (4 warps)     TMEM -> registers, bias + ReLU -> fp16 activation panel in smem (A operand of
//                          the next layer), Fourier features of the view direction (layers.py:21-25), final
//                          -> 3 layer + sigmoid (+ L1 loss gradient) on CUDA cores; when training, stashes
//                          h_l / dz_l tiles for the wgrad kernel.
#include <cstdlib>
#include "common.cuh"
#include "ptx.cuh"

namespace inf {

namespace {

// Six warpgroups.  Registers are re-partitioned at kernel start (setmaxnreg): the epilogue warpgroup needs
// ~150 (32-column TMEM slices in flight), a producer warp ~80 (two gather passes in flight), the four
// single-thread roles almost none.
constexpr int NUM_THREADS_MAIN = 24 * 32;
constexpr int EPI_WARPS = 4;      // warps 0..3: TMEM epilogue (warp w owns TMEM lanes 32w..32w+31)
constexpr int PROD_WARPS = 16;    // warps 4..19: gather producers (8 tile rows each)
constexpr int W_G1 = 20;          // GEMM1 issuer + TMEM allocator
constexpr int W_CH = 21;          // chain issuer
constexpr int W_B1 = 22;          // [W0;Wy] chunk loader
constexpr int W_B2 = 23;          // hidden-weight loader
constexpr int REGS_EPI = 128, REGS_UTIL = 32;   // pool = 768 x 80 (launch value): 128*128 + 512*80 + 128*32
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int PROD_THREADS = PROD_WARPS * 32;

constexpr int MAX_STAGE_B1 = 4;   // half-stages of the [W0;Wy] ring
constexpr int MAX_STAGE_A = 5;

// shared memory carve-up (bytes from a 1024-aligned base), at most 9 panels of 16 KB + small tables = 163 KB:
//   training   A ring 3 x 16 KB (2 with the view branch: one stage goes to the Fourier panel), B1 one whole chunk
//              (2 half-stages), H 32 KB, B2 32 KB
//   inference  A ring 2, B1 3 half-stages (2 with the view branch), H 32 KB, B2 32 KB
struct Smem {
  uint32_t a, b1, h, b2, e, dzlt, bias, wlast, vid, bary, red, bars, tmemptr, total;
  int nstage_a, nstage_b1, nh;
};
__host__ __device__ constexpr Smem smem_layout(bool view, bool train) {
  Smem s{};
  s.nh = 1;
  // Inference stays within the 164 KB shared-memory carve-out: the gather's register loads are bounded by what is left
  // of the 256 KB array for L1 (measured with scripts/probes/gather_probe.cu: 7.3 TB/s of random 4 KB rows at <= 164 KB of
  // shared memory, 5.1-5.6 TB/s at 196/228 KB), and in inference the gather is the whole kernel.
  // Training takes the same diet: one activation buffer (every stash store's shared-memory read is waited for before
  // the chain continues) and ONE whole [W0;Wy] stage (its refill, ~1 us, hides behind the >= 2 us the producers need
  // per k-chunk), which keeps the single N = 256 instruction per k-step.
  s.nstage_a = train ? (view ? 2 : 3) : 2;
  s.nstage_b1 = train ? 2 : (view ? 2 : 3);       // 16 KB half-stages: [W0 chunk] and [Wy chunk] are released separately
  s.a = 0;
  s.b1 = s.a + static_cast<uint32_t>(s.nstage_a) * PANEL_BYTES;
  s.h = s.b1 + static_cast<uint32_t>(s.nstage_b1) * PANEL_BYTES;
  s.b2 = s.h + static_cast<uint32_t>(s.nh) * BLOB_BYTES;
  s.e = s.b2 + BLOB_BYTES;
  s.dzlt = s.e + (view ? PANEL_BYTES : 0u);
  s.bias = s.dzlt + (train ? DZLT_BYTES : 0u);    // f32 [MAX_MMA_LAYERS][128]
  s.wlast = s.bias + MAX_MMA_LAYERS * HID * 4;    // f32 [3][128] + blast[4]
  s.vid = s.wlast + OUT_DIM * HID * 4 + 16;       // u32 [3][384]
  s.bary = s.vid + 3 * 384 * 4;                   // f32 [3][384]
  s.red = s.bary + 3 * 384 * 4;                   // f32 [20]
  s.bars = s.red + 96;
  s.tmemptr = s.bars + 32 * 8;
  s.total = s.tmemptr + 16;
  return s;
}
static_assert(smem_layout(false, false).total <= 166912 && smem_layout(true, false).total <= 166912 &&
              smem_layout(false, true).total <= 166912 && smem_layout(true, true).total <= 166912,
              "every variant must fit the 164 KB carve-out (1 KB of it is reserved per CTA)");

// barrier indices
constexpr int BAR_A_FULL = 0;                          // [5]
constexpr int BAR_A_EMPTY = BAR_A_FULL + MAX_STAGE_A;  // [5]
constexpr int BAR_B1_FULL = BAR_A_EMPTY + MAX_STAGE_A; // [4]
constexpr int BAR_B1_EMPTY = BAR_B1_FULL + MAX_STAGE_B1;  // [4]
constexpr int BAR_B2_FULL = BAR_B1_EMPTY + MAX_STAGE_B1;
constexpr int BAR_B2_EMPTY = BAR_B2_FULL + 1;
constexpr int BAR_G_FULL = BAR_B2_EMPTY + 1;           // [2] GEMM1 accumulator buffer complete
constexpr int BAR_G_EMPTY = BAR_G_FULL + 2;            // [2] chain of the tile that used the buffer is done
constexpr int BAR_ACT_READY = BAR_G_EMPTY + 2;
constexpr int BAR_MMA_DONE = BAR_ACT_READY + 1;
static_assert(BAR_MMA_DONE < 32, "barrier table overflow");

constexpr uint32_t TMEM_COLS = 512;   // two GEMM1 buffers: [256 buf, +128) z0 / chain accumulators, [+128, +256) zy

constexpr uint32_t IDESC_N256 = umma_idesc_f16(128, 256, 0, 0);
constexpr uint32_t IDESC_N128 = umma_idesc_f16(128, 128, 0, 0);
constexpr uint32_t IDESC_N128_BMN = umma_idesc_f16(128, 128, 0, 1);  // dgrad: B = weight image read MN-major

constexpr uint32_t DESC_HI = umma_desc_hi(1024);   // every operand image here: SWIZZLE_128B, 1024 B between 8-row groups

constexpr float PI_F = 3.14159274101257324f;  // float32(torch.pi), layers.py:14

__device__ __forceinline__ float sigmoidf_(float v) { return __fdividef(1.f, 1.f + __expf(-v)); }

// 16 packed f16x2 words (columns 32*cb .. 32*cb+31 of row r) into the 2-panel activation image.
__device__ __forceinline__ void store_act_words(uint32_t hbuf, int r, int cb, const uint32_t (&w)[16]) {
  const uint32_t panel = hbuf + static_cast<uint32_t>(cb >> 1) * PANEL_BYTES + static_cast<uint32_t>(r) * 128u;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t chunk = static_cast<uint32_t>((cb & 1) * 4 + i);
    const uint32_t addr = panel + ((chunk ^ (static_cast<uint32_t>(r) & 7u)) << 4);
    st_shared_v4(addr, w[4 * i + 0], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
  }
}
// ReLU mask of 16 packed non-negative f16x2 words: bit i = (low half of word i != 0), bit 16+i = (high half != 0).
// (h + 0x7FFF sets bit 15 exactly when the non-negative half h is non-zero; halves cannot carry into each other.)
__device__ __forceinline__ uint32_t relu_mask16(const uint32_t (&w)[16]) {
  uint32_t m = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) m |= ((w[i] + 0x7FFF7FFFu) >> (15 - i)) & (0x00010001u << i);
  return m;
}
// keep / zero the halves of word i according to the mask layout of relu_mask16
__device__ __forceinline__ uint32_t apply_mask(uint32_t w, uint32_t m, int i) {
  return w & (((m >> i) & 0x00010001u) * 0xFFFFu);
}

// A producer warp owns 8 rows of every A panel: two "quads" of 4 rows; lanes 8g..8g+7 serve row
// {0,4,1,5}[g] (+2 for the second quad) of the warp's 8-row swizzle atom, which makes the panel stores
// bank-conflict free.  A pass is what one RowLoader holds in registers between issue and consume.
template <typename TableT>
struct RowLoader;

// fp32 table: a pass = one quad; per k-chunk a ray needs 3 x 256 B = 8 lanes x 2 float4 per row.
template <>
struct RowLoader<float> {
  static constexpr int QUADS = 1;        // quads per pass
  float4 a[3], b[3];
  __device__ __forceinline__ void load(int, const float* base, const int64_t (&roff)[3], int koff, int s) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float* p = base + roff[j] + koff + 4 * s;
      a[j] = ldg_nc_f4(p);
      b[j] = ldg_nc_f4(p + 32);
    }
  }
  // writes fp16 k-offsets [4s,4s+4) and [32+4s, 32+4s+4) of panel row r
  __device__ __forceinline__ void store(int, uint32_t panel, int r, int s, const float (&w)[3]) const {
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

// fp16 table: a pass = both quads (same bytes in flight as the fp32 pass); per k-chunk a ray needs
// 3 x 128 B = 8 lanes x one 16-byte load per row.
template <>
struct RowLoader<__half> {
  static constexpr int QUADS = 2;
  uint4 a[2][3];
  __device__ __forceinline__ void load(int u, const __half* base, const int64_t (&roff)[3], int koff, int s) {
#pragma unroll
    for (int j = 0; j < 3; ++j) a[u][j] = ldg_nc_u4(base + roff[j] + koff + 8 * s);
  }
  __device__ __forceinline__ void store(int u, uint32_t panel, int r, int s, const float (&w)[3]) const {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const uint32_t v[4] = {a[u][j].x, a[u][j].y, a[u][j].z, a[u][j].w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&v[i]));
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

template <int N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// Element `idx` of FourierFeatEnc(x) (layers.py:21-25, use_logspace=True -> freq_j = 2^j * pi):
// cat(cos(x*f) [input-major, freq-minor], sin(x*f), x); zero beyond the embedding.
__device__ __forceinline__ float fourier_feature(int idx, float x0, float x1, float x2, int d, int emb, int include) {
  const int nf = d * emb;
  if (idx < 2 * nf) {
    const int q = idx < nf ? idx : idx - nf;
    const int dim = q / emb, j = q - dim * emb;
    const float x = dim == 0 ? x0 : (dim == 1 ? x1 : x2);
    const float arg = x * ldexpf(PI_F, j);
    return idx < nf ? cosf(arg) : sinf(arg);
  }
  if (include && idx < 2 * nf + d) {
    const int dim = idx - 2 * nf;
    return dim == 0 ? x0 : (dim == 1 ? x1 : x2);
  }
  return 0.f;
}

}  // namespace

template <bool kTrain, typename TableT>
__global__ void __launch_bounds__(NUM_THREADS_MAIN, 1) field_main_kernel(const FieldParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const Geo g = p.geo;
  const bool view = g.view != INF_VIEW_NONE;
  const Smem sm = smem_layout(view, kTrain);
  const int nstage_a = sm.nstage_a, nstage_b1 = sm.nstage_b1;
  const int nb = g.nb, nhid = g.nhid;
  // Cluster of `cs` CTAs (1 or 2): every [W0;Wy] chunk is fetched from L2 ONCE per cluster -- each CTA loads 1/cs of it
  // and multicasts that slice to all of them.  The CTAs of a cluster therefore run the same number of tile iterations;
  // a CTA whose tile index is past the end processes a null tile (all rays invalid, nothing stored).
  const int cs = p.cluster;
  const uint32_t crank = cs > 1 ? cluster_ctarank() : 0u;
  const uint16_t cmask = static_cast<uint16_t>((1u << cs) - 1u);

  auto bar = [&](int i) -> uint32_t { return smem_base + sm.bars + 8u * static_cast<uint32_t>(i); };
  // measurement only (INF_DEBUG & 8): per-role cycle accounting of CTA 0, written once at the end of the kernel
  const bool tracing = p.trace != nullptr && blockIdx.x == 0;
  auto now = [&]() -> long long { return tracing ? clock64() : 0ll; };
  auto put = [&](int slot, long long v) { if (tracing) p.trace[slot] = static_cast<unsigned long long>(v); };

  if ((smem_base & 1023u) != 0) {  // SWIZZLE_128B operand images need a 1024-byte aligned base
    if (tid == 0 && p.err) atomicExch(p.err, 0xBAD00001u);
    return;
  }

  // ---- one-time setup ---------------------------------------------------------------------
  if (tid == 0) {
    for (int i = 0; i < MAX_STAGE_A; ++i) {
      mbar_init(bar(BAR_A_FULL + i), PROD_WARPS);
      mbar_init(bar(BAR_A_EMPTY + i), kTrain ? 2 : 1);  // MMA commit (+ x-panel stash store finished reading)
    }
    for (int i = 0; i < MAX_STAGE_B1; ++i) { mbar_init(bar(BAR_B1_FULL + i), 1); mbar_init(bar(BAR_B1_EMPTY + i), cs); }
    mbar_init(bar(BAR_B2_FULL), 1);
    mbar_init(bar(BAR_B2_EMPTY), 1);
    for (int i = 0; i < 2; ++i) { mbar_init(bar(BAR_G_FULL + i), 1); mbar_init(bar(BAR_G_EMPTY + i), EPI_WARPS); }
    mbar_init(bar(BAR_ACT_READY), 1);
    mbar_init(bar(BAR_MMA_DONE), 1);
    mbar_fence_init();
  }
  if (warp == W_G1) {
    tmem_alloc(smem_base + sm.tmemptr, TMEM_COLS);
    tmem_relinquish();
  }
  if (p.debug & DEBUG_PDL_EARLY) pdl_launch_dependents();
  pdl_wait();   // the operand images (previous step's Adam refresh), the stash's previous readers
  {
    // f32 section of the packed buffer -> smem (bias rows, then wlast + blast which are contiguous in both)
    const float* f32sec = reinterpret_cast<const float*>(p.packed + packed_f32_offset(g));
    float* s_bias = reinterpret_cast<float*>(smem + sm.bias);
    float* s_wl = reinterpret_cast<float*>(smem + sm.wlast);
    const int nbias = (nb + 1) * HID;
    const int nf = packed_f32_count(g);
    for (int i = tid; i < nf; i += NUM_THREADS_MAIN) {
      const float v = f32sec[i];
      if (i < nbias) s_bias[i] = v;
      else s_wl[i - nbias] = v;
    }
    if (kTrain) {
      uint32_t* dz = reinterpret_cast<uint32_t*>(smem + sm.dzlt);   // rows 3..15 of dzl^T stay zero
      for (int i = tid; i < DZLT_BYTES / 4; i += NUM_THREADS_MAIN) dz[i] = 0u;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (cs > 1) cluster_sync_all();   // every CTA's barriers are initialised before a peer multicasts into / arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + sm.tmemptr);

  const int ntiles = p.ntiles;
  // loop bound of this CTA: as many iterations as the first CTA of its cluster needs
  int nt_loop = ntiles;
  if (cs > 1) {
    const int first = (static_cast<int>(blockIdx.x) / cs) * cs;
    const int iters = first < ntiles ? (ntiles - first + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x) : 0;
    nt_loop = iters > 0 ? static_cast<int>(blockIdx.x) + (iters - 1) * static_cast<int>(gridDim.x) + 1 : 0;
  }
  const int nchunks = g.nchunks;
  const size_t tile_stride = p.stash_stride;
  // stash tiles are written once, read once by the next kernel (INF_DEBUG & 64: no hint, measurement)
  const uint64_t stash_policy = (p.debug & 64) ? 0ull : l2_policy_evict_first();

  // Register re-partition: ONE setmaxnreg instance per warpgroup (all four warps of a warpgroup must execute the
  // same instruction), placed so that it dominates the role's code -- ptxas budgets each region from it.
  if (warp >= EPI_WARPS && warp < EPI_WARPS + PROD_WARPS) {
    // =========================== gather producers ===========================================
    const int pt = tid - EPI_THREADS;
    const int pw = pt >> 5;
    const int s = lane & 7, gq = lane >> 3;
    const TableT* table = reinterpret_cast<const TableT*>(p.table);
    uint32_t* s_vid = reinterpret_cast<uint32_t*>(smem + sm.vid);
    float* s_bary = reinterpret_cast<float*>(smem + sm.bary);
    int stage = 0;
    uint32_t phase = 0;
    // tile metadata ring (3 deep): entry i = (ray i/3, vertex i%3) -> vertex id, barycentric weight
    auto load_meta = [&](int tile, int slot) {
      const int64_t ray0 = static_cast<int64_t>(tile) * TILE_M;
      for (int i = pt; i < 384; i += PROD_THREADS) {
        const int64_t ray = ray0 + i / 3;
        uint32_t v = 0;
        float w = 0.f;
        if (ray < p.n_rays) {
          if (p.dense) { v = static_cast<uint32_t>(ray); w = (i % 3 == 0) ? 1.f : 0.f; }
          else {
            // batch selection folded in (ray_dataloader.py:122-126): row of the split's arrays this ray reads
            int64_t src = ray;
            if (p.ridx != nullptr)
              src = p.ridx64 ? static_cast<int64_t>(reinterpret_cast<const long long*>(p.ridx)[ray])
                             : static_cast<int64_t>(reinterpret_cast<const int*>(p.ridx)[ray]);
            const int64_t e = src * 3 + i % 3;
            v = p.idx64 ? static_cast<uint32_t>(reinterpret_cast<const long long*>(p.vids)[e])
                        : static_cast<uint32_t>(reinterpret_cast<const int*>(p.vids)[e]);
            w = p.bary[e];
          }
        }
        s_vid[slot * 384 + i] = v;
        s_bary[slot * 384 + i] = w;
      }
    };
    // Software-pipelined gather: two passes are always in flight per warp.  The loads of pass n+2 are issued
    // right after pass n has been consumed -- also across k-chunk and tile boundaries, because a load needs
    // registers only, not a free ring slot.
    constexpr int QUADS = RowLoader<TableT>::QUADS;
    constexpr int PPC = 2 / QUADS;   // passes per k-chunk
    struct Cursor { int tile, slot, c, q; };
    auto advance = [&](Cursor& k) {
      if (++k.q == PPC) {
        k.q = 0;
        if (++k.c == nchunks) { k.c = 0; k.tile += gridDim.x; k.slot = (k.slot + 1) % 3; }
      }
    };
    auto quad_row = [&](int j) { return 8 * pw + 2 * j + (gq >> 1) + 4 * (gq & 1); };
    // DRAM -> L2 prefetch, `pf` k-chunks at a time and `pf` chunks ahead of the loads: every table row is pulled
    // from HBM in (pf * 64 * sizeof(TableT))-byte pieces (one DRAM page visit instead of pf separate 128/256-byte
    // ones), through the bulk-copy engine, which does not occupy the LSU's outstanding-load slots.
    const int pf = p.prefetch;
    auto issue = [&](const Cursor& k, RowLoader<TableT>& ld) {
      if (k.tile >= nt_loop) return;
#pragma unroll
      for (int u = 0; u < QUADS; ++u) {
        const int e = k.slot * 384 + 3 * quad_row(k.q * QUADS + u);
        int64_t roff[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) roff[j] = static_cast<int64_t>(s_vid[e + j]) * p.row_stride;
        ld.load(u, table, roff, k.c * KC, s);
        if (pf > 0 && s == 0 && (k.c % pf) == 0) {
          int c2 = k.c + pf;
          if (c2 < nchunks) {
            const uint32_t bytes = static_cast<uint32_t>(min(pf, nchunks - c2) * KC * sizeof(TableT));
#pragma unroll
            for (int j = 0; j < 3; ++j) bulk_prefetch_l2(table + roff[j] + c2 * KC, bytes);
          } else if (k.tile + static_cast<int>(gridDim.x) < ntiles) {
            // first piece of the next tile's rows (its metadata slot was filled when this tile started)
            const int e2 = ((k.slot + 1) % 3) * 384 + 3 * quad_row(k.q * QUADS + u);
            const uint32_t bytes = static_cast<uint32_t>(min(pf, nchunks) * KC * sizeof(TableT));
#pragma unroll
            for (int j = 0; j < 3; ++j)
              bulk_prefetch_l2(table + static_cast<int64_t>(s_vid[e2 + j]) * p.row_stride, bytes);
          }
        }
      }
    };
    auto consume = [&](int slot, int q, const RowLoader<TableT>& ld, uint32_t panel) {
#pragma unroll
      for (int u = 0; u < QUADS; ++u) {
        const int r = quad_row(q * QUADS + u);
        const float w[3] = {s_bary[slot * 384 + 3 * r], s_bary[slot * 384 + 3 * r + 1], s_bary[slot * 384 + 3 * r + 2]};
        ld.store(u, panel, r, s, w);
      }
    };
    RowLoader<TableT> ld0, ld1;
    Cursor L{static_cast<int>(blockIdx.x), 0, 0, 0};
    // The load cursor runs two passes ahead of the consumer: one tile ahead at most when a tile has >= 2 passes per
    // warp, two tiles ahead when it has a single pass (fp16 table with k_pad = 64).  Metadata is loaded `ahead` tiles
    // ahead accordingly (three ring slots: consuming, +1, +2).
    const int ahead = (nchunks * PPC >= 2) ? 1 : 2;
    if (static_cast<int>(blockIdx.x) < nt_loop) {
      load_meta(blockIdx.x, 0);
      if (ahead == 2 && static_cast<int>(blockIdx.x + gridDim.x) < nt_loop) load_meta(blockIdx.x + gridDim.x, 1);
      named_bar_sync(2, PROD_THREADS);
      issue(L, ld0); advance(L);
      issue(L, ld1); advance(L);
    }
    int it = 0;
    uint32_t n = 0;   // running pass index: pass n lives in ld0 (even) / ld1 (odd)
    long long t_prod_wait = 0;
    const long long t_prod0 = now();
    for (int tile = blockIdx.x; tile < nt_loop; tile += gridDim.x, ++it) {
      const int slot = it % 3;
      const int next_tile = tile + ahead * static_cast<int>(gridDim.x);
      if (ahead == 2) named_bar_sync(2, PROD_THREADS);   // slot (it+2)%3 was the previous tile's: everyone has left it
      if (next_tile < nt_loop) load_meta(next_tile, (it + ahead) % 3);   // the load cursor enters it 2 passes early
      named_bar_sync(2, PROD_THREADS);
      for (int c = 0; c < nchunks; ++c) {
        const long long tw0 = now();
        mbar_wait(bar(BAR_A_EMPTY + stage), phase ^ 1u, p.err, 0x10);
        t_prod_wait += now() - tw0;
        const uint32_t panel = smem_base + sm.a + static_cast<uint32_t>(stage) * PANEL_BYTES;
#pragma unroll
        for (int q = 0; q < PPC; ++q, ++n) {
          if (n & 1u) { consume(slot, q, ld1, panel); issue(L, ld1); }
          else { consume(slot, q, ld0, panel); issue(L, ld0); }
          advance(L);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(BAR_A_FULL + stage));
        if (++stage == nstage_a) { stage = 0; phase ^= 1u; }
      }
    }
    if (pt == 0) { put(0, now() - t_prod0); put(1, t_prod_wait); }
  } else if (warp >= W_G1) {
   setmaxnreg_dec<REGS_UTIL>();
   if (warp == W_B1) {
    // =========================== [W0;Wy] chunk loader =========================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t slice = PANEL_BYTES / static_cast<uint32_t>(cs);
      const int nhalf = 2 * nchunks;   // a chunk blob = [W0 rows | Wy rows], one 16 KB half-stage each
      for (int tile = blockIdx.x; tile < nt_loop; tile += gridDim.x) {
        for (int c = 0; c < nhalf; ++c) {
          mbar_wait(bar(BAR_B1_EMPTY + stage), phase ^ 1u, p.err, 0x20);   // released by every CTA of the cluster
          mbar_arrive_expect_tx(bar(BAR_B1_FULL + stage), PANEL_BYTES);
          const uint32_t dst = smem_base + sm.b1 + static_cast<uint32_t>(stage) * PANEL_BYTES;
          const uint8_t* src = p.packed + static_cast<size_t>(c) * PANEL_BYTES;
          if (cs == 1) bulk_g2s(dst, src, PANEL_BYTES, bar(BAR_B1_FULL + stage));
          else bulk_g2s_multicast(dst + crank * slice, src + crank * slice, slice, bar(BAR_B1_FULL + stage), cmask);
          if (++stage == nstage_b1) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == W_B2) {
    // =========================== hidden-weight loader ========================================
    if (lane == 0) {
      uint32_t phase = 0;
      auto load = [&](int blob, uint32_t bytes) {
        mbar_wait(bar(BAR_B2_EMPTY), phase ^ 1u, p.err, 0x21);
        mbar_arrive_expect_tx(bar(BAR_B2_FULL), bytes);
        bulk_g2s(smem_base + sm.b2, p.packed + static_cast<size_t>(blob) * BLOB_BYTES, bytes, bar(BAR_B2_FULL));
        phase ^= 1u;
      };
      for (int tile = blockIdx.x; tile < nt_loop; tile += gridDim.x) {
        for (int l = 1; l <= nhid; ++l) load(blob_hidden(g, l), BLOB_BYTES);
        if (view) { load(blob_wd1z(g), BLOB_BYTES); load(blob_wd1e(g), PANEL_BYTES); }
        if (kTrain) {
          if (p.debug & 4) continue;
          if (view) load(blob_wd1z(g), BLOB_BYTES);
          for (int l = nhid; l >= 1; --l) load(blob_hidden(g, l), BLOB_BYTES);
        }
      }
    }
  } else if (warp == W_G1) {
    // =========================== GEMM1 issuer ================================================
    if (lane == 0) {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      uint32_t gph[2] = {0u, 0u};
      int it = 0;
      long long t_g = 0, t_a = 0, t_b = 0, t_st = 0;
      const long long t_g1_0 = now();
      for (int tile = blockIdx.x; tile < nt_loop; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const long long tg0 = now();
        mbar_wait(bar(BAR_G_EMPTY + buf), gph[buf] ^ 1u, p.err, 0x33);  // chain of tile it-2 left this buffer
        t_g += now() - tg0;
        gph[buf] ^= 1u;
        tc_fence_after();
        const uint32_t dcol = tmem_base + 256u * static_cast<uint32_t>(buf);
        uint8_t* xdst = (kTrain && tile < ntiles) ? p.stash + static_cast<size_t>(tile) * tile_stride + stash_off_x(g, 0) : nullptr;
        for (int c = 0; c < nchunks; ++c) {
          const long long ta0 = now();
          mbar_wait(bar(BAR_A_FULL + sa), pa, p.err, 0x30);
          t_a += now() - ta0;
          const uint32_t abase = smem_base + sm.a + static_cast<uint32_t>(sa) * PANEL_BYTES;
          const uint32_t alo = umma_desc_lo(abase, 0);
          if (kTrain) {
            // four half-stages = two whole [W0;Wy] chunks, each contiguous: one N = 256 instruction per k-step
            const long long tb0 = now();
            mbar_wait(bar(BAR_B1_FULL + sb), pb, p.err, 0x32);
            mbar_wait(bar(BAR_B1_FULL + sb + 1), pb, p.err, 0x35);
            t_b += now() - tb0;
            tc_fence_after();
            const uint32_t blo = umma_desc_lo(smem_base + sm.b1 + static_cast<uint32_t>(sb) * PANEL_BYTES, 0);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              umma_f16(dcol, umma_desc_join(alo + 2u * kk, DESC_HI), umma_desc_join(blo + 2u * kk, DESC_HI), IDESC_N256,
                       (c > 0 || kk > 0) ? 1u : 0u);
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              if (cs == 1) umma_commit(bar(BAR_B1_EMPTY + sb + half));
              else umma_commit_multicast(bar(BAR_B1_EMPTY + sb + half), cmask);
            }
            sb += 2;
            if (sb == nstage_b1) { sb = 0; pb ^= 1u; }
          } else {
#pragma unroll
          for (int half = 0; half < 2; ++half) {   // z0 += x W0^T into columns [0,128), zy += x Wy^T into [128,256)
            const long long tb0 = now();
            mbar_wait(bar(BAR_B1_FULL + sb), pb, p.err, 0x32);
            t_b += now() - tb0;
            tc_fence_after();
            const uint32_t blo = umma_desc_lo(smem_base + sm.b1 + static_cast<uint32_t>(sb) * PANEL_BYTES, 0);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              umma_f16(dcol + 128u * static_cast<uint32_t>(half), umma_desc_join(alo + 2u * kk, DESC_HI),
                       umma_desc_join(blo + 2u * kk, DESC_HI), IDESC_N128, (c > 0 || kk > 0) ? 1u : 0u);
            }
            if (cs == 1) umma_commit(bar(BAR_B1_EMPTY + sb));
            else umma_commit_multicast(bar(BAR_B1_EMPTY + sb), cmask);   // the stage is refilled by every CTA of the cluster
            if (++sb == nstage_b1) { sb = 0; pb ^= 1u; }
          }
          }
          umma_commit(bar(BAR_A_EMPTY + sa));
          if (kTrain) {
            // x panel -> stash (the B operand of dW0 / dWy in the wgrad kernel)
            if (!(p.debug & 1) && xdst != nullptr) bulk_s2g_hint(xdst + static_cast<size_t>(c) * PANEL_BYTES, abase, PANEL_BYTES, stash_policy);
            bulk_commit();
            const long long ts0 = now();
            bulk_wait_read<0>();
            t_st += now() - ts0;
            mbar_arrive(bar(BAR_A_EMPTY + sa));
          }
          if (++sa == nstage_a) { sa = 0; pa ^= 1u; }
        }
        umma_commit(bar(BAR_G_FULL + buf));
      }
      put(4, now() - t_g1_0); put(5, t_g); put(6, t_a); put(7, t_b); put(8, t_st);
      if (kTrain) bulk_wait<0>();
    }
  } else if (warp == W_CH) {
    // =========================== chain issuer ================================================
    if (lane == 0) {
      uint32_t pact = 0, pb2 = 0;
      uint32_t wr = 0;  // running index of the activation-buffer write being consumed (buffer = wr & 1)
      int it = 0;
      long long t_act = 0, t_b2 = 0;
      const long long t_ch0 = now();
      auto wait_act = [&]() {
        const long long t0 = now();
        mbar_wait(bar(BAR_ACT_READY), pact, p.err, 0x31); pact ^= 1u; tc_fence_after();
        t_act += now() - t0;
      };
      auto wait_b2 = [&]() {
        const long long t0 = now();
        mbar_wait(bar(BAR_B2_FULL), pb2, p.err, 0x34); pb2 ^= 1u; tc_fence_after();
        t_b2 += now() - t0;
      };
      const uint32_t bbase = smem_base + sm.b2;
      auto hbuf = [&](uint32_t) { return smem_base + sm.h; };
      const uint32_t blo_k = umma_desc_lo(bbase, 0);             // weight image read K-major (forward)
      const uint32_t blo_mn = umma_desc_lo(bbase, PANEL_BYTES);   // the same image read MN-major (dgrad)
      // D[128 x 128] (+)= A[128 x 64*npanel] * B^T, both K-major panel images
      auto mma_fwd = [&](uint32_t dcol, uint32_t abase, int npanel, bool acc_first) {
        const uint32_t alo = umma_desc_lo(abase, 0);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          if (kk < 4 * npanel) {
            const uint32_t o = static_cast<uint32_t>(kk >> 2) * (PANEL_BYTES >> 4) + static_cast<uint32_t>(kk & 3) * 2u;
            umma_f16(dcol, umma_desc_join(alo + o, DESC_HI), umma_desc_join(blo_k + o, DESC_HI), IDESC_N128,
                     (acc_first || kk > 0) ? 1u : 0u);
          }
        }
      };
      // dh[128 x 128] = dz[128 x 128] * W: A K-major, B = the forward weight image read MN-major (K = row n)
      auto mma_dgrad = [&](uint32_t dcol, uint32_t abase) {
        const uint32_t alo = umma_desc_lo(abase, 0);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t oa = static_cast<uint32_t>(kk >> 2) * (PANEL_BYTES >> 4) + static_cast<uint32_t>(kk & 3) * 2u;
          umma_f16(dcol, umma_desc_join(alo + oa, DESC_HI), umma_desc_join(blo_mn + static_cast<uint32_t>(kk) * 128u, DESC_HI),
                   IDESC_N128_BMN, kk > 0 ? 1u : 0u);
        }
      };
      for (int tile = blockIdx.x; tile < nt_loop; tile += gridDim.x, ++it) {
        const uint32_t base = tmem_base + 256u * static_cast<uint32_t>(it & 1);
        for (int l = 1; l <= nhid; ++l, ++wr) {
          wait_act();
          wait_b2();
          const bool is_skip = (l == g.skip);
          mma_fwd(is_skip ? base + 128u : base, hbuf(wr), 2, is_skip);   // skip: Lx(h) on top of Ly(x)
          umma_commit(bar(BAR_B2_EMPTY));
          umma_commit(bar(BAR_MMA_DONE));
        }
        if (view) {
          wait_act();
          wait_b2();
          mma_fwd(base, hbuf(wr), 2, false);                 // bottleneck part of directional_mlp.0
          umma_commit(bar(BAR_B2_EMPTY));
          wait_b2();
          mma_fwd(base, smem_base + sm.e, 1, true);          // Fourier part
          umma_commit(bar(BAR_B2_EMPTY));
          umma_commit(bar(BAR_MMA_DONE));
          ++wr;
        }
        if (kTrain) {
          wait_act();   // h_nb in buffer wr, dz_nb in buffer wr+1
          ++wr;
          for (int l = (p.debug & 4) ? 0 : nb; l >= 1; --l, ++wr) {
            wait_b2();
            mma_dgrad(base, hbuf(wr));                       // dh_{l-1} = dz_l * W_l
            umma_commit(bar(BAR_B2_EMPTY));
            umma_commit(bar(BAR_MMA_DONE));
            if (l > 1) wait_act();                           // dz_{l-1} written
          }
          ++wr;                                              // dz_0: written by the epilogue, consumed by nobody here
        }
      }
      put(10, now() - t_ch0); put(11, t_act); put(12, t_b2);
    }
   }
  } else {
    // =========================== epilogue warps ================================================
    setmaxnreg_inc<REGS_EPI>();
    const int r = tid;  // tile row == TMEM lane
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const float* s_bias = reinterpret_cast<const float*>(smem + sm.bias);
    const float* s_wlast = reinterpret_cast<const float*>(smem + sm.wlast);
    const float* s_blast = s_wlast + OUT_DIM * HID;
    uint32_t pdone = 0;
    uint32_t gph[2] = {0u, 0u};
    uint32_t wr = 0;
    float loss_abs = 0.f, loss_sq = 0.f;
    float gb[3] = {0.f, 0.f, 0.f};
    uint32_t mk[MAX_MMA_LAYERS][4];   // ReLU masks of this row (local memory; indexed by layer)

    long long t_done = 0, t_gfull = 0, t_sync = 0, t_store = 0, t_fwd = 0, t_final = 0, t_bwd = 0, t_pro = 0;
    const long long t_epi0 = now();
    auto wait_done = [&]() {
      const long long t0 = now();
      mbar_wait(bar(BAR_MMA_DONE), pdone, p.err, 0x40); pdone ^= 1u; tc_fence_after();
      t_done += now() - t0;
    };
    auto sync_epi = [&]() {
      const long long t0 = now();
      fence_proxy_async_smem();
      tc_fence_before();
      named_bar_sync(1, EPI_THREADS);
      t_sync += now() - t0;
    };
    auto hbuf = [&](uint32_t) { return smem_base + sm.h; };   // ONE activation buffer: see the bulk_wait_read<0> below

    int it = 0;
    for (int tile = blockIdx.x; tile < nt_loop; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t tb = tlane + 256u * static_cast<uint32_t>(buf);
      const int64_t ray = static_cast<int64_t>(tile) * TILE_M + r;
      const bool valid = ray < p.n_rays;
      const bool store_tile = tile < ntiles;   // a null tile (cluster padding) leaves no trace in the stash
      int64_t src = ray;   // row of target / dirs / face_idxs this ray reads (batch selection folded in)
      if (valid && p.ridx != nullptr)
        src = p.ridx64 ? static_cast<int64_t>(reinterpret_cast<const long long*>(p.ridx)[ray])
                       : static_cast<int64_t>(reinterpret_cast<const int*>(p.ridx)[ray]);
      uint8_t* stash_tile = kTrain ? p.stash + static_cast<size_t>(tile) * tile_stride : nullptr;

      // ---- view branch prologue: Fourier features of this ray -> E panel (model.py:175-186) --------
      const long long tp0 = now();
      if (view) {
        float x0 = 0.f, x1 = 0.f, x2 = 0.f;
        if (valid) {
          const float dx = p.dirs[src * 3 + 0], dy = p.dirs[src * 3 + 1], dz = p.dirs[src * 3 + 2];
          if (g.view == INF_VIEW_INTRINSIC) {
            const int64_t f = p.idx64 ? static_cast<int64_t>(reinterpret_cast<const long long*>(p.face_idxs)[src])
                                      : static_cast<int64_t>(reinterpret_cast<const int*>(p.face_idxs)[src]);
            const float nx = p.face_normals[f * 3 + 0], ny = p.face_normals[f * 3 + 1], nz = p.face_normals[f * 3 + 2];
            // F.cosine_similarity(-dir, n): each vector divided by max(|v|, 1e-8), then the dot product
            const float ia = 1.f / fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-8f);
            const float ib = 1.f / fmaxf(sqrtf(nx * nx + ny * ny + nz * nz), 1e-8f);
            float c = (-dx * ia) * (nx * ib) + (-dy * ia) * (ny * ib) + (-dz * ia) * (nz * ib);
            c = fminf(fmaxf(c, -1.f), 1.f);   // the reference yields NaN when rounding pushes |c| above 1
            x0 = acosf(c);
          } else {
            x0 = dx; x1 = dy; x2 = dz;
          }
        }
        const uint32_t row = smem_base + sm.e + static_cast<uint32_t>(r) * 128u;
        for (int ch = 0; ch < 8; ++ch) {
          float f[8];
          if (8 * ch < g.edim) {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = fourier_feature(8 * ch + i, x0, x1, x2, g.vdim, g.emb, g.include);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = 0.f;
          }
          st_shared_v4(row + ((static_cast<uint32_t>(ch) ^ (static_cast<uint32_t>(r) & 7u)) << 4), pack_half2(f[0], f[1]),
                       pack_half2(f[2], f[3]), pack_half2(f[4], f[5]), pack_half2(f[6], f[7]));
        }
      }

      t_pro += now() - tp0;
      // ---- forward layers 0..nb-1: h_l = relu(z_l + b_l) -> activation panel -------------------
      for (int l = 0; l < nb; ++l, ++wr) {
        if (l == 0) {
          const long long t0 = now();
          mbar_wait(bar(BAR_G_FULL + buf), gph[buf], p.err, 0x41);
          gph[buf] ^= 1u;
          tc_fence_after();
          t_gfull += now() - t0;
        } else {
          wait_done();
        }
        const long long tf0 = now();
        const uint32_t hb = hbuf(wr);
        const uint32_t col0 = (l == g.skip) ? tb + 128u : tb;
        uint32_t mbits[4];
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
          uint32_t v[32];
          tmem_ld32(col0 + 32u * cb, v);
          tmem_ld_wait();
          const float4* b4 = reinterpret_cast<const float4*>(s_bias + l * HID + 32 * cb);
          uint32_t w[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bb = b4[i];
            w[2 * i] = pack_half2_relu_sat(__uint_as_float(v[4 * i]) + bb.x, __uint_as_float(v[4 * i + 1]) + bb.y);
            w[2 * i + 1] = pack_half2_relu_sat(__uint_as_float(v[4 * i + 2]) + bb.z, __uint_as_float(v[4 * i + 3]) + bb.w);
          }
          if (kTrain) mbits[cb] = relu_mask16(w);
          store_act_words(hb, r, cb, w);
        }
        if (kTrain) { mk[l][0] = mbits[0]; mk[l][1] = mbits[1]; mk[l][2] = mbits[2]; mk[l][3] = mbits[3]; }
        sync_epi();
        if (tid == 0) {
          if (kTrain) {
            if (!(p.debug & 2) && store_tile) bulk_s2g_hint(stash_tile + stash_off_h(g, l), hb, BLOB_BYTES, stash_policy);
            if (l == 0 && view && store_tile) bulk_s2g_hint(stash_tile + stash_off_e(g), smem_base + sm.e, PANEL_BYTES, stash_policy);
            bulk_commit();
            const long long t0 = now();
            bulk_wait_read<0>();   // the next step overwrites this buffer: its stash store has left shared memory
            t_store += now() - t0;
          }
          mbar_arrive(bar(BAR_ACT_READY));
        }
        t_fwd += now() - tf0;
      }

      // ---- layer nb epilogue + output layer (-> 3, sigmoid) on CUDA cores -----------------------
      {
        wait_done();
        const long long tl0 = now();
        const uint32_t ha = hbuf(wr);
        const uint32_t col0 = (nb == g.skip) ? tb + 128u : tb;
        float y[3] = {s_blast[0], s_blast[1], s_blast[2]};
        uint32_t mbits[4];
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
          uint32_t v[32];
          tmem_ld32(col0 + 32u * cb, v);
          tmem_ld_wait();
          const float4* b4 = reinterpret_cast<const float4*>(s_bias + nb * HID + 32 * cb);
          const float4* w0 = reinterpret_cast<const float4*>(s_wlast + 0 * HID + 32 * cb);
          const float4* w1 = reinterpret_cast<const float4*>(s_wlast + 1 * HID + 32 * cb);
          const float4* w2 = reinterpret_cast<const float4*>(s_wlast + 2 * HID + 32 * cb);
          uint32_t w[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bb = b4[i], a0 = w0[i], a1 = w1[i], a2 = w2[i];
            const float h0 = fmaxf(__uint_as_float(v[4 * i]) + bb.x, 0.f), h1 = fmaxf(__uint_as_float(v[4 * i + 1]) + bb.y, 0.f);
            const float h2 = fmaxf(__uint_as_float(v[4 * i + 2]) + bb.z, 0.f), h3 = fmaxf(__uint_as_float(v[4 * i + 3]) + bb.w, 0.f);
            y[0] = fmaf(a0.w, h3, fmaf(a0.z, h2, fmaf(a0.y, h1, fmaf(a0.x, h0, y[0]))));
            y[1] = fmaf(a1.w, h3, fmaf(a1.z, h2, fmaf(a1.y, h1, fmaf(a1.x, h0, y[1]))));
            y[2] = fmaf(a2.w, h3, fmaf(a2.z, h2, fmaf(a2.y, h1, fmaf(a2.x, h0, y[2]))));
            if (kTrain) { w[2 * i] = pack_half2_sat(h0, h1); w[2 * i + 1] = pack_half2_sat(h2, h3); }
          }
          if (kTrain) {
            mbits[cb] = relu_mask16(w);
            store_act_words(ha, r, cb, w);
          }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) y[c] = sigmoidf_(y[c]);
        if (valid) {
          p.pred[ray * 3 + 0] = y[0];
          p.pred[ray * 3 + 1] = y[1];
          p.pred[ray * 3 + 2] = y[2];
        }
        if (!kTrain) {
          // all TMEM reads of this tile are complete: hand the accumulator buffer back to the GEMM1 issuer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(BAR_G_EMPTY + buf));
        } else {
          float dzl[3] = {0.f, 0.f, 0.f};
          if (valid) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              float gy;
              if (p.target != nullptr) {
                const float d = y[c] - p.target[src * 3 + c];
                loss_abs += fabsf(d);
                loss_sq += d * d;
                gy = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);  // d|d|/dd, sign(0) = 0 as F.l1_loss
              } else {
                gy = p.grad_pred[ray * 3 + c] * p.ext_scale;
              }
              dzl[c] = gy * y[c] * (1.f - y[c]) * DZ_SCALE;
              gb[c] += dzl[c];
            }
          }
          // dzl^T tile [16 x 128 rays] (rows 3..15 stay zero): B operand of the output-layer weight gradient
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            *reinterpret_cast<__half*>(smem + sm.dzlt + static_cast<uint32_t>(r >> 6) * 2048u + sw128_off(c, r & 63)) =
                __float2half_rn(dzl[c]);
          }
          // h_nb goes to the stash, then dz_nb takes its place in the activation buffer
          sync_epi();
          if (tid == 0) {
            if (!(p.debug & 2) && store_tile) bulk_s2g_hint(stash_tile + stash_off_h(g, nb), ha, BLOB_BYTES, stash_policy);
            bulk_commit();
            bulk_wait_read<0>();
          }
          sync_epi();
          // dz_nb = (dz_last * W_last) masked by relu'
          const uint32_t hb2 = hbuf(wr + 1);
#pragma unroll
          for (int cb = 0; cb < 4; ++cb) {
            const float4* w0 = reinterpret_cast<const float4*>(s_wlast + 0 * HID + 32 * cb);
            const float4* w1 = reinterpret_cast<const float4*>(s_wlast + 1 * HID + 32 * cb);
            const float4* w2 = reinterpret_cast<const float4*>(s_wlast + 2 * HID + 32 * cb);
            uint32_t w[16];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 a0 = w0[i], a1 = w1[i], a2 = w2[i];
              const float d0 = fmaf(dzl[2], a2.x, fmaf(dzl[1], a1.x, dzl[0] * a0.x));
              const float d1 = fmaf(dzl[2], a2.y, fmaf(dzl[1], a1.y, dzl[0] * a0.y));
              const float d2 = fmaf(dzl[2], a2.z, fmaf(dzl[1], a1.z, dzl[0] * a0.z));
              const float d3 = fmaf(dzl[2], a2.w, fmaf(dzl[1], a1.w, dzl[0] * a0.w));
              w[2 * i] = apply_mask(pack_half2_sat(d0, d1), mbits[cb], 2 * i);
              w[2 * i + 1] = apply_mask(pack_half2_sat(d2, d3), mbits[cb], 2 * i + 1);
            }
            store_act_words(hb2, r, cb, w);
          }
          sync_epi();
          if (tid == 0) {
            if (!(p.debug & 2) && store_tile) bulk_s2g_hint(stash_tile + stash_off_dz(g, nb), hb2, BLOB_BYTES, stash_policy);
            if (store_tile) bulk_s2g_hint(stash_tile + stash_off_dzlt(g), smem_base + sm.dzlt, DZLT_BYTES, stash_policy);
            bulk_commit();
            bulk_wait_read<0>();
            mbar_arrive(bar(BAR_ACT_READY));
          }
          wr += 2;
          t_final += now() - tl0;
          if (p.debug & 4) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(BAR_G_EMPTY + buf));
          }
          // ---- dgrad chain: dz_l = dh_l * relu'(z_l), l = nb-1..0 --------------------------------
          for (int lb = (p.debug & 4) ? -1 : nb - 1; lb >= 0; --lb, ++wr) {
            wait_done();
            const long long tb0 = now();
            const uint32_t hb = hbuf(wr);
            const uint32_t mm[4] = {mk[lb][0], mk[lb][1], mk[lb][2], mk[lb][3]};
#pragma unroll
            for (int cb = 0; cb < 4; ++cb) {
              uint32_t v[32];
              tmem_ld32(tb + 32u * cb, v);
              tmem_ld_wait();
              uint32_t w[16];
#pragma unroll
              for (int i = 0; i < 16; ++i)
                w[i] = apply_mask(pack_half2_sat(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), mm[cb], i);
              store_act_words(hb, r, cb, w);
            }
            if (lb == 0) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar(BAR_G_EMPTY + buf));   // last TMEM read of this tile
            }
            sync_epi();
            if (tid == 0) {
              if (!(p.debug & 2) && store_tile) bulk_s2g_hint(stash_tile + stash_off_dz(g, lb), hb, BLOB_BYTES, stash_policy);
              bulk_commit();
              bulk_wait_read<0>();
              if (lb > 0) mbar_arrive(bar(BAR_ACT_READY));
            }
            // no MMA gates the next tile's first write into the activation buffer: everyone waits for dz_0 to leave it
            if (lb == 0) named_bar_sync(1, EPI_THREADS);
            t_bwd += now() - tb0;
          }
        }
      }
    }
    if (tid == 0) {
      put(16, now() - t_epi0); put(17, t_gfull); put(18, t_done); put(19, t_sync); put(20, t_store);
      put(21, t_fwd); put(22, t_final); put(23, t_bwd); put(24, t_pro);
    }
    if (kTrain) {
      // ---- per-CTA partials: output-layer bias gradient, loss sums ------------------------------
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        loss_abs += __shfl_xor_sync(0xffffffffu, loss_abs, o);
        loss_sq += __shfl_xor_sync(0xffffffffu, loss_sq, o);
#pragma unroll
        for (int c = 0; c < 3; ++c) gb[c] += __shfl_xor_sync(0xffffffffu, gb[c], o);
      }
      float* s_red = reinterpret_cast<float*>(smem + sm.red);   // [5][4]: per-warp partial sums
      if (lane == 0) {
        s_red[warp] = loss_abs; s_red[4 + warp] = loss_sq;
        s_red[8 + warp] = gb[0]; s_red[12 + warp] = gb[1]; s_red[16 + warp] = gb[2];
      }
      named_bar_sync(1, EPI_THREADS);
      if (tid == 0) {
        float* part = p.part1 + static_cast<size_t>(blockIdx.x) * K1_PART_FLOATS;
        part[0] = s_red[8] + s_red[9] + s_red[10] + s_red[11];
        part[1] = s_red[12] + s_red[13] + s_red[14] + s_red[15];
        part[2] = s_red[16] + s_red[17] + s_red[18] + s_red[19];
        part[3] = 0.f;
        part[4] = s_red[0] + s_red[1] + s_red[2] + s_red[3];
        part[5] = s_red[4] + s_red[5] + s_red[6] + s_red[7];
        bulk_wait<0>();  // stash stores fully written before the grid completes
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (cs > 1) cluster_sync_all();   // no CTA leaves while a peer may still multicast into it or arrive on its barriers
  if (warp == W_G1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---- host launcher (kernels stay private to this translation unit) -----------------------------
template <bool kTrain, typename TableT>
static cudaError_t launch_one(const FieldParams& p, int grid, cudaStream_t stream) {
  const int smem = static_cast<int>(smem_layout(p.geo.view != INF_VIEW_NONE, kTrain).total);
  // exactly this launch's footprint: the driver sizes the L1 / shared-memory split from it
  cudaError_t e = cudaFuncSetAttribute(field_main_kernel<kTrain, TableT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  {
    // ask for the smallest carve-out that holds one CTA (+1 KB the system reserves): what is left of the 256 KB array
    // is the L1 the gather's loads are tracked in
    static const int env_carve = getenv("INF_CARVEOUT") ? atoi(getenv("INF_CARVEOUT")) : -2;
    if (env_carve != -2) {
      e = cudaFuncSetAttribute(field_main_kernel<kTrain, TableT>, cudaFuncAttributePreferredSharedMemoryCarveout, env_carve);
      if (e != cudaSuccess) return e;
    }
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(NUM_THREADS_MAIN);
  cfg.dynamicSmemBytes = static_cast<size_t>(smem);
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(p.cluster);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  add_pdl_attr(cfg, attr);
  return cudaLaunchKernelEx(&cfg, field_main_kernel<kTrain, TableT>, p);
}

// Largest grid (a multiple of `cluster`) whose clusters are all co-resident with this kernel's footprint, 0 on error.
template <bool kTrain, typename TableT>
static int max_grid_one(int cluster, bool view) {
  cudaFuncSetAttribute(field_main_kernel<kTrain, TableT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       static_cast<int>(smem_layout(view, kTrain).total));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(cluster));
  cfg.blockDim = dim3(NUM_THREADS_MAIN);
  cfg.dynamicSmemBytes = smem_layout(view, kTrain).total;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(cluster);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, field_main_kernel<kTrain, TableT>, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n * cluster;
}

int field_main_max_grid(bool train, bool table_f16, int cluster, bool view) {
  if (train) return table_f16 ? max_grid_one<true, __half>(cluster, view) : max_grid_one<true, float>(cluster, view);
  return table_f16 ? max_grid_one<false, __half>(cluster, view) : max_grid_one<false, float>(cluster, view);
}

cudaError_t launch_field_main(const FieldParams& p, bool train, bool table_f16, int grid, cudaStream_t stream) {
  if (train) return table_f16 ? launch_one<true, __half>(p, grid, stream) : launch_one<true, float>(p, grid, stream);
  return table_f16 ? launch_one<false, __half>(p, grid, stream) : launch_one<false, float>(p, grid, stream);
}

}  // namespace inf
