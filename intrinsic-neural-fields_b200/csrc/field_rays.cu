// Per-ray kernel of the vertex-projected path, sm_100a: gather -> interpolate -> hidden chain -> RGB
// (+ L1 loss + dgrad chain when training).
//
// vertex_project.cu has already contracted the table with [W0;Wy] per vertex, so a ray's layer-0 pre-activation and
// its skip-branch term are   [z0 | zy] = sum_j bary_j P[v_j]   (mesh.py:313-324 + model.py:98-104 + layers.py:60-62 by
// linearity): three coalesced 512-byte row reads from a table that lives in L2, interpolated in fp32.  What is left
// per ray is the 128-wide chain (model.py:105-112 / 168-173), which is latency-bound per tile (MMA 0.3 us, epilogue
// ~1 us per layer), so a CTA runs TWO 128-ray tiles in lockstep: one hidden-weight image streamed from L2 feeds both
// tiles' MMAs, and the two epilogue warpgroups overlap each other and the other tile's MMA.
//
//   warps 0..3 / 4..7   slot 0 / slot 1: gather + interpolation of the tile's rays (lane = 16-byte chunk of a P row),
//                       layer-0 bias + ReLU -> fp16 activation image X, zy -> fp16 image Y (added in the skip layer's
//                       epilogue); then per layer TMEM -> registers -> bias/ReLU -> next A operand; output layer,
//                       sigmoid, L1 gradient and the dgrad epilogues; the backward chain ping-pongs between X and Y so
//                       that no step waits for the stash store of the image it overwrites
//   warp 8              chain issuer (+ TMEM allocator): per layer step, the MMAs of both slots from one weight stage
//   warp 9              weight loader: 32 KB hidden / view-branch operand images through a 2-stage ring
//   warps 10 / 11       training: stash writers of slot 0 / 1 -- every h_l / dz_l image is bulk-stored to the stash
//                       (hidden weight gradients: field_wgrad.cu; dz_0 / dz_skip also feed the vertex gradient G,
//                       vertex_grad.cu) off the epilogue's critical path
#include "common.cuh"
#include "ptx.cuh"

namespace inf {

namespace {

// Two flavours of the same kernel (template parameter kSlots):
//   2  one CTA per SM running two tiles in lockstep (one weight stage feeds both tiles' MMAs), 12 warps
//   1  two CTAs per SM, one tile each, 8 warps: the CTAs drift apart, so one's gather overlaps the other's chain
constexpr int RY_SLOTS = 2;   // maximum (sizes the barrier table)
// measured variants: request the next tile's gather metadata one tile ahead / the L1 target at
// the head of the tile / wait for the previous tile's last stash stores only after the first P rows are requested
#ifndef RY_META_PF
#define RY_META_PF 0
#endif
#ifndef RY_TGT_PF
#define RY_TGT_PF 0
#endif
#ifndef RY_LATE_FREE
#define RY_LATE_FREE 0
#endif
// A tile is served by EIGHT gather/epilogue warps: warp (quarter wq, column half ch) owns TMEM lanes 32 wq .. 32 wq + 31 and
// columns 64 ch .. 64 ch + 63 of every layer, so a layer step is two 32-column TMEM loads per thread instead of four and
// twice as many warps hide the TMEM / shared-memory / barrier latencies of the (latency-bound) chain.
__host__ __device__ constexpr int ry_threads(int slots) { return (8 * slots + 4) * 32; }
// registers: launch value 96 x 640 (or 80 x 384, two CTAs per SM); the epilogue warpgroups grow, the utility warpgroup
// shrinks to 40
constexpr int RY_REGS_UTIL = 40;
__host__ __device__ constexpr int ry_regs_epi(int slots) { return slots == 2 ? 104 : 96; }
constexpr int RY_WG_THREADS = 256;

struct RySmem {
  uint32_t h, zy, w, dzlt, bias, wlast, ysum, red, bars, tmemptr, total;
};
__host__ __device__ constexpr RySmem ry_smem(bool train, int slots) {
  RySmem s{};
  s.h = 0;                                                   // [slots] activation images X (32 KB each)
  s.zy = s.h + slots * BLOB_BYTES;                           // [slots] images Y: skip-branch term zy; view: Fourier panel; backward ping-pong
  s.w = s.zy + slots * BLOB_BYTES;                           // weight ring, `slots` stages
  s.dzlt = s.w + slots * BLOB_BYTES;                         // [slots] dzl^T tiles (training)
  s.bias = s.dzlt + (train ? slots * DZLT_BYTES : 0u);       // f32 [MAX_MMA_LAYERS][128]
  s.wlast = s.bias + MAX_MMA_LAYERS * HID * 4;               // f32 [3][128] + blast[4]
  s.ysum = s.wlast + OUT_DIM * HID * 4 + 16;                 // f32 [slots][2 column halves][128 rows][3]: output-layer partial sums
  s.red = s.ysum + slots * 2 * TILE_M * OUT_DIM * 4;         // f32 [slots][24]
  s.bars = s.red + slots * 96;
  s.tmemptr = s.bars + 32 * 8;
  s.total = s.tmemptr + 16;
  return s;
}
static_assert(ry_smem(true, 2).total <= 232448 && 2 * (ry_smem(true, 1).total + 1024) <= 233472, "shared memory budget");

constexpr int RB_ACT_READY = 0;                    // [2] activation image of the slot written (A operand of the next step)
constexpr int RB_MMA_DONE = RB_ACT_READY + RY_SLOTS;   // [2] the slot's accumulator is complete
constexpr int RB_FREE_X = RB_MMA_DONE + RY_SLOTS;      // [2] training: the stash store has finished reading image X of the slot
constexpr int RB_FREE_Y = RB_FREE_X + RY_SLOTS;        // [2] ... image Y
// [2][2] an image of the slot is ready to be stashed.  Requests alternate between two barriers: an mbarrier must not
// complete two phases before its waiter has observed the first, and request i+2 is only ever posted after the image of
// request i was handed back (FREE_X / FREE_Y), i.e. after the stash writer consumed request i.
constexpr int RB_STASH_REQ = RB_FREE_Y + RY_SLOTS;
constexpr int RB_W_FULL = RB_STASH_REQ + 2 * RY_SLOTS; // [stages]
constexpr int RB_W_EMPTY = RB_W_FULL + RY_SLOTS;
static_assert(RB_W_EMPTY + RY_SLOTS <= 32, "barrier table overflow");

constexpr uint32_t RY_IDESC_N128 = umma_idesc_f16(128, 128, 0, 0);
constexpr uint32_t RY_IDESC_N128_BMN = umma_idesc_f16(128, 128, 0, 1);
constexpr uint32_t RY_DESC_HI = umma_desc_hi(1024);
constexpr float RY_PI_F = 3.14159274101257324f;   // float32(torch.pi), layers.py:14

__device__ __forceinline__ float ry_sigmoid(float v) { return __fdividef(1.f, 1.f + __expf(-v)); }

// 16 packed f16x2 words (columns 32*cb .. 32*cb+31 of row r) of a 2-panel [128 x 128] image
__device__ __forceinline__ uint32_t act_chunk_addr(uint32_t img, int r, int cb, int i) {
  const uint32_t panel = img + static_cast<uint32_t>(cb >> 1) * PANEL_BYTES + static_cast<uint32_t>(r) * 128u;
  const uint32_t chunk = static_cast<uint32_t>((cb & 1) * 4 + i);
  return panel + ((chunk ^ (static_cast<uint32_t>(r) & 7u)) << 4);
}
__device__ __forceinline__ void ry_store_words(uint32_t img, int r, int cb, const uint32_t (&w)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) st_shared_v4(act_chunk_addr(img, r, cb, i), w[4 * i + 0], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
}
__device__ __forceinline__ void ry_load_words(uint32_t img, int r, int cb, uint32_t (&w)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(w[4 * i + 0]), "=r"(w[4 * i + 1]), "=r"(w[4 * i + 2]), "=r"(w[4 * i + 3])
                 : "r"(act_chunk_addr(img, r, cb, i)));
}
// bit i = (low half of word i != 0), bit 16+i = (high half != 0), for non-negative halves (see field_main.cu)
__device__ __forceinline__ uint32_t ry_relu_mask16(const uint32_t (&w)[16]) {
  uint32_t m = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) m |= ((w[i] + 0x7FFF7FFFu) >> (15 - i)) & (0x00010001u << i);
  return m;
}
__device__ __forceinline__ uint32_t ry_apply_mask(uint32_t w, uint32_t m, int i) {
  return w & (((m >> i) & 0x00010001u) * 0xFFFFu);
}
__device__ __forceinline__ float2 ry_unpack(uint32_t w) { return __half22float2(*reinterpret_cast<const __half2*>(&w)); }

// layers.py:21-25 (use_logspace=True): cat(cos(x f) [input-major, freq-minor], sin(x f), x), zero beyond the embedding
__device__ __forceinline__ float ry_fourier(int idx, float x0, float x1, float x2, int d, int emb, int include) {
  const int nf = d * emb;
  if (idx < 2 * nf) {
    const int q = idx < nf ? idx : idx - nf;
    const int dim = q / emb, j = q - dim * emb;
    const float x = dim == 0 ? x0 : (dim == 1 ? x1 : x2);
    const float arg = x * ldexpf(RY_PI_F, j);
    return idx < nf ? cosf(arg) : sinf(arg);
  }
  if (include && idx < 2 * nf + d) {
    const int dim = idx - 2 * nf;
    return dim == 0 ? x0 : (dim == 1 ? x1 : x2);
  }
  return 0.f;
}

template <int N>
__device__ __forceinline__ void ry_setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void ry_setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

}  // namespace

template <bool kTrain, int kSlots>
__global__ void __launch_bounds__(ry_threads(kSlots), kSlots == 1 ? 2 : 1) field_rays_kernel(const FieldParams p) {
  constexpr int RY_THREADS = ry_threads(kSlots);
  constexpr int RY_EPI_WARPS = 8 * kSlots;
  constexpr int RY_W_CH = RY_EPI_WARPS, RY_W_LD = RY_EPI_WARPS + 1, RY_W_ST = RY_EPI_WARPS + 2;   // stash writer: + slot
  constexpr int RY_NSTAGE_W = kSlots;
  constexpr uint32_t RY_TMEM_COLS = 128u * kSlots;   // one 128-column accumulator per slot
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const Geo g = p.geo;
  const bool view = g.view != INF_VIEW_NONE;
  const RySmem sm = ry_smem(kTrain, kSlots);
  const int nb = g.nb, nhid = g.nhid;
  auto bar = [&](int i) -> uint32_t { return smem_base + sm.bars + 8u * static_cast<uint32_t>(i); };
  // measurement only (INF_DEBUG & 8): per-role cycle accounting of CTA 0 (scripts/trace_rays.py)
  const bool tracing = p.trace != nullptr && blockIdx.x == 0;
  auto now = [&]() -> long long { return tracing ? clock64() : 0ll; };
  auto put = [&](int slot_, long long v) { if (tracing) p.trace[slot_] = static_cast<unsigned long long>(v); };

  if ((smem_base & 1023u) != 0) {
    if (tid == 0 && p.err) atomicExch(p.err, 0xBAD00004u);
    return;
  }
  if (tid == 0) {
    for (int s = 0; s < kSlots; ++s) {
      mbar_init(bar(RB_ACT_READY + s), 1);
      mbar_init(bar(RB_MMA_DONE + s), 1);
      mbar_init(bar(RB_FREE_X + s), 1);
      mbar_init(bar(RB_FREE_Y + s), 1);
      mbar_init(bar(RB_STASH_REQ + 2 * s), 1);
      mbar_init(bar(RB_STASH_REQ + 2 * s + 1), 1);
    }
    for (int i = 0; i < RY_NSTAGE_W; ++i) { mbar_init(bar(RB_W_FULL + i), 1); mbar_init(bar(RB_W_EMPTY + i), 1); }
    mbar_fence_init();
  }
  if (warp == RY_W_CH) {
    tmem_alloc(smem_base + sm.tmemptr, RY_TMEM_COLS);
    tmem_relinquish();
  }
  if (p.debug & DEBUG_PDL_EARLY) pdl_launch_dependents();
  pdl_wait();   // P (vertex_project), the operand images, the stash's previous readers
  {
    const float* f32sec = reinterpret_cast<const float*>(p.packed + packed_f32_offset(g));
    float* s_bias = reinterpret_cast<float*>(smem + sm.bias);
    float* s_wl = reinterpret_cast<float*>(smem + sm.wlast);
    const int nbias = (nb + 1) * HID;
    const int nf = packed_f32_count(g);
    for (int i = tid; i < nf; i += RY_THREADS) {
      const float v = f32sec[i];
      if (i < nbias) s_bias[i] = v;
      else s_wl[i - nbias] = v;
    }
    if (kTrain) {
      uint32_t* dz = reinterpret_cast<uint32_t*>(smem + sm.dzlt);   // rows 3..15 of both dzl^T tiles stay zero
      for (int i = tid; i < kSlots * DZLT_BYTES / 4; i += RY_THREADS) dz[i] = 0u;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + sm.tmemptr);

  const int ntiles = p.ntiles;
  const int npairs = (ntiles + kSlots - 1) / kSlots;
  const size_t tile_stride = p.stash_stride;
  const uint64_t stash_policy = 0ull;   // the stash is re-read by the wgrad / vertex-gradient kernels right after this launch

  // weight images of one tile pair, in consumption order: hidden 1..nhid, [view: Wd1z, Wd1e], then the dgrad chain
  // [view: Wd1z], hidden nhid..1
  const int steps_fwd = nhid + (view ? 2 : 0);
  const int steps_all = steps_fwd + (kTrain ? nb : 0);
  auto step_blob = [&](int st, uint32_t& bytes) -> int {
    bytes = BLOB_BYTES;
    if (st < nhid) return blob_hidden(g, st + 1);
    if (st < steps_fwd) {
      if (st == nhid) return blob_wd1z(g);
      bytes = PANEL_BYTES;
      return blob_wd1e(g);
    }
    const int l = nb - (st - steps_fwd);   // dgrad of MMA layer l = nb..1
    if (view && l == nb) return blob_wd1z(g);
    return blob_hidden(g, l);
  };

  if (warp >= RY_EPI_WARPS) {
   ry_setmaxnreg_dec<RY_REGS_UTIL>();
   if (warp == RY_W_LD) {
    // =========================== weight loader ==============================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
        for (int st = 0; st < steps_all; ++st) {
          uint32_t bytes;
          const int blob = step_blob(st, bytes);
          mbar_wait(bar(RB_W_EMPTY + stage), phase ^ 1u, p.err, 0xB0);
          mbar_arrive_expect_tx(bar(RB_W_FULL + stage), bytes);
          bulk_g2s(smem_base + sm.w + static_cast<uint32_t>(stage) * BLOB_BYTES, p.packed + static_cast<size_t>(blob) * BLOB_BYTES, bytes,
                   bar(RB_W_FULL + stage));
          if (++stage == RY_NSTAGE_W) { stage = 0; phase ^= 1u; }
        }
      }
    }
   } else if (warp == RY_W_CH) {
    // =========================== chain issuer ===============================================
    if (lane == 0) {
      int stage = 0;
      uint32_t wph = 0;
      uint32_t aph[2] = {0u, 0u};
      long long t_w = 0, t_act = 0;
      const long long t_ch0 = now();
      // D[128 x 128] (+)= A[128 x 64*npanel] * B^T, both K-major panel images
      auto mma_fwd = [&](uint32_t dcol, uint32_t abase, uint32_t bbase, int npanel, bool acc_first) {
        const uint32_t alo = umma_desc_lo(abase, 0), blo = umma_desc_lo(bbase, 0);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          if (kk < 4 * npanel) {
            const uint32_t o = static_cast<uint32_t>(kk >> 2) * (PANEL_BYTES >> 4) + static_cast<uint32_t>(kk & 3) * 2u;
            umma_f16(dcol, umma_desc_join(alo + o, RY_DESC_HI), umma_desc_join(blo + o, RY_DESC_HI), RY_IDESC_N128,
                     (acc_first || kk > 0) ? 1u : 0u);
          }
        }
      };
      // dh[128 x 128] = dz[128 x 128] * W: A K-major, B = the forward weight image read MN-major (K = row n)
      auto mma_dgrad = [&](uint32_t dcol, uint32_t abase, uint32_t bbase) {
        const uint32_t alo = umma_desc_lo(abase, 0), blo_mn = umma_desc_lo(bbase, PANEL_BYTES);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t oa = static_cast<uint32_t>(kk >> 2) * (PANEL_BYTES >> 4) + static_cast<uint32_t>(kk & 3) * 2u;
          umma_f16(dcol, umma_desc_join(alo + oa, RY_DESC_HI), umma_desc_join(blo_mn + static_cast<uint32_t>(kk) * 128u, RY_DESC_HI),
                   RY_IDESC_N128_BMN, kk > 0 ? 1u : 0u);
        }
      };
      for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
        const int nslot = (kSlots == 2 && pair * kSlots + 1 < ntiles) ? 2 : 1;
        for (int st = 0; st < steps_all; ++st) {
          const long long tw0 = now();
          mbar_wait(bar(RB_W_FULL + stage), wph, p.err, 0xB1);
          t_w += now() - tw0;
          tc_fence_after();
          const uint32_t bbase = smem_base + sm.w + static_cast<uint32_t>(stage) * BLOB_BYTES;
          const bool fourier = view && st == nhid + 1;   // second half of the view layer: no new activation, same accumulator
          const bool view_z = view && st == nhid;        // first half: the accumulator is completed by the Fourier part
          for (int s = 0; s < nslot; ++s) {
            const uint32_t dcol = tmem_base + 128u * static_cast<uint32_t>(s);
            const uint32_t ximg = smem_base + sm.h + static_cast<uint32_t>(s) * BLOB_BYTES;
            const uint32_t yimg = smem_base + sm.zy + static_cast<uint32_t>(s) * BLOB_BYTES;
            if (!fourier) {
              const long long ta0 = now();
              mbar_wait(bar(RB_ACT_READY + s), aph[s], p.err, 0xB2);
              t_act += now() - ta0;
              aph[s] ^= 1u;
              tc_fence_after();
            }
            if (st < steps_fwd) {
              if (fourier) mma_fwd(dcol, yimg, bbase, 1, true);
              else mma_fwd(dcol, ximg, bbase, 2, false);
            } else {
              // the backward chain ping-pongs: dz_nb sits in Y, dz_{nb-1} in X, ...
              mma_dgrad(dcol, ((st - steps_fwd) & 1) ? ximg : yimg, bbase);
            }
            if (!view_z) umma_commit(bar(RB_MMA_DONE + s));
          }
          umma_commit(bar(RB_W_EMPTY + stage));
          if (++stage == RY_NSTAGE_W) { stage = 0; wph ^= 1u; }
        }
      }
      put(0, now() - t_ch0); put(1, t_w); put(2, t_act);
    }
   } else if (kTrain && warp < RY_W_ST + kSlots) {
    // =========================== stash writer of slot (warp - RY_W_ST) ==========================
    // Mirrors the epilogue's sequence of finished images; every request is one bulk store (two where noted) whose
    // shared-memory read is waited for HERE, then the image is handed back (FREE_X / FREE_Y).
    if (lane == 0) {
      const int slot = warp - RY_W_ST;
      const uint32_t ximg = smem_base + sm.h + static_cast<uint32_t>(slot) * BLOB_BYTES;
      const uint32_t yimg = smem_base + sm.zy + static_cast<uint32_t>(slot) * BLOB_BYTES;
      const uint32_t dzlt = smem_base + sm.dzlt + static_cast<uint32_t>(slot) * DZLT_BYTES;
      uint32_t nreq = 0;   // running request index: barrier nreq & 1, phase (nreq >> 1) & 1
      auto req = [&]() { mbar_wait(bar(RB_STASH_REQ + 2 * slot + static_cast<int>(nreq & 1u)), (nreq >> 1) & 1u, p.err, 0xD0); ++nreq; };
      auto done = [&](bool x, bool y) {
        bulk_commit();
        bulk_wait_read<0>();
        if (x) mbar_arrive(bar(RB_FREE_X + slot));
        if (y) mbar_arrive(bar(RB_FREE_Y + slot));
      };
      for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
        const int tile = pair * kSlots + slot;
        if (tile >= ntiles) break;
        uint8_t* st = p.stash + static_cast<size_t>(tile) * tile_stride;
        for (int l = 0; l < nb; ++l) {            // h_l from X (+ the Fourier panel from Y with the last of them)
          req();
          bulk_s2g_hint(st + stash_off_h(g, l), ximg, BLOB_BYTES, stash_policy);
          const bool e = view && l == nb - 1;
          if (e) bulk_s2g_hint(st + stash_off_e(g), yimg, PANEL_BYTES, stash_policy);
          done(true, e);
        }
        req();                                      // h_nb from X, dz_nb from Y, dzl^T: one request
        bulk_s2g_hint(st + stash_off_h(g, nb), ximg, BLOB_BYTES, stash_policy);
        bulk_s2g_hint(st + stash_off_dz(g, nb), yimg, BLOB_BYTES, stash_policy);
        bulk_s2g_hint(st + stash_off_dzlt(g), dzlt, DZLT_BYTES, stash_policy);
        done(true, true);
        for (int i = 0; i < nb; ++i) {              // dz_{nb-1-i} from X (i even) / Y (i odd)
          req();
          bulk_s2g_hint(st + stash_off_dz(g, nb - 1 - i), (i & 1) ? yimg : ximg, BLOB_BYTES, stash_policy);
          done(!(i & 1), (i & 1) != 0);
        }
      }
      bulk_wait<0>();   // stash stores fully written before the grid completes
    }
   }
  } else {
    // =========================== gather + epilogue warpgroups ==================================
    ry_setmaxnreg_inc<ry_regs_epi(kSlots)>();
    const int slot = warp >> 3;                 // eight warps = tile slot
    const int ws = warp & 7;
    const int wq = ws & 3;                      // TMEM lane quarter of this warp
    const int ch = ws >> 2;                     // column half of every layer this warp owns
    const int r = 32 * wq + lane;               // tile row == TMEM lane
    const bool leader = ws == 0 && lane == 0;   // one thread per slot talks to the MMA issuer / stash writer
    const uint32_t tb = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + 128u * static_cast<uint32_t>(slot);
    const uint32_t ximg = smem_base + sm.h + static_cast<uint32_t>(slot) * BLOB_BYTES;
    const uint32_t yimg = smem_base + sm.zy + static_cast<uint32_t>(slot) * BLOB_BYTES;
    const uint32_t dzlt = sm.dzlt + static_cast<uint32_t>(slot) * DZLT_BYTES;
    const float* s_bias = reinterpret_cast<const float*>(smem + sm.bias);
    const float* s_wlast = reinterpret_cast<const float*>(smem + sm.wlast);
    const float* s_blast = s_wlast + OUT_DIM * HID;
    const uint32_t bar_id = 1u + static_cast<uint32_t>(slot);
    uint32_t pdone = 0, pfx = 0, pfy = 0, nreq = 0;
    bool stored_x = false, stored_y = false;   // a stash store of the image is outstanding (training)
    float loss_abs = 0.f, loss_sq = 0.f;
    float gb[3] = {0.f, 0.f, 0.f};
    uint32_t mk[MAX_MMA_LAYERS][2];   // ReLU masks of this row's 64 columns (local memory; indexed by layer)
    float* s_ysum = reinterpret_cast<float*>(smem + sm.ysum) + slot * 2 * TILE_M * OUT_DIM;

    long long t_done = 0, t_free = 0, t_sync = 0, t_gather = 0, t_fwd = 0, t_final = 0, t_bwd = 0;
    const long long t_epi0 = now();
    auto wait_done = [&]() {
      const long long t0 = now();
      mbar_wait(bar(RB_MMA_DONE + slot), pdone, p.err, 0xC0); pdone ^= 1u; tc_fence_after();
      t_done += now() - t0;
    };
    // before an image is overwritten: its stash store (if any) has finished reading it
    auto free_x = [&]() {
      if (kTrain && stored_x) { const long long t0 = now(); mbar_wait(bar(RB_FREE_X + slot), pfx, p.err, 0xC1); pfx ^= 1u; stored_x = false; t_free += now() - t0; }
    };
    auto free_y = [&]() {
      if (kTrain && stored_y) { const long long t0 = now(); mbar_wait(bar(RB_FREE_Y + slot), pfy, p.err, 0xC2); pfy ^= 1u; stored_y = false; t_free += now() - t0; }
    };
    auto sync_epi = [&]() {
      const long long t0 = now();
      fence_proxy_async_smem();
      tc_fence_before();
      named_bar_sync(bar_id, RY_WG_THREADS);
      t_sync += now() - t0;
    };
    // after sync_epi: hand the finished image(s) to the stash writer (training) and / or release the next MMA
    auto publish = [&](bool x, bool y, bool arrive_act) {
      if (kTrain) {
        stored_x = stored_x || x;
        stored_y = stored_y || y;
        if (leader) mbar_arrive(bar(RB_STASH_REQ + 2 * slot + static_cast<int>(nreq & 1u)));
        ++nreq;
      }
      if (arrive_act && leader) mbar_arrive(bar(RB_ACT_READY + slot));
    };

    // Gather metadata of a tile (warp ws gathers rays 16 ws .. 16 ws + 15: lanes 0..15 and 16..31 hold the same ray's vertex
    // ids and barycentrics).  It is a chain of dependent DRAM reads (ray_idxs -> vids / bary), so the NEXT tile's metadata
    // is requested while this tile's chain runs and has long arrived when its gather starts.
    auto load_meta = [&](int tile_, uint32_t (&mv)[3], float (&mb)[3]) {
#pragma unroll
      for (int j = 0; j < 3; ++j) { mv[j] = 0u; mb[j] = 0.f; }
      const int64_t gray = static_cast<int64_t>(tile_) * TILE_M + 16 * ws + (lane & 15);
      if (tile_ < ntiles && gray < p.n_rays) {
        int64_t gsrc = gray;
        if (p.ridx != nullptr)
          gsrc = p.ridx64 ? static_cast<int64_t>(reinterpret_cast<const long long*>(p.ridx)[gray])
                          : static_cast<int64_t>(reinterpret_cast<const int*>(p.ridx)[gray]);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int64_t e = gsrc * 3 + j;
          mv[j] = p.idx64 ? static_cast<uint32_t>(reinterpret_cast<const long long*>(p.vids)[e])
                          : static_cast<uint32_t>(reinterpret_cast<const int*>(p.vids)[e]);
          mb[j] = p.bary[e];
        }
      }
    };
    uint32_t nxv[3];
    float nxb[3];
    if (RY_META_PF) load_meta(static_cast<int>(blockIdx.x) * kSlots + slot, nxv, nxb);

    for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
      const int tile = pair * kSlots + slot;
      if (tile >= ntiles) break;   // odd tile count: slot 1 sits out the last pair
      const int64_t ray = static_cast<int64_t>(tile) * TILE_M + r;
      const bool valid = ray < p.n_rays;
      int64_t src = ray;   // row of the per-ray arrays this ray reads (batch selection folded in, ray_dataloader.py:122-126)
      if (valid && p.ridx != nullptr)
        src = p.ridx64 ? static_cast<int64_t>(reinterpret_cast<const long long*>(p.ridx)[ray])
                       : static_cast<int64_t>(reinterpret_cast<const int*>(p.ridx)[ray]);
      // the L1 target of this thread's ray: requested now (second link of a dependent DRAM chain), used after the last layer
      float tgt[3] = {0.f, 0.f, 0.f};
      if (RY_TGT_PF && kTrain && valid && p.target != nullptr) {
#pragma unroll
        for (int c = 0; c < 3; ++c) tgt[c] = p.target[src * 3 + c];
      }

      // ---- gather + interpolate: [z0 | zy] = sum_j bary_j P[v_j] ------------------------------------
      // The warp walks its 16 rays, lane = 16-byte chunk of the 512-byte P rows (fully coalesced), RU rays = 3 RU loads in
      // flight per lane.
      const long long tg0 = now();
      uint32_t myv[3];
      float myb[3];
      if (RY_META_PF) {
#pragma unroll
        for (int j = 0; j < 3; ++j) { myv[j] = nxv[j]; myb[j] = nxb[j]; }
      } else {
        load_meta(tile, myv, myb);
      }
      if (!RY_LATE_FREE) {
        free_x();   // the previous tile's last dz image has left X / Y
        free_y();
      }
      {
        constexpr int RU = 4;
        const bool z0_lane = lane < 16;
        const uint32_t blk = static_cast<uint32_t>(lane & 15);      // 8 columns of the 128-column half
        float b0[8];   // bias of layer 0 for this lane's columns (z0 lanes only)
#pragma unroll
        for (int i = 0; i < 8; ++i) b0[i] = z0_lane ? s_bias[8 * blk + i] : 0.f;
        const uint32_t dst_img = (z0_lane ? ximg : yimg) + (blk >> 3) * PANEL_BYTES;
#pragma unroll 1
        for (int rr0 = 0; rr0 < 16; rr0 += RU) {
          uint4 q[RU][3];
          float w[RU][3];
#pragma unroll
          for (int u = 0; u < RU; ++u) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const uint32_t v = __shfl_sync(0xffffffffu, myv[j], rr0 + u);
              w[u][j] = __shfl_sync(0xffffffffu, myb[j], rr0 + u);
              q[u][j] = ldg_nc_u4(p.proj + static_cast<size_t>(v) * PROJ_ROW_BYTES + proj_chunk_off(v, static_cast<uint32_t>(lane)));
            }
          }
          if (RY_LATE_FREE && rr0 == 0) {
            // the previous tile's last dz images must have left X / Y before the first store below; waiting here, with the
            // first P rows already requested, hides the tail of that stash store behind the loads' latency
            free_x();
            free_y();
          }
#pragma unroll
          for (int u = 0; u < RU; ++u) {
            float acc[8];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const uint32_t vv[4] = {q[u][j].x, q[u][j].y, q[u][j].z, q[u][j].w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 f = ry_unpack(vv[i]);
                if (j == 0) { acc[2 * i] = w[u][0] * f.x; acc[2 * i + 1] = w[u][0] * f.y; }
                else { acc[2 * i] = fmaf(w[u][j], f.x, acc[2 * i]); acc[2 * i + 1] = fmaf(w[u][j], f.y, acc[2 * i + 1]); }
              }
            }
            uint32_t o[4];
            if (z0_lane) {
#pragma unroll
              for (int i = 0; i < 4; ++i) o[i] = pack_half2_relu_sat(acc[2 * i] + b0[2 * i], acc[2 * i + 1] + b0[2 * i + 1]);
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i) o[i] = pack_half2_sat(acc[2 * i], acc[2 * i + 1]);
            }
            const uint32_t row = static_cast<uint32_t>(16 * ws + rr0 + u);
            st_shared_v4(dst_img + row * 128u + (((blk & 7u) ^ (row & 7u)) << 4), o[0], o[1], o[2], o[3]);
          }
        }
      }
      if (RY_META_PF) load_meta((pair + static_cast<int>(gridDim.x)) * kSlots + slot, nxv, nxb);   // next tile's metadata: in flight during the chain
      sync_epi();
      if (kTrain) {
        // ReLU mask of layer 0 for this thread's row, read back from the image the warps just wrote
#pragma unroll
        for (int cq = 0; cq < 2; ++cq) {
          uint32_t w[16];
          ry_load_words(ximg, r, 2 * ch + cq, w);
          mk[0][cq] = ry_relu_mask16(w);
        }
      }
      publish(true, false, true);   // h_0
      t_gather += now() - tg0;
      const long long tf0 = now();

      // ---- forward layers 1..nb-1: h_l = relu(z_l + b_l) -> image X -----------------------------------
      for (int l = 1; l < nb; ++l) {
        wait_done();
        free_x();
        const bool is_skip = (l == g.skip);
        uint32_t mbits[2];
#pragma unroll
        for (int cq = 0; cq < 2; ++cq) {
          const int cb = 2 * ch + cq;
          uint32_t v[32];
          tmem_ld32(tb + 32u * cb, v);
          tmem_ld_wait();
          const float4* b4 = reinterpret_cast<const float4*>(s_bias + l * HID + 32 * cb);
          uint32_t zw[16];
          if (is_skip) ry_load_words(yimg, r, cb, zw);   // Lx(h) + Ly(x): the interpolated zy of this row
          uint32_t w[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bb = b4[i];
            float a0 = __uint_as_float(v[4 * i]) + bb.x, a1 = __uint_as_float(v[4 * i + 1]) + bb.y;
            float a2 = __uint_as_float(v[4 * i + 2]) + bb.z, a3 = __uint_as_float(v[4 * i + 3]) + bb.w;
            if (is_skip) {
              const float2 z01 = ry_unpack(zw[2 * i]), z23 = ry_unpack(zw[2 * i + 1]);
              a0 += z01.x; a1 += z01.y; a2 += z23.x; a3 += z23.y;
            }
            w[2 * i] = pack_half2_relu_sat(a0, a1);
            w[2 * i + 1] = pack_half2_relu_sat(a2, a3);
          }
          if (kTrain) mbits[cq] = ry_relu_mask16(w);
          ry_store_words(ximg, r, cb, w);
        }
        if (kTrain) { mk[l][0] = mbits[0]; mk[l][1] = mbits[1]; }
        const bool fourier = view && l == nb - 1;
        if (fourier) {
          // Fourier features of this ray's view direction (model.py:175-186) -> [128 x 64] panel in Y, whose last reader
          // was the skip layer's epilogue (an earlier iteration: skip < nb - 1 with the view branch)
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
          const uint32_t row = yimg + static_cast<uint32_t>(r) * 128u;
          for (int ck = 4 * ch; ck < 4 * ch + 4; ++ck) {      // this thread's half of the row's eight 16-byte chunks
            float f[8];
            if (8 * ck < g.edim) {
#pragma unroll
              for (int i = 0; i < 8; ++i) f[i] = ry_fourier(8 * ck + i, x0, x1, x2, g.vdim, g.emb, g.include);
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) f[i] = 0.f;
            }
            st_shared_v4(row + ((static_cast<uint32_t>(ck) ^ (static_cast<uint32_t>(r) & 7u)) << 4), pack_half2(f[0], f[1]),
                         pack_half2(f[2], f[3]), pack_half2(f[4], f[5]), pack_half2(f[6], f[7]));
          }
        }
        sync_epi();
        publish(true, fourier, true);   // h_l (+ the Fourier panel)
      }

      t_fwd += now() - tf0;
      // ---- layer nb epilogue + output layer (-> 3, sigmoid) on CUDA cores --------------------------------
      {
        const long long tl0 = now();
        wait_done();
        free_x();
        const bool is_skip = (nb == g.skip);
        float y[3] = {0.f, 0.f, 0.f};       // this thread's 64 columns of the output layer's dot products
        uint32_t mbits[2];
#pragma unroll
        for (int cq = 0; cq < 2; ++cq) {
          const int cb = 2 * ch + cq;
          uint32_t v[32];
          tmem_ld32(tb + 32u * cb, v);
          tmem_ld_wait();
          const float4* b4 = reinterpret_cast<const float4*>(s_bias + nb * HID + 32 * cb);
          const float4* w0 = reinterpret_cast<const float4*>(s_wlast + 0 * HID + 32 * cb);
          const float4* w1 = reinterpret_cast<const float4*>(s_wlast + 1 * HID + 32 * cb);
          const float4* w2 = reinterpret_cast<const float4*>(s_wlast + 2 * HID + 32 * cb);
          uint32_t zw[16];
          if (is_skip) ry_load_words(yimg, r, cb, zw);
          uint32_t w[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bb = b4[i], a0 = w0[i], a1 = w1[i], a2 = w2[i];
            float h0 = __uint_as_float(v[4 * i]) + bb.x, h1 = __uint_as_float(v[4 * i + 1]) + bb.y;
            float h2 = __uint_as_float(v[4 * i + 2]) + bb.z, h3 = __uint_as_float(v[4 * i + 3]) + bb.w;
            if (is_skip) {
              const float2 z01 = ry_unpack(zw[2 * i]), z23 = ry_unpack(zw[2 * i + 1]);
              h0 += z01.x; h1 += z01.y; h2 += z23.x; h3 += z23.y;
            }
            h0 = fmaxf(h0, 0.f); h1 = fmaxf(h1, 0.f); h2 = fmaxf(h2, 0.f); h3 = fmaxf(h3, 0.f);
            y[0] = fmaf(a0.w, h3, fmaf(a0.z, h2, fmaf(a0.y, h1, fmaf(a0.x, h0, y[0]))));
            y[1] = fmaf(a1.w, h3, fmaf(a1.z, h2, fmaf(a1.y, h1, fmaf(a1.x, h0, y[1]))));
            y[2] = fmaf(a2.w, h3, fmaf(a2.z, h2, fmaf(a2.y, h1, fmaf(a2.x, h0, y[2]))));
            if (kTrain) { w[2 * i] = pack_half2_sat(h0, h1); w[2 * i + 1] = pack_half2_sat(h2, h3); }
          }
          if (kTrain) {
            mbits[cq] = ry_relu_mask16(w);
            ry_store_words(ximg, r, cb, w);
          }
        }
        // the two column halves of a row meet in shared memory; both then hold the same y (same summation order)
#pragma unroll
        for (int c = 0; c < 3; ++c) s_ysum[(ch * TILE_M + r) * OUT_DIM + c] = y[c];
        named_bar_sync(bar_id, RY_WG_THREADS);
#pragma unroll
        for (int c = 0; c < 3; ++c)
          y[c] = ry_sigmoid((s_blast[c] + s_ysum[r * OUT_DIM + c]) + s_ysum[(TILE_M + r) * OUT_DIM + c]);
        if (valid && ch == 0) {
          p.pred[ray * 3 + 0] = y[0];
          p.pred[ray * 3 + 1] = y[1];
          p.pred[ray * 3 + 2] = y[2];
        }
        if (kTrain) {
          float dzl[3] = {0.f, 0.f, 0.f};
          if (valid) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              float gy;
              if (p.target != nullptr) {
                const float d = y[c] - (RY_TGT_PF ? tgt[c] : p.target[src * 3 + c]);
                if (ch == 0) {      // one thread per ray counts the loss
                  loss_abs += fabsf(d);
                  loss_sq += d * d;
                }
                gy = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);  // d|d|/dd, sign(0) = 0 as F.l1_loss
              } else {
                gy = p.grad_pred[ray * 3 + c] * p.ext_scale;
              }
              dzl[c] = gy * y[c] * (1.f - y[c]) * DZ_SCALE;
              if (ch == 0) gb[c] += dzl[c];
            }
          }
          // Y (zy / Fourier panel; the previous tile's dz_nb and dzl^T left with the same store) is free for dz_nb
          free_y();
          // dzl^T tile [16 x 128 rays] (rows 3..15 stay zero): B operand of the output-layer weight gradient
          if (ch == 0) {
#pragma unroll
            for (int c = 0; c < 3; ++c)
              *reinterpret_cast<__half*>(smem + dzlt + static_cast<uint32_t>(r >> 6) * 2048u + sw128_off(c, r & 63)) = __float2half_rn(dzl[c]);
          }
          // dz_nb = (dz_last * W_last) masked by relu' -> Y, while h_nb (X) goes to the stash
#pragma unroll
          for (int cq = 0; cq < 2; ++cq) {
            const int cb = 2 * ch + cq;
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
              w[2 * i] = ry_apply_mask(pack_half2_sat(d0, d1), mbits[cq], 2 * i);
              w[2 * i + 1] = ry_apply_mask(pack_half2_sat(d2, d3), mbits[cq], 2 * i + 1);
            }
            ry_store_words(yimg, r, cb, w);
          }
          sync_epi();
          publish(true, true, true);   // h_nb (X), dz_nb + dzl^T (Y): one stash request; releases the first dgrad MMA
          t_final += now() - tl0;
          const long long tb0 = now();
          // ---- dgrad chain: dz_l = dh_l * relu'(z_l), l = nb-1..0, alternating X / Y ------------------------
          for (int i = 0; i < nb; ++i) {
            const int lb = nb - 1 - i;
            const bool to_x = (i & 1) == 0;
            wait_done();
            if (to_x) free_x(); else free_y();
            const uint32_t dimg = to_x ? ximg : yimg;
            const uint32_t mm[2] = {mk[lb][0], mk[lb][1]};
#pragma unroll
            for (int cq = 0; cq < 2; ++cq) {
              const int cb = 2 * ch + cq;
              uint32_t v[32];
              tmem_ld32(tb + 32u * cb, v);
              tmem_ld_wait();
              uint32_t w[16];
#pragma unroll
              for (int q = 0; q < 16; ++q)
                w[q] = ry_apply_mask(pack_half2_sat(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1])), mm[cq], q);
              ry_store_words(dimg, r, cb, w);
            }
            sync_epi();
            publish(to_x, !to_x, lb > 0);
          }
          t_bwd += now() - tb0;
        } else {
          t_final += now() - tl0;
        }
      }
    }
    if (leader) {
      const int o = 8 + 12 * slot;
      put(o, now() - t_epi0); put(o + 1, t_done); put(o + 2, t_free); put(o + 3, t_sync); put(o + 4, t_gather); put(o + 5, t_fwd);
      put(o + 6, t_final); put(o + 7, t_bwd);
    }
    if (kTrain) {
      // ---- per-warpgroup partials: output-layer bias gradient, loss sums ------------------------------
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        loss_abs += __shfl_xor_sync(0xffffffffu, loss_abs, o);
        loss_sq += __shfl_xor_sync(0xffffffffu, loss_sq, o);
#pragma unroll
        for (int c = 0; c < 3; ++c) gb[c] += __shfl_xor_sync(0xffffffffu, gb[c], o);
      }
      float* s_red = reinterpret_cast<float*>(smem + sm.red) + slot * 24;   // [5][4]: per-warp partial sums
      if (lane == 0 && ch == 0) {      // the column-half-0 warps carry the per-ray sums
        s_red[wq] = loss_abs; s_red[4 + wq] = loss_sq;
        s_red[8 + wq] = gb[0]; s_red[12 + wq] = gb[1]; s_red[16 + wq] = gb[2];
      }
      named_bar_sync(bar_id, RY_WG_THREADS);
      if (leader) {
        float* part = p.part1 + static_cast<size_t>(blockIdx.x * kSlots + slot) * K1_PART_FLOATS;
        part[0] = s_red[8] + s_red[9] + s_red[10] + s_red[11];
        part[1] = s_red[12] + s_red[13] + s_red[14] + s_red[15];
        part[2] = s_red[16] + s_red[17] + s_red[18] + s_red[19];
        part[3] = 0.f;
        part[4] = s_red[0] + s_red[1] + s_red[2] + s_red[3];
        part[5] = s_red[4] + s_red[5] + s_red[6] + s_red[7];
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == RY_W_CH) {
    tc_fence_after();
    tmem_dealloc(tmem_base, RY_TMEM_COLS);
  }
}

template <bool kTrain, int kSlots>
static cudaError_t launch_rays_one(const FieldParams& p, int grid, cudaStream_t stream) {
  const int smem = static_cast<int>(ry_smem(kTrain, kSlots).total);
  cudaError_t e = cudaFuncSetAttribute(field_rays_kernel<kTrain, kSlots>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(ry_threads(kSlots));
  cfg.dynamicSmemBytes = static_cast<size_t>(smem);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  add_pdl_attr(cfg, attr);
  return cudaLaunchKernelEx(&cfg, field_rays_kernel<kTrain, kSlots>, p);
}

// slots = 2: one CTA per SM, two tiles in lockstep; slots = 1: two CTAs per SM, one tile each
cudaError_t launch_field_rays(const FieldParams& p, bool train, int slots, int grid, cudaStream_t stream) {
  if (slots == 1) return train ? launch_rays_one<true, 1>(p, grid, stream) : launch_rays_one<false, 1>(p, grid, stream);
  return train ? launch_rays_one<true, 2>(p, grid, stream) : launch_rays_one<false, 2>(p, grid, stream);
}

}  // namespace inf
