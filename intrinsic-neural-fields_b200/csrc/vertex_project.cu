// Vertex projection, sm_100a:  P = E [W0;Wy]^T  -- layer 0 (model.py:98-104) and the skip branch Ly (layers.py:60-62)
// evaluated once per VERTEX instead of once per ray.
//
// Both are linear in the table rows, and the per-ray input is x = sum_j bary_j E[v_j] (mesh.py:313-324), so
//     x W^T = sum_j bary_j (E[v_j] W^T) = sum_j bary_j P[v_j].
// The K = k_pad contraction therefore runs over V rows that stream through the SM exactly once per step (a dense
// [V x k_pad] x [k_pad x 256] GEMM: every A panel is a linear 16 KB bulk copy of the table's fp16 operand image, no
// per-thread loads, no gather), and the per-ray kernel (field_rays.cu) gathers three 512-byte rows of P instead of
// three 4 KB table rows.  The vertices a batch touches at least once each, so for V <= ~2 B this is strictly less
// first-layer work and 8x less gathered traffic than the per-ray path of field_main.cu (kept for V >> B).
//
// Roles of a CTA (persistent, one per SM, clusters of 2 sharing the [W0;Wy] stream by multicast):
//   warps 0..3  epilogue: TMEM accumulator [128 vertices x 256] -> fp16 -> row-major staging tile in shared memory
//               (16-byte chunks XOR-permuted per row, proj_chunk_off) -> ONE 64 KB linear bulk store into P
//   warp 4      A loader: 16 KB panels of the table image (read-once stream, L2 evict_first)
//   warp 5      B loader: 32 KB [W0;Wy] chunks, each fetched once per cluster and multicast
//   warp 6      MMA issuer: tcgen05.mma kind::f16 M=128 N=256 K=16, two 256-column TMEM buffers (the epilogue of
//               tile t overlaps the MMAs of tile t+1)
#include "common.cuh"
#include "ptx.cuh"

namespace inf {

namespace {

constexpr int VP_EPI_WARPS = 4;
constexpr int VP_W_A = 4, VP_W_B = 5, VP_W_MMA = 6;
constexpr int VP_THREADS = 7 * 32;
constexpr int VP_EPI_THREADS = VP_EPI_WARPS * 32;
// The loaders are bound by bytes in flight: at the full MMA rate a k-chunk lasts 0.26 us, a table panel comes from DRAM (~2 us)
// and a weight chunk from L2 (~1 us), i.e. ~8 panels and ~4 chunks would have to be in flight; with 4 + 2 stages the kernel sat
// at 40 % of the tensor pipe.  The staging tile of the epilogue therefore holds HALF a vertex tile (the two warp pairs take
// turns), which pays for 6 panel stages and 3 chunk stages.
#ifndef VP_STAGES_A
#define VP_STAGES_A 6
#endif
#ifndef VP_STAGES_B
#define VP_STAGES_B 3
#endif
constexpr int VP_NSTAGE_A = VP_STAGES_A;   // 16 KB panels
constexpr int VP_NSTAGE_B = VP_STAGES_B;   // 32 KB chunks
constexpr uint32_t VP_HALF_TILE_BYTES = PROJ_TILE_BYTES / 2;   // 64 vertex rows

constexpr uint32_t VP_SM_A = 0;
constexpr uint32_t VP_SM_B = VP_SM_A + VP_NSTAGE_A * PANEL_BYTES;
constexpr uint32_t VP_SM_OUT = VP_SM_B + VP_NSTAGE_B * BLOB_BYTES;   // [64 rows][512 B] staging of half a P tile
constexpr uint32_t VP_SM_BARS = VP_SM_OUT + VP_HALF_TILE_BYTES;
constexpr uint32_t VP_SM_TMEMPTR = VP_SM_BARS + 32 * 8;
constexpr uint32_t VP_SM_TOTAL = VP_SM_TMEMPTR + 16;

constexpr int VB_A_FULL = 0;
constexpr int VB_A_EMPTY = VB_A_FULL + VP_NSTAGE_A;
constexpr int VB_B_FULL = VB_A_EMPTY + VP_NSTAGE_A;
constexpr int VB_B_EMPTY = VB_B_FULL + VP_NSTAGE_B;
constexpr int VB_G_FULL = VB_B_EMPTY + VP_NSTAGE_B;   // [2]
constexpr int VB_G_EMPTY = VB_G_FULL + 2;             // [2]
constexpr int VB_S_FREE = VB_G_EMPTY + 2;             // [2] the staging tile is free for warp pair 0 / 1
static_assert(VB_S_FREE + 2 <= 32, "barrier table overflow");
static_assert(VP_SM_TOTAL <= 232448, "shared memory budget");

constexpr uint32_t VP_IDESC_N256 = umma_idesc_f16(128, 256, 0, 0);
constexpr uint32_t VP_DESC_HI = umma_desc_hi(1024);

}  // namespace

__global__ void __launch_bounds__(VP_THREADS, 1) vertex_project_kernel(const FieldParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const Geo g = p.geo;
  const int cs = p.cluster;
  const uint32_t crank = cs > 1 ? cluster_ctarank() : 0u;
  const uint16_t cmask = static_cast<uint16_t>((1u << cs) - 1u);
  auto bar = [&](int i) -> uint32_t { return smem_base + VP_SM_BARS + 8u * static_cast<uint32_t>(i); };

  if ((smem_base & 1023u) != 0) {
    if (tid == 0 && p.err) atomicExch(p.err, 0xBAD00003u);
    return;
  }
  if (tid == 0) {
    for (int i = 0; i < VP_NSTAGE_A; ++i) { mbar_init(bar(VB_A_FULL + i), 1); mbar_init(bar(VB_A_EMPTY + i), 1); }
    for (int i = 0; i < VP_NSTAGE_B; ++i) { mbar_init(bar(VB_B_FULL + i), 1); mbar_init(bar(VB_B_EMPTY + i), cs); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar(VB_G_FULL + i), 1); mbar_init(bar(VB_G_EMPTY + i), VP_EPI_WARPS); mbar_init(bar(VB_S_FREE + i), 1); }
    mbar_fence_init();
  }
  if (warp == VP_W_MMA) {
    tmem_alloc(smem_base + VP_SM_TMEMPTR, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (cs > 1) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + VP_SM_TMEMPTR);
  if (p.debug & DEBUG_PDL_EARLY) pdl_launch_dependents();
  pdl_wait();   // the operand images (previous step's Adam refresh) and P's previous readers

  // Vertex tiles are dealt round-robin; the CTAs of a cluster run the same number of iterations (a CTA past the end
  // recomputes the last tile and stores nothing) because they release each other's [W0;Wy] stages.
  const int ntv = p.ntiles_v;
  int nt_loop = ntv;
  if (cs > 1) {
    const int first = (static_cast<int>(blockIdx.x) / cs) * cs;
    const int iters = first < ntv ? (ntv - first + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x) : 0;
    nt_loop = iters > 0 ? static_cast<int>(blockIdx.x) + (iters - 1) * static_cast<int>(gridDim.x) + 1 : 0;
  }
  const int nchunks = g.nchunks;

  if (warp == VP_W_A) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint64_t stream_policy = l2_policy_evict_first();   // the image is read once per launch; keep the weights in L2
      for (int tile = blockIdx.x; tile < nt_loop; tile += gridDim.x) {
        const int t = tile < ntv ? tile : ntv - 1;
        const uint8_t* src = p.e_image + static_cast<size_t>(t) * nchunks * PANEL_BYTES;
        for (int c = 0; c < nchunks; ++c) {
          mbar_wait(bar(VB_A_EMPTY + stage), phase ^ 1u, p.err, 0xA0);
          mbar_arrive_expect_tx(bar(VB_A_FULL + stage), PANEL_BYTES);
          bulk_g2s_hint(smem_base + VP_SM_A + static_cast<uint32_t>(stage) * PANEL_BYTES, src + static_cast<size_t>(c) * PANEL_BYTES,
                        PANEL_BYTES, bar(VB_A_FULL + stage), stream_policy);
          if (++stage == VP_NSTAGE_A) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == VP_W_B) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t slice = BLOB_BYTES / static_cast<uint32_t>(cs);
      for (int tile = blockIdx.x; tile < nt_loop; tile += gridDim.x) {
        for (int c = 0; c < nchunks; ++c) {
          mbar_wait(bar(VB_B_EMPTY + stage), phase ^ 1u, p.err, 0xA1);   // released by every CTA of the cluster
          mbar_arrive_expect_tx(bar(VB_B_FULL + stage), BLOB_BYTES);
          const uint32_t dst = smem_base + VP_SM_B + static_cast<uint32_t>(stage) * BLOB_BYTES;
          const uint8_t* src = p.packed + static_cast<size_t>(c) * BLOB_BYTES;
          if (cs == 1) bulk_g2s(dst, src, BLOB_BYTES, bar(VB_B_FULL + stage));
          else bulk_g2s_multicast(dst + crank * slice, src + crank * slice, slice, bar(VB_B_FULL + stage), cmask);
          if (++stage == VP_NSTAGE_B) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == VP_W_MMA) {
    if (lane == 0) {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      uint32_t gph[2] = {0u, 0u};
      int it = 0;
      // measurement only (INF_DEBUG & 8): where the issuer of CTA 0 waits (scripts/trace_project.py)
      const bool tracing = p.trace != nullptr && blockIdx.x == 0;
      long long t_a = 0, t_b = 0, t_g = 0;
      const long long t_0 = tracing ? clock64() : 0;
      for (int tile = blockIdx.x; tile < nt_loop; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const long long tg0 = tracing ? clock64() : 0;
        mbar_wait(bar(VB_G_EMPTY + buf), gph[buf] ^ 1u, p.err, 0xA2);   // epilogue of tile it-2 has drained this buffer
        if (tracing) t_g += clock64() - tg0;
        gph[buf] ^= 1u;
        tc_fence_after();
        const uint32_t dcol = tmem_base + 256u * static_cast<uint32_t>(buf);
        for (int c = 0; c < nchunks; ++c) {
          const long long ta0 = tracing ? clock64() : 0;
          mbar_wait(bar(VB_A_FULL + sa), pa, p.err, 0xA3);
          const long long tb0 = tracing ? clock64() : 0;
          mbar_wait(bar(VB_B_FULL + sb), pb, p.err, 0xA4);
          if (tracing) { t_a += tb0 - ta0; t_b += clock64() - tb0; }
          tc_fence_after();
          const uint32_t alo = umma_desc_lo(smem_base + VP_SM_A + static_cast<uint32_t>(sa) * PANEL_BYTES, 0);
          const uint32_t blo = umma_desc_lo(smem_base + VP_SM_B + static_cast<uint32_t>(sb) * BLOB_BYTES, 0);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_f16(dcol, umma_desc_join(alo + 2u * kk, VP_DESC_HI), umma_desc_join(blo + 2u * kk, VP_DESC_HI), VP_IDESC_N256,
                     (c > 0 || kk > 0) ? 1u : 0u);
          if (cs == 1) umma_commit(bar(VB_B_EMPTY + sb));
          else umma_commit_multicast(bar(VB_B_EMPTY + sb), cmask);
          umma_commit(bar(VB_A_EMPTY + sa));
          if (++sa == VP_NSTAGE_A) { sa = 0; pa ^= 1u; }
          if (++sb == VP_NSTAGE_B) { sb = 0; pb ^= 1u; }
        }
        umma_commit(bar(VB_G_FULL + buf));
      }
      if (tracing) {
        p.trace[50] = static_cast<unsigned long long>(clock64() - t_0);
        p.trace[51] = static_cast<unsigned long long>(t_a);
        p.trace[52] = static_cast<unsigned long long>(t_b);
        p.trace[53] = static_cast<unsigned long long>(t_g);
        p.trace[54] = static_cast<unsigned long long>(it);
      }
    }
  } else {
    // =========================== epilogue warps: thread = vertex row = TMEM lane =====================
    // Warp pair 0 (rows 0..63) and pair 1 (rows 64..127) take turns on the 32 KB staging tile: a pair converts its rows,
    // its first thread bulk-stores the half tile, waits until the copy engine has read it and hands the tile to the other pair.
    const int r = tid;
    const int pair = warp >> 1;
    const bool pair_lead = (tid & 63) == 0;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const uint32_t out_row = smem_base + VP_SM_OUT + static_cast<uint32_t>(r & 63) * PROJ_ROW_BYTES;
    uint32_t gph[2] = {0u, 0u};
    uint32_t sph = pair == 0 ? 1u : 0u;   // pair 0 finds the tile free at the start (parity trick of a fresh barrier)
    int it = 0;
    for (int tile = blockIdx.x; tile < nt_loop; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(bar(VB_G_FULL + buf), gph[buf], p.err, 0xA5);
      gph[buf] ^= 1u;
      tc_fence_after();
      mbar_wait(bar(VB_S_FREE + pair), sph, p.err, 0xA6);   // the other pair's store has finished READING the staging tile
      sph ^= 1u;
      const uint32_t tb = tlane + 256u * static_cast<uint32_t>(buf);
#pragma unroll 2
      for (int cb = 0; cb < 8; ++cb) {
        uint32_t v[32];
        tmem_ld32(tb + 32u * cb, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t q = static_cast<uint32_t>(4 * cb + i);
          st_shared_v4(out_row + proj_chunk_off(static_cast<uint32_t>(r), q),
                       pack_half2_sat(__uint_as_float(v[8 * i + 0]), __uint_as_float(v[8 * i + 1])),
                       pack_half2_sat(__uint_as_float(v[8 * i + 2]), __uint_as_float(v[8 * i + 3])),
                       pack_half2_sat(__uint_as_float(v[8 * i + 4]), __uint_as_float(v[8 * i + 5])),
                       pack_half2_sat(__uint_as_float(v[8 * i + 6]), __uint_as_float(v[8 * i + 7])));
        }
      }
      // all TMEM reads of this buffer by this warp are complete: hand it back to the MMA issuer (4 arrivals)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(VB_G_EMPTY + buf));
      fence_proxy_async_smem();
      named_bar_sync(1 + static_cast<uint32_t>(pair), 64);
      if (pair_lead) {
        if (tile < ntv) {
          // P stays L2-resident for the per-ray kernel that follows: plain store, no eviction hint
          bulk_s2g_hint(p.proj + static_cast<size_t>(tile) * PROJ_TILE_BYTES + static_cast<size_t>(pair) * VP_HALF_TILE_BYTES,
                        smem_base + VP_SM_OUT, VP_HALF_TILE_BYTES, 0ull);
          bulk_commit();
          bulk_wait_read<0>();
        }
        mbar_arrive(bar(VB_S_FREE + (pair ^ 1)));
      }
    }
    if (pair_lead) bulk_wait<0>();
    tc_fence_before();
  }

  __syncthreads();
  if (cs > 1) cluster_sync_all();
  if (warp == VP_W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// The same projection on CTA PAIRS (cta_group::2).  With one CTA per MMA the shared memory of an SM moves, per 64-wide
// k-chunk, 48 KB in (table panel 16 KB + the whole [W0;Wy] chunk 32 KB) and 48 KB out to the tensor core: 192 bytes per
// clock at the full MMA rate against the 128 the SM has -- the kernel sat at 40 % of the tensor pipe whatever the ring depth.
// A pair computes two vertex tiles with ONE M = 256 MMA: each CTA holds its own tile's A panel and only ITS HALF of the
// weight chunk (N/2 = 128 rows), so per CTA 32 KB go in and 32 KB out per chunk.
//   warps 0..3  epilogue of this CTA's tile (as above)
//   warp 4      loader: one 32 KB stage per k-chunk = [own table panel 16 KB][own half of the weight chunk 16 KB]
//   warp 5      relay (peer CTA only): tells the leader that the peer's stage has landed
//   warp 6      TMEM allocation (both CTAs); MMA issuer (leader only)
// ---------------------------------------------------------------------------------------------
namespace {
constexpr int V2_NSTAGE = 6;
constexpr uint32_t V2_STAGE_BYTES = PANEL_BYTES + BLOB_BYTES / 2;   // 32 KB
constexpr uint32_t V2_SM_STAGES = 0;
constexpr uint32_t V2_SM_OUT = V2_SM_STAGES + V2_NSTAGE * V2_STAGE_BYTES;
constexpr uint32_t V2_SM_BARS = V2_SM_OUT + VP_HALF_TILE_BYTES;
constexpr uint32_t V2_SM_TMEMPTR = V2_SM_BARS + 32 * 8;
constexpr uint32_t V2_SM_TOTAL = V2_SM_TMEMPTR + 16;
static_assert(V2_SM_TOTAL <= 232448, "shared memory budget");
constexpr int V2_FULL = 0;                       // [stages] this CTA's stage has landed
constexpr int V2_EMPTY = V2_FULL + V2_NSTAGE;    // [stages] the pair MMA has read the stage (both CTAs, multicast commit)
constexpr int V2_PFULL = V2_EMPTY + V2_NSTAGE;   // [stages] leader: the PEER's stage has landed (remote arrive by its relay)
constexpr int V2_G_FULL = V2_PFULL + V2_NSTAGE;  // [2] accumulator complete (both CTAs, multicast commit)
constexpr int V2_G_EMPTY = V2_G_FULL + 2;        // [2] leader: both CTAs' epilogues have drained the buffer (4 + 4 warps)
constexpr int V2_S_FREE = V2_G_EMPTY + 2;        // [2] staging tile free for warp pair 0 / 1
static_assert(V2_S_FREE + 2 <= 32, "barrier table overflow");
constexpr uint32_t V2_IDESC = umma_idesc_f16(256, 256, 0, 0);
constexpr int V2_W_LD = 4, V2_W_RELAY = 5, V2_W_MMA = 6;
}  // namespace

__global__ void __launch_bounds__(VP_THREADS, 1) vertex_project2_kernel(const FieldParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const Geo g = p.geo;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  auto bar = [&](int i) -> uint32_t { return smem_base + V2_SM_BARS + 8u * static_cast<uint32_t>(i); };

  if ((smem_base & 1023u) != 0) {
    if (tid == 0 && p.err) atomicExch(p.err, 0xBAD00005u);
    return;
  }
  if (tid == 0) {
    for (int i = 0; i < V2_NSTAGE; ++i) { mbar_init(bar(V2_FULL + i), 1); mbar_init(bar(V2_EMPTY + i), 1); mbar_init(bar(V2_PFULL + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar(V2_G_FULL + i), 1); mbar_init(bar(V2_G_EMPTY + i), 2 * VP_EPI_WARPS); mbar_init(bar(V2_S_FREE + i), 1); }
    mbar_fence_init();
  }
  if (warp == V2_W_MMA) {
    tmem_alloc2(smem_base + V2_SM_TMEMPTR, 512);
    tmem_relinquish2();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs' barriers exist before anything arrives on them remotely
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + V2_SM_TMEMPTR);
  if (p.debug & DEBUG_PDL_EARLY) pdl_launch_dependents();
  pdl_wait();   // the operand images (previous step's Adam refresh) and P's previous readers

  // Tile pairs are dealt round-robin to the clusters; this CTA takes tile 2 * pair + rank (past the end: the last tile
  // again, nothing stored, so that both CTAs of a pair always run the same chunks).
  const int ntv = p.ntiles_v;
  const int npairs = (ntv + 1) / 2;
  const int cluster_id = static_cast<int>(blockIdx.x) >> 1, nclusters = static_cast<int>(gridDim.x) >> 1;
  const int nchunks = g.nchunks;

  if (warp == V2_W_LD) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint64_t stream_policy = l2_policy_evict_first();   // the image is read once per launch; keep the weights in L2
      for (int tp = cluster_id; tp < npairs; tp += nclusters) {
        const int tile = 2 * tp + static_cast<int>(crank);
        const int t = tile < ntv ? tile : ntv - 1;
        const uint8_t* src = p.e_image + static_cast<size_t>(t) * nchunks * PANEL_BYTES;
        for (int c = 0; c < nchunks; ++c) {
          mbar_wait_cluster(bar(V2_EMPTY + stage), phase ^ 1u, p.err, 0xE0);
          mbar_arrive_expect_tx(bar(V2_FULL + stage), V2_STAGE_BYTES);
          const uint32_t dst = smem_base + V2_SM_STAGES + static_cast<uint32_t>(stage) * V2_STAGE_BYTES;
          bulk_g2s_hint(dst, src + static_cast<size_t>(c) * PANEL_BYTES, PANEL_BYTES, bar(V2_FULL + stage), stream_policy);
          bulk_g2s(dst + PANEL_BYTES, p.packed + static_cast<size_t>(c) * BLOB_BYTES + crank * (BLOB_BYTES / 2), BLOB_BYTES / 2,
                   bar(V2_FULL + stage));
          if (++stage == V2_NSTAGE) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == V2_W_RELAY) {
    if (lane == 0 && !leader) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tp = cluster_id; tp < npairs; tp += nclusters) {
        for (int c = 0; c < nchunks; ++c) {
          mbar_wait(bar(V2_FULL + stage), phase, p.err, 0xE1);
          mbar_arrive_cluster(mapa_cluster(bar(V2_PFULL + stage), 0u));
          if (++stage == V2_NSTAGE) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == V2_W_MMA) {
    if (lane == 0 && leader) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t gph[2] = {0u, 0u};
      int it = 0;
      for (int tp = cluster_id; tp < npairs; tp += nclusters, ++it) {
        const int buf = it & 1;
        mbar_wait_cluster(bar(V2_G_EMPTY + buf), gph[buf] ^ 1u, p.err, 0xE2);   // both epilogues of pair it-2 have drained this buffer
        gph[buf] ^= 1u;
        tc_fence_after();
        const uint32_t dcol = tmem_base + 256u * static_cast<uint32_t>(buf);
        for (int c = 0; c < nchunks; ++c) {
          mbar_wait(bar(V2_FULL + stage), phase, p.err, 0xE3);
          mbar_wait_cluster(bar(V2_PFULL + stage), phase, p.err, 0xE4);
          tc_fence_after();
          const uint32_t sbase = smem_base + V2_SM_STAGES + static_cast<uint32_t>(stage) * V2_STAGE_BYTES;
          const uint32_t alo = umma_desc_lo(sbase, 0), blo = umma_desc_lo(sbase + PANEL_BYTES, 0);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma2_f16(dcol, umma_desc_join(alo + 2u * kk, VP_DESC_HI), umma_desc_join(blo + 2u * kk, VP_DESC_HI), V2_IDESC,
                      (c > 0 || kk > 0) ? 1u : 0u);
          umma2_commit_multicast(bar(V2_EMPTY + stage), 3);
          if (++stage == V2_NSTAGE) { stage = 0; phase ^= 1u; }
        }
        umma2_commit_multicast(bar(V2_G_FULL + buf), 3);
      }
    }
  } else if (warp < VP_EPI_WARPS) {
    // =========================== epilogue warps: thread = vertex row = TMEM lane of THIS CTA =====================
    const int r = tid;
    const int pair = warp >> 1;
    const bool pair_lead = (tid & 63) == 0;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const uint32_t out_row = smem_base + V2_SM_OUT + static_cast<uint32_t>(r & 63) * PROJ_ROW_BYTES;
    uint32_t gph[2] = {0u, 0u};
    uint32_t sph = pair == 0 ? 1u : 0u;
    uint32_t gempty[2];
    gempty[0] = mapa_cluster(bar(V2_G_EMPTY + 0), 0u);
    gempty[1] = mapa_cluster(bar(V2_G_EMPTY + 1), 0u);
    int it = 0;
    for (int tp = cluster_id; tp < npairs; tp += nclusters, ++it) {
      const int tile = 2 * tp + static_cast<int>(crank);
      const int buf = it & 1;
      mbar_wait_cluster(bar(V2_G_FULL + buf), gph[buf], p.err, 0xE5);
      gph[buf] ^= 1u;
      tc_fence_after();
      mbar_wait(bar(V2_S_FREE + pair), sph, p.err, 0xE6);
      sph ^= 1u;
      const uint32_t tb = tlane + 256u * static_cast<uint32_t>(buf);
#pragma unroll 2
      for (int cb = 0; cb < 8; ++cb) {
        uint32_t v[32];
        tmem_ld32(tb + 32u * cb, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t q = static_cast<uint32_t>(4 * cb + i);
          st_shared_v4(out_row + proj_chunk_off(static_cast<uint32_t>(r), q),
                       pack_half2_sat(__uint_as_float(v[8 * i + 0]), __uint_as_float(v[8 * i + 1])),
                       pack_half2_sat(__uint_as_float(v[8 * i + 2]), __uint_as_float(v[8 * i + 3])),
                       pack_half2_sat(__uint_as_float(v[8 * i + 4]), __uint_as_float(v[8 * i + 5])),
                       pack_half2_sat(__uint_as_float(v[8 * i + 6]), __uint_as_float(v[8 * i + 7])));
        }
      }
      // this warp's TMEM reads of the buffer are complete: tell the leader's MMA issuer (8 arrivals per buffer)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(gempty[buf]);
      fence_proxy_async_smem();
      named_bar_sync(1 + static_cast<uint32_t>(pair), 64);
      if (pair_lead) {
        if (tile < ntv) {
          bulk_s2g_hint(p.proj + static_cast<size_t>(tile) * PROJ_TILE_BYTES + static_cast<size_t>(pair) * VP_HALF_TILE_BYTES,
                        smem_base + V2_SM_OUT, VP_HALF_TILE_BYTES, 0ull);
          bulk_commit();
          bulk_wait_read<0>();
        }
        mbar_arrive(bar(V2_S_FREE + (pair ^ 1)));
      }
    }
    if (pair_lead) bulk_wait<0>();
    tc_fence_before();
  }

  __syncthreads();
  cluster_sync_all();   // no CTA of the pair leaves (or frees its tensor memory) while the other may still use it
  if (warp == V2_W_MMA) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

static void vp_launch_config(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int grid, int cluster, cudaStream_t stream) {
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(VP_THREADS);
  cfg.dynamicSmemBytes = VP_SM_TOTAL;
  cfg.stream = stream;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(cluster);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
}

int vertex_project_max_grid(int cluster) {
  cudaFuncSetAttribute(vertex_project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(VP_SM_TOTAL));
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  vp_launch_config(cfg, attr, cluster, cluster, nullptr);
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, vertex_project_kernel, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n * cluster;
}

cudaError_t launch_vertex_project(const FieldParams& p, int grid, int cluster, cudaStream_t stream) {
  // measured (scripts/trace_project.py, gpurun_out/t_v2.log): the pair kernel is correct (same tests) but takes 67 us against
  // 39 us -- the single-CTA main loop is only 23 of those 39 us (tensor pipe ~50 % busy inside it; the rest is launch,
  // first-panel latency and the last tile's epilogue + store, fixed costs of a kernel with 2.65 tiles per CTA), and the pair's
  // stage hand-off (peer relay -> leader, multicast commits back) costs more than the halved operand traffic returns.  Off.
  static const bool pair_mma = [] { const char* e = getenv("INF_VP_2CTA"); return e != nullptr && atoi(e) != 0; }();
  if (pair_mma && cluster == 2 && grid >= 2 && (grid & 1) == 0 && p.ntiles_v >= 2) {
    cudaError_t e2 = cudaFuncSetAttribute(vertex_project2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(V2_SM_TOTAL));
    if (e2 != cudaSuccess) return e2;
    FieldParams q = p;
    q.cluster = 2;
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[2];
    vp_launch_config(cfg, attr, grid, 2, stream);
    cfg.dynamicSmemBytes = V2_SM_TOTAL;
    add_pdl_attr(cfg, attr);
    return cudaLaunchKernelEx(&cfg, vertex_project2_kernel, q);
  }
  cudaError_t e = cudaFuncSetAttribute(vertex_project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(VP_SM_TOTAL));
  if (e != cudaSuccess) return e;
  FieldParams q = p;
  q.cluster = cluster;
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[2];
  vp_launch_config(cfg, attr, grid, cluster, stream);
  add_pdl_attr(cfg, attr);
  return cudaLaunchKernelEx(&cfg, vertex_project_kernel, q);
}

// ---------------------------------------------------------------------------------------------
// Table operand image: (V, k_pad) row-major f32|f16 -> [ceil(V/128)][k_pad/64] panels of [128 rows][64 fp16],
// SWIZZLE_128B.  Built once where the table goes to the device (ray_dataloader.py:73, renderer.py:43); the same bytes
// are the K-major A operand of the projection GEMM and the MN-major B operand of the first-layer weight gradient.
// One thread per 16-byte chunk of the image.
// ---------------------------------------------------------------------------------------------
template <typename TableT>
__global__ void table_image_kernel(const TableT* __restrict__ table, int64_t n_rows, int k_pad, uint8_t* __restrict__ image) {
  const int nchunks = k_pad / KC;
  const int64_t ntv = (n_rows + TILE_M - 1) / TILE_M;
  const int64_t total = ntv * nchunks * (PANEL_BYTES / 16);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int q = static_cast<int>(i & 7);                 // 16-byte chunk of the 128-byte panel row
    const int r = static_cast<int>((i >> 3) & 127);        // row of the tile
    const int64_t pc = i >> 10;                            // tile * nchunks + c
    const int c = static_cast<int>(pc % nchunks);
    const int64_t row = (pc / nchunks) * TILE_M + r;
    uint4 out = make_uint4(0u, 0u, 0u, 0u);
    if (row < n_rows) {
      const TableT* src = table + row * k_pad + c * KC + q * 8;
      if constexpr (sizeof(TableT) == 4) {
        const float4 a = *reinterpret_cast<const float4*>(src);
        const float4 b = *reinterpret_cast<const float4*>(src + 4);
        out = make_uint4(pack_half2_sat(a.x, a.y), pack_half2_sat(a.z, a.w), pack_half2_sat(b.x, b.y), pack_half2_sat(b.z, b.w));
      } else {
        out = *reinterpret_cast<const uint4*>(src);
      }
    }
    *reinterpret_cast<uint4*>(image + static_cast<size_t>(pc) * PANEL_BYTES + static_cast<size_t>(r) * 128 +
                              ((static_cast<uint32_t>(q) ^ (static_cast<uint32_t>(r) & 7u)) << 4)) = out;
  }
}

// P as a plain (V, 256) row-major fp16 matrix (inf_project_vertices): undo the per-row chunk permutation
__global__ void proj_unpermute_kernel(const uint8_t* __restrict__ proj, int64_t n_rows, uint8_t* __restrict__ out) {
  const int64_t total = n_rows * (PROJ_ROW_BYTES / 16);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t v = i >> 5;
    const uint32_t q = static_cast<uint32_t>(i & 31);
    *reinterpret_cast<uint4*>(out + v * PROJ_ROW_BYTES + q * 16) =
        *reinterpret_cast<const uint4*>(proj + v * PROJ_ROW_BYTES + proj_chunk_off(static_cast<uint32_t>(v), q));
  }
}
void launch_proj_unpermute(const uint8_t* proj, int64_t n_rows, void* out, cudaStream_t stream) {
  if (n_rows == 0) return;
  const int64_t want = (n_rows * 32 + 255) / 256;
  proj_unpermute_kernel<<<static_cast<unsigned>(want < 148 * 16 ? want : 148 * 16), 256, 0, stream>>>(proj, n_rows,
                                                                                                     reinterpret_cast<uint8_t*>(out));
}

void launch_table_image(const void* table, int table_f16, int64_t n_rows, int k_pad, void* image, cudaStream_t stream) {
  if (n_rows == 0) return;
  const int64_t ntv = (n_rows + TILE_M - 1) / TILE_M;
  const int64_t total = ntv * (k_pad / KC) * (PANEL_BYTES / 16);
  const int threads = 256;
  const int64_t want = (total + threads - 1) / threads;
  const unsigned blocks = static_cast<unsigned>(want < 148 * 16 ? want : 148 * 16);
  if (table_f16)
    table_image_kernel<__half><<<blocks, threads, 0, stream>>>(reinterpret_cast<const __half*>(table), n_rows, k_pad,
                                                               reinterpret_cast<uint8_t*>(image));
  else
    table_image_kernel<float><<<blocks, threads, 0, stream>>>(reinterpret_cast<const float*>(table), n_rows, k_pad,
                                                              reinterpret_cast<uint8_t*>(image));
}

}  // namespace inf
