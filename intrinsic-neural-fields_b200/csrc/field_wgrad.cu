// Weight-gradient kernel + deterministic partial reduction, sm_100a.
//
// dW_l = sum_rays dz_l[ray,:]^T * input_l[ray,:] is a GEMM whose contraction dimension is the ray
// index.  Each CTA owns a column slice of the gradient for the whole launch and keeps it in TMEM
// (512 fp32 columns x 128 lanes) while it walks its share of the ray tiles:
//   "big" slice s < nbig : 256 columns j of dW0 (TMEM [0,256)) and of dWy (TMEM [256,512)); the x operand is
//                          the fp16 panel image the forward kernel stashed (no second gather);
//   "hidden" slices      : the work units of make_wplan(): 128x128 hidden / view-branch weight gradients,
//                          every bias gradient (an extra N=16 MMA against an all-ones tile: free column sums
//                          on the tensor core) and the output-layer weight gradient (h_nb^T x dzl^T).
// Both operands are read MN-major (UMMA transposes on the fly) from the same swizzled images the
// forward/backward kernel stashed, so no transposed copy is ever materialised; everything moves
// HBM/L2 -> smem with 1-D bulk copies (TMA engine), there are no per-thread loads at all.
// Ray tiles are dealt round-robin to `groups` CTAs per slice; every CTA finally dumps its
// accumulators to a partial buffer and field_reduce_kernel sums the partials in a fixed order
// (deterministic -> the N-GPU == 1-GPU check can be bit-stable per rank).
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "ptx.cuh"

namespace inf {

namespace {

constexpr int WG_THREADS = 6 * 32;   // warps 0..3: accumulator dump, warp 4: MMA issuer, warp 5: loader
constexpr int WG_MMA_WARP = 4;
constexpr int WG_LOAD_WARP = 5;

constexpr int NSTAGE = 3;
constexpr uint32_t STAGE_BYTES = 2 * BLOB_BYTES;  // [A 32 KB][B 32 KB]
constexpr uint32_t SM_STAGES = 0;
constexpr uint32_t SM_ONES = SM_STAGES + NSTAGE * STAGE_BYTES;   // [16 x 64] fp16 ones, K-major
constexpr uint32_t SM_BARS = SM_ONES + 2048;
constexpr uint32_t SM_TMEMPTR = SM_BARS + 16 * 8;
constexpr uint32_t SM_TOTAL = SM_TMEMPTR + 16;

constexpr int BAR_FULL = 0;
constexpr int BAR_EMPTY = NSTAGE;
constexpr int BAR_DONE = 2 * NSTAGE;

}  // namespace

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void __launch_bounds__(WG_THREADS, 1) field_wgrad_kernel(const FieldParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const Geo g = p.geo;

  if ((smem_base & 1023u) != 0) {
    if (tid == 0 && p.err) atomicExch(p.err, 0xBAD00002u);
    return;
  }

  // CTA -> (slice, ray group): big slices first (groups_big CTAs each), then the hidden slices.  A launch may cover a
  // sub-range of that CTA space (cta_base): the vertex path launches the hidden slices while the vertex gradient that the
  // big slices need is still being gathered.
  const int cta = static_cast<int>(blockIdx.x) + p.cta_base;
  const int nbig_ctas = g.nbig * p.groups_big;
  const bool big = cta < nbig_ctas;
  int slice, grp, groups;
  if (big) {
    // group-major: the CTAs of one ray group sit next to each other, so a cluster = `cw` slices of the SAME tiles
    groups = p.groups_big;
    slice = cta % g.nbig;
    grp = cta / g.nbig;
  } else {
    int rel = cta - nbig_ctas;
    int hs = 0;
    while (hs < p.plan.nslice - 1 && rel >= p.groups_hid[hs]) { rel -= p.groups_hid[hs]; ++hs; }
    slice = hs;
    groups = p.groups_hid[hs];
    grp = rel;
  }
  const int ncols = big ? min(256, g.k_pad - 256 * slice) : 0;   // big: x columns of this slice
  // hidden: units [u0, u1) of the plan belong to this slice (units are ordered by slice)
  int u0 = 0, u1 = 0;
  if (!big) {
    while (u0 < p.plan.n && p.plan.u[u0].slice != slice) ++u0;
    u1 = u0;
    while (u1 < p.plan.n && p.plan.u[u1].slice == slice) ++u1;
  }

  auto bar = [&](int i) -> uint32_t { return smem_base + SM_BARS + 8u * static_cast<uint32_t>(i); };
  // Clusters of `cw` CTAs: the big slices of a cluster walk the same ray tiles and need the same dz_0 / dz_skip
  // half-panels, so each CTA fetches 1/cw of them and multicasts -- per-SM ingest (what bounds this kernel, ~42 GB/s
  // through the bulk-copy engine) drops from 64 KB to (32 + 32/cw) KB per 64-ray unit.  Hidden-slice CTAs share nothing.
  const int cw = p.wg_cluster;
  const uint32_t crank = cw > 1 ? cluster_ctarank() : 0u;
  const uint16_t cmask = static_cast<uint16_t>((1u << cw) - 1u);
  const bool mcast = big && cw > 1;

  if (tid == 0) {
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(bar(BAR_FULL + i), 1);
      mbar_init(bar(BAR_EMPTY + i), mcast ? cw : 1);   // a stage is refilled by every CTA of the cluster
    }
    mbar_init(bar(BAR_DONE), 1);
    mbar_fence_init();
  }
  if (warp == WG_MMA_WARP) {
    tmem_alloc(smem_base + SM_TMEMPTR, WG_COLS);
    tmem_relinquish();
  }
  {
    uint32_t* ones = reinterpret_cast<uint32_t*>(smem + SM_ONES);
    for (int i = tid; i < 512; i += WG_THREADS) ones[i] = 0x3C003C00u;  // fp16 1.0 x2
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (cw > 1) cluster_sync_all();   // peers' barriers exist before anyone multicasts into them
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + SM_TMEMPTR);
  if (p.debug & DEBUG_PDL_EARLY) pdl_launch_dependents();
  pdl_wait();   // the stash / vertex-gradient image of the kernels before; the partials' previous reader

  // Vertex-projected path: the big slices contract over VERTEX tiles -- A = the vertex-gradient image G (laid out as a
  // tile's dz_0 / dz_skip blobs), B = the table's operand image (laid out as a tile's x panels); [dW0;dWy] = G^T E.
  const bool vp = p.e_image != nullptr;
  const int ntiles = (big && vp) ? p.ntiles_v : p.ntiles;
  const int my_tiles = (ntiles > grp) ? (ntiles - grp + groups - 1) / groups : 0;
  const int units_per_tile = big ? 2 : (u1 - u0);
  const int nunits = my_tiles * units_per_tile;
  const size_t tile_stride = p.stash_stride;

  if (warp == WG_LOAD_WARP) {
    // ======================= stash loader ======================================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint64_t stash_policy = (p.debug & 128) ? 0ull : l2_policy_evict_first();   // read-once stream (INF_DEBUG & 128: no hint)
      // (L2 prefetches of upcoming units through the same bulk-copy engine, and cp.async copy warps sharing the ring,
      //  were both measured slower than this plain 3-stage stream -- see DESIGN.md section 4)
      for (int u = 0; u < nunits; ++u) {
        // newest tiles first: the forward/backward kernel wrote them last, so part of their stash is still in L2
        const int tile = ntiles - 1 - (grp + (u / units_per_tile) * groups);
        const uint8_t* st = p.stash + static_cast<size_t>(tile) * tile_stride;
        mbar_wait(bar(BAR_EMPTY + stage), phase ^ 1u, p.err, 0x60);
        const uint32_t sbase = smem_base + SM_STAGES + static_cast<uint32_t>(stage) * STAGE_BYTES;
        const uint32_t fb = bar(BAR_FULL + stage);
        if (big) {
          // half tile hh (64 rays): dz_0 and dz_skip halves (two n-panels each), x halves of this slice's panels
          const int hh = u & 1;
          const int npanels = ncols / KC;
          const uint8_t* dz0 = st + stash_off_dz(g, 0) + hh * HALF_PANEL;
          const uint8_t* dzs = st + stash_off_dz(g, g.skip) + hh * HALF_PANEL;
          const uint8_t* xs = st + stash_off_x(g, 4 * slice) + hh * HALF_PANEL;
          if (vp) {
            dz0 = p.g_image + static_cast<size_t>(tile) * GIMG_TILE_BYTES + hh * HALF_PANEL;
            dzs = dz0 + BLOB_BYTES;
            xs = p.e_image + (static_cast<size_t>(tile) * g.nchunks + 4 * slice) * PANEL_BYTES + hh * HALF_PANEL;
          }
          mbar_arrive_expect_tx(fb, static_cast<uint32_t>(4 + npanels) * HALF_PANEL);
          const uint8_t* dzsrc[4] = {dz0, dz0 + PANEL_BYTES, dzs, dzs + PANEL_BYTES};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (!mcast) bulk_g2s_hint(sbase + static_cast<uint32_t>(j) * HALF_PANEL, dzsrc[j], HALF_PANEL, fb, stash_policy);
            else if (static_cast<uint32_t>(j % cw) == crank)
              bulk_g2s_multicast(sbase + static_cast<uint32_t>(j) * HALF_PANEL, dzsrc[j], HALF_PANEL, fb, cmask);
          }
          for (int pn = 0; pn < npanels; ++pn)
            bulk_g2s_hint(sbase + BLOB_BYTES + static_cast<uint32_t>(pn) * HALF_PANEL, xs + static_cast<size_t>(pn) * PANEL_BYTES,
                          HALF_PANEL, fb, stash_policy);
        } else {
          const WUnit w = p.plan.u[u0 + (u % units_per_tile)];
          mbar_arrive_expect_tx(fb, static_cast<uint32_t>(BLOB_BYTES + w.b_bytes));
          bulk_g2s_hint(sbase, st + w.a_off, BLOB_BYTES, fb, stash_policy);
          if (w.b_bytes > 0) bulk_g2s_hint(sbase + BLOB_BYTES, st + w.b_off, static_cast<uint32_t>(w.b_bytes), fb, stash_policy);
        }
        if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == WG_MMA_WARP) {
    // ======================= MMA issuer ===========================================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t idesc_big = umma_idesc_f16(128, ncols > 0 ? ncols : 16, 1, 1);
      const uint32_t idesc_ones = umma_idesc_f16(128, 16, 1, 0);   // A MN-major, B K-major (ones / dzl^T)
      const uint64_t ones_desc = umma_desc_sw128(smem_base + SM_ONES, 0, 1024);
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
          const WUnit w = p.plan.u[u0 + (u % units_per_tile)];
          const bool first = u < units_per_tile;   // first tile of this CTA: overwrite instead of accumulate
#pragma unroll 1
          for (int kk = 0; kk < 8; ++kk) {
            const uint32_t ko = static_cast<uint32_t>(kk) * 2048u;
            const uint64_t ad = umma_desc_sw128(abase + ko, PANEL_BYTES, 1024);
            const uint32_t acc = (!first || kk > 0) ? 1u : 0u;
            if (w.kind == UNIT_W) {
              umma_f16(tmem_base + static_cast<uint32_t>(w.col), ad, umma_desc_sw128(bbase + ko, PANEL_BYTES, 1024),
                       umma_idesc_f16(128, w.ncols, 1, 1), acc);
            } else if (w.kind == UNIT_LAST) {
              // B = dzl^T [16 x 128 rays] K-major: rays 16kk.. live in panel kk/4 at byte (kk%4)*32
              umma_f16(tmem_base + static_cast<uint32_t>(w.col), ad,
                       umma_desc_sw128(bbase + static_cast<uint32_t>(kk >> 2) * 2048u + static_cast<uint32_t>(kk & 3) * 32u, 0, 1024),
                       idesc_ones, acc);
            }
            if (w.bias_col >= 0) umma_f16(tmem_base + static_cast<uint32_t>(w.bias_col), ad, ones_desc, idesc_ones, acc);
          }
        }
        if (mcast) umma_commit_multicast(bar(BAR_EMPTY + stage), cmask);
        else umma_commit(bar(BAR_EMPTY + stage));
        if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
      }
      umma_commit(bar(BAR_DONE));
    }
  } else if (warp < 4) {
    // ======================= accumulator dump ====================================================
    mbar_wait(bar(BAR_DONE), 0, p.err, 0x80);
    tc_fence_after();
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    float* dst = p.partw + (static_cast<size_t>(cta) * HID + tid) * WG_COLS;
    const int used = big ? 512 : ((p.plan.cols_used[slice] + 31) & ~31);
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
  if (p.trace != nullptr && tid == 0) atomicMax(p.trace + 48, globaltimer_ns());   // measurement: last wgrad CTA done
  if (cw > 1) cluster_sync_all();   // no CTA leaves while a peer may still multicast into it or arrive on its barriers
  if (warp == WG_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, WG_COLS);
  }
}

static void wgrad_launch_config(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int grid, int cluster, cudaStream_t stream) {
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(WG_THREADS);
  cfg.dynamicSmemBytes = SM_TOTAL;
  cfg.stream = stream;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(cluster);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
}

cudaError_t launch_field_wgrad(const FieldParams& p, int grid, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(field_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(SM_TOTAL));
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[2];
  wgrad_launch_config(cfg, attr, grid, p.wg_cluster, stream);
  add_pdl_attr(cfg, attr);
  return cudaLaunchKernelEx(&cfg, field_wgrad_kernel, p);
}

// Largest number of CTAs (a multiple of `cluster`) that are co-resident with clusters of that size, 0 on error.
int field_wgrad_max_grid(int cluster) {
  cudaFuncSetAttribute(field_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(SM_TOTAL));
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  wgrad_launch_config(cfg, attr, cluster, cluster, nullptr);
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, field_wgrad_kernel, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n * cluster;
}

// ---------------------------------------------------------------------------------------------
// Partial reduction (+ optional multi-GPU sum, Adam step and operand refresh), fixed summation order.  Shared helpers first,
// then the generic kernel (addresses derived per thread), then the table-driven kernel training steps use.
// ---------------------------------------------------------------------------------------------
// INF_ADAM_STEP_AUTOINC: every block has read the step counter in its prologue; the last block to leave advances it
__device__ __forceinline__ void step_autoinc_finish(const AdamParams& a) {
  if (!a.step_autoinc || a.params == nullptr) return;
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(a.done_ctr, 1u) == gridDim.x - 1) {
    atomicExch(a.done_ctr, 0u);
    *a.step_dev = *a.step_dev + 1;
  }
}

struct AdamState { float m, v, p; };
__device__ __forceinline__ AdamState adam_load(const AdamParams& a, int64_t i) {
  AdamState st;
  st.m = a.exp_avg[i]; st.v = a.exp_avg_sq[i]; st.p = a.params[i];
  return st;
}
__device__ __forceinline__ void adam_finish(const AdamParams& a, float step_size, float bc2_sqrt, int64_t i, float g, const AdamState& st,
                                            float& p_new) {
  // torch/optim/adam.py (single-tensor math, also the fused kernel's): lerp, mul+addcmul, bias-corrected addcdiv
  const float m = st.m + (g - st.m) * a.omb1;
  const float v = st.v * a.beta2 + a.omb2 * g * g;
  const float denom = sqrtf(v) / bc2_sqrt + a.eps;
  p_new = st.p - step_size * (m / denom);
  a.exp_avg[i] = m;
  a.exp_avg_sq[i] = v;
  a.params[i] = p_new;
}
__device__ __forceinline__ void adam_update(const AdamParams& a, float step_size, float bc2_sqrt, int64_t i, float g, float& p_new) {
  const AdamState st = adam_load(a, i);
  adam_finish(a, step_size, bc2_sqrt, i, g, st, p_new);
}

// system-scope accesses for the peer-memory exchange (peer lines are cached in L1 only: never trust a plain load).
// (value, step number) pairs travel as ONE 64-bit access each -- single-copy atomic, so a reader never sees a torn pair.
__device__ __forceinline__ void st_sys_pair(float* p, float v, uint32_t seq) {
  const unsigned long long w = (static_cast<unsigned long long>(seq) << 32) | __float_as_uint(v);
  asm volatile("st.relaxed.sys.global.b64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ bool ld_sys_pair(const float* p, uint32_t seq, float& v) {
  unsigned long long w;
  asm volatile("ld.relaxed.sys.global.b64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
  v = __uint_as_float(static_cast<uint32_t>(w));
  return static_cast<uint32_t>(w >> 32) == seq;
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}


__device__ __forceinline__ void reduce_item(const ReduceParams& rp, int sidx, uint32_t row, uint32_t c0, int nc, int p, int par, float (&g4)[4]);
__device__ __forceinline__ void apply_update(const ReduceParams& rp, int64_t i, int sidx, float grad, float dev_step_size, float dev_bc2_sqrt,
                                             AdamState st);

// beta^t by square-and-multiply in double: <= 31 dependent multiplications instead of a pow() call (t < 2^31; the relative
// error grows like t * 2^-53, far below what the float results below resolve)
__device__ __forceinline__ double pow_int(double b, uint32_t t) {
  double r = 1.0;
  while (t != 0u) {
    if (t & 1u) r *= b;
    b *= b;
    t >>= 1;
  }
  return r;
}

// Work slot -> (segment, row, first column, #columns, lane p of the item).  An item is up to four consecutive columns of one
// row of one parameter tensor (the wide tensors read their partials with 16-byte loads); `par` adjacent lanes share an item
// and split its ray / vertex groups between them (a hidden weight is the sum of ~50 CTAs' partials: one lane walking them
// eight loads at a time was the long pole of this kernel).  Returns -1 for the padding slots between tensors.
__device__ __forceinline__ int item_decode(const ReduceParams& rp, int64_t slot, uint32_t& row, uint32_t& c0, int& nc, int& p, int& par) {
  int sidx = 0;
  for (int s = 1; s < rp.nseg; ++s)
    if (slot >= rp.seg[s].item_off) sidx = s;
  par = rp.seg[sidx].par;
  const uint32_t ls = static_cast<uint32_t>(slot - rp.seg[sidx].item_off);
  const uint32_t li = ls / static_cast<uint32_t>(par);
  p = static_cast<int>(ls - li * static_cast<uint32_t>(par));
  const uint32_t ipr = static_cast<uint32_t>(rp.seg[sidx].ipr);
  const int cpi = rp.seg[sidx].cpi;
  row = li / ipr;
  c0 = (li - row * ipr) * static_cast<uint32_t>(cpi);
  const int left = rp.seg[sidx].cols - static_cast<int>(c0);
  nc = left < cpi ? left : cpi;
  return row < static_cast<uint32_t>(rp.seg[sidx].rows) ? sidx : -1;
}

// The lanes of an item add up their shares in lane order (every lane ends up with the same totals): fixed order, so the
// gradient stays bit-identical from run to run.  Executed by all 32 lanes of the warp.
__device__ __forceinline__ void item_combine(float (&g4)[4], int par) {
  const int lane = threadIdx.x & 31;
  const int base = lane & ~(par - 1);
  float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int q = 0; q < 4; ++q) {   // par <= 4
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float o = __shfl_sync(0xffffffffu, g4[j], (base + q) & 31);
      if (q < par) t[j] += o;
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) g4[j] = t[j];
}

// sums of one slot's item -> flat / exchange slot `dst` -> (optionally) Adam; element j of the item is finished by lane j % par
__device__ __forceinline__ void process_slot(const ReduceParams& rp, int64_t slot, float* dst, bool update, float step_size, float bc2_sqrt) {
  uint32_t row = 0, c0 = 0;
  int nc = 0, p = 0, par = 1;
  int sidx = -1;
  if (slot < rp.nitems) sidx = item_decode(rp, slot, row, c0, nc, p, par);
  if (sidx >= 0 && rp.do_reduce && rp.seg[sidx].kind == 4) sidx = -1;   // block 0 did the output-layer bias
  float g4[4] = {0.f, 0.f, 0.f, 0.f};
  int64_t i0 = 0;
  if (sidx >= 0) {
    i0 = rp.seg[sidx].off + static_cast<int64_t>(row) * rp.seg[sidx].cols + c0;
    if (rp.do_reduce) reduce_item(rp, sidx, row, c0, nc, p, par, g4);
  }
  if (rp.do_reduce) item_combine(g4, par);
  if (sidx < 0) return;
  AdamState st[4];
  const bool adam = update && rp.adam.params != nullptr;
  const int pm = par < 4 ? par : 4;   // lanes of the item that finish elements: element j belongs to lane j % pm
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (j < nc && (j & (pm - 1)) == p) {
      if (!rp.do_reduce) g4[j] = rp.flat[i0 + j];
      if (adam) st[j] = adam_load(rp.adam, i0 + j);
    }
  }
  // Rolled on purpose: every launch starts with a cold instruction cache and each warp runs its path exactly once, so this
  // kernel is bound by its own instruction fetch (ncu: stall_no_inst is the top stall; ~11 us from block start to the last
  // sum whatever the number of partials) -- four inlined copies of the optimiser tail were most of the executed bytes.
#pragma unroll 1
  for (int j = 0; j < 4; ++j) {
    if (j < nc && (j & (pm - 1)) == p) {
      float gj = j == 0 ? g4[0] : (j == 1 ? g4[1] : (j == 2 ? g4[2] : g4[3]));
      if (rp.do_reduce) {
        gj *= rp.scale;
        // the table holds E[:,j] * 2^e_j, so dW0 / dWy came out multiplied by 2^e_j (exact: powers of two)
        if (rp.seg[sidx].kind == 0 && rp.col_exp != nullptr) gj = ldexpf(gj, -static_cast<int>(rp.col_exp[c0 + j]));
        dst[i0 + j] = gj;
      }
      if (adam) {
        AdamState sj;
        sj.m = j == 0 ? st[0].m : (j == 1 ? st[1].m : (j == 2 ? st[2].m : st[3].m));
        sj.v = j == 0 ? st[0].v : (j == 1 ? st[1].v : (j == 2 ? st[2].v : st[3].v));
        sj.p = j == 0 ? st[0].p : (j == 1 ? st[1].p : (j == 2 ? st[2].p : st[3].p));
        apply_update(rp, i0 + j, sidx, gj, step_size, bc2_sqrt, sj);
      }
    }
  }
}

// Generic reduction kernel: every address derived from the parameter block.  Used for the optimiser tail on an already
// (all-)reduced gradient (inf_adam_pack_step: do_reduce = 0) and as the fallback when the tables of the table-driven kernel
// below cannot be built (stream capture in progress, table budget exhausted); single GPU only.
__global__ void __launch_bounds__(256, 4) field_reduce_kernel(const ReduceParams rp) {   // 4 CTAs per SM: one resident wave at the bench shape
  const int64_t tid0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t nthreads = static_cast<int64_t>(gridDim.x) * blockDim.x;
  // Adam's bias corrections from the device-resident step counter, once per CTA (FP64 transcendental code in every thread
  // was most of this kernel's first version; a pow() call at the head of every CTA still delayed all its warps)
  __shared__ float s_bias_corr[2];
  if (threadIdx.x == 0) {
    if (rp.adam.params != nullptr && rp.adam.step_dev != nullptr) {
      const int32_t t = *rp.adam.step_dev + rp.adam.step_autoinc;
      const uint32_t tu = t > 0 ? static_cast<uint32_t>(t) : 0u;
      s_bias_corr[0] = static_cast<float>(static_cast<double>(rp.adam.lr) / (1.0 - pow_int(static_cast<double>(rp.adam.beta1), tu)));
      s_bias_corr[1] = static_cast<float>(sqrt(1.0 - pow_int(static_cast<double>(rp.adam.beta2), tu)));
    }
  }
  if (rp.pdl_early) pdl_launch_dependents();
  pdl_wait();   // the partials (field_wgrad, and through it every kernel of the step)
  __syncthreads();
  // The forward/backward kernel's per-CTA partials (output-layer bias gradient, loss sums: ncta1 rows of 8 floats): block 0
  // sums them as a block -- thread t takes rows t, t + 256, ..., then a fixed shuffle tree and a fixed order over the warps,
  // so the result is deterministic.  (One thread walking the ~300 rows, four loads in flight, was 25 of this kernel's 30 us.)
  __shared__ float s_part[5][8];
  __shared__ float s_tot[5];   // gb[0..2] (unscaled), loss sums
  if (blockIdx.x == 0 && rp.do_reduce) {
    float loss_a, loss_b;
    float a[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int c = threadIdx.x; c < rp.ncta1; c += blockDim.x) {
      const float4* pl = reinterpret_cast<const float4*>(rp.part1 + static_cast<size_t>(c) * K1_PART_FLOATS);
      const float4 lo = pl[0], hi = pl[1];
      a[0] += lo.x; a[1] += lo.y; a[2] += lo.z; a[3] += hi.x; a[4] += hi.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int j = 0; j < 5; ++j) a[j] += __shfl_xor_sync(0xffffffffu, a[j], o);
    if ((threadIdx.x & 31) == 0)
#pragma unroll
      for (int j = 0; j < 5; ++j) s_part[j][threadIdx.x >> 5] = a[j];
    __syncthreads();
    if (threadIdx.x < 5) {
      float t = 0.f;
      for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) t += s_part[threadIdx.x][w];
      s_tot[threadIdx.x] = t;
    }
    __syncthreads();
    loss_a = s_tot[3];
    loss_b = s_tot[4];
    if (threadIdx.x == 0 && rp.loss_sums != nullptr) { rp.loss_sums[0] = loss_a; rp.loss_sums[1] = loss_b; rp.loss_sums[2] = loss_a * rp.loss_scale; }
    if (threadIdx.x < OUT_DIM) {   // the output-layer bias gradient (the last tensor, kind 4) is finished here
      const int sidx = rp.nseg - 1;
      const int64_t i = rp.seg[sidx].off + threadIdx.x;
      const float gbias = s_tot[threadIdx.x] * rp.scale;
      rp.flat[i] = gbias;
      if (rp.adam.params != nullptr) apply_update(rp, i, sidx, gbias, s_bias_corr[0], s_bias_corr[1], adam_load(rp.adam, i));
    }
  }
  const int64_t nslots32 = (rp.nitems + 31) & ~static_cast<int64_t>(31);   // whole warps walk the slots (item_combine shuffles)
  for (int64_t slot = tid0; slot < nslots32; slot += nthreads) process_slot(rp, slot, rp.flat, true, s_bias_corr[0], s_bias_corr[1]);
  step_autoinc_finish(rp.adam);
}

// Lane p's share (groups p, p + par, ...) of the sums of an item's up to four consecutive columns, unscaled.  The wide tensors
// (W0 / Wy, hidden weights) read their partials with 16-byte loads; all of a lane's loads are in flight before its first
// addition.
__device__ __forceinline__ void reduce_item(const ReduceParams& rp, const int sidx, const uint32_t row, const uint32_t c0, const int nc,
                                            const int p, const int par, float (&g4)[4]) {
  const ReduceSeg sg = rp.seg[sidx];
  int cta0 = -1, ng = 0, gmul = 1, c = 0;
  if (sg.kind == 0) {
    cta0 = static_cast<int>(c0 >> 8);
    ng = rp.groups_big;
    gmul = rp.geo.nbig;
    c = static_cast<int>(c0 & 255u) + (sg.a ? 256 : 0);
  } else if (sg.kind == 1) {
    const int split = sg.split;
    const bool first = static_cast<int>(c0) + nc <= split, second = static_cast<int>(c0) >= split;
    if (first || second) {
      const WUnit w = rp.plan.u[first ? sg.a : sg.b];
      const int cc = w.col + static_cast<int>(c0) - (first ? 0 : split);
      if ((cc & 3) == 0 && cc + 4 <= WG_COLS) {
        cta0 = rp.cta_base_hid[w.slice];
        ng = rp.groups_hid[w.slice];
        c = cc;
      }
    }
  }
  if (cta0 < 0) {
    // narrow tensors (biases, output-layer weight, a straddled or misaligned view-layer split): one partial address per
    // element; the lane's loads of all (<= 4) elements are in flight together (these few warps were the kernel's long pole
    // when they walked their elements one after the other)
    const float* sp[4] = {rp.partw, rp.partw, rp.partw, rp.partw};
    int ngj[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j >= nc) continue;
      const int col = static_cast<int>(c0) + j;
      int ui, ln, cc;
      if (sg.kind == 1) {
        const bool first = col < sg.split;
        ui = first ? sg.a : sg.b;
        ln = static_cast<int>(row);
        cc = rp.plan.u[ui].col + (first ? col : col - sg.split);
      } else if (sg.kind == 2) {
        ui = sg.a; ln = static_cast<int>(row); cc = rp.plan.u[ui].bias_col;
      } else {   // kind 3: lane = input feature, column = channel
        ui = sg.a; ln = col; cc = rp.plan.u[ui].col + static_cast<int>(row);
      }
      const int slice = rp.plan.u[ui].slice;
      ngj[j] = rp.groups_hid[slice];
      sp[j] = rp.partw + (static_cast<size_t>(rp.cta_base_hid[slice]) * HID + ln) * WG_COLS + cc;
    }
    const size_t GS = static_cast<size_t>(HID) * WG_COLS;
    const int ngmax = max(max(ngj[0], ngj[1]), max(ngj[2], ngj[3]));
    for (int g0 = p; g0 < ngmax; g0 += 8 * par) {
      float v[8][4];
#pragma unroll
      for (int q = 0; q < 8; ++q)
#pragma unroll
        for (int j = 0; j < 4; ++j) v[q][j] = g0 + q * par < ngj[j] ? sp[j][static_cast<size_t>(g0 + q * par) * GS] : 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (g0 + q * par < ngj[j]) g4[j] += v[q][j];
    }
    return;
  }
  const float4* src = reinterpret_cast<const float4*>(rp.partw + (static_cast<size_t>(cta0) * HID + row) * WG_COLS + c);
  const size_t GSTRIDE4 = static_cast<size_t>(HID) * WG_COLS * static_cast<size_t>(gmul) / 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int g0 = p; g0 < ng; g0 += 8 * par) {
    float4 v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = g0 + q * par < ng ? src[static_cast<size_t>(g0 + q * par) * GSTRIDE4] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (g0 + q * par < ng) { acc.x += v[q].x; acc.y += v[q].y; acc.z += v[q].z; acc.w += v[q].w; }
  }
  g4[0] = acc.x; g4[1] = acc.y; g4[2] = acc.z; g4[3] = acc.w;
}

// Adam on the master parameter of flat element i, then its entry of the fp16 / f32 operand images
__device__ __forceinline__ void apply_update(const ReduceParams& rp, const int64_t i, const int sidx, const float grad,
                                             const float dev_step_size, const float dev_bc2_sqrt, const AdamState st) {
  const AdamParams& ad = rp.adam;
  if (ad.params == nullptr) return;
  const ReduceSeg sg = rp.seg[sidx];
  const uint32_t e = static_cast<uint32_t>(i - sg.off);
  const int row = static_cast<int>(e / static_cast<uint32_t>(sg.cols));
  const int col = static_cast<int>(e - static_cast<uint32_t>(row) * static_cast<uint32_t>(sg.cols));
  if (sg.kind == 2 && sg.pair_off == -2) return;   // Ly.bias: handled by the thread of Lx.bias
  float step_size = ad.step_size, bc2_sqrt = ad.bc2_sqrt;
  if (ad.step_dev != nullptr) {
    step_size = dev_step_size;
    bc2_sqrt = dev_bc2_sqrt;
  }
  float pn;
  adam_finish(ad, step_size, bc2_sqrt, i, grad, st, pn);
  const Geo& g = rp.geo;
  float* f32sec = reinterpret_cast<float*>(ad.packed + packed_f32_offset(g));
  const int nbias = (g.nb + 1) * HID;
  if (sg.kind == 0) {
    const uint32_t off = sw128_off(static_cast<uint32_t>(row + (sg.a ? HID : 0)), static_cast<uint32_t>(col & 63));
    const float img = ad.col_exp != nullptr ? ldexpf(pn, -static_cast<int>(ad.col_exp[col])) : pn;
    *reinterpret_cast<__half*>(ad.packed + static_cast<size_t>(col >> 6) * BLOB_BYTES + off) = __float2half_rn(img);
  } else if (sg.kind == 1) {
    int blob, jj = col;
    if (g.view && sg.layer == g.nb) {
      if (col < sg.split) blob = blob_wd1z(g);
      else { blob = blob_wd1e(g); jj = col - sg.split; }
    } else {
      blob = blob_hidden(g, sg.layer);
    }
    const uint32_t off = static_cast<uint32_t>(jj >> 6) * PANEL_BYTES + sw128_off(static_cast<uint32_t>(row), static_cast<uint32_t>(jj & 63));
    *reinterpret_cast<__half*>(ad.packed + static_cast<size_t>(blob) * BLOB_BYTES + off) = __float2half_rn(pn);
  } else if (sg.kind == 2) {
    float bsum = pn;
    if (sg.pair_off >= 0) {   // skip layer: the image holds Lx.bias + Ly.bias; both share this gradient
      float p2;
      adam_update(ad, step_size, bc2_sqrt, sg.pair_off + row, grad, p2);
      bsum += p2;
    }
    f32sec[sg.layer * HID + row] = bsum;
  } else if (sg.kind == 3) {
    f32sec[nbias + row * HID + col] = pn;
  } else {
    f32sec[nbias + OUT_DIM * HID + row] = pn;
  }
}

// ---------------------------------------------------------------------------------------------
// Table-driven reduction (the one training steps use).  The generic kernel above derives every address from the parameter
// block: segment searches, divisions, unit / slice lookups with run-time-indexed constant loads, several inlined copies of
// the optimiser tail -- thousands of instructions that each warp runs exactly once, on a cold instruction and constant
// cache at every launch; measured, that (not the 25 MB of partials) is what the kernel's ~17 us were made of, twice over
// with the two phases of the multi-GPU exchange.  Here reduce_plan_kernel evaluates the same address logic ONCE per (model
// geometry, wgrad CTA split) into two device tables, and the per-step kernel is: one 16-byte descriptor load, the partial
// loads, a shuffle combine, one 8-byte descriptor per element for the optimiser tail.
// ---------------------------------------------------------------------------------------------
__global__ void reduce_plan_kernel(const ReduceParams rp, ItemDesc* __restrict__ items, ElemDesc* __restrict__ elems) {
  const int64_t tid0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t nthreads = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t nslots32 = (rp.nitems + 31) & ~static_cast<int64_t>(31);
  for (int64_t slot = tid0; slot < nslots32; slot += nthreads) {
    ItemDesc d{};
    uint32_t row = 0, c0 = 0;
    int nc = 0, p = 0, par = 1;
    int sidx = -1;
    if (slot < rp.nitems) sidx = item_decode(rp, slot, row, c0, nc, p, par);
    if (sidx >= 0 && rp.seg[sidx].kind != 4) {   // kind 4 (output-layer bias) is finished by block 0 of the reduction
      const ReduceSeg sg = rp.seg[sidx];
      size_t src = 0, gs = 0;
      int ngt = 0;
      bool vec = false;
      if (sg.kind == 0) {
        src = (static_cast<size_t>(c0 >> 8) * HID + row) * WG_COLS + (c0 & 255u) + (sg.a ? 256 : 0);
        gs = static_cast<size_t>(HID) * WG_COLS * static_cast<size_t>(rp.geo.nbig);
        ngt = rp.groups_big;
        vec = true;
      } else {
        const int col = static_cast<int>(c0);
        int ui, ln, cc;
        if (sg.kind == 1) {
          const bool first = col < sg.split;
          ui = first ? sg.a : sg.b;
          ln = static_cast<int>(row);
          cc = rp.plan.u[ui].col + (first ? col : col - sg.split);
          vec = sg.cpi == 4;   // the host vectorises a split tensor only when the split is a multiple of four
        } else if (sg.kind == 2) {
          ui = sg.a; ln = static_cast<int>(row); cc = rp.plan.u[ui].bias_col;
        } else {   // kind 3: lane = input feature, column = channel
          ui = sg.a; ln = col; cc = rp.plan.u[ui].col + static_cast<int>(row);
        }
        const int slice = rp.plan.u[ui].slice;
        src = (static_cast<size_t>(rp.cta_base_hid[slice]) * HID + ln) * WG_COLS + cc;
        gs = static_cast<size_t>(HID) * WG_COLS;
        ngt = rp.groups_hid[slice];
      }
      int lg = 0;
      while ((1 << lg) < par) ++lg;
      d.src = static_cast<uint32_t>(src + static_cast<size_t>(p) * gs);
      d.gstride = static_cast<uint32_t>(gs * static_cast<size_t>(par));
      d.i0 = static_cast<uint32_t>(sg.off + static_cast<int64_t>(row) * sg.cols + c0);
      d.ng = static_cast<uint16_t>(p < ngt ? (ngt - p + par - 1) / par : 0);
      d.nc = static_cast<uint8_t>(nc);
      d.flags = static_cast<uint8_t>((vec ? 1 : 0) | (lg << 2) | (p << 4));
    }
    items[slot] = d;
  }
  const Geo& g = rp.geo;
  const uint32_t f32off = static_cast<uint32_t>(packed_f32_offset(g));
  const int nbias = (g.nb + 1) * HID;
  for (int64_t i = tid0; i < rp.total; i += nthreads) {
    int sidx = 0;
    for (int s = 1; s < rp.nseg; ++s) if (i >= rp.seg[s].off) sidx = s;
    const ReduceSeg sg = rp.seg[sidx];
    const uint32_t e = static_cast<uint32_t>(i - sg.off);
    const int row = static_cast<int>(e / static_cast<uint32_t>(sg.cols));
    const int col = static_cast<int>(e - static_cast<uint32_t>(row) * static_cast<uint32_t>(sg.cols));
    ElemDesc ed{};
    if (sg.kind == 0) {
      ed.poff = static_cast<uint32_t>(static_cast<size_t>(col >> 6) * BLOB_BYTES +
                                      sw128_off(static_cast<uint32_t>(row + (sg.a ? HID : 0)), static_cast<uint32_t>(col & 63)));
      ed.aux = 0u | 4u | (static_cast<uint32_t>(col) << 3);
    } else if (sg.kind == 1) {
      int blob, jj = col;
      if (g.view && sg.layer == g.nb) {
        if (col < sg.split) blob = blob_wd1z(g);
        else { blob = blob_wd1e(g); jj = col - sg.split; }
      } else {
        blob = blob_hidden(g, sg.layer);
      }
      ed.poff = static_cast<uint32_t>(static_cast<size_t>(blob) * BLOB_BYTES + static_cast<uint32_t>(jj >> 6) * PANEL_BYTES +
                                      sw128_off(static_cast<uint32_t>(row), static_cast<uint32_t>(jj & 63)));
      ed.aux = 0u;
    } else if (sg.kind == 2) {
      ed.poff = f32off + static_cast<uint32_t>(sg.layer * HID + row) * 4u;
      if (sg.pair_off == -2) ed.aux = 3u;
      else if (sg.pair_off >= 0) ed.aux = 2u | (static_cast<uint32_t>(sg.pair_off + row) << 3);
      else ed.aux = 1u;
    } else if (sg.kind == 3) {
      ed.poff = f32off + static_cast<uint32_t>(nbias + row * HID + col) * 4u;
      ed.aux = 1u;
    } else {
      ed.poff = f32off + static_cast<uint32_t>(nbias + OUT_DIM * HID + row) * 4u;
      ed.aux = 1u;
    }
    elems[i] = ed;
  }
}

// optimiser tail of one element from its descriptor (same arithmetic as apply_update)
__device__ __forceinline__ void lean_update(const AdamParams& ad, const ElemDesc ed, const int64_t i, const float grad, const float step_size,
                                            const float bc2_sqrt, const AdamState st) {
  const uint32_t type = ed.aux & 3u;
  if (type == 3u) return;   // Ly.bias: updated by the thread of Lx.bias
  float pn;
  adam_finish(ad, step_size, bc2_sqrt, i, grad, st, pn);
  if (type == 0u) {
    const float img = ((ed.aux & 4u) && ad.col_exp != nullptr) ? ldexpf(pn, -static_cast<int>(ad.col_exp[ed.aux >> 3])) : pn;
    *reinterpret_cast<__half*>(ad.packed + ed.poff) = __float2half_rn(img);
  } else {
    if (type == 2u) {   // skip layer: the image holds Lx.bias + Ly.bias; both share this gradient
      float p2;
      adam_update(ad, step_size, bc2_sqrt, static_cast<int64_t>(ed.aux >> 3), grad, p2);
      pn += p2;
    }
    *reinterpret_cast<float*>(ad.packed + ed.poff) = pn;
  }
}

// Sum over ranks of one element, exchanged as (value, step number) pairs: push this rank's value into slot (parity, rank) of
// every peer's region, then add the ranks' values IN RANK ORDER (bit-identical on every rank), spinning on the local copy of a
// peer's pair until it carries this step's number.  `i` = element index inside a slot (flat index; total, total + 1: loss sums).
__device__ __forceinline__ void exch_push(const LeanParams& lp, const uint32_t seq, const int64_t i, const float g) {
  const CommParams& cm = lp.comm;
  const int64_t off = comm_slot_off(cm.slot_floats, seq & 1u, cm.rank) + 2 * i;
  for (int q = 0; q < cm.world; ++q)
    if (q != cm.rank) st_sys_pair(cm.region[q] + off, g, seq);   // posted store over NVLink
}
// local copy of slot `src`'s pair of element i, once it carries this step's number
__device__ __forceinline__ float exch_wait(const LeanParams& lp, const uint32_t seq, const int src, const int64_t i) {
  const CommParams& cm = lp.comm;
  const float* p = cm.region[cm.rank] + comm_slot_off(cm.slot_floats, seq & 1u, src) + 2 * i;
  float v = 0.f;
  const long long t0 = clock64();
  while (!ld_sys_pair(p, seq, v)) {
    if (clock64() - t0 > (1ll << 33)) {   // ~4 s: a peer died; surface it instead of hanging the GPU
      if (lp.err) atomicExch(lp.err, 0xDEAD0C00u | static_cast<unsigned>(src));
      break;
    }
  }
  return v;
}
__device__ __forceinline__ float exch_sum(const LeanParams& lp, const uint32_t seq, const int64_t i, const float g) {
  const CommParams& cm = lp.comm;
  float sum = 0.f;
  for (int q = 0; q < cm.world; ++q) sum += q == cm.rank ? g : exch_wait(lp, seq, q, i);
  return sum;
}
// two-shot exchange: the rank that sums element i
__device__ __forceinline__ int exch_owner(const CommParams& cm, const int64_t i) {
  const int o = static_cast<int>(i / cm.own_chunk);
  return o < cm.world ? o : cm.world - 1;
}

// one work slot: this lane's share of the item's sums, combined over the item's lanes, then the elements this lane finishes
// (`seq` != 0: multi-GPU step -- the scaled gradient is summed over the ranks before it reaches flat + optimiser)
// In a multi-GPU step the block stages its (value, index) pairs in shared memory and pushes them with consecutive threads on
// consecutive elements: an item's lane owns up to four neighbouring elements, so pushing from the owner would put 8-byte
// stores 32 bytes apart -- 330 k separate NVLink write requests per peer, measured at +15 us per step, whereas whole
// 256-byte runs per warp instruction cost about what the 2.6 MB payload does.
__device__ __forceinline__ void lean_slot(const LeanParams& lp, const int64_t slot, const uint32_t seq, const float step_size,
                                          const float bc2_sqrt, float* s_val, int* s_idx) {
  ItemDesc d{};
  if (slot < lp.nslots) d = lp.items[slot];
  if (lp.pdl_early & 2) d.ng = 0;   // measurement (INF_REDUCE_DBG bit 0): no partial loads -- wrong results
  float g4[4] = {0.f, 0.f, 0.f, 0.f};
  if (d.nc != 0) {
    const float* src = lp.partw + d.src;
    if (d.flags & 1u) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int g0 = 0; g0 < d.ng; g0 += 8) {
        float4 v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q)
          v[q] = g0 + q < d.ng ? *reinterpret_cast<const float4*>(src + static_cast<size_t>(g0 + q) * d.gstride) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (g0 + q < d.ng) { acc.x += v[q].x; acc.y += v[q].y; acc.z += v[q].z; acc.w += v[q].w; }
      }
      g4[0] = acc.x; g4[1] = acc.y; g4[2] = acc.z; g4[3] = acc.w;
    } else {
      float a = 0.f;
      for (int g0 = 0; g0 < d.ng; g0 += 8) {
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = g0 + q < d.ng ? src[static_cast<size_t>(g0 + q) * d.gstride] : 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (g0 + q < d.ng) a += v[q];
      }
      g4[0] = a;
    }
  }
  const int par = 1 << ((d.flags >> 2) & 3u);
  item_combine(g4, par);
  if (d.nc == 0 && seq == 0u) return;   // padding slot (a multi-GPU step still meets the block at the barriers below)
  const int p = (d.flags >> 4) & 3u;
  const bool adam = lp.adam.params != nullptr && !(lp.pdl_early & 4);   // (INF_REDUCE_DBG bit 1: no optimiser tail)
  ElemDesc ed[4];
  AdamState st[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (j < d.nc && (j & (par - 1)) == p) {
      ed[j] = lp.elems[d.i0 + j];
      g4[j] *= lp.scale;
      // the table holds E[:,j] * 2^e_j, so dW0 / dWy came out multiplied by 2^e_j (exact: powers of two)
      if ((ed[j].aux & 4u) && (ed[j].aux & 3u) == 0u && lp.col_exp != nullptr) g4[j] = ldexpf(g4[j], -static_cast<int>(lp.col_exp[ed[j].aux >> 3]));
      if (adam) st[j] = adam_load(lp.adam, d.i0 + j);
    }
  }
  const CommParams& cm = lp.comm;
  const bool two_shot = seq != 0u && cm.own_chunk > 0;
  if (seq != 0u) {
    // every thread of the block is here (the slot loop's trip count is uniform per block); entries without an owner are -1
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool own = j < d.nc && (j & (par - 1)) == p;
      s_val[threadIdx.x * 4 + j] = g4[j];
      s_idx[threadIdx.x * 4 + j] = own ? static_cast<int>(d.i0) + j : -1;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int e = k * static_cast<int>(blockDim.x) + static_cast<int>(threadIdx.x);
      const int i = s_idx[e];
      if (i >= 0) {   // on its way while the optimiser state arrives
        if (!two_shot) {
          exch_push(lp, seq, i, s_val[e]);
        } else {       // first shot: this rank's value goes to the element's summing rank only
          const int o = exch_owner(cm, i);
          if (o != cm.rank) st_sys_pair(cm.region[o] + comm_slot_off(cm.slot_floats, seq & 1u, cm.rank) + 2 * static_cast<int64_t>(i), s_val[e], seq);
        }
      }
    }
    __syncthreads();   // the staging buffers are free again
  }
  float tot[4] = {g4[0], g4[1], g4[2], g4[3]};
  if (seq != 0u) {
    bool mine[4] = {false, false, false, false};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j < d.nc && (j & (par - 1)) == p) {
        const int64_t i = static_cast<int64_t>(d.i0) + j;
        mine[j] = !two_shot || exch_owner(cm, i) == cm.rank;
        if (mine[j]) tot[j] = exch_sum(lp, seq, i, g4[j]);   // rank order: the same bits wherever the sum is made
      }
    }
    if (two_shot) {
      // second shot: the summing rank sends the finished sum to everyone (again as whole runs per warp)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s_val[threadIdx.x * 4 + j] = tot[j];
        s_idx[threadIdx.x * 4 + j] = (mine[j] && j < d.nc && (j & (par - 1)) == p) ? static_cast<int>(d.i0) + j : -1;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int e = k * static_cast<int>(blockDim.x) + static_cast<int>(threadIdx.x);
        const int i = s_idx[e];
        if (i >= 0) {
          const int64_t off = comm_slot_off(cm.slot_floats, seq & 1u, COMM_SLOT_RESULT) + 2 * static_cast<int64_t>(i);
          for (int q = 0; q < cm.world; ++q)
            if (q != cm.rank) st_sys_pair(cm.region[q] + off, s_val[e], seq);
        }
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < d.nc && (j & (par - 1)) == p && !mine[j]) tot[j] = exch_wait(lp, seq, COMM_SLOT_RESULT, static_cast<int64_t>(d.i0) + j);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (j < d.nc && (j & (par - 1)) == p) {
      const int64_t i = static_cast<int64_t>(d.i0) + j;
      lp.flat[i] = tot[j];
      if (adam) lean_update(lp.adam, ed[j], i, tot[j], step_size, bc2_sqrt, st[j]);
    }
  }
}

__global__ void __launch_bounds__(256, 4) field_reduce_lean_kernel(const LeanParams lp) {
  const int64_t tid0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t nthreads = static_cast<int64_t>(gridDim.x) * blockDim.x;
  __shared__ float s_bias_corr[2];
  __shared__ uint32_t s_seq;
  __shared__ float s_part[5][8];
  __shared__ float s_tot[5];   // gb[0..2] (unscaled), loss sums
  const CommParams& cm = lp.comm;
  const bool exchange = cm.world > 1;
  uint32_t* ctl = exchange ? reinterpret_cast<uint32_t*>(cm.region[cm.rank] + comm_ctl_off(cm.slot_floats)) : nullptr;
  if (threadIdx.x == 0) {
    s_bias_corr[0] = lp.adam.step_size;
    s_bias_corr[1] = lp.adam.bc2_sqrt;
    if (lp.adam.params != nullptr && lp.adam.step_dev != nullptr) {
      const int32_t t = *lp.adam.step_dev + lp.adam.step_autoinc;
      const uint32_t tu = t > 0 ? static_cast<uint32_t>(t) : 0u;
      s_bias_corr[0] = static_cast<float>(static_cast<double>(lp.adam.lr) / (1.0 - pow_int(static_cast<double>(lp.adam.beta1), tu)));
      s_bias_corr[1] = static_cast<float>(sqrt(1.0 - pow_int(static_cast<double>(lp.adam.beta2), tu)));
    }
    // this step's number (> 0; advanced by the last block to leave, so every block of the launch reads the same value)
    s_seq = exchange ? ld_acquire_sys_u32(ctl + CTL_SEQ) + 1u : 0u;
  }
  if (lp.pdl_early & 1) pdl_launch_dependents();
  pdl_wait();   // the partials (field_wgrad, and through it every kernel of the step)
  __syncthreads();
  const float step_size = s_bias_corr[0], bc2_sqrt = s_bias_corr[1];
  const uint32_t seq = s_seq;
  const bool tracing = lp.trace != nullptr && threadIdx.x == 0;
  if (tracing) {   // measurement (INF_DEBUG & 8): earliest block start, latest block start
    atomicMax(lp.trace + 44, (1ull << 62) - globaltimer_ns());
    atomicMax(lp.trace + 45, globaltimer_ns());
  }
  // block 0: the forward/backward kernel's per-CTA partials (output-layer bias gradient, loss sums) -- thread t takes rows t,
  // t + 256, ..., then a fixed shuffle tree and a fixed order over the warps (deterministic)
  if (blockIdx.x == 0) {
    float a[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int c = threadIdx.x; c < lp.ncta1; c += blockDim.x) {
      const float4* pl = reinterpret_cast<const float4*>(lp.part1 + static_cast<size_t>(c) * K1_PART_FLOATS);
      const float4 lo = pl[0], hi = pl[1];
      a[0] += lo.x; a[1] += lo.y; a[2] += lo.z; a[3] += hi.x; a[4] += hi.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int j = 0; j < 5; ++j) a[j] += __shfl_xor_sync(0xffffffffu, a[j], o);
    if ((threadIdx.x & 31) == 0)
#pragma unroll
      for (int j = 0; j < 5; ++j) s_part[j][threadIdx.x >> 5] = a[j];
    __syncthreads();
    if (threadIdx.x < 5) {
      float t = 0.f;
      for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) t += s_part[threadIdx.x][w];
      s_tot[threadIdx.x] = t;
    }
    __syncthreads();
    if (threadIdx.x < OUT_DIM) {   // output-layer bias gradient (the last tensor of the flat vector)
      const int64_t i = lp.out_bias_off + threadIdx.x;
      float gbias = s_tot[threadIdx.x] * lp.scale;
      if (exchange) { exch_push(lp, seq, i, gbias); gbias = exch_sum(lp, seq, i, gbias); }
      lp.flat[i] = gbias;
      if (lp.adam.params != nullptr) lean_update(lp.adam, lp.elems[i], i, gbias, step_size, bc2_sqrt, adam_load(lp.adam, i));
    } else if (threadIdx.x < OUT_DIM + 2) {   // the two loss sums travel as elements total, total + 1 of the slot
      const int j = threadIdx.x - OUT_DIM;
      float v = s_tot[3 + j];
      if (exchange) { exch_push(lp, seq, lp.total + j, v); v = exch_sum(lp, seq, lp.total + j, v); }
      if (lp.loss_sums != nullptr) {
        lp.loss_sums[j] = v;
        if (j == 0) lp.loss_sums[2] = v * lp.loss_scale;
      }
    }
  }
  // whole warps walk the slots (item_combine shuffles); whole blocks in a multi-GPU step (staged pushes)
  __shared__ float s_val[1024];
  __shared__ int s_idx[1024];
  const int64_t nslots_up = (lp.nslots + 255) & ~static_cast<int64_t>(255);
  for (int64_t slot = tid0; slot < nslots_up; slot += nthreads) lean_slot(lp, slot, seq, step_size, bc2_sqrt, s_val, s_idx);
  step_autoinc_finish(lp.adam);
  if (lp.trace != nullptr || exchange) {
    __syncthreads();
    if (tracing) atomicMax(lp.trace + 46, globaltimer_ns());   // latest block end
    if (exchange && threadIdx.x == 0 && atomicAdd(ctl + CTL_DONE, 1u) == gridDim.x - 1) {
      atomicExch(ctl + CTL_DONE, 0u);
      st_release_sys_u32(ctl + CTL_SEQ, seq);   // every block has read this step's number long ago
    }
  }
}

namespace {

// work slots of a launch: items x lanes per item; a tensor's slots start at a multiple of 8 so that the lanes of an item sit
// in one warp.  A CTA lives several microseconds whatever it does, so the launch should be ONE resident wave: the lanes per
// item are capped until the slots fit it.  `narrow1`: biases / output-layer weight / a misaligned view-layer split become
// one-column items (the table-driven kernel has no per-element fallback inside an item).
int64_t plan_slots(ReduceParams& rp, int threads, int resident, bool narrow1) {
  int64_t blocks = 0;
  for (int par_max = 4; par_max >= 1; par_max >>= 1) {
    int64_t slots = 0;
    for (int s = 0; s < rp.nseg; ++s) {
      ReduceSeg& sg = rp.seg[s];
      int ng = 1;
      if (rp.do_reduce) {
        if (sg.kind == 0) ng = rp.groups_big;
        else if (sg.kind == 1) {
          ng = rp.groups_hid[rp.plan.u[sg.a].slice];
          if (sg.b >= 0 && rp.groups_hid[rp.plan.u[sg.b].slice] > ng) ng = rp.groups_hid[rp.plan.u[sg.b].slice];
        } else if (sg.kind == 2 || sg.kind == 3) ng = rp.groups_hid[rp.plan.u[sg.a].slice];
      }
      sg.par = ng <= 8 ? 1 : (ng <= 16 ? 2 : 4);   // > 32 groups: two rounds of eight loads per lane
      if (sg.par > par_max) sg.par = par_max;
      const bool wide = sg.kind == 0 || (sg.kind == 1 && (sg.split >= sg.cols || sg.split % 4 == 0));
      sg.cpi = (wide || !narrow1) ? 4 : 1;
      sg.item_off = slots;
      sg.ipr = (sg.cols + sg.cpi - 1) / sg.cpi;
      slots += static_cast<int64_t>(sg.rows) * sg.ipr * sg.par;
      slots = (slots + 7) & ~static_cast<int64_t>(7);
    }
    rp.nitems = slots;
    blocks = (slots + threads - 1) / threads;
    if (blocks <= resident) break;
  }
  return blocks;
}

// Device tables of the table-driven kernel, keyed by everything they are computed from.  Owned by the library and never
// freed (a captured CUDA graph may hold their addresses); capped in total size, beyond which (and while a stream capture is
// in progress, where allocation is not allowed) the generic kernel runs instead.
struct PlanEntry {
  int dev;
  std::string sig;
  ItemDesc* items;
  ElemDesc* elems;
};
std::mutex g_plan_mutex;
std::vector<PlanEntry> g_plans;
size_t g_plan_bytes = 0;
constexpr size_t PLAN_BYTES_MAX = static_cast<size_t>(512) << 20;

std::string plan_signature(const ReduceParams& rp) {
  // field by field (struct padding is not part of the value)
  std::string sig;
  auto put = [&](int64_t v) { sig.append(reinterpret_cast<const char*>(&v), sizeof(v)); };
  put(rp.nseg); put(rp.total); put(rp.nitems); put(rp.groups_big);
  for (int s = 0; s < MAX_HID_SLICES; ++s) { put(rp.groups_hid[s]); put(rp.cta_base_hid[s]); }
  for (int s = 0; s < rp.nseg; ++s) {
    const ReduceSeg& g = rp.seg[s];
    put(g.off); put(g.rows); put(g.cols); put(g.kind); put(g.a); put(g.b); put(g.split); put(g.layer); put(g.pair_off);
    put(g.item_off); put(g.ipr); put(g.par); put(g.cpi);
  }
  sig.append(reinterpret_cast<const char*>(&rp.geo), sizeof(rp.geo));   // all ints: no padding
  put(rp.plan.n); put(rp.plan.nslice);
  for (int u = 0; u < rp.plan.n; ++u) { put(rp.plan.u[u].col); put(rp.plan.u[u].bias_col); put(rp.plan.u[u].slice); put(rp.plan.u[u].layer); }
  return sig;
}

const PlanEntry* plan_get(const ReduceParams& rp, cudaStream_t stream) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  const std::string sig = plan_signature(rp);
  std::lock_guard<std::mutex> lock(g_plan_mutex);
  for (const PlanEntry& e : g_plans)
    if (e.dev == dev && e.sig == sig) return &e;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) { cudaGetLastError(); return nullptr; }
  const int64_t nslots32 = (rp.nitems + 31) & ~static_cast<int64_t>(31);
  const size_t ib = static_cast<size_t>(nslots32) * sizeof(ItemDesc), eb = static_cast<size_t>(rp.total) * sizeof(ElemDesc);
  if (g_plan_bytes + ib + eb > PLAN_BYTES_MAX) return nullptr;
  PlanEntry e{dev, sig, nullptr, nullptr};
  if (cudaMalloc(&e.items, ib) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (cudaMalloc(&e.elems, eb) != cudaSuccess) { cudaGetLastError(); cudaFree(e.items); return nullptr; }
  const int64_t work = nslots32 > rp.total ? nslots32 : rp.total;
  reduce_plan_kernel<<<static_cast<unsigned>((work + 255) / 256), 256, 0, stream>>>(rp, e.items, e.elems);
  // once per configuration: later calls may come on other streams
  if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(e.items);
    cudaFree(e.elems);
    return nullptr;
  }
  g_plan_bytes += ib + eb;
  g_plans.push_back(e);
  return &g_plans.back();
}

bool lean_enabled() {
  static const bool on = [] { const char* e = getenv("INF_REDUCE_LEAN"); return e == nullptr || atoi(e) != 0; }();
  return on;
}

}  // namespace

cudaError_t launch_field_reduce(const ReduceParams& rp_in, cudaStream_t stream) {
  ReduceParams rp = rp_in;
  const int threads = 256;
  static int resident = 0;   // CTAs of one wave (both kernels: 256 threads, 4 CTAs per SM)
  if (resident == 0) {
    int dev = 0, nsm = 0, per_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || nsm < 1)
      nsm = 132;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, field_reduce_lean_kernel, threads, 0) != cudaSuccess || per_sm < 1) per_sm = 3;
    int per_sm2 = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2, field_reduce_kernel, threads, 0) != cudaSuccess || per_sm2 < 1) per_sm2 = 3;
    resident = nsm * (per_sm < per_sm2 ? per_sm : per_sm2);
  }
  const bool exchange = rp.comm.world > 1 && rp.do_reduce;
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(threads);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  if (exchange) {   // a block spins on its peers' data, which needs every block of every rank to run: all blocks resident
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  } else {
    add_pdl_attr(cfg, attr);
  }
  if (rp.do_reduce && lean_enabled()) {
    int64_t blocks = plan_slots(rp, threads, resident, true);
    const PlanEntry* pe = plan_get(rp, stream);
    if (pe != nullptr) {
      if (blocks > resident) blocks = resident;   // grid-stride beyond one wave
      LeanParams lp{};
      lp.items = pe->items;
      lp.elems = pe->elems;
      lp.nslots = rp.nitems;
      lp.total = rp.total;
      lp.out_bias_off = rp.seg[rp.nseg - 1].off;
      lp.partw = rp.partw;
      lp.part1 = rp.part1;
      lp.flat = rp.flat;
      lp.loss_sums = rp.loss_sums;
      lp.ncta1 = rp.ncta1;
      lp.scale = rp.scale;
      lp.loss_scale = rp.loss_scale;
      lp.col_exp = rp.col_exp;
      static const int dbg = getenv("INF_REDUCE_DBG") ? atoi(getenv("INF_REDUCE_DBG")) : 0;   // measurement only
      lp.pdl_early = (rp.pdl_early ? 1 : 0) | (dbg << 1);
      lp.comm = rp.comm;
      lp.err = rp.err;
      lp.trace = rp.trace;
      lp.adam = rp.adam;
      cfg.gridDim = dim3(static_cast<unsigned>(blocks));
      return cudaLaunchKernelEx(&cfg, field_reduce_lean_kernel, lp);
    }
  }
  if (exchange) return cudaErrorNotSupported;   // the exchange protocol lives in the table-driven kernel (every rank must take it)
  int64_t blocks = plan_slots(rp, threads, resident, false);
  if (blocks > resident) blocks = resident;   // grid-stride beyond one wave
  cfg.gridDim = dim3(static_cast<unsigned>(blocks));
  return cudaLaunchKernelEx(&cfg, field_reduce_kernel, rp);
}

}  // namespace inf
