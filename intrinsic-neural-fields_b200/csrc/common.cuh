// Shared geometry of the field kernels: tile sizes, operand-image ("blob") layout in HBM,
// activation stash layout, partial-gradient workspace layout.  Host + device.
#pragma once
#include <stdint.h>

#include "../../include/inf_field.h"

namespace inf {

constexpr int HID = 128;         // mlp_hidden_dim (compiled set)
constexpr int TILE_M = 128;      // rays per tile = UMMA M = TMEM lanes
constexpr int KC = 64;           // k elements per chunk: 64 x fp16 = one 128-byte swizzle row
constexpr int PANEL_BYTES = TILE_M * KC * 2;  // 16 KB: [128 rows][64 fp16], SWIZZLE_128B
constexpr int BLOB_BYTES = 2 * PANEL_BYTES;   // 32 KB: [256 x 64] (GEMM1 weight chunk) or
                                              //        [128 x 128] as two k-panels (hidden operand)
constexpr int OUT_DIM = 3;       // RGB_COLOR_DIM, model.py:9
constexpr float DZ_SCALE = 32.f; // backward signals are carried x32 in fp16 (range centring)
constexpr int MAX_CTAS = 160;    // sizing bound for per-CTA partials (B200: 148 SMs)
constexpr int K1_PART_FLOATS = 1536;  // per-CTA partials of the forward/backward kernel
constexpr int WG_COLS = 512;     // TMEM columns = accumulator columns of a wgrad CTA

// Block shape shared by the forward/backward kernel and the wgrad kernel.
constexpr int EPI_WARPS = 4;     // warps 0..3: TMEM epilogue (warp w owns TMEM lanes 32w..32w+31)
constexpr int PROD_WARPS = 8;    // warps 4..11: gather producers
constexpr int MMA_WARP = 12;     // warp 12: tcgen05.mma issuer + TMEM allocator
constexpr int LOAD_WARP = 13;    // warp 13: bulk-copy (TMA engine) loader
constexpr int NUM_THREADS = 14 * 32;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int PROD_THREADS = PROD_WARPS * 32;

struct Geo {
  int k, k_pad, nchunks;  // nchunks = k_pad / 64
  int L, skip;
  int nbig;               // wgrad slices over k_pad, 256 columns each
  int nhid;               // wgrad slices over the hidden layers, 4 layers each
  int nslots;             // stash blobs per tile: dz_0..dz_{L-2}, h_0..h_{L-3}
  int nblobs_fwd, nblobs_train;
};

__host__ __device__ inline Geo make_geo(int k, int k_pad, int L, int skip) {
  Geo g;
  g.k = k; g.k_pad = k_pad; g.nchunks = k_pad / KC;
  g.L = L; g.skip = skip;
  g.nbig = (k_pad + 255) / 256;
  g.nhid = (L - 2 + 3) / 4;
  g.nslots = 2 * L - 3;
  g.nblobs_fwd = g.nchunks + (L - 2);
  g.nblobs_train = g.nchunks + 2 * (L - 2);
  return g;
}

// ---- packed operand images (inf_pack_weights) -------------------------------------------------
// blob b (32 KB each), in the order the forward/backward kernel consumes them per tile:
//   [0, nchunks)                 chunk c of [W0;Wy]: rows 0..127 = W0[:, 64c..], 128..255 = Wy
//   nchunks + (l-1), l=1..L-2    W_l  (Lx for the skip layer): [n][k]
//   nchunks + (L-2) + i          W_l^T for l = L-2-i (i = 0..L-3): [j][n]
// followed by an f32 section: bias[L-1][128] (skip row = Lx.bias + Ly.bias), wlast[3][128],
// blast[4].
__host__ __device__ inline size_t packed_f32_offset(const Geo& g) {
  return static_cast<size_t>(g.nblobs_train) * BLOB_BYTES;
}
__host__ __device__ inline size_t packed_total_bytes(const Geo& g) {
  return packed_f32_offset(g) + static_cast<size_t>((g.L - 1) * HID + OUT_DIM * HID + 4) * 4;
}
// Byte offset of element (row r, 16-bit column c) inside a [rows x 64] SWIZZLE_128B panel.
__host__ __device__ inline uint32_t sw128_off(uint32_t r, uint32_t c) {
  return r * 128u + ((((c >> 3) ^ (r & 7u)) << 4) | ((c & 7u) << 1));
}

// ---- activation stash (forward/backward kernel -> wgrad kernel) --------------------------------
// per tile `nslots` blobs: slot l = dz_l (l = 0..L-2), slot (L-1)+l = h_l (l = 0..L-3); each blob
// is the smem image [128 rays][128 features] as two k-panels.
__host__ __device__ inline int stash_slot_dz(const Geo& g, int l) { (void)g; return l; }
__host__ __device__ inline int stash_slot_h(const Geo& g, int l) { return (g.L - 1) + l; }

struct WorkspaceLayout {
  size_t stash_off, partw_off, part1_off, err_off, total;
};
__host__ inline WorkspaceLayout workspace_layout(const Geo& g, int64_t n_rays, bool training) {
  WorkspaceLayout w{};
  size_t off = 0;
  w.err_off = off; off += 256;
  if (training) {
    const int64_t ntiles = (n_rays + TILE_M - 1) / TILE_M;
    w.stash_off = off; off += static_cast<size_t>(ntiles) * g.nslots * BLOB_BYTES;
    w.partw_off = off; off += static_cast<size_t>(MAX_CTAS) * WG_COLS * HID * 4;
    w.part1_off = off; off += static_cast<size_t>(MAX_CTAS) * K1_PART_FLOATS * 4;
  }
  w.total = off;
  return w;
}

// ---- kernel parameter blocks --------------------------------------------------------------------
struct FieldParams {
  const void* table;
  int64_t row_stride;  // elements
  const void* vids;
  const float* bary;
  const float* target;
  const float* grad_pred;
  float* pred;
  const uint8_t* packed;
  uint8_t* stash;
  float* part1;
  float* partw;
  unsigned int* err;
  int64_t n_rays;
  int ntiles;
  int idx64;
  int dense;
  float ext_scale;  // multiplies grad_pred before it enters the fp16 backward chain
  int prefetch;     // L2 prefetch distance of the gather (k-chunks / units ahead); 0 = off
  int groups_big;   // wgrad: ray groups (CTAs) per big slice
  int groups_hid;   // wgrad: ray groups (CTAs) per hidden slice
  Geo geo;
};

struct PackParams {
  const float* w[INF_MAX_LAYERS];  // layer weight (Lx.weight for the skip layer)
  const float* b[INF_MAX_LAYERS];  // layer bias   (Lx.bias   for the skip layer)
  const float* wy;                 // Ly.weight (128 x k)
  const float* by;                 // Ly.bias
  uint8_t* packed;
  Geo geo;
};

struct ReduceSeg {
  int64_t off;   // first element in the flat gradient vector
  int32_t rows, cols;
  int32_t kind;  // 0 big W (a: 0=W0, 1=Wy)  1 hidden W (a: layer)  2 bias (a: layer)  3 last W  4 last b
  int32_t a;
};
struct ReduceParams {
  ReduceSeg seg[INF_MAX_PARAMS];
  int nseg;
  int64_t total;
  const float* partw;
  const float* part1;
  float* flat;
  float* loss_sums;
  int ncta1;   // CTAs of the forward/backward kernel
  int groups_big, groups_hid;  // ray groups per wgrad slice (big / hidden)
  float scale;
  Geo geo;
};

}  // namespace inf
