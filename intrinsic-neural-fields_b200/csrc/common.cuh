// Shared geometry of the field kernels: tile sizes, operand-image ("blob") layout in HBM,
// activation stash layout, wgrad work plan, partial-gradient workspace layout.  Host + device.
#pragma once
#include <stdint.h>

#include "../../include/inf_field.h"

namespace inf {

constexpr int HID = 128;         // mlp_hidden_dim (compiled set); narrower view layers are zero-padded to it
constexpr int TILE_M = 128;      // rays per tile = UMMA M = TMEM lanes
constexpr int KC = 64;           // k elements per chunk: 64 x fp16 = one 128-byte swizzle row
constexpr int PANEL_BYTES = TILE_M * KC * 2;  // 16 KB: [128 rows][64 fp16], SWIZZLE_128B
constexpr int BLOB_BYTES = 2 * PANEL_BYTES;   // 32 KB: [256 x 64] (GEMM1 weight chunk) or
                                              //        [128 x 128] as two k-panels (hidden operand)
constexpr int HALF_PANEL = 64 * 128;          // 8 KB: 64 rows of a panel
constexpr int DZLT_BYTES = 4096;              // [16 x 128 rays] fp16, K-major: two [16 x 64] panels
constexpr int OUT_DIM = 3;       // RGB_COLOR_DIM, model.py:9
constexpr float DZ_SCALE = 32.f; // backward signals are carried x32 in fp16 (range centring)
constexpr int MAX_CTAS = 160;    // sizing bound for per-CTA partials of one launch (B200: 148 SMs)
constexpr int MAX_WGRAD_CTAS = 2 * MAX_CTAS;   // the vertex path runs the hidden and the big slices as two launches
constexpr int MAX_MMA_LAYERS = 8;  // MMA layers 0..nb, nb <= 7
constexpr int K1_PART_FLOATS = 8;  // per-CTA partials of the forward/backward kernel: gb[3], pad, loss sums[2]
constexpr int WG_COLS = 512;     // TMEM columns = accumulator columns of a wgrad CTA
constexpr int MAX_UNITS = 16;    // wgrad work units of the hidden slices
constexpr int MAX_HID_SLICES = 4;
constexpr int ERR_AREA_BYTES = 65536;

// ---- vertex-projected path ------------------------------------------------------------------------
// Layer 0 and the skip branch Ly are linear in the table rows, so  x W^T = sum_j bary_j (E[v_j] W^T):
// the K = k_pad contraction runs ONCE per vertex per step (a streaming GEMM over the table's fp16 operand image),
// and a ray gathers three 512-byte rows of the projected table P = E [W0;Wy]^T instead of three 4 KB table rows.
// The backward mirrors it: G[v] = sum over incident (ray, corner) of bary * [dz_0 | dz_skip][ray], then
// [dW0;dWy] = G^T E, again a streaming GEMM with K = vertices.
constexpr int PROJ_COLS = 2 * HID;                        // [z0 | zy] columns of a projected row
constexpr int PROJ_ROW_BYTES = PROJ_COLS * 2;             // fp16
constexpr int PROJ_TILE_BYTES = TILE_M * PROJ_ROW_BYTES;  // 64 KB per 128 vertices
// 16-byte chunk q (8 columns, q = 0..31) of projected row v is stored at chunk q ^ (v & 7) of the row: the
// projection kernel's epilogue stages whole rows in shared memory (thread = row) without bank conflicts and
// ships a tile with one linear bulk store.
__host__ __device__ inline uint32_t proj_chunk_off(uint32_t v, uint32_t q) { return (q ^ (v & 7u)) << 4; }
// Vertex-gradient image G, per 128 vertices: [dz_0 part: two 64-feature panels][dz_skip part: two panels], every
// panel [128 vertices][64 features] SWIZZLE_128B -- the layout of two stashed dz blobs, so the wgrad kernel reads it
// exactly as it reads a ray tile's dz_0 / dz_skip.
constexpr int GIMG_TILE_BYTES = 2 * BLOB_BYTES;

// Model geometry.  MMA "layers" are numbered 0..nb:
//   layer 0            x (k_pad) -> 128, fused with the skip branch Ly in GEMM1
//   layers 1..nhid     128 -> 128 ReLU layers.  nhid = L-2, or L-1 with the view branch: its spatial MLP ends
//                      in a ReLU bottleneck (model.py:92,156-157), zero-padded to 128 outputs
//   layer nb = nhid+1  view branch only: [bottleneck | Fourier(view dir)] (128 + 64 padded) -> dh (padded to 128),
//                      ReLU (model.py:168-170)
// The output layer (-> 3, sigmoid) runs on CUDA cores from the registers of layer nb's epilogue.
struct Geo {
  int k, k_pad, nchunks;  // nchunks = k_pad / 64
  int L, skip;
  int view;               // INF_VIEW_*
  int emb, include;       // Fourier bands per input dim, include raw input
  int vdim;               // d: 1 (intrinsic: angle) or 3 (extrinsic: direction)
  int edim;               // E = 2*d*emb (+ d) <= 64
  int bneck, dh;          // true widths (<= 128) of the bottleneck / directional hidden layer
  int nhid;               // number of 128->128 MMA layers
  int nb;                 // index of the last MMA layer
  int wout;               // true input width of the output layer (128 or dh)
  int nbig;               // wgrad slices over k_pad, 256 columns each
  int nblobs;             // packed operand blobs: nchunks + nhid (+2 view)
  int stash_blobs;        // 32 KB stash blobs per tile before the x panels
};

__host__ __device__ inline Geo make_geo(int k, int k_pad, int L, int skip, int view = 0, int emb = 0, int include = 0,
                                        int bneck = 0, int dh = 0) {
  Geo g;
  g.k = k; g.k_pad = k_pad; g.nchunks = k_pad / KC;
  g.L = L; g.skip = skip;
  g.view = view; g.emb = view ? emb : 0; g.include = view ? include : 0;
  g.vdim = (view == INF_VIEW_EXTRINSIC) ? 3 : 1;
  g.edim = view ? g.vdim * emb * 2 + (include ? g.vdim : 0) : 0;
  g.bneck = view ? bneck : 0; g.dh = view ? dh : 0;
  g.nhid = (L - 2) + (view ? 1 : 0);
  g.nb = g.nhid + (view ? 1 : 0);
  g.wout = view ? dh : HID;
  g.nbig = (k_pad + 255) / 256;
  g.nblobs = g.nchunks + g.nhid + (view ? 2 : 0);
  g.stash_blobs = 2 * (g.nb + 1) + 1;
  return g;
}

// ---- packed operand images (inf_pack_weights) -------------------------------------------------
// blob b (32 KB each):
//   [0, nchunks)                 chunk c of [W0;Wy]: rows 0..127 = W0[:, 64c..], 128..255 = Wy  ([256][64])
//   nchunks + (l-1), l=1..nhid   W_l (Lx for the skip layer; the bottleneck layer zero-padded to 128 rows):
//                                [n][j] as two j-panels.  The forward GEMM reads it K-major (K = j), the dgrad
//                                GEMM reads the SAME bytes MN-major (K = n), so no transposed copy exists.
//   view: nchunks + nhid         Wd1[:, :bneck]  zero-padded to [128][128]
//         nchunks + nhid + 1     Wd1[:, bneck:bneck+E] zero-padded to [128][64] (first 16 KB of the blob)
// followed by an f32 section: bias[nb+1][128] (skip row = Lx.bias + Ly.bias; rows zero-padded),
// wlast[3][128] (output weight, zero-padded beyond wout), blast[4].
__host__ __device__ inline size_t packed_f32_offset(const Geo& g) { return static_cast<size_t>(g.nblobs) * BLOB_BYTES; }
__host__ __device__ inline int packed_f32_count(const Geo& g) { return (g.nb + 1) * HID + OUT_DIM * HID + 4; }
__host__ __device__ inline size_t packed_total_bytes(const Geo& g) {
  return packed_f32_offset(g) + static_cast<size_t>(packed_f32_count(g)) * 4;
}
__host__ __device__ inline int blob_hidden(const Geo& g, int l) { return g.nchunks + (l - 1); }  // l = 1..nhid
__host__ __device__ inline int blob_wd1z(const Geo& g) { return g.nchunks + g.nhid; }
__host__ __device__ inline int blob_wd1e(const Geo& g) { return g.nchunks + g.nhid + 1; }

// Byte offset of element (row r, 16-bit column c) inside a [rows x 64] SWIZZLE_128B panel.
__host__ __device__ inline uint32_t sw128_off(uint32_t r, uint32_t c) {
  return r * 128u + ((((c >> 3) ^ (r & 7u)) << 4) | ((c & 7u) << 1));
}

// ---- activation stash (forward/backward kernel -> wgrad kernel) --------------------------------
// per tile: `stash_blobs` blobs of 32 KB followed by nchunks x-panels of 16 KB, all smem images (fp16):
//   blob l                 dz_l, l = 0..nb           [128 rays][128 features], two feature-panels
//   blob (nb+1)+l          h_l,  l = 0..nb           (input of layer l+1; h_nb feeds the output layer)
//   blob 2(nb+1)           misc: [0,16K) Fourier panel e [128 rays][64]; [16K,20K) dzl^T [16][128 rays]
//   then x-panel c         x[:, 64c..64c+64)         [128 rays][64], the interpolated embedding in fp16
__host__ __device__ inline size_t stash_tile_bytes(const Geo& g) {
  return static_cast<size_t>(g.stash_blobs) * BLOB_BYTES + static_cast<size_t>(g.nchunks) * PANEL_BYTES;
}
__host__ __device__ inline size_t stash_off_dz(const Geo& g, int l) { (void)g; return static_cast<size_t>(l) * BLOB_BYTES; }
__host__ __device__ inline size_t stash_off_h(const Geo& g, int l) { return static_cast<size_t>(g.nb + 1 + l) * BLOB_BYTES; }
__host__ __device__ inline size_t stash_off_e(const Geo& g) { return static_cast<size_t>(2 * (g.nb + 1)) * BLOB_BYTES; }
__host__ __device__ inline size_t stash_off_dzlt(const Geo& g) { return stash_off_e(g) + PANEL_BYTES; }
__host__ __device__ inline size_t stash_off_x(const Geo& g, int c) {
  return static_cast<size_t>(g.stash_blobs) * BLOB_BYTES + static_cast<size_t>(c) * PANEL_BYTES;
}

// stash of the vertex-projected path: no x panels (the first-layer weight gradient is G^T E)
__host__ __device__ inline size_t stash_tile_bytes_vp(const Geo& g) { return static_cast<size_t>(g.stash_blobs) * BLOB_BYTES; }

constexpr int SCAN_BLOCK = 1024;   // elements per block of the CSR offset scan
struct WorkspaceLayout {
  size_t stash_off, partw_off, part1_off, err_off, total;
  size_t proj_off, gimg_off, vstart_off, vcursor_off, pairs_off, scan_off;   // vertex-projected path (n_rows > 0)
};
// n_rows > 0 sizes the workspace for BOTH paths (the larger stash of the per-ray path plus the vertex buffers), so the
// path can be chosen per call; n_rows == 0: per-ray gather path only.
__host__ inline WorkspaceLayout workspace_layout(const Geo& g, int64_t n_rays, bool training, int64_t n_rows = 0) {
  WorkspaceLayout w{};
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t at = off; off += (bytes + 1023) & ~static_cast<size_t>(1023); return at; };
  w.err_off = take(ERR_AREA_BYTES);   // error word, then the optional timing trace (INF_DEBUG & 8)
  const int64_t ntiles = (n_rays + TILE_M - 1) / TILE_M;
  if (training) {
    w.stash_off = take(static_cast<size_t>(ntiles) * stash_tile_bytes(g));
    w.partw_off = take(static_cast<size_t>(n_rows > 0 ? MAX_WGRAD_CTAS : MAX_CTAS) * WG_COLS * HID * 4);
    w.part1_off = take(static_cast<size_t>(2 * MAX_CTAS) * K1_PART_FLOATS * 4);
  }
  if (n_rows > 0) {
    const int64_t ntv = (n_rows + TILE_M - 1) / TILE_M;
    w.proj_off = take(static_cast<size_t>(ntv) * PROJ_TILE_BYTES);
    if (training) {
      w.gimg_off = take(static_cast<size_t>(ntv) * GIMG_TILE_BYTES);
      w.vstart_off = take(static_cast<size_t>(n_rows + 1) * 4);
      w.vcursor_off = take(static_cast<size_t>(n_rows + 1) * 4);
      w.pairs_off = take(static_cast<size_t>(ntiles) * TILE_M * 3 * 8);   // (ray, bary) per incident pair
      w.scan_off = take(static_cast<size_t>((n_rows + SCAN_BLOCK) / SCAN_BLOCK + 1) * 4);
    }
  }
  w.total = off;
  return w;
}

// ---- wgrad work plan of the "hidden" slices -------------------------------------------------------
// A unit accumulates, over the rays of a tile, products of one stashed A blob (read MN-major: TMEM lane =
// feature of A) into TMEM columns of its slice:
//   kind UNIT_W     D[n, col + j]      += sum_r A[r, n] * B[r, j],  B = stash blob/panel, j < ncols (MN-major)
//   kind UNIT_LAST  D[j, col + c]      += sum_r A[r, j] * dzl^T[c, r]            (A = h_nb, B K-major, N = 16)
//   bias_col >= 0   D[n, bias_col + *] += sum_r A[r, n] * 1                       (B = ones tile, N = 16)
enum : int16_t { UNIT_W = 0, UNIT_LAST = 1, UNIT_BIAS_ONLY = 2 };
struct WUnit {
  int32_t a_off;      // byte offset of the A blob inside the tile's stash
  int32_t b_off;      // byte offset of the B operand inside the tile's stash
  int32_t b_bytes;    // bytes to load for B (0 for UNIT_BIAS_ONLY)
  int16_t kind, ncols, col, bias_col, slice, layer;
};
struct WPlan {
  WUnit u[MAX_UNITS];
  int n;                        // units
  int nslice;                   // hidden slices
  int cols_used[MAX_HID_SLICES];
  int bytes_per_tile[MAX_HID_SLICES];
  int unit_w[MAX_MMA_LAYERS];   // unit accumulating the weight gradient of MMA layer l (1..nb)
  int unit_e;                   // view: unit of dWd1[:, bneck:]
  int unit_bias[MAX_MMA_LAYERS];// unit whose bias_col holds the bias gradient of layer l (0..nb)
  int unit_last;                // output-layer weight gradient
};
__host__ inline WPlan make_wplan(const Geo& g) {
  WPlan p{};
  int slice = 0, col = 0;
  auto place = [&](int need) {
    if (col + need > WG_COLS) { ++slice; col = 0; }
  };
  auto add = [&](int kind, size_t a_off, size_t b_off, int b_bytes, int ncols, bool bias, int layer) -> int {
    place(ncols + (bias ? 16 : 0));
    WUnit w{};
    w.a_off = static_cast<int32_t>(a_off); w.b_off = static_cast<int32_t>(b_off); w.b_bytes = b_bytes;
    w.kind = static_cast<int16_t>(kind); w.ncols = static_cast<int16_t>(ncols); w.col = static_cast<int16_t>(col);
    col += ncols;
    w.bias_col = -1;
    if (bias) { w.bias_col = static_cast<int16_t>(col); col += 16; }
    w.slice = static_cast<int16_t>(slice); w.layer = static_cast<int16_t>(layer);
    p.u[p.n] = w;
    p.cols_used[slice] = col;
    p.bytes_per_tile[slice] += BLOB_BYTES + b_bytes;
    return p.n++;
  };
  p.unit_e = -1;
  for (int l = 1; l <= g.nhid; ++l) {
    p.unit_w[l] = add(UNIT_W, stash_off_dz(g, l), stash_off_h(g, l - 1), BLOB_BYTES, HID, true, l);
    p.unit_bias[l] = p.unit_w[l];
  }
  if (g.view) {
    p.unit_w[g.nb] = add(UNIT_W, stash_off_dz(g, g.nb), stash_off_h(g, g.nhid), BLOB_BYTES, HID, true, g.nb);
    p.unit_bias[g.nb] = p.unit_w[g.nb];
    p.unit_e = add(UNIT_W, stash_off_dz(g, g.nb), stash_off_e(g), PANEL_BYTES, KC, false, g.nb);
  }
  p.unit_bias[0] = add(UNIT_BIAS_ONLY, stash_off_dz(g, 0), 0, 0, 0, true, 0);
  p.unit_last = add(UNIT_LAST, stash_off_h(g, g.nb), stash_off_dzlt(g), DZLT_BYTES, 16, false, g.nb);
  p.nslice = slice + 1;
  return p;
}
constexpr int BIG_BYTES_PER_TILE = 4 * BLOB_BYTES;  // dz_0 + dz_skip + 256 x-columns, per 128 rays

// ---- launch helper: programmatic dependent launch (INF_PDL=0 turns it off) -----------------------------------------
// The kernels of a step form a chain on one stream; with this attribute the next kernel's CTAs are scheduled while the
// previous kernel drains and run their prologue (barrier init, TMEM allocation, address arithmetic) up to pdl_wait().
// Measured on the bench step (scripts/pdl_ab.sh) and left OFF: with the early trigger the training step went 0.333 -> 0.352 ms
// and inference 0.091 -> 0.095 ms (INF_PDL=1); INF_PDL=2 sets the attribute without the early trigger.
inline int pdl_mode() {
  static const int m = [] { const char* e = getenv("INF_PDL"); return e == nullptr ? 0 : atoi(e); }();
  return m;
}
inline bool pdl_enabled() { return pdl_mode() != 0; }
constexpr int DEBUG_PDL_EARLY = 1 << 16;   // FieldParams.debug bit: kernels call griddepcontrol.launch_dependents in their prologue
inline void add_pdl_attr(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr) {
  if (!pdl_enabled()) return;
  attr[cfg.numAttrs].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[cfg.numAttrs].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  ++cfg.numAttrs;
}

// ---- kernel parameter blocks --------------------------------------------------------------------
struct FieldParams {
  const void* table;
  int64_t row_stride;  // elements
  const void* vids;
  const float* bary;
  const float* target;
  const float* grad_pred;
  const float* dirs;          // (B,3) unit ray directions (view branch)
  const void* face_idxs;      // (B,)  hit face ids (intrinsic view strategy)
  const float* face_normals;  // (F,3)
  const void* ridx;           // optional (B,) batch selection into the per-ray arrays
  int ridx64;
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
  unsigned long long* trace;  // (INF_DEBUG & 8) per-role cycle counters of CTA 0, 32 slots (scripts/trace_chain.py)
  int prefetch;     // gather L2 prefetch distance in k-chunks (0 = off)
  int cluster;      // CTAs per cluster of the forward/backward kernel (1 or 2): [W0;Wy] chunks are multicast within it
  int wg_cluster;   // CTAs per cluster of the wgrad kernel (1, 2 or 4): dz_0 / dz_skip half-panels are multicast within it
  int debug;        // measurement knobs (INF_DEBUG env): 1 no x stash, 2 no h/dz stash, 4 no dgrad chain -- wrong results
  // vertex-projected path (e_image == nullptr: per-ray gather of table rows)
  const uint8_t* e_image;           // fp16 operand image of the table: [ntiles_v][nchunks] panels [128 rows][64] SWIZZLE_128B
  uint8_t* proj;                    // P [ntiles_v * 128][256] fp16 (chunks permuted, proj_chunk_off)
  uint8_t* g_image;                 // G [ntiles_v][GIMG_TILE_BYTES]
  int32_t* vstart;                  // [n_rows + 1] CSR offsets: incident (ray, corner) pairs of every vertex
  int32_t* vcursor;                 // [n_rows + 1] counts, then fill cursors
  int32_t* pairs;                   // [3 * n_rays] int2 (ray, bits of the barycentric weight), grouped by vertex
  int32_t* scan_tmp;                // block sums of the offset scan
  int64_t n_rows;                   // V
  int ntiles_v;                     // ceil(V / 128)
  size_t stash_stride;              // bytes per ray tile of the activation stash
  int cta_base;                     // wgrad: index of this launch's first CTA in the combined (big slices, hidden slices) CTA space
  int groups_big;                   // wgrad: ray groups (CTAs) per big slice
  int groups_hid[MAX_HID_SLICES];   // wgrad: ray groups (CTAs) per hidden slice
  Geo geo;
  WPlan plan;
};

struct PackParams {
  const float* w[MAX_MMA_LAYERS];  // weight of MMA layer l (Lx.weight for the skip layer; Wd1 for the view layer)
  const float* b[MAX_MMA_LAYERS];  // bias of MMA layer l
  const float* wy;                 // Ly.weight (128 x k)
  const float* by;                 // Ly.bias
  const float* wlast;              // output layer weight (3 x wout)
  const float* blast;              // output layer bias (3)
  const int8_t* col_exp;           // optional column exponents of the table: W0 / Wy images hold W[:,j] * 2^-col_exp[j]
  uint8_t* packed;
  Geo geo;
};

// One parameter tensor of the flat gradient vector.
//   kind 0  W0 / Wy (a = 0/1): element (n, j) = big slice j/256, lane n, column (j%256) + 256a
//   kind 1  weight from hidden unit(s): (n, j) = unit a column j for j < split, unit b column j-split beyond
//   kind 2  bias: (n) = bias column of unit a, lane n
//   kind 3  output weight: (c, j) = unit a, lane j, column c
//   kind 4  output bias from the forward/backward kernel's per-CTA partials
struct ReduceSeg {
  int64_t off;   // first element in the flat gradient vector
  int32_t rows, cols;
  int32_t kind;
  int32_t a, b, split;
  int32_t layer;      // MMA layer of the tensor (kinds 0..2); -1 otherwise
  int64_t pair_off;   // kind 2, skip layer: offset of the partner bias (Lx.bias owns the update of both, its image entry
                      // is their sum); -2 marks the partner itself (Ly.bias: nothing to do here); -1 otherwise
  int64_t item_off;   // first work item of the tensor in the reduction kernel (an item = up to 4 consecutive columns of a
  int32_t ipr;        // row; ipr items per row) -- filled by launch_field_reduce
  int32_t par;        // lanes sharing an item (1, 2 or 4): they split the item's ray / vertex groups
  int32_t cpi;        // columns per item: 4, or 1 for the narrow tensors when the table-driven kernel is used
};
// Optional optimiser tail of the reduction kernel: Adam on the flat fp32 master parameters + refresh of the fp16
// operand images, fused with the partial sums (config.py:108 torch.optim.Adam(lr), defaults).
struct AdamParams {
  float* params;      // flat master parameters, state_dict order (null = no optimiser step)
  float* exp_avg;
  float* exp_avg_sq;
  uint8_t* packed;    // operand images to refresh
  const int8_t* col_exp;     // column exponents of the table (W0 / Wy images hold W[:,j] * 2^-col_exp[j]); null = none
  int32_t* step_dev;        // optional device-resident step count t (CUDA-graph replays); else the host values below
  int step_autoinc;         // step_dev counts COMPLETED updates: t = *step_dev + 1, written back by the last block to leave
  unsigned int* done_ctr;   // block counter for that (a zero-initialised word of the workspace)
  float step_size;    // lr / (1 - beta1^t)
  float bc2_sqrt;     // sqrt(1 - beta2^t)
  float lr, beta1, beta2, eps;
  float omb1, omb2;   // 1 - beta, rounded from double as torch does
};
// Peer-memory exchange of the flat gradient (multi-GPU step): every rank owns an exchange region that all
// ranks of the node have mapped (CUDA IPC over NVLink).  A region = 2 x MAX_PEERS gradient slots -- slot (parity, q) receives
// rank q's gradient of a step of that parity, PUSHED by rank q with posted remote stores, so that the sum over ranks reads
// local memory only; steps alternate between the two parities, so a rank may run one step ahead of a peer that is still
// reading -- followed by a control block.  Every element travels as an 8-byte (value, step number) pair written with one
// store: the reader spins on the element itself until it carries this step's number, so there is no fence, no flag round
// and no grid-wide barrier between a rank's sums and the sum over ranks (the exchange costs one NVLink store latency).
constexpr int MAX_PEERS = 8;
constexpr int COMM_CTL_WORDS = 64;       // u32 words of the control block
enum : int { CTL_SEQ = 16,               // steps this rank has completed (device-resident: a CUDA-graph replay advances it)
             CTL_DONE = 18 };            // block counter of the reduction kernel (the last block to leave advances CTL_SEQ)
__host__ __device__ inline int64_t comm_slot_floats(int64_t total) { return 2 * (((total + 2 + 31) / 32) * 32); }
// float offset of slot (parity, source rank) / of the control block inside a region
// slots per parity: one per source rank (contributions) + one for the finished sums of the two-shot exchange
constexpr int COMM_SLOTS = MAX_PEERS + 1;
constexpr int COMM_SLOT_RESULT = MAX_PEERS;
__host__ __device__ inline int64_t comm_slot_off(int64_t slot_floats, uint32_t parity, int src) {
  return (static_cast<int64_t>(parity) * COMM_SLOTS + src) * slot_floats;
}
__host__ __device__ inline int64_t comm_ctl_off(int64_t slot_floats) { return 2 * COMM_SLOTS * slot_floats; }
struct CommParams {
  int world, rank;                       // world <= 1: no exchange
  float* region[MAX_PEERS];              // region[q] = rank q's exchange region as mapped in this process
  int64_t slot_floats;                   // floats per gradient slot: 2 per element (value, step number), total + 2 elements, padded
  int64_t own_chunk;                     // two-shot exchange (world >= 4): rank i / own_chunk sums element i and sends the sum to
                                         // everyone (each rank moves 2 x 7/8 of the vector instead of 7 x); 0 = one shot
};
__host__ __device__ inline size_t comm_region_bytes(int64_t total) {
  return static_cast<size_t>(2 * COMM_SLOTS * comm_slot_floats(total)) * 4 + COMM_CTL_WORDS * 4;
}

// Tables of the table-driven reduction kernel (field_reduce_lean_kernel), generated on the device by reduce_plan_kernel once
// per (model geometry, wgrad CTA split) and kept by the library: what the generic kernel derives per thread from the parameter
// block with run-time-indexed constant loads and a lot of code, the lean kernel reads with one coalesced load.
struct ItemDesc {        // one per work slot (item x lane)
  uint32_t src;          // float offset into the wgrad partials of (first group of this lane, first column)
  uint32_t gstride;      // floats between consecutive groups OF THIS LANE (group stride x lanes per item)
  uint32_t i0;           // flat index of the item's first element
  uint16_t ng;           // groups this lane adds up
  uint8_t nc;            // columns (1..4); 0 = padding slot
  uint8_t flags;         // bit 0: 16-byte loads; bits 2-3: log2(lanes per item); bits 4-5: lane of the item
};
struct ElemDesc {        // one per flat element: where the optimiser tail writes the refreshed operand
  uint32_t poff;         // byte offset into the packed operand buffer
  uint32_t aux;          // bits 0-1: 0 fp16 image entry, 1 f32 entry, 2 f32 entry = this + partner (skip-layer biases), 3 nothing
                         // bit 2: first-layer / skip-branch weight (bits 3..: its table column, for the column exponents)
                         // type 2: bits 3..: flat index of the partner
};
struct LeanParams {
  const ItemDesc* items;
  const ElemDesc* elems;
  int64_t nslots, total;
  int64_t out_bias_off;   // flat index of the output-layer bias (finished by block 0)
  const float* partw;
  const float* part1;
  float* flat;
  float* loss_sums;
  int ncta1;
  float scale;
  float loss_scale;       // loss_sums[2] = loss_sums[0] * loss_scale (desc.grad_scale: the mean L1 loss)
  const int8_t* col_exp;
  int pdl_early;
  CommParams comm;
  unsigned int* err;
  unsigned long long* trace;
  AdamParams adam;
};

struct ReduceParams {
  ReduceSeg seg[INF_MAX_PARAMS];
  int nseg;
  int64_t total;
  int64_t nitems;   // work slots (items x lanes per item, padded per tensor) of the reduction kernel (launch_field_reduce)
  const float* partw;
  const float* part1;
  float* flat;
  float* loss_sums;
  int ncta1;   // CTAs of the forward/backward kernel
  int groups_big;
  int groups_hid[MAX_HID_SLICES];
  int cta_base_hid[MAX_HID_SLICES];
  float scale;
  float loss_scale;        // loss_sums[2] = loss_sums[0] * loss_scale (desc.grad_scale)
  const int8_t* col_exp;   // column exponents of the table: the W0 / Wy gradients come out multiplied by 2^col_exp[j]
  int pdl_early;   // call griddepcontrol.launch_dependents in the prologue (INF_PDL=1)
  int do_reduce;   // 0: flat already holds the (all-reduced) gradient, only the optimiser tail runs
  CommParams comm; // multi-GPU: exchange + sum of the per-rank gradients inside this kernel (world <= 1: none)
  unsigned int* err;
  unsigned long long* trace;   // (INF_DEBUG & 8) globaltimer stamps of the exchange: start, published, peers seen, end
  AdamParams adam;
  Geo geo;
  WPlan plan;
};

}  // namespace inf
