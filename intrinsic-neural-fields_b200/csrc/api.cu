// C ABI of libinf_field.so (see include/inf_field.h).  Validation, workspace carving, launches.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <mutex>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>

#include "common.cuh"

namespace inf {

// kernels / launchers defined in the other translation units
cudaError_t launch_field_main(const FieldParams& p, bool train, bool table_f16, int grid, cudaStream_t stream);
int field_main_max_grid(bool train, bool table_f16, int cluster, bool view);
cudaError_t launch_field_wgrad(const FieldParams& p, int grid, cudaStream_t stream);
int field_wgrad_max_grid(int cluster);
cudaError_t launch_field_reduce(const ReduceParams& rp, cudaStream_t stream);

cudaError_t launch_vertex_project(const FieldParams& p, int grid, int cluster, cudaStream_t stream);
int vertex_project_max_grid(int cluster);
void launch_table_image(const void* table, int table_f16, int64_t n_rows, int k_pad, void* image, cudaStream_t stream);
void launch_proj_unpermute(const uint8_t* proj, int64_t n_rows, void* out, cudaStream_t stream);
cudaError_t launch_field_rays(const FieldParams& p, bool train, int slots, int grid, cudaStream_t stream);
cudaError_t launch_csr_build(const FieldParams& p, cudaStream_t stream);
cudaError_t launch_vertex_gather(const FieldParams& p, int nsm, cudaStream_t stream);

void launch_pack_weights(const PackParams& pp, cudaStream_t stream);
void launch_relayout(const float* src, int64_t n_rows, int k, int64_t src_stride, void* dst, int k_pad, int dst_dtype,
                     const int8_t* col_exp, cudaStream_t stream);
void launch_column_exponents(const float* src, int64_t n_rows, int k, int64_t src_stride, int k_pad, int8_t* col_exp,
                             unsigned int* scratch, cudaStream_t stream);
void launch_gather_interpolate(const void* table, int table_f16, int k, int k_pad, const void* vids, int idx64,
                               const float* bary, int64_t n_rays, float* out, const int8_t* col_exp, cudaStream_t stream);
void launch_umma_probe(int mode, int n, int k, const float* a, const float* b, float* d, cudaStream_t stream);

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};   // kernels of this library enqueued so far (inf_launch_count)

// Environment knobs are read ONCE per process (not on every launch).  INF_DEBUG bits 1|2|4 are measurement switches
// that skip work and produce WRONG gradients: bench.py refuses to run with them and records the value.
struct Env {
  int debug, prefetch, cluster, verbose, wg_cluster, force_path, rays_slots, side_stream, wgrad_hid_ctas, wgrad_split, wgrad_min_tiles;
  double wgrad_bigw;
  Env() {
    auto geti = [](const char* n, int d) { const char* e = getenv(n); return e ? atoi(e) : d; };
    debug = geti("INF_DEBUG", 0);
    prefetch = geti("INF_PREFETCH", 0);
    cluster = geti("INF_CLUSTER", 2);
    verbose = geti("INF_VERBOSE", 0);
    wg_cluster = geti("INF_WGRAD_CLUSTER", 0);   // 0: 4 on the per-ray path, 2 on the vertex path (see run_forward_backward)
    rays_slots = geti("INF_RAYS_SLOTS", 1) == 2 ? 2 : 1;   // flavour of the per-ray kernel (field_rays.cu)
    // Vertex path, measured and left off: hidden-slice wgrad as its own launch, concurrent with the vertex gather on the
    // side stream, big slices after (0.410 vs 0.394 ms/step: both kernels are bound by the same memory system)
    wgrad_split = geti("INF_WGRAD_SPLIT", 0);
    // measured at B = 4096 (32 tiles): 1 -> wgrad 23 us + reduce 20 us; 4 -> 31 + 15; 8 -> 44 + 15: fewer, longer CTAs lose
    // more in the wgrad kernel than the smaller partial set saves in the reduction, so no cap by default
    wgrad_min_tiles = geti("INF_WGRAD_MIN_TILES", 1);
    if (wgrad_min_tiles < 1) wgrad_min_tiles = 1;
    wgrad_hid_ctas = geti("INF_WGRAD_HID_CTAS", 0);   // vertex path: CTAs of the hidden-slice wgrad launch (0 = all SMs)
    side_stream = geti("INF_SIDE_STREAM", 1);   // 0: incidence lists on the caller's stream
    force_path = geti("INF_PATH", 0);   // 1 = per-ray, 2 = vertex (when a table image is given): overrides InfDesc.path AUTO
    const char* b = getenv("INF_WGRAD_BIGW");
    // (0.85 on the vertex path: wgrad 92 -> 88.8 us, step -1 %, within noise -- and the 2000-step view-branch PSNR protocol,
    // a chaotic map, then lands 0.1-0.4 dB lower although one-step gradients differ by 9e-8: scripts/grad_split_check.py)
    wgrad_bigw = b ? atof(b) : 1.0;
  }
};
const Env& env() {
  static const Env e;
  return e;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_desc(const InfDesc* d) {
  if (d == nullptr) return fail(INF_ERR_BAD_ARG, "desc is null");
  if (d->abi_version != INF_ABI_VERSION)
    return fail(INF_ERR_BAD_ARG, "abi_version %d != %d", d->abi_version, INF_ABI_VERSION);
  if (d->hidden != HID)
    return fail(INF_ERR_UNSUPPORTED, "mlp_hidden_dim=%d outside the compiled set {128}", d->hidden);
  if (d->num_layers < 3 || d->num_layers > INF_MAX_LAYERS)
    return fail(INF_ERR_UNSUPPORTED, "num_layers=%d outside [3,%d]", d->num_layers, INF_MAX_LAYERS);
  if (!(d->skip > 0 && d->skip < d->num_layers - 1))
    return fail(INF_ERR_BAD_ARG, "skip_layer_idx=%d must satisfy 0 < skip < num_layers-1 (model.py:27)", d->skip);
  if (d->k < 1 || d->k_pad < d->k || d->k_pad % KC != 0 || d->k_pad > 16384)
    return fail(INF_ERR_BAD_ARG, "k=%d k_pad=%d: k_pad must be a multiple of 64, >= k, <= 16384", d->k, d->k_pad);
  if (d->n_rays < 0 || d->n_rays > (int64_t(1) << 31) - TILE_M)
    return fail(INF_ERR_BAD_ARG, "n_rays=%lld out of range", static_cast<long long>(d->n_rays));
  if (d->idx_dtype != INF_IDX_I32 && d->idx_dtype != INF_IDX_I64) return fail(INF_ERR_BAD_ARG, "bad idx_dtype");
  if (d->table_dtype != INF_TABLE_F32 && d->table_dtype != INF_TABLE_F16) return fail(INF_ERR_BAD_ARG, "bad table_dtype");
  if (d->input_mode != INF_INPUT_GATHER && d->input_mode != INF_INPUT_DENSE) return fail(INF_ERR_BAD_ARG, "bad input_mode");
  if (d->view_strategy != INF_VIEW_NONE) {
    if (d->view_strategy != INF_VIEW_INTRINSIC && d->view_strategy != INF_VIEW_EXTRINSIC)
      return fail(INF_ERR_BAD_ARG, "bad view_strategy %d", d->view_strategy);
    if (d->num_layers > MAX_MMA_LAYERS - 1)
      return fail(INF_ERR_UNSUPPORTED, "num_layers=%d with the view branch exceeds %d", d->num_layers, MAX_MMA_LAYERS - 1);
    if (d->view_bneck < 1 || d->view_bneck > HID || d->view_hidden < 1 || d->view_hidden > HID)
      return fail(INF_ERR_UNSUPPORTED, "bottleneck_vec_dim=%d / directional_hidden_dim=%d outside [1,%d]", d->view_bneck,
                  d->view_hidden, HID);
    const int dd = d->view_strategy == INF_VIEW_EXTRINSIC ? 3 : 1;
    const int e = dd * d->view_embed * 2 + (d->view_include ? dd : 0);
    if (d->view_embed < 0 || d->view_embed > 24 || e < 1 || e > KC)
      return fail(INF_ERR_UNSUPPORTED, "view embedding of %d dims (embed_size=%d) outside [1,%d]", e, d->view_embed, KC);
  } else if (d->num_layers - 2 > MAX_MMA_LAYERS - 1) {
    return fail(INF_ERR_UNSUPPORTED, "num_layers=%d outside the compiled set", d->num_layers);
  }
  if (d->input_mode == INF_INPUT_GATHER && d->n_rows > (int64_t(1) << 32) - 1)
    return fail(INF_ERR_BAD_ARG, "n_rows exceeds 2^32-1");
  if (d->path != INF_PATH_AUTO && d->path != INF_PATH_RAY && d->path != INF_PATH_VERTEX) return fail(INF_ERR_BAD_ARG, "bad path %d", d->path);
  return INF_OK;
}

// Which evaluation of the two K = k contractions a call takes (include/inf_field.h, INF_PATH_*).  The vertex path costs
// ~2 V k 256 MACs + one streaming pass over the table image per contraction (forward P, backward G^T E), the per-ray
// path ~B k 256 MACs + 12 KB of gathered rows per ray per contraction: the vertex path wins while V <= ~2 B.
bool vertex_path(const InfDesc* d, const InfTensors* t) {
  if (d->input_mode != INF_INPUT_GATHER || t == nullptr || d->n_rows < 1 || d->n_rays < 1) return false;
  int path = d->path;
  if (path == INF_PATH_AUTO && env().force_path != 0) path = env().force_path;
  if (path == INF_PATH_RAY) return false;
  if (t->proj != nullptr && t->target == nullptr && t->grad_pred == nullptr) return true;   // a ready projection: nothing left per vertex
  if (t->table_image == nullptr) return false;
  if (path == INF_PATH_VERTEX) return true;
  return d->n_rows <= 2 * d->n_rays;
}

Geo geo_of(const InfDesc* d) {
  return make_geo(d->k, d->k_pad, d->num_layers, d->skip, d->view_strategy, d->view_embed, d->view_include, d->view_bneck,
                  d->view_hidden);
}

struct ParamInfo {
  int64_t off;
  int rows, cols;
  int kind, a, b, split;  // ReduceSeg fields
  int layer;
  int64_t pair_off;
};

// state_dict order (model.py:42-96, layers.py:50-58; with the view branch the spatial MLP comes first under
// "spatial_mlp." followed by directional_mlp.0 / directional_mlp.2, model.py:134-173)
int build_param_table(const InfDesc* d, ParamInfo* out, int64_t* total) {
  const Geo g = geo_of(d);
  const WPlan plan = make_wplan(g);
  int n = 0;
  int64_t off = 0;
  const int L = d->num_layers;
  auto add = [&](int rows, int cols, int kind, int a, int b = -1, int split = 1 << 30) {
    int layer = -1;
    if (kind == 0) layer = (a == 0) ? 0 : d->skip;
    else if (kind == 1 || kind == 2) layer = plan.u[a].layer;
    out[n++] = ParamInfo{off, rows, cols, kind, a, b, split, layer, -1};
    off += static_cast<int64_t>(rows) * cols;
  };
  for (int i = 0; i < L; ++i) {
    if (i == d->skip) {
      add(HID, HID, 1, plan.unit_w[i]);         // Lx.weight
      add(HID, 1, 2, plan.unit_bias[i]);        // Lx.bias
      add(HID, d->k, 0, 1);                     // Ly.weight
      add(HID, 1, 2, plan.unit_bias[i]);        // Ly.bias (same gradient as Lx.bias)
      out[n - 3].pair_off = out[n - 1].off;     // the optimiser tail updates both biases from the Lx.bias thread
      out[n - 1].pair_off = -2;
    } else if (i == 0) {
      add(HID, d->k, 0, 0);
      add(HID, 1, 2, plan.unit_bias[0]);
    } else if (i == L - 1) {
      if (g.view) {                             // bottleneck layer (ReLU), an MMA layer
        add(g.bneck, HID, 1, plan.unit_w[i]);
        add(g.bneck, 1, 2, plan.unit_bias[i]);
      } else {
        add(OUT_DIM, HID, 3, plan.unit_last);
        add(OUT_DIM, 1, 4, 0);
      }
    } else {
      add(HID, HID, 1, plan.unit_w[i]);
      add(HID, 1, 2, plan.unit_bias[i]);
    }
  }
  if (g.view) {
    add(g.dh, g.bneck + g.edim, 1, plan.unit_w[g.nb], plan.unit_e, g.bneck);   // directional_mlp.0.weight
    add(g.dh, 1, 2, plan.unit_bias[g.nb]);
    add(OUT_DIM, g.dh, 3, plan.unit_last);                                      // directional_mlp.2.weight
    add(OUT_DIM, 1, 4, 0);
  }
  if (total) *total = off;
  return n;
}

// ---- optional per-kernel timing (inf_profile_*): library-owned CUDA events around the launches ----
constexpr int PROF_SLOTS = 256;
constexpr int PROF_EVENTS = INF_PROFILE_STAGES + 1;
struct Profile {
  bool on = false;
  int n = 0;                       // recorded calls since inf_profile_enable(1)
  // start, after incidence lists, after vertex projection, after the per-ray kernel, after the vertex gather,
  // after wgrad, after reduce (stages a call does not run record a zero-length interval)
  cudaEvent_t ev[PROF_SLOTS][PROF_EVENTS];
  bool created = false;
} g_prof;

void prof_record(int slot, int i, cudaStream_t st) {
  if (g_prof.on && slot < PROF_SLOTS) cudaEventRecord(g_prof.ev[slot][i], st);
}

int cuda_fail(cudaError_t e, const char* what) {
  return fail(INF_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

int sm_count(int* out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  int n = 0;
  e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute");
  *out = n < MAX_CTAS ? n : MAX_CTAS;
  return INF_OK;
}

int check_tensors(const InfDesc* d, const InfTensors* t, bool training) {
  if (t == nullptr) return fail(INF_ERR_BAD_ARG, "tensors is null");
  const bool vp = vertex_path(d, t);
  const bool ready = vp && !training && t->proj != nullptr;   // the caller's projection: no table image, no P in the workspace
  if (vp && !ready && (t->table_image == nullptr || (reinterpret_cast<uintptr_t>(t->table_image) & 127) != 0))
    return fail(INF_ERR_BAD_ARG, "table_image must be non-null and 128-byte aligned");
  if (ready && (reinterpret_cast<uintptr_t>(t->proj) & 127) != 0) return fail(INF_ERR_BAD_ARG, "proj must be 128-byte aligned");
  if (training && t->proj != nullptr && t->table_image == nullptr)
    return fail(INF_ERR_BAD_ARG, "proj is an inference input (the weights change every step): training needs table_image");
  if (d->n_rays == 0) return INF_OK;
  if (t->table == nullptr || t->packed == nullptr || t->pred == nullptr)
    return fail(INF_ERR_BAD_ARG, "table / packed / pred must be non-null");
  if ((reinterpret_cast<uintptr_t>(t->table) & 15) != 0) return fail(INF_ERR_BAD_ARG, "table must be 16-byte aligned");
  if ((reinterpret_cast<uintptr_t>(t->packed) & 127) != 0) return fail(INF_ERR_BAD_ARG, "packed must be 128-byte aligned");
  if (d->input_mode == INF_INPUT_GATHER && (t->vids == nullptr || t->bary == nullptr))
    return fail(INF_ERR_BAD_ARG, "vids / bary must be non-null in gather mode");
  if (t->ray_idxs != nullptr) {
    if (d->input_mode != INF_INPUT_GATHER) return fail(INF_ERR_BAD_ARG, "ray_idxs needs input_mode GATHER");
    if (d->ray_idx_dtype != INF_IDX_I32 && d->ray_idx_dtype != INF_IDX_I64) return fail(INF_ERR_BAD_ARG, "bad ray_idx_dtype");
    if (d->n_src_rays < 1) return fail(INF_ERR_BAD_ARG, "n_src_rays must be given with ray_idxs");
  }
  if (d->view_strategy != INF_VIEW_NONE) {
    if (t->unit_ray_dirs == nullptr) return fail(INF_ERR_BAD_ARG, "unit_ray_dirs must be non-null with the view branch");
    if (d->view_strategy == INF_VIEW_INTRINSIC && (t->face_idxs == nullptr || t->face_normals == nullptr))
      return fail(INF_ERR_BAD_ARG, "face_idxs / face_normals must be non-null for the intrinsic view strategy");
  }
  if (training) {
    if ((t->target == nullptr) == (t->grad_pred == nullptr))
      return fail(INF_ERR_BAD_ARG, "exactly one of target / grad_pred must be given");
  }
  if (training || vp) {
    const WorkspaceLayout w = workspace_layout(geo_of(d), d->n_rays, training, (vp && !ready) ? d->n_rows : 0);
    if (t->workspace == nullptr || t->workspace_bytes < w.total)
      return fail(INF_ERR_WORKSPACE, "workspace %zu B < required %zu B", t->workspace_bytes, w.total);
    if ((reinterpret_cast<uintptr_t>(t->workspace) & 255) != 0)
      return fail(INF_ERR_BAD_ARG, "workspace must be 256-byte aligned");
  }
  return INF_OK;
}

// Grid and cluster size of the forward/backward kernel: clusters of 2 (weight chunks multicast) when the call has at
// least two tiles and the device can co-schedule them; INF_CLUSTER=1 forces single CTAs.
int pick_main_grid(const InfDesc* d, int ntiles, int nsm, bool training, int* cluster) {
  static int cached[2][2][2] = {{{-1, -1}, {-1, -1}}, {{-1, -1}, {-1, -1}}};   // [train][f16][view] -> max grid at cluster 2
  int cs = env().cluster;
  if (cs != 2 || ntiles < 2) cs = 1;
  int grid = ntiles < nsm ? ntiles : nsm;
  if (cs == 2) {
    const bool f16 = d->table_dtype == INF_TABLE_F16, view = d->view_strategy != INF_VIEW_NONE;
    int& mg = cached[training ? 1 : 0][f16 ? 1 : 0][view ? 1 : 0];
    if (mg < 0) mg = field_main_max_grid(training, f16, 2, view);
    if (mg < 2) cs = 1;
    else {
      if (grid > mg) grid = mg;
      grid &= ~1;
    }
  }
  *cluster = cs;
  if (env().verbose) fprintf(stderr, "[inf] field_main grid %d cluster %d (ntiles %d)\n", grid, cs, ntiles);
  return grid;
}

FieldParams make_params(const InfDesc* d, const InfTensors* t, bool training) {
  FieldParams p{};
  p.table = t->table;
  p.row_stride = d->k_pad;
  p.vids = t->vids;
  p.bary = t->bary;
  p.target = t->target;
  p.grad_pred = t->grad_pred;
  p.dirs = t->unit_ray_dirs;
  p.face_idxs = t->face_idxs;
  p.face_normals = t->face_normals;
  p.ridx = (d->input_mode == INF_INPUT_GATHER) ? t->ray_idxs : nullptr;
  p.ridx64 = d->ray_idx_dtype == INF_IDX_I64;
  p.pred = t->pred;
  p.packed = reinterpret_cast<const uint8_t*>(t->packed);
  p.n_rays = d->n_rays;
  p.ntiles = static_cast<int>((d->n_rays + TILE_M - 1) / TILE_M);
  p.idx64 = d->idx_dtype == INF_IDX_I64;
  p.dense = d->input_mode == INF_INPUT_DENSE;
  p.ext_scale = 1.f;
  {
    p.debug = env().debug;   // measurement only; bits 1|2|4 produce wrong gradients
    // L2 prefetch of the next row pieces through the bulk-copy engine (distance in k-chunks).  Off by default: it paid
    // while the gather was throttled by a 28 KB L1 (227 KB shared-memory carve-out); with the 164 KB carve-out the plain
    // loads run at the memory system's rate and the extra requests cost 3-13 % at every table size.
    p.prefetch = env().prefetch;
    if (p.prefetch < 0 || p.prefetch > 8 || d->input_mode == INF_INPUT_DENSE) p.prefetch = 0;
  }
  p.geo = geo_of(d);
  p.plan = make_wplan(p.geo);
  const bool vp = vertex_path(d, t);
  p.stash_stride = vp ? stash_tile_bytes_vp(p.geo) : stash_tile_bytes(p.geo);
  p.n_rows = d->n_rows;
  if (t->workspace != nullptr && t->workspace_bytes >= ERR_AREA_BYTES) {
    const bool ready = vp && !training && t->proj != nullptr;
    const WorkspaceLayout w = workspace_layout(p.geo, d->n_rays, training, (vp && !ready) ? d->n_rows : 0);
    uint8_t* ws = reinterpret_cast<uint8_t*>(t->workspace);
    p.err = reinterpret_cast<unsigned int*>(ws + w.err_off);
    if (pdl_mode() == 1) p.debug |= DEBUG_PDL_EARLY;
    p.trace = (p.debug & 8) ? reinterpret_cast<unsigned long long*>(ws + w.err_off + 1024) : nullptr;
    if (training) {
      p.stash = ws + w.stash_off;
      p.partw = reinterpret_cast<float*>(ws + w.partw_off);
      p.part1 = reinterpret_cast<float*>(ws + w.part1_off);
    }
    if (vp) {
      p.e_image = reinterpret_cast<const uint8_t*>(t->table_image);
      p.ntiles_v = static_cast<int>((d->n_rows + TILE_M - 1) / TILE_M);
      p.proj = ready ? const_cast<uint8_t*>(reinterpret_cast<const uint8_t*>(t->proj)) : ws + w.proj_off;
      if (training) {
        p.g_image = ws + w.gimg_off;
        p.vstart = reinterpret_cast<int32_t*>(ws + w.vstart_off);
        p.vcursor = reinterpret_cast<int32_t*>(ws + w.vcursor_off);
        p.pairs = reinterpret_cast<int32_t*>(ws + w.pairs_off);
        p.scan_tmp = reinterpret_cast<int32_t*>(ws + w.scan_off);
      }
    }
  }
  return p;
}

// The incidence lists of a training call depend on its inputs only, so they are built on a library-owned side stream,
// forked from and joined back into the caller's stream with events (capturable: a CUDA graph records the fork/join).
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  int device = -1;
};
std::mutex g_side_mutex;
SideStream g_side[16];
SideStream* side_stream() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  SideStream& s = g_side[dev];
  if (s.stream == nullptr) {
    if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    s.device = dev;
  }
  return &s;
}

// Vertex projection P = E [W0;Wy]^T: persistent CTAs in clusters of 2 (the weight stream is multicast).
int launch_projection(const FieldParams& p, int nsm, cudaStream_t st) {
  static int max2 = -1;
  if (max2 < 0) max2 = vertex_project_max_grid(2);
  int cs = (env().cluster == 2 && max2 >= 2 && p.ntiles_v >= 2) ? 2 : 1;
  int grid = p.ntiles_v < nsm ? p.ntiles_v : nsm;
  if (cs == 2) {
    if (grid > max2) grid = max2;
    grid &= ~1;
  }
  cudaError_t e = launch_vertex_project(p, grid, cs, st);
  if (e != cudaSuccess) return cuda_fail(e, "vertex_project_kernel launch");
  return INF_OK;
}

// CTAs of the per-ray kernel: one per SM walking pairs of 128-ray tiles, or two per SM walking single tiles
int rays_grid(const FieldParams& p, int nsm) {
  const int slots = env().rays_slots;
  const int units = (p.ntiles + slots - 1) / slots;
  const int cap = slots == 1 ? 2 * nsm : nsm;
  return units < cap ? units : cap;
}

}  // namespace
}  // namespace inf

using namespace inf;

extern "C" {

int inf_version(void) { return INF_ABI_VERSION; }

long long inf_launch_count(void) { return g_launches.load(); }

int inf_debug_flags(void) {
  // INF_DEBUG bits 1|2|4 and INF_REDUCE_DBG (reported as bits 8, 9) skip work and produce wrong results
  static const int rdbg = getenv("INF_REDUCE_DBG") ? atoi(getenv("INF_REDUCE_DBG")) : 0;
  return env().debug | ((rdbg & 3) << 8);
}

const char* inf_last_error(void) { return g_err; }

int64_t inf_num_params(const InfDesc* desc) {
  if (check_desc(desc) != INF_OK) return -1;
  ParamInfo pi[INF_MAX_PARAMS];
  int64_t total = 0;
  build_param_table(desc, pi, &total);
  return total;
}

int32_t inf_num_param_tensors(const InfDesc* desc) {
  if (check_desc(desc) != INF_OK) return -1;
  ParamInfo pi[INF_MAX_PARAMS];
  return build_param_table(desc, pi, nullptr);
}

int inf_param_info(const InfDesc* desc, int32_t i, int64_t* offset, int32_t* rows, int32_t* cols) {
  int rc = check_desc(desc);
  if (rc != INF_OK) return rc;
  ParamInfo pi[INF_MAX_PARAMS];
  const int n = build_param_table(desc, pi, nullptr);
  if (i < 0 || i >= n) return fail(INF_ERR_BAD_ARG, "parameter index %d out of range [0,%d)", i, n);
  if (offset) *offset = pi[i].off;
  if (rows) *rows = pi[i].rows;
  if (cols) *cols = pi[i].cols;
  return INF_OK;
}

size_t inf_packed_bytes(const InfDesc* desc) {
  if (check_desc(desc) != INF_OK) return 0;
  return packed_total_bytes(geo_of(desc));
}

size_t inf_workspace_bytes(const InfDesc* desc, int training) {
  if (check_desc(desc) != INF_OK) return 0;
  // n_rows > 0 (and a path that may use it): sized for either path, so the caller need not know which one a call takes
  const bool may_vp = desc->input_mode == INF_INPUT_GATHER && desc->path != INF_PATH_RAY && desc->n_rows > 0;
  return workspace_layout(geo_of(desc), desc->n_rays, training != 0, may_vp ? desc->n_rows : 0).total;
}

size_t inf_table_image_bytes(int64_t n_rows, int32_t k_pad) {
  if (n_rows < 0 || k_pad < KC || k_pad % KC != 0) return 0;
  return static_cast<size_t>((n_rows + TILE_M - 1) / TILE_M) * static_cast<size_t>(k_pad / KC) * PANEL_BYTES;
}

int inf_build_table_image(const void* table, int32_t table_dtype, int64_t n_rows, int32_t k_pad, void* image, void* stream) {
  if (table == nullptr || image == nullptr) return fail(INF_ERR_BAD_ARG, "table / image must be non-null");
  if (n_rows < 0 || k_pad < KC || k_pad % KC != 0) return fail(INF_ERR_BAD_ARG, "bad table shape: rows=%lld k_pad=%d", static_cast<long long>(n_rows), k_pad);
  if (table_dtype != INF_TABLE_F32 && table_dtype != INF_TABLE_F16) return fail(INF_ERR_BAD_ARG, "bad table_dtype");
  if ((reinterpret_cast<uintptr_t>(table) & 15) != 0 || (reinterpret_cast<uintptr_t>(image) & 127) != 0)
    return fail(INF_ERR_BAD_ARG, "table must be 16-byte and image 128-byte aligned");
  launch_table_image(table, table_dtype == INF_TABLE_F16, n_rows, k_pad, image, static_cast<cudaStream_t>(stream));
  g_launches += 1;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "table_image launch");
  return INF_OK;
}

int inf_project_vertices(const InfDesc* desc, const InfTensors* t, void* out, void* stream) {
  int rc = check_desc(desc);
  if (rc != INF_OK) return rc;
  if (t == nullptr || t->table_image == nullptr || t->packed == nullptr || out == nullptr)
    return fail(INF_ERR_BAD_ARG, "table_image / packed / out must be non-null");
  if (desc->n_rows < 1) return fail(INF_ERR_BAD_ARG, "n_rows must be given");
  InfDesc d = *desc;
  d.path = INF_PATH_VERTEX;
  d.input_mode = INF_INPUT_GATHER;
  if (d.n_rays < 1) d.n_rays = 1;
  const WorkspaceLayout w = workspace_layout(geo_of(&d), d.n_rays, false, d.n_rows);
  if (t->workspace == nullptr || t->workspace_bytes < w.total)
    return fail(INF_ERR_WORKSPACE, "workspace %zu B < required %zu B", t->workspace_bytes, w.total);
  int nsm = 0;
  rc = sm_count(&nsm);
  if (rc != INF_OK) return rc;
  FieldParams p = make_params(&d, t, false);
  if (p.e_image == nullptr) return fail(INF_ERR_BAD_ARG, "vertex path not available for this call");
  rc = launch_projection(p, nsm, static_cast<cudaStream_t>(stream));
  if (rc != INF_OK) return rc;
  launch_proj_unpermute(p.proj, d.n_rows, out, static_cast<cudaStream_t>(stream));
  g_launches += 2;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "proj_unpermute launch");
  return INF_OK;
}

size_t inf_projection_bytes(int64_t n_rows) {
  if (n_rows < 1) return 0;
  return static_cast<size_t>((n_rows + TILE_M - 1) / TILE_M) * PROJ_TILE_BYTES;
}

int inf_project_vertices_raw(const InfDesc* desc, const InfTensors* t, void* proj, void* stream) {
  int rc = check_desc(desc);
  if (rc != INF_OK) return rc;
  if (t == nullptr || t->table_image == nullptr || t->packed == nullptr || proj == nullptr)
    return fail(INF_ERR_BAD_ARG, "table_image / packed / proj must be non-null");
  if (desc->n_rows < 1) return fail(INF_ERR_BAD_ARG, "n_rows must be given");
  int nsm = 0;
  rc = sm_count(&nsm);
  if (rc != INF_OK) return rc;
  FieldParams p{};
  p.packed = reinterpret_cast<const uint8_t*>(t->packed);
  p.e_image = reinterpret_cast<const uint8_t*>(t->table_image);
  p.n_rows = desc->n_rows;
  p.ntiles_v = static_cast<int>((desc->n_rows + TILE_M - 1) / TILE_M);
  p.proj = reinterpret_cast<uint8_t*>(proj);
  p.geo = geo_of(desc);
  if (t->workspace != nullptr && t->workspace_bytes >= ERR_AREA_BYTES) p.err = reinterpret_cast<unsigned int*>(t->workspace);
  rc = launch_projection(p, nsm, static_cast<cudaStream_t>(stream));
  if (rc != INF_OK) return rc;
  g_launches += 1;
  return INF_OK;
}

int inf_uses_vertex_path(const InfDesc* desc, const InfTensors* t) {
  if (check_desc(desc) != INF_OK) return 0;
  return vertex_path(desc, t) ? 1 : 0;
}

int inf_relayout_table_scaled(const float* src, int64_t n_rows, int32_t k, int64_t src_row_stride, const int8_t* col_exp, void* dst,
                              int32_t k_pad, int32_t dst_dtype, void* stream);

int inf_relayout_table(const float* src, int64_t n_rows, int32_t k, int64_t src_row_stride, void* dst, int32_t k_pad,
                       int32_t dst_dtype, void* stream) {
  return inf_relayout_table_scaled(src, n_rows, k, src_row_stride, nullptr, dst, k_pad, dst_dtype, stream);
}

int inf_column_exponents(const float* src, int64_t n_rows, int32_t k, int64_t src_row_stride, int32_t k_pad, int8_t* col_exp,
                         void* scratch, void* stream) {
  if (src == nullptr || col_exp == nullptr || scratch == nullptr) return fail(INF_ERR_BAD_ARG, "src / col_exp / scratch must be non-null");
  if (n_rows < 0 || k < 1 || k_pad < k || k_pad % KC != 0 || src_row_stride < k) return fail(INF_ERR_BAD_ARG, "bad table shape");
  launch_column_exponents(src, n_rows, k, src_row_stride, k_pad, col_exp, reinterpret_cast<unsigned int*>(scratch),
                          static_cast<cudaStream_t>(stream));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "column_exponents launch");
  g_launches += 2;
  return INF_OK;
}

int inf_relayout_table_scaled(const float* src, int64_t n_rows, int32_t k, int64_t src_row_stride, const int8_t* col_exp, void* dst,
                              int32_t k_pad, int32_t dst_dtype, void* stream) {
  if (src == nullptr || dst == nullptr) return fail(INF_ERR_BAD_ARG, "src / dst must be non-null");
  if (n_rows < 0 || k < 1 || k_pad < k || k_pad % KC != 0 || src_row_stride < k)
    return fail(INF_ERR_BAD_ARG, "bad table shape: rows=%lld k=%d k_pad=%d stride=%lld", static_cast<long long>(n_rows), k,
                k_pad, static_cast<long long>(src_row_stride));
  if (dst_dtype != INF_TABLE_F32 && dst_dtype != INF_TABLE_F16) return fail(INF_ERR_BAD_ARG, "bad dst_dtype");
  launch_relayout(src, n_rows, k, src_row_stride, dst, k_pad, dst_dtype, col_exp, static_cast<cudaStream_t>(stream));
  g_launches += 1;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "relayout_table launch");
  return INF_OK;
}

int inf_pack_weights_scaled(const InfDesc* desc, const float* const* params, const int8_t* col_exp, void* packed, void* stream);

int inf_pack_weights(const InfDesc* desc, const float* const* params, void* packed, void* stream) {
  return inf_pack_weights_scaled(desc, params, nullptr, packed, stream);
}

int inf_pack_weights_scaled(const InfDesc* desc, const float* const* params, const int8_t* col_exp, void* packed, void* stream) {
  int rc = check_desc(desc);
  if (rc != INF_OK) return rc;
  if (params == nullptr || packed == nullptr) return fail(INF_ERR_BAD_ARG, "params / packed must be non-null");
  PackParams pp{};
  pp.geo = geo_of(desc);
  pp.col_exp = col_exp;
  pp.packed = reinterpret_cast<uint8_t*>(packed);
  const int L = desc->num_layers;
  const bool view = desc->view_strategy != INF_VIEW_NONE;
  int idx = 0;
  for (int i = 0; i < L; ++i) {
    const float* w = params[idx++];
    const float* b = params[idx++];
    if (i == L - 1 && !view) { pp.wlast = w; pp.blast = b; }
    else { pp.w[i] = w; pp.b[i] = b; }
    if (i == desc->skip) {
      pp.wy = params[idx++];
      pp.by = params[idx++];
    }
  }
  if (view) {
    pp.w[pp.geo.nb] = params[idx++];
    pp.b[pp.geo.nb] = params[idx++];
    pp.wlast = params[idx++];
    pp.blast = params[idx++];
  }
  for (int i = 0; i < idx; ++i)
    if (params[i] == nullptr) return fail(INF_ERR_BAD_ARG, "params[%d] is null", i);
  launch_pack_weights(pp, static_cast<cudaStream_t>(stream));
  g_launches += 1;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "pack_weights launch");
  return INF_OK;
}

int inf_forward(const InfDesc* desc, const InfTensors* t, void* stream) {
  int rc = check_desc(desc);
  if (rc != INF_OK) return rc;
  rc = check_tensors(desc, t, false);
  if (rc != INF_OK) return rc;
  if (desc->n_rays == 0) return INF_OK;
  int nsm = 0;
  rc = sm_count(&nsm);
  if (rc != INF_OK) return rc;
  FieldParams p = make_params(desc, t, false);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int slot = g_prof.n;
  prof_record(slot, 0, st);
  prof_record(slot, 1, st);
  if (p.proj != nullptr) {
    const bool ready = t->proj != nullptr;   // the caller's projection of these weights: the per-ray kernel only
    if (!ready) {
      rc = launch_projection(p, nsm, st);
      if (rc != INF_OK) return rc;
    }
    prof_record(slot, 2, st);
    cudaError_t e = launch_field_rays(p, false, env().rays_slots, rays_grid(p, nsm), st);
    if (e != cudaSuccess) return cuda_fail(e, "field_rays_kernel<inference> launch");
    g_launches += ready ? 1 : 2;   // [vertex_project,] field_rays
  } else {
    prof_record(slot, 2, st);
    const int grid = pick_main_grid(desc, p.ntiles, nsm, false, &p.cluster);
    cudaError_t e = launch_field_main(p, false, desc->table_dtype == INF_TABLE_F16, grid, st);
    if (e != cudaSuccess) return cuda_fail(e, "field_main_kernel<inference> launch");
    g_launches += 1;
  }
  for (int i = 3; i < PROF_EVENTS; ++i) prof_record(slot, i, st);
  if (g_prof.on && slot < PROF_SLOTS) g_prof.n = slot + 1;
  return INF_OK;
}

static int check_adam(const InfAdam* a) {
  if (a == nullptr) return INF_OK;
  if (!a->params || !a->exp_avg || !a->exp_avg_sq || !a->packed)
    return fail(INF_ERR_BAD_ARG, "adam: params / exp_avg / exp_avg_sq / packed must be non-null");
  if ((a->step_dev == nullptr && a->step < 1) || !(a->lr >= 0.f) || !(a->beta1 >= 0.f && a->beta1 < 1.f) || !(a->beta2 >= 0.f && a->beta2 < 1.f) || !(a->eps >= 0.f))
    return fail(INF_ERR_BAD_ARG, "adam: bad hyper-parameters (step=%d lr=%g betas=%g,%g eps=%g)", a->step, a->lr, a->beta1, a->beta2, a->eps);
  return INF_OK;
}

static AdamParams adam_params(const InfAdam* a) {
  AdamParams ap{};
  if (a == nullptr) return ap;
  ap.params = a->params; ap.exp_avg = a->exp_avg; ap.exp_avg_sq = a->exp_avg_sq;
  ap.packed = reinterpret_cast<uint8_t*>(a->packed);
  ap.step_dev = a->step_dev;
  ap.step_autoinc = (a->step_dev != nullptr && (a->flags & INF_ADAM_STEP_AUTOINC)) ? 1 : 0;
  ap.col_exp = a->col_exp;
  ap.lr = a->lr;
  const int tstep = a->step < 1 ? 1 : a->step;
  const double bc1 = 1.0 - pow(static_cast<double>(a->beta1), tstep);
  const double bc2 = 1.0 - pow(static_cast<double>(a->beta2), tstep);
  ap.step_size = static_cast<float>(static_cast<double>(a->lr) / bc1);
  ap.bc2_sqrt = static_cast<float>(sqrt(bc2));
  ap.beta1 = a->beta1; ap.beta2 = a->beta2; ap.eps = a->eps;
  ap.omb1 = static_cast<float>(1.0 - static_cast<double>(a->beta1));
  ap.omb2 = static_cast<float>(1.0 - static_cast<double>(a->beta2));
  return ap;
}

static int run_forward_backward(const InfDesc* desc, const InfTensors* t, const InfGrads* g, const InfAdam* adam, void* stream,
                                const InfComm* comm = nullptr) {
  int rc = check_desc(desc);
  if (rc != INF_OK) return rc;
  rc = check_tensors(desc, t, true);
  if (rc != INF_OK) return rc;
  if (g == nullptr || g->flat == nullptr) return fail(INF_ERR_BAD_ARG, "grads.flat must be non-null");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ParamInfo pi[INF_MAX_PARAMS];
  int64_t total = 0;
  const int nparam = build_param_table(desc, pi, &total);
  if (desc->n_rays == 0 && adam != nullptr) return fail(INF_ERR_BAD_ARG, "inf_train_step needs at least one ray");
  if (desc->n_rays == 0) {
    cudaError_t e = cudaMemsetAsync(g->flat, 0, static_cast<size_t>(total) * 4, st);
    if (e == cudaSuccess && g->loss_sums) e = cudaMemsetAsync(g->loss_sums, 0, 12, st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync");
    return INF_OK;
  }
  int nsm = 0;
  rc = sm_count(&nsm);
  if (rc != INF_OK) return rc;
  FieldParams p = make_params(desc, t, true);
  const bool vp = p.e_image != nullptr;
  // External upstream gradients are typically O(1/(3B)); bring them to O(1) for the fp16 chain.
  const float ext = (t->grad_pred != nullptr) ? 3.f * static_cast<float>(desc->n_rays) : 1.f;
  p.ext_scale = ext;
  const Geo geo = p.geo;
  const int grid1 = vp ? rays_grid(p, nsm) : pick_main_grid(desc, p.ntiles, nsm, true, &p.cluster);
  // wgrad CTAs: every slice gets ray groups in proportion to the bytes it moves over the whole launch
  const WPlan& plan = p.plan;
  // clusters of cw big-slice CTAs share the dz_0 / dz_skip half-panels (multicast): cw = 4, 2 or 1, dividing nbig
  int cw = 4;
  {
    // per-ray path: clusters of 4 share the dz_0 / dz_skip half-panels (0.152 -> 0.127 ms, r1f); vertex path: clusters of 2 --
    // only 33 clusters of 4 fit the 148 SMs (132 CTAs), 74 clusters of 2 use all of them (wgrad 96 -> 91.6 us, r2)
    const int env_cw = env().wg_cluster > 0 ? env().wg_cluster : (vp ? 2 : 4);
    cw = env_cw >= 4 ? 4 : (env_cw >= 2 ? 2 : 1);
    while (cw > 1 && geo.nbig % cw != 0) cw >>= 1;
  }
  int nsm_w = nsm;
  if (cw > 1) {
    static int cached[5] = {-1, -1, -1, -1, -1};
    if (cached[cw] < 0) cached[cw] = field_wgrad_max_grid(cw);
    if (cached[cw] < cw) cw = 1;
    else if (cached[cw] < nsm_w) nsm_w = cached[cw];
  }
  // tiles a big slice walks: ray tiles, or vertex tiles on the vertex path ([dW0;dWy] = G^T E)
  const int tiles_big = vp ? p.ntiles_v : p.ntiles;
  int used = 0;
  int gh[MAX_HID_SLICES] = {0, 0, 0, 0};
  int gb = 1;
  const bool split = vp && env().wgrad_split != 0;
  if (split) {
    // Two launches: the hidden slices (all SMs, no clusters) run while the vertex gradient is gathered on the side
    // stream; the big slices follow.  Each launch splits its SMs in proportion to the bytes a slice ingests.
    double hsum = 0;
    for (int s = 0; s < plan.nslice; ++s) hsum += plan.bytes_per_tile[s];
    int hid_ctas = env().wgrad_hid_ctas > 0 ? env().wgrad_hid_ctas : nsm;
    if (hid_ctas > nsm) hid_ctas = nsm;
    for (int s = 0; s < plan.nslice; ++s) {
      int v = static_cast<int>(hid_ctas * (plan.bytes_per_tile[s] / hsum));
      if (v < 1) v = 1;
      if (v > p.ntiles) v = p.ntiles;
      gh[s] = v;
      used += v;
    }
    gb = nsm_w / geo.nbig;
  } else {
    // bytes a CTA ingests over the launch: the x panels of its 256 columns + its share of dz_0 / dz_skip per (ray or
    // vertex) tile (INF_WGRAD_BIGW scales the big slices' share: they also carry 2.5x the tensor work per ingested byte)
    const double bigw = env().wgrad_bigw > 0.0 ? env().wgrad_bigw : 1.0;
    const double big_bytes = (2.0 * BLOB_BYTES + 2.0 * BLOB_BYTES / cw) * bigw * tiles_big;
    double wsum = static_cast<double>(geo.nbig) * big_bytes;
    for (int s = 0; s < plan.nslice; ++s) wsum += static_cast<double>(plan.bytes_per_tile[s]) * p.ntiles;
    int smax = 0;
    for (int s = 0; s < plan.nslice; ++s) {
      int v = static_cast<int>(nsm_w * (static_cast<double>(plan.bytes_per_tile[s]) * p.ntiles / wsum));
      if (v < 1) v = 1;
      if (v > p.ntiles) v = p.ntiles;
      gh[s] = v;
      used += v;
      if (plan.bytes_per_tile[s] > plan.bytes_per_tile[smax]) smax = s;
    }
    if (cw > 1 && used % cw != 0) {   // whole clusters only: pad the heaviest hidden slice
      const int add = cw - used % cw;
      if (plan.nslice > 0 && gh[smax] + add <= p.ntiles) { gh[smax] += add; used += add; }
      else cw = 1;
    }
    gb = (nsm_w - used) / geo.nbig;
  }
  if (gb < 1) gb = 1;
  if (gb > tiles_big) gb = tiles_big;
  {
    // Small batches: every wgrad CTA dumps its whole 256 KB accumulator and the reduction reads it back, whatever it
    // accumulated; INF_WGRAD_MIN_TILES = m makes every CTA own at least m tiles (measurement knob, see Env)
    const int mt = env().wgrad_min_tiles;
    const int cap_big = (tiles_big + mt - 1) / mt, cap_hid = (p.ntiles + mt - 1) / mt;
    if (gb > cap_big) gb = cap_big;
    int used2 = 0;
    for (int s = 0; s < plan.nslice; ++s) {
      if (gh[s] > cap_hid) gh[s] = cap_hid;
      used2 += gh[s];
    }
    used = used2;
    if (!split && cw > 1 && used % cw != 0) {   // whole clusters only (the big CTAs are a multiple of cw: nbig % cw == 0)
      const int add = cw - used % cw;
      int s_pad = -1;
      for (int s = 0; s < plan.nslice; ++s)
        if (gh[s] + add <= p.ntiles && (s_pad < 0 || plan.bytes_per_tile[s] > plan.bytes_per_tile[s_pad])) s_pad = s;
      if (s_pad >= 0) { gh[s_pad] += add; used += add; }
      else cw = 1;
    }
  }
  p.wg_cluster = cw;
  p.groups_big = gb;
  for (int s = 0; s < MAX_HID_SLICES; ++s) p.groups_hid[s] = gh[s];
  const int grid_big = geo.nbig * gb;
  const int grid2 = grid_big + used;
  if (grid2 > (split ? MAX_WGRAD_CTAS : MAX_CTAS))
    return fail(INF_ERR_UNSUPPORTED, "k_pad=%d needs %d wgrad CTAs > %d", desc->k_pad, grid2, split ? MAX_WGRAD_CTAS : MAX_CTAS);

  const bool f16 = desc->table_dtype == INF_TABLE_F16;
  const int slot = g_prof.n;
  cudaError_t e;
  prof_record(slot, 0, st);
  if (vp) {
    std::lock_guard<std::mutex> lock(g_side_mutex);
    SideStream* side = env().side_stream ? side_stream() : nullptr;
    if (side != nullptr) {
      e = cudaEventRecord(side->fork, st);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(side->stream, side->fork, 0);
      if (e == cudaSuccess) e = launch_csr_build(p, side->stream);
      if (e == cudaSuccess) e = cudaEventRecord(side->join, side->stream);
    } else {
      e = launch_csr_build(p, st);
    }
    if (e != cudaSuccess) return cuda_fail(e, "incidence-list (CSR) kernels launch");
    prof_record(slot, 1, st);
    rc = launch_projection(p, nsm, st);
    if (rc != INF_OK) return rc;
    prof_record(slot, 2, st);
    if (side != nullptr && !split) {
      // The incidence lists (~25 us on the side stream, behind the ~40 us projection) join the caller's stream HERE, before
      // the per-ray kernel, so that rays -> gather -> wgrad -> reduce is an unbroken chain of programmatic dependent launches
      e = cudaStreamWaitEvent(st, side->join, 0);
      if (e != cudaSuccess) return cuda_fail(e, "stream join");
    }
    e = launch_field_rays(p, true, env().rays_slots, grid1, st);
    if (e != cudaSuccess) return cuda_fail(e, "field_rays_kernel<train> launch");
    prof_record(slot, 3, st);
    // (INF_WGRAD_SPLIT) vertex gradient on the side stream (after the incidence lists, which are already queued there),
    // concurrently with the hidden-slice weight gradients on the caller's stream; the big slices ([dW0;dWy] = G^T E) wait for it
    cudaStream_t gs = st;
    if (side != nullptr && split) {
      e = cudaEventRecord(side->fork, st);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(side->stream, side->fork, 0);
      if (e != cudaSuccess) return cuda_fail(e, "stream fork");
      gs = side->stream;
    }
    e = launch_vertex_gather(p, nsm, gs);
    if (e != cudaSuccess) return cuda_fail(e, "vertex_gather_kernel launch");
    if (split) {
      FieldParams ph = p;
      ph.cta_base = grid_big;
      ph.wg_cluster = 1;
      e = launch_field_wgrad(ph, used, st);
      if (e != cudaSuccess) return cuda_fail(e, "field_wgrad_kernel (hidden slices) launch");
    }
    prof_record(slot, 4, st);
    if (side != nullptr && split) {
      e = cudaEventRecord(side->join, side->stream);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(st, side->join, 0);
      if (e != cudaSuccess) return cuda_fail(e, "stream join");
    }
  } else {
    prof_record(slot, 1, st);
    prof_record(slot, 2, st);
    e = launch_field_main(p, true, f16, grid1, st);
    if (e != cudaSuccess) return cuda_fail(e, "field_main_kernel<train> launch");
    prof_record(slot, 3, st);
    prof_record(slot, 4, st);
  }
  e = launch_field_wgrad(p, split ? grid_big : grid2, st);   // split: the big slices only (cta_base = 0)
  prof_record(slot, 5, st);
  if (e != cudaSuccess) return cuda_fail(e, "field_wgrad_kernel launch");

  ReduceParams rp{};
  rp.nseg = nparam;
  for (int i = 0; i < nparam; ++i)
    rp.seg[i] = ReduceSeg{pi[i].off, pi[i].rows, pi[i].cols, pi[i].kind, pi[i].a, pi[i].b, pi[i].split, pi[i].layer, pi[i].pair_off};
  rp.total = total;
  rp.partw = p.partw;
  rp.part1 = p.part1;
  rp.flat = g->flat;
  rp.loss_sums = (t->target != nullptr) ? g->loss_sums : nullptr;
  rp.ncta1 = vp ? grid1 * env().rays_slots : grid1;   // rows of the forward/backward kernel's partials
  rp.groups_big = gb;
  int base = geo.nbig * gb;
  for (int s = 0; s < MAX_HID_SLICES; ++s) {
    rp.groups_hid[s] = gh[s];
    rp.cta_base_hid[s] = base;
    base += gh[s];
  }
  rp.scale = desc->grad_scale / (DZ_SCALE * ext);
  rp.loss_scale = desc->grad_scale;
  rp.do_reduce = 1;
  rp.pdl_early = pdl_mode() == 1;
  rp.col_exp = t->col_exp;
  rp.err = p.err;
  rp.trace = p.trace;
  if (comm != nullptr && comm->world > 1) {
    rp.comm.world = comm->world;
    rp.comm.rank = comm->rank;
    for (int q = 0; q < comm->world; ++q) rp.comm.region[q] = reinterpret_cast<float*>(comm->regions[q]);
    rp.comm.slot_floats = comm_slot_floats(total);
    {
      static const int two_shot_min = getenv("INF_EXCH_TWO_SHOT_MIN") ? atoi(getenv("INF_EXCH_TWO_SHOT_MIN")) : 4;
      rp.comm.own_chunk = comm->world >= two_shot_min ? (((total + comm->world - 1) / comm->world + 31) / 32) * 32 : 0;
    }
  }
  rp.adam = adam_params(adam);
  // the block counter of the step auto-increment lives in the workspace's (zero-initialised) error area
  rp.adam.done_ctr = p.err != nullptr ? p.err + 4 : nullptr;
  if (rp.adam.done_ctr == nullptr) rp.adam.step_autoinc = 0;
  rp.geo = geo;
  rp.plan = plan;
  e = launch_field_reduce(rp, st);
  // vertex path: 5 incidence-list kernels, vertex_project, field_rays, vertex_gather, field_wgrad (x2 when split), reduce
  g_launches += vp ? (split ? 11 : 10) : 3;
  prof_record(slot, 6, st);
  if (g_prof.on && slot < PROF_SLOTS) g_prof.n = slot + 1;
  if (e == cudaErrorNotSupported && comm != nullptr && comm->world > 1)
    return fail(INF_ERR_UNSUPPORTED, "the multi-GPU step needs the reduction tables, which cannot be built while a stream capture is in "
                                     "progress (or beyond the table budget): run one eager step with this configuration first");
  if (e != cudaSuccess) return cuda_fail(e, "field_reduce_kernel launch");
  return INF_OK;
}

int inf_forward_backward(const InfDesc* desc, const InfTensors* t, const InfGrads* g, void* stream) {
  return run_forward_backward(desc, t, g, nullptr, stream);
}

int inf_train_step(const InfDesc* desc, const InfTensors* t, const InfGrads* g, const InfAdam* adam, void* stream) {
  if (adam == nullptr) return fail(INF_ERR_BAD_ARG, "adam is null");
  int rc = check_adam(adam);
  if (rc != INF_OK) return rc;
  return run_forward_backward(desc, t, g, adam, stream);
}

int inf_train_step_comm(const InfDesc* desc, const InfTensors* t, const InfGrads* g, const InfAdam* adam, const InfComm* comm,
                        void* stream) {
  if (adam == nullptr) return fail(INF_ERR_BAD_ARG, "adam is null");
  int rc = check_adam(adam);
  if (rc != INF_OK) return rc;
  if (comm != nullptr && comm->world > 1) {
    if (comm->world > INF_MAX_PEERS || comm->rank < 0 || comm->rank >= comm->world) return fail(INF_ERR_BAD_ARG, "comm: bad world / rank");
    if (check_desc(desc) != INF_OK) return INF_ERR_BAD_ARG;
    if (comm->region_bytes < inf_comm_region_bytes(desc)) return fail(INF_ERR_WORKSPACE, "comm: exchange region too small");
    for (int q = 0; q < comm->world; ++q)
      if (comm->regions[q] == nullptr) return fail(INF_ERR_BAD_ARG, "comm: region of rank %d is not mapped", q);
    if (t != nullptr && t->target == nullptr) return fail(INF_ERR_BAD_ARG, "comm: the exchanged step uses the built-in L1 loss (target)");
  }
  return run_forward_backward(desc, t, g, adam, stream, comm);
}

size_t inf_comm_region_bytes(const InfDesc* desc) {
  if (check_desc(desc) != INF_OK) return 0;
  ParamInfo pi[INF_MAX_PARAMS];
  int64_t total = 0;
  build_param_table(desc, pi, &total);
  return comm_region_bytes(total);
}

int inf_comm_create(size_t bytes, void** region, void* ipc_handle_out) {
  if (region == nullptr || ipc_handle_out == nullptr || bytes == 0) return fail(INF_ERR_BAD_ARG, "comm_create: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == INF_IPC_HANDLE_BYTES, "IPC handle size");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc (exchange region)");
  e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "cudaIpcGetMemHandle"); }
  memcpy(ipc_handle_out, &h, sizeof(h));
  *region = p;
  return INF_OK;
}

int inf_comm_open(const void* ipc_handle, void** mapped) {
  if (ipc_handle == nullptr || mapped == nullptr) return fail(INF_ERR_BAD_ARG, "comm_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle, sizeof(h));
  cudaError_t e = cudaIpcOpenMemHandle(mapped, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return cuda_fail(e, "cudaIpcOpenMemHandle");
  return INF_OK;
}

int inf_comm_close(void* mapped) {
  cudaError_t e = cudaIpcCloseMemHandle(mapped);
  if (e != cudaSuccess) return cuda_fail(e, "cudaIpcCloseMemHandle");
  return INF_OK;
}

int inf_comm_free(void* region) {
  cudaError_t e = cudaFree(region);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFree");
  return INF_OK;
}

int inf_adam_pack_step(const InfDesc* desc, float* flat_grad, const InfAdam* adam, void* stream) {
  int rc = check_desc(desc);
  if (rc != INF_OK) return rc;
  if (flat_grad == nullptr || adam == nullptr) return fail(INF_ERR_BAD_ARG, "flat_grad / adam must be non-null");
  rc = check_adam(adam);
  if (rc != INF_OK) return rc;
  ParamInfo pi[INF_MAX_PARAMS];
  int64_t total = 0;
  const int nparam = build_param_table(desc, pi, &total);
  ReduceParams rp{};
  rp.nseg = nparam;
  for (int i = 0; i < nparam; ++i)
    rp.seg[i] = ReduceSeg{pi[i].off, pi[i].rows, pi[i].cols, pi[i].kind, pi[i].a, pi[i].b, pi[i].split, pi[i].layer, pi[i].pair_off};
  rp.total = total;
  rp.flat = flat_grad;
  rp.do_reduce = 0;
  rp.adam = adam_params(adam);
  if (rp.adam.step_autoinc) return fail(INF_ERR_BAD_ARG, "INF_ADAM_STEP_AUTOINC is a flag of inf_train_step[_comm]");
  rp.geo = geo_of(desc);
  rp.plan = make_wplan(rp.geo);
  cudaError_t e = launch_field_reduce(rp, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "field_reduce_kernel<adam> launch");
  g_launches += 1;
  return INF_OK;
}

int inf_gather_interpolate(const void* table, int32_t table_dtype, int32_t k, int32_t k_pad, const void* vids,
                           int32_t idx_dtype, const float* bary, int64_t n_rays, float* out, void* stream) {
  return inf_gather_interpolate_scaled(table, table_dtype, k, k_pad, nullptr, vids, idx_dtype, bary, n_rays, out, stream);
}

int inf_gather_interpolate_scaled(const void* table, int32_t table_dtype, int32_t k, int32_t k_pad, const int8_t* col_exp,
                                  const void* vids, int32_t idx_dtype, const float* bary, int64_t n_rays, float* out, void* stream) {
  if (n_rays < 0 || k < 1 || k_pad < k || k_pad % KC != 0) return fail(INF_ERR_BAD_ARG, "bad shape");
  if (n_rays > 0 && (!table || !vids || !bary || !out)) return fail(INF_ERR_BAD_ARG, "null pointer");
  launch_gather_interpolate(table, table_dtype == INF_TABLE_F16, k, k_pad, vids, idx_dtype == INF_IDX_I64, bary, n_rays, out,
                            col_exp, static_cast<cudaStream_t>(stream));
  g_launches += 1;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "gather_interpolate launch");
  return INF_OK;
}

int inf_profile_enable(int on) {
  if (on && !g_prof.created) {
    for (int s = 0; s < PROF_SLOTS; ++s)
      for (int i = 0; i < PROF_EVENTS; ++i) {
        cudaError_t e = cudaEventCreate(&g_prof.ev[s][i]);
        if (e != cudaSuccess) return cuda_fail(e, "cudaEventCreate");
      }
    g_prof.created = true;
  }
  g_prof.on = on != 0;
  g_prof.n = 0;
  return INF_OK;
}

int inf_profile_read_stages(float* stage_ms, int* calls) {
  const int n = g_prof.n < PROF_SLOTS ? g_prof.n : PROF_SLOTS;
  double acc[INF_PROFILE_STAGES] = {0, 0, 0, 0, 0, 0};
  for (int s = 0; s < n; ++s) {
    cudaError_t e = cudaEventSynchronize(g_prof.ev[s][PROF_EVENTS - 1]);
    if (e != cudaSuccess) return cuda_fail(e, "cudaEventSynchronize");
    for (int i = 0; i < INF_PROFILE_STAGES; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, g_prof.ev[s][i], g_prof.ev[s][i + 1]);
      acc[i] += ms;
    }
  }
  if (stage_ms)
    for (int i = 0; i < INF_PROFILE_STAGES; ++i) stage_ms[i] = n ? static_cast<float>(acc[i] / n) : 0.f;
  if (calls) *calls = n;
  return INF_OK;
}

int inf_profile_read(float* main_ms, float* wgrad_ms, float* reduce_ms, int* calls) {
  float st[INF_PROFILE_STAGES];
  int rc = inf_profile_read_stages(st, calls);
  if (rc != INF_OK) return rc;
  if (main_ms) *main_ms = st[2];
  if (wgrad_ms) *wgrad_ms = st[4];
  if (reduce_ms) *reduce_ms = st[5];
  return INF_OK;
}

int inf_debug_umma(int mode, int n, int k, const float* a, const float* b, float* d, void* stream) {
  if (mode < 0 || mode > 3 || n < 16 || n > 256 || n % 16 != 0 || (k != 64 && k != 128) || !a || !b || !d)
    return fail(INF_ERR_BAD_ARG, "inf_debug_umma: mode in [0,3], n in {16..256 step 16}, k in {64,128}");
  if (mode >= 2 && n % 64 != 0) return fail(INF_ERR_BAD_ARG, "inf_debug_umma: MN-major B needs n %% 64 == 0");
  launch_umma_probe(mode, n, k, a, b, d, static_cast<cudaStream_t>(stream));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "umma_probe launch");
  return INF_OK;
}

}  // extern "C"
