/*
 * inf_field.h -- C ABI of the B200-native intrinsic-neural-field evaluation path.
 *
 * One shared library (libinf_field.so, built from intrinsic-neural-fields_b200/csrc) exports
 * exactly the entry points below.  They replace, for the one hot path
 *
 *     gather 3 eigenfunction rows per ray -> barycentric interpolation -> MLP -> RGB
 *     (+ the L1-loss training backward pass producing the MLP weight gradients)
 *
 * the following reference interfaces (paths relative to tum-vision/intrinsic-neural-fields):
 *
 *   inf_forward            mesh.py:313-324  get_k_eigenfunc_vec_vals(E, vids, bary)
 *                          model.py:98-112  TextureField.forward(batch)
 *                          layers.py:60-62  LinearWithConcatAndActivation.forward
 *                          model.py:175-191 TextureFieldWithViewDependency.forward (view-direction branch:
 *                          face-normal angle / raw direction -> layers.py:21-25 FourierFeatEnc -> directional_mlp)
 *                          renderer.py:113-118 per-chunk model(batch) of Renderer.render
 *   inf_forward_backward   trainer.py:71-84 Trainer._train_step (forward, loss, backward; Adam
 *                          stays torch.optim.Adam), config.py:117-118 (L1 = F.l1_loss, mean)
 *   inf_train_step         the same plus config.py:108 torch.optim.Adam(model.parameters(), lr).step() and the refresh
 *   inf_adam_pack_step     of the operand images, fused into the gradient-reduction kernel (trainer.py:78-82)
 *   inf_relayout_table     mesh.py:53-108   tail of load_first_k_eigenfunctions (dense (V,k) f32
 *                          table -> device-resident padded table the kernels gather from)
 *   inf_pack_weights       model.py:42-96   parameter tensors of TextureField (state_dict order)
 *                          -> tensor-core operand images
 *
 * The reference has no FFI layer of its own (pure PyTorch); the binding a
 * maintainer adds is the ctypes stub shown in INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is a CUDA device pointer owned by the caller unless stated otherwise; the
 *     library never frees or retains a buffer past the call and allocates nothing persistent;
 *   - work is enqueued asynchronously on `stream` (a cudaStream_t passed as void*); no hidden
 *     device synchronisation; safe to call from any host thread (no thread-local CUDA state);
 *   - return value 0 = success, negative = error (INF_ERR_*); the message of the last error on
 *     the calling thread is returned by inf_last_error();
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef INF_FIELD_H_
#define INF_FIELD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define INF_ABI_VERSION 6

/* error codes */
#define INF_OK 0
#define INF_ERR_BAD_ARG (-1)      /* null pointer, bad shape, misaligned buffer            */
#define INF_ERR_UNSUPPORTED (-2)  /* config outside the compiled set (hidden!=128, ...)     */
#define INF_ERR_WORKSPACE (-3)    /* workspace too small                                   */
#define INF_ERR_CUDA (-4)         /* CUDA runtime error (message has cudaGetErrorString)   */

/* InfDesc.idx_dtype */
#define INF_IDX_I32 0
#define INF_IDX_I64 1
/* InfDesc.table_dtype */
#define INF_TABLE_F32 0
#define INF_TABLE_F16 1
/* InfDesc.input_mode */
#define INF_INPUT_GATHER 0 /* table (V,k_pad) + vids (B,3) + bary (B,3): the fused path          */
#define INF_INPUT_DENSE 1  /* table IS the materialised embedding (B,k_pad), no vids/bary        */
/* InfDesc.view_strategy */
#define INF_VIEW_NONE 0
#define INF_VIEW_INTRINSIC 1 /* model.py:176-182 */
#define INF_VIEW_EXTRINSIC 2 /* model.py:183-184 */

/* InfDesc.path: how layer 0 + the skip branch Ly (the two K = k contractions) are evaluated */
#define INF_PATH_AUTO 0   /* vertex path when a table image is given and n_rows <= 2 * n_rays, else per-ray        */
#define INF_PATH_RAY 1    /* per ray: gather 3 table rows, interpolate, GEMM over the interpolated embedding       */
#define INF_PATH_VERTEX 2 /* per vertex: P = E [W0;Wy]^T once per call (needs InfTensors.table_image), rays gather */
                          /* 3 rows of P; backward: G = scatter of bary * dz by vertex, [dW0;dWy] = G^T E          */

#define INF_MAX_LAYERS 8
#define INF_MAX_PARAMS 20

/* Shape / mode descriptor of one call (plain data, passed by pointer from host memory). */
typedef struct InfDesc {
  int32_t abi_version;   /* INF_ABI_VERSION                                                   */
  int32_t idx_dtype;     /* dtype of vids / face_idxs                                         */
  int64_t n_rays;        /* B: rays in this call (any value >= 0; ragged tail tiles handled)  */
  int64_t n_rows;        /* V: rows of the table (GATHER) -- unused for DENSE                 */
  int32_t k;             /* embedding dim = in_dim of layer 0 (model.py:211-215)              */
  int32_t k_pad;         /* row stride of the table in elements, multiple of 64, >= k         */
  int32_t hidden;        /* mlp_hidden_dim; compiled set: 128                                 */
  int32_t num_layers;    /* 3..INF_MAX_LAYERS (model.py:27); <= 7 with the view branch        */
  int32_t skip;          /* skip_layer_idx, 0 < skip < num_layers-1                           */
  int32_t table_dtype;   /* INF_TABLE_F32 | INF_TABLE_F16                                     */
  int32_t input_mode;    /* INF_INPUT_GATHER | INF_INPUT_DENSE                                */
  int32_t view_strategy; /* INF_VIEW_*                                                        */
  int32_t view_embed;    /* embed_size of the view-direction Fourier features; the embedding  */
                         /* d*embed*2 (+d) must fit 64 dims (d = 1 intrinsic, 3 extrinsic)     */
  int32_t view_include;  /* include_view_dir (0/1)                                            */
  int32_t view_bneck;    /* bottleneck_vec_dim, 1..128 (zero-padded to 128 inside)            */
  int32_t view_hidden;   /* directional_hidden_dim, 1..128 (zero-padded to 128 inside)        */
  int64_t n_faces;       /* rows of face_normals (intrinsic view strategy)                    */
  int64_t n_src_rays;    /* rows of the per-ray arrays when InfTensors.ray_idxs selects from them (else 0)  */
  int32_t ray_idx_dtype; /* dtype of ray_idxs (INF_IDX_*)                                      */
  int32_t path;          /* INF_PATH_*                                                         */
  float grad_scale;      /* multiplies every summed gradient; 1/(3*B_global) for the L1 mean  */
  float reserved0;
} InfDesc;

/* Device buffers of one call. */
typedef struct InfTensors {
  const void* table;          /* (V,k_pad) f32|f16 row-major, 16-byte aligned (inf_relayout_table) */
  const void* vids;           /* (B,3) i32|i64 vertex ids of the hit faces (bit-exact pass-through) */
  const float* bary;          /* (B,3) f32 barycentric weights                                      */
  const float* target;        /* (B,3) f32 expected_rgbs, training with the built-in L1 loss        */
  const float* grad_pred;     /* (B,3) f32 dLoss/dpred supplied by autograd (exclusive with target) */
  const float* unit_ray_dirs; /* (B,3) f32, view branch                                             */
  const void* face_idxs;      /* (B,)  i32|i64, intrinsic view branch                               */
  const float* face_normals;  /* (F,3) f32, intrinsic view branch                                   */
  const void* ray_idxs;       /* optional (B,) i32|i64 batch selection (ray_dataloader.py:103-145 folded into the   */
                              /* kernels): ray r of this call reads row ray_idxs[r] of vids / bary / target /       */
                              /* unit_ray_dirs / face_idxs, which are then the WHOLE split's arrays; pred, grad_pred */
                              /* stay in batch order                                                                */
  const void* table_image;    /* optional fp16 operand image of the table (inf_build_table_image); enables the      */
                              /* vertex path; the workspace must then be sized with desc.n_rows set                 */
  const int8_t* col_exp;      /* optional (k_pad,) column exponents of the table (inf_column_exponents): the table  */
                              /* holds E[:,j] * 2^col_exp[j]; `packed` must have been built with the same exponents */
  const void* packed;         /* operand images written by inf_pack_weights                         */
  float* pred;                /* (B,3) f32 out                                                      */
  void* workspace;            /* >= inf_workspace_bytes(desc, training)                             */
  size_t workspace_bytes;
  const void* proj;           /* optional, inference only: the vertex projection P of THESE weights and THIS table,   */
                              /* written by inf_project_vertices_raw (inf_projection_bytes(n_rows) bytes, the         */
                              /* library's layout).  inf_forward then takes the vertex path whatever the batch size   */
                              /* and launches the per-ray kernel only -- a frozen model (renderer.py:36-146,          */
                              /* render_video.py:117-134) projects its vertices once, not once per chunk.             */
} InfTensors;

/* Gradient outputs of inf_forward_backward. */
typedef struct InfGrads {
  float* flat;       /* inf_num_params(desc) f32: all parameter gradients, state_dict order   */
  float* loss_sums;  /* 3 f32: sum|pred-target|, sum (pred-target)^2 over this call's rays (over all ranks' rays after  */
                     /* inf_train_step_comm), and the first times desc.grad_scale = the mean L1 loss when grad_scale  */
                     /* is 1/(3 B_global) -- the training log needs no extra kernel                                    */
} InfGrads;

/* Optimiser state of inf_train_step / inf_adam_pack_step: torch.optim.Adam defaults semantics (no weight decay, no
 * amsgrad).  params / exp_avg / exp_avg_sq are flat f32 vectors of inf_num_params(desc) elements in state_dict order. */
#define INF_ADAM_STEP_AUTOINC 1 /* step_dev counts completed updates; the kernel increments it (inf_train_step[_comm] only) */
typedef struct InfAdam {
  float* params;      /* master parameters, updated in place                                          */
  float* exp_avg;     /* first moment                                                                  */
  float* exp_avg_sq;  /* second moment                                                                 */
  void* packed;       /* operand images (inf_pack_weights layout); the live entries are rewritten      */
  const int8_t* col_exp; /* column exponents of the table these operand images are used with (null: none)        */
  int32_t* step_dev;  /* optional DEVICE counter (lets a CUDA graph replay the step with a changing t); when null,    */
                      /* `step` below is used.  It holds t -- or, with INF_ADAM_STEP_AUTOINC, the number of COMPLETED  */
                      /* updates: inf_train_step[_comm] then uses t = *step_dev + 1 and writes it back itself          */
  float lr, beta1, beta2, eps;
  int32_t step;       /* 1-based step count t of THIS update (bias corrections 1 - beta^t)             */
  int32_t flags;      /* INF_ADAM_*                                                                     */
} InfAdam;

/* Peer-memory exchange of the multi-GPU training step (one process per GPU on one node; replaces the reference's
 * nn.DataParallel, train.py:46-48, and the NCCL allreduce + separate optimiser launch of a plain data-parallel step).
 * Every rank creates one exchange region (inf_comm_create: device memory + an IPC handle), the ranks swap the 64-byte
 * handles by any means (torch.distributed all_gather_object in the Python mirror), and each opens its peers' regions
 * (inf_comm_open).  inf_train_step_comm then sums the per-rank gradients INSIDE the reduction kernel: every rank writes
 * its gradient into its own region, publishes a flag in every peer's region with a system-scope release store over
 * NVLink, waits for the peers' flags, reads their slots and adds them in rank order -- so all ranks obtain bit-identical
 * sums -- and applies Adam + the operand refresh in the same launch.  All ranks must make the same sequence of calls. */
#define INF_MAX_PEERS 8
#define INF_IPC_HANDLE_BYTES 64
typedef struct InfComm {
  int32_t world;                /* ranks on this node, 2..INF_MAX_PEERS (<= 1: no exchange)            */
  int32_t rank;
  void* regions[INF_MAX_PEERS]; /* regions[q]: rank q's exchange region as mapped in THIS process      */
  size_t region_bytes;          /* >= inf_comm_region_bytes(desc)                                      */
} InfComm;
size_t inf_comm_region_bytes(const InfDesc* desc);
int inf_comm_create(size_t bytes, void** region, void* ipc_handle_out /* INF_IPC_HANDLE_BYTES */);
int inf_comm_open(const void* ipc_handle, void** mapped);
int inf_comm_close(void* mapped);
int inf_comm_free(void* region);

/* version / errors */
int inf_version(void);
/* kernels of this library enqueued by the process so far (bench.py's gpu_launches is a difference of two readings;
 * a CUDA-graph replay enqueues without passing through the library and is not counted) */
long long inf_launch_count(void);
/* value of the INF_DEBUG measurement knob read at first use (bits 1|2|4 skip work and produce WRONG gradients) */
int inf_debug_flags(void);
const char* inf_last_error(void);

/* sizes (host-only, no CUDA calls) */
int64_t inf_num_params(const InfDesc* desc);  /* number of f32 in the flat parameter vector   */
int32_t inf_num_param_tensors(const InfDesc* desc);
/* offset (in f32 elements) and shape of parameter tensor `i` (state_dict order,
 * model.py:42-96: layers.0.0.weight, layers.0.0.bias, ..., layers.<skip>.Lx.weight, .Lx.bias,
 * .Ly.weight, .Ly.bias, ..., then directional_mlp.0/.2 for the view branch). */
int inf_param_info(const InfDesc* desc, int32_t i, int64_t* offset, int32_t* rows, int32_t* cols);
size_t inf_packed_bytes(const InfDesc* desc);
size_t inf_workspace_bytes(const InfDesc* desc, int training);

/* (V,k) f32 table with arbitrary row stride -> (V,k_pad) f32|f16, zero padded columns. */
int inf_relayout_table(const float* src, int64_t n_rows, int32_t k, int64_t src_row_stride,
                       void* dst, int32_t k_pad, int32_t dst_dtype, void* stream);

/* fp16 range guard.  The kernels round the interpolated embedding (and the table image) to fp16, which is harmless for
 * the reference's default `standard` rescale (|E| <= 1, mesh.py:99-102) but loses `unscaled` / HKS / functional-map
 * tables whose columns sit far from 1 (mesh.py:53-108): below 6e-5 they fall into fp16's subnormals, above 65504 they
 * overflow.  inf_column_exponents writes, per column j, the power of two that brings max|E[:,j]| into [0.5, 1) when it
 * lies outside [2^-6, 2^6] (0 otherwise: tables in the usual range stay bit for bit); inf_relayout_table_scaled stores
 * E[:,j] * 2^col_exp[j]; inf_pack_weights_scaled stores W0[:,j] * 2^-col_exp[j] (and Wy alike) in the operand images, so
 * every product x_j W[n,j] is unchanged -- scaling by powers of two is exact -- and the weight gradients of layer 0 / Ly
 * are scaled back in the reduction kernel.  `scratch` = k_pad u32 of device memory. */
int inf_column_exponents(const float* src, int64_t n_rows, int32_t k, int64_t src_row_stride, int32_t k_pad, int8_t* col_exp,
                         void* scratch, void* stream);
int inf_relayout_table_scaled(const float* src, int64_t n_rows, int32_t k, int64_t src_row_stride, const int8_t* col_exp,
                              void* dst, int32_t k_pad, int32_t dst_dtype, void* stream);
int inf_pack_weights_scaled(const InfDesc* desc, const float* const* params, const int8_t* col_exp, void* packed, void* stream);

/* Table operand image for the vertex path: (V,k_pad) f32|f16 row-major table -> fp16 [ceil(V/128)][k_pad/64] panels of
 * [128 rows][64 columns] in the tensor core's SWIZZLE_128B layout (values rounded to fp16 once).  Built once where the
 * reference moves the table to the device (ray_dataloader.py:73, renderer.py:43). */
size_t inf_table_image_bytes(int64_t n_rows, int32_t k_pad);
int inf_build_table_image(const void* table, int32_t table_dtype, int64_t n_rows, int32_t k_pad, void* image, void* stream);

/* The per-vertex projection on its own: out (V, 256) f16 row-major = [E W0^T | E Wy^T], i.e. layer 0's and the skip
 * branch's pre-activations (without bias) at every VERTEX (model.py:98-104, layers.py:60-62 evaluated at bary = e_i).
 * Needs t->table_image, t->packed and a workspace of inf_workspace_bytes(desc with n_rows set, 0). */
int inf_project_vertices(const InfDesc* desc, const InfTensors* t, void* out, void* stream);
/* The same projection kept in the layout the per-ray kernel gathers from (InfTensors.proj): `proj` receives
 * inf_projection_bytes(desc->n_rows) bytes.  Needs t->table_image and t->packed; no workspace. */
size_t inf_projection_bytes(int64_t n_rows);
int inf_project_vertices_raw(const InfDesc* desc, const InfTensors* t, void* proj, void* stream);

/* fp32 master parameters -> operand images.  `params[i]` = device pointer of tensor i. */
int inf_pack_weights(const InfDesc* desc, const float* const* params, void* packed, void* stream);

/* gather -> interpolate -> MLP -> RGB (inference; no gradients, no workspace needed). */
int inf_forward(const InfDesc* desc, const InfTensors* t, void* stream);

/* fused training pass: pred, loss sums and all parameter gradients (scaled by grad_scale). */
int inf_forward_backward(const InfDesc* desc, const InfTensors* t, const InfGrads* g, void* stream);

/* inf_forward_backward followed by an Adam step on the flat master parameters and the refresh of the operand images,
 * all inside the gradient-reduction kernel (single-GPU training step in three launches).  g->flat still receives the
 * gradient.  `t->packed` (read by this step's forward) and `adam->packed` (written at the end) may be the same buffer. */
int inf_train_step(const InfDesc* desc, const InfTensors* t, const InfGrads* g, const InfAdam* adam, void* stream);

/* inf_train_step across the GPUs of a node: the gradient exchange runs inside the reduction kernel over peer memory
 * (see InfComm).  g->flat and g->loss_sums receive the SUMS over ranks (desc->grad_scale carries 1/(3*B_global)). */
int inf_train_step_comm(const InfDesc* desc, const InfTensors* t, const InfGrads* g, const InfAdam* adam, const InfComm* comm,
                        void* stream);

/* Adam + operand refresh from an already reduced gradient (multi-GPU: call after the allreduce of g->flat). */
int inf_adam_pack_step(const InfDesc* desc, float* flat_grad, const InfAdam* adam, void* stream);

/* Materialising gather, mesh.py:313-324 get_k_eigenfunc_vec_vals as one kernel:
 * out (B,k) f32 = sum_j bary[:,j] * table[vids[:,j], :k].  For callers that need the dense embedding
 * (UV baking, notebooks); the fused path above never builds it. */
int inf_gather_interpolate(const void* table, int32_t table_dtype, int32_t k, int32_t k_pad,
                           const void* vids, int32_t idx_dtype, const float* bary, int64_t n_rays,
                           float* out, void* stream);
/* the same for a table stored with column exponents: the output is scaled back to the reference's values */
int inf_gather_interpolate_scaled(const void* table, int32_t table_dtype, int32_t k, int32_t k_pad, const int8_t* col_exp,
                                  const void* vids, int32_t idx_dtype, const float* bary, int64_t n_rays,
                                  float* out, void* stream);

/* Measurement hooks (bench.py): while enabled, inf_forward / inf_forward_backward bracket their
 * kernels with library-owned CUDA events on the launching stream (up to 256 calls, no sync);
 * inf_profile_read synchronises on those events and returns the average device time per call of the
 * forward/backward kernel, the wgrad kernel and the reduction.  Single-threaded use only. */
int inf_profile_enable(int on);
int inf_profile_read(float* main_ms, float* wgrad_ms, float* reduce_ms, int* calls);
/* The same per stage: stage_ms[0..5] = incidence lists (CSR), vertex projection, per-ray forward/backward kernel,
 * vertex-gradient gather, weight-gradient kernel, reduction (+Adam); stages a call did not run read 0. */
#define INF_PROFILE_STAGES 6
int inf_profile_read_stages(float* stage_ms, int* calls);
/* 1 if the call described by (desc, tensors) takes the vertex path, 0 for the per-ray path (host only). */
int inf_uses_vertex_path(const InfDesc* desc, const InfTensors* t);

/* Self-test hook used by tests/: D(128,n) = A(128,k) * B(n,k)^T on the tensor cores with the
 * operand-layout variant `mode` (0: A,B K-major; 1: A MN-major; 2: A,B MN-major; 3: B MN-major).  A and
 * B are f32 row-major on the device ((128,k)/(n,k) for K-major, (k,128)/(k,n) for MN-major). */
int inf_debug_umma(int mode, int n, int k, const float* a, const float* b, float* d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* INF_FIELD_H_ */
