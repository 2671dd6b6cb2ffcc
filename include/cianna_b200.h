/*
 * cianna_b200.h - C-ABI of the B200-native compute core (libcianna_b200.so).
 *
 * This is the drop-in boundary for CIANNA's C_CUDA back-end: every entry point below
 * replaces one (group of) function(s) the reference binds through its per-layer
 * `forward`/`backprop` function pointers and `cuda_*` helpers.  Signatures carry only
 * plain pointers, sizes and POD descriptors (no C++/torch types) so that host code in
 * C (cianna_b200/host, or a patched upstream src/*.c) can call it.
 *
 * Reference interface replaced (paths relative to the upstream tree):
 *   src/prototypes.h:217-295            the `cuda_*` extern "C" surface
 *   src/structs.h:88-262                the per-precision kernel function tables
 *   src/cuda/cuda_{conv,pool,norm,lrn,dense}_layer.cu, cuda_activ_functions.cu, cuda_main.cu
 *
 * Conventions
 *   - Every function returns 0 on success, non-zero on failure (cb200_last_error() gives
 *     the text).  Nothing here falls back to the CPU: without a CUDA device every compute
 *     entry point fails with CB200_ERR_NO_DEVICE.
 *   - `stream` is a cudaStream_t passed as void* (NULL = the core's default compute stream).
 *   - Device tensors use the core's INTERNAL layout: channels-last
 *         act[b][y][x][c],  c in [0, Cp),  Cp = cb200_round_channels(C) (multiple of 8),
 *     pad channels are kept at zero.  The reference's layouts ([C][B][H*W] activations,
 *     [B][C*H*W+1] dataset rows, [N][K+pad] filters) only exist at the boundary and are
 *     converted by the cb200_import_* / cb200_export_* entry points.
 *   - dtype is the storage/compute type of activations and deltas; accumulation is always
 *     FP32 (reference modes FP32C_FP32A, FP16C_FP32A, BF16C_FP32A, src/structs.h:70).
 */
#ifndef CIANNA_B200_H
#define CIANNA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ enums */
typedef enum { CB200_FP32 = 0, CB200_FP16 = 1, CB200_BF16 = 2 } cb200_dtype;
/* same order as the reference's activation_functions_enum (src/structs.h:33) */
typedef enum { CB200_RELU = 0, CB200_LOGISTIC = 1, CB200_SOFTMAX = 2, CB200_YOLO = 3, CB200_LINEAR = 4 } cb200_activ_type;
typedef enum { CB200_POOL_MAX = 0, CB200_POOL_AVG = 1 } cb200_pool_type;

enum {
	CB200_OK = 0,
	CB200_ERR_NO_DEVICE = 1,
	CB200_ERR_CUDA = 2,
	CB200_ERR_UNSUPPORTED = 3,
	CB200_ERR_ARG = 4,
	CB200_ERR_NCCL = 5
};

/* Activation applied in a producing kernel's epilogue (forward) or as the
 * "previous->deriv_activation" hook of a backward kernel.
 * Replaces: ReLU/logistic/linear kernels + wrappers, src/cuda/cuda_activ_functions.cu:37-111,203-270. */
typedef struct {
	int   type;        /* cb200_activ_type (SOFTMAX / YOLO are separate passes) */
	float leak;        /* RELU leaking factor (default 0.05) */
	float saturation;  /* RELU / LOGISTIC saturation (800 / 6) */
	float beta;        /* LOGISTIC beta */
} cb200_activ;

/* ------------------------------------------------------------------ runtime */
/* Replaces init_cuda (src/cuda/cuda_main.cu:922-1067). device < 0: keep current device. */
int  cb200_init(int device);
int  cb200_device_count(void);
const char* cb200_last_error(void);
const char* cb200_version(void);
/* which implementation the last conv/dense call dispatched to: "tcgen05" or "simt" */
const char* cb200_last_conv_impl(void);
/* bit 0: force the generic SIMT kernels (parity cross-checks of the tcgen05 path); bit 1: keep the tcgen05 path but
 * route every layer through the per-tap kernel instead of the halo-reuse kernel (A/B checks); bit 2: enable the
 * 2-CTA cluster variant of the per-tap kernel (filter halves TMA-multicast to both CTAs; measured neutral, off by
 * default); bit 3 / bit 4: force the CTA-pair kernels (cta_group::2) on / off - forward / data gradient with N tiles of
 * 256 on layers with enough tiles (default on, env CB200_CTA_PAIR=0) and the weight gradient of layers with more than
 * 128 channels on both sides (default off, env CB200_WGRAD_PAIR=1); 0 = auto */
void cb200_force_simt(int on);
/* number of kernels launched by this library since the last reset */
long long cb200_launch_count(int reset);

/* opt-in timing of kernel families with CUDA event pairs recorded on the launching stream (no per-layer sync:
 * replaces the always-on cuda_perf_eval_* of src/cuda/cuda_main.cu:533-588).  family: 0/1/2 = tcgen05 conv
 * forward / data-grad / weight-grad, 3/4/5 = same on the SIMT kernels, 6 = pooling, 7 = group-norm.
 * `work` accumulates the algorithmic FLOPs (conv) or bytes (pool / norm) of the timed launches. */
void cb200_profile_enable(int on);
void cb200_profile_reset(void);
int  cb200_profile_collect(int family, double* ms, double* work, long long* launches);

int  cb200_round_channels(int c);
size_t cb200_dtype_size(int dtype);

/* memory: replaces cuda_create_table*, cuda_free_table, cuda_put/get_table*, cuda_set_mem_value
 * (src/cuda/cuda_main.cu:108-379) */
int  cb200_malloc(void** dev_ptr, size_t bytes);          /* zero-initialised */
int  cb200_free(void* dev_ptr);
int  cb200_host_alloc(void** host_ptr, size_t bytes);     /* pinned host memory */
int  cb200_host_free(void* host_ptr);
int  cb200_memset(void* dev_ptr, int byte, size_t bytes, void* stream);
int  cb200_h2d(void* dev_dst, const void* host_src, size_t bytes, void* stream);
int  cb200_d2h(void* host_dst, const void* dev_src, size_t bytes, void* stream);
int  cb200_d2d(void* dev_dst, const void* dev_src, size_t bytes, void* stream);
int  cb200_stream_create(void** stream);
/* a stream of the lowest priority; the core's default compute stream has the highest, so kernels put here (the host
 * library's weight gradients) only take the SMs the critical path leaves free */
int  cb200_stream_create_low_priority(void** stream);
int  cb200_stream_destroy(void* stream);
int  cb200_stream_sync(void* stream);                     /* NULL: default compute stream */
int  cb200_device_sync(void);
/* make `stream` wait for everything enqueued so far on `on_stream` */
int  cb200_stream_wait(void* stream, void* on_stream);
/* events (replace cuda_perf_eval_*, src/cuda/cuda_main.cu:533-588, without forcing a sync per layer) */
int  cb200_event_create(void** ev);
int  cb200_event_destroy(void* ev);
int  cb200_event_record(void* ev, void* stream);
int  cb200_stream_wait_event(void* stream, void* ev);   /* work queued on `stream` after this call waits for the event */
int  cb200_event_sync(void* ev);                      /* host waits for the work recorded before the event */
int  cb200_event_elapsed_ms(void* ev_start, void* ev_stop, float* ms);  /* synchronises on ev_stop */

/* typed conversion helpers: FP32 host/device arrays <-> dtype arrays on device.
 * Replaces cuda_convert_table / cuda_get_table_to_FP32 (src/cuda/cuda_main.cu:153-300). */
int  cb200_cast_from_f32(void* dev_dst, int dtype, const float* dev_src, size_t n, void* stream);
int  cb200_cast_to_f32(float* dev_dst, const void* dev_src, int dtype, size_t n, void* stream);
/* host-side conversion with round-toward-zero, as the reference's dataset conversion does
 * (src/cuda/cuda_main.cu:790,813) */
int  cb200_host_cast_from_f32(void* host_dst, int dtype, const float* host_src, size_t n);
/* the way back on the host (exact), as cuda_get_typed_host_table does (src/cuda/cuda_main.cu:245-256) */
int  cb200_host_cast_to_f32(float* host_dst, const void* host_src, int dtype, size_t n);
/* the same round-toward-zero conversion on the device: an FP32 batch staged in device memory -> `dtype`
 * (moves the dataset conversion loop of cuda_convert_batched_table, src/cuda/cuda_main.cu:355-371, off the host) */
int  cb200_cast_from_f32_rz(void* dev_dst, int dtype, const float* dev_src, size_t n, void* stream);
/* FP32 sample rows [rows][src_row] (device) -> dataset rows [rows][dst_row] in `dtype` (device), dst_row >= src_row, the
 * extra slots (the bias slot of an input row, src/auxil.c:320-329) set to tail_value; round toward zero.
 * The per-element host loop of python_module.c:140-165 + copy_to_FP16 (cuda_main.cu:108-113) as one kernel. */
int  cb200_dataset_pack(void* dev_dst, int dtype, const float* dev_src, size_t rows, size_t src_row, size_t dst_row,
                        float tail_value, void* stream);
/* dataset shuffle on the device: row i of the source batches -> row index[i] of the destination batches; both are
 * device arrays of nb_batch device pointers, each batch [batch_size][row_bytes]; index (device int32 [n_rows]) may be
 * NULL for a plain copy (the way back).  Replaces shfl_kern_* + get_back_shuffle_* (src/cuda/cuda_main.cu:590-666). */
int  cb200_rows_permute(void* const* dst_batches_dev, void* const* src_batches_dev, const int* index_dev, long long n_rows,
                        int batch_size, size_t row_bytes, void* stream);

/* ------------------------------------------------------------------ layout import / export */
/* Dataset rows -> internal.  src: device array [batch][C*H*W + 1] in `dtype` (reference dataset
 * layout incl. the trailing bias slot, src/auxil.c:320-329) ; dst: act[batch][H][W][Cp].
 * Replaces the first-layer branch of im2col (src/cuda/cuda_conv_layer.cu:36-103, bias_in=1). */
int  cb200_import_input(void* dst, const void* src, int dtype, int batch, int c, int h, int w, void* stream);
/* First-layer variant for inputs with very few channels (RGB / grey images): writes, for every OUTPUT pixel of the
 * first convolution, its receptive field as one row [c*f_h*f_w values in the reference's column order c*taps + tap |
 * bias_value | zero pad] of cb200_patch_width(c, f_h, f_w) elements: dst[batch][out_h][out_w][patch_width].
 * This is the one place where the reference's explicit unrolling (im2col_kernel with bias_in = 1,
 * src/cuda/cuda_conv_layer.cu:36-103) is kept, fused into the layout import: with 3 input channels the unrolled row
 * (28 values) is no larger than the layer's own output row, so the layer is HBM-bound either way, and it lets the first
 * layer use the same tensor-core GEMM kernels as every other layer. */
int  cb200_patch_width(int c, int f_h, int f_w);
int  cb200_import_input_patches(void* dst, const void* src, int dtype, int batch, int c, int h, int w,
                                int f_h, int f_w, int stride_h, int stride_w, int pad_h, int pad_w,
                                int out_h, int out_w, float bias_value, void* stream);
/* Reference activation layout [C][B][H*W] (FP32, device) <-> internal (dtype). */
int  cb200_import_cbhw(void* dst, int dtype, const float* src, int batch, int c, int h, int w, void* stream);
int  cb200_export_cbhw(float* dst, const void* src, int dtype, int batch, int c, int h, int w, void* stream);
/* pool argmax map: internal uint8 [B][Ho][Wo][Cp] -> reference int32 [C][B][Ho*Wo] */
int  cb200_export_pool_map(int32_t* dst, const uint8_t* src, int batch, int c, int h, int w, void* stream);
/* dense layout [B][n+1] (reference, FP32, bias node last) <-> internal [B][1][1][Cp(n)] */
int  cb200_export_dense(float* dst, const void* src, int dtype, int batch, int n, float bias_node, void* stream);

/* ------------------------------------------------------------------ convolution */
/* Geometry of one conv layer. Replaces conv_param (src/structs.h:392-415). */
typedef struct {
	int dtype;
	int batch;            /* net->batch_size */
	int length;           /* net->length: samples >= length are forced to zero by non-linear activations */
	int in_c, in_h, in_w;
	int out_c, out_h, out_w;
	int f_h, f_w;
	int stride_h, stride_w;
	int pad_h, pad_w;
	float bias_value;     /* constant input of the bias column (layer->bias_value) */
	cb200_activ activ;    /* this layer's activation (fused in the forward epilogue) */
	int input_is_patches; /* 1: `x` is the patch tensor made by cb200_import_input_patches (first layer with few
	                         input channels); the layer then runs as a 1x1 GEMM over cb200_patch_width() columns
	                         whose last real column is the bias input, w_fwd/grad are [out_c][patch_width];
	                         2: `x` is the dataset batch itself, [batch][in_c*in_h*in_w + 1] values of `dtype`: the patch
	                         rows are built in shared memory inside the forward / weight-gradient kernels and never
	                         touch HBM (allowed when cb200_conv_first_direct() returns 1; weight buffers as in mode 1) */
	/* third dimension and internal padding (src/structs.h:392-415: f_size[2], stride[2], padding[2], int_padding[3]); all
	 * zero for a plain 2-D layer (0 depth / stride means 1).  Activations are then [batch][D][H][W][Cp], the filter taps
	 * run depth-major (tap = (kz*f_h + ky)*f_w + kx, upstream's column order), and internal padding q-1 puts input pixel i
	 * at position i*q of the zero-stuffed grid the filter slides over (transposed convolution,
	 * src/cuda/cuda_conv_layer.cu:66-68).  Such layers run on the generic CUDA-core kernels (conv_simt.cu). */
	int in_d, out_d, f_d, stride_d, pad_d;
	int ipad_w, ipad_h, ipad_d;
} cb200_conv_desc;
/* 1 when a first layer described by `d` (input_is_patches != 0) can run in mode 2: 16-bit compute type, 3x3 filters on
 * 1-3 channels or 5x5 on one channel, 8..64 filters, tensor-core path not disabled by cb200_force_simt. */
int cb200_conv_first_direct(const cb200_conv_desc* d);

/* Compute-side weights of one conv (or dense) layer, all device pointers owned by the caller:
 *   master : FP32 [out_c][k_ref] in the REFERENCE layout (k_ref = f_h*f_w*in_c + 1, column order
 *            c-major / tap-minor, bias column last; src/cuda/cuda_conv_layer.cu:91-94); this is
 *            what save/load files hold (src/conv_layer.c:491-526)
 *   moment : FP32 [out_c][k_ref] momentum buffer ("update", src/structs.h:413)
 *   w_fwd  : dtype [out_c][f_h*f_w][in_cp]   K-major operand of the forward implicit GEMM
 *   w_bwd  : dtype [in_c][f_h*f_w][out_cp]   180-degree rotated + transposed operand of the
 *            data-gradient GEMM (replaces cuda_rotate_filter_matrix, cuda_conv_layer.cu:106-128)
 *   bias_w : FP32 [out_c] = master[f][k_ref-1]
 *   grad   : FP32 [out_c][f_h*f_w][in_cp] raw weight gradient sum_m col*delta (all-reduced in DP)
 *   grad_b : FP32 [out_c] raw bias-column gradient sum_m delta
 */
typedef struct {
	float* master;
	float* moment;
	void*  w_fwd;
	void*  w_bwd;
	float* bias_w;
	float* grad;
	float* grad_b;
} cb200_conv_weights;

size_t cb200_conv_wfwd_elems(const cb200_conv_desc* d);
size_t cb200_conv_wbwd_elems(const cb200_conv_desc* d);
size_t cb200_conv_grad_elems(const cb200_conv_desc* d);
size_t cb200_conv_master_elems(const cb200_conv_desc* d);

/* (re)build w_fwd / w_bwd / bias_w from master. Replaces cuda_master_weight_copy
 * (src/cuda/cuda_main.cu:433-452) + the per-step filter rotation. */
int cb200_conv_prepare_weights(const cb200_conv_desc* d, const cb200_conv_weights* w, void* stream);

/* y = act( conv(x, W) + bias_value*W[:,bias] ).  Replaces cuda_forward_conv_layer
 * (src/cuda/cuda_conv_layer.cu:319-423: cast + im2col + cublasGemmEx + activation). */
int cb200_conv_forward(const cb200_conv_desc* d, const cb200_conv_weights* w,
                       const void* x, void* y, void* stream);

/* dx = fullconv(dy, rot(W)) * act'(prev).  `prev_activ`/`prev_out` describe the PREVIOUS layer's
 * activation and activated output (NULL prev_out or LINEAR: no hook).  Replaces the error-propagation
 * half of cuda_backward_conv_layer (cuda_conv_layer.cu:456-541). */
int cb200_conv_backward_data(const cb200_conv_desc* d, const cb200_conv_weights* w,
                             const void* dy, void* dx,
                             const cb200_activ* prev_activ, const void* prev_out, void* stream);

/* grad = sum over batch/pixels of im2col(x)^T * dy (raw, no lr), grad_b = sum dy.
 * Replaces the cublasGemmEx at cuda_conv_layer.cu:551-557 (optimizer split out so that the raw
 * gradient can be all-reduced across GPUs before momentum is applied). */
int cb200_conv_backward_weights(const cb200_conv_desc* d, const cb200_conv_weights* w,
                                const void* x, const void* dy, void* stream);
/* same, when grad_b was already produced by the kernel that wrote dy (cb200_norm_backward's dx_colsum) */
int cb200_conv_backward_weights_ex(const cb200_conv_desc* d, const cb200_conv_weights* w,
                                   const void* x, const void* dy, int have_grad_b, void* stream);

/* Optimizer hyper-parameters live in a small device buffer (a new learning rate is one small copy, no kernel argument
 * changes, and the whole-network update plan keeps pointing at it): hyper[0]=lr/batch_total, [1]=momentum, [2]=lr*weight_decay,
 * [3]=TC_scale_factor, [4]=lr (plain). */
#define CB200_HYPER_LEN 8
/* moment = hyper0*grad + mom*moment ; moment += wd_lr*master*S ; master -= moment/S ; then refresh
 * w_fwd / w_bwd / bias_w.  Replaces the alpha/beta of the wgrad GEMM + cuda_update_weights
 * (cuda_conv_layer.cu:551-560, cuda_main.cu:455-475) + next step's cuda_master_weight_copy. */
int cb200_conv_update(const cb200_conv_desc* d, const cb200_conv_weights* w, const float* hyper,
                      int is_pivot, void* stream);

/* ---- the optimizer sweep of a whole network in three launches (update_plan.cu): every conv layer's
 * cb200_conv_update and every group-norm layer's cb200_norm_reduce_update / cb200_norm_update, bit-identical results.
 * A plan captures the device pointers of the layers it is given: rebuild it when they change (or when a layer is frozen).
 * Layers cb200_update_plan_accepts() refuses (first layer on patch rows, filters whose FP32 row pair exceeds 96 KB of
 * shared memory) and dense layers keep their own cb200_conv_update / cb200_dense_update call. */
typedef struct {
	const float* d_gamma; const float* d_beta;   /* [batch][nb_group] per-sample gradients (reduce = 1) */
	float* gsum;                                 /* [2][nb_group]: written when reduce = 1, read (all-reduced sums) when 0 */
	float* gamma; float* beta; float* gamma_upd; float* beta_upd;
	int batch, nb_group, set_off, reduce;
} cb200_norm_update_ref;
int cb200_update_plan_accepts(const cb200_conv_desc* d);
int cb200_update_plan_create(void** plan, int dtype, const cb200_conv_desc* const* conv_desc, const cb200_conv_weights* const* conv_w,
                             int n_conv, const cb200_norm_update_ref* norms, int n_norm);
int cb200_update_plan_run(const void* plan, const float* hyper, void* stream);
int cb200_update_plan_destroy(void* plan);

/* ------------------------------------------------------------------ dense */
/* A dense layer runs through the convolution entry points above with a descriptor whose filter covers
 * the whole input map (f_h = in_h, f_w = in_w, no padding, 1x1 output): the reference's flatten order
 * flat[b][c*A + a] (cuda_flat_dense, src/cuda/cuda_dense_layer.cu:33-54) is exactly the conv column order
 * c*taps + tap, so no flatten / reroll pass exists.  Only the FP32 master layout differs: upstream keeps
 * W[in_size][n + 1] (input-major, pivot column last; src/dense_layer.c:253-268).  These two variants of
 * prepare / update read that layout; `master`/`moment` then hold in_size*(n+1) floats and the pivot
 * column is never touched (upstream: is_pivot, cuda_dense_layer.cu:489-495).
 * Replaces cuda_forward/backward_dense_layer's three cublasGemmEx calls (cuda_dense_layer.cu:370,420,489). */
int cb200_dense_prepare_weights(const cb200_conv_desc* d, const cb200_conv_weights* w, void* stream);
int cb200_dense_update(const cb200_conv_desc* d, const cb200_conv_weights* w, const float* hyper, void* stream);

/* ------------------------------------------------------------------ pooling */
typedef struct {
	int dtype;
	int batch;
	int c, in_h, in_w, out_h, out_w;
	int p_h, p_w, stride_h, stride_w, pad_h, pad_w;
	int pool_type;        /* cb200_pool_type */
	int length;
	cb200_activ activ;    /* layer activation applied to the pooled output (LINEAR / RELU / LOGISTIC) */
	/* third dimension (src/structs.h:418-436: p_size[2], stride[2], padding[2]); all zero for a 2-D layer (0 = 1).
	 * Activations are then [batch][D][H][W][Cp]; map value (z*p_h + y)*p_w + x (cuda_pool_layer.cu:108). */
	int in_d, out_d, p_d, stride_d, pad_d;
} cb200_pool_desc;

/* Replaces cuda_forward_pool_layer (src/cuda/cuda_pool_layer.cu:429-492). map: uint8 [B][Ho][Wo][Cp],
 * local window index y*p_w+x of the first strict maximum, 255 when the window is empty (reference: -1). */
int cb200_pool_forward(const cb200_pool_desc* d, const void* x, void* y, uint8_t* map, void* stream);
/* Replaces cuda_backward_pool_layer (cuda_pool_layer.cu:495-547) incl. the previous->deriv_activation hook. */
int cb200_pool_backward(const cb200_pool_desc* d, const void* dy, const uint8_t* map, void* dx,
                        const cb200_activ* prev_activ, const void* prev_out, void* stream);

/* ------------------------------------------------------------------ group normalisation */
typedef struct {
	int dtype;
	int batch, length;
	int c, h, w;
	int group_size, nb_group, set_off;
	float eps;            /* 1e-3 upstream */
} cb200_norm_desc;

/* mean/var: FP32 [batch][nb_group]; gamma/beta: FP32 [nb_group] (device).
 * Replaces cuda_forward_norm_layer (src/cuda/cuda_norm_layer.cu:361-397). */
size_t cb200_norm_workspace_bytes(const cb200_norm_desc* d);   /* FP64 partial sums, caller-owned */
/* The forward / backward entry points below (and the fused norm + pool pair) run as two launches over the whole batch -
 * statistics, then apply with the finalize step folded into its blocks - (on = 0, default), or walk the batch in L2-sized chunks (on = 1, env CB200_GN_PIPELINE=1): launch i
 * holds the statistics blocks of chunk i and the apply blocks of chunk i-1, which re-read their chunk from L2 instead of
 * HBM - same arithmetic, measured slower on B200 (csrc/norm.cu), kept as a tested option.
 * chunk_kb: bytes read by the statistics blocks of one launch (0: keep; default 24 MB, env CB200_GN_CHUNK_MB);
 * blocks_per_sm: blocks per SM and role in one launch (0: keep; default 6, env CB200_GN_BLOCKS_PER_SM). */
void cb200_norm_set_pipeline(int on, int chunk_kb, int blocks_per_sm);
/* grid sizing of the group-norm kernels: blocks wanted per SM over the whole batch (default 8) and the most pixels a
 * block takes (default 4096); 0 keeps a value.  Measurement hook (scripts/exp/gn_apply_sweep.py). */
void cb200_norm_set_tuning(int want_blocks_per_sm, int ppb_max);
int cb200_norm_forward(const cb200_norm_desc* d, const void* x, void* y,
                       const float* gamma, const float* beta, float* mean, float* var,
                       void* workspace, void* stream);
/* ---- statistics pass taken out of HBM: the convolution in front of a group-norm layer accumulates the layer's
 * (sum, sum of squares) per (sample, group) in its epilogue, from the values it has just rounded for the store - the
 * numbers cuda_group_mean / cuda_group_var (src/cuda/cuda_norm_layer.cu:65-138) would read back - and leaves them in the
 * norm layer's FP64 workspace.  cb200_conv_forward_stats = cb200_conv_forward + that; *stats_done = 1 when the kernel that
 * ran could do it (tcgen05 kernels, group size 8 / 16 / a multiple of 32, tile rows of a sample >= sums per chunk), 0
 * when the workspace is untouched.  The _ex norm entry points take that flag: stats_ready = 1 skips their own
 * statistics launch and the workspace reset (forward: 3 -> 2 passes over the tensor; fused with the max-pool: 2 -> 1). */
int cb200_conv_forward_stats(const cb200_conv_desc* d, const cb200_conv_weights* w, const void* x, void* y,
                             const cb200_norm_desc* gn, void* gn_workspace, int* stats_done, void* stream);
/* 0: never; 1 (default, env CB200_GN_EPILOGUE_STATS): only where the sums hide behind the tensor pipe (K loops of >= 16
 * blocks: 3x3 filters on >= 128 channels - elsewhere they cost what the statistics pass costs, csrc/conv_tc.cu); 2: always */
void cb200_set_gn_epilogue_stats(int mode);
int cb200_norm_forward_ex(const cb200_norm_desc* d, const void* x, void* y,
                          const float* gamma, const float* beta, float* mean, float* var,
                          void* workspace, int stats_ready, void* stream);
/* d_gamma/d_beta: FP32 [batch][nb_group] per-sample sums (as upstream); dx includes the
 * previous->deriv_activation hook (prev_out == x of this layer).
 * Replaces cuda_backward_norm_layer's device half (cuda_norm_layer.cu:399-432). */
int cb200_norm_backward(const cb200_norm_desc* d, const void* x, const void* dy, void* dx,
                        const float* gamma, const float* mean, const float* var,
                        float* d_gamma, float* d_beta,
                        const cb200_activ* prev_activ, float* dx_colsum, void* workspace, void* stream);
/* dx_colsum (optional, FP32 [c]): receives sum over batch and pixels of dx per channel - the raw bias-column gradient
 * (grad_b) of the convolution that feeds this norm layer, produced for free while dx is written; the caller then uses
 * cb200_conv_backward_weights_ex(..., have_grad_b = 1) for that convolution. */
/* gamma_upd = mom*gamma_upd + lr*sum_b(d_gamma)/batch_total ; gamma -= gamma_upd/S (same for beta).
 * gsum: FP32 [2][nb_group] batch-summed (d_gamma, d_beta) (the buffer that is all-reduced in DP).
 * Replaces the host loop + 4 blocking memcpys of cuda_norm_layer.cu:434-457. */
int cb200_norm_reduce_grads(const cb200_norm_desc* d, const float* d_gamma, const float* d_beta,
                            float* gsum, void* stream);
int cb200_norm_update(const cb200_norm_desc* d, float* gamma, float* beta, float* gamma_upd, float* beta_upd,
                      const float* gsum, const float* hyper, void* stream);
/* cb200_norm_reduce_grads + cb200_norm_update in one launch, for runs without a gradient all-reduce between them
 * (single GPU): same arithmetic, gsum is still written. */
int cb200_norm_reduce_update(const cb200_norm_desc* d, const float* d_gamma, const float* d_beta, float* gsum,
                             float* gamma, float* beta, float* gamma_upd, float* beta_upd, const float* hyper, void* stream);

/* ---- group-norm followed by a max-pool over disjoint 2x2 windows, fused (no upstream counterpart: upstream runs
 * cuda_forward_norm_layer then cuda_forward_pool_layer and the reverse pair, cuda_norm_layer.cu:361-461,
 * cuda_pool_layer.cu:429-547).  Same results as cb200_norm_forward + cb200_pool_forward (values rounded to the
 * storage type before they are compared, so the argmax map is identical) and as cb200_pool_backward +
 * cb200_norm_backward, without ever writing or reading the full-resolution normalised tensor / its delta:
 * forward moves 2.4 instead of 4.4 passes over the input-sized tensor, backward 3.4 instead of 6.4.
 * Supported when cb200_norm_pool_fusable() returns 1: MAX pool, 2x2 window, stride 2, no padding, even input
 * size, LINEAR pool activation, nd->{c,h,w} == pd->{c,in_h,in_w}. */
int cb200_norm_pool_fusable(const cb200_norm_desc* nd, const cb200_pool_desc* pd);
int cb200_norm_pool_forward(const cb200_norm_desc* nd, const cb200_pool_desc* pd, const void* x, void* pooled,
                            uint8_t* pool_map, const float* gamma, const float* beta, float* mean, float* var,
                            void* workspace, void* stream);
int cb200_norm_pool_forward_ex(const cb200_norm_desc* nd, const cb200_pool_desc* pd, const void* x, void* pooled,
                               uint8_t* pool_map, const float* gamma, const float* beta, float* mean, float* var,
                               void* workspace, int stats_ready, void* stream);
/* d_pooled: delta of the pool OUTPUT; dx: delta of the norm INPUT (previous layer's derivative hook included). */
int cb200_norm_pool_backward(const cb200_norm_desc* nd, const cb200_pool_desc* pd, const void* x, const void* d_pooled,
                             const uint8_t* pool_map, void* dx, const float* gamma, const float* mean,
                             const float* var, float* d_gamma, float* d_beta, const cb200_activ* prev_activ,
                             float* dx_colsum, void* workspace, void* stream);
/* The same with the pool layer's OUTPUT and beta at hand: the backward reductions sum(d), sum(d*x) then read the pooled
 * delta and the pooled output only (x = (y - shift) / scale at the selected position) instead of the input-sized
 * tensor, its map and the pooled delta; groups where that inversion is ill-conditioned (|gamma| < 1e-4 or
 * |beta| > 16 |gamma|) and dead samples keep the gather from x.  pooled == NULL or beta == NULL: as above. */
int cb200_norm_pool_backward_ex(const cb200_norm_desc* nd, const cb200_pool_desc* pd, const void* x, const void* d_pooled,
                                const uint8_t* pool_map, void* dx, const float* gamma, const float* mean,
                                const float* var, float* d_gamma, float* d_beta, const cb200_activ* prev_activ,
                                float* dx_colsum, void* workspace, const void* pooled, const float* beta, void* stream);

/* ------------------------------------------------------------------ local response normalisation */
typedef struct {
	int dtype;
	int batch, length;
	int c, h, w;
	int range;
	float k, alpha, beta;
} cb200_lrn_desc;
/* Replaces cuda_forward/backward_lrn_layer (src/cuda/cuda_lrn_layer.cu:35-101,172-215). */
int cb200_lrn_forward(const cb200_lrn_desc* d, const void* x, void* y, float* local_scale, void* stream);
int cb200_lrn_backward(const cb200_lrn_desc* d, const void* x, const void* y, const void* dy, void* dx,
                       const float* local_scale, const cb200_activ* prev_activ, const void* prev_out, void* stream);

/* ------------------------------------------------------------------ dropout (conv / pool / dense outputs) */
/* Replaces cuda_dropout_apply_{conv,dense,pool} / cuda_dropout_scale_* and the cuRAND mask tensors they read
 * (src/cuda/cuda_conv_layer.cu:131-163,399-420,440-447; cuda_dense_layer.cu:78-112,375-411; cuda_pool_layer.cu:280-312,
 * 472-508).  Same arithmetic: the PRE-activation output is multiplied by a 0/1 mask (kept when a uniform draw >=
 * drop_rate, no rescale while training) or, at inference in AVG_MODEL, by (1 - drop_rate); then the layer's activation
 * runs; the backward pass multiplies the layer's delta by the same mask.  The mask is not stored: it is a function of
 * (seed, stream_id, draw, element position) that both passes evaluate, so the layer that produced `y` must have run
 * with a LINEAR epilogue and `activ` is applied here, in the same pass.  A dense layer is c = nb_neurons, h = w = 1
 * (its bias node is not part of the tensor, so it is "always kept" as upstream requires). */
typedef struct {
	int dtype;
	int batch, length;
	int c, h, w;
	float drop_rate;
	unsigned int stream_id;        /* which layer: decorrelates the layers of one network */
	unsigned long long seed;       /* per network (and per data-parallel rank) */
	unsigned long long draw;       /* which forward pass: the backward pass reuses the value of its forward pass */
	cb200_activ activ;             /* applied after the mask / scale in the forward call */
} cb200_dropout_desc;
/* y (tensor in internal layout [B][h][w][Cp]) is updated in place. scale_only != 0: the AVG_MODEL inference branch. */
int cb200_dropout_forward(const cb200_dropout_desc* d, void* y, int scale_only, void* stream);
int cb200_dropout_backward(const cb200_dropout_desc* d, void* dy, void* stream);

/* ------------------------------------------------------------------ output layer: softmax / losses */
/* All three take the last layer's tensor in internal layout [B][h][w][Cp] and the target batch in the
 * reference's target layout [B][c*h*w] (per sample c-major, src/cuda/cuda_activ_functions.cu:114-194),
 * stored in `dtype` like upstream (src/cuda/cuda_main.cu:398). */
/* in-place softmax over the c*h*w values of each sample (zero for b >= length).
 * Replaces softmax_activation_kernel (cuda_activ_functions.cu:280-381). */
int cb200_softmax(void* y, int dtype, int batch, int length, int c, int h, int w, void* stream);
/* delta = (y - t) * scale (quadratic and cross-entropy share it upstream), zero for b >= length.
 * Replaces quadratic/cross_entropy_deriv_output_error (cuda_activ_functions.cu:114-155,384-425). */
int cb200_output_delta(void* delta, const void* y, const void* target, int dtype,
                       int batch, int length, int c, int h, int w, float scale, void* stream);
/* the same for an output layer with a RELU / LOGISTIC activation: upstream stores (y - t) * scale and then runs the
 * layer's own derivative kernel on it (cuda_ReLU_deriv_output_error / cuda_logistic_deriv_output_error,
 * cuda_activ_functions.cu:2221-2233,2273-2285); one pass here, same intermediate rounding. activ NULL / LINEAR /
 * SOFTMAX = cb200_output_delta. */
int cb200_output_delta_activ(void* delta, const void* y, const void* target, int dtype, int batch, int length,
                             int c, int h, int w, float scale, const cb200_activ* activ, void* stream);
/* per-sample loss summed over the sample's outputs -> loss[batch] (FP32), kind 0: 0.5*(y-t)^2,
 * kind 1: -t*log(max(y,1e-6)).  Replaces *_output_error kernels + the host-side summation of the
 * whole per-element tensor (cuda_activ_functions.cu:157-194,427-470; src/auxil.c:1851-1917). */
int cb200_output_loss(float* loss, const void* y, const void* target, int dtype,
                      int batch, int length, int c, int h, int w, int kind, void* stream);

/* per-ELEMENT loss in upstream's table layout (conv / pool outputs err[c][batch][h*w], dense_layout: err[batch][c + 1]
 * with the bias node untouched), FP32, rows b >= length untouched (the caller zeroes the table as upstream does,
 * src/auxil.c:1853-1856).  For the upstream-side back-end shim: upstream's host code sums this table itself
 * (cuda_output_error_fct, src/cuda/cuda_activ_functions.cu:2585-2602 + src/auxil.c:1871-1913). */
int cb200_output_error_elems(float* err, const void* y, const void* target, int dtype, int batch, int length,
                             int c, int h, int w, int kind, int dense_layout, void* stream);
/* YOLO twin: writes the six loss parts of cb200_yolo_loss ([batch][6]) into upstream's table err[ch][batch][cells], each
 * on the first element of its channel class, so that the per-class sums of src/auxil.c:1429-1455 come out the same. */
int cb200_yolo_scatter_parts(float* err, const float* parts, int batch, int length, int cells, int nb_class, int nb_param, void* stream);

/* per-sample argmax (class-major index over the c*h*w outputs, first maximum wins) of the output and of the target
 * row into two device int32 [batch] arrays (-1 for b >= length): the inputs of the confusion matrix of
 * compute_error (src/auxil.c:1365-1426, 1562-1658) without copying the output tensor to the host. */
int cb200_output_argmax(int* pred, int* truth, const void* y, const void* target, int dtype,
                        int batch, int length, int c, int h, int w, void* stream);

/* ------------------------------------------------------------------ YOLO output layer */
/* Detection head of the reference (src/activ_functions.c:970-1477 for the parameters,
 * src/cuda/cuda_activ_functions.cu:477-597 activation, :700-1406 association + error signal,
 * :1409-2075 loss monitor).  The last conv layer carries nb_box*(8+nb_class+nb_param) filters; per box:
 * [x y z | w h d | prob | objectness | classes.. | params..].  Tensors are in internal layout
 * [B][gh][gw][Cp]; the target batch keeps the reference layout, one row per image,
 * [n_obj, (class, x0, y0, z0, x1, y1, z1, params.., (difficult))...] of target_stride values, stored
 * in `dtype` like upstream.  All tables below are plain values; the three pointers are DEVICE arrays. */
enum { CB200_IOU = 0, CB200_GIOU = 1, CB200_DIOU = 2, CB200_DIOU2 = 3 };           /* src/structs.h:40 */
enum { CB200_DIST_IOU = 0, CB200_DIST_SIZE = 1, CB200_DIST_OFFSET = 2 };            /* src/structs.h:41 */
enum { CB200_ERR_COMPLETE = 0, CB200_ERR_NATURAL = 1 };                             /* src/structs.h:42 */
#define CB200_YOLO_MAX_BOX 32
typedef struct {
	int dtype;
	int batch, length;
	int grid_h, grid_w;                   /* nb_area[1], nb_area[0] of the last conv layer */
	int nb_box, nb_class, nb_param;
	int max_nb_obj;                       /* max_nb_obj_per_image */
	int target_stride;                    /* values per image in the target batch: 1+max_nb_obj*(7+nb_param+diff_flag) */
	int fit_dim;
	int IoU_type, prior_dist_type, error_type;
	int class_softmax, diff_flag;
	int strict_box_size_association;
	int rand_startup;
	float rand_prob_best_box_assoc, rand_prob, min_prior_forced_scaling;
	int cell_size[3];                     /* in_dims[i] / nb_area[i] (third = depth, 1 grid cell) */
	float scale_tab[6];                   /* pos, size, prob, obj, class, param */
	float slopes_and_maxes[18];           /* [6][slope, max, min] */
	float IoU_limits[8];
	int fit_parts[6];
	const float* prior_size;              /* device [nb_box][3], already clamped to >= 1 */
	const float* noobj_prob_prior;        /* device [nb_box] */
	const float* param_ind_scale;         /* device [nb_param] (may be NULL when nb_param == 0) */
} cb200_yolo_desc;
/* FP32 scratch shared by the two association passes: [batch][max_nb_obj][nb_box + 1] */
size_t cb200_yolo_workspace_bytes(const cb200_yolo_desc* d);
/* in place on the linear output of the last convolution.  Replaces YOLO_activation_kernel (:477-597). */
int cb200_yolo_activation(const cb200_yolo_desc* d, void* y, void* stream);
/* target association + error signal delta (already multiplied by the activation derivative and by
 * tc_scale; zero for b >= length and for pad channels).  nb_im_iter = iter * train.size drives the
 * random start-up phase; seed/step select the draw of the counter-based generator used by the random
 * association branches (upstream: curand states seeded with time(NULL)).  box_state (optional, device
 * int32 [B][gh*gw][nb_box]) receives upstream's box_locked values: 0 background, 1 good-but-not-best,
 * 2 associated.  Replaces YOLO_deriv_error_kernel (:700-1406). */
int cb200_yolo_delta(const cb200_yolo_desc* d, void* delta, const void* y, const void* target,
                     float tc_scale, long long nb_im_iter, unsigned long long seed, unsigned long long step,
                     int* box_state, float* workspace, void* stream);
/* loss monitor: loss[b] = sum of the image's per-element YOLO errors; parts (optional) FP32 [batch][6] =
 * the same sum split into position/size/probability/objectness/class/param; monitor (optional) FP32
 * [B][gh*gw][nb_box][2] = (objectness, IoU) of each associated box, -1 elsewhere (upstream IoU_monitor).
 * Replaces YOLO_error_kernel (:1409-2075) + the host-side sums of src/auxil.c:1429-1486. */
int cb200_yolo_loss(const cb200_yolo_desc* d, float* loss, float* parts, float* monitor, const void* y,
                    const void* target, float* workspace, void* stream);
/* decoded boxes for a forward pass written in the reference's [C][B][gh*gw] FP32 layout
 * (x0,y0,z0,x1,y1,z1 in pixels, then prob, objectness, classes, params), src/auxil.c:1304-1344. */
int cb200_yolo_export_boxes(const cb200_yolo_desc* d, float* dst, const void* y, void* stream);

/* ------------------------------------------------------------------ data parallel (NCCL) */
/* One process per GPU.  Rank 0 creates the id (128 bytes) and ships it to the others through any
 * side channel (torch.distributed store, MPI, a file); then every rank calls cb200_dp_init.
 * The reference has no multi-GPU path (src/cuda/cuda_main.cu:1066). */
int cb200_dp_unique_id(void* id128);
int cb200_dp_init(const void* id128, int rank, int world);
int cb200_dp_world(void);
/* in-place sum all-reduce of an FP32 device buffer on the core's communication stream, ordered
 * after everything enqueued so far on `after_stream`; cb200_dp_join makes `stream` wait for all
 * all-reduces issued so far.  No-ops when world == 1. */
int cb200_dp_allreduce(float* buf, size_t n, void* after_stream);
/* the communication stream additionally waits for everything enqueued so far on `stream` (a bucket whose
 * gradients were produced on two streams: cb200_dp_after(a); cb200_dp_allreduce(buf, n, b)) */
int cb200_dp_after(void* stream);
int cb200_dp_join(void* stream);
/* in-place broadcast of `bytes` bytes from rank `root`, enqueued on `stream` (replica initialisation:
 * every rank must start from rank 0's parameters and optimizer state) */
int cb200_dp_broadcast(void* buf, size_t bytes, int root, void* stream);
int cb200_dp_rank(void);
int cb200_dp_finalize(void);

#ifdef __cplusplus
}
#endif
#endif /* CIANNA_B200_H */
