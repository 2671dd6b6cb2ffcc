// conv.cu - C-ABI entry points of the convolution layer: operand preparation, dispatch between the
// tcgen05 implicit-GEMM kernels (conv_tc.cu) and the generic SIMT kernels (conv_simt.cu), and the
// fused optimizer.  Reference: src/cuda/cuda_conv_layer.cu:319-562, src/cuda/cuda_main.cu:433-475.
#include "common.cuh"

namespace cb200 {
extern const char* g_last_conv_impl;
extern int g_force_simt;
int conv_forward_simt(const cb200_conv_desc*, const cb200_conv_weights*, const void*, void*, cudaStream_t);
int conv_dgrad_simt(const cb200_conv_desc*, const cb200_conv_weights*, const void*, void*, const cb200_activ*, const void*, cudaStream_t);
int conv_wgrad_simt(const cb200_conv_desc*, const cb200_conv_weights*, const void*, const void*, cudaStream_t);
int conv_colsum(int dtype, const void* dy, float* out, long long P, int c, cudaStream_t st);
// tcgen05 path (conv_tc.cu); *_supported() says whether the geometry is taken
bool conv_tc_fwd_supported(const cb200_conv_desc*);
bool conv_tc_dgrad_supported(const cb200_conv_desc*);
bool conv_tc_wgrad_supported(const cb200_conv_desc*);
extern int g_gn_epilogue_mode;
int conv_forward_tc(const cb200_conv_desc*, const cb200_conv_weights*, const void*, void*, cudaStream_t,
                    const cb200_norm_desc* gn = nullptr, void* gn_ws = nullptr, int* gn_fused = nullptr);
int conv_dgrad_tc(const cb200_conv_desc*, const cb200_conv_weights*, const void*, void*, const cb200_activ*, const void*, cudaStream_t);
int conv_wgrad_tc(const cb200_conv_desc*, const cb200_conv_weights*, const void*, const void*, cudaStream_t);
// first layer straight from the dataset batch, patch rows built in shared memory (conv_first.cu)
bool conv_first_supported(const cb200_conv_desc*);
int conv_first_forward(const cb200_conv_desc*, const cb200_conv_weights*, const void*, void*, cudaStream_t, const cb200_norm_desc*, void*, int*);
int conv_first_wgrad(const cb200_conv_desc*, const cb200_conv_weights*, const void*, const void*, cudaStream_t);

// master [out_c][taps*in_c + 1] (column = c*taps + tap, bias last) -> compute operands.
// One thread per master element; the same mapping is used by the optimizer below.
// w_bwd row of input channel c, filter tap `tap`: (c, rotated tap) for the data-gradient correlation, or - for a filter
// that covers its whole input map (conv_whole_map, common.cuh: dense layers behind conv / pool) - (tap, c), the order of
// the input tensor itself, so that the layer runs as ONE 1x1 GEMM over f_h*f_w*Cp "channels" in all three passes
__device__ __forceinline__ size_t wbwd_row(int c, int tap, int taps, int in_cp, int wb_dense) {
	return wb_dense ? (size_t)tap * in_cp + c : (size_t)c * taps + (taps - 1 - tap);
}
template <typename T>
__device__ __forceinline__ void scatter_weight(float wv, int f, int col, int taps, int in_c, int in_cp, int out_cp,
                                               T* __restrict__ w_fwd, T* __restrict__ w_bwd, float* __restrict__ bias_w, int wb_dense) {
	if (col == taps * in_c) { bias_w[f] = wv; return; }
	const int c = col / taps, tap = col - c * taps;
	w_fwd[((size_t)f * taps + tap) * in_cp + c] = from_f32<T>(wv);
	w_bwd[wbwd_row(c, tap, taps, in_cp, wb_dense) * out_cp + f] = from_f32<T>(wv);
}

template <typename T>
__global__ void conv_prepare_kernel(const float* __restrict__ master, T* __restrict__ w_fwd, T* __restrict__ w_bwd,
                                    float* __restrict__ bias_w, int out_c, int taps, int in_c, int in_cp, int out_cp,
                                    size_t ms_f, size_t ms_c, int wb_dense) {
	// master element (filter f, column col) lives at f*ms_f + col*ms_c:
	//   conv  layout [out_c][kref]      -> ms_f = kref, ms_c = 1
	//   dense layout [in_size][n + 1]   -> ms_f = 1,    ms_c = n + 1   (src/dense_layer.c:253-268)
	const int kref = taps * in_c + 1;
	const size_t total = (size_t)out_c * kref;
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
		const int f = (int)(i / kref), col = (int)(i - (size_t)f * kref);
		scatter_weight<T>(master[f * ms_f + col * ms_c], f, col, taps, in_c, in_cp, out_cp, w_fwd, w_bwd, bias_w, wb_dense);
	}
}

// moment = (lr/B)*g + mom*moment ; moment += lr*wd*w*S ; w -= moment/S   (cuda_conv_layer.cu:551-560,
// cuda_main.cu:455-467), then the refreshed 16-bit operands are written in the same pass.
template <typename T>
__global__ void conv_update_kernel(float* __restrict__ master, float* __restrict__ moment,
                                   const float* __restrict__ grad, const float* __restrict__ grad_b,
                                   const float* __restrict__ hyper, float bias_value, int is_pivot,
                                   T* __restrict__ w_fwd, T* __restrict__ w_bwd, float* __restrict__ bias_w,
                                   int out_c, int taps, int in_c, int in_cp, int out_cp, size_t ms_f, size_t ms_c, int wb_dense) {
	const int kref = taps * in_c + 1;
	const size_t total = (size_t)out_c * kref;
	const float alpha = hyper[0], mom = hyper[1], wdlr = hyper[2], S = hyper[3];
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
		const int f = (int)(i / kref), col = (int)(i - (size_t)f * kref);
		const size_t mi = f * ms_f + col * ms_c;
		float wv = master[mi];
		if (i < total - (size_t)is_pivot) {
			float g;
			if (col == kref - 1) g = bias_value * grad_b[f];
			else { const int c = col / taps, tap = col - c * taps; g = grad[((size_t)f * taps + tap) * in_cp + c]; }
			float m = alpha * g + mom * moment[mi];
			m += wdlr * wv * S;
			wv -= m / S;
			moment[mi] = m;
			master[mi] = wv;
		}
		scatter_weight<T>(wv, f, col, taps, in_c, in_cp, out_cp, w_fwd, w_bwd, bias_w, wb_dense);
	}
}

// Dense layers (master [in_size][n + 1]: one ROW per input, neurons contiguous): the same update on 32 x 32 tiles.  The raw
// gradient and w_fwd are [n][K] (K = taps * in_cp contiguous), master / momentum / w_bwd are [K][n]-oriented, so the tile
// goes through shared memory once each way and every global access is a full 64 - 128 B row segment - the generic kernel
// above reads 4 B out of every 32 B sector of master and momentum (measured on the 12289 x 3072 layer of the
// extinction-profile network: 3.4 ms, 27 GB of L2 traffic for 1 GB of data).
template <typename T>
__global__ void __launch_bounds__(256)
dense_update_tiled_kernel(float* __restrict__ master, float* __restrict__ moment, const float* __restrict__ grad,
                          const float* __restrict__ grad_b, const float* __restrict__ hyper, float bias_value,
                          T* __restrict__ w_fwd, T* __restrict__ w_bwd, float* __restrict__ bias_w,
                          int n, int taps, int in_c, int in_cp, int out_cp, int wb_dense) {
	__shared__ float gt[32][33];      // gradient tile [neuron][k]
	__shared__ float wt[32][33];      // updated weights [k][neuron]
	const int k0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
	const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
	const int K = taps * in_cp;
	const size_t ms_c = (size_t)n + 1;
	const float alpha = hyper[0], mom = hyper[1], wdlr = hyper[2], S = hyper[3];
#pragma unroll
	for (int r = ty; r < 32; r += 8) {
		const int f = f0 + r, k = k0 + tx;
		gt[r][tx] = (f < n && k < K) ? grad[(size_t)f * K + k] : 0.0f;
	}
	__syncthreads();
#pragma unroll
	for (int r = ty; r < 32; r += 8) {
		const int k = k0 + r, f = f0 + tx;
		float wv = 0.0f;
		if (k < K && f < n) {
			const int tap = k / in_cp, c = k - tap * in_cp;
			if (c < in_c) {
				const size_t mi = ((size_t)c * taps + tap) * ms_c + f;      // master row of (channel c, tap): upstream's flatten order
				wv = master[mi];
				float m = alpha * gt[tx][r] + mom * moment[mi];
				m += wdlr * wv * S;
				wv -= m / S;
				moment[mi] = m;
				master[mi] = wv;
				w_bwd[wbwd_row(c, tap, taps, in_cp, wb_dense) * out_cp + f] = from_f32<T>(wv);
			}
		}
		wt[r][tx] = wv;
	}
	__syncthreads();
#pragma unroll
	for (int r = ty; r < 32; r += 8) {
		const int f = f0 + r, k = k0 + tx;
		if (f < n && k < K && (k % in_cp) < in_c) w_fwd[(size_t)f * K + k] = from_f32<T>(wt[tx][r]);
	}
	if (blockIdx.x == 0 && ty == 0) {         // the bias row (input in_size - 1)
		const int f = f0 + tx;
		if (f < n) {
			const size_t mi = (size_t)taps * in_c * ms_c + f;
			float wv = master[mi];
			float m = alpha * (bias_value * grad_b[f]) + mom * moment[mi];
			m += wdlr * wv * S;
			wv -= m / S;
			moment[mi] = m;
			master[mi] = wv;
			bias_w[f] = wv;
		}
	}
}

// Conv layers (master in the conv layout, ms_c == 1): the same update with threads in OPERAND order (f, tap, c) instead
// of master order, so the raw gradient is read and w_fwd written fully coalesced; the master / momentum accesses of a
// warp then stride by `taps` floats but stay inside one filter's contiguous span (L1 resident).  The transposed +
// rotated copy for the data gradient is produced by a tiled transpose of w_fwd (both sides coalesced) instead of one
// scattered 2-byte store per weight.
template <typename T>
__global__ void __launch_bounds__(256)
conv_update_operand_kernel(float* __restrict__ master, float* __restrict__ moment, const float* __restrict__ grad,
                           const float* __restrict__ grad_b, const float* __restrict__ hyper, float bias_value,
                           T* __restrict__ w_fwd, float* __restrict__ bias_w, int out_c, int taps, int in_c, int in_cp) {
	const int kref = taps * in_c + 1;
	const size_t total = (size_t)out_c * taps * in_cp;
	const float alpha = hyper[0], mom = hyper[1], wdlr = hyper[2], S = hyper[3];
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
		const int c = (int)(i % in_cp);
		const size_t r = i / in_cp;
		const int tap = (int)(r % taps), f = (int)(r / taps);
		if (c < in_c) {
			const size_t mi = (size_t)f * kref + (size_t)c * taps + tap;
			float wv = master[mi];
			float m = alpha * grad[i] + mom * moment[mi];
			m += wdlr * wv * S;
			wv -= m / S;
			moment[mi] = m;
			master[mi] = wv;
			w_fwd[i] = from_f32<T>(wv);
		}
		if (c == 0 && tap == 0) {      // this filter's bias column
			const size_t mi = (size_t)f * kref + kref - 1;
			float wv = master[mi];
			float m = alpha * (bias_value * grad_b[f]) + mom * moment[mi];
			m += wdlr * wv * S;
			wv -= m / S;
			moment[mi] = m;
			master[mi] = wv;
			bias_w[f] = wv;
		}
	}
}

// The same update with one block per filter: the filter's master and momentum rows (k_ref contiguous floats each) are
// staged in shared memory with coalesced loads, updated there in operand order (tap, c) - coalesced gradient reads and
// w_fwd writes, conflict-free strided shared-memory accesses (stride `taps` is odd for 1x1 / 3x3 / 5x5) - and written
// back coalesced.  22 bytes per weight, all of them in full lines (the one-thread-per-weight forms above cost a
// 32-byte sector per 2- or 4-byte access on one side or the other).
template <typename T>
__global__ void __launch_bounds__(256)
conv_update_rows_kernel(float* __restrict__ master, float* __restrict__ moment, const float* __restrict__ grad,
                        const float* __restrict__ grad_b, const float* __restrict__ hyper, float bias_value,
                        T* __restrict__ w_fwd, float* __restrict__ bias_w, int taps, int in_c, int in_cp) {
	extern __shared__ float rows[];                 // [2][kref]: master row, momentum row
	const int kref = taps * in_c + 1;
	const int f = blockIdx.x;
	float* sm_w = rows;
	float* sm_m = rows + kref;
	const float alpha = hyper[0], mom = hyper[1], wdlr = hyper[2], S = hyper[3];
	for (int i = threadIdx.x; i < kref; i += blockDim.x) {
		sm_w[i] = master[(size_t)f * kref + i];
		sm_m[i] = moment[(size_t)f * kref + i];
	}
	__syncthreads();
	const int n_op = taps * in_cp;
	const float* __restrict__ g = grad + (size_t)f * n_op;
	T* __restrict__ wf = w_fwd + (size_t)f * n_op;
	for (int j = threadIdx.x; j < n_op; j += blockDim.x) {
		const int tap = j / in_cp, c = j - tap * in_cp;
		if (c >= in_c) continue;
		const int mi = c * taps + tap;
		float wv = sm_w[mi], m = sm_m[mi];
		sgd_momentum_step(alpha, mom, wdlr, S, g[j], m, wv);
		sm_m[mi] = m;
		sm_w[mi] = wv;
		wf[j] = from_f32<T>(wv);
	}
	if (threadIdx.x == 0) {
		const int mi = kref - 1;
		float wv = sm_w[mi], m = sm_m[mi];
		sgd_momentum_step(alpha, mom, wdlr, S, __fmul_rn(bias_value, grad_b[f]), m, wv);
		sm_m[mi] = m;
		sm_w[mi] = wv;
		bias_w[f] = wv;
	}
	__syncthreads();
	for (int i = threadIdx.x; i < kref; i += blockDim.x) {
		master[(size_t)f * kref + i] = sm_w[i];
		moment[(size_t)f * kref + i] = sm_m[i];
	}
}

// w_bwd[c][taps-1-tap][f] = w_fwd[f][tap][c] through a 32x32 shared-memory tile; grid (c tiles, f tiles, taps)
template <typename T>
__global__ void __launch_bounds__(256)
conv_wbwd_transpose_kernel(const T* __restrict__ w_fwd, T* __restrict__ w_bwd, int out_c, int out_cp, int taps, int in_c, int in_cp, int wb_dense) {
	__shared__ T tile[32][33];
	const int tap = blockIdx.z, c0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
	const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
	for (int k = 0; k < 4; k++) {
		const int f = f0 + ty + 8 * k, c = c0 + tx;
		tile[ty + 8 * k][tx] = (f < out_c && c < in_cp) ? w_fwd[((size_t)f * taps + tap) * in_cp + c] : from_f32<T>(0.0f);
	}
	__syncthreads();
#pragma unroll
	for (int k = 0; k < 4; k++) {
		const int c = c0 + ty + 8 * k, f = f0 + tx;
		if (c < in_c && f < out_cp) w_bwd[wbwd_row(c, tap, taps, in_cp, wb_dense) * out_cp + f] = tile[tx][ty + 8 * k];
	}
}

// ---- first layer run on patch rows (cb200_import_input_patches): a 1x1 GEMM over kp columns, bias inside the GEMM
static cb200_conv_desc effective_desc(const cb200_conv_desc* d) {
	cb200_conv_desc e = *d;
	if (d->input_is_patches) {
		e.in_c = cb200_patch_width(d->in_c, d->f_h, d->f_w);
		e.in_h = d->out_h; e.in_w = d->out_w;
		e.f_h = 1; e.f_w = 1; e.stride_h = 1; e.stride_w = 1; e.pad_h = 0; e.pad_w = 0;
		e.bias_value = 0.0f;
		e.input_is_patches = 0;
	}
	return e;
}
template <typename T>
__global__ void patch_prepare_kernel(const float* __restrict__ master, T* __restrict__ w_fwd, float* __restrict__ bias_w, int out_c, int kref, int kp) {
	const size_t total = (size_t)out_c * kp;
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
		const int f = (int)(i / kp), col = (int)(i - (size_t)f * kp);
		w_fwd[i] = from_f32<T>(col < kref ? master[(size_t)f * kref + col] : 0.0f);
		if (col == 0) bias_w[f] = 0.0f;
	}
}
template <typename T>
__global__ void patch_update_kernel(float* __restrict__ master, float* __restrict__ moment, const float* __restrict__ grad,
                                    const float* __restrict__ hyper, T* __restrict__ w_fwd, int out_c, int kref, int kp) {
	const size_t total = (size_t)out_c * kref;
	const float alpha = hyper[0], mom = hyper[1], wdlr = hyper[2], S = hyper[3];
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
		const int f = (int)(i / kref), col = (int)(i - (size_t)f * kref);
		float wv = master[i];
		float m = alpha * grad[(size_t)f * kp + col] + mom * moment[i];   // the bias column's gradient comes out of the GEMM itself
		m += wdlr * wv * S;
		wv -= m / S;
		moment[i] = m;
		master[i] = wv;
		w_fwd[(size_t)f * kp + col] = from_f32<T>(wv);
	}
}

static int check_desc(const cb200_conv_desc* d) {
	CB_ARG(d != nullptr);
	CB_ARG(d->batch > 0 && d->in_c > 0 && d->out_c > 0 && d->in_h > 0 && d->in_w > 0 && d->out_h > 0 && d->out_w > 0);
	CB_ARG(d->f_h > 0 && d->f_w > 0 && d->stride_h > 0 && d->stride_w > 0 && d->pad_h >= 0 && d->pad_w >= 0);
	CB_ARG(d->ipad_w >= 0 && d->ipad_h >= 0 && d->ipad_d >= 0 && d->pad_d >= 0);
	// output size of the reference (nb_area_comp, src/conv_layer.c:32-41): the filter slides over the zero-stuffed, padded input
	CB_ARG(d->out_h == (d->in_h + (d->in_h - 1) * d->ipad_h + 2 * d->pad_h - d->f_h) / d->stride_h + 1);
	CB_ARG(d->out_w == (d->in_w + (d->in_w - 1) * d->ipad_w + 2 * d->pad_w - d->f_w) / d->stride_w + 1);
	CB_ARG(conv_out_d(d) == (conv_in_d(d) + (conv_in_d(d) - 1) * d->ipad_d + 2 * d->pad_d - conv_f_d(d)) / (d->stride_d > 0 ? d->stride_d : 1) + 1);
	CB_ARG(!(conv_generic(d) && d->input_is_patches));
	CB_ARG(d->length >= 0 && d->length <= d->batch);
	return CB200_OK;
}
}  // namespace cb200
using namespace cb200;

extern "C" {

int cb200_conv_first_direct(const cb200_conv_desc* d) {
	return d != nullptr && d->input_is_patches != 0 && !g_force_simt && !conv_generic(d) && conv_first_supported(d) ? 1 : 0;
}

size_t cb200_conv_wfwd_elems(const cb200_conv_desc* d) {
	if (d->input_is_patches) return (size_t)d->out_c * cb200_patch_width(d->in_c, d->f_h, d->f_w);
	return (size_t)d->out_c * conv_taps(d) * round8(d->in_c);
}
size_t cb200_conv_wbwd_elems(const cb200_conv_desc* d) {
	// (whole-map filters keep one row per element of the padded input tensor, see wbwd_row)
	return (size_t)(conv_whole_map(d) ? round8(d->in_c) : d->in_c) * conv_taps(d) * round8(d->out_c);
}
size_t cb200_conv_grad_elems(const cb200_conv_desc* d) { return cb200_conv_wfwd_elems(d); }
size_t cb200_conv_master_elems(const cb200_conv_desc* d) { return (size_t)d->out_c * ((size_t)conv_taps(d) * d->in_c + 1); }

static int prepare_weights_impl(const cb200_conv_desc* d, const cb200_conv_weights* w, size_t ms_f, size_t ms_c, void* s) {
	CB_REQUIRE_DEVICE();
	int rc = check_desc(d); if (rc) return rc;
	cudaStream_t st = as_stream(s);
	if (d->input_is_patches) {
		const int kref = d->f_h * d->f_w * d->in_c + 1, kp = cb200_patch_width(d->in_c, d->f_h, d->f_w);
		CB_DISPATCH_DTYPE(d->dtype, T, (patch_prepare_kernel<T><<<grid_for((long long)d->out_c * kp, 256), 256, 0, st>>>(
			w->master, (T*)w->w_fwd, w->bias_w, d->out_c, kref, kp)));
		CB_LAUNCH_CHECK();
		return CB200_OK;
	}
	// pad lanes of the operands must be zero: clear, then scatter
	CB_CUDA(cudaMemsetAsync(w->w_fwd, 0, cb200_conv_wfwd_elems(d) * cb200_dtype_size(d->dtype), st));
	CB_CUDA(cudaMemsetAsync(w->w_bwd, 0, cb200_conv_wbwd_elems(d) * cb200_dtype_size(d->dtype), st));
	const int taps = conv_taps(d);
	long long total = (long long)cb200_conv_master_elems(d);
	CB_DISPATCH_DTYPE(d->dtype, T, (conv_prepare_kernel<T><<<grid_for(total, 256), 256, 0, st>>>(
		w->master, (T*)w->w_fwd, (T*)w->w_bwd, w->bias_w, d->out_c, taps, d->in_c, round8(d->in_c), round8(d->out_c), ms_f, ms_c,
		conv_whole_map(d) ? 1 : 0)));
	CB_LAUNCH_CHECK();
	return CB200_OK;
}
int cb200_conv_prepare_weights(const cb200_conv_desc* d, const cb200_conv_weights* w, void* s) {
	return prepare_weights_impl(d, w, (size_t)conv_taps(d) * d->in_c + 1, 1, s);
}
int cb200_dense_prepare_weights(const cb200_conv_desc* d, const cb200_conv_weights* w, void* s) {
	return prepare_weights_impl(d, w, 1, (size_t)d->out_c + 1, s);
}

void cb200_set_gn_epilogue_stats(int mode) { g_gn_epilogue_mode = mode < 0 ? -1 : (mode > 2 ? 2 : mode); }

int cb200_conv_forward(const cb200_conv_desc* d_in, const cb200_conv_weights* w, const void* x, void* y, void* s) {
	return cb200_conv_forward_stats(d_in, w, x, y, nullptr, nullptr, nullptr, s);
}

int cb200_conv_forward_stats(const cb200_conv_desc* d_in, const cb200_conv_weights* w, const void* x, void* y,
                             const cb200_norm_desc* gn, void* gn_workspace, int* stats_done, void* s) {
	CB_REQUIRE_DEVICE();
	if (stats_done) *stats_done = 0;
	if (stats_done == nullptr) { gn = nullptr; gn_workspace = nullptr; }
	int rc = check_desc(d_in); if (rc) return rc;
	const cb200_conv_desc eff = effective_desc(d_in);
	const cb200_conv_desc* d = &eff;
	// algorithmic FLOPs: 2*M*N*K with K including the bias column, excluding any channel padding
	const double flops = 2.0 * d_in->batch * conv_out_d(d_in) * d_in->out_h * d_in->out_w * (double)d_in->out_c * ((double)conv_taps(d_in) * d_in->in_c + 1);
	if (d_in->input_is_patches == 2) {
		if (!conv_first_supported(d_in)) { set_error("cb200_conv_forward: input_is_patches = 2 is not available for this layer (cb200_conv_first_direct)"); return CB200_ERR_UNSUPPORTED; }
		g_last_conv_impl = "tcgen05";
		prof_begin(PROF_CONV_FWD_TC, flops, as_stream(s));
		rc = conv_first_forward(d_in, w, x, y, as_stream(s), gn, gn_workspace, stats_done);
		prof_end(as_stream(s));
		return rc;
	}
	const bool tc = !g_force_simt && !conv_generic(d) && conv_tc_fwd_supported(d);
	g_last_conv_impl = tc ? "tcgen05" : "simt";
	prof_begin(tc ? PROF_CONV_FWD_TC : PROF_CONV_FWD_SIMT, flops, as_stream(s));
	rc = tc ? conv_forward_tc(d, w, x, y, as_stream(s), gn, gn_workspace, stats_done) : conv_forward_simt(d, w, x, y, as_stream(s));
	prof_end(as_stream(s));
	return rc;
}

int cb200_conv_backward_data(const cb200_conv_desc* d, const cb200_conv_weights* w, const void* dy, void* dx,
                             const cb200_activ* prev_activ, const void* prev_out, void* s) {
	CB_REQUIRE_DEVICE();
	int rc = check_desc(d); if (rc) return rc;
	if (d->input_is_patches) { set_error("cb200_conv_backward_data: a patch-input (first) layer has no data gradient"); return CB200_ERR_UNSUPPORTED; }
	const double flops = 2.0 * d->batch * conv_in_d(d) * d->in_h * d->in_w * (double)d->in_c * ((double)conv_taps(d) * d->out_c);
	const bool tc = !g_force_simt && !conv_generic(d) && conv_tc_dgrad_supported(d);
	g_last_conv_impl = tc ? "tcgen05" : "simt";
	prof_begin(tc ? PROF_CONV_DGRAD_TC : PROF_CONV_DGRAD_SIMT, flops, as_stream(s));
	rc = tc ? conv_dgrad_tc(d, w, dy, dx, prev_activ, prev_out, as_stream(s)) : conv_dgrad_simt(d, w, dy, dx, prev_activ, prev_out, as_stream(s));
	prof_end(as_stream(s));
	return rc;
}

int cb200_conv_backward_weights(const cb200_conv_desc* d_in, const cb200_conv_weights* w, const void* x, const void* dy, void* s) {
	return cb200_conv_backward_weights_ex(d_in, w, x, dy, 0, s);
}

int cb200_conv_backward_weights_ex(const cb200_conv_desc* d_in, const cb200_conv_weights* w, const void* x, const void* dy,
                                   int have_grad_b, void* s) {
	CB_REQUIRE_DEVICE();
	int rc = check_desc(d_in); if (rc) return rc;
	const cb200_conv_desc eff = effective_desc(d_in);
	const cb200_conv_desc* d = &eff;
	cudaStream_t st = as_stream(s);
	long long P = (long long)d->batch * conv_out_d(d) * d->out_h * d->out_w;
	if (!d_in->input_is_patches && !have_grad_b) {      // (patch rows carry the bias input as a column: its gradient comes out of the GEMM)
		rc = conv_colsum(d->dtype, dy, w->grad_b, P, d->out_c, st);
		if (rc) return rc;
	}
	const double flops = 2.0 * P * (double)d_in->out_c * ((double)conv_taps(d_in) * d_in->in_c + 1);
	if (d_in->input_is_patches == 2) {
		if (!conv_first_supported(d_in)) { set_error("cb200_conv_backward_weights: input_is_patches = 2 is not available for this layer"); return CB200_ERR_UNSUPPORTED; }
		g_last_conv_impl = "tcgen05";
		prof_begin(PROF_CONV_WGRAD_TC, flops, st);
		rc = conv_first_wgrad(d_in, w, x, dy, st);
		prof_end(st);
		return rc;
	}
	const bool tc = !g_force_simt && !conv_generic(d) && conv_tc_wgrad_supported(d);
	g_last_conv_impl = tc ? "tcgen05" : "simt";
	prof_begin(tc ? PROF_CONV_WGRAD_TC : PROF_CONV_WGRAD_SIMT, flops, st);
	rc = tc ? conv_wgrad_tc(d, w, x, dy, st) : conv_wgrad_simt(d, w, x, dy, st);
	prof_end(st);
	return rc;
}

static int update_impl(const cb200_conv_desc* d, const cb200_conv_weights* w, const float* hyper, int is_pivot,
                       size_t ms_f, size_t ms_c, void* s) {
	CB_REQUIRE_DEVICE();
	int rc = check_desc(d); if (rc) return rc;
	if (d->input_is_patches) {
		const int kref = d->f_h * d->f_w * d->in_c + 1, kp = cb200_patch_width(d->in_c, d->f_h, d->f_w);
		CB_DISPATCH_DTYPE(d->dtype, T, (patch_update_kernel<T><<<grid_for((long long)d->out_c * kref, 256), 256, 0, as_stream(s)>>>(
			w->master, w->moment, w->grad, hyper, (T*)w->w_fwd, d->out_c, kref, kp)));
		CB_LAUNCH_CHECK();
		return CB200_OK;
	}
	const int taps = conv_taps(d);
	long long total = (long long)cb200_conv_master_elems(d);
	static const bool old_update = getenv("CB200_OLD_UPDATE") != nullptr;
	if (!old_update && ms_c == 1 && !is_pivot && ms_f == (size_t)taps * d->in_c + 1) {
		const int in_cp = round8(d->in_c), out_cp = round8(d->out_c);
		const long long op_total = (long long)d->out_c * taps * in_cp;
		const size_t row_smem = 2 * ((size_t)taps * d->in_c + 1) * sizeof(float);
		if (row_smem <= 96 * 1024) {
			static bool configured = false;
			if (!configured) {
				CB_CUDA(cudaFuncSetAttribute(conv_update_rows_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
				CB_CUDA(cudaFuncSetAttribute(conv_update_rows_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
				CB_CUDA(cudaFuncSetAttribute(conv_update_rows_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
				configured = true;
			}
			CB_DISPATCH_DTYPE(d->dtype, T, (conv_update_rows_kernel<T><<<d->out_c, 256, row_smem, as_stream(s)>>>(
				w->master, w->moment, w->grad, w->grad_b, hyper, d->bias_value, (T*)w->w_fwd, w->bias_w, taps, d->in_c, in_cp)));
		} else {
			CB_DISPATCH_DTYPE(d->dtype, T, (conv_update_operand_kernel<T><<<grid_for(op_total, 256), 256, 0, as_stream(s)>>>(
				w->master, w->moment, w->grad, w->grad_b, hyper, d->bias_value, (T*)w->w_fwd, w->bias_w, d->out_c, taps, d->in_c, in_cp)));
		}
		CB_LAUNCH_CHECK();
		dim3 tgrid((unsigned)ceil_div(d->in_c, 32), (unsigned)ceil_div(out_cp, 32), (unsigned)taps);
		CB_DISPATCH_DTYPE(d->dtype, T, (conv_wbwd_transpose_kernel<T><<<tgrid, 256, 0, as_stream(s)>>>(
			(const T*)w->w_fwd, (T*)w->w_bwd, d->out_c, out_cp, taps, d->in_c, in_cp, conv_whole_map(d) ? 1 : 0)));
		CB_LAUNCH_CHECK();
		return CB200_OK;
	}
	if (!old_update && ms_f == 1 && ms_c == (size_t)d->out_c + 1 && !is_pivot) {
		const int in_cp = round8(d->in_c), K = taps * in_cp;
		const dim3 tgrid((unsigned)ceil_div(K, 32), (unsigned)ceil_div(d->out_c, 32));
		CB_DISPATCH_DTYPE(d->dtype, T, (dense_update_tiled_kernel<T><<<tgrid, 256, 0, as_stream(s)>>>(
			w->master, w->moment, w->grad, w->grad_b, hyper, d->bias_value, (T*)w->w_fwd, (T*)w->w_bwd, w->bias_w,
			d->out_c, taps, d->in_c, in_cp, round8(d->out_c), conv_whole_map(d) ? 1 : 0)));
		CB_LAUNCH_CHECK();
		return CB200_OK;
	}
	CB_DISPATCH_DTYPE(d->dtype, T, (conv_update_kernel<T><<<grid_for(total, 256), 256, 0, as_stream(s)>>>(
		w->master, w->moment, w->grad, w->grad_b, hyper, d->bias_value, is_pivot,
		(T*)w->w_fwd, (T*)w->w_bwd, w->bias_w, d->out_c, taps, d->in_c, round8(d->in_c), round8(d->out_c), ms_f, ms_c,
		conv_whole_map(d) ? 1 : 0)));
	CB_LAUNCH_CHECK();
	return CB200_OK;
}
int cb200_conv_update(const cb200_conv_desc* d, const cb200_conv_weights* w, const float* hyper, int is_pivot, void* s) {
	return update_impl(d, w, hyper, is_pivot, (size_t)conv_taps(d) * d->in_c + 1, 1, s);
}
int cb200_dense_update(const cb200_conv_desc* d, const cb200_conv_weights* w, const float* hyper, void* s) {
	return update_impl(d, w, hyper, 0, 1, (size_t)d->out_c + 1, s);
}

}  // extern "C"
