// output.cu - output-layer passes: softmax, error signal (delta) and loss monitor.
// Reference kernels: softmax_activation_kernel (src/cuda/cuda_activ_functions.cu:280-381),
// quadratic_* / cross_entropy_* (:114-194, :384-470) and their CPU twins in src/activ_functions.c.
// The target batch keeps the reference layout [B][c*h*w] (per sample: channel-major, then pixel).
#include "common.cuh"

namespace cb200 {

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
	v = is_max ? warp_max(v) : warp_sum(v);
	if (lane == 0) red[wid] = v;
	__syncthreads();
	float r = threadIdx.x < nw ? red[threadIdx.x] : (is_max ? -INFINITY : 0.0f);
	if (wid == 0) {
		r = is_max ? warp_max(r) : warp_sum(r);
		if (lane == 0) red[0] = r;
	}
	__syncthreads();
	r = red[0];
	__syncthreads();
	return r;
}

// one block per sample; softmax over all c*hw real values of the sample (upstream semantics for the
// conv layout: a single softmax per sample across filters AND positions)
template <typename T>
__global__ void softmax_kernel(T* __restrict__ y, int length, int c, int cp, int hw) {
	__shared__ float red[32];
	const int b = blockIdx.x;
	T* row = y + (size_t)b * hw * cp;
	const int n = hw * cp;
	if (b >= length) {
		for (int i = threadIdx.x; i < n; i += blockDim.x) row[i] = from_f32<T>(0.0f);
		return;
	}
	float vmax = -INFINITY;
	for (int i = threadIdx.x; i < n; i += blockDim.x)
		if (i % cp < c) vmax = fmaxf(vmax, to_f32<T>(row[i]));
	vmax = block_reduce(vmax, red, true);
	float sum = 0.0f;
	for (int i = threadIdx.x; i < n; i += blockDim.x)
		if (i % cp < c) sum += expf(to_f32<T>(row[i]) - vmax);
	sum = block_reduce(sum, red, false);
	for (int i = threadIdx.x; i < n; i += blockDim.x)
		row[i] = (i % cp < c) ? from_f32<T>(expf(to_f32<T>(row[i]) - vmax) / sum) : from_f32<T>(0.0f);
}

template <typename T>
__global__ void output_delta_kernel(T* __restrict__ delta, const T* __restrict__ y, const T* __restrict__ target,
                                    int batch, int length, int c, int cp, int hw, float scale, cb200_activ act) {
	const size_t total = (size_t)batch * hw * cp;
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
		const int ch = (int)(i % cp);
		const size_t pix = i / cp;
		const int p = (int)(pix % hw);
		const int b = (int)(pix / hw);
		float v = 0.0f;
		if (ch < c && b < length) {
			const float o = to_f32<T>(y[i]);
			v = (o - to_f32<T>(target[((size_t)b * c + ch) * hw + p])) * scale;
			// RELU / LOGI output layers: upstream stores (o - t) * S, then runs the layer's own derivative kernel on it
			if (act.type != CB200_LINEAR) v = activ_deriv_mul(act, to_f32<T>(from_f32<T>(v)), o);
		}
		delta[i] = from_f32<T>(v);
	}
}

// loss[b] = sum over the sample's outputs; one block per sample
template <typename T>
__global__ void output_loss_kernel(float* __restrict__ loss, const T* __restrict__ y, const T* __restrict__ target,
                                   int length, int c, int cp, int hw, int kind) {
	__shared__ float red[32];
	const int b = blockIdx.x;
	float s = 0.0f;
	if (b < length) {
		const int n = hw * cp;
		for (int i = threadIdx.x; i < n; i += blockDim.x) {
			const int ch = i % cp, p = i / cp;
			if (ch >= c) continue;
			const float o = to_f32<T>(y[(size_t)b * n + i]);
			const float t = to_f32<T>(target[((size_t)b * c + ch) * hw + p]);
			if (kind == 0) s += 0.5f * (o - t) * (o - t);
			else s += -t * logf(o > 0.000001f ? o : 0.000001f);
		}
	}
	s = block_reduce(s, red, false);
	if (threadIdx.x == 0) loss[b] = s;
}

// per-ELEMENT loss in upstream's own table layout (what *_output_error kernels fill: conv / pool outputs
// err[ch][batch][hw], dense outputs err[b][c + 1] with the bias node left at zero) - only the upstream-side back-end shim
// needs it (upstream's host code sums the table itself, src/auxil.c:1871-1913); the host library reduces on the device.
template <typename T>
__global__ void output_error_elems_kernel(float* __restrict__ err, const T* __restrict__ y, const T* __restrict__ target,
                                          int batch, int length, int c, int cp, int hw, int kind, int dense) {
	const size_t total = (size_t)length * hw * c;
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
		const int ch = (int)(i % c);
		const size_t pix = i / c;
		const int p = (int)(pix % hw);
		const int b = (int)(pix / hw);
		const float o = to_f32<T>(y[((size_t)b * hw + p) * cp + ch]);
		const float t = to_f32<T>(target[((size_t)b * c + ch) * hw + p]);
		const float e = kind == 0 ? 0.5f * (o - t) * (o - t) : -t * logf(o > 0.000001f ? o : 0.000001f);
		if (dense) err[(size_t)b * (c + 1) + ch] = e;
		else err[((size_t)ch * batch + b) * hw + p] = e;
	}
}

// YOLO: the loss monitor of this library reduces to one value + six parts per image; upstream's host code re-derives
// the six parts by summing the per-element table by channel class (src/auxil.c:1429-1455).  Each part is written to the
// first element of its class (box 0, cell 0: position x / size w / probability / objectness / first class / first
// parameter channel) of upstream's table err[ch][batch][cells]; every sum upstream forms is preserved.
__global__ void yolo_scatter_parts_kernel(float* __restrict__ err, const float* __restrict__ parts, int batch, int length,
                                          int cells, int nb_class, int nb_param) {
	const int b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= length) return;
	const int ch_of_part[6] = {0, 3, 6, 7, 8, 8 + nb_class};
	for (int k = 0; k < 6; k++) {
		if ((k == 4 && nb_class <= 0) || (k == 5 && nb_param <= 0)) continue;
		err[((size_t)ch_of_part[k] * batch + b) * cells] = parts[(size_t)b * 6 + k];
	}
}
}  // namespace cb200
using namespace cb200;

extern "C" {

int cb200_softmax(void* y, int dtype, int batch, int length, int c, int h, int w, void* s) {
	CB_REQUIRE_DEVICE();
	CB_ARG(batch > 0 && c > 0);
	CB_DISPATCH_DTYPE(dtype, T, (softmax_kernel<T><<<batch, 256, 0, as_stream(s)>>>((T*)y, length, c, round8(c), h * w)));
	CB_LAUNCH_CHECK();
	return CB200_OK;
}

int cb200_output_delta_activ(void* delta, const void* y, const void* target, int dtype, int batch, int length,
                             int c, int h, int w, float scale, const cb200_activ* activ, void* s) {
	CB_REQUIRE_DEVICE();
	long long total = (long long)batch * h * w * round8(c);
	cb200_activ act; act.type = CB200_LINEAR; act.leak = 0; act.saturation = 0; act.beta = 0;
	if (activ != nullptr && (activ->type == CB200_RELU || activ->type == CB200_LOGISTIC)) act = *activ;
	CB_DISPATCH_DTYPE(dtype, T, (output_delta_kernel<T><<<grid_for(total, 256), 256, 0, as_stream(s)>>>(
		(T*)delta, (const T*)y, (const T*)target, batch, length, c, round8(c), h * w, scale, act)));
	CB_LAUNCH_CHECK();
	return CB200_OK;
}
int cb200_output_delta(void* delta, const void* y, const void* target, int dtype, int batch, int length,
                       int c, int h, int w, float scale, void* s) {
	return cb200_output_delta_activ(delta, y, target, dtype, batch, length, c, h, w, scale, nullptr, s);
}

int cb200_output_loss(float* loss, const void* y, const void* target, int dtype, int batch, int length,
                      int c, int h, int w, int kind, void* s) {
	CB_REQUIRE_DEVICE();
	CB_DISPATCH_DTYPE(dtype, T, (output_loss_kernel<T><<<batch, 256, 0, as_stream(s)>>>(
		loss, (const T*)y, (const T*)target, length, c, round8(c), h * w, kind)));
	CB_LAUNCH_CHECK();
	return CB200_OK;
}

int cb200_output_error_elems(float* err, const void* y, const void* target, int dtype, int batch, int length,
                             int c, int h, int w, int kind, int dense_layout, void* s) {
	CB_REQUIRE_DEVICE();
	CB_ARG(err != nullptr && batch > 0 && c > 0);
	long long total = (long long)length * h * w * c;
	if (total <= 0) return CB200_OK;
	CB_DISPATCH_DTYPE(dtype, T, (output_error_elems_kernel<T><<<grid_for(total, 256), 256, 0, as_stream(s)>>>(
		err, (const T*)y, (const T*)target, batch, length, c, round8(c), h * w, kind, dense_layout)));
	CB_LAUNCH_CHECK();
	return CB200_OK;
}

int cb200_yolo_scatter_parts(float* err, const float* parts, int batch, int length, int cells, int nb_class, int nb_param, void* s) {
	CB_REQUIRE_DEVICE();
	CB_ARG(err != nullptr && parts != nullptr && batch > 0 && cells > 0);
	yolo_scatter_parts_kernel<<<(batch + 127) / 128, 128, 0, as_stream(s)>>>(err, parts, batch, length, cells, nb_class, nb_param);
	CB_LAUNCH_CHECK();
	return CB200_OK;
}

}  // extern "C"

// ---------------------------------------------------------------- classification read-out
// Per-sample argmax of the network output and of the target row (first maximum wins, like upstream's argmax /
// conv_argmax, src/auxil.c:1365-1426): what the confusion matrix of compute_error needs - two ints per sample come back
// instead of the whole output tensor.
namespace cb200 {
template <typename T>
__global__ void output_argmax_kernel(int* __restrict__ pred, int* __restrict__ truth, const T* __restrict__ y,
                                     const T* __restrict__ target, int length, int c, int cp, int hw) {
	__shared__ float best_v[2][32];
	__shared__ int best_i[2][32];
	const int b = blockIdx.x;
	const int n = c * hw;
	float v[2] = {-INFINITY, -INFINITY};
	int idx[2] = {0x7fffffff, 0x7fffffff};
	if (b < length) {
		for (int o = threadIdx.x; o < n; o += blockDim.x) {      // o = class-major index ch*hw + p, the target's own order
			const int ch = o / hw, p = o - ch * hw;
			const float a = to_f32<T>(y[((size_t)b * hw + p) * cp + ch]);
			const float t = to_f32<T>(target[(size_t)b * n + o]);
			if (a > v[0]) { v[0] = a; idx[0] = o; }
			if (t > v[1]) { v[1] = t; idx[1] = o; }
		}
	}
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
	for (int k = 0; k < 2; k++) {
		for (int off = 16; off > 0; off >>= 1) {
			const float ov = __shfl_xor_sync(0xffffffffu, v[k], off);
			const int oi = __shfl_xor_sync(0xffffffffu, idx[k], off);
			if (ov > v[k] || (ov == v[k] && oi < idx[k])) { v[k] = ov; idx[k] = oi; }
		}
		if (lane == 0) { best_v[k][wid] = v[k]; best_i[k][wid] = idx[k]; }
	}
	__syncthreads();
	if (threadIdx.x < 2) {
		const int k = threadIdx.x;
		float bv = best_v[k][0];
		int bi = best_i[k][0];
		for (int w = 1; w < (int)(blockDim.x >> 5); w++)
			if (best_v[k][w] > bv || (best_v[k][w] == bv && best_i[k][w] < bi)) { bv = best_v[k][w]; bi = best_i[k][w]; }
		(k == 0 ? pred : truth)[b] = b < length ? bi : -1;
	}
}
}  // namespace cb200

extern "C" int cb200_output_argmax(int* pred, int* truth, const void* y, const void* target, int dtype, int batch, int length,
                                   int c, int h, int w, void* s) {
	CB_REQUIRE_DEVICE();
	CB_ARG(pred != nullptr && truth != nullptr && batch > 0 && c > 0);
	CB_DISPATCH_DTYPE(dtype, T, (cb200::output_argmax_kernel<T><<<batch, 256, 0, cb200::as_stream(s)>>>(
		pred, truth, (const T*)y, (const T*)target, length, c, cb200::round8(c), h * w)));
	CB_LAUNCH_CHECK();
	return CB200_OK;
}
