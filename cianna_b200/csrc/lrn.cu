// lrn.cu - local response normalisation across channels, forward / backward, channels-last.
// Reference (CUDA only upstream, no CPU twin): src/cuda/cuda_lrn_layer.cu:35-101 kernels, :172-215 layer.
//   s_i = k + alpha/range * sum_{j in [i-range/2, i+range/2] ∩ [0,C)} x_j^2 ;  y_i = x_i / s_i^beta
//   dx_i = dy_i / s_i^beta - 2*alpha*beta/range * x_i * sum_j dy_j * y_j / s_j
// In the channels-last layout the channel window of a pixel is one contiguous span, so a thread that owns
// 8 channels reads at most 8 + range values from a single cache line pair instead of `range` strided planes.
#include "common.cuh"

namespace cb200 {

struct LrnGeom { int batch, length, c, cp, hw, range; float k, alpha, beta; };

template <typename T>
__global__ void __launch_bounds__(256)
lrn_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, float* __restrict__ scale, LrnGeom g) {
	const int cv = g.cp >> 3, half = g.range / 2;
	const long long total = (long long)g.batch * g.hw * cv;
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
		const int v = (int)(i % cv);
		const long long pix = i / cv;
		const T* row = x + pix * g.cp;
		float out[8];
#pragma unroll
		for (int j = 0; j < 8; j++) {
			const int ch = v * 8 + j;
			float o = 0.0f, s = 0.0f;
			if (ch < g.c) {
				const int lo = max(0, ch - half), hi = min(g.c - 1, ch + half);
				float sum = 0.0f;
				for (int q = lo; q <= hi; q++) { const float t = to_f32<T>(row[q]); sum += t * t; }
				s = g.k + g.alpha * sum / g.range;
				o = to_f32<T>(row[ch]) / powf(s, g.beta);
			}
			out[j] = o;
			if (scale != nullptr) scale[pix * g.cp + ch] = s;
		}
		store8<T>(y + pix * g.cp + v * 8, out);
	}
}

template <typename T>
__global__ void __launch_bounds__(256)
lrn_bwd_kernel(const T* __restrict__ x, const T* __restrict__ y, const T* __restrict__ dy, T* __restrict__ dx,
               const float* __restrict__ scale, const T* __restrict__ prev_out, cb200_activ prev_activ, LrnGeom g) {
	const int cv = g.cp >> 3, half = g.range / 2;
	const long long total = (long long)g.batch * g.hw * cv;
	const bool hook = prev_out != nullptr && prev_activ.type != CB200_LINEAR;
	const bool mask_tail = hook && activ_masks_tail(prev_activ);
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
		const int v = (int)(i % cv);
		const long long pix = i / cv;
		const int b = (int)(pix / g.hw);
		const long long base = pix * g.cp;
		float out[8];
#pragma unroll
		for (int j = 0; j < 8; j++) {
			const int ch = v * 8 + j;
			float o = 0.0f;
			if (ch < g.c) {
				const int lo = max(0, ch - half), hi = min(g.c - 1, ch + half);
				float sum = 0.0f;
				for (int q = lo; q <= hi; q++) sum += to_f32<T>(dy[base + q]) * to_f32<T>(y[base + q]) / scale[base + q];
				o = to_f32<T>(dy[base + ch]) / powf(scale[base + ch], g.beta) - 2.0f * g.alpha * g.beta * to_f32<T>(x[base + ch]) * sum / g.range;
				if (hook) o = (mask_tail && b >= g.length) ? 0.0f : activ_deriv_mul(prev_activ, o, to_f32<T>(prev_out[base + ch]));
			}
			out[j] = o;
		}
		store8<T>(dx + base + v * 8, out);
	}
}

static int fill(const cb200_lrn_desc* d, LrnGeom& g) {
	CB_ARG(d != nullptr && d->batch > 0 && d->c > 0 && d->range > 0);
	g.batch = d->batch; g.length = d->length; g.c = d->c; g.cp = round8(d->c); g.hw = d->h * d->w;
	g.range = d->range; g.k = d->k; g.alpha = d->alpha; g.beta = d->beta;
	return CB200_OK;
}
}  // namespace cb200
using namespace cb200;

extern "C" {
int cb200_lrn_forward(const cb200_lrn_desc* d, const void* x, void* y, float* local_scale, void* s) {
	CB_REQUIRE_DEVICE();
	LrnGeom g;
	int rc = fill(d, g); if (rc) return rc;
	long long total = (long long)g.batch * g.hw * (g.cp >> 3);
	CB_DISPATCH_DTYPE(d->dtype, T, (lrn_fwd_kernel<T><<<grid_for(total, 256), 256, 0, as_stream(s)>>>((const T*)x, (T*)y, local_scale, g)));
	CB_LAUNCH_CHECK();
	return CB200_OK;
}
int cb200_lrn_backward(const cb200_lrn_desc* d, const void* x, const void* y, const void* dy, void* dx, const float* local_scale,
                       const cb200_activ* prev_activ, const void* prev_out, void* s) {
	CB_REQUIRE_DEVICE();
	LrnGeom g;
	int rc = fill(d, g); if (rc) return rc;
	CB_ARG(local_scale != nullptr);
	cb200_activ pa; pa.type = CB200_LINEAR; pa.leak = 0; pa.saturation = 0; pa.beta = 0;
	if (prev_activ) pa = *prev_activ;
	long long total = (long long)g.batch * g.hw * (g.cp >> 3);
	CB_DISPATCH_DTYPE(d->dtype, T, (lrn_bwd_kernel<T><<<grid_for(total, 256), 256, 0, as_stream(s)>>>(
		(const T*)x, (const T*)y, (const T*)dy, (T*)dx, local_scale, (const T*)prev_out, pa, g)));
	CB_LAUNCH_CHECK();
	return CB200_OK;
}
}
