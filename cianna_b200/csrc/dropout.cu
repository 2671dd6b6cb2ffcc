// dropout.cu - dropout of conv / pool / dense layer outputs, in place, without a mask tensor.
// Reference: cuda_dropout_apply_* / cuda_dropout_scale_* and their call sites (src/cuda/cuda_conv_layer.cu:131-163,399-420,440-447,
// cuda_dense_layer.cu:78-112,375-390,405-411, cuda_pool_layer.cu:280-312,472-488,501-508): the layer's PRE-activation output is
// multiplied by a 0/1 mask (kept when a uniform draw >= drop_rate; no 1/(1-p) rescale while training), the activation runs on
// the result, the backward pass multiplies the layer's delta by the same mask first; inference multiplies by (1 - drop_rate)
// instead (AVG_MODEL) or draws a mask as in training (MC_MODEL).
// Upstream fills a FP32 mask tensor with cuRAND (4 B written + 2 x 4 B read per activation and step).  Here the mask is a pure
// function of (seed, draw counter, element position): forward and backward recompute it from the counter, nothing is stored,
// and the layer's activation is applied in the same pass (the producing kernel runs with a LINEAR epilogue).
#include "common.cuh"

namespace cb200 {

struct DropGeom { int batch, length, c, cp, hw; uint32_t threshold; float keep_scale; unsigned long long key; };

__host__ __device__ __forceinline__ unsigned long long drop_mix64(unsigned long long z) {
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}
// 8 independent 16-bit uniforms for the 8-channel packet `packet`: bit j of the result = element j is kept
__device__ __forceinline__ uint32_t keep_bits(unsigned long long key, long long packet, uint32_t threshold) {
	const unsigned long long a = drop_mix64(key + 0x9E3779B97F4A7C15ULL * (unsigned long long)(2 * packet + 1));
	const unsigned long long b = drop_mix64(a ^ 0xD1B54A32D192ED03ULL);
	uint32_t bits = 0;
#pragma unroll
	for (int j = 0; j < 4; j++) {
		bits |= (uint32_t)(((a >> (16 * j)) & 0xFFFFu) >= threshold) << j;
		bits |= (uint32_t)(((b >> (16 * j)) & 0xFFFFu) >= threshold) << (4 + j);
	}
	return bits;
}

// MODE 0: y = act(y * mask)   MODE 1: y = act(y * (1 - p))   MODE 2: dy = dy * mask (no activation)
template <typename T, int MODE>
__global__ void __launch_bounds__(256)
dropout_kernel(T* __restrict__ y, DropGeom g, cb200_activ act) {
	const int cv = g.cp >> 3;
	const long long total = (long long)g.batch * g.hw * cv;
	const bool zero_tail = MODE != 2 && activ_masks_tail(act);
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
		const int v = (int)(i % cv);
		const int b = (int)(i / ((long long)cv * g.hw));
		float val[8];
		unpack8(load_raw8<T>(y + i * 8), val);
		const uint32_t keep = MODE == 1 ? 0xFFu : keep_bits(g.key, i, g.threshold);
#pragma unroll
		for (int j = 0; j < 8; j++) {
			float z = ((keep >> j) & 1u) ? val[j] * g.keep_scale : 0.0f;
			if (MODE != 2) z = activ_forward(act, z);
			if (v * 8 + j >= g.c || (zero_tail && b >= g.length)) z = 0.0f;
			val[j] = z;
		}
		store8<T>(y + i * 8, val);
	}
}

static int fill(const cb200_dropout_desc* d, DropGeom& g, int mode) {
	CB_ARG(d != nullptr && d->batch > 0 && d->c > 0 && d->h > 0 && d->w > 0);
	CB_ARG(d->drop_rate >= 0.0f && d->drop_rate < 1.0f);
	g.batch = d->batch; g.length = d->length; g.c = d->c; g.cp = round8(d->c); g.hw = d->h * d->w;
	g.threshold = (uint32_t)(d->drop_rate * 65536.0f + 0.5f);
	g.keep_scale = mode == 1 ? 1.0f - d->drop_rate : 1.0f;
	g.key = drop_mix64(d->seed ^ drop_mix64(d->draw * 0x632BE59BD9B4E019ULL + (unsigned long long)d->stream_id));
	return CB200_OK;
}
}  // namespace cb200
using namespace cb200;

extern "C" {
int cb200_dropout_forward(const cb200_dropout_desc* d, void* y, int scale_only, void* s) {
	CB_REQUIRE_DEVICE();
	DropGeom g;
	int rc = fill(d, g, scale_only ? 1 : 0); if (rc) return rc;
	const long long total = (long long)g.batch * g.hw * (g.cp >> 3);
	if (scale_only) {
		CB_DISPATCH_DTYPE(d->dtype, T, (dropout_kernel<T, 1><<<grid_for(total, 256), 256, 0, as_stream(s)>>>((T*)y, g, d->activ)));
	} else {
		CB_DISPATCH_DTYPE(d->dtype, T, (dropout_kernel<T, 0><<<grid_for(total, 256), 256, 0, as_stream(s)>>>((T*)y, g, d->activ)));
	}
	CB_LAUNCH_CHECK();
	return CB200_OK;
}
int cb200_dropout_backward(const cb200_dropout_desc* d, void* dy, void* s) {
	CB_REQUIRE_DEVICE();
	DropGeom g;
	int rc = fill(d, g, 2); if (rc) return rc;
	const long long total = (long long)g.batch * g.hw * (g.cp >> 3);
	CB_DISPATCH_DTYPE(d->dtype, T, (dropout_kernel<T, 2><<<grid_for(total, 256), 256, 0, as_stream(s)>>>((T*)dy, g, d->activ)));
	CB_LAUNCH_CHECK();
	return CB200_OK;
}
}
