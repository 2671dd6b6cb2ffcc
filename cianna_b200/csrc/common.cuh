// common.cuh - shared device/host helpers of the B200-native CIANNA compute core.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/cianna_b200.h"

namespace cb200 {

// ---------------------------------------------------------------- error plumbing
void set_error(const char* fmt, ...);
extern bool g_have_device;
extern cudaStream_t g_stream;       // default compute stream of the core
extern int g_num_sms;
extern long long g_launches;

// opt-in per-kernel-family timing (runtime.cu). Families: see cb200_profile_collect in the header.
void prof_begin(int family, double work, cudaStream_t st);
void prof_end(cudaStream_t st);
enum { PROF_CONV_FWD_TC = 0, PROF_CONV_DGRAD_TC = 1, PROF_CONV_WGRAD_TC = 2, PROF_CONV_FWD_SIMT = 3, PROF_CONV_DGRAD_SIMT = 4,
       PROF_CONV_WGRAD_SIMT = 5, PROF_POOL = 6, PROF_NORM = 7, PROF_OPTIM = 8, PROF_OTHER = 9 };

inline cudaStream_t as_stream(void* s) { return s ? (cudaStream_t)s : g_stream; }

#define CB_REQUIRE_DEVICE()                                                               \
	do { if (!cb200::g_have_device) { cb200::set_error("%s: no CUDA device initialised (cb200_init); this core has no CPU fallback", __func__); \
	     return CB200_ERR_NO_DEVICE; } } while (0)

#define CB_CUDA(call)                                                                     \
	do { cudaError_t e__ = (call); if (e__ != cudaSuccess) {                               \
	     cb200::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
	     return CB200_ERR_CUDA; } } while (0)

#define CB_LAUNCH_CHECK()                                                                 \
	do { cb200::g_launches++; cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) { \
	     cb200::set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
	     return CB200_ERR_CUDA; } } while (0)

#define CB_ARG(cond)                                                                      \
	do { if (!(cond)) { cb200::set_error("%s: bad argument: %s", __func__, #cond); return CB200_ERR_ARG; } } while (0)

// dispatch a templated callable on the storage dtype
#define CB_DISPATCH_DTYPE(dtype, T, ...)                                                  \
	switch (dtype) {                                                                      \
		case CB200_FP32: { using T = float; __VA_ARGS__; break; }                         \
		case CB200_FP16: { using T = __half; __VA_ARGS__; break; }                        \
		case CB200_BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; }                 \
		default: cb200::set_error("%s: unknown dtype %d", __func__, (int)(dtype)); return CB200_ERR_ARG; \
	}

__host__ __device__ inline int round8(int c) { return (c + 7) & ~7; }
// A filter that covers its whole (unpadded) input map - a dense layer behind a conv / pool layer.  In the channels-last
// layout one image of the input is then a single run of f_h*f_w*Cp values in exactly the order of the filter's operand
// rows, so the layer is a 1x1 convolution over that many "channels"; its data-gradient operand keeps the same order
// (conv.cu: wbwd_row).
// depth geometry of cb200_conv_desc: 0 means 1 (a 2-D layer described by a zero-initialised tail of the struct)
__host__ __device__ inline int conv_in_d(const cb200_conv_desc* d) { return d->in_d > 0 ? d->in_d : 1; }
__host__ __device__ inline int conv_out_d(const cb200_conv_desc* d) { return d->out_d > 0 ? d->out_d : 1; }
__host__ __device__ inline int conv_f_d(const cb200_conv_desc* d) { return d->f_d > 0 ? d->f_d : 1; }
__host__ __device__ inline int conv_taps(const cb200_conv_desc* d) { return conv_f_d(d) * d->f_h * d->f_w; }
// 3-D geometry or internal padding (transposed convolution): served by the generic CUDA-core kernels only
__host__ __device__ inline bool conv_generic(const cb200_conv_desc* d) {
	return conv_in_d(d) > 1 || conv_out_d(d) > 1 || conv_f_d(d) > 1 || d->stride_d > 1 || d->pad_d > 0 ||
	       d->ipad_w > 0 || d->ipad_h > 0 || d->ipad_d > 0;
}
__host__ __device__ inline bool conv_whole_map(const cb200_conv_desc* d) {
	return !conv_generic(d) && d->out_h == 1 && d->out_w == 1 && d->pad_h == 0 && d->pad_w == 0 && d->f_h == d->in_h && d->f_w == d->in_w &&
	       d->f_h * d->f_w > 1 && d->input_is_patches == 0;
}
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------- scalar conversions
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ---------------------------------------------------------------- 8-channel vectors
// Every activation row is a multiple of 8 channels, so the natural access unit of the
// bandwidth kernels is "8 channels of one pixel": 16 B for 16-bit types, 32 B for FP32.
template <typename T> struct Vec8 { T v[8]; };

template <typename T>
__device__ __forceinline__ void load8(const T* __restrict__ p, float (&out)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float* __restrict__ p, float (&out)[8]) {
	float4 a = __ldg(reinterpret_cast<const float4*>(p));
	float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
	out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w;
	out[4] = b.x; out[5] = b.y; out[6] = b.z; out[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__half>(const __half* __restrict__ p, float (&out)[8]) {
	uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
	const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
	for (int i = 0; i < 4; i++) { float2 f = __half22float2(h[i]); out[2 * i] = f.x; out[2 * i + 1] = f.y; }
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* __restrict__ p, float (&out)[8]) {
	uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
	const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
	for (int i = 0; i < 4; i++) { float2 f = __bfloat1622float2(h[i]); out[2 * i] = f.x; out[2 * i + 1] = f.y; }
}

template <typename T>
__device__ __forceinline__ void store8(T* __restrict__ p, const float (&in)[8]);
template <>
__device__ __forceinline__ void store8<float>(float* __restrict__ p, const float (&in)[8]) {
	reinterpret_cast<float4*>(p)[0] = make_float4(in[0], in[1], in[2], in[3]);
	reinterpret_cast<float4*>(p)[1] = make_float4(in[4], in[5], in[6], in[7]);
}
template <>
__device__ __forceinline__ void store8<__half>(__half* __restrict__ p, const float (&in)[8]) {
	uint4 r;
	__half2* h = reinterpret_cast<__half2*>(&r);
#pragma unroll
	for (int i = 0; i < 4; i++) h[i] = __floats2half2_rn(in[2 * i], in[2 * i + 1]);
	*reinterpret_cast<uint4*>(p) = r;
}
template <>
__device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* __restrict__ p, const float (&in)[8]) {
	uint4 r;
	__nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
	for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(in[2 * i], in[2 * i + 1]);
	*reinterpret_cast<uint4*>(p) = r;
}

// raw (unconverted) 8-channel packet: lets a thread issue several independent 128-bit loads before it
// starts converting, which is what keeps enough bytes in flight per SM for HBM3e (Little's law)
template <typename T> struct Raw8 { uint4 a; };
template <> struct Raw8<float> { float4 a, b; };
template <typename T> __device__ __forceinline__ Raw8<T> load_raw8(const T* __restrict__ p) {
	Raw8<T> r; r.a = __ldg(reinterpret_cast<const uint4*>(p)); return r;
}
template <> __device__ __forceinline__ Raw8<float> load_raw8<float>(const float* __restrict__ p) {
	Raw8<float> r; r.a = __ldg(reinterpret_cast<const float4*>(p)); r.b = __ldg(reinterpret_cast<const float4*>(p) + 1); return r;
}
__device__ __forceinline__ void unpack8(const Raw8<float>& r, float (&o)[8]) {
	o[0] = r.a.x; o[1] = r.a.y; o[2] = r.a.z; o[3] = r.a.w; o[4] = r.b.x; o[5] = r.b.y; o[6] = r.b.z; o[7] = r.b.w;
}
__device__ __forceinline__ void unpack8(const Raw8<__half>& r, float (&o)[8]) {
	const __half2* h = reinterpret_cast<const __half2*>(&r.a);
#pragma unroll
	for (int i = 0; i < 4; i++) { float2 f = __half22float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
}
__device__ __forceinline__ void unpack8(const Raw8<__nv_bfloat16>& r, float (&o)[8]) {
	const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r.a);
#pragma unroll
	for (int i = 0; i < 4; i++) { float2 f = __bfloat1622float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
}

// ---------------------------------------------------------------- activations
// Forward: reference ReLU_activation_kernel / logistic_activation_kernel
// (src/cuda/cuda_activ_functions.cu:37-70, 203-235); LINEAR is the identity.
__device__ __forceinline__ float activ_forward(const cb200_activ& a, float z) {
	if (a.type == CB200_RELU) {
		if (z <= 0.0f) return z * a.leak;
		if (z > a.saturation) return a.saturation + (z - a.saturation) * a.leak;
		return z;
	}
	if (a.type == CB200_LOGISTIC) {
		float t = -a.beta * z;
		if (t > a.saturation) t = a.saturation;   // upstream clamps the exponent argument
		return 1.0f / (1.0f + expf(t));
	}
	return z;
}
// Backward hook: multiply a delta by act'(.) evaluated from the ACTIVATED output y, the way the
// reference's deriv kernels do (cuda_activ_functions.cu:72-111, 237-270).
__device__ __forceinline__ float activ_deriv_mul(const cb200_activ& a, float delta, float y) {
	if (a.type == CB200_RELU) {
		if (y <= 0.0f || y > a.saturation) return delta * a.leak;
		return delta;
	}
	if (a.type == CB200_LOGISTIC) return delta * a.beta * y * (1.0f - y);
	return delta;
}
// non-linear activations zero the samples of a partially filled batch
__host__ __device__ __forceinline__ bool activ_masks_tail(const cb200_activ& a) {
	return a.type == CB200_RELU || a.type == CB200_LOGISTIC || a.type == CB200_SOFTMAX;
}

// One SGD + momentum + weight-decay step on one weight (upstream: cuda_update_weights, src/cuda/cuda_main.cu:486-509, with
// the loss-scale factor S carried by gradient and momentum).  Written with explicit round-to-nearest intrinsics so that
// every kernel that calls it - the per-layer update kernels and the whole-network plan kernels (update_plan.cu) - gets the
// same FMA contraction and therefore the same bits.
__device__ __forceinline__ void sgd_momentum_step(float alpha, float mom, float wdlr, float S, float g, float& m, float& w) {
	float t = __fmaf_rn(alpha, g, __fmul_rn(mom, m));
	t = __fmaf_rn(__fmul_rn(wdlr, w), S, t);
	w = __fsub_rn(w, __fdiv_rn(t, S));
	m = t;
}
// group-norm parameter step (no weight decay): upd = mom * upd + alpha * sum ; param -= upd / S
__device__ __forceinline__ void norm_param_step(float alpha, float mom, float S, float sum, float& upd, float& param) {
	const float u = __fmaf_rn(mom, upd, __fmul_rn(alpha, sum));
	upd = u;
	param = __fsub_rn(param, __fdiv_rn(u, S));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
	return v;
}

inline int grid_for(long long work_items, int threads, int max_waves = 8) {
	long long blocks = ceil_div_ll(work_items, threads);
	long long cap = (long long)g_num_sms * max_waves * (2048 / threads);
	if (blocks > cap) blocks = cap;
	if (blocks < 1) blocks = 1;
	return (int)blocks;
}

}  // namespace cb200
