// runtime.cu - device bring-up, memory/stream/event plumbing and layout import/export
// of the B200-native CIANNA core.  Replaces src/cuda/cuda_main.cu of the reference
// (init_cuda :922-1067, typed alloc/copy helpers :108-379, event timers :533-588).
#include <stdarg.h>
#include "common.cuh"

namespace cb200 {
bool g_have_device = false;
cudaStream_t g_stream = nullptr;
int g_num_sms = 148;
long long g_launches = 0;
static char g_error[1024] = "no error";
const char* g_last_conv_impl = "none";
int g_force_simt = 0;
extern int g_disable_halo;
extern int g_enable_cluster;
extern int g_enable_pair;
extern int g_enable_pair_wgrad;

// ---- opt-in per-kernel-family timing (CUDA event pairs on the launching stream, harvested lazily:
// replaces the always-on per-layer cudaEventSynchronize of upstream's perf_eval, src/auxil.c:698-766)
struct ProfRec { cudaEvent_t a, b; int family; double work; };
static const int PROF_MAX = 16384;
static ProfRec* g_prof = nullptr;
static int g_prof_n = 0;
int g_prof_on = 0;

void prof_begin(int family, double work, cudaStream_t st) {
	if (!g_prof_on || g_prof_n >= PROF_MAX) return;
	if (!g_prof) g_prof = (ProfRec*)calloc(PROF_MAX, sizeof(ProfRec));
	ProfRec& r = g_prof[g_prof_n];
	if (!r.a) { cudaEventCreate(&r.a); cudaEventCreate(&r.b); }
	r.family = family; r.work = work;
	cudaEventRecord(r.a, st);
}
void prof_end(cudaStream_t st) {
	if (!g_prof_on || g_prof_n >= PROF_MAX) return;
	cudaEventRecord(g_prof[g_prof_n].b, st);
	g_prof_n++;
}

void set_error(const char* fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_error, sizeof(g_error), fmt, ap);
	va_end(ap);
}
}  // namespace cb200
using namespace cb200;

extern "C" {

const char* cb200_last_error(void) { return g_error; }
const char* cb200_version(void) { return "cianna_b200 0.1 (sm_100a)"; }
const char* cb200_last_conv_impl(void) { return g_last_conv_impl; }
static int g_pair_default = 1;     // env CB200_CTA_PAIR=0: keep the wide-N layers on the one-SM kernel
static int g_pair_wgrad_default = 0;   // env CB200_WGRAD_PAIR=1: weight gradient of the wide layers on CTA pairs (measured 3 % slower)
void cb200_force_simt(int on) {
	g_force_simt = on & 1; g_disable_halo = (on >> 1) & 1; g_enable_cluster = (on >> 2) & 1;
	g_enable_pair = (on & 16) ? 0 : ((on & 8) ? 1 : g_pair_default);
	g_enable_pair_wgrad = (on & 16) ? 0 : ((on & 8) ? 1 : g_pair_wgrad_default);
}
long long cb200_launch_count(int reset) { long long v = g_launches; if (reset) g_launches = 0; return v; }
void cb200_profile_enable(int on) { g_prof_on = on; }
int cb200_profile_collect(int family, double* ms, double* work, long long* launches) {
	CB_REQUIRE_DEVICE();
	CB_CUDA(cudaDeviceSynchronize());
	double t = 0.0, w = 0.0; long long n = 0;
	for (int i = 0; i < g_prof_n; i++) {
		if (g_prof[i].family != family) continue;
		float e = 0.0f;
		if (cudaEventElapsedTime(&e, g_prof[i].a, g_prof[i].b) == cudaSuccess) { t += e; w += g_prof[i].work; n++; }
	}
	if (ms) *ms = t;
	if (work) *work = w;
	if (launches) *launches = n;
	return CB200_OK;
}
void cb200_profile_reset(void) { g_prof_n = 0; }
int cb200_round_channels(int c) { return round8(c); }
size_t cb200_dtype_size(int dtype) { return dtype == CB200_FP32 ? 4 : 2; }

int cb200_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

int cb200_init(int device) {
	int n = cb200_device_count();
	if (n <= 0) { set_error("cb200_init: no CUDA device visible; this core has no CPU fallback"); return CB200_ERR_NO_DEVICE; }
	if (device >= 0) CB_CUDA(cudaSetDevice(device));
	int dev = 0;
	CB_CUDA(cudaGetDevice(&dev));
	cudaDeviceProp prop;
	CB_CUDA(cudaGetDeviceProperties(&prop, dev));
	if (prop.major != 10) {
		set_error("cb200_init: device %d is sm_%d%d; this library only carries sm_100a code", dev, prop.major, prop.minor);
		return CB200_ERR_UNSUPPORTED;
	}
	g_num_sms = prop.multiProcessorCount;
	// debugging aid: CB200_FORCE_SIMT=1 routes every conv through the generic kernels (see cb200_force_simt)
	const char* fs = getenv("CB200_FORCE_SIMT");
	if (fs && fs[0] == '1') g_force_simt = 1;
	const char* cp = getenv("CB200_CTA_PAIR");
	g_pair_default = (cp && cp[0] == '0') ? 0 : 1;
	g_enable_pair = g_pair_default;
	const char* wp = getenv("CB200_WGRAD_PAIR");
	g_pair_wgrad_default = (wp && wp[0] == '1') ? 1 : 0;
	g_enable_pair_wgrad = g_pair_wgrad_default;
	if (!g_stream) {
		// the compute stream carries the critical path: highest priority, so that work put on side streams (weight gradients)
		// only takes the SMs the critical path leaves free
		int lo = 0, hi = 0;
		CB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
		CB_CUDA(cudaStreamCreateWithPriority(&g_stream, cudaStreamNonBlocking, hi));
	}
	g_have_device = true;
	return CB200_OK;
}

// ---------------------------------------------------------------- memory
int cb200_malloc(void** p, size_t bytes) {
	CB_REQUIRE_DEVICE();
	if (bytes == 0) bytes = 16;
	CB_CUDA(cudaMalloc(p, bytes));
	CB_CUDA(cudaMemsetAsync(*p, 0, bytes, g_stream));
	return CB200_OK;
}
int cb200_free(void* p) { CB_REQUIRE_DEVICE(); CB_CUDA(cudaFree(p)); return CB200_OK; }
int cb200_host_alloc(void** p, size_t bytes) {
	CB_REQUIRE_DEVICE();
	CB_CUDA(cudaMallocHost(p, bytes ? bytes : 16));
	memset(*p, 0, bytes);
	return CB200_OK;
}
int cb200_host_free(void* p) { CB_REQUIRE_DEVICE(); CB_CUDA(cudaFreeHost(p)); return CB200_OK; }
int cb200_memset(void* p, int byte, size_t bytes, void* s) {
	CB_REQUIRE_DEVICE(); CB_CUDA(cudaMemsetAsync(p, byte, bytes, as_stream(s))); return CB200_OK;
}
int cb200_h2d(void* d, const void* h, size_t bytes, void* s) {
	CB_REQUIRE_DEVICE(); CB_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, as_stream(s))); return CB200_OK;
}
int cb200_d2h(void* h, const void* d, size_t bytes, void* s) {
	CB_REQUIRE_DEVICE(); CB_CUDA(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, as_stream(s))); return CB200_OK;
}
int cb200_d2d(void* d, const void* s_, size_t bytes, void* s) {
	CB_REQUIRE_DEVICE(); CB_CUDA(cudaMemcpyAsync(d, s_, bytes, cudaMemcpyDeviceToDevice, as_stream(s))); return CB200_OK;
}
int cb200_stream_create(void** s) {
	CB_REQUIRE_DEVICE(); cudaStream_t st; CB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)); *s = st; return CB200_OK;
}
int cb200_stream_create_low_priority(void** s) {
	CB_REQUIRE_DEVICE();
	int lo = 0, hi = 0;
	CB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));     // lo = numerically largest = least urgent
	cudaStream_t st; CB_CUDA(cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, lo)); *s = st; return CB200_OK;
}
int cb200_stream_destroy(void* s) { CB_REQUIRE_DEVICE(); CB_CUDA(cudaStreamDestroy((cudaStream_t)s)); return CB200_OK; }
int cb200_stream_sync(void* s) { CB_REQUIRE_DEVICE(); CB_CUDA(cudaStreamSynchronize(as_stream(s))); return CB200_OK; }
int cb200_device_sync(void) { CB_REQUIRE_DEVICE(); CB_CUDA(cudaDeviceSynchronize()); return CB200_OK; }
int cb200_stream_wait(void* s, void* on) {
	CB_REQUIRE_DEVICE();
	cudaEvent_t ev;
	CB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
	CB_CUDA(cudaEventRecord(ev, as_stream(on)));
	CB_CUDA(cudaStreamWaitEvent(as_stream(s), ev, 0));
	CB_CUDA(cudaEventDestroy(ev));
	return CB200_OK;
}
int cb200_event_create(void** ev) { CB_REQUIRE_DEVICE(); cudaEvent_t e; CB_CUDA(cudaEventCreate(&e)); *ev = e; return CB200_OK; }
int cb200_event_destroy(void* ev) { CB_REQUIRE_DEVICE(); CB_CUDA(cudaEventDestroy((cudaEvent_t)ev)); return CB200_OK; }
int cb200_event_record(void* ev, void* s) { CB_REQUIRE_DEVICE(); CB_CUDA(cudaEventRecord((cudaEvent_t)ev, as_stream(s))); return CB200_OK; }
int cb200_stream_wait_event(void* s, void* ev) { CB_REQUIRE_DEVICE(); CB_CUDA(cudaStreamWaitEvent(as_stream(s), (cudaEvent_t)ev, 0)); return CB200_OK; }
int cb200_event_sync(void* ev) { CB_REQUIRE_DEVICE(); CB_CUDA(cudaEventSynchronize((cudaEvent_t)ev)); return CB200_OK; }
int cb200_event_elapsed_ms(void* a, void* b, float* ms) {
	CB_REQUIRE_DEVICE();
	CB_CUDA(cudaEventSynchronize((cudaEvent_t)b));
	CB_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)a, (cudaEvent_t)b));
	return CB200_OK;
}

}  // extern "C"

// ---------------------------------------------------------------- casts
template <typename T>
__global__ void cast_from_f32_kernel(T* __restrict__ dst, const float* __restrict__ src, size_t n) {
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
		dst[i] = from_f32<T>(src[i]);
}
template <typename T>
__global__ void cast_to_f32_kernel(float* __restrict__ dst, const T* __restrict__ src, size_t n) {
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
		dst[i] = to_f32<T>(src[i]);
}

// dataset conversion on the device: round toward zero, value for value what the reference's host loop produces
// (copy_to_FP16 / copy_to_BF16 with __float2half_rz / __float2bfloat16_rz, src/cuda/cuda_main.cu:108-113,790,813)
template <typename T> __device__ __forceinline__ T from_f32_rz(float v);
template <> __device__ __forceinline__ float from_f32_rz<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32_rz<__half>(float v) { return __float2half_rz(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32_rz<__nv_bfloat16>(float v) { return __float2bfloat16_rz(v); }
template <typename T>
__global__ void cast_from_f32_rz_kernel(T* __restrict__ dst, const float* __restrict__ src, size_t n) {
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
		dst[i] = from_f32_rz<T>(src[i]);
}

// FP32 sample rows [rows][src_row] -> typed dataset rows [rows][dst_row] (dst_row >= src_row; the extra slots, the bias
// slot of an input row, take tail_value), round toward zero
template <typename T>
__global__ void dataset_pack_kernel(T* __restrict__ dst, const float* __restrict__ src, size_t rows, size_t src_row, size_t dst_row, float tail_value) {
	const size_t total = rows * dst_row;
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
		const size_t r = i / dst_row, c = i - r * dst_row;
		dst[i] = from_f32_rz<T>(c < src_row ? src[r * src_row + c] : tail_value);
	}
}

// dataset shuffle: row i of the source batches goes to row index[i] of the destination batches (index == nullptr:
// identity = the copy back).  One warp per row, 16-byte chunks when the row pitch allows, else 2- or 4-byte elements.
// Replaces shfl_kern / get_back_shuffle (src/cuda/cuda_main.cu:590-640: one THREAD per row there, uncoalesced).
template <typename V>
__global__ void rows_permute_kernel(void* const* __restrict__ dst_batches, void* const* __restrict__ src_batches,
                                    const int* __restrict__ index, long long n_rows, int batch_size, size_t row_units) {
	const int lane = threadIdx.x & 31;
	const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
	const long long nwarp = (gridDim.x * (long long)blockDim.x) >> 5;
	for (long long i = warp0; i < n_rows; i += nwarp) {
		const long long j = index != nullptr ? (long long)index[i] : i;
		const V* s = (const V*)src_batches[i / batch_size] + (size_t)(i % batch_size) * row_units;
		V* d = (V*)dst_batches[j / batch_size] + (size_t)(j % batch_size) * row_units;
		for (size_t k = lane; k < row_units; k += 32) d[k] = s[k];
	}
}

// ---------------------------------------------------------------- layout conversions
// dataset row [C*H*W + 1] (channel-major planes, bias slot last) -> act[b][y][x][Cp]
template <typename T>
__global__ void import_input_kernel(T* __restrict__ dst, const T* __restrict__ src, int batch, int c, int hw, int cp) {
	size_t total = (size_t)batch * hw * cp;
	size_t row = (size_t)c * hw + 1;
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
		int ch = (int)(i % cp);
		size_t pix = i / cp;
		int p = (int)(pix % hw);
		size_t b = pix / hw;
		dst[i] = ch < c ? src[b * row + (size_t)ch * hw + p] : from_f32<T>(0.0f);
	}
}
// dataset rows -> first-layer patch rows (see cb200_import_input_patches in the header)
// grid: (x: tiles of output columns, y: output row, z: sample).  A block first stages the f_h input row segments of
// every channel that its output columns need in shared memory (coalesced reads of the planar dataset row), then each
// thread assembles one 8-column packet of a patch row from shared memory; the column -> (channel, ky, kx) decode is a
// small table, so there is no integer division and no scattered global load in the inner loop.
template <typename T>
__global__ void __launch_bounds__(256)
import_patches_kernel(T* __restrict__ dst, const T* __restrict__ src, int c, int h, int w,
                      int f_h, int f_w, int s_h, int s_w, int p_h, int p_w, int out_h, int out_w, int kp, float bias_value,
                      int tile_x, int seg_w) {
	extern __shared__ unsigned char smem_u8[];
	int* tab = reinterpret_cast<int*>(smem_u8);                 // [kp]: (channel << 16 | ky << 8 | kx), -1 = bias, -2 = zero pad
	T* seg = reinterpret_cast<T*>(smem_u8 + 1024);              // [c][f_h][seg_w]
	const int taps = f_h * f_w, kreal = c * taps;
	const int oy = blockIdx.y;
	const size_t b = blockIdx.z;
	const int ox0 = blockIdx.x * tile_x;
	const int ix0 = ox0 * s_w - p_w;                            // first input column of the segment
	for (int col = threadIdx.x; col < kp; col += blockDim.x) {
		int code = -2;
		if (col < kreal) { const int ch = col / taps, tap = col - ch * taps; const int ky = tap / f_w, kx = tap - ky * f_w; code = (ch << 16) | (ky << 8) | kx; }
		else if (col == kreal) code = -1;
		tab[col] = code;
	}
	const T* img = src + b * ((size_t)c * h * w + 1);
	const int rows = c * f_h;
	for (int i = threadIdx.x; i < rows * seg_w; i += blockDim.x) {
		const int r = i / seg_w, xx = i - r * seg_w;
		const int ch = r / f_h, ky = r - ch * f_h;
		const int iy = oy * s_h - p_h + ky, ix = ix0 + xx;
		seg[i] = (iy >= 0 && iy < h && ix >= 0 && ix < w) ? img[((size_t)ch * h + iy) * w + ix] : from_f32<T>(0.0f);
	}
	__syncthreads();
	const int kv = kp >> 3;
	const int v = threadIdx.x % kv, lx = threadIdx.x / kv;
	const int ox = ox0 + lx;
	if (lx >= tile_x || ox >= out_w) return;
	float o[8];
#pragma unroll
	for (int j = 0; j < 8; j++) {
		const int code = tab[v * 8 + j];
		float val = 0.0f;
		if (code >= 0) {
			const int ch = code >> 16, ky = (code >> 8) & 0xff, kx = code & 0xff;
			val = to_f32<T>(seg[(ch * f_h + ky) * seg_w + lx * s_w + kx]);
		} else if (code == -1) val = bias_value;
		o[j] = val;
	}
	store8<T>(dst + (((b * out_h + oy) * out_w + ox) * (size_t)kp) + v * 8, o);
}

// reference activation layout [C][B][HW] (FP32) -> internal
template <typename T>
__global__ void import_cbhw_kernel(T* __restrict__ dst, const float* __restrict__ src, int batch, int c, int hw, int cp) {
	size_t total = (size_t)batch * hw * cp;
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
		int ch = (int)(i % cp);
		size_t pix = i / cp;
		int p = (int)(pix % hw);
		size_t b = pix / hw;
		dst[i] = ch < c ? from_f32<T>(src[((size_t)ch * batch + b) * hw + p]) : from_f32<T>(0.0f);
	}
}
template <typename T>
__global__ void export_cbhw_kernel(float* __restrict__ dst, const T* __restrict__ src, int batch, int c, int hw, int cp) {
	size_t total = (size_t)c * batch * hw;
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
		int p = (int)(i % hw);
		size_t r = i / hw;
		int b = (int)(r % batch);
		int ch = (int)(r / batch);
		dst[i] = to_f32<T>(src[((size_t)b * hw + p) * cp + ch]);
	}
}
__global__ void export_pool_map_kernel(int32_t* __restrict__ dst, const uint8_t* __restrict__ src, int batch, int c, int hw, int cp) {
	size_t total = (size_t)c * batch * hw;
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
		int p = (int)(i % hw);
		size_t r = i / hw;
		int b = (int)(r % batch);
		int ch = (int)(r / batch);
		uint8_t v = src[((size_t)b * hw + p) * cp + ch];
		dst[i] = v == 255 ? -1 : (int32_t)v;
	}
}
template <typename T>
__global__ void export_dense_kernel(float* __restrict__ dst, const T* __restrict__ src, int batch, int n, int cp, float bias_node) {
	size_t total = (size_t)batch * (n + 1);
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
		int j = (int)(i % (n + 1));
		size_t b = i / (n + 1);
		dst[i] = j < n ? to_f32<T>(src[b * cp + j]) : bias_node;
	}
}

extern "C" {

int cb200_cast_from_f32(void* dst, int dtype, const float* src, size_t n, void* s) {
	CB_REQUIRE_DEVICE();
	if (n == 0) return CB200_OK;
	CB_DISPATCH_DTYPE(dtype, T, (cast_from_f32_kernel<T><<<grid_for((long long)n, 256), 256, 0, as_stream(s)>>>((T*)dst, src, n)));
	CB_LAUNCH_CHECK();
	return CB200_OK;
}
int cb200_cast_to_f32(float* dst, const void* src, int dtype, size_t n, void* s) {
	CB_REQUIRE_DEVICE();
	if (n == 0) return CB200_OK;
	CB_DISPATCH_DTYPE(dtype, T, (cast_to_f32_kernel<T><<<grid_for((long long)n, 256), 256, 0, as_stream(s)>>>(dst, (const T*)src, n)));
	CB_LAUNCH_CHECK();
	return CB200_OK;
}

int cb200_cast_from_f32_rz(void* dst, int dtype, const float* src, size_t n, void* s) {
	CB_REQUIRE_DEVICE();
	if (n == 0) return CB200_OK;
	CB_DISPATCH_DTYPE(dtype, T, (cast_from_f32_rz_kernel<T><<<grid_for((long long)n, 256), 256, 0, as_stream(s)>>>((T*)dst, src, n)));
	CB_LAUNCH_CHECK();
	return CB200_OK;
}

int cb200_dataset_pack(void* dst, int dtype, const float* src, size_t rows, size_t src_row, size_t dst_row, float tail_value, void* s) {
	CB_REQUIRE_DEVICE();
	CB_ARG(dst_row >= src_row);
	if (rows == 0 || dst_row == 0) return CB200_OK;
	CB_DISPATCH_DTYPE(dtype, T, (dataset_pack_kernel<T><<<grid_for((long long)(rows * dst_row), 256), 256, 0, as_stream(s)>>>((T*)dst, src, rows, src_row, dst_row, tail_value)));
	CB_LAUNCH_CHECK();
	return CB200_OK;
}

int cb200_rows_permute(void* const* dst_batches, void* const* src_batches, const int* index, long long n_rows, int batch_size,
                       size_t row_bytes, void* s) {
	CB_REQUIRE_DEVICE();
	CB_ARG(dst_batches != nullptr && src_batches != nullptr && batch_size > 0);
	if (n_rows <= 0 || row_bytes == 0) return CB200_OK;
	const int threads = 256;
	long long blocks = (n_rows * 32 + threads - 1) / threads;
	if (blocks > 148 * 16) blocks = 148 * 16;
	// every batch comes from cudaMalloc (256-byte aligned): the row pitch alone decides the widest safe access
	if (row_bytes % 16 == 0) rows_permute_kernel<uint4><<<(unsigned)blocks, threads, 0, as_stream(s)>>>(dst_batches, src_batches, index, n_rows, batch_size, row_bytes / 16);
	else if (row_bytes % 4 == 0) rows_permute_kernel<uint32_t><<<(unsigned)blocks, threads, 0, as_stream(s)>>>(dst_batches, src_batches, index, n_rows, batch_size, row_bytes / 4);
	else if (row_bytes % 2 == 0) rows_permute_kernel<uint16_t><<<(unsigned)blocks, threads, 0, as_stream(s)>>>(dst_batches, src_batches, index, n_rows, batch_size, row_bytes / 2);
	else rows_permute_kernel<uint8_t><<<(unsigned)blocks, threads, 0, as_stream(s)>>>(dst_batches, src_batches, index, n_rows, batch_size, row_bytes);
	CB_LAUNCH_CHECK();
	return CB200_OK;
}

// round-toward-zero host conversion (the reference converts datasets on the host with
// __float2half_rz / __float2bfloat16_rz, src/cuda/cuda_main.cu:790,813)
static inline uint16_t f32_to_bf16_rz(float f) { uint32_t u; memcpy(&u, &f, 4); return (uint16_t)(u >> 16); }
static inline uint16_t f32_to_f16_rz(float f) {
	uint32_t u; memcpy(&u, &f, 4);
	uint32_t sign = (u >> 16) & 0x8000u;
	int32_t exp = (int32_t)((u >> 23) & 0xff) - 127 + 15;
	uint32_t man = u & 0x7fffffu;
	if (((u >> 23) & 0xff) == 0xff) return (uint16_t)(sign | 0x7c00u | (man ? 0x200u : 0));  // inf / nan
	if (exp >= 31) return (uint16_t)(sign | 0x7bffu);                                       // RZ: clamp to max finite
	if (exp <= 0) {
		if (exp < -10) return (uint16_t)sign;
		man |= 0x800000u;
		return (uint16_t)(sign | (man >> (14 - exp)));                                      // subnormal, truncated
	}
	return (uint16_t)(sign | ((uint32_t)exp << 10) | (man >> 13));
}
int cb200_host_cast_from_f32(void* dst, int dtype, const float* src, size_t n) {
	if (dtype == CB200_FP32) { memcpy(dst, src, n * 4); return CB200_OK; }
	uint16_t* o = (uint16_t*)dst;
	if (dtype == CB200_FP16) for (size_t i = 0; i < n; i++) o[i] = f32_to_f16_rz(src[i]);
	else if (dtype == CB200_BF16) for (size_t i = 0; i < n; i++) o[i] = f32_to_bf16_rz(src[i]);
	else { set_error("cb200_host_cast_from_f32: unknown dtype %d", dtype); return CB200_ERR_ARG; }
	return CB200_OK;
}

static inline float f16_to_f32(uint16_t h) {
	uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1f, man = h & 0x3ffu, u;
	if (exp == 0) {
		if (man == 0) u = sign;
		else { int e = -1; do { e++; man <<= 1; } while (!(man & 0x400u)); u = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ffu) << 13); }
	} else if (exp == 31) u = sign | 0x7f800000u | (man << 13);
	else u = sign | ((exp - 15 + 127) << 23) | (man << 13);
	float f; memcpy(&f, &u, 4); return f;
}
int cb200_host_cast_to_f32(float* dst, const void* src, int dtype, size_t n) {
	if (dtype == CB200_FP32) { memcpy(dst, src, n * 4); return CB200_OK; }
	const uint16_t* in = (const uint16_t*)src;
	if (dtype == CB200_FP16) for (size_t i = 0; i < n; i++) dst[i] = f16_to_f32(in[i]);
	else if (dtype == CB200_BF16) for (size_t i = 0; i < n; i++) { uint32_t u = (uint32_t)in[i] << 16; memcpy(&dst[i], &u, 4); }
	else { set_error("cb200_host_cast_to_f32: unknown dtype %d", dtype); return CB200_ERR_ARG; }
	return CB200_OK;
}

int cb200_import_input(void* dst, const void* src, int dtype, int batch, int c, int h, int w, void* s) {
	CB_REQUIRE_DEVICE();
	int cp = round8(c);
	long long total = (long long)batch * h * w * cp;
	CB_DISPATCH_DTYPE(dtype, T, (import_input_kernel<T><<<grid_for(total, 256), 256, 0, as_stream(s)>>>((T*)dst, (const T*)src, batch, c, h * w, cp)));
	CB_LAUNCH_CHECK();
	return CB200_OK;
}
int cb200_patch_width(int c, int f_h, int f_w) { int k = c * f_h * f_w + 1; return (k + 15) & ~15; }
int cb200_import_input_patches(void* dst, const void* src, int dtype, int batch, int c, int h, int w, int f_h, int f_w,
                               int stride_h, int stride_w, int pad_h, int pad_w, int out_h, int out_w, float bias_value, void* s) {
	CB_REQUIRE_DEVICE();
	const int kp = cb200_patch_width(c, f_h, f_w);
	CB_ARG(kp <= 256 && f_h < 256 && f_w < 256);
	const int tile_x = 256 / (kp >> 3);
	const int seg_w = (tile_x - 1) * stride_w + f_w;
	const size_t smem = 1024 + (size_t)c * f_h * seg_w * cb200_dtype_size(dtype);
	CB_ARG(smem <= 48 * 1024);
	dim3 grid((unsigned)ceil_div(out_w, tile_x), (unsigned)out_h, (unsigned)batch);
	CB_DISPATCH_DTYPE(dtype, T, (import_patches_kernel<T><<<grid, 256, smem, as_stream(s)>>>(
		(T*)dst, (const T*)src, c, h, w, f_h, f_w, stride_h, stride_w, pad_h, pad_w, out_h, out_w, kp, bias_value, tile_x, seg_w)));
	CB_LAUNCH_CHECK();
	return CB200_OK;
}
int cb200_import_cbhw(void* dst, int dtype, const float* src, int batch, int c, int h, int w, void* s) {
	CB_REQUIRE_DEVICE();
	int cp = round8(c);
	long long total = (long long)batch * h * w * cp;
	CB_DISPATCH_DTYPE(dtype, T, (import_cbhw_kernel<T><<<grid_for(total, 256), 256, 0, as_stream(s)>>>((T*)dst, src, batch, c, h * w, cp)));
	CB_LAUNCH_CHECK();
	return CB200_OK;
}
int cb200_export_cbhw(float* dst, const void* src, int dtype, int batch, int c, int h, int w, void* s) {
	CB_REQUIRE_DEVICE();
	int cp = round8(c);
	long long total = (long long)batch * h * w * c;
	CB_DISPATCH_DTYPE(dtype, T, (export_cbhw_kernel<T><<<grid_for(total, 256), 256, 0, as_stream(s)>>>(dst, (const T*)src, batch, c, h * w, cp)));
	CB_LAUNCH_CHECK();
	return CB200_OK;
}
int cb200_export_pool_map(int32_t* dst, const uint8_t* src, int batch, int c, int h, int w, void* s) {
	CB_REQUIRE_DEVICE();
	int cp = round8(c);
	long long total = (long long)batch * h * w * c;
	export_pool_map_kernel<<<grid_for(total, 256), 256, 0, as_stream(s)>>>(dst, src, batch, c, h * w, cp);
	CB_LAUNCH_CHECK();
	return CB200_OK;
}
int cb200_export_dense(float* dst, const void* src, int dtype, int batch, int n, float bias_node, void* s) {
	CB_REQUIRE_DEVICE();
	int cp = round8(n);
	long long total = (long long)batch * (n + 1);
	CB_DISPATCH_DTYPE(dtype, T, (export_dense_kernel<T><<<grid_for(total, 256), 256, 0, as_stream(s)>>>(dst, (const T*)src, batch, n, cp, bias_node)));
	CB_LAUNCH_CHECK();
	return CB200_OK;
}

}  // extern "C"
