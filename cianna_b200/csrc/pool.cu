// pool.cu - max / average pooling, forward and backward, on channels-last tensors.
// One thread handles 8 consecutive channels of one pixel (128-bit accesses for 16-bit types),
// so a warp always touches one contiguous span of memory.
// Reference: src/cuda/cuda_pool_layer.cu:31-277 (kernels), :429-547 (layer functions);
// CPU twin src/naiv/naiv_pool_layer.c:30-318.
#include "common.cuh"

namespace cb200 {

struct PoolGeom {
	int batch, length, c, cp, in_h, in_w, out_h, out_w, p_h, p_w, s_h, s_w, pad_h, pad_w, type;
	cb200_activ activ;
};

template <typename T>
__global__ void __launch_bounds__(256)
pool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, uint8_t* __restrict__ map, PoolGeom g) {
	// grid: (x: output columns x channel vectors, y: output row, z: sample) - 32-bit index math only
	const unsigned cv = (unsigned)g.cp >> 3;
	const bool mask_tail = activ_masks_tail(g.activ);
	const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx < (unsigned)g.out_w * cv) {
		const int v = (int)(idx % cv);
		const int ox = (int)(idx / cv);
		const int oy = blockIdx.y;
		const int b = blockIdx.z;
		float best[8];
		int arg[8];
		int count = 0;
#pragma unroll
		for (int j = 0; j < 8; j++) { best[j] = 0.0f; arg[j] = 255; }
		// window scan order y then x (z,y,x upstream), first strict maximum wins
		for (int py = 0; py < g.p_h; py++) {
			const int iy = oy * g.s_h + py - g.pad_h;
			if (iy < 0 || iy >= g.in_h) continue;
			for (int px = 0; px < g.p_w; px++) {
				const int ix = ox * g.s_w + px - g.pad_w;
				if (ix < 0 || ix >= g.in_w) continue;
				float val[8];
				load8<T>(x + (((long long)b * g.in_h + iy) * g.in_w + ix) * g.cp + v * 8, val);
				if (g.type == CB200_POOL_MAX) {
					const int loc = py * g.p_w + px;
					if (count == 0) {
#pragma unroll
						for (int j = 0; j < 8; j++) { best[j] = val[j]; arg[j] = loc; }
					} else {
#pragma unroll
						for (int j = 0; j < 8; j++) if (val[j] > best[j]) { best[j] = val[j]; arg[j] = loc; }
					}
				} else {
#pragma unroll
					for (int j = 0; j < 8; j++) best[j] += val[j];
				}
				count++;
			}
		}
		if (g.type == CB200_POOL_AVG) {
			const float inv = 1.0f / (float)count;   // count == 0 gives inf/nan exactly like upstream's r_avg/sum_elem
#pragma unroll
			for (int j = 0; j < 8; j++) best[j] *= inv;
		}
		const bool dead = mask_tail && b >= g.length;
#pragma unroll
		for (int j = 0; j < 8; j++) {
			const bool real = (v * 8 + j) < g.c;
			best[j] = (real && !dead) ? activ_forward(g.activ, best[j]) : 0.0f;
			if (!real) arg[j] = 255;
		}
		const long long o = (((long long)b * g.out_h + oy) * g.out_w + ox) * g.cp + v * 8;
		store8<T>(y + o, best);
		if (map != nullptr && g.type == CB200_POOL_MAX) {
			uint2 packed;
			packed.x = (uint32_t)arg[0] | ((uint32_t)arg[1] << 8) | ((uint32_t)arg[2] << 16) | ((uint32_t)arg[3] << 24);
			packed.y = (uint32_t)arg[4] | ((uint32_t)arg[5] << 8) | ((uint32_t)arg[6] << 16) | ((uint32_t)arg[7] << 24);
			*reinterpret_cast<uint2*>(map + o) = packed;
		}
	}
}

// gather form of the backward pass: each input pixel sums the deltas of the windows that selected it
// (deltah_max_pool_cont / deltah_avg_pool_cont upstream), then the previous layer's deriv hook.
template <typename T>
__global__ void __launch_bounds__(256)
pool_bwd_kernel(const T* __restrict__ dy, const uint8_t* __restrict__ map, T* __restrict__ dx,
                const T* __restrict__ prev_out, cb200_activ prev_activ, PoolGeom g) {
	const unsigned cv = (unsigned)g.cp >> 3;
	const bool hook = prev_out != nullptr && prev_activ.type != CB200_LINEAR;
	const bool mask_tail = hook && activ_masks_tail(prev_activ);
	const float inv_vol = 1.0f / (float)(g.p_h * g.p_w);
	const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx < (unsigned)g.in_w * cv) {
		const int v = (int)(idx % cv);
		const int ix = (int)(idx / cv);
		const int iy = blockIdx.y;
		const int b = blockIdx.z;
		float acc[8];
#pragma unroll
		for (int j = 0; j < 8; j++) acc[j] = 0.0f;
		const int py_pos = iy + g.pad_h, px_pos = ix + g.pad_w;
		for (int oy = py_pos / g.s_h; oy >= 0 && py_pos - oy * g.s_h < g.p_h; oy--) {
			if (oy >= g.out_h) continue;
			const int fy = py_pos - oy * g.s_h;
			for (int ox = px_pos / g.s_w; ox >= 0 && px_pos - ox * g.s_w < g.p_w; ox--) {
				if (ox >= g.out_w) continue;
				const int fx = px_pos - ox * g.s_w;
				const long long o = (((long long)b * g.out_h + oy) * g.out_w + ox) * g.cp + v * 8;
				float d[8];
				load8<T>(dy + o, d);
				if (g.type == CB200_POOL_MAX) {
					const uint2 packed = *reinterpret_cast<const uint2*>(map + o);
					const int loc = fy * g.p_w + fx;
#pragma unroll
					for (int j = 0; j < 8; j++) {
						const uint32_t m = ((j < 4 ? packed.x : packed.y) >> (8 * (j & 3))) & 0xffu;
						if ((int)m == loc) acc[j] += d[j];
					}
				} else {
#pragma unroll
					for (int j = 0; j < 8; j++) acc[j] += d[j] * inv_vol;
				}
			}
		}
		const long long o_in = (((long long)b * g.in_h + iy) * g.in_w + ix) * g.cp + v * 8;
		if (hook) {
			float pv[8];
			load8<T>(prev_out + o_in, pv);
			const bool dead = mask_tail && b >= g.length;
#pragma unroll
			for (int j = 0; j < 8; j++) acc[j] = dead ? 0.0f : activ_deriv_mul(prev_activ, acc[j], pv[j]);
		}
		store8<T>(dx + o_in, acc);
	}
}

// fast path of the backward pass for non-overlapping windows (stride == size, no padding: every pooling layer of the
// BASELINE configs): an input pixel belongs to exactly one window, four pixels are kept in flight per thread
template <typename T>
__global__ void __launch_bounds__(256)
pool_bwd_disjoint_kernel(const T* __restrict__ dy, const uint8_t* __restrict__ map, T* __restrict__ dx,
                         const T* __restrict__ prev_out, cb200_activ prev_activ, PoolGeom g) {
	const unsigned cv = (unsigned)g.cp >> 3;
	const bool hook = prev_out != nullptr && prev_activ.type != CB200_LINEAR;
	const bool mask_tail = hook && activ_masks_tail(prev_activ);
	const float inv_vol = 1.0f / (float)(g.p_h * g.p_w);
	const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= (unsigned)g.in_w * cv) return;
	const int v = (int)(idx % cv);
	const int ix = (int)(idx / cv);
	const int iy = blockIdx.y, b = blockIdx.z;
	const int oy = iy / g.p_h, ox = ix / g.p_w;
	const int loc = (iy - oy * g.p_h) * g.p_w + (ix - ox * g.p_w);
	const long long o_in = (((long long)b * g.in_h + iy) * g.in_w + ix) * g.cp + v * 8;
	float acc[8];
#pragma unroll
	for (int j = 0; j < 8; j++) acc[j] = 0.0f;
	Raw8<T> rp;
	if (hook) rp = load_raw8<T>(prev_out + o_in);
	if (oy < g.out_h && ox < g.out_w) {
		const long long o = (((long long)b * g.out_h + oy) * g.out_w + ox) * g.cp + v * 8;
		const Raw8<T> rd = load_raw8<T>(dy + o);
		float d[8];
		if (g.type == CB200_POOL_MAX) {
			const uint2 rm = __ldg(reinterpret_cast<const uint2*>(map + o));
			unpack8(rd, d);
#pragma unroll
			for (int j = 0; j < 8; j++) {
				const uint32_t m = ((j < 4 ? rm.x : rm.y) >> (8 * (j & 3))) & 0xffu;
				if ((int)m == loc) acc[j] = d[j];
			}
		} else {
			unpack8(rd, d);
#pragma unroll
			for (int j = 0; j < 8; j++) acc[j] = d[j] * inv_vol;
		}
	}
	if (hook) {
		float pv[8];
		unpack8(rp, pv);
		const bool dead = mask_tail && b >= g.length;
#pragma unroll
		for (int j = 0; j < 8; j++) acc[j] = dead ? 0.0f : activ_deriv_mul(prev_activ, acc[j], pv[j]);
	}
	store8<T>(dx + o_in, acc);
}

// global average pooling (window == whole map): one block per sample, threads = channel vectors x pixel lanes,
// shared-memory reduction over the pixel lanes.  (Darknet19 head: 14x14x1000 -> 1000 per image.)
template <typename T>
__global__ void __launch_bounds__(256)
pool_global_avg_kernel(const T* __restrict__ x, T* __restrict__ y, PoolGeom g) {
	extern __shared__ float red[];                 // [lanes_p][lanes_c * 8]
	const int cv = g.cp >> 3, hw = g.in_h * g.in_w;
	const int b = blockIdx.x;
	const int lanes_c = cv < 256 ? cv : 256;
	const int lanes_p = 256 / lanes_c;
	const int lane_c = threadIdx.x % lanes_c, lane_p = threadIdx.x / lanes_c;
	const bool mask_tail = activ_masks_tail(g.activ);
	const bool dead = mask_tail && b >= g.length;
	constexpr int U = 4;
	for (int v0 = 0; v0 < cv; v0 += lanes_c) {
		const int v = v0 + lane_c;
		float acc[8];
#pragma unroll
		for (int j = 0; j < 8; j++) acc[j] = 0.0f;
		if (lane_p < lanes_p && v < cv) {
			const T* base = x + (long long)b * hw * g.cp + v * 8;
			for (int p = lane_p; p < hw; p += lanes_p * U) {
				Raw8<T> raw[U];
#pragma unroll
				for (int u = 0; u < U; u++) { const int pp = p + u * lanes_p; if (pp < hw) raw[u] = load_raw8<T>(base + (long long)pp * g.cp); }
#pragma unroll
				for (int u = 0; u < U; u++) {
					if (p + u * lanes_p < hw) {
						float t[8];
						unpack8(raw[u], t);
#pragma unroll
						for (int j = 0; j < 8; j++) acc[j] += t[j];
					}
				}
			}
		}
		if (lane_p < lanes_p) {
#pragma unroll
			for (int j = 0; j < 8; j++) red[(lane_p * lanes_c + lane_c) * 8 + j] = acc[j];
		}
		__syncthreads();
		if (lane_p == 0 && v < cv) {
			float out[8];
			const float inv = 1.0f / (float)hw;
#pragma unroll
			for (int j = 0; j < 8; j++) {
				float t = 0.0f;
				for (int q = 0; q < lanes_p; q++) t += red[(q * lanes_c + lane_c) * 8 + j];
				const bool real = (v * 8 + j) < g.c;
				out[j] = (real && !dead) ? activ_forward(g.activ, t * inv) : 0.0f;
			}
			store8<T>(y + (long long)b * g.cp + v * 8, out);
		}
		__syncthreads();
	}
}

// ---------------------------------------------------------------- 3-D windows (depth > 1)
// Same conventions one dimension up: activations [B][D][H][W][Cp], window scan order z, y, x, map value
// (z*p_h + y)*p_w + x of the first strict maximum (max_pooling_kernel, src/cuda/cuda_pool_layer.cu:31-114), 255 for an
// empty window.  One thread per 8 channels of one output voxel (forward) / input voxel (backward, gather form).
struct Pool3Geom { int in_d, out_d, p_d, s_d, pad_d; };

template <typename T>
__global__ void __launch_bounds__(256)
pool3d_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, uint8_t* __restrict__ map, PoolGeom g, Pool3Geom g3) {
	const int cv = g.cp >> 3;
	const bool mask_tail = activ_masks_tail(g.activ);
	const long long total = (long long)g.batch * g3.out_d * g.out_h * g.out_w * cv;
	for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
		const int v = (int)(idx % cv);
		long long r = idx / cv;
		const int ox = (int)(r % g.out_w); r /= g.out_w;
		const int oy = (int)(r % g.out_h); r /= g.out_h;
		const int oz = (int)(r % g3.out_d);
		const int b = (int)(r / g3.out_d);
		float best[8];
		int arg[8], count = 0;
#pragma unroll
		for (int j = 0; j < 8; j++) { best[j] = 0.0f; arg[j] = 255; }
		for (int pz = 0; pz < g3.p_d; pz++) {
			const int iz = oz * g3.s_d + pz - g3.pad_d;
			if (iz < 0 || iz >= g3.in_d) continue;
			for (int py = 0; py < g.p_h; py++) {
				const int iy = oy * g.s_h + py - g.pad_h;
				if (iy < 0 || iy >= g.in_h) continue;
				for (int px = 0; px < g.p_w; px++) {
					const int ix = ox * g.s_w + px - g.pad_w;
					if (ix < 0 || ix >= g.in_w) continue;
					float val[8];
					load8<T>(x + ((((long long)b * g3.in_d + iz) * g.in_h + iy) * g.in_w + ix) * g.cp + v * 8, val);
					if (g.type == CB200_POOL_MAX) {
						const int loc = (pz * g.p_h + py) * g.p_w + px;
#pragma unroll
						for (int j = 0; j < 8; j++) if (count == 0 || val[j] > best[j]) { best[j] = val[j]; arg[j] = loc; }
					} else {
#pragma unroll
						for (int j = 0; j < 8; j++) best[j] += val[j];
					}
					count++;
				}
			}
		}
		if (g.type == CB200_POOL_AVG) {
			const float inv = 1.0f / (float)count;
#pragma unroll
			for (int j = 0; j < 8; j++) best[j] *= inv;
		}
		const bool dead = mask_tail && b >= g.length;
#pragma unroll
		for (int j = 0; j < 8; j++) {
			const bool real = (v * 8 + j) < g.c;
			best[j] = (real && !dead) ? activ_forward(g.activ, best[j]) : 0.0f;
			if (!real) arg[j] = 255;
		}
		const long long o = ((((long long)b * g3.out_d + oz) * g.out_h + oy) * g.out_w + ox) * g.cp + v * 8;
		store8<T>(y + o, best);
		if (map != nullptr && g.type == CB200_POOL_MAX) {
			uint2 packed;
			packed.x = (uint32_t)arg[0] | ((uint32_t)arg[1] << 8) | ((uint32_t)arg[2] << 16) | ((uint32_t)arg[3] << 24);
			packed.y = (uint32_t)arg[4] | ((uint32_t)arg[5] << 8) | ((uint32_t)arg[6] << 16) | ((uint32_t)arg[7] << 24);
			*reinterpret_cast<uint2*>(map + o) = packed;
		}
	}
}

template <typename T>
__global__ void __launch_bounds__(256)
pool3d_bwd_kernel(const T* __restrict__ dy, const uint8_t* __restrict__ map, T* __restrict__ dx,
                  const T* __restrict__ prev_out, cb200_activ prev_activ, PoolGeom g, Pool3Geom g3) {
	const int cv = g.cp >> 3;
	const bool hook = prev_out != nullptr && prev_activ.type != CB200_LINEAR;
	const bool mask_tail = hook && activ_masks_tail(prev_activ);
	const float inv_vol = 1.0f / (float)(g3.p_d * g.p_h * g.p_w);
	const long long total = (long long)g.batch * g3.in_d * g.in_h * g.in_w * cv;
	for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
		const int v = (int)(idx % cv);
		long long r = idx / cv;
		const int ix = (int)(r % g.in_w); r /= g.in_w;
		const int iy = (int)(r % g.in_h); r /= g.in_h;
		const int iz = (int)(r % g3.in_d);
		const int b = (int)(r / g3.in_d);
		float acc[8];
#pragma unroll
		for (int j = 0; j < 8; j++) acc[j] = 0.0f;
		const int pz_pos = iz + g3.pad_d, py_pos = iy + g.pad_h, px_pos = ix + g.pad_w;
		for (int oz = pz_pos / g3.s_d; oz >= 0 && pz_pos - oz * g3.s_d < g3.p_d; oz--) {
			if (oz >= g3.out_d) continue;
			const int fz = pz_pos - oz * g3.s_d;
			for (int oy = py_pos / g.s_h; oy >= 0 && py_pos - oy * g.s_h < g.p_h; oy--) {
				if (oy >= g.out_h) continue;
				const int fy = py_pos - oy * g.s_h;
				for (int ox = px_pos / g.s_w; ox >= 0 && px_pos - ox * g.s_w < g.p_w; ox--) {
					if (ox >= g.out_w) continue;
					const int fx = px_pos - ox * g.s_w;
					const long long o = ((((long long)b * g3.out_d + oz) * g.out_h + oy) * g.out_w + ox) * g.cp + v * 8;
					float d[8];
					load8<T>(dy + o, d);
					if (g.type == CB200_POOL_MAX) {
						const uint2 packed = *reinterpret_cast<const uint2*>(map + o);
						const int loc = (fz * g.p_h + fy) * g.p_w + fx;
#pragma unroll
						for (int j = 0; j < 8; j++) {
							const uint32_t m = ((j < 4 ? packed.x : packed.y) >> (8 * (j & 3))) & 0xffu;
							if ((int)m == loc) acc[j] += d[j];
						}
					} else {
#pragma unroll
						for (int j = 0; j < 8; j++) acc[j] += d[j] * inv_vol;
					}
				}
			}
		}
		const long long o_in = ((((long long)b * g3.in_d + iz) * g.in_h + iy) * g.in_w + ix) * g.cp + v * 8;
		if (hook) {
			float pv[8];
			load8<T>(prev_out + o_in, pv);
			const bool dead = mask_tail && b >= g.length;
#pragma unroll
			for (int j = 0; j < 8; j++) acc[j] = dead ? 0.0f : activ_deriv_mul(prev_activ, acc[j], pv[j]);
		}
		store8<T>(dx + o_in, acc);
	}
}

static bool pool_is_3d(const cb200_pool_desc* d) { return d->in_d > 1 || d->out_d > 1 || d->p_d > 1 || d->stride_d > 1 || d->pad_d > 0; }
static Pool3Geom fill_geom3(const cb200_pool_desc* d) {
	Pool3Geom g3;
	g3.in_d = d->in_d > 0 ? d->in_d : 1; g3.out_d = d->out_d > 0 ? d->out_d : 1; g3.p_d = d->p_d > 0 ? d->p_d : 1;
	g3.s_d = d->stride_d > 0 ? d->stride_d : 1; g3.pad_d = d->pad_d;
	return g3;
}

static int fill_geom(const cb200_pool_desc* d, PoolGeom& g) {
	CB_ARG(d != nullptr && d->batch > 0 && d->c > 0);
	CB_ARG(d->p_h > 0 && d->p_w > 0 && d->stride_h > 0 && d->stride_w > 0);
	CB_ARG((d->p_d > 0 ? d->p_d : 1) * d->p_h * d->p_w < 255);
	g.batch = d->batch; g.length = d->length; g.c = d->c; g.cp = round8(d->c);
	g.in_h = d->in_h; g.in_w = d->in_w; g.out_h = d->out_h; g.out_w = d->out_w;
	g.p_h = d->p_h; g.p_w = d->p_w; g.s_h = d->stride_h; g.s_w = d->stride_w; g.pad_h = d->pad_h; g.pad_w = d->pad_w;
	g.type = d->pool_type; g.activ = d->activ;
	return CB200_OK;
}
}  // namespace cb200
using namespace cb200;

extern "C" {

int cb200_pool_forward(const cb200_pool_desc* d, const void* x, void* y, uint8_t* map, void* s) {
	CB_REQUIRE_DEVICE();
	PoolGeom g;
	int rc = fill_geom(d, g); if (rc) return rc;
	if (pool_is_3d(d)) {
		const Pool3Geom g3 = fill_geom3(d);
		const long long total = (long long)g.batch * g3.out_d * g.out_h * g.out_w * (g.cp >> 3);
		CB_DISPATCH_DTYPE(d->dtype, T, (pool3d_fwd_kernel<T><<<grid_for(total, 256), 256, 0, as_stream(s)>>>((const T*)x, (T*)y, map, g, g3)));
		CB_LAUNCH_CHECK();
		return CB200_OK;
	}
	dim3 grid((unsigned)ceil_div(g.out_w * (g.cp >> 3), 256), (unsigned)g.out_h, (unsigned)g.batch);
	const double es = (double)cb200_dtype_size(d->dtype);
	// algorithmic bytes: read the input once, write output + 1-byte argmax
	prof_begin(PROF_POOL, (double)g.batch * g.c * ((double)g.in_h * g.in_w * es + (double)g.out_h * g.out_w * (es + 1)), as_stream(s));
	const bool global_avg = g.type == CB200_POOL_AVG && g.out_h == 1 && g.out_w == 1 && g.p_h == g.in_h && g.p_w == g.in_w &&
	                        g.pad_h == 0 && g.pad_w == 0 && g.in_h * g.in_w >= 16;
	if (global_avg) {
		CB_DISPATCH_DTYPE(d->dtype, T, (pool_global_avg_kernel<T><<<g.batch, 256, 256 * 8 * sizeof(float), as_stream(s)>>>((const T*)x, (T*)y, g)));
	} else {
		CB_DISPATCH_DTYPE(d->dtype, T, (pool_fwd_kernel<T><<<grid, 256, 0, as_stream(s)>>>((const T*)x, (T*)y, map, g)));
	}
	CB_LAUNCH_CHECK();
	prof_end(as_stream(s));
	return CB200_OK;
}

int cb200_pool_backward(const cb200_pool_desc* d, const void* dy, const uint8_t* map, void* dx,
                        const cb200_activ* prev_activ, const void* prev_out, void* s) {
	CB_REQUIRE_DEVICE();
	PoolGeom g;
	int rc = fill_geom(d, g); if (rc) return rc;
	CB_ARG(d->pool_type != CB200_POOL_MAX || map != nullptr);
	cb200_activ pa; pa.type = CB200_LINEAR; pa.leak = 0; pa.saturation = 0; pa.beta = 0;
	if (prev_activ) pa = *prev_activ;
	if (pool_is_3d(d)) {
		const Pool3Geom g3 = fill_geom3(d);
		const long long total = (long long)g.batch * g3.in_d * g.in_h * g.in_w * (g.cp >> 3);
		CB_DISPATCH_DTYPE(d->dtype, T, (pool3d_bwd_kernel<T><<<grid_for(total, 256), 256, 0, as_stream(s)>>>((const T*)dy, map, (T*)dx, (const T*)prev_out, pa, g, g3)));
		CB_LAUNCH_CHECK();
		return CB200_OK;
	}
	dim3 grid((unsigned)ceil_div(g.in_w * (g.cp >> 3), 256), (unsigned)g.in_h, (unsigned)g.batch);
	const double es = (double)cb200_dtype_size(d->dtype);
	prof_begin(PROF_POOL, (double)g.batch * g.c * ((double)g.in_h * g.in_w * es + (double)g.out_h * g.out_w * (es + 1)), as_stream(s));
	const bool disjoint = g.s_h == g.p_h && g.s_w == g.p_w && g.pad_h == 0 && g.pad_w == 0;
	if (disjoint) {
		CB_DISPATCH_DTYPE(d->dtype, T, (pool_bwd_disjoint_kernel<T><<<grid, 256, 0, as_stream(s)>>>((const T*)dy, map, (T*)dx, (const T*)prev_out, pa, g)));
	} else {
		CB_DISPATCH_DTYPE(d->dtype, T, (pool_bwd_kernel<T><<<grid, 256, 0, as_stream(s)>>>((const T*)dy, map, (T*)dx, (const T*)prev_out, pa, g)));
	}
	CB_LAUNCH_CHECK();
	prof_end(as_stream(s));
	return CB200_OK;
}

}  // extern "C"
