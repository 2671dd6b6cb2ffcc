// conv_simt.cu - generic implicit-GEMM convolution on the CUDA cores (FP32 accumulate).
//
// Serves (a) the FP32C_FP32A mode, whose 1e-5 tolerance rules out 16-bit tensor-core inputs,
// (b) every geometry the tcgen05 kernels do not take (stride > 1, channel counts < 16, ...),
// (c) the on-GPU cross-check of the tcgen05 path (cb200_force_simt).
// No im2col buffer is ever materialised: the three GEMMs of a conv layer
// (reference: src/cuda/cuda_conv_layer.cu:391-397 fwd, :521-527 dgrad, :551-557 wgrad) gather their
// operands straight from the channels-last activations with the reference's address map
// (im2col_kernel, cuda_conv_layer.cu:36-103) folded into the loaders below.
#include "common.cuh"

namespace cb200 {

constexpr int TM = 64, TN = 64, TK = 16, NTHREADS = 256;

struct ConvGeom {
	int batch, length;
	int in_c, in_cp, in_d, in_h, in_w;
	int out_c, out_cp, out_d, out_h, out_w;
	int f_d, f_h, f_w, s_d, s_h, s_w, p_d, p_h, p_w;
	int q_d, q_h, q_w;           // 1 + internal padding: input pixel i sits at i*q of the zero-stuffed grid the filter slides over
	int wb_dense;                // row order of w_bwd (conv_whole_map)
	float bias_value;
	cb200_activ activ;
};

static ConvGeom make_geom(const cb200_conv_desc* d) {
	ConvGeom g;
	g.batch = d->batch; g.length = d->length;
	g.in_c = d->in_c; g.in_cp = round8(d->in_c); g.in_h = d->in_h; g.in_w = d->in_w;
	g.out_c = d->out_c; g.out_cp = round8(d->out_c); g.out_h = d->out_h; g.out_w = d->out_w;
	g.f_h = d->f_h; g.f_w = d->f_w; g.s_h = d->stride_h; g.s_w = d->stride_w; g.p_h = d->pad_h; g.p_w = d->pad_w;
	g.in_d = conv_in_d(d); g.out_d = conv_out_d(d); g.f_d = conv_f_d(d); g.s_d = d->stride_d > 0 ? d->stride_d : 1; g.p_d = d->pad_d;
	g.q_d = 1 + d->ipad_d; g.q_h = 1 + d->ipad_h; g.q_w = 1 + d->ipad_w;
	g.bias_value = d->bias_value; g.activ = d->activ;
	g.wb_dense = conv_whole_map(d) ? 1 : 0;
	return g;
}

// The reference's address map (im2col_kernel, src/cuda/cuda_conv_layer.cu:36-103) in gather form, all three dimensions
// and internal padding included: output position o, filter tap k -> position v = o*stride - pad + k on the zero-stuffed
// grid; it holds input pixel v / q when v >= 0, v % q == 0 and v / q < size, a zero otherwise.
__device__ __forceinline__ bool tap_src(int o, int k, int stride, int pad, int q, int size, int& i) {
	const int v = o * stride - pad + k;
	if (v < 0) return false;
	if (q == 1) { i = v; return v < size; }
	if (v % q != 0) return false;
	i = v / q;
	return i < size;
}
// the transposed map of the data gradient: input pixel i, tap k -> the output position o with o*stride - pad + k == i*q
__device__ __forceinline__ bool tap_dst(int i, int k, int stride, int pad, int q, int size, int& o) {
	const int n = i * q + pad - k;
	if (n < 0 || n % stride != 0) return false;
	o = n / stride;
	return o < size;
}
__device__ __forceinline__ void split_pixel(long long m, int d, int h, int w, int& b, int& z, int& y, int& x) {
	x = (int)(m % w); long long r = m / w;
	y = (int)(r % h); r /= h;
	z = (int)(r % d); b = (int)(r / d);
}

// ---------------------------------------------------------------- forward
// C[m][n] = sum_k A[m][k] B[k][n]; m = output pixel (b,oy,ox), n = filter, k = (tap, c)
template <typename T>
__global__ void __launch_bounds__(NTHREADS)
conv_fwd_simt_kernel(const T* __restrict__ x, const T* __restrict__ wf, const float* __restrict__ bias_w,
                     T* __restrict__ y, ConvGeom g) {
	__shared__ float As[TK][TM + 4];
	__shared__ float Bs[TK][TN + 4];
	const int tid = threadIdx.x;
	const long long M = (long long)g.batch * g.out_d * g.out_h * g.out_w;
	const int K = g.f_d * g.f_h * g.f_w * g.in_cp;
	const long long m0 = (long long)blockIdx.x * TM;
	const int n0 = blockIdx.y * TN;

	// loader roles: each thread owns one row (m or n) and 4 consecutive k
	const int lrow = tid & 63, lk = (tid >> 6) * 4;
	const long long lm = m0 + lrow;
	int lb = 0, loz = 0, loy = 0, lox = 0;
	const bool lm_ok = lm < M;
	if (lm_ok) split_pixel(lm, g.out_d, g.out_h, g.out_w, lb, loz, loy, lox);
	const int ln = n0 + lrow;
	const bool ln_ok = ln < g.out_c;

	const int tx = tid & 15, ty = tid >> 4;   // 16 x 16 threads, 4x4 outputs each
	float acc[4][4] = {};

	for (int k0 = 0; k0 < K; k0 += TK) {
		const int k = k0 + lk;                       // 4 consecutive k share a tap (in_cp % 8 == 0)
		float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
		if (k < K) {
			const int tap = k / g.in_cp, c = k - tap * g.in_cp;
			if (lm_ok) {
				const int kz = tap / (g.f_h * g.f_w), kyx = tap - kz * g.f_h * g.f_w;
				const int ky = kyx / g.f_w, kx = kyx - ky * g.f_w;
				int iz, iy, ix;
				if (tap_src(loz, kz, g.s_d, g.p_d, g.q_d, g.in_d, iz) && tap_src(loy, ky, g.s_h, g.p_h, g.q_h, g.in_h, iy)
				    && tap_src(lox, kx, g.s_w, g.p_w, g.q_w, g.in_w, ix)) {
					const T* p = x + ((((long long)lb * g.in_d + iz) * g.in_h + iy) * g.in_w + ix) * g.in_cp + c;
#pragma unroll
					for (int i = 0; i < 4; i++) av[i] = to_f32<T>(p[i]);
				}
			}
			if (ln_ok) {
				const T* p = wf + (long long)ln * K + k;
#pragma unroll
				for (int i = 0; i < 4; i++) bv[i] = to_f32<T>(p[i]);
			}
		}
#pragma unroll
		for (int i = 0; i < 4; i++) { As[lk + i][lrow] = av[i]; Bs[lk + i][lrow] = bv[i]; }
		__syncthreads();
#pragma unroll
		for (int kk = 0; kk < TK; kk++) {
			float a[4], b[4];
#pragma unroll
			for (int i = 0; i < 4; i++) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
			for (int i = 0; i < 4; i++)
#pragma unroll
				for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
		}
		__syncthreads();
	}

	const bool mask_tail = activ_masks_tail(g.activ);
#pragma unroll
	for (int i = 0; i < 4; i++) {
		const long long m = m0 + ty * 4 + i;
		if (m >= M) continue;
		const int b = (int)(m / ((long long)g.out_d * g.out_h * g.out_w));
		const bool dead = mask_tail && b >= g.length;
#pragma unroll
		for (int j = 0; j < 4; j++) {
			const int n = n0 + tx * 4 + j;
			if (n >= g.out_cp) continue;
			float v = 0.0f;
			if (n < g.out_c && !dead) v = activ_forward(g.activ, acc[i][j] + g.bias_value * bias_w[n]);
			y[m * g.out_cp + n] = from_f32<T>(v);
		}
	}
}

// ---------------------------------------------------------------- data gradient
// dx[m][c] = sum_{tap',f} dy[shift(m,tap')][f] * w_bwd[c][tap'][f];  m = input pixel
template <typename T>
__global__ void __launch_bounds__(NTHREADS)
conv_dgrad_simt_kernel(const T* __restrict__ dy, const T* __restrict__ wb, T* __restrict__ dx,
                       const T* __restrict__ prev_out, cb200_activ prev_activ, ConvGeom g) {
	__shared__ float As[TK][TM + 4];
	__shared__ float Bs[TK][TN + 4];
	const int tid = threadIdx.x;
	const long long M = (long long)g.batch * g.in_d * g.in_h * g.in_w;
	const int taps = g.f_d * g.f_h * g.f_w;
	const int K = taps * g.out_cp;
	const long long m0 = (long long)blockIdx.x * TM;
	const int n0 = blockIdx.y * TN;

	const int lrow = tid & 63, lk = (tid >> 6) * 4;
	const long long lm = m0 + lrow;
	int lb = 0, liz = 0, liy = 0, lix = 0;
	const bool lm_ok = lm < M;
	if (lm_ok) split_pixel(lm, g.in_d, g.in_h, g.in_w, lb, liz, liy, lix);
	const int ln = n0 + lrow;
	const bool ln_ok = ln < g.in_c;

	const int tx = tid & 15, ty = tid >> 4;
	float acc[4][4] = {};

	for (int k0 = 0; k0 < K; k0 += TK) {
		const int k = k0 + lk;
		float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
		if (k < K) {
			const int tapr = k / g.out_cp, f = k - tapr * g.out_cp;
			if (lm_ok) {
				// rotated tap index <-> original tap: tap = taps - 1 - tap' (reversed in every dimension)
				const int tap = taps - 1 - tapr;
				const int kz = tap / (g.f_h * g.f_w), kyx = tap - kz * g.f_h * g.f_w;
				const int ky = kyx / g.f_w, kx = kyx - ky * g.f_w;
				int oz, oy, ox;
				if (tap_dst(liz, kz, g.s_d, g.p_d, g.q_d, g.out_d, oz) && tap_dst(liy, ky, g.s_h, g.p_h, g.q_h, g.out_h, oy)
				    && tap_dst(lix, kx, g.s_w, g.p_w, g.q_w, g.out_w, ox)) {
					const T* p = dy + ((((long long)lb * g.out_d + oz) * g.out_h + oy) * g.out_w + ox) * g.out_cp + f;
#pragma unroll
					for (int i = 0; i < 4; i++) av[i] = to_f32<T>(p[i]);
				}
			}
			if (ln_ok) {
				// rows (c, rotated tap), or (tap, c) for a whole-map filter (conv.cu: wbwd_row)
				const T* p = g.wb_dense ? wb + ((long long)(taps - 1 - tapr) * g.in_cp + ln) * g.out_cp + f
				                        : wb + (long long)ln * K + k;
#pragma unroll
				for (int i = 0; i < 4; i++) bv[i] = to_f32<T>(p[i]);
			}
		}
#pragma unroll
		for (int i = 0; i < 4; i++) { As[lk + i][lrow] = av[i]; Bs[lk + i][lrow] = bv[i]; }
		__syncthreads();
#pragma unroll
		for (int kk = 0; kk < TK; kk++) {
			float a[4], b[4];
#pragma unroll
			for (int i = 0; i < 4; i++) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
			for (int i = 0; i < 4; i++)
#pragma unroll
				for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
		}
		__syncthreads();
	}

	const bool hook = prev_out != nullptr && prev_activ.type != CB200_LINEAR;
	const bool mask_tail = hook && activ_masks_tail(prev_activ);
#pragma unroll
	for (int i = 0; i < 4; i++) {
		const long long m = m0 + ty * 4 + i;
		if (m >= M) continue;
		const int b = (int)(m / ((long long)g.in_d * g.in_h * g.in_w));
		const bool dead = mask_tail && b >= g.length;
#pragma unroll
		for (int j = 0; j < 4; j++) {
			const int n = n0 + tx * 4 + j;
			if (n >= g.in_cp) continue;
			float v = 0.0f;
			if (n < g.in_c && !dead) {
				v = acc[i][j];
				if (hook) v = activ_deriv_mul(prev_activ, v, to_f32<T>(prev_out[m * g.in_cp + n]));
			}
			dx[m * g.in_cp + n] = from_f32<T>(v);
		}
	}
}

// ---------------------------------------------------------------- weight gradient
// grad[f][tap][c] += sum_pix dy[pix][f] * x[shift(pix,tap)][c]   (split over pix chunks, FP32 atomics)
template <typename T>
__global__ void __launch_bounds__(NTHREADS)
conv_wgrad_simt_kernel(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ grad,
                       ConvGeom g, long long pix_per_split) {
	__shared__ float As[TK][TM + 4];   // [pix][f]
	__shared__ float Bs[TK][TN + 4];   // [pix][(tap,c)]
	const int tid = threadIdx.x;
	const long long P = (long long)g.batch * g.out_d * g.out_h * g.out_w;
	const int NN = g.f_d * g.f_h * g.f_w * g.in_cp;
	const int f0 = blockIdx.x * TM;
	const int n0 = blockIdx.y * TN;
	const long long p_begin = (long long)blockIdx.z * pix_per_split;
	long long p_end = p_begin + pix_per_split;
	if (p_end > P) p_end = P;

	const int lrow = tid & 63, lk = (tid >> 6) * 4;
	const int lf = f0 + lrow;
	const bool lf_ok = lf < g.out_c;
	const int ln = n0 + lrow;
	const bool ln_ok = ln < NN;
	int ltap = 0, lc = 0, lkz = 0, lky = 0, lkx = 0;
	if (ln_ok) {
		ltap = ln / g.in_cp; lc = ln - ltap * g.in_cp;
		lkz = ltap / (g.f_h * g.f_w);
		const int kyx = ltap - lkz * g.f_h * g.f_w;
		lky = kyx / g.f_w; lkx = kyx - lky * g.f_w;
	}

	const int tx = tid & 15, ty = tid >> 4;
	float acc[4][4] = {};

	for (long long q0 = p_begin; q0 < p_end; q0 += TK) {
#pragma unroll
		for (int i = 0; i < 4; i++) {
			const long long pix = q0 + lk + i;
			float a = 0.0f, b = 0.0f;
			if (pix < p_end) {
				if (lf_ok) a = to_f32<T>(dy[pix * g.out_cp + lf]);
				if (ln_ok) {
					int bb, oz, oy, ox, iz, iy, ix;
					split_pixel(pix, g.out_d, g.out_h, g.out_w, bb, oz, oy, ox);
					if (tap_src(oz, lkz, g.s_d, g.p_d, g.q_d, g.in_d, iz) && tap_src(oy, lky, g.s_h, g.p_h, g.q_h, g.in_h, iy)
					    && tap_src(ox, lkx, g.s_w, g.p_w, g.q_w, g.in_w, ix))
						b = to_f32<T>(x[((((long long)bb * g.in_d + iz) * g.in_h + iy) * g.in_w + ix) * g.in_cp + lc]);
				}
			}
			As[lk + i][lrow] = a;
			Bs[lk + i][lrow] = b;
		}
		__syncthreads();
#pragma unroll
		for (int kk = 0; kk < TK; kk++) {
			float a[4], b[4];
#pragma unroll
			for (int i = 0; i < 4; i++) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
			for (int i = 0; i < 4; i++)
#pragma unroll
				for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
		}
		__syncthreads();
	}
#pragma unroll
	for (int i = 0; i < 4; i++) {
		const int f = f0 + ty * 4 + i;
		if (f >= g.out_c) continue;
#pragma unroll
		for (int j = 0; j < 4; j++) {
			const int n = n0 + tx * 4 + j;
			if (n >= NN) continue;
			if (gridDim.z == 1) grad[(long long)f * NN + n] = acc[i][j];
			else atomicAdd(&grad[(long long)f * NN + n], acc[i][j]);
		}
	}
}

// grad_b[f] = sum over all pixels of dy[pix][f] : column sums of a [P][Cp] matrix
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ dy, float* __restrict__ out, long long P, int c, int cp, long long rows_per_block) {
	// block = 256 threads: 32 channel lanes x 8 row lanes
	__shared__ float red[8][33];
	const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
	const int ch = blockIdx.x * 32 + cx;
	const long long r0 = (long long)blockIdx.y * rows_per_block;
	long long r1 = r0 + rows_per_block;
	if (r1 > P) r1 = P;
	float s = 0.0f;
	if (ch < c)
		for (long long r = r0 + ry; r < r1; r += 8) s += to_f32<T>(dy[r * cp + ch]);
	red[ry][cx] = s;
	__syncthreads();
	if (ry == 0 && ch < c) {
		float t = 0.0f;
#pragma unroll
		for (int i = 0; i < 8; i++) t += red[i][cx];
		atomicAdd(&out[ch], t);
	}
}

// ---------------------------------------------------------------- host launchers
int conv_forward_simt(const cb200_conv_desc* d, const cb200_conv_weights* w, const void* x, void* y, cudaStream_t st) {
	ConvGeom g = make_geom(d);
	long long M = (long long)g.batch * g.out_d * g.out_h * g.out_w;
	dim3 grid((unsigned)ceil_div_ll(M, TM), (unsigned)ceil_div(g.out_cp, TN));
	CB_DISPATCH_DTYPE(d->dtype, T, (conv_fwd_simt_kernel<T><<<grid, NTHREADS, 0, st>>>((const T*)x, (const T*)w->w_fwd, w->bias_w, (T*)y, g)));
	CB_LAUNCH_CHECK();
	return CB200_OK;
}

int conv_dgrad_simt(const cb200_conv_desc* d, const cb200_conv_weights* w, const void* dy, void* dx,
                    const cb200_activ* prev_activ, const void* prev_out, cudaStream_t st) {
	ConvGeom g = make_geom(d);
	long long M = (long long)g.batch * g.in_d * g.in_h * g.in_w;
	dim3 grid((unsigned)ceil_div_ll(M, TM), (unsigned)ceil_div(g.in_cp, TN));
	cb200_activ pa; pa.type = CB200_LINEAR; pa.leak = 0; pa.saturation = 0; pa.beta = 0;
	if (prev_activ) pa = *prev_activ;
	CB_DISPATCH_DTYPE(d->dtype, T, (conv_dgrad_simt_kernel<T><<<grid, NTHREADS, 0, st>>>((const T*)dy, (const T*)w->w_bwd, (T*)dx, (const T*)prev_out, pa, g)));
	CB_LAUNCH_CHECK();
	return CB200_OK;
}

int conv_colsum(int dtype, const void* dy, float* out, long long P, int c, cudaStream_t st) {
	int cp = round8(c);
	if (cudaMemsetAsync(out, 0, sizeof(float) * c, st) != cudaSuccess) { set_error("colsum memset failed"); return CB200_ERR_CUDA; }
	int cblocks = ceil_div(c, 32);
	long long want = (long long)g_num_sms * 8 / cblocks;
	if (want < 1) want = 1;
	long long rows_per_block = ceil_div_ll(P, want);
	if (rows_per_block < 64) rows_per_block = 64;
	dim3 grid((unsigned)cblocks, (unsigned)ceil_div_ll(P, rows_per_block));
	CB_DISPATCH_DTYPE(dtype, T, (colsum_kernel<T><<<grid, 256, 0, st>>>((const T*)dy, out, P, c, cp, rows_per_block)));
	CB_LAUNCH_CHECK();
	return CB200_OK;
}

int conv_wgrad_simt(const cb200_conv_desc* d, const cb200_conv_weights* w, const void* x, const void* dy, cudaStream_t st) {
	ConvGeom g = make_geom(d);
	long long P = (long long)g.batch * g.out_d * g.out_h * g.out_w;
	int NN = g.f_d * g.f_h * g.f_w * g.in_cp;
	int tiles = ceil_div(g.out_c, TM) * ceil_div(NN, TN);
	long long splits = ceil_div_ll((long long)g_num_sms * 4, tiles);
	long long max_splits = ceil_div_ll(P, 4 * TK);
	if (splits > max_splits) splits = max_splits;
	if (splits < 1) splits = 1;
	if (splits > 65535) splits = 65535;
	long long pps = ceil_div_ll(ceil_div_ll(P, splits), TK) * TK;
	splits = ceil_div_ll(P, pps);
	if (splits > 1) {
		if (cudaMemsetAsync(w->grad, 0, sizeof(float) * (size_t)g.out_c * NN, st) != cudaSuccess) { set_error("wgrad memset failed"); return CB200_ERR_CUDA; }
	}
	dim3 grid((unsigned)ceil_div(g.out_c, TM), (unsigned)ceil_div(NN, TN), (unsigned)splits);
	CB_DISPATCH_DTYPE(d->dtype, T, (conv_wgrad_simt_kernel<T><<<grid, NTHREADS, 0, st>>>((const T*)x, (const T*)dy, w->grad, g, pps)));
	CB_LAUNCH_CHECK();
	return CB200_OK;
}

}  // namespace cb200
