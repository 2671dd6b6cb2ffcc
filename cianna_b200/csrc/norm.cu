// norm.cu - group normalisation forward / backward / parameter update on channels-last tensors.
//
// Statistics are per (sample, group of `group_size` channels) exactly as upstream
// (src/cuda/cuda_norm_layer.cu:37-46 index map, :65-178 reductions, :180-265 apply kernels;
// CPU twin src/naiv/naiv_norm_layer.c).  Differences in mechanism, not in maths:
//   - one streaming pass produces sum and sum-of-squares (FP32 per thread over a short run,
//     FP64 across threads/blocks), instead of a mean pass followed by a variance pass;
//   - gamma/beta and their momentum live on the device; the update is a kernel, not a host
//     loop behind four blocking memcpys (cuda_norm_layer.cu:434-457).
#include <cstdlib>
#include "common.cuh"

namespace cb200 {

struct NormGeom {
	int batch, length, c, cp, hw, group_size, nb_group, set_off, ppb;
	float eps;
};

constexpr int NORM_THREADS = 256;
// pixels handled by one block: sized on the host so that every launch has several waves of blocks
// (deep layers have few pixels per sample) while a thread still streams a few 128-bit packets
// blocks wanted per SM over the batch and the most pixels a block takes: 8 / 4096 measured best on the Darknet19 shapes at
// batch 128 (profiles/r2_gn_apply_sweep.txt: 1024 pixels per block cost the 448 px layer 8-11 %, more blocks per SM 3-20 %)
static int g_norm_want_blocks = 8, g_norm_ppb_max = 4096;
static int norm_pix_per_block(int hw, int batch, int cv) {
	const int lanes_c = cv < NORM_THREADS ? cv : NORM_THREADS;
	const int lanes_p = NORM_THREADS / lanes_c;
	long long want_blocks = (long long)g_num_sms * g_norm_want_blocks;
	long long ppb = ((long long)hw * batch + want_blocks - 1) / want_blocks;
	const int min_ppb = lanes_p * 4;
	if (ppb < min_ppb) ppb = min_ppb;
	if (ppb > g_norm_ppb_max) ppb = g_norm_ppb_max;
	return (int)ppb;
}

// Per-thread sums of one channel vector -> the block's per-group accumulators in shared memory.
// Lanes of a warp that own the same channel vector (lanes_c < 32, a power of two) are summed by shuffles first, so
// that only lanes_c lanes per warp touch the shared accumulators; then the 8 channels are folded into their groups
// (a vector spans one group when group_size % 8 == 0).  Every lane of the warp must call this.
__device__ __forceinline__ void fold_into_groups(float (&s0)[8], float (&s1)[8], int v, int lanes_c, const NormGeom& g, float* sm_acc) {
	bool writer = true;
	if (lanes_c < 32 && (lanes_c & (lanes_c - 1)) == 0) {
#pragma unroll
		for (int j = 0; j < 8; j++) {
			for (int off = lanes_c; off < 32; off <<= 1) {
				s0[j] += __shfl_xor_sync(0xffffffffu, s0[j], off);
				s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], off);
			}
		}
		writer = (threadIdx.x & 31) < lanes_c;
	}
	if (writer) {
		int cur = -1;
		float g0 = 0.0f, g1 = 0.0f;
#pragma unroll
		for (int j = 0; j < 8; j++) {
			const int ch = v * 8 + j;
			if (ch >= g.c) break;
			const int grp = ch / g.group_size;
			if (grp != cur) {
				if (cur >= 0 && cur < g.nb_group) { atomicAdd(&sm_acc[cur * 2], g0); atomicAdd(&sm_acc[cur * 2 + 1], g1); }
				cur = grp; g0 = 0.0f; g1 = 0.0f;
			}
			g0 += s0[j];
			g1 += s1[j];
		}
		if (cur >= 0 && cur < g.nb_group) { atomicAdd(&sm_acc[cur * 2], g0); atomicAdd(&sm_acc[cur * 2 + 1], g1); }
	}
}

// Accumulate, for every channel vector owned by the thread, sum(a) and sum(a*b) over the block's
// pixel range; b == a gives (sum x, sum x^2), b == x with a == delta gives (sum d, sum d*x).
// ws layout: double [batch][nb_group][2].
// (vbx, b) = the block's pixel range and sample: blockIdx of the plain kernels, a virtual block of the pipelined ones
template <typename T, bool TWO_INPUTS>
__device__ __forceinline__ void norm_stats_body(const T* __restrict__ a_in, const T* __restrict__ b_in, double* __restrict__ ws,
                                                const NormGeom& g, int vbx, int b, float* sm_acc /* [nb_group][2] block-level partial sums */) {
	const int cv = g.cp >> 3;
	for (int i = threadIdx.x; i < g.nb_group * 2; i += blockDim.x) sm_acc[i] = 0.0f;
	__syncthreads();

	const int lanes_c = cv < NORM_THREADS ? cv : NORM_THREADS;
	const int lanes_p = NORM_THREADS / lanes_c;
	const int lane_c = threadIdx.x % lanes_c, lane_p = threadIdx.x / lanes_c;
	const int p0 = vbx * g.ppb;
	int p1 = p0 + g.ppb;
	if (p1 > g.hw) p1 = g.hw;

	{
		// (threads beyond lanes_c * lanes_p keep zero sums: every lane must reach the shuffles below)
		const bool active = lane_p < lanes_p;
		for (int v = lane_c; v < cv; v += lanes_c) {
			float s0[8], s1[8];
#pragma unroll
			for (int j = 0; j < 8; j++) { s0[j] = 0.0f; s1[j] = 0.0f; }
			constexpr int U = TWO_INPUTS ? 2 : 4;        // independent 128-bit loads in flight per thread: 4
			for (int p = p0 + lane_p; active && p < p1; p += lanes_p * U) {
				Raw8<T> ra[U], rb[U];
#pragma unroll
				for (int u = 0; u < U; u++) {
					const int pp = p + u * lanes_p;
					if (pp < p1) {
						const long long o = ((long long)b * g.hw + pp) * g.cp + v * 8;
						ra[u] = load_raw8<T>(a_in + o);
						if (TWO_INPUTS) rb[u] = load_raw8<T>(b_in + o);
					}
				}
#pragma unroll
				for (int u = 0; u < U; u++) {
					if (p + u * lanes_p < p1) {
						float av[8], bv[8];
						unpack8(ra[u], av);
						if (TWO_INPUTS) unpack8(rb[u], bv);
#pragma unroll
						for (int j = 0; j < 8; j++) {
							s0[j] += av[j];
							s1[j] += av[j] * (TWO_INPUTS ? bv[j] : av[j]);
						}
					}
				}
			}
			fold_into_groups(s0, s1, v, lanes_c, g, sm_acc);
		}
	}
	__syncthreads();
	for (int i = threadIdx.x; i < g.nb_group * 2; i += blockDim.x)
		atomicAdd(&ws[(size_t)b * g.nb_group * 2 + i], (double)sm_acc[i]);
}

template <typename T, bool TWO_INPUTS>
__global__ void __launch_bounds__(NORM_THREADS)
norm_stats_kernel(const T* __restrict__ a_in, const T* __restrict__ b_in, double* __restrict__ ws, NormGeom g) {
	extern __shared__ float sm_acc[];
	norm_stats_body<T, TWO_INPUTS>(a_in, b_in, ws, g, blockIdx.x, blockIdx.y, sm_acc);
}

// Apply kernels: same thread -> (channel vector, pixel lane) mapping as the statistics kernel, one sample per
// blockIdx.y.  A thread keeps ONE channel vector, so the per-channel affine constants are computed once outside
// the pixel loop and the loop body is: 128-bit load(s), 8 FMAs, 128-bit store - no integer division, several loads in flight.
template <typename T>
__device__ __forceinline__ void norm_apply_body(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                const float* mean, const float* var, const NormGeom& g, int vbx, int b) {
	const int cv = g.cp >> 3;
	const int lanes_c = cv < NORM_THREADS ? cv : NORM_THREADS;
	const int lanes_p = NORM_THREADS / lanes_c;
	const int lane_c = threadIdx.x % lanes_c, lane_p = threadIdx.x / lanes_c;
	if (lane_p >= lanes_p) return;
	const int p0 = vbx * g.ppb;
	int p1 = p0 + g.ppb;
	if (p1 > g.hw) p1 = g.hw;
	const bool dead = b >= g.length;
	constexpr int U = 4;
	for (int v = lane_c; v < cv; v += lanes_c) {
		float sc[8], sh[8];
#pragma unroll
		for (int j = 0; j < 8; j++) {
			const int ch = v * 8 + j;
			sc[j] = 0.0f; sh[j] = 0.0f;
			if (ch < g.c && !dead) {
				const int grp = ch / g.group_size;
				if (grp < g.nb_group - g.set_off) {
					const float rstd = 1.0f / sqrtf(var[b * g.nb_group + grp] + g.eps);
					sc[j] = gamma[grp] * rstd;
					sh[j] = beta[grp] - mean[b * g.nb_group + grp] * sc[j];
				} else sc[j] = 1.0f;
			}
		}
		const long long base = (long long)b * g.hw * g.cp + v * 8;
		for (int p = p0 + lane_p; p < p1; p += lanes_p * U) {
			Raw8<T> raw[U];
#pragma unroll
			for (int u = 0; u < U; u++) { const int pp = p + u * lanes_p; if (pp < p1) raw[u] = load_raw8<T>(x + base + (long long)pp * g.cp); }
#pragma unroll
			for (int u = 0; u < U; u++) {
				const int pp = p + u * lanes_p;
				if (pp >= p1) continue;
				float xv[8], out[8];
				unpack8(raw[u], xv);
#pragma unroll
				for (int j = 0; j < 8; j++) out[j] = fmaf(xv[j], sc[j], sh[j]);
				store8<T>(y + base + (long long)pp * g.cp, out);
			}
		}
	}
}

// per-channel constants of the backward apply: dx = k0*(n*d - d_beta - (x - mu)*rstd*d_gamma) rewritten per channel as
// ca*d + cx*x + cc; pass-through groups (set_off): ca = 1; dead samples / pad channels: all zero.  d_gamma / d_beta
// (forward: mean / var) may point into shared memory (chunked kernels): generic loads, no __restrict__.
__device__ __forceinline__ void bwd_constants(const NormGeom& g, int b, int v, const float* __restrict__ gamma, const float* mean, const float* var,
                                              const float* d_gamma, const float* d_beta, float (&ca)[8], float (&cx)[8], float (&cc)[8]) {
	const bool dead = b >= g.length;
	const float n = (float)(g.group_size * g.hw);
	const float inv_n = 1.0f / n;
#pragma unroll
	for (int j = 0; j < 8; j++) {
		const int ch = v * 8 + j;
		ca[j] = 0.0f; cx[j] = 0.0f; cc[j] = 0.0f;
		if (ch < g.c && !dead) {
			const int grp = ch / g.group_size;
			if (grp < g.nb_group - g.set_off) {
				const int s = b * g.nb_group + grp;
				const float rstd = 1.0f / sqrtf(var[s] + g.eps);
				const float k0 = inv_n * gamma[grp] * rstd;
				const float dg = d_gamma[s];
				ca[j] = k0 * n;
				cx[j] = -k0 * rstd * dg;
				cc[j] = k0 * (mean[s] * rstd * dg - d_beta[s]);
			} else ca[j] = 1.0f;
		}
	}
}

// a thread's column sums of dx -> the CTA's shared accumulators (lanes that own the same channel vector first)
__device__ __forceinline__ void fold_colsum(float (&csum)[8], int v, int lanes_c, bool active_thread, float* cs_acc) {
	bool writer = active_thread;
	if (lanes_c < 32 && (lanes_c & (lanes_c - 1)) == 0) {
#pragma unroll
		for (int j = 0; j < 8; j++)
			for (int off = lanes_c; off < 32; off <<= 1) csum[j] += __shfl_xor_sync(0xffffffffu, csum[j], off);
		writer = (threadIdx.x & 31) < lanes_c;
	}
	if (writer) {
#pragma unroll
		for (int j = 0; j < 8; j++) atomicAdd(&cs_acc[v * 8 + j], csum[j]);
	}
}

// dx = gamma*rstd/n * (n*d - d_beta - xhat*d_gamma), then the previous layer's deriv hook on x.
// cs_acc: [cp] column sums of dx accumulated in shared memory (nullptr: not wanted); zeroed and flushed by the caller.
template <typename T>
__device__ __forceinline__ void norm_bwd_apply_body(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx, const float* __restrict__ gamma,
                                                    const float* mean, const float* var, const float* d_gamma, const float* d_beta,
                                                    const cb200_activ& prev_activ, float* cs_acc, const NormGeom& g, int vbx, int b) {
	const int cv = g.cp >> 3;
	const int lanes_c = cv < NORM_THREADS ? cv : NORM_THREADS;
	const int lanes_p = NORM_THREADS / lanes_c;
	const int lane_c = threadIdx.x % lanes_c, lane_p = threadIdx.x / lanes_c;
	const bool active_thread = lane_p < lanes_p;
	const int p0 = vbx * g.ppb;
	int p1 = p0 + g.ppb;
	if (p1 > g.hw) p1 = g.hw;
	const int act = prev_activ.type;
	const float leak = prev_activ.leak, sat = prev_activ.saturation, abeta = prev_activ.beta;
	constexpr int U = 2;
	for (int v = lane_c; v < cv; v += lanes_c) {
		float ca[8], cx[8], cc[8];
		bwd_constants(g, b, v, gamma, mean, var, d_gamma, d_beta, ca, cx, cc);
		const long long base = (long long)b * g.hw * g.cp + v * 8;
		float csum[8];
#pragma unroll
		for (int j = 0; j < 8; j++) csum[j] = 0.0f;
		for (int p = p0 + lane_p; active_thread && p < p1; p += lanes_p * U) {
			Raw8<T> rx[U], rd[U];
#pragma unroll
			for (int u = 0; u < U; u++) {
				const int pp = p + u * lanes_p;
				if (pp < p1) { rx[u] = load_raw8<T>(x + base + (long long)pp * g.cp); rd[u] = load_raw8<T>(dy + base + (long long)pp * g.cp); }
			}
#pragma unroll
			for (int u = 0; u < U; u++) {
				const int pp = p + u * lanes_p;
				if (pp >= p1) continue;
				float xv[8], dv[8], out[8];
				unpack8(rx[u], xv);
				unpack8(rd[u], dv);
#pragma unroll
				for (int j = 0; j < 8; j++) out[j] = fmaf(ca[j], dv[j], fmaf(cx[j], xv[j], cc[j]));
				if (act == CB200_RELU) {
#pragma unroll
					for (int j = 0; j < 8; j++) out[j] = (xv[j] <= 0.0f || xv[j] > sat) ? out[j] * leak : out[j];
				} else if (act == CB200_LOGISTIC) {
#pragma unroll
					for (int j = 0; j < 8; j++) out[j] = out[j] * abeta * xv[j] * (1.0f - xv[j]);
				}
				store8<T>(dx + base + (long long)pp * g.cp, out);
#pragma unroll
				for (int j = 0; j < 8; j++) csum[j] += out[j];
			}
		}
		// column sums of the delta just produced = raw bias-column gradient of the preceding convolution
		if (cs_acc != nullptr) fold_colsum(csum, v, lanes_c, active_thread, cs_acc);
	}
}

// gsum[0][g] = sum_b d_gamma[b][g], gsum[1][g] = sum_b d_beta[b][g]; one warp per group, lanes over the batch
__global__ void norm_reduce_grads_kernel(const float* __restrict__ d_gamma, const float* __restrict__ d_beta,
                                         float* __restrict__ gsum, int batch, int nb_group) {
	const int grp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (grp >= nb_group) return;
	double sg = 0.0, sb = 0.0;
	for (int b = lane; b < batch; b += 32) { sg += d_gamma[b * nb_group + grp]; sb += d_beta[b * nb_group + grp]; }
	sg = warp_sum(sg);
	sb = warp_sum(sb);
	if (lane == 0) { gsum[grp] = (float)sg; gsum[nb_group + grp] = (float)sb; }
}

// upd = mom*upd + lr*(sum/B) ; param -= upd/S     (hyper[0] = lr/B_total, hyper[3] = S)
__global__ void norm_update_kernel(float* __restrict__ gamma, float* __restrict__ beta, float* __restrict__ gamma_upd,
                                   float* __restrict__ beta_upd, const float* __restrict__ gsum,
                                   const float* __restrict__ hyper, int nb_group, int set_off) {
	const int grp = blockIdx.x * blockDim.x + threadIdx.x;
	if (grp >= nb_group - set_off) return;
	const float alpha = hyper[0], mom = hyper[1], S = hyper[3];
	norm_param_step(alpha, mom, S, gsum[grp], gamma_upd[grp], gamma[grp]);
	norm_param_step(alpha, mom, S, gsum[nb_group + grp], beta_upd[grp], beta[grp]);
}

// single-GPU form of the two kernels above in one launch (no all-reduce between them): one warp per group sums the
// per-sample gradients over the batch and its first lane applies the update; same arithmetic, gsum still written
__global__ void norm_reduce_update_kernel(const float* __restrict__ d_gamma, const float* __restrict__ d_beta, float* __restrict__ gsum,
                                          float* __restrict__ gamma, float* __restrict__ beta, float* __restrict__ gamma_upd,
                                          float* __restrict__ beta_upd, const float* __restrict__ hyper, int batch, int nb_group, int set_off) {
	const int grp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (grp >= nb_group) return;
	double sg = 0.0, sb = 0.0;
	for (int b = lane; b < batch; b += 32) { sg += d_gamma[b * nb_group + grp]; sb += d_beta[b * nb_group + grp]; }
	sg = warp_sum(sg);
	sb = warp_sum(sb);
	if (lane != 0) return;
	const float fg = (float)sg, fb = (float)sb;
	gsum[grp] = fg; gsum[nb_group + grp] = fb;
	if (grp >= nb_group - set_off) return;
	const float alpha = hyper[0], mom = hyper[1], S = hyper[3];
	norm_param_step(alpha, mom, S, fg, gamma_upd[grp], gamma[grp]);
	norm_param_step(alpha, mom, S, fb, beta_upd[grp], beta[grp]);
}

// ---------------------------------------------------------------- group-norm + 2x2 max-pool, fused
// Geometry of the pooled side; the norm geometry `n` describes the input-sized tensor (n.hw = in_h * in_w).
struct FusedGeom {
	NormGeom n;
	int in_w, out_w, out_hw, ppb_out;
};

template <typename T> __device__ __forceinline__ float round_to_storage(float v) { return to_f32<T>(from_f32<T>(v)); }
template <> __device__ __forceinline__ float round_to_storage<float>(float v) { return v; }

__device__ __forceinline__ uint32_t map_byte(const uint2& m, int j) { return ((j < 4 ? m.x : m.y) >> (8 * (j & 3))) & 0xffu; }

// per-channel affine constants of the forward apply (zero for pad channels and samples beyond `length`)
__device__ __forceinline__ void fwd_constants(const NormGeom& g, int b, int v, const float* __restrict__ gamma, const float* __restrict__ beta,
                                              const float* mean, const float* var, float (&sc)[8], float (&sh)[8]) {
	const bool dead = b >= g.length;
#pragma unroll
	for (int j = 0; j < 8; j++) {
		const int ch = v * 8 + j;
		sc[j] = 0.0f; sh[j] = 0.0f;
		if (ch < g.c && !dead) {
			const int grp = ch / g.group_size;
			if (grp < g.nb_group - g.set_off) {
				const float rstd = 1.0f / sqrtf(var[b * g.nb_group + grp] + g.eps);
				sc[j] = gamma[grp] * rstd;
				sh[j] = beta[grp] - mean[b * g.nb_group + grp] * sc[j];
			} else sc[j] = 1.0f;
		}
	}
}

// The window maximum of four 8-channel candidates in PACKED 16-bit arithmetic (the kernel is instruction-bound: 61 % ALU
// pipe, 56 % issue slots on the first Darknet19 layer).  The candidates are rounded to the storage type first - two
// values per cvt - and compared as stored, so value and index equal the scalar form below: first strict maximum in scan
// order = (c1 > c0 ? 1 : 0) / (c3 > c2 ? 3 : 2), the second pair winning only when strictly larger.  Comparison masks
// (0xffff per 16-bit lane) are combined into the 2-bit index per lane and the eight indices packed with byte permutes.
__device__ __forceinline__ __half2 pk2(float a, float b, const __half*) { return __floats2half2_rn(a, b); }
__device__ __forceinline__ __nv_bfloat162 pk2(float a, float b, const __nv_bfloat16*) { return __floats2bfloat162_rn(a, b); }
template <typename T>
__device__ __forceinline__ void pool4_select_packed(const float (&y0)[8], const float (&y1)[8], const float (&y2)[8], const float (&y3)[8],
                                                    uint4& best, uint2& arg) {
	uint32_t bw[4], aw[4];
#pragma unroll
	for (int i = 0; i < 4; i++) {
		const auto c0 = pk2(y0[2 * i], y0[2 * i + 1], (const T*)nullptr), c1 = pk2(y1[2 * i], y1[2 * i + 1], (const T*)nullptr);
		const auto c2 = pk2(y2[2 * i], y2[2 * i + 1], (const T*)nullptr), c3 = pk2(y3[2 * i], y3[2 * i + 1], (const T*)nullptr);
		const auto m01 = __hmax2(c0, c1), m23 = __hmax2(c2, c3);
		const uint32_t g1 = __hgt2_mask(c1, c0), g3 = __hgt2_mask(c3, c2), gb = __hgt2_mask(m23, m01);
		const auto m = __hmax2(m01, m23);
		bw[i] = *reinterpret_cast<const uint32_t*>(&m);
		const uint32_t a = (gb & 0x00020002u) | (((g3 & gb) | (g1 & ~gb)) & 0x00010001u);      // index per 16-bit lane
		aw[i] = a;
	}
	best = make_uint4(bw[0], bw[1], bw[2], bw[3]);
	arg.x = __byte_perm(aw[0], aw[1], 0x6420);      // bytes: lane 0 / 1 of pair 0, lane 0 / 1 of pair 1
	arg.y = __byte_perm(aw[2], aw[3], 0x6420);
}

// y = x*sc + sh rounded to the storage type (what the unfused apply would have stored), then the first strict maximum
// of the window in scan order (0,0) (0,1) (1,0) (1,1); only the pooled value and its window index are written
template <typename T>
__device__ __forceinline__ void norm_pool_fwd_body(const T* __restrict__ x, T* __restrict__ pooled, uint8_t* __restrict__ map, const float* __restrict__ gamma,
                                                   const float* __restrict__ beta, const float* mean, const float* var, const FusedGeom& f, int vbx, int b) {
	const NormGeom& g = f.n;
	const int cv = g.cp >> 3;
	const int lanes_c = cv < NORM_THREADS ? cv : NORM_THREADS;
	const int lanes_p = NORM_THREADS / lanes_c;
	const int lane_c = threadIdx.x % lanes_c, lane_p = threadIdx.x / lanes_c;
	if (lane_p >= lanes_p) return;
	const int q0 = vbx * f.ppb_out;
	int q1 = q0 + f.ppb_out;
	if (q1 > f.out_hw) q1 = f.out_hw;
	const long long row = (long long)f.in_w * g.cp;
	for (int v = lane_c; v < cv; v += lanes_c) {
		float sc[8], sh[8];
		fwd_constants(g, b, v, gamma, beta, mean, var, sc, sh);
		const T* xb = x + (long long)b * g.hw * g.cp + v * 8;
		const long long ob = (long long)b * f.out_hw * g.cp + v * 8;
		for (int q = q0 + lane_p; q < q1; q += lanes_p) {
			const int oy = q / f.out_w, ox = q - oy * f.out_w;
			const T* p = xb + (long long)(2 * oy) * row + (long long)(2 * ox) * g.cp;
			const Raw8<T> r0 = load_raw8<T>(p), r1 = load_raw8<T>(p + g.cp), r2 = load_raw8<T>(p + row), r3 = load_raw8<T>(p + row + g.cp);
			float a0[8], a1[8], a2[8], a3[8], best[8];
			unpack8(r0, a0); unpack8(r1, a1); unpack8(r2, a2); unpack8(r3, a3);
			const long long o = ob + (long long)q * g.cp;
			if constexpr (sizeof(T) == 2) {
				if (v * 8 + 8 <= g.c) {                 // (a vector with pad channels takes the scalar form below)
#pragma unroll
					for (int j = 0; j < 8; j++) {
						a0[j] = fmaf(a0[j], sc[j], sh[j]); a1[j] = fmaf(a1[j], sc[j], sh[j]);
						a2[j] = fmaf(a2[j], sc[j], sh[j]); a3[j] = fmaf(a3[j], sc[j], sh[j]);
					}
					uint4 bp; uint2 ap;
					pool4_select_packed<T>(a0, a1, a2, a3, bp, ap);
					*reinterpret_cast<uint4*>(pooled + o) = bp;
					if (map != nullptr) *reinterpret_cast<uint2*>(map + o) = ap;
					continue;
				}
			}
			uint32_t arg[8];
#pragma unroll
			for (int j = 0; j < 8; j++) {
				best[j] = round_to_storage<T>(fmaf(a0[j], sc[j], sh[j]));
				arg[j] = 0;
				float t = round_to_storage<T>(fmaf(a1[j], sc[j], sh[j]));
				if (t > best[j]) { best[j] = t; arg[j] = 1; }
				t = round_to_storage<T>(fmaf(a2[j], sc[j], sh[j]));
				if (t > best[j]) { best[j] = t; arg[j] = 2; }
				t = round_to_storage<T>(fmaf(a3[j], sc[j], sh[j]));
				if (t > best[j]) { best[j] = t; arg[j] = 3; }
				if (v * 8 + j >= g.c) { best[j] = 0.0f; arg[j] = 255; }
			}
			store8<T>(pooled + o, best);
			uint2 packed;
			packed.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
			packed.y = arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24);
			if (map != nullptr) *reinterpret_cast<uint2*>(map + o) = packed;
		}
	}
}

// backward reductions: the delta of the normalised tensor is the pooled delta at the selected window position and
// zero elsewhere, so sum(d) and sum(d*x) run over the pooled pixels, reading x at the position the map names
template <typename T>
__device__ __forceinline__ void norm_pool_bwd_stats_body(const T* __restrict__ x, const T* __restrict__ dp, const uint8_t* __restrict__ map,
                                                         double* __restrict__ ws, const FusedGeom& f, int vbx, int b, float* sm_acc) {
	const NormGeom& g = f.n;
	const int cv = g.cp >> 3;
	for (int i = threadIdx.x; i < g.nb_group * 2; i += blockDim.x) sm_acc[i] = 0.0f;
	__syncthreads();
	const int lanes_c = cv < NORM_THREADS ? cv : NORM_THREADS;
	const int lanes_p = NORM_THREADS / lanes_c;
	const int lane_c = threadIdx.x % lanes_c, lane_p = threadIdx.x / lanes_c;
	const bool active = lane_p < lanes_p;
	const int q0 = vbx * f.ppb_out;
	int q1 = q0 + f.ppb_out;
	if (q1 > f.out_hw) q1 = f.out_hw;
	const long long row = (long long)f.in_w * g.cp;
	for (int v = lane_c; v < cv; v += lanes_c) {
		float s0[8], s1[8];
#pragma unroll
		for (int j = 0; j < 8; j++) { s0[j] = 0.0f; s1[j] = 0.0f; }
		const T* xb = x + (long long)b * g.hw * g.cp + v * 8;
		const long long ob = (long long)b * f.out_hw * g.cp + v * 8;
		for (int q = q0 + lane_p; active && q < q1; q += lanes_p) {
			const int oy = q / f.out_w, ox = q - oy * f.out_w;
			const T* p = xb + (long long)(2 * oy) * row + (long long)(2 * ox) * g.cp;
			const long long o = ob + (long long)q * g.cp;
			const Raw8<T> rd = load_raw8<T>(dp + o);
			const uint2 rm = __ldg(reinterpret_cast<const uint2*>(map + o));
			const Raw8<T> r0 = load_raw8<T>(p), r1 = load_raw8<T>(p + g.cp), r2 = load_raw8<T>(p + row), r3 = load_raw8<T>(p + row + g.cp);
			float d[8], a0[8], a1[8], a2[8], a3[8];
			unpack8(rd, d); unpack8(r0, a0); unpack8(r1, a1); unpack8(r2, a2); unpack8(r3, a3);
#pragma unroll
			for (int j = 0; j < 8; j++) {
				const uint32_t m = map_byte(rm, j);
				const float xs = m == 0 ? a0[j] : (m == 1 ? a1[j] : (m == 2 ? a2[j] : a3[j]));
				if (m < 4) { s0[j] += d[j]; s1[j] += d[j] * xs; }
			}
		}
		fold_into_groups(s0, s1, v, lanes_c, g, sm_acc);
	}
	__syncthreads();
	for (int i = threadIdx.x; i < g.nb_group * 2; i += blockDim.x)
		atomicAdd(&ws[(size_t)b * g.nb_group * 2 + i], (double)sm_acc[i]);
}

template <typename T>
__global__ void __launch_bounds__(NORM_THREADS)
norm_pool_bwd_stats_kernel(const T* __restrict__ x, const T* __restrict__ dp, const uint8_t* __restrict__ map, double* __restrict__ ws, FusedGeom f) {
	extern __shared__ float sm_acc[];
	norm_pool_bwd_stats_body<T>(x, dp, map, ws, f, blockIdx.x, blockIdx.y, sm_acc);
}

// The same two sums WITHOUT reading the full-resolution input: at the selected position the forward pass stored
// y = x*sc + sh (the pooled output, kept for the next layer's weight gradient), so x = (y - sh) / sc there and
// sum(d*x) runs over the pooled delta and the pooled output alone - 2 quarter-size tensors instead of the input-sized
// one + delta + map (first Darknet19 layer at batch 128: 0.82 GB instead of 2.27 GB).  y carries one rounding to the
// storage type, like x does; dividing by sc amplifies it by |y| / |x*sc| ~ 1 + |beta / gamma|, so groups with
// |beta| > 16 |gamma| or |gamma| < 1e-4, and dead samples (whose pooled output is zero), keep the gather from x.
template <typename T>
__device__ __forceinline__ void norm_pool_bwd_stats_y_body(const T* __restrict__ x, const T* __restrict__ dp, const uint8_t* __restrict__ map,
                                                           const T* __restrict__ yp, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const float* __restrict__ mean, const float* __restrict__ var,
                                                           double* __restrict__ ws, const FusedGeom& f, int vbx, int b, float* sm_acc) {
	const NormGeom& g = f.n;
	const int cv = g.cp >> 3;
	for (int i = threadIdx.x; i < g.nb_group * 2; i += blockDim.x) sm_acc[i] = 0.0f;
	__syncthreads();
	const int lanes_c = cv < NORM_THREADS ? cv : NORM_THREADS;
	const int lanes_p = NORM_THREADS / lanes_c;
	const int lane_c = threadIdx.x % lanes_c, lane_p = threadIdx.x / lanes_c;
	const bool active = lane_p < lanes_p;
	const int q0 = vbx * f.ppb_out;
	int q1 = q0 + f.ppb_out;
	if (q1 > f.out_hw) q1 = f.out_hw;
	const long long row = (long long)f.in_w * g.cp;
	const bool dead = b >= g.length;
	for (int v = lane_c; v < cv; v += lanes_c) {
		float isc[8], off[8];
		uint32_t slow = 0;
#pragma unroll
		for (int j = 0; j < 8; j++) {
			const int ch = v * 8 + j;
			isc[j] = 0.0f; off[j] = 0.0f;
			if (ch >= g.c) continue;
			const int grp = ch / g.group_size;
			if (dead) slow |= 1u << j;
			else if (grp < g.nb_group - g.set_off) {
				const float ga = gamma[grp], be = beta[grp];
				if (fabsf(ga) < 1e-4f || fabsf(be) > 16.0f * fabsf(ga)) slow |= 1u << j;
				else {
					const float rstd = 1.0f / sqrtf(var[b * g.nb_group + grp] + g.eps);
					const float sc = ga * rstd;
					const float sh = be - mean[b * g.nb_group + grp] * sc;
					isc[j] = 1.0f / sc;
					off[j] = -sh * isc[j];
				}
			} else isc[j] = 1.0f;                      // pass-through group: y == x
		}
		float s0[8], s1[8];
#pragma unroll
		for (int j = 0; j < 8; j++) { s0[j] = 0.0f; s1[j] = 0.0f; }
		const T* xb = x + (long long)b * g.hw * g.cp + v * 8;
		const long long ob = (long long)b * f.out_hw * g.cp + v * 8;
		constexpr int U = 4;
		for (int q = q0 + lane_p; active && q < q1; q += lanes_p * U) {
			Raw8<T> rd[U], ry[U];
#pragma unroll
			for (int u = 0; u < U; u++) {
				const int qq = q + u * lanes_p;
				if (qq < q1) { const long long o = ob + (long long)qq * g.cp; rd[u] = load_raw8<T>(dp + o); ry[u] = load_raw8<T>(yp + o); }
			}
#pragma unroll
			for (int u = 0; u < U; u++) {
				const int qq = q + u * lanes_p;
				if (qq >= q1) continue;
				float d[8], yv[8], xs[8];
				unpack8(rd[u], d); unpack8(ry[u], yv);
#pragma unroll
				for (int j = 0; j < 8; j++) xs[j] = fmaf(yv[j], isc[j], off[j]);
				if (slow != 0) {
					const int oy = qq / f.out_w, ox = qq - oy * f.out_w;
					const T* p = xb + (long long)(2 * oy) * row + (long long)(2 * ox) * g.cp;
					const uint2 rm = __ldg(reinterpret_cast<const uint2*>(map + ob + (long long)qq * g.cp));
					float a0[8], a1[8], a2[8], a3[8];
					unpack8(load_raw8<T>(p), a0); unpack8(load_raw8<T>(p + g.cp), a1);
					unpack8(load_raw8<T>(p + row), a2); unpack8(load_raw8<T>(p + row + g.cp), a3);
#pragma unroll
					for (int j = 0; j < 8; j++) {
						const uint32_t m = map_byte(rm, j);
						if ((slow >> j) & 1u) xs[j] = m == 0 ? a0[j] : (m == 1 ? a1[j] : (m == 2 ? a2[j] : a3[j]));
					}
				}
#pragma unroll
				for (int j = 0; j < 8; j++) { s0[j] += d[j]; s1[j] += d[j] * xs[j]; }
			}
		}
		fold_into_groups(s0, s1, v, lanes_c, g, sm_acc);
	}
	__syncthreads();
	for (int i = threadIdx.x; i < g.nb_group * 2; i += blockDim.x)
		atomicAdd(&ws[(size_t)b * g.nb_group * 2 + i], (double)sm_acc[i]);
}

template <typename T>
__global__ void __launch_bounds__(NORM_THREADS)
norm_pool_bwd_stats_y_kernel(const T* __restrict__ x, const T* __restrict__ dp, const uint8_t* __restrict__ map, const T* __restrict__ yp,
                             const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean,
                             const float* __restrict__ var, double* __restrict__ ws, FusedGeom f) {
	extern __shared__ float sm_acc[];
	norm_pool_bwd_stats_y_body<T>(x, dp, map, yp, gamma, beta, mean, var, ws, f, blockIdx.x, blockIdx.y, sm_acc);
}

// dx of the four input pixels of each window: ca*d + cx*x + cc with d = pooled delta at the selected position, else 0;
// then the previous layer's derivative hook and the column sums, exactly like norm_bwd_apply_kernel
template <typename T>
__device__ __forceinline__ void norm_pool_bwd_apply_body(const T* __restrict__ x, const T* __restrict__ dp, const uint8_t* __restrict__ map, T* __restrict__ dx,
                                                         const float* __restrict__ gamma, const float* mean, const float* var,
                                                         const float* d_gamma, const float* d_beta, const cb200_activ& prev_activ,
                                                         float* cs_acc, const FusedGeom& f, int vbx, int b) {
	const NormGeom& g = f.n;
	const int cv = g.cp >> 3;
	const int lanes_c = cv < NORM_THREADS ? cv : NORM_THREADS;
	const int lanes_p = NORM_THREADS / lanes_c;
	const int lane_c = threadIdx.x % lanes_c, lane_p = threadIdx.x / lanes_c;
	const bool active_thread = lane_p < lanes_p;
	const int q0 = vbx * f.ppb_out;
	int q1 = q0 + f.ppb_out;
	if (q1 > f.out_hw) q1 = f.out_hw;
	const int act = prev_activ.type;
	const float leak = prev_activ.leak, sat = prev_activ.saturation, abeta = prev_activ.beta;
	const long long row = (long long)f.in_w * g.cp;
	for (int v = lane_c; v < cv; v += lanes_c) {
		float ca[8], cx[8], cc[8];
		bwd_constants(g, b, v, gamma, mean, var, d_gamma, d_beta, ca, cx, cc);
		const long long xb = (long long)b * g.hw * g.cp + v * 8;
		const long long ob = (long long)b * f.out_hw * g.cp + v * 8;
		float csum[8];
#pragma unroll
		for (int j = 0; j < 8; j++) csum[j] = 0.0f;
		for (int q = q0 + lane_p; active_thread && q < q1; q += lanes_p) {
			const int oy = q / f.out_w, ox = q - oy * f.out_w;
			const long long pi = xb + (long long)(2 * oy) * row + (long long)(2 * ox) * g.cp;
			const long long o = ob + (long long)q * g.cp;
			const Raw8<T> rd = load_raw8<T>(dp + o);
			const uint2 rm = __ldg(reinterpret_cast<const uint2*>(map + o));
			Raw8<T> rx[4];
			rx[0] = load_raw8<T>(x + pi); rx[1] = load_raw8<T>(x + pi + g.cp);
			rx[2] = load_raw8<T>(x + pi + row); rx[3] = load_raw8<T>(x + pi + row + g.cp);
			float d[8];
			unpack8(rd, d);
#pragma unroll
			for (int k = 0; k < 4; k++) {
				float xv[8], out[8];
				unpack8(rx[k], xv);
#pragma unroll
				for (int j = 0; j < 8; j++) {
					const float dv = map_byte(rm, j) == (uint32_t)k ? d[j] : 0.0f;
					out[j] = fmaf(ca[j], dv, fmaf(cx[j], xv[j], cc[j]));
				}
				if (act == CB200_RELU) {
#pragma unroll
					for (int j = 0; j < 8; j++) out[j] = (xv[j] <= 0.0f || xv[j] > sat) ? out[j] * leak : out[j];
				} else if (act == CB200_LOGISTIC) {
#pragma unroll
					for (int j = 0; j < 8; j++) out[j] = out[j] * abeta * xv[j] * (1.0f - xv[j]);
				}
				store8<T>(dx + pi + (k >> 1) * row + (k & 1) * g.cp, out);
#pragma unroll
				for (int j = 0; j < 8; j++) csum[j] += out[j];
			}
		}
		if (cs_acc != nullptr) fold_colsum(csum, v, lanes_c, active_thread, cs_acc);
	}
}

// ---------------------------------------------------------------- launches: statistics role / apply role in one grid
// DEFAULT (two passes): a statistics launch over the whole batch, then ONE launch whose blocks all take the apply role and
// turn the FP64 sums of their sample into mean / var (forward) or d_gamma / d_beta (backward) themselves - no separate
// finalize launch on the critical path (36 tiny launches per Darknet19 step).
// OPTIONAL (chunked, cb200_norm_set_pipeline): the two-pass form streams the tensor from HBM twice; the chunked form
// walks the batch in chunks of a few images sized to stay in L2 (126 MB): launch i runs, side by side in ONE grid,
// the statistics blocks of chunk i and the apply blocks of chunk i-1, so the apply re-reads its chunk from L2 while the
// statistics blocks stream the next one from HBM - forward 3 -> 2 passes over HBM, backward 5 -> 3 - and the finalize
// launch disappears: every apply block turns the FP64 sums of its sample into mean / var (or d_gamma / d_beta) in
// shared memory itself (the first block of a sample also stores them for the backward pass / the optimizer).
// Dependencies are plain stream order (statistics of chunk i-1 were launch i-1); blocks keep the occupancy of the
// separate kernels.
// MEASURED (B200, batch 128, Darknet19 shapes, profiles/r1_gn_chunked_sweep.txt): SLOWER than the two passes at every
// chunk size (1.3-2x on the 448 / 224 px layers, 1.2-1.5x on the 112 / 56 px ones): a launch that moves 25-75 MB lasts
// 15-35 us where the bytes alone would take 5-14 us - short grids never reach the steady state in which the two-pass
// kernels stream at 4.5-5.5 TB/s - so what the apply saves by finding its chunk in L2 is lost several times over.  A
// persistent cooperative kernel with per-chunk counters was measured before that and lost 2-4x (2-3 CTAs per SM, one
// batch of loads per phase in flight, hand-over latency exposed: profiles/r1_gn_pipeline_sweep.txt).  The chunked form
// therefore stays OFF by default (cb200_norm_set_pipeline / CB200_GN_PIPELINE=1), kept as a tested option; getting
// under two HBM passes needs the tile to stay on chip (TMA-staged slabs per CTA cluster), not a second launch.
struct ChunkGeom {
	int nblk_a, nbx_a, a_b0;     // statistics role: blocks in this launch, blocks per sample, first sample
	int nbx_b, b_b0;             // apply role: blocks per sample, first sample
};

// statistics of sample b -> shared memory [mean | var] (forward): mean = S1/n, var = S2/n - mean^2 (biased, as upstream);
// the first block of the sample keeps them for backward
__device__ __forceinline__ void chunk_finalize_fwd(const double* ws, float* mean, float* var, const NormGeom& g, int b, bool keep, float* sm) {
	const double n = (double)g.group_size * g.hw;
	for (int i = threadIdx.x; i < g.nb_group; i += blockDim.x) {
		const int s = b * g.nb_group + i;
		const double m = ws[2 * s] / n;
		double v = ws[2 * s + 1] / n - m * m;
		if (v < 0.0) v = 0.0;
		sm[i] = (float)m;
		sm[g.nb_group + i] = (float)v;
		if (keep) { mean[s] = (float)m; var[s] = (float)v; }
	}
	__syncthreads();
}
// sums of sample b -> shared memory [d_gamma | d_beta] (backward): d_beta = sum d, d_gamma = (sum d*x - mean*sum d) / sqrt(var+eps);
// the first block stores them for the optimizer
__device__ __forceinline__ void chunk_finalize_bwd(const double* ws, const float* mean, const float* var, float* d_gamma, float* d_beta,
                                                   const NormGeom& g, int b, bool keep, float* sm) {
	for (int i = threadIdx.x; i < g.nb_group; i += blockDim.x) {
		const int s = b * g.nb_group + i;
		const double sd = ws[2 * s], sdx = ws[2 * s + 1];
		const float db = (float)sd;
		const float dg = (float)((sdx - (double)mean[s] * sd) / sqrt((double)var[s] + (double)g.eps));
		sm[i] = dg;
		sm[g.nb_group + i] = db;
		if (keep) { d_gamma[s] = dg; d_beta[s] = db; }
	}
	__syncthreads();
}

// shared memory: [2 * nb_group] (block sums of the statistics role, per-sample constants of the apply role), then [cp]
// column sums of dx in the backward kernels
template <typename T>
__global__ void __launch_bounds__(NORM_THREADS)
norm_fwd_chunk_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ gamma, const float* __restrict__ beta,
                      float* mean, float* var, double* ws, NormGeom g, ChunkGeom cg) {
	extern __shared__ float sm[];
	if ((int)blockIdx.x < cg.nblk_a) {
		const int bi = blockIdx.x / cg.nbx_a;
		norm_stats_body<T, false>(x, nullptr, ws, g, blockIdx.x - bi * cg.nbx_a, cg.a_b0 + bi, sm);
		return;
	}
	const int vb = blockIdx.x - cg.nblk_a, bi = vb / cg.nbx_b, vbx = vb - bi * cg.nbx_b, b = cg.b_b0 + bi;
	chunk_finalize_fwd(ws, mean, var, g, b, vbx == 0, sm);
	norm_apply_body<T>(x, y, gamma, beta, sm - b * g.nb_group, sm + g.nb_group - b * g.nb_group, g, vbx, b);
}

// (occupancy, measured on the Darknet19 shapes: four blocks per SM for the fused forward - 64 registers after the packed
//  maximum - gain 3 %, three for the fused backward apply - 80 registers, 48 bytes spilled - gain 5-8 %; forcing the
//  plain backward apply or the statistics kernels to four, or the pooled-output statistics to three, spills and loses)
template <typename T>
__global__ void __launch_bounds__(NORM_THREADS, 4)
norm_pool_fwd_chunk_kernel(const T* __restrict__ x, T* __restrict__ pooled, uint8_t* __restrict__ map, const float* __restrict__ gamma,
                           const float* __restrict__ beta, float* mean, float* var, double* ws, FusedGeom f, ChunkGeom cg) {
	extern __shared__ float sm[];
	const NormGeom& g = f.n;
	if ((int)blockIdx.x < cg.nblk_a) {
		const int bi = blockIdx.x / cg.nbx_a;
		norm_stats_body<T, false>(x, nullptr, ws, g, blockIdx.x - bi * cg.nbx_a, cg.a_b0 + bi, sm);
		return;
	}
	const int vb = blockIdx.x - cg.nblk_a, bi = vb / cg.nbx_b, vbx = vb - bi * cg.nbx_b, b = cg.b_b0 + bi;
	chunk_finalize_fwd(ws, mean, var, g, b, vbx == 0, sm);
	norm_pool_fwd_body<T>(x, pooled, map, gamma, beta, sm - b * g.nb_group, sm + g.nb_group - b * g.nb_group, f, vbx, b);
}

template <typename T>
__global__ void __launch_bounds__(NORM_THREADS, 3)
norm_bwd_chunk_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx, const float* __restrict__ gamma,
                      const float* __restrict__ mean, const float* __restrict__ var, float* d_gamma, float* d_beta,
                      cb200_activ prev_activ, float* __restrict__ colsum, double* ws, NormGeom g, ChunkGeom cg) {
	extern __shared__ float sm[];
	if ((int)blockIdx.x < cg.nblk_a) {
		const int bi = blockIdx.x / cg.nbx_a;
		norm_stats_body<T, true>(dy, x, ws, g, blockIdx.x - bi * cg.nbx_a, cg.a_b0 + bi, sm);
		return;
	}
	const int vb = blockIdx.x - cg.nblk_a, bi = vb / cg.nbx_b, vbx = vb - bi * cg.nbx_b, b = cg.b_b0 + bi;
	float* cs_acc = colsum != nullptr ? sm + 2 * g.nb_group : nullptr;
	if (cs_acc != nullptr) for (int i = threadIdx.x; i < g.cp; i += blockDim.x) cs_acc[i] = 0.0f;
	chunk_finalize_bwd(ws, mean, var, d_gamma, d_beta, g, b, vbx == 0, sm);
	norm_bwd_apply_body<T>(x, dy, dx, gamma, mean, var, sm - b * g.nb_group, sm + g.nb_group - b * g.nb_group, prev_activ, cs_acc, g, vbx, b);
	if (cs_acc != nullptr) {
		__syncthreads();
		for (int i = threadIdx.x; i < g.c; i += blockDim.x) atomicAdd(&colsum[i], cs_acc[i]);
	}
}

template <typename T>
__global__ void __launch_bounds__(NORM_THREADS, 3)
norm_pool_bwd_chunk_kernel(const T* __restrict__ x, const T* __restrict__ dp, const uint8_t* __restrict__ map, T* __restrict__ dx,
                           const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ var,
                           float* d_gamma, float* d_beta, cb200_activ prev_activ, float* __restrict__ colsum, double* ws,
                           FusedGeom f, ChunkGeom cg) {
	extern __shared__ float sm[];
	const NormGeom& g = f.n;
	if ((int)blockIdx.x < cg.nblk_a) {
		const int bi = blockIdx.x / cg.nbx_a;
		norm_pool_bwd_stats_body<T>(x, dp, map, ws, f, blockIdx.x - bi * cg.nbx_a, cg.a_b0 + bi, sm);
		return;
	}
	const int vb = blockIdx.x - cg.nblk_a, bi = vb / cg.nbx_b, vbx = vb - bi * cg.nbx_b, b = cg.b_b0 + bi;
	float* cs_acc = colsum != nullptr ? sm + 2 * g.nb_group : nullptr;
	if (cs_acc != nullptr) for (int i = threadIdx.x; i < g.cp; i += blockDim.x) cs_acc[i] = 0.0f;
	chunk_finalize_bwd(ws, mean, var, d_gamma, d_beta, g, b, vbx == 0, sm);
	norm_pool_bwd_apply_body<T>(x, dp, map, dx, gamma, mean, var, sm - b * g.nb_group, sm + g.nb_group - b * g.nb_group, prev_activ, cs_acc, f, vbx, b);
	if (cs_acc != nullptr) {
		__syncthreads();
		for (int i = threadIdx.x; i < g.c; i += blockDim.x) atomicAdd(&colsum[i], cs_acc[i]);
	}
}

// ---- host side of the chunked launches
static int env_int(const char* name, int dflt) {
	const char* e = getenv(name);
	return (e != nullptr && *e != '\0') ? atoi(e) : dflt;
}
// on = 0: the statistics / finalize / apply launches over the whole batch (kept as the reference implementation of the
// same arithmetic; tests compare both)
static int g_norm_pipeline = -1, g_norm_chunk_kb = 0, g_norm_blocks_per_sm = 0;
static bool norm_pipeline_on() {
	if (g_norm_pipeline < 0) {
		g_norm_pipeline = env_int("CB200_GN_PIPELINE", 0) != 0;
		if (g_norm_chunk_kb <= 0) g_norm_chunk_kb = env_int("CB200_GN_CHUNK_MB", 24) * 1024;
		if (g_norm_blocks_per_sm <= 0) g_norm_blocks_per_sm = env_int("CB200_GN_BLOCKS_PER_SM", 6);
	}
	return g_norm_pipeline != 0;
}

// samples per chunk: as many as fit the chunk size in bytes read by the statistics role (x, or x + dy)
static int chunk_samples(int batch, double bytes_per_sample) {
	int k = (int)((double)g_norm_chunk_kb * 1024.0 / bytes_per_sample);
	if (k < 1) k = 1;
	return k > batch ? batch : k;
}
// pixels per block such that one chunk of k samples gives each role about blocks_per_sm blocks per SM
static int chunk_ppb(int hw, int k, int cv) {
	const int lanes_c = cv < NORM_THREADS ? cv : NORM_THREADS;
	const int lanes_p = NORM_THREADS / lanes_c;
	const long long want = (long long)g_num_sms * g_norm_blocks_per_sm;
	long long ppb = ((long long)hw * k + want - 1) / want;
	const int min_ppb = lanes_p * 4;
	if (ppb < min_ppb) ppb = min_ppb;
	if (ppb > 1024) ppb = 1024;
	return (int)ppb;
}
// whole batch, apply role only (the default two-pass form: statistics launch, then finalize + apply in one launch)
static ChunkGeom apply_only_geom(int batch, int nbx, unsigned& grid) {
	ChunkGeom cg;
	cg.nblk_a = 0; cg.nbx_a = 1; cg.a_b0 = 0; cg.nbx_b = nbx; cg.b_b0 = 0;
	grid = (unsigned)(batch * nbx);
	return cg;
}
// launch i of nchunks + 1: statistics of chunk i, apply of chunk i - 1
static bool chunk_launch_geom(int i, int batch, int k, int nbx_a, int nbx_b, ChunkGeom& cg, unsigned& grid) {
	const int nchunks = (batch + k - 1) / k;
	if (i > nchunks) return false;
	const int a_nb = i < nchunks ? (batch - i * k < k ? batch - i * k : k) : 0;
	const int b_nb = i >= 1 ? (batch - (i - 1) * k < k ? batch - (i - 1) * k : k) : 0;
	cg.nbx_a = nbx_a; cg.nblk_a = a_nb * nbx_a; cg.a_b0 = i * k;
	cg.nbx_b = nbx_b; cg.b_b0 = (i - 1) * k;
	grid = (unsigned)(cg.nblk_a + b_nb * nbx_b);
	return true;
}

static int fill_geom(const cb200_norm_desc* d, NormGeom& g) {
	CB_ARG(d != nullptr && d->batch > 0 && d->c > 0 && d->group_size > 0 && d->nb_group > 0);
	CB_ARG(d->nb_group * d->group_size >= d->c);
	g.batch = d->batch; g.length = d->length; g.c = d->c; g.cp = round8(d->c); g.hw = d->h * d->w;
	g.group_size = d->group_size; g.nb_group = d->nb_group; g.set_off = d->set_off; g.eps = d->eps;
	g.ppb = norm_pix_per_block(g.hw, g.batch, g.cp >> 3);
	return CB200_OK;
}
}  // namespace cb200
using namespace cb200;

extern "C" {

// FP64 sums [batch][nb_group][2]
size_t cb200_norm_workspace_bytes(const cb200_norm_desc* d) { return sizeof(double) * 2 * (size_t)d->batch * d->nb_group; }

void cb200_norm_set_tuning(int want_blocks_per_sm, int ppb_max) {
	if (want_blocks_per_sm > 0) g_norm_want_blocks = want_blocks_per_sm;
	if (ppb_max > 0) g_norm_ppb_max = ppb_max;
}

void cb200_norm_set_pipeline(int on, int chunk_kb, int blocks_per_sm) {
	norm_pipeline_on();                    // environment defaults first
	g_norm_pipeline = on != 0;
	if (chunk_kb > 0) g_norm_chunk_kb = chunk_kb;
	if (blocks_per_sm > 0) g_norm_blocks_per_sm = blocks_per_sm;
}

int cb200_norm_forward(const cb200_norm_desc* d, const void* x, void* y, const float* gamma, const float* beta,
                       float* mean, float* var, void* workspace, void* s) {
	return cb200_norm_forward_ex(d, x, y, gamma, beta, mean, var, workspace, 0, s);
}

int cb200_norm_forward_ex(const cb200_norm_desc* d, const void* x, void* y, const float* gamma, const float* beta,
                          float* mean, float* var, void* workspace, int stats_ready, void* s) {
	CB_REQUIRE_DEVICE();
	NormGeom g;
	int rc = fill_geom(d, g); if (rc) return rc;
	CB_ARG(workspace != nullptr);
	cudaStream_t st = as_stream(s);
	double* ws = (double*)workspace;
	// algorithmic bytes: read x (stats, unless the producing convolution has left them in the workspace) + read x + write y
	prof_begin(PROF_NORM, (stats_ready ? 2.0 : 3.0) * g.batch * g.hw * (double)g.c * cb200_dtype_size(d->dtype), st);
	if (!stats_ready) CB_CUDA(cudaMemsetAsync(ws, 0, cb200_norm_workspace_bytes(d), st));
	if (norm_pipeline_on() && !stats_ready) {
		const int k = chunk_samples(g.batch, (double)g.hw * g.cp * cb200_dtype_size(d->dtype));
		g.ppb = chunk_ppb(g.hw, k, g.cp >> 3);
		const int nbx = ceil_div(g.hw, g.ppb);
		ChunkGeom cg; unsigned grid;
		for (int i = 0; chunk_launch_geom(i, g.batch, k, nbx, nbx, cg, grid); i++) {
			CB_DISPATCH_DTYPE(d->dtype, T, (norm_fwd_chunk_kernel<T><<<grid, NORM_THREADS, sizeof(float) * 2 * g.nb_group, st>>>(
				(const T*)x, (T*)y, gamma, beta, mean, var, ws, g, cg)));
			CB_LAUNCH_CHECK();
		}
		prof_end(st);
		return CB200_OK;
	}
	dim3 grid((unsigned)ceil_div(g.hw, g.ppb), (unsigned)g.batch);
	size_t smem = sizeof(float) * 2 * g.nb_group;
	if (!stats_ready) {
		CB_DISPATCH_DTYPE(d->dtype, T, (norm_stats_kernel<T, false><<<grid, NORM_THREADS, smem, st>>>((const T*)x, nullptr, ws, g)));
		CB_LAUNCH_CHECK();
	}
	// finalize folded into the apply blocks (each turns its sample's FP64 sums into mean / var in shared memory): one tiny
	// launch less on the critical path per layer and pass
	unsigned agrid;
	const ChunkGeom cg = apply_only_geom(g.batch, (int)grid.x, agrid);
	CB_DISPATCH_DTYPE(d->dtype, T, (norm_fwd_chunk_kernel<T><<<agrid, NORM_THREADS, smem, st>>>((const T*)x, (T*)y, gamma, beta, mean, var, ws, g, cg)));
	CB_LAUNCH_CHECK();
	prof_end(st);
	return CB200_OK;
}

int cb200_norm_backward(const cb200_norm_desc* d, const void* x, const void* dy, void* dx, const float* gamma,
                        const float* mean, const float* var, float* d_gamma, float* d_beta,
                        const cb200_activ* prev_activ, float* dx_colsum, void* workspace, void* s) {
	CB_REQUIRE_DEVICE();
	NormGeom g;
	int rc = fill_geom(d, g); if (rc) return rc;
	CB_ARG(workspace != nullptr);
	cudaStream_t st = as_stream(s);
	double* ws = (double*)workspace;
	cb200_activ pa; pa.type = CB200_LINEAR; pa.leak = 0; pa.saturation = 0; pa.beta = 0;
	if (prev_activ) pa = *prev_activ;
	// algorithmic bytes: (dy, x) for the reductions + (dy, x) + write dx = 5 passes
	prof_begin(PROF_NORM, 5.0 * g.batch * g.hw * (double)g.c * cb200_dtype_size(d->dtype), st);
	CB_CUDA(cudaMemsetAsync(ws, 0, cb200_norm_workspace_bytes(d), st));
	if (norm_pipeline_on()) {
		if (dx_colsum != nullptr) CB_CUDA(cudaMemsetAsync(dx_colsum, 0, sizeof(float) * g.c, st));
		const int k = chunk_samples(g.batch, 2.0 * g.hw * g.cp * cb200_dtype_size(d->dtype));
		g.ppb = chunk_ppb(g.hw, k, g.cp >> 3);
		const int nbx = ceil_div(g.hw, g.ppb);
		const size_t smem_c = sizeof(float) * (2 * g.nb_group + (dx_colsum ? g.cp : 0));
		ChunkGeom cg; unsigned grid;
		for (int i = 0; chunk_launch_geom(i, g.batch, k, nbx, nbx, cg, grid); i++) {
			CB_DISPATCH_DTYPE(d->dtype, T, (norm_bwd_chunk_kernel<T><<<grid, NORM_THREADS, smem_c, st>>>(
				(const T*)x, (const T*)dy, (T*)dx, gamma, mean, var, d_gamma, d_beta, pa, dx_colsum, ws, g, cg)));
			CB_LAUNCH_CHECK();
		}
		prof_end(st);
		return CB200_OK;
	}
	dim3 grid((unsigned)ceil_div(g.hw, g.ppb), (unsigned)g.batch);
	size_t smem = sizeof(float) * 2 * g.nb_group;
	CB_DISPATCH_DTYPE(d->dtype, T, (norm_stats_kernel<T, true><<<grid, NORM_THREADS, smem, st>>>((const T*)dy, (const T*)x, ws, g)));
	CB_LAUNCH_CHECK();
	if (dx_colsum != nullptr) CB_CUDA(cudaMemsetAsync(dx_colsum, 0, sizeof(float) * g.c, st));
	unsigned agrid;
	const ChunkGeom cg = apply_only_geom(g.batch, (int)grid.x, agrid);
	CB_DISPATCH_DTYPE(d->dtype, T, (norm_bwd_chunk_kernel<T><<<agrid, NORM_THREADS, sizeof(float) * (2 * g.nb_group + (dx_colsum ? g.cp : 0)), st>>>(
		(const T*)x, (const T*)dy, (T*)dx, gamma, mean, var, d_gamma, d_beta, pa, dx_colsum, ws, g, cg)));
	CB_LAUNCH_CHECK();
	prof_end(st);
	return CB200_OK;
}

int cb200_norm_pool_fusable(const cb200_norm_desc* nd, const cb200_pool_desc* pd) {
	if (nd == nullptr || pd == nullptr) return 0;
	return pd->pool_type == CB200_POOL_MAX && pd->p_h == 2 && pd->p_w == 2 && pd->stride_h == 2 && pd->stride_w == 2 &&
	       pd->pad_h == 0 && pd->pad_w == 0 && pd->activ.type == CB200_LINEAR && nd->c == pd->c && nd->h == pd->in_h &&
	       nd->w == pd->in_w && (pd->in_h & 1) == 0 && (pd->in_w & 1) == 0 && pd->out_h * 2 == pd->in_h &&
	       pd->out_w * 2 == pd->in_w && nd->dtype == pd->dtype && nd->batch == pd->batch;
}

static int fill_fused(const cb200_norm_desc* nd, const cb200_pool_desc* pd, FusedGeom& f) {
	int rc = fill_geom(nd, f.n); if (rc) return rc;
	if (!cb200_norm_pool_fusable(nd, pd)) { set_error("cb200_norm_pool_*: this norm / pool pair cannot be fused (see cb200_norm_pool_fusable)"); return CB200_ERR_ARG; }
	f.in_w = pd->in_w; f.out_w = pd->out_w; f.out_hw = pd->out_h * pd->out_w;
	f.ppb_out = norm_pix_per_block(f.out_hw, f.n.batch, f.n.cp >> 3);
	return CB200_OK;
}

int cb200_norm_pool_forward(const cb200_norm_desc* nd, const cb200_pool_desc* pd, const void* x, void* pooled, uint8_t* pool_map,
                            const float* gamma, const float* beta, float* mean, float* var, void* workspace, void* s) {
	return cb200_norm_pool_forward_ex(nd, pd, x, pooled, pool_map, gamma, beta, mean, var, workspace, 0, s);
}

int cb200_norm_pool_forward_ex(const cb200_norm_desc* nd, const cb200_pool_desc* pd, const void* x, void* pooled, uint8_t* pool_map,
                               const float* gamma, const float* beta, float* mean, float* var, void* workspace, int stats_ready, void* s) {
	CB_REQUIRE_DEVICE();
	FusedGeom f;
	int rc = fill_fused(nd, pd, f); if (rc) return rc;
	CB_ARG(workspace != nullptr);
	const NormGeom& g = f.n;
	cudaStream_t st = as_stream(s);
	double* ws = (double*)workspace;
	const double es = (double)cb200_dtype_size(nd->dtype), E = (double)g.batch * g.hw * g.c;
	// algorithmic bytes: read x (statistics, unless the producing convolution has left them in the workspace) + read x +
	// write pooled values and their 1-byte window index
	prof_begin(PROF_NORM, (stats_ready ? 1.0 : 2.0) * E * es + 0.25 * E * (es + 1.0), st);
	if (!stats_ready) CB_CUDA(cudaMemsetAsync(ws, 0, cb200_norm_workspace_bytes(nd), st));
	if (norm_pipeline_on() && !stats_ready) {
		const int k = chunk_samples(g.batch, (double)g.hw * g.cp * es);
		f.n.ppb = chunk_ppb(g.hw, k, g.cp >> 3);
		f.ppb_out = chunk_ppb(f.out_hw, k, g.cp >> 3);
		ChunkGeom cg; unsigned grid;
		for (int i = 0; chunk_launch_geom(i, g.batch, k, ceil_div(g.hw, g.ppb), ceil_div(f.out_hw, f.ppb_out), cg, grid); i++) {
			CB_DISPATCH_DTYPE(nd->dtype, T, (norm_pool_fwd_chunk_kernel<T><<<grid, NORM_THREADS, sizeof(float) * 2 * g.nb_group, st>>>(
				(const T*)x, (T*)pooled, pool_map, gamma, beta, mean, var, ws, f, cg)));
			CB_LAUNCH_CHECK();
		}
		prof_end(st);
		return CB200_OK;
	}
	dim3 grid((unsigned)ceil_div(g.hw, g.ppb), (unsigned)g.batch);
	if (!stats_ready) {
		CB_DISPATCH_DTYPE(nd->dtype, T, (norm_stats_kernel<T, false><<<grid, NORM_THREADS, sizeof(float) * 2 * g.nb_group, st>>>((const T*)x, nullptr, ws, g)));
		CB_LAUNCH_CHECK();
	}
	unsigned agrid;
	const ChunkGeom cg = apply_only_geom(g.batch, ceil_div(f.out_hw, f.ppb_out), agrid);
	CB_DISPATCH_DTYPE(nd->dtype, T, (norm_pool_fwd_chunk_kernel<T><<<agrid, NORM_THREADS, sizeof(float) * 2 * g.nb_group, st>>>(
		(const T*)x, (T*)pooled, pool_map, gamma, beta, mean, var, ws, f, cg)));
	CB_LAUNCH_CHECK();
	prof_end(st);
	return CB200_OK;
}

int cb200_norm_pool_backward(const cb200_norm_desc* nd, const cb200_pool_desc* pd, const void* x, const void* d_pooled,
                             const uint8_t* pool_map, void* dx, const float* gamma, const float* mean, const float* var,
                             float* d_gamma, float* d_beta, const cb200_activ* prev_activ, float* dx_colsum, void* workspace, void* s) {
	return cb200_norm_pool_backward_ex(nd, pd, x, d_pooled, pool_map, dx, gamma, mean, var, d_gamma, d_beta, prev_activ, dx_colsum,
	                                   workspace, nullptr, nullptr, s);
}

int cb200_norm_pool_backward_ex(const cb200_norm_desc* nd, const cb200_pool_desc* pd, const void* x, const void* d_pooled,
                                const uint8_t* pool_map, void* dx, const float* gamma, const float* mean, const float* var,
                                float* d_gamma, float* d_beta, const cb200_activ* prev_activ, float* dx_colsum, void* workspace,
                                const void* pooled, const float* beta, void* s) {
	CB_REQUIRE_DEVICE();
	FusedGeom f;
	int rc = fill_fused(nd, pd, f); if (rc) return rc;
	CB_ARG(workspace != nullptr && pool_map != nullptr);
	const NormGeom& g = f.n;
	cudaStream_t st = as_stream(s);
	double* ws = (double*)workspace;
	cb200_activ pa; pa.type = CB200_LINEAR; pa.leak = 0; pa.saturation = 0; pa.beta = 0;
	if (prev_activ) pa = *prev_activ;
	const double es = (double)cb200_dtype_size(nd->dtype), E = (double)g.batch * g.hw * g.c;
	// algorithmic bytes: the pooled delta + pooled output for the reductions (x + pooled delta + map without `pooled`), then
	// x + dx + the pooled delta and its map for the apply
	prof_begin(PROF_NORM, (pooled != nullptr && beta != nullptr ? 2.0 * E * es + 0.5 * E * es : 3.0 * E * es + 0.25 * E * (es + 1.0)) + 0.25 * E * (es + 1.0), st);
	CB_CUDA(cudaMemsetAsync(ws, 0, cb200_norm_workspace_bytes(nd), st));
	if (norm_pipeline_on()) {
		if (dx_colsum != nullptr) CB_CUDA(cudaMemsetAsync(dx_colsum, 0, sizeof(float) * g.c, st));
		const int k = chunk_samples(g.batch, (double)g.hw * g.cp * es + 0.25 * g.hw * g.cp * (es + 1.0));
		f.ppb_out = chunk_ppb(f.out_hw, k, g.cp >> 3);
		const int nbx = ceil_div(f.out_hw, f.ppb_out);
		const size_t smem_c = sizeof(float) * (2 * g.nb_group + (dx_colsum ? g.cp : 0));
		ChunkGeom cg; unsigned grid;
		for (int i = 0; chunk_launch_geom(i, g.batch, k, nbx, nbx, cg, grid); i++) {
			CB_DISPATCH_DTYPE(nd->dtype, T, (norm_pool_bwd_chunk_kernel<T><<<grid, NORM_THREADS, smem_c, st>>>(
				(const T*)x, (const T*)d_pooled, pool_map, (T*)dx, gamma, mean, var, d_gamma, d_beta, pa, dx_colsum, ws, f, cg)));
			CB_LAUNCH_CHECK();
		}
		prof_end(st);
		return CB200_OK;
	}
	dim3 grid_o((unsigned)ceil_div(f.out_hw, f.ppb_out), (unsigned)g.batch);
	static const bool stats_from_y = env_int("CB200_GN_POOL_STATS_Y", 1) != 0;
	if (pooled != nullptr && beta != nullptr && stats_from_y) {
		// reductions over the pooled delta and the pooled OUTPUT only (see norm_pool_bwd_stats_y_body)
		CB_DISPATCH_DTYPE(nd->dtype, T, (norm_pool_bwd_stats_y_kernel<T><<<grid_o, NORM_THREADS, sizeof(float) * 2 * g.nb_group, st>>>(
			(const T*)x, (const T*)d_pooled, pool_map, (const T*)pooled, gamma, beta, mean, var, ws, f)));
	} else
		CB_DISPATCH_DTYPE(nd->dtype, T, (norm_pool_bwd_stats_kernel<T><<<grid_o, NORM_THREADS, sizeof(float) * 2 * g.nb_group, st>>>(
			(const T*)x, (const T*)d_pooled, pool_map, ws, f)));
	CB_LAUNCH_CHECK();
	if (dx_colsum != nullptr) CB_CUDA(cudaMemsetAsync(dx_colsum, 0, sizeof(float) * g.c, st));
	unsigned agrid;
	const ChunkGeom cg = apply_only_geom(g.batch, (int)grid_o.x, agrid);
	CB_DISPATCH_DTYPE(nd->dtype, T, (norm_pool_bwd_chunk_kernel<T><<<agrid, NORM_THREADS, sizeof(float) * (2 * g.nb_group + (dx_colsum ? g.cp : 0)), st>>>(
		(const T*)x, (const T*)d_pooled, pool_map, (T*)dx, gamma, mean, var, d_gamma, d_beta, pa, dx_colsum, ws, f, cg)));
	CB_LAUNCH_CHECK();
	prof_end(st);
	return CB200_OK;
}

int cb200_norm_reduce_grads(const cb200_norm_desc* d, const float* d_gamma, const float* d_beta, float* gsum, void* s) {
	CB_REQUIRE_DEVICE();
	norm_reduce_grads_kernel<<<ceil_div(d->nb_group * 32, 128), 128, 0, as_stream(s)>>>(d_gamma, d_beta, gsum, d->batch, d->nb_group);
	CB_LAUNCH_CHECK();
	return CB200_OK;
}

int cb200_norm_reduce_update(const cb200_norm_desc* d, const float* d_gamma, const float* d_beta, float* gsum, float* gamma, float* beta,
                             float* gamma_upd, float* beta_upd, const float* hyper, void* s) {
	CB_REQUIRE_DEVICE();
	norm_reduce_update_kernel<<<ceil_div(d->nb_group * 32, 128), 128, 0, as_stream(s)>>>(d_gamma, d_beta, gsum, gamma, beta, gamma_upd, beta_upd,
	                                                                                    hyper, d->batch, d->nb_group, d->set_off);
	CB_LAUNCH_CHECK();
	return CB200_OK;
}

int cb200_norm_update(const cb200_norm_desc* d, float* gamma, float* beta, float* gamma_upd, float* beta_upd,
                      const float* gsum, const float* hyper, void* s) {
	CB_REQUIRE_DEVICE();
	norm_update_kernel<<<ceil_div(d->nb_group, 128), 128, 0, as_stream(s)>>>(gamma, beta, gamma_upd, beta_upd, gsum, hyper, d->nb_group, d->set_off);
	CB_LAUNCH_CHECK();
	return CB200_OK;
}

}  // extern "C"
