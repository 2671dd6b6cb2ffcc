// norm.cu - group normalisation forward / backward / parameter update on channels-last tensors.
//
// Statistics are per (sample, group of `group_size` channels) exactly as upstream
// (src/cuda/cuda_norm_layer.cu:37-46 index map, :65-178 reductions, :180-265 apply kernels;
// CPU twin src/naiv/naiv_norm_layer.c).  Differences in mechanism, not in maths:
//   - one streaming pass produces sum and sum-of-squares (FP32 per thread over a short run,
//     FP64 across threads/blocks), instead of a mean pass followed by a variance pass;
//   - gamma/beta and their momentum live on the device; the update is a kernel, not a host
//     loop behind four blocking memcpys (cuda_norm_layer.cu:434-457).
#include "common.cuh"

namespace cb200 {

struct NormGeom {
	int batch, length, c, cp, hw, group_size, nb_group, set_off;
	float eps;
};

constexpr int NORM_THREADS = 256;
constexpr int NORM_PIX_PER_BLOCK = 1024;

// Accumulate, for every channel vector owned by the thread, sum(a) and sum(a*b) over the block's
// pixel range; b == a gives (sum x, sum x^2), b == x with a == delta gives (sum d, sum d*x).
// ws layout: double [batch][nb_group][2].
template <typename T, bool TWO_INPUTS>
__global__ void __launch_bounds__(NORM_THREADS)
norm_stats_kernel(const T* __restrict__ a_in, const T* __restrict__ b_in, double* __restrict__ ws, NormGeom g) {
	extern __shared__ double sm_acc[];   // [nb_group][2]
	const int cv = g.cp >> 3;
	const int b = blockIdx.y;
	for (int i = threadIdx.x; i < g.nb_group * 2; i += blockDim.x) sm_acc[i] = 0.0;
	__syncthreads();

	const int lanes_c = cv < NORM_THREADS ? cv : NORM_THREADS;
	const int lanes_p = NORM_THREADS / lanes_c;
	const int lane_c = threadIdx.x % lanes_c, lane_p = threadIdx.x / lanes_c;
	const int p0 = blockIdx.x * NORM_PIX_PER_BLOCK;
	int p1 = p0 + NORM_PIX_PER_BLOCK;
	if (p1 > g.hw) p1 = g.hw;

	if (lane_p < lanes_p) {
		for (int v = lane_c; v < cv; v += lanes_c) {
			float s0[8], s1[8];
#pragma unroll
			for (int j = 0; j < 8; j++) { s0[j] = 0.0f; s1[j] = 0.0f; }
			for (int p = p0 + lane_p; p < p1; p += lanes_p) {
				const long long o = ((long long)b * g.hw + p) * g.cp + v * 8;
				float av[8], bv[8];
				load8<T>(a_in + o, av);
				if (TWO_INPUTS) load8<T>(b_in + o, bv);
#pragma unroll
				for (int j = 0; j < 8; j++) {
					s0[j] += av[j];
					s1[j] += av[j] * (TWO_INPUTS ? bv[j] : av[j]);
				}
			}
			// fold the 8 channels into their groups (a vector spans one group when group_size % 8 == 0)
			int cur = -1;
			double g0 = 0.0, g1 = 0.0;
#pragma unroll
			for (int j = 0; j < 8; j++) {
				const int ch = v * 8 + j;
				if (ch >= g.c) break;
				const int grp = ch / g.group_size;
				if (grp != cur) {
					if (cur >= 0 && cur < g.nb_group) { atomicAdd(&sm_acc[cur * 2], g0); atomicAdd(&sm_acc[cur * 2 + 1], g1); }
					cur = grp; g0 = 0.0; g1 = 0.0;
				}
				g0 += (double)s0[j];
				g1 += (double)s1[j];
			}
			if (cur >= 0 && cur < g.nb_group) { atomicAdd(&sm_acc[cur * 2], g0); atomicAdd(&sm_acc[cur * 2 + 1], g1); }
		}
	}
	__syncthreads();
	for (int i = threadIdx.x; i < g.nb_group * 2; i += blockDim.x)
		atomicAdd(&ws[(size_t)b * g.nb_group * 2 + i], sm_acc[i]);
}

// mean = S1/n ; var = S2/n - mean^2  (biased variance, as upstream)
__global__ void norm_finalize_fwd_kernel(const double* __restrict__ ws, float* __restrict__ mean, float* __restrict__ var, NormGeom g) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= g.batch * g.nb_group) return;
	const double n = (double)g.group_size * g.hw;
	const double m = ws[2 * i] / n;
	double v = ws[2 * i + 1] / n - m * m;
	if (v < 0.0) v = 0.0;
	mean[i] = (float)m;
	var[i] = (float)v;
}

// d_beta = sum d ; d_gamma = (sum d*x - mean*sum d) / sqrt(var+eps)
__global__ void norm_finalize_bwd_kernel(const double* __restrict__ ws, const float* __restrict__ mean, const float* __restrict__ var,
                                         float* __restrict__ d_gamma, float* __restrict__ d_beta, NormGeom g) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= g.batch * g.nb_group) return;
	const double sd = ws[2 * i], sdx = ws[2 * i + 1];
	d_beta[i] = (float)sd;
	d_gamma[i] = (float)((sdx - (double)mean[i] * sd) / sqrt((double)var[i] + (double)g.eps));
}

template <typename T>
__global__ void __launch_bounds__(256)
norm_apply_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ gamma, const float* __restrict__ beta,
                  const float* __restrict__ mean, const float* __restrict__ var, NormGeom g) {
	const int cv = g.cp >> 3;
	const long long total = (long long)g.batch * g.hw * cv;
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
		const int v = (int)(i % cv);
		const long long pix = i / cv;
		const int b = (int)(pix / g.hw);
		float xv[8], out[8];
		load8<T>(x + pix * g.cp + v * 8, xv);
		if (b >= g.length) {
#pragma unroll
			for (int j = 0; j < 8; j++) out[j] = 0.0f;
		} else {
			int cur = -1;
			float sc = 1.0f, sh = 0.0f;
#pragma unroll
			for (int j = 0; j < 8; j++) {
				const int ch = v * 8 + j;
				if (ch >= g.c) { out[j] = 0.0f; continue; }
				const int grp = ch / g.group_size;
				if (grp != cur) {
					cur = grp;
					if (grp < g.nb_group - g.set_off) {
						const float rstd = 1.0f / sqrtf(var[b * g.nb_group + grp] + g.eps);
						sc = gamma[grp] * rstd;
						sh = beta[grp] - mean[b * g.nb_group + grp] * sc;
					} else { sc = 1.0f; sh = 0.0f; }
				}
				out[j] = xv[j] * sc + sh;
			}
		}
		store8<T>(y + pix * g.cp + v * 8, out);
	}
}

// dx = gamma*rstd/n * (n*d - d_beta - xhat*d_gamma), then the previous layer's deriv hook on x
template <typename T>
__global__ void __launch_bounds__(256)
norm_bwd_apply_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx, const float* __restrict__ gamma,
                      const float* __restrict__ mean, const float* __restrict__ var,
                      const float* __restrict__ d_gamma, const float* __restrict__ d_beta,
                      cb200_activ prev_activ, NormGeom g) {
	const int cv = g.cp >> 3;
	const long long total = (long long)g.batch * g.hw * cv;
	const float n = (float)(g.group_size * g.hw);
	const float inv_n = 1.0f / n;
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
		const int v = (int)(i % cv);
		const long long pix = i / cv;
		const int b = (int)(pix / g.hw);
		float out[8];
		if (b >= g.length) {
#pragma unroll
			for (int j = 0; j < 8; j++) out[j] = 0.0f;
		} else {
			float xv[8], dv[8];
			load8<T>(x + pix * g.cp + v * 8, xv);
			load8<T>(dy + pix * g.cp + v * 8, dv);
			int cur = -1;
			bool active = false;
			float mu = 0.0f, rstd = 0.0f, gm = 0.0f, dg = 0.0f, db = 0.0f;
#pragma unroll
			for (int j = 0; j < 8; j++) {
				const int ch = v * 8 + j;
				if (ch >= g.c) { out[j] = 0.0f; continue; }
				const int grp = ch / g.group_size;
				if (grp != cur) {
					cur = grp;
					active = grp < g.nb_group - g.set_off;
					if (active) {
						const int s = b * g.nb_group + grp;
						mu = mean[s]; rstd = 1.0f / sqrtf(var[s] + g.eps); gm = gamma[grp]; dg = d_gamma[s]; db = d_beta[s];
					}
				}
				float r = dv[j];
				if (active) r = inv_n * gm * rstd * (n * dv[j] - db - (xv[j] - mu) * rstd * dg);
				out[j] = activ_deriv_mul(prev_activ, r, xv[j]);
			}
		}
		store8<T>(dx + pix * g.cp + v * 8, out);
	}
}

// gsum[0][g] = sum_b d_gamma[b][g], gsum[1][g] = sum_b d_beta[b][g]
__global__ void norm_reduce_grads_kernel(const float* __restrict__ d_gamma, const float* __restrict__ d_beta,
                                         float* __restrict__ gsum, int batch, int nb_group) {
	const int grp = blockIdx.x * blockDim.x + threadIdx.x;
	if (grp >= nb_group) return;
	double sg = 0.0, sb = 0.0;
	for (int b = 0; b < batch; b++) { sg += d_gamma[b * nb_group + grp]; sb += d_beta[b * nb_group + grp]; }
	gsum[grp] = (float)sg;
	gsum[nb_group + grp] = (float)sb;
}

// upd = mom*upd + lr*(sum/B) ; param -= upd/S     (hyper[0] = lr/B_total, hyper[3] = S)
__global__ void norm_update_kernel(float* __restrict__ gamma, float* __restrict__ beta, float* __restrict__ gamma_upd,
                                   float* __restrict__ beta_upd, const float* __restrict__ gsum,
                                   const float* __restrict__ hyper, int nb_group, int set_off) {
	const int grp = blockIdx.x * blockDim.x + threadIdx.x;
	if (grp >= nb_group - set_off) return;
	const float alpha = hyper[0], mom = hyper[1], S = hyper[3];
	const float gu = mom * gamma_upd[grp] + alpha * gsum[grp];
	const float bu = mom * beta_upd[grp] + alpha * gsum[nb_group + grp];
	gamma_upd[grp] = gu;
	beta_upd[grp] = bu;
	gamma[grp] -= gu / S;
	beta[grp] -= bu / S;
}

static int fill_geom(const cb200_norm_desc* d, NormGeom& g) {
	CB_ARG(d != nullptr && d->batch > 0 && d->c > 0 && d->group_size > 0 && d->nb_group > 0);
	CB_ARG(d->nb_group * d->group_size >= d->c);
	g.batch = d->batch; g.length = d->length; g.c = d->c; g.cp = round8(d->c); g.hw = d->h * d->w;
	g.group_size = d->group_size; g.nb_group = d->nb_group; g.set_off = d->set_off; g.eps = d->eps;
	return CB200_OK;
}
}  // namespace cb200
using namespace cb200;

extern "C" {

size_t cb200_norm_workspace_bytes(const cb200_norm_desc* d) { return sizeof(double) * 2 * (size_t)d->batch * d->nb_group; }

int cb200_norm_forward(const cb200_norm_desc* d, const void* x, void* y, const float* gamma, const float* beta,
                       float* mean, float* var, void* workspace, void* s) {
	CB_REQUIRE_DEVICE();
	NormGeom g;
	int rc = fill_geom(d, g); if (rc) return rc;
	CB_ARG(workspace != nullptr);
	cudaStream_t st = as_stream(s);
	double* ws = (double*)workspace;
	// algorithmic bytes: read x (stats) + read x + write y = 3 passes over the real elements
	prof_begin(PROF_NORM, 3.0 * g.batch * g.hw * (double)g.c * cb200_dtype_size(d->dtype), st);
	CB_CUDA(cudaMemsetAsync(ws, 0, cb200_norm_workspace_bytes(d), st));
	dim3 grid((unsigned)ceil_div(g.hw, NORM_PIX_PER_BLOCK), (unsigned)g.batch);
	size_t smem = sizeof(double) * 2 * g.nb_group;
	CB_DISPATCH_DTYPE(d->dtype, T, (norm_stats_kernel<T, false><<<grid, NORM_THREADS, smem, st>>>((const T*)x, nullptr, ws, g)));
	CB_LAUNCH_CHECK();
	norm_finalize_fwd_kernel<<<ceil_div(g.batch * g.nb_group, 128), 128, 0, st>>>(ws, mean, var, g);
	CB_LAUNCH_CHECK();
	long long total = (long long)g.batch * g.hw * (g.cp >> 3);
	CB_DISPATCH_DTYPE(d->dtype, T, (norm_apply_kernel<T><<<grid_for(total, 256), 256, 0, st>>>((const T*)x, (T*)y, gamma, beta, mean, var, g)));
	CB_LAUNCH_CHECK();
	prof_end(st);
	return CB200_OK;
}

int cb200_norm_backward(const cb200_norm_desc* d, const void* x, const void* dy, void* dx, const float* gamma,
                        const float* mean, const float* var, float* d_gamma, float* d_beta,
                        const cb200_activ* prev_activ, void* workspace, void* s) {
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
	dim3 grid((unsigned)ceil_div(g.hw, NORM_PIX_PER_BLOCK), (unsigned)g.batch);
	size_t smem = sizeof(double) * 2 * g.nb_group;
	CB_DISPATCH_DTYPE(d->dtype, T, (norm_stats_kernel<T, true><<<grid, NORM_THREADS, smem, st>>>((const T*)dy, (const T*)x, ws, g)));
	CB_LAUNCH_CHECK();
	norm_finalize_bwd_kernel<<<ceil_div(g.batch * g.nb_group, 128), 128, 0, st>>>(ws, mean, var, d_gamma, d_beta, g);
	CB_LAUNCH_CHECK();
	long long total = (long long)g.batch * g.hw * (g.cp >> 3);
	CB_DISPATCH_DTYPE(d->dtype, T, (norm_bwd_apply_kernel<T><<<grid_for(total, 256), 256, 0, st>>>(
		(const T*)x, (const T*)dy, (T*)dx, gamma, mean, var, d_gamma, d_beta, pa, g)));
	CB_LAUNCH_CHECK();
	prof_end(st);
	return CB200_OK;
}

int cb200_norm_reduce_grads(const cb200_norm_desc* d, const float* d_gamma, const float* d_beta, float* gsum, void* s) {
	CB_REQUIRE_DEVICE();
	norm_reduce_grads_kernel<<<ceil_div(d->nb_group, 128), 128, 0, as_stream(s)>>>(d_gamma, d_beta, gsum, d->batch, d->nb_group);
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
