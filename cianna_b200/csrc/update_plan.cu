// update_plan.cu - the optimizer sweep of a whole network in THREE launches.
//
// Upstream updates layer by layer (cuda_update_weights + cuda_master_weight_copy per conv / dense layer,
// src/cuda/cuda_conv_layer.cu:559-562, src/cuda/cuda_main.cu:486-509; the group-norm parameters on the host,
// src/cuda/cuda_norm_layer.cu:437-459).  Layer-by-layer launches of the per-layer kernels in conv.cu / norm.cu cost
// Darknet19 61 launches per step, most of them a few microseconds of work behind a launch gap, at 0.13 of the HBM
// roofline (profiles/r1_step_metrics_b128_summary.txt).  A plan is a device table with one entry per layer; each of
// the three kernels below covers every layer of its kind at once, a block (or warp) finding its layer by binary search
// in the table's prefix sums:
//   1  conv_update_rows_many    one block per FILTER of any conv layer: master + momentum rows staged in shared memory,
//                               SGD + momentum + weight decay, FP32 master -> 16-bit forward operand  (= conv_update_rows_kernel)
//   2  conv_wbwd_transpose_many one block per 32x32 tile of any layer's operand: forward operand -> rotated, transposed
//                               data-gradient operand                                                  (= conv_wbwd_transpose_kernel)
//   3  norm_update_many         one warp per group of any group-norm layer: batch sum of the per-sample gradients
//                               (or the all-reduced sums) -> gamma / beta update                       (= norm_reduce_update / norm_update)
// Arithmetic is the per-layer kernels' own, statement for statement: results are bit-identical (tests/test_gpu_update_plan.py).
#include "common.cuh"

namespace cb200 {

struct ConvUpdItem {
	float* master; float* moment; const float* grad; const float* grad_b;
	void* w_fwd; void* w_bwd; float* bias_w;
	float bias_value;
	int taps, in_c, in_cp, out_c, out_cp, wb_dense;
	int first_block;      // of kernel 1 (one block per filter)
	int first_tile;       // of kernel 2
	int tiles_c, tiles_f; // 32-wide tiles over input channels / filters
};

struct NormUpdItem {
	const float* d_gamma; const float* d_beta; float* gsum;
	float* gamma; float* beta; float* gamma_upd; float* beta_upd;
	int batch, nb_group, set_off, reduce;
	int first_group;
};

struct UpdatePlan {
	int dtype;
	int n_conv, n_norm;
	int conv_blocks, conv_tiles, norm_groups;
	size_t smem;
	ConvUpdItem* conv_dev;
	NormUpdItem* norm_dev;
};

// index of the last item whose `first` is <= v
template <typename Item, int Item::*First>
__device__ __forceinline__ int find_item(const Item* __restrict__ items, int n, int v) {
	int lo = 0, hi = n - 1;
	while (lo < hi) {
		const int mid = (lo + hi + 1) >> 1;
		if (items[mid].*First <= v) lo = mid; else hi = mid - 1;
	}
	return lo;
}

__device__ __forceinline__ size_t plan_wbwd_row(int c, int tap, int taps, int in_cp, int wb_dense) {
	// (conv.cu: wbwd_row) rows (c, rotated tap), or (tap, c) for a whole-map filter
	return wb_dense ? (size_t)tap * in_cp + c : (size_t)c * taps + (taps - 1 - tap);
}

template <typename T>
__global__ void __launch_bounds__(256)
conv_update_rows_many_kernel(const ConvUpdItem* __restrict__ items, int n, const float* __restrict__ hyper) {
	extern __shared__ float rows[];                 // [2][kref]: master row, momentum row
	const ConvUpdItem it = items[find_item<ConvUpdItem, &ConvUpdItem::first_block>(items, n, (int)blockIdx.x)];
	const int f = (int)blockIdx.x - it.first_block;
	const int taps = it.taps, in_c = it.in_c, in_cp = it.in_cp;
	const int kref = taps * in_c + 1;
	float* sm_w = rows;
	float* sm_m = rows + kref;
	float* __restrict__ master = it.master;
	float* __restrict__ moment = it.moment;
	const float alpha = hyper[0], mom = hyper[1], wdlr = hyper[2], S = hyper[3];
	for (int i = threadIdx.x; i < kref; i += blockDim.x) {
		sm_w[i] = master[(size_t)f * kref + i];
		sm_m[i] = moment[(size_t)f * kref + i];
	}
	__syncthreads();
	const int n_op = taps * in_cp;
	const float* __restrict__ g = it.grad + (size_t)f * n_op;
	T* __restrict__ wf = (T*)it.w_fwd + (size_t)f * n_op;
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
		sgd_momentum_step(alpha, mom, wdlr, S, __fmul_rn(it.bias_value, it.grad_b[f]), m, wv);
		sm_m[mi] = m;
		sm_w[mi] = wv;
		it.bias_w[f] = wv;
	}
	__syncthreads();
	for (int i = threadIdx.x; i < kref; i += blockDim.x) {
		master[(size_t)f * kref + i] = sm_w[i];
		moment[(size_t)f * kref + i] = sm_m[i];
	}
}

template <typename T>
__global__ void __launch_bounds__(256)
conv_wbwd_transpose_many_kernel(const ConvUpdItem* __restrict__ items, int n) {
	__shared__ T tile[32][33];
	const ConvUpdItem it = items[find_item<ConvUpdItem, &ConvUpdItem::first_tile>(items, n, (int)blockIdx.x)];
	int t = (int)blockIdx.x - it.first_tile;
	const int tc = t % it.tiles_c; t /= it.tiles_c;
	const int tf = t % it.tiles_f;
	const int tap = t / it.tiles_f;
	const int c0 = tc * 32, f0 = tf * 32;
	const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
	const T* __restrict__ w_fwd = (const T*)it.w_fwd;
	T* __restrict__ w_bwd = (T*)it.w_bwd;
#pragma unroll
	for (int k = 0; k < 4; k++) {
		const int f = f0 + ty + 8 * k, c = c0 + tx;
		tile[ty + 8 * k][tx] = (f < it.out_c && c < it.in_cp) ? w_fwd[((size_t)f * it.taps + tap) * it.in_cp + c] : from_f32<T>(0.0f);
	}
	__syncthreads();
#pragma unroll
	for (int k = 0; k < 4; k++) {
		const int c = c0 + ty + 8 * k, f = f0 + tx;
		if (c < it.in_c && f < it.out_cp) w_bwd[plan_wbwd_row(c, tap, it.taps, it.in_cp, it.wb_dense) * it.out_cp + f] = tile[tx][ty + 8 * k];
	}
}

__global__ void __launch_bounds__(128)
norm_update_many_kernel(const NormUpdItem* __restrict__ items, int n, int total_groups, const float* __restrict__ hyper) {
	const int gw = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
	const int lane = threadIdx.x & 31;
	if (gw >= total_groups) return;
	const NormUpdItem it = items[find_item<NormUpdItem, &NormUpdItem::first_group>(items, n, gw)];
	const int grp = gw - it.first_group;
	float fg, fb;
	if (it.reduce) {
		double sg = 0.0, sb = 0.0;
		for (int b = lane; b < it.batch; b += 32) { sg += it.d_gamma[b * it.nb_group + grp]; sb += it.d_beta[b * it.nb_group + grp]; }
		sg = warp_sum(sg);
		sb = warp_sum(sb);
		if (lane != 0) return;
		fg = (float)sg; fb = (float)sb;
		it.gsum[grp] = fg; it.gsum[it.nb_group + grp] = fb;
	} else {
		if (lane != 0) return;
		fg = it.gsum[grp]; fb = it.gsum[it.nb_group + grp];
	}
	if (grp >= it.nb_group - it.set_off) return;
	const float alpha = hyper[0], mom = hyper[1], S = hyper[3];
	norm_param_step(alpha, mom, S, fg, it.gamma_upd[grp], it.gamma[grp]);
	norm_param_step(alpha, mom, S, fb, it.beta_upd[grp], it.beta[grp]);
}

}  // namespace cb200
using namespace cb200;

extern "C" {

int cb200_update_plan_accepts(const cb200_conv_desc* d) {
	if (d == nullptr || d->input_is_patches) return 0;
	const size_t row_smem = 2 * ((size_t)conv_taps(d) * d->in_c + 1) * sizeof(float);
	return row_smem <= 96 * 1024 ? 1 : 0;
}

int cb200_update_plan_create(void** plan_out, int dtype, const cb200_conv_desc* const* conv_desc, const cb200_conv_weights* const* conv_w,
                             int n_conv, const cb200_norm_update_ref* norms, int n_norm) {
	CB_REQUIRE_DEVICE();
	CB_ARG(plan_out != nullptr && n_conv >= 0 && n_norm >= 0);
	UpdatePlan* p = (UpdatePlan*)calloc(1, sizeof(UpdatePlan));
	p->dtype = dtype; p->n_conv = n_conv; p->n_norm = n_norm;
	ConvUpdItem* ci = (ConvUpdItem*)calloc(n_conv > 0 ? n_conv : 1, sizeof(ConvUpdItem));
	NormUpdItem* ni = (NormUpdItem*)calloc(n_norm > 0 ? n_norm : 1, sizeof(NormUpdItem));
	for (int i = 0; i < n_conv; i++) {
		const cb200_conv_desc* d = conv_desc[i];
		const cb200_conv_weights* w = conv_w[i];
		if (!cb200_update_plan_accepts(d) || d->dtype != dtype) {
			free(ci); free(ni); free(p);
			set_error("cb200_update_plan_create: layer %d is not eligible (cb200_update_plan_accepts) or has another dtype", i);
			return CB200_ERR_ARG;
		}
		ConvUpdItem& it = ci[i];
		it.master = w->master; it.moment = w->moment; it.grad = w->grad; it.grad_b = w->grad_b;
		it.w_fwd = w->w_fwd; it.w_bwd = w->w_bwd; it.bias_w = w->bias_w; it.bias_value = d->bias_value;
		it.taps = conv_taps(d); it.in_c = d->in_c; it.in_cp = round8(d->in_c); it.out_c = d->out_c; it.out_cp = round8(d->out_c);
		it.wb_dense = conv_whole_map(d) ? 1 : 0;
		it.first_block = p->conv_blocks; p->conv_blocks += d->out_c;
		it.tiles_c = ceil_div(d->in_c, 32); it.tiles_f = ceil_div(it.out_cp, 32);
		it.first_tile = p->conv_tiles; p->conv_tiles += it.tiles_c * it.tiles_f * it.taps;
		const size_t row_smem = 2 * ((size_t)it.taps * it.in_c + 1) * sizeof(float);
		if (row_smem > p->smem) p->smem = row_smem;
	}
	for (int i = 0; i < n_norm; i++) {
		NormUpdItem& it = ni[i];
		it.d_gamma = norms[i].d_gamma; it.d_beta = norms[i].d_beta; it.gsum = norms[i].gsum;
		it.gamma = norms[i].gamma; it.beta = norms[i].beta; it.gamma_upd = norms[i].gamma_upd; it.beta_upd = norms[i].beta_upd;
		it.batch = norms[i].batch; it.nb_group = norms[i].nb_group; it.set_off = norms[i].set_off; it.reduce = norms[i].reduce;
		it.first_group = p->norm_groups; p->norm_groups += norms[i].nb_group;
	}
	cudaError_t e = cudaSuccess;
	if (n_conv > 0) {
		e = cudaMalloc(&p->conv_dev, n_conv * sizeof(ConvUpdItem));
		if (e == cudaSuccess) e = cudaMemcpy(p->conv_dev, ci, n_conv * sizeof(ConvUpdItem), cudaMemcpyHostToDevice);
	}
	if (e == cudaSuccess && n_norm > 0) {
		e = cudaMalloc(&p->norm_dev, n_norm * sizeof(NormUpdItem));
		if (e == cudaSuccess) e = cudaMemcpy(p->norm_dev, ni, n_norm * sizeof(NormUpdItem), cudaMemcpyHostToDevice);
	}
	free(ci); free(ni);
	if (e != cudaSuccess) { set_error("cb200_update_plan_create: %s", cudaGetErrorString(e)); cudaFree(p->conv_dev); cudaFree(p->norm_dev); free(p); return CB200_ERR_CUDA; }
	static bool configured = false;
	if (!configured) {
		CB_CUDA(cudaFuncSetAttribute(conv_update_rows_many_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
		CB_CUDA(cudaFuncSetAttribute(conv_update_rows_many_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
		CB_CUDA(cudaFuncSetAttribute(conv_update_rows_many_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
		configured = true;
	}
	*plan_out = p;
	return CB200_OK;
}

int cb200_update_plan_run(const void* plan, const float* hyper, void* s) {
	CB_REQUIRE_DEVICE();
	const UpdatePlan* p = (const UpdatePlan*)plan;
	CB_ARG(p != nullptr && hyper != nullptr);
	if (p->n_conv > 0) {
		CB_DISPATCH_DTYPE(p->dtype, T, (conv_update_rows_many_kernel<T><<<p->conv_blocks, 256, p->smem, as_stream(s)>>>(p->conv_dev, p->n_conv, hyper)));
		CB_LAUNCH_CHECK();
		CB_DISPATCH_DTYPE(p->dtype, T, (conv_wbwd_transpose_many_kernel<T><<<p->conv_tiles, 256, 0, as_stream(s)>>>(p->conv_dev, p->n_conv)));
		CB_LAUNCH_CHECK();
	}
	if (p->n_norm > 0) {
		norm_update_many_kernel<<<ceil_div(p->norm_groups * 32, 128), 128, 0, as_stream(s)>>>(p->norm_dev, p->n_norm, p->norm_groups, hyper);
		CB_LAUNCH_CHECK();
	}
	return CB200_OK;
}

int cb200_update_plan_destroy(void* plan) {
	UpdatePlan* p = (UpdatePlan*)plan;
	if (p == nullptr) return CB200_OK;
	cudaFree(p->conv_dev); cudaFree(p->norm_dev);
	free(p);
	return CB200_OK;
}

}  // extern "C"
