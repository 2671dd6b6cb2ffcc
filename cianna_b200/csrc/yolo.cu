// yolo.cu - YOLO detection head: activation, target association + error signal, loss monitor, box decoding.
//
// Behaviour follows the reference's YOLO_activation_kernel / YOLO_deriv_error_kernel / YOLO_error_kernel
// (src/cuda/cuda_activ_functions.cu:477-597, :700-1406, :1409-2075) and their CPU twins
// (src/activ_functions.c:1480-2985); the organisation is new:
//  * tensors are channels-last, so the nb_box*(8+nb_class+nb_param) values of one grid cell are one
//    contiguous run instead of values batch*grid apart;
//  * one block per image, threads stride over the grid cells; the association pass and the loss pass share
//    one templated routine (upstream keeps two 700-line copies);
//  * scratch is [image][target][box] - a target belongs to exactly one cell, so the rows a cell uses are
//    addressed by the target's own index: 1/(grid cells) of upstream's per-cell tables and no clearing pass;
//    per-box state (corners, lock flags, allowed-prior flags) lives in registers / local memory;
//  * the loss monitor is reduced on the device to one value per image (+ the six-part split) instead of
//    copying the whole per-element tensor back to the host;
//  * the random association branches draw from a counter-based generator keyed on (seed, step, cell, draw)
//    instead of per-cell curand states.
// Deliberate deviations from upstream, all on code that is out-of-bounds / ill-defined there:
//  * "forced smallest prior" compares prior k with the smallest prior (upstream indexes c_prior_size[k+l] after
//    already offsetting by 3k, reading outside the prior table, :1021-1031);
//  * random box indices are drawn in [0, nb_box) (curand_uniform's (0,1] can yield nb_box upstream);
//  * the loss pass visits cell (x, y) at (x, y) for any grid (upstream swaps the two coordinates, :1469-1471,
//    which is a permutation of the cells on square grids and out of bounds otherwise).
#include "common.cuh"

namespace cb200 {

constexpr int MAXB = CB200_YOLO_MAX_BOX;

struct YoloRng {
	unsigned long long key;
	unsigned int draw;
	__device__ float uniform() {   // [0, 1)
		unsigned long long z = key + 0x9E3779B97F4A7C15ULL * (unsigned long long)(++draw);
		z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
		z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
		z ^= z >> 31;
		return (float)(z >> 40) * (1.0f / 16777216.0f);
	}
};

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}

// ---- the four box-overlap measures (corner order: x0 y0 z0 x1 y1 z1), operation order as upstream :599-698
__device__ __forceinline__ void inter_union(const float* o, const float* t, float& inter, float& uni) {
	const float iw = fmaxf(0.0f, fminf(o[3], t[3]) - fmaxf(o[0], t[0]));
	const float ih = fmaxf(0.0f, fminf(o[4], t[4]) - fmaxf(o[1], t[1]));
	const float id = fmaxf(0.0f, fminf(o[5], t[5]) - fmaxf(o[2], t[2]));
	inter = iw * ih * id;
	uni = fabsf(o[3] - o[0]) * fabsf(o[4] - o[1]) * fabsf(o[5] - o[2])
	    + fabsf(t[3] - t[0]) * fabsf(t[4] - t[1]) * fabsf(t[5] - t[2]) - inter;
}

__device__ float overlap(int type, const float* o, const float* t) {
	float inter, uni;
	inter_union(o, t, inter, uni);
	if (type == CB200_IOU) return inter / uni;
	const float ew = fmaxf(o[3], t[3]) - fminf(o[0], t[0]);
	const float eh = fmaxf(o[4], t[4]) - fminf(o[1], t[1]);
	const float ed = fmaxf(o[5], t[5]) - fminf(o[2], t[2]);
	if (type == CB200_GIOU) {
		const float enclose = ew * eh * ed;
		return inter / uni - (enclose - uni) / enclose;
	}
	const float dx = (o[3] + o[0]) * 0.5f - (t[3] + t[0]) * 0.5f;
	const float dy = (o[4] + o[1]) * 0.5f - (t[4] + t[1]) * 0.5f;
	const float dz = (o[5] + o[2]) * 0.5f - (t[5] + t[2]) * 0.5f;
	float dist = dx * dx + dy * dy + dz * dz;
	float diag = ew * ew + eh * eh + ed * ed;
	if (type == CB200_DIOU) { dist = sqrtf(dist); diag = sqrtf(diag); }
	return inter / uni - dist / diag;
}

// distance between a target (size ts[3]) and prior k, the three flavours of upstream :896-935
__device__ float prior_distance(const cb200_yolo_desc& d, const float* __restrict__ prior, const float* ts,
                                float size_min_sat, float size_max_sat) {
	if (d.prior_dist_type == CB200_DIST_IOU) {
		float a[6], b[6];
#pragma unroll
		for (int l = 0; l < 6; l++) {
			a[l] = copysignf(0.5f, l - 2.5f) * prior[l % 3];
			b[l] = copysignf(0.5f, l - 2.5f) * ts[l % 3];
		}
		return 1.0f - overlap(d.IoU_type, a, b);
	}
	if (d.prior_dist_type == CB200_DIST_OFFSET) {
		float s = 0.0f;
#pragma unroll
		for (int l = 0; l < 3; l++) {
			float r = ts[l] / prior[l];
			r = r < size_min_sat ? logf(size_min_sat) : (r > size_max_sat ? logf(size_max_sat) : logf(r));
			s += fabsf(r);
		}
		return s;
	}
	const float a = ts[0] - prior[0], b = ts[1] - prior[1], c = ts[2] - prior[2];
	return sqrtf(a * a + b * b + c * c);
}

template <typename T>
__device__ __forceinline__ bool target_in_cell(const cb200_yolo_desc& d, const T* __restrict__ t, int cx, int cy) {
	// t points at the target's class slot; corners follow. Cell of the box centre, truncated like upstream :818-825
	const int ox = (int)((to_f32<T>(t[4]) + to_f32<T>(t[1])) * 0.5f / d.cell_size[0]);
	const int oy = (int)((to_f32<T>(t[5]) + to_f32<T>(t[2])) * 0.5f / d.cell_size[1]);
	const int oz = (int)((to_f32<T>(t[6]) + to_f32<T>(t[3])) * 0.5f / d.cell_size[2]);
	return ox == cx && oy == cy && oz == 0;
}

// One grid cell: decode its boxes, find which targets it owns, associate boxes and targets, then either write the
// error signal (LOSS = false) or accumulate the loss monitor (LOSS = true).
template <typename T, bool LOSS>
__device__ void yolo_cell(const cb200_yolo_desc& d, const T* __restrict__ out, const T* __restrict__ tg_row,
                          int cx, int cy, float* __restrict__ ws, YoloRng& rng, long long nb_im_iter, float tc_scale,
                          T* __restrict__ delta, float (&acc)[6], float* __restrict__ monitor, int* __restrict__ box_state) {
	const int nb_box = d.nb_box, nb_class = d.nb_class, nb_param = d.nb_param;
	const int per = 8 + nb_class + nb_param;
	const int tlen = 7 + nb_param + d.diff_flag;
	const float* __restrict__ prior = d.prior_size;
	const float (*sm)[3] = reinterpret_cast<const float (*)[3]>(d.slopes_and_maxes);
	const float size_max_sat = expf(sm[1][1]), size_min_sat = expf(sm[1][2]);
	const float good_lim = d.IoU_limits[0], low_best_lim = d.IoU_limits[1];
	const bool complete = !LOSS || d.error_type == CB200_ERR_COMPLETE;
	const bool natural = LOSS && d.error_type == CB200_ERR_NATURAL;
	const int cell_pos[3] = {cx, cy, 0};
	float* __restrict__ ws_iou = ws;                                            // [max_nb_obj][nb_box]
	unsigned int* __restrict__ ws_allowed = reinterpret_cast<unsigned int*>(ws + (size_t)d.max_nb_obj * nb_box);

	float bx[MAXB][6];
	unsigned int lock1 = 0u, lock2 = 0u;
	int s_p_i = 0;

	int nb_obj = (int)to_f32<T>(tg_row[0]);
	float class_only = -2.0f;
	if (nb_obj == -1) { nb_obj = 1; class_only = good_lim; }
	if (nb_obj > d.max_nb_obj) nb_obj = d.max_nb_obj;
	const T* __restrict__ tg = tg_row + 1;

	{
		float best = 100000000.0f;
		for (int k = 0; k < nb_box; k++) {
			float c[6];
#pragma unroll
			for (int l = 0; l < 3; l++) {
				c[l] = (to_f32<T>(out[k * per + l]) + cell_pos[l]) * d.cell_size[l];
				c[l + 3] = prior[k * 3 + l] * expf(to_f32<T>(out[k * per + l + 3]));
			}
#pragma unroll
			for (int l = 0; l < 6; l++) bx[k][l] = c[l % 3] + copysignf(0.5f, l - 2.5f) * c[3 + l % 3];
			const float dist = sqrtf(prior[k * 3] * prior[k * 3] + prior[k * 3 + 1] * prior[k * 3 + 1] + prior[k * 3 + 2] * prior[k * 3 + 2]);
			if (dist < best) { best = dist; s_p_i = k; }
			if (LOSS && monitor != nullptr) { monitor[k * 2] = -1.0f; monitor[k * 2 + 1] = -1.0f; }
		}
	}

	// ---- every target of the image: good-but-not-best flags; for the targets this cell owns, the overlap row and
	//      the set of priors the target may be associated with
	int nb_in_cell = 0;
	for (int j = 0; j < nb_obj; j++) {
		const T* __restrict__ t = tg + (size_t)j * tlen;
		float ti[6];
#pragma unroll
		for (int l = 0; l < 6; l++) ti[l] = to_f32<T>(t[1 + l]);
		const bool mine = target_in_cell<T>(d, t, cx, cy);
		for (int k = 0; k < nb_box; k++) {
			const float v = overlap(d.IoU_type, bx[k], ti);
			if (v > good_lim) lock1 |= 1u << k;
			if (mine) ws_iou[(size_t)j * nb_box + k] = v;
		}
		if (!mine) continue;
		nb_in_cell++;
		unsigned int allowed = 0xFFFFFFFFu;
		if (complete && d.strict_box_size_association > 0) {
			float dp[MAXB];
			const float ts[3] = {ti[3] - ti[0], ti[4] - ti[1], ti[5] - ti[2]};
			for (int k = 0; k < nb_box; k++) dp[k] = prior_distance(d, prior + k * 3, ts, size_min_sat, size_max_sat);
			for (int l = 0; l < d.strict_box_size_association; l++) {
				float best = 1000000.0f;
				for (int k = 0; k < nb_box; k++)
					if (dp[k] > 0.0f && dp[k] < best) best = dp[k];
				for (int k = 0; k < nb_box; k++)
					if (fabsf(dp[k] - best) < 0.001f) dp[k] = -2.0f;
			}
			allowed = 0u;
			for (int k = 0; k < nb_box; k++)
				if (dp[k] < -1.0f) allowed |= 1u << k;
		}
		ws_allowed[j] = allowed;
	}

	// ---- greedy association, one target per round
	for (int round = 0; round < nb_in_cell; round++) {
		int resp_box = -1, resp_j = -1;
		float max_iou = -2.0f;
		if (!LOSS && nb_im_iter <= (long long)d.rand_startup) {
			// start-up phase: the round-th owned target takes any box that is still free
			for (int k = 0; k < 2 * nb_box; k++) {
				const int rb = min((int)(rng.uniform() * nb_box), nb_box - 1);
				if (!(lock2 >> rb & 1u)) { resp_box = rb; break; }
			}
			if (resp_box == -1) continue;
			int seen = 0;
			for (int j = 0; j < nb_obj; j++)
				if (target_in_cell<T>(d, tg + (size_t)j * tlen, cx, cy) && seen++ == round) { resp_j = j; break; }
		} else {
			for (int j = 0; j < nb_obj; j++) {
				if (!target_in_cell<T>(d, tg + (size_t)j * tlen, cx, cy)) continue;
				const unsigned int allowed = ws_allowed[j];
				for (int k = 0; k < nb_box; k++) {
					const float v = ws_iou[(size_t)j * nb_box + k];
					if (v > max_iou && (allowed >> k & 1u)) { max_iou = v; resp_j = j; resp_box = k; }
				}
			}
			if (resp_box == -1) continue;   // no usable prior left, or more targets than boxes
			const T* __restrict__ t = tg + (size_t)resp_j * tlen;
			const float ts[3] = {to_f32<T>(t[4]) - to_f32<T>(t[1]), to_f32<T>(t[5]) - to_f32<T>(t[2]), to_f32<T>(t[6]) - to_f32<T>(t[3])};
			const float* row = ws_iou + (size_t)resp_j * nb_box;
			if (!LOSS && d.rand_prob > 0.0f && rng.uniform() < d.rand_prob) {
				for (int k = 0; k < 2 * nb_box; k++) {
					const int rb = min((int)(rng.uniform() * nb_box), nb_box - 1);
					if (!(lock2 >> rb & 1u)) { resp_box = rb; break; }
				}
			} else if (complete && ts[0] < d.min_prior_forced_scaling * prior[s_p_i * 3] &&
			           ts[1] < d.min_prior_forced_scaling * prior[s_p_i * 3 + 1] &&
			           ts[2] < d.min_prior_forced_scaling * prior[s_p_i * 3 + 2]) {
				// target smaller than the smallest prior: give it to that prior (or a twin of it)
				float best_v = -2.0f;
				for (int k = 0; k < nb_box; k++)
					if (prior[s_p_i * 3] == prior[k * 3] && prior[s_p_i * 3 + 1] == prior[k * 3 + 1] &&
					    prior[s_p_i * 3 + 2] == prior[k * 3 + 2] && row[k] > best_v) { best_v = row[k]; resp_box = k; }
			} else if (complete && (max_iou < low_best_lim ||
			           (!LOSS && d.rand_prob_best_box_assoc > 0.0f && rng.uniform() < d.rand_prob_best_box_assoc))) {
				// poor prediction: hand the target to its closest prior (or a twin), the best-overlapping free one
				float dp[MAXB], best = 100000.0f;
				for (int k = 0; k < nb_box; k++) {
					dp[k] = prior_distance(d, prior + k * 3, ts, size_min_sat, size_max_sat);
					if (dp[k] < best) best = dp[k];
				}
				float best_v = -2.0f;
				for (int k = 0; k < nb_box; k++)
					if (fabsf(dp[k] - best) < 0.001f && row[k] > best_v) { best_v = row[k]; resp_box = k; }
			}
		}

		// the target leaves the table whatever happens next
		for (int k = 0; k < nb_box; k++) ws_iou[(size_t)resp_j * nb_box + k] = -2.0f;

		const T* __restrict__ t = tg + (size_t)resp_j * tlen;
		float ti[6], ts[3];
#pragma unroll
		for (int l = 0; l < 6; l++) ti[l] = to_f32<T>(t[1 + l]);
#pragma unroll
		for (int l = 0; l < 3; l++) ts[l] = ti[l + 3] - ti[l];
		max_iou = overlap(d.IoU_type, bx[resp_box], ti);
		if (max_iou > 0.98f) max_iou = 0.98f;
		if (class_only > -2.0f) max_iou = class_only;

		const int l_o = resp_box * per;
		const int diff = d.diff_flag ? (int)to_f32<T>(t[7 + nb_param]) : 0;
		// "difficult" targets only train a box that already predicts them well
		if (d.diff_flag && diff > 0 && (natural || max_iou < d.IoU_limits[6] || to_f32<T>(out[l_o + 7]) < d.IoU_limits[7]))
			continue;

		for (int j = 0; j < nb_obj; j++)
			if (target_in_cell<T>(d, tg + (size_t)j * tlen, cx, cy)) ws_iou[(size_t)j * nb_box + resp_box] = -2.0f;
		lock2 |= 1u << resp_box;

		float want[6];
#pragma unroll
		for (int l = 0; l < 3; l++) {
			want[l] = ((ti[l + 3] + ti[l]) * 0.5f - cell_pos[l] * d.cell_size[l]) / (float)d.cell_size[l];
			float r = ts[l] / prior[resp_box * 3 + l];
			want[l + 3] = r < size_min_sat ? logf(size_min_sat) : (r > size_max_sat ? logf(size_max_sat) : logf(r));
		}
		const bool geom_ok = class_only < -1.9f && (d.diff_flag == 0 || diff < 3);
		const bool cls_ok = d.diff_flag == 0 || diff < 2;
		const int cls = (int)to_f32<T>(t[0]) - 1;

		if (LOSS) {
			if (monitor != nullptr) { monitor[resp_box * 2] = to_f32<T>(out[l_o + 7]); monitor[resp_box * 2 + 1] = max_iou; }
			for (int k = 0; k < 3; k++) {
				const float o = to_f32<T>(out[l_o + k]), s = to_f32<T>(out[l_o + k + 3]);
				if (d.fit_parts[0] == 1 && d.fit_dim > k && geom_ok) acc[0] += 0.5f * d.scale_tab[0] * (o - want[k]) * (o - want[k]);
				else if (d.fit_parts[0] == 0 && d.fit_dim > k) acc[0] += 0.5f * d.scale_tab[0] * (o - 0.0f) * (o - 0.0f);
				if (d.fit_parts[1] == 1 && d.fit_dim > k && geom_ok) acc[1] += 0.5f * d.scale_tab[1] * (s - want[k + 3]) * (s - want[k + 3]);
				else if (d.fit_parts[1] == 0 && d.fit_dim > k) acc[1] += 0.5f * d.scale_tab[1] * (s - 0.0f) * (s - 0.0f);
			}
			{
				const float o = to_f32<T>(out[l_o + 6]);
				if (d.fit_parts[2] == 1 && (max_iou > d.IoU_limits[2] || natural)) acc[2] += 0.5f * d.scale_tab[2] * (o - 0.98f) * (o - 0.98f);
				else if (d.fit_parts[2] == 0) acc[2] += 0.5f * d.scale_tab[2] * (o - 0.5f) * (o - 0.5f);
			}
			{
				const float o = to_f32<T>(out[l_o + 7]);
				if (d.fit_parts[3] == 1 && (max_iou > d.IoU_limits[3] || natural)) {
					const double e = (double)o - (1.0 + (double)max_iou) * 0.5;
					acc[3] += (float)((double)(0.5f * d.scale_tab[3]) * e * e);
				} else if (d.fit_parts[3] == 0) {
					const double e = (double)o - 0.5;
					acc[3] += (float)((double)(0.5f * d.scale_tab[3]) * e * e);
				}
			}
			if (d.fit_parts[4] == 1 && ((max_iou > d.IoU_limits[4] && cls_ok) || natural)) {
				for (int k = 0; k < nb_class; k++) {
					const float o = to_f32<T>(out[l_o + 8 + k]);
					if (d.class_softmax) {
						if (k == cls) acc[4] += d.scale_tab[4] * (-logf(o > 0.0000001f ? o : 0.0000001f));
					} else {
						const float w = k == cls ? 0.98f : 0.02f;
						acc[4] += 0.5f * d.scale_tab[4] * (o - w) * (o - w);
					}
				}
			} else if (d.fit_parts[4] == 0 && !d.class_softmax) {
				for (int k = 0; k < nb_class; k++) {
					const float o = to_f32<T>(out[l_o + 8 + k]);
					acc[4] += 0.5f * d.scale_tab[4] * (o - 0.5f) * (o - 0.5f);
				}
			}
			if (d.fit_parts[5] == 1 && ((max_iou > d.IoU_limits[5] && cls_ok) || natural)) {
				for (int k = 0; k < nb_param; k++) {
					const float o = to_f32<T>(out[l_o + 8 + nb_class + k]), w = to_f32<T>(t[7 + k]);
					acc[5] += d.param_ind_scale[k] * 0.5f * d.scale_tab[5] * (o - w) * (o - w);
				}
			} else if (d.fit_parts[5] == 0) {
				for (int k = 0; k < nb_param; k++) {
					const float o = to_f32<T>(out[l_o + 8 + nb_class + k]);
					acc[5] += d.param_ind_scale[k] * 0.5f * d.scale_tab[5] * (o - 0.5f) * (o - 0.5f);
				}
			}
		} else {
			const float S = tc_scale;
			for (int k = 0; k < 3; k++) {
				const float o = to_f32<T>(out[l_o + k]), s = to_f32<T>(out[l_o + k + 3]);
				float dpos = 0.0f, dsize = 0.0f;
				if (d.fit_parts[0] == 1 && d.fit_dim > k && geom_ok) dpos = S * sm[0][0] * d.scale_tab[0] * o * (1.0f - o) * (o - want[k]);
				else if (d.fit_parts[0] == 0 && d.fit_dim > k) dpos = S * sm[0][0] * d.scale_tab[0] * o * (1.0f - o) * (o - 0.5f);
				if (d.fit_parts[1] == 1 && d.fit_dim > k && geom_ok) dsize = S * sm[1][0] * d.scale_tab[1] * (s - want[k + 3]);
				else if (d.fit_parts[1] == 0 && d.fit_dim > k) dsize = S * sm[1][0] * d.scale_tab[1] * (s - 0.0f);
				delta[l_o + k] = from_f32<T>(dpos);
				delta[l_o + k + 3] = from_f32<T>(dsize);
			}
			{
				const float o = to_f32<T>(out[l_o + 6]);
				float v = 0.0f;
				if (d.fit_parts[2] == 1 && max_iou > d.IoU_limits[2]) v = S * sm[2][0] * d.scale_tab[2] * o * (1.0f - o) * (o - 0.98f);
				else if (d.fit_parts[2] == 0) v = S * sm[2][0] * d.scale_tab[2] * o * (1.0f - o) * (o - 0.5f);
				delta[l_o + 6] = from_f32<T>(v);
			}
			{
				const float o = to_f32<T>(out[l_o + 7]);
				float v = 0.0f;
				if (d.fit_parts[3] == 1 && max_iou > d.IoU_limits[3])
					v = (float)((double)(S * sm[3][0] * d.scale_tab[3] * o * (1.0f - o)) * ((double)o - (1.0 + (double)max_iou) * 0.5));
				else if (d.fit_parts[3] == 0) v = S * sm[3][0] * d.scale_tab[3] * o * (1.0f - o) * (o - 0.5f);
				delta[l_o + 7] = from_f32<T>(v);
			}
			for (int k = 0; k < nb_class; k++) {
				const float o = to_f32<T>(out[l_o + 8 + k]);
				float v = 0.0f;
				if (d.fit_parts[4] == 1 && max_iou > d.IoU_limits[4] && cls_ok) {
					if (d.class_softmax) v = S * d.scale_tab[4] * (o - (k == cls ? 1.0f : 0.0f));
					else v = S * sm[4][0] * d.scale_tab[4] * o * (1.0f - o) * (o - (k == cls ? 0.98f : 0.02f));
				} else if (d.fit_parts[4] == 0 && !d.class_softmax)
					v = S * sm[4][0] * d.scale_tab[4] * o * (1.0f - o) * (o - 0.5f);
				delta[l_o + 8 + k] = from_f32<T>(v);
			}
			for (int k = 0; k < nb_param; k++) {
				const float o = to_f32<T>(out[l_o + 8 + nb_class + k]);
				float v = 0.0f;
				if (d.fit_parts[5] == 1 && max_iou > d.IoU_limits[5] && cls_ok) v = d.param_ind_scale[k] * S * sm[5][0] * d.scale_tab[5] * (o - to_f32<T>(t[7 + k]));
				else if (d.fit_parts[5] == 0) v = d.param_ind_scale[k] * S * sm[5][0] * d.scale_tab[5] * (o - 0.5f);
				delta[l_o + 8 + nb_class + k] = from_f32<T>(v);
			}
		}
	}

	// ---- boxes without a target: pull probability / objectness down unless the box is "good but not best"
	for (int k = 0; k < nb_box; k++) {
		const int state = (lock2 >> k & 1u) ? 2 : ((lock1 >> k & 1u) ? 1 : 0);
		if (!LOSS && box_state != nullptr) box_state[k] = state;
		if (state == 2) continue;
		const int l_o = k * per;
		const float op = to_f32<T>(out[l_o + 6]), oo = to_f32<T>(out[l_o + 7]);
		const float lam = d.noobj_prob_prior[k];
		if (LOSS) {
			if (state == 0) {
				if (d.fit_parts[2] == 1) acc[2] += 0.5f * lam * d.scale_tab[2] * (op - 0.02f) * (op - 0.02f);
				else if (d.fit_parts[2] == 0) acc[2] += 0.5f * lam * d.scale_tab[2] * (op - 0.5f) * (op - 0.5f);
				if (d.fit_parts[3] == 1) acc[3] += 0.5f * lam * d.scale_tab[3] * (oo - 0.02f) * (oo - 0.02f);
				else if (d.fit_parts[3] == 0) acc[3] += 0.5f * lam * d.scale_tab[3] * (oo - 0.5f) * (oo - 0.5f);
			}
		} else {
			float vp = 0.0f, vo = 0.0f;
			if (state == 0) {
				if (d.fit_parts[2] == 1) vp = tc_scale * sm[2][0] * lam * d.scale_tab[2] * op * (1.0f - op) * (op - 0.02f);
				else if (d.fit_parts[2] == 0) vp = tc_scale * sm[2][0] * lam * d.scale_tab[2] * op * (1.0f - op) * (op - 0.5f);
				if (d.fit_parts[3] == 1) vo = tc_scale * sm[3][0] * lam * d.scale_tab[3] * oo * (1.0f - oo) * (oo - 0.02f);
				else if (d.fit_parts[3] == 0) vo = tc_scale * sm[3][0] * lam * d.scale_tab[3] * oo * (1.0f - oo) * (oo - 0.5f);
			}
			for (int c = 0; c < per; c++) delta[l_o + c] = from_f32<T>(0.0f);
			delta[l_o + 6] = from_f32<T>(vp);
			delta[l_o + 7] = from_f32<T>(vo);
		}
	}
}

template <typename T>
__global__ void __launch_bounds__(128) yolo_delta_kernel(cb200_yolo_desc d, T* __restrict__ delta, const T* __restrict__ y,
                                                         const T* __restrict__ target, float tc_scale, long long nb_im_iter,
                                                         unsigned long long seed, unsigned long long step,
                                                         int* __restrict__ box_state, float* __restrict__ ws) {
	const int b = blockIdx.x, cells = d.grid_h * d.grid_w;
	const int C = d.nb_box * (8 + d.nb_class + d.nb_param), cp = round8(C);
	float acc[6];
	// grid (image, cell chunk): cells are independent (a target belongs to exactly one cell, the scratch is per target)
	for (int cell = blockIdx.y * blockDim.x + threadIdx.x; cell < cells; cell += gridDim.y * blockDim.x) {
		T* dl = delta + ((size_t)b * cells + cell) * cp;
		for (int c = C; c < cp; c++) dl[c] = from_f32<T>(0.0f);
		if (b >= d.length) {
			for (int c = 0; c < C; c++) dl[c] = from_f32<T>(0.0f);
			if (box_state != nullptr)
				for (int k = 0; k < d.nb_box; k++) box_state[((size_t)b * cells + cell) * d.nb_box + k] = 0;
			continue;
		}
		YoloRng rng;
		rng.key = mix64(seed ^ mix64(step * 0x632BE59BD9B4E019ULL + (unsigned long long)((size_t)b * cells + cell)));
		rng.draw = 0;
		yolo_cell<T, false>(d, y + ((size_t)b * cells + cell) * cp, target + (size_t)b * d.target_stride, cell % d.grid_w, cell / d.grid_w,
			ws + (size_t)b * d.max_nb_obj * (d.nb_box + 1), rng, nb_im_iter, tc_scale, dl, acc, nullptr,
			box_state != nullptr ? box_state + ((size_t)b * cells + cell) * d.nb_box : nullptr);
	}
}

template <typename T>
__global__ void __launch_bounds__(128) yolo_loss_kernel(cb200_yolo_desc d, float* __restrict__ loss, float* __restrict__ parts,
                                                        float* __restrict__ monitor, const T* __restrict__ y,
                                                        const T* __restrict__ target, float* __restrict__ ws) {
	__shared__ float red[4][6];
	const int b = blockIdx.x, cells = d.grid_h * d.grid_w;
	const int cp = round8(d.nb_box * (8 + d.nb_class + d.nb_param));
	float acc[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
	YoloRng rng;
	rng.key = 0; rng.draw = 0;
	for (int cell = blockIdx.y * blockDim.x + threadIdx.x; cell < cells; cell += gridDim.y * blockDim.x) {
		float* mon = monitor != nullptr ? monitor + ((size_t)b * cells + cell) * d.nb_box * 2 : nullptr;
		if (b >= d.length) {
			if (mon != nullptr)
				for (int k = 0; k < 2 * d.nb_box; k++) mon[k] = -1.0f;
			continue;
		}
		yolo_cell<T, true>(d, y + ((size_t)b * cells + cell) * cp, target + (size_t)b * d.target_stride, cell % d.grid_w, cell / d.grid_w,
			ws + (size_t)b * d.max_nb_obj * (d.nb_box + 1), rng, 0, 1.0f, nullptr, acc, mon, nullptr);
	}
#pragma unroll
	for (int p = 0; p < 6; p++) {
		const float v = warp_sum(acc[p]);
		if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][p] = v;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		// (loss / parts are cleared by the launcher; one block per image when the grid has a single chunk)
		float total = 0.0f;
		for (int p = 0; p < 6; p++) {
			const float v = red[0][p] + red[1][p] + red[2][p] + red[3][p];
			if (parts != nullptr) atomicAdd(parts + b * 6 + p, v);
			total += v;
		}
		atomicAdd(loss + b, total);
	}
}

// in-place activation of the raw convolution output, one thread per value (classes under softmax: one thread per box)
template <typename T>
__global__ void yolo_activation_kernel(cb200_yolo_desc d, T* __restrict__ y, long long total, int C, int cp) {
	const float (*sm)[3] = reinterpret_cast<const float (*)[3]>(d.slopes_and_maxes);
	const int per = 8 + d.nb_class + d.nb_param;
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
		const int ch = (int)(i % cp);
		if (ch >= C) { y[i] = from_f32<T>(0.0f); continue; }
		const int in_col = ch % per;
		float v = to_f32<T>(y[i]);
		if (in_col < 3) {
			if (d.fit_dim > in_col) {
				v = fminf(fmaxf(-sm[0][0] * v, sm[0][2]), sm[0][1]);
				v = 1.0f / (1.0f + expf(v));
			} else v = 0.5f;
		} else if (in_col < 6) {
			if (d.fit_dim > in_col - 3) v = fminf(fmaxf(sm[1][0] * v, sm[1][2]), sm[1][1]);
			else v = 0.0f;
		} else if (in_col < 8) {
			const int r = in_col - 4;   // 2: probability, 3: objectness
			v = fminf(fmaxf(-sm[r][0] * v, sm[r][2]), sm[r][1]);
			v = 1.0f / (1.0f + expf(v));
		} else if (in_col < 8 + d.nb_class) {
			if (d.class_softmax) {
				if (in_col != 8) continue;
				float vmax = v;
				for (int j = 1; j < d.nb_class; j++) vmax = fmaxf(vmax, to_f32<T>(y[i + j]));
				float normal = 0.0f;
				for (int j = 0; j < d.nb_class; j++) {
					// the exponentials go through the storage type like upstream (:566-570)
					const T e = from_f32<T>(expf(to_f32<T>(y[i + j]) - vmax));
					y[i + j] = e;
					normal += to_f32<T>(e);
				}
				for (int j = 0; j < d.nb_class; j++) y[i + j] = from_f32<T>(to_f32<T>(y[i + j]) / normal);
				continue;
			}
			v = fminf(fmaxf(-sm[4][0] * v, sm[4][2]), sm[4][1]);
			v = 1.0f / (1.0f + expf(v));
		} else {
			v = fminf(fmaxf(sm[5][0] * v, sm[5][2]), sm[5][1]);
		}
		y[i] = from_f32<T>(v);
	}
}

// decoded forward output in the reference's [C][B][cells] FP32 layout: corners in pixels, then the raw values
template <typename T>
__global__ void yolo_export_kernel(cb200_yolo_desc d, float* __restrict__ dst, const T* __restrict__ y, long long total, int C, int cp) {
	const int per = 8 + d.nb_class + d.nb_param, cells = d.grid_h * d.grid_w;
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
		const int ch = (int)(i % C);
		const long long pix = i / C;
		const int cell = (int)(pix % cells), b = (int)(pix / cells);
		const T* row = y + (size_t)pix * cp;
		const int box = ch / per, in_col = ch % per;
		const int g[3] = {cell % d.grid_w, cell / d.grid_w, 0};
		float v;
		if (in_col < 6) {
			const int a = in_col % 3;
			v = (float)(g[a] * d.cell_size[a]);
			v += to_f32<T>(row[box * per + a]) * d.cell_size[a];
			const float half_size = 0.5f * d.prior_size[box * 3 + a] * expf(to_f32<T>(row[box * per + 3 + a]));
			v = in_col < 3 ? v - half_size : v + half_size;
		} else v = to_f32<T>(row[ch]);
		dst[((size_t)ch * d.batch + b) * cells + cell] = v;
	}
}

static int check_desc(const cb200_yolo_desc* d, const char* who) {
	if (d == nullptr || d->batch <= 0 || d->grid_h <= 0 || d->grid_w <= 0 || d->nb_box <= 0 || d->nb_box > MAXB ||
	    d->nb_class < 0 || d->nb_param < 0 || d->max_nb_obj < 0 || d->prior_size == nullptr || d->noobj_prob_prior == nullptr ||
	    (d->nb_param > 0 && d->param_ind_scale == nullptr) || d->cell_size[0] <= 0 || d->cell_size[1] <= 0 || d->cell_size[2] <= 0 ||
	    d->target_stride < 1 + d->max_nb_obj * (7 + d->nb_param + (d->diff_flag ? 1 : 0))) {
		set_error("%s: inconsistent YOLO descriptor (nb_box must be 1..%d, tables non-NULL, target_stride >= 1+max_nb_obj*(7+nb_param+diff_flag))", who, MAXB);
		return CB200_ERR_ARG;
	}
	return CB200_OK;
}
}  // namespace cb200
using namespace cb200;

extern "C" {

size_t cb200_yolo_workspace_bytes(const cb200_yolo_desc* d) {
	if (d == nullptr) return 0;
	size_t n = (size_t)d->batch * (d->max_nb_obj > 0 ? d->max_nb_obj : 1) * (d->nb_box + 1) * sizeof(float);
	return n;
}

int cb200_yolo_activation(const cb200_yolo_desc* d, void* y, void* s) {
	CB_REQUIRE_DEVICE();
	int rc = check_desc(d, __func__);
	if (rc != CB200_OK) return rc;
	const int C = d->nb_box * (8 + d->nb_class + d->nb_param), cp = round8(C);
	const long long total = (long long)d->batch * d->grid_h * d->grid_w * cp;
	CB_DISPATCH_DTYPE(d->dtype, T, (yolo_activation_kernel<T><<<grid_for(total, 256), 256, 0, as_stream(s)>>>(*d, (T*)y, total, C, cp)));
	CB_LAUNCH_CHECK();
	return CB200_OK;
}

int cb200_yolo_delta(const cb200_yolo_desc* d, void* delta, const void* y, const void* target, float tc_scale, long long nb_im_iter,
                     unsigned long long seed, unsigned long long step, int* box_state, float* workspace, void* s) {
	CB_REQUIRE_DEVICE();
	int rc = check_desc(d, __func__);
	if (rc != CB200_OK) return rc;
	CB_ARG(delta != nullptr && y != nullptr && target != nullptr && workspace != nullptr);
	const dim3 grid((unsigned)d->batch, (unsigned)ceil_div(d->grid_h * d->grid_w, 128));
	CB_DISPATCH_DTYPE(d->dtype, T, (yolo_delta_kernel<T><<<grid, 128, 0, as_stream(s)>>>(*d, (T*)delta, (const T*)y, (const T*)target,
		tc_scale, nb_im_iter, seed, step, box_state, workspace)));
	CB_LAUNCH_CHECK();
	return CB200_OK;
}

int cb200_yolo_loss(const cb200_yolo_desc* d, float* loss, float* parts, float* monitor, const void* y, const void* target,
                    float* workspace, void* s) {
	CB_REQUIRE_DEVICE();
	int rc = check_desc(d, __func__);
	if (rc != CB200_OK) return rc;
	CB_ARG(loss != nullptr && y != nullptr && target != nullptr && workspace != nullptr);
	const dim3 grid((unsigned)d->batch, (unsigned)ceil_div(d->grid_h * d->grid_w, 128));
	CB_CUDA(cudaMemsetAsync(loss, 0, sizeof(float) * (size_t)d->batch, as_stream(s)));
	if (parts != nullptr) CB_CUDA(cudaMemsetAsync(parts, 0, sizeof(float) * 6 * (size_t)d->batch, as_stream(s)));
	CB_DISPATCH_DTYPE(d->dtype, T, (yolo_loss_kernel<T><<<grid, 128, 0, as_stream(s)>>>(*d, loss, parts, monitor, (const T*)y,
		(const T*)target, workspace)));
	CB_LAUNCH_CHECK();
	return CB200_OK;
}

int cb200_yolo_export_boxes(const cb200_yolo_desc* d, float* dst, const void* y, void* s) {
	CB_REQUIRE_DEVICE();
	int rc = check_desc(d, __func__);
	if (rc != CB200_OK) return rc;
	const int C = d->nb_box * (8 + d->nb_class + d->nb_param), cp = round8(C);
	const long long total = (long long)d->batch * d->grid_h * d->grid_w * C;
	CB_DISPATCH_DTYPE(d->dtype, T, (yolo_export_kernel<T><<<grid_for(total, 256), 256, 0, as_stream(s)>>>(*d, dst, (const T*)y, total, C, cp)));
	CB_LAUNCH_CHECK();
	return CB200_OK;
}

}  // extern "C"
