// conv_tc.cu - implicit-GEMM convolution on the 5th-generation tensor cores (sm_100a).
//
// Replaces the reference's "im2col scatter kernel + cublasGemmEx" pipeline for the three GEMMs of a
// conv layer (src/cuda/cuda_conv_layer.cu:379-397 forward, :505-527 data gradient, :551-557 weight
// gradient).  Nothing is unrolled to memory: the GEMM A operand is fetched tap by tap straight from
// the channels-last activation with 4-D TMA boxes (c, x, y, image); out-of-bound box elements are
// zero-filled by the TMA unit, which IS the convolution's zero padding.
//
//   forward / data-gradient kernel (conv_igemm_kernel)
//     D[128 pixels][BN out-channels] += A[pixels][64 ch of tap t] * W[out-ch][tap t][64 ch]^T
//     M tile  = a TWxTHxTN rectangle of output pixels (TW*TH*TN = 128) -> one TMA box per tap
//     K loop  = taps x channel blocks, multi-stage smem ring, mbarrier full/empty pipeline
//     MMA     = tcgen05.mma kind::f16, K-major A and B (128B/64B swizzle), FP32 accumulator in TMEM,
//               double-buffered so the epilogue of tile i overlaps the main loop of tile i+1
//     epilogue= tcgen05.ld -> bias + activation (+ previous layer's derivative for dgrad) -> cast -> store
//     The data gradient is the same kernel run on dy with the rotated/transposed weights (w_bwd)
//     and padding f-1-p.
//   weight-gradient kernel (conv_wgrad_kernel)
//     G[128 out-ch][BN in-ch] (one tap) += dy[pixels][out-ch]^T * x[shifted pixels][in-ch]
//     both operands are MN-major (the contraction index, pixels, is the slow one in memory),
//     split over pixel ranges across CTAs, FP32 atomics into the raw-gradient buffer.
//
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, then the epilogue warps
// (forward/dgrad kernel: 8 = two per TMEM lane quadrant, a warp may only touch lanes 32*(warp%4)..+31;
// weight-gradient kernel: 4).
#include <cuda.h>
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace cb200 {

using namespace ptx;

// ---------------------------------------------------------------- host: tensor-map encoding
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int load_encode() {
	if (g_encode) return CB200_OK;
	void* fn = nullptr;
	cudaDriverEntryPointQueryResult qres;
	cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
	if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
		set_error("cuTensorMapEncodeTiled not available from the driver (%s)", cudaGetErrorString(e));
		return CB200_ERR_CUDA;
	}
	g_encode = (EncodeTiledFn)fn;
	return CB200_OK;
}

// rank-4 map over act[n][y][x][c] (c fastest).  box = (bc, bw, bh, bn)
// pix_stride > 1 (strided convolutions): the box still DELIVERS bw x bh pixels, taken every pix_stride-th pixel of the
// tensor in x and y (TMA traversal stride: a box dimension of N * stride loads N elements)
int make_act_map(CUtensorMap* m, const void* base, int dtype, int cp, int w, int h, int n,
                        int bc, int bw, int bh, int bn, CUtensorMapSwizzle sw, int pix_stride) {
	int rc = load_encode(); if (rc) return rc;
	const cuuint64_t es = 2;
	const int s = pix_stride < 1 ? 1 : pix_stride;
	if (bw * s > 256 || bh * s > 256) { set_error("make_act_map: box %dx%d with pixel stride %d exceeds the TMA box limit", bw, bh, s); return CB200_ERR_UNSUPPORTED; }
	cuuint64_t dims[4] = {(cuuint64_t)cp, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
	cuuint64_t strides[3] = {(cuuint64_t)cp * es, (cuuint64_t)w * cp * es, (cuuint64_t)h * w * cp * es};
	cuuint32_t box[4] = {(cuuint32_t)bc, (cuuint32_t)(bw * s), (cuuint32_t)(bh * s), (cuuint32_t)bn};
	cuuint32_t estr[4] = {1, (cuuint32_t)s, (cuuint32_t)s, 1};
	CUresult r = g_encode(m, dtype == CB200_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
	                      const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
	                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(act %dx%dx%dx%d box %dx%dx%dx%d) failed: %d", cp, w, h, n, bc, bw, bh, bn, (int)r); return CB200_ERR_CUDA; }
	return CB200_OK;
}
// rank-3 map over w[row][tap][c] (c fastest). box = (bc, 1, brow)
int make_w_map(CUtensorMap* m, const void* base, int dtype, int cp, int taps, int rows, int bc, int brow, CUtensorMapSwizzle sw) {
	int rc = load_encode(); if (rc) return rc;
	const cuuint64_t es = 2;
	cuuint64_t dims[3] = {(cuuint64_t)cp, (cuuint64_t)taps, (cuuint64_t)rows};
	cuuint64_t strides[2] = {(cuuint64_t)cp * es, (cuuint64_t)taps * cp * es};
	cuuint32_t box[3] = {(cuuint32_t)bc, 1, (cuuint32_t)brow};
	cuuint32_t estr[3] = {1, 1, 1};
	CUresult r = g_encode(m, dtype == CB200_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
	                      const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
	                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(weights %dx%dx%d box %dx1x%d) failed: %d", cp, taps, rows, bc, brow, (int)r); return CB200_ERR_CUDA; }
	return CB200_OK;
}

// choose the TWxTHxTN pixel rectangle (product = npix, all powers of two) that wastes the least work
void choose_rect(int W, int H, int N, int npix, int& tw, int& th, int& tn) {
	double best = 1e30;
	tw = npix; th = 1; tn = 1;
	for (int a = 1; a <= npix; a <<= 1)
		for (int b = 1; a * b <= npix; b <<= 1) {
			int c = npix / (a * b);
			if (a > 256 || b > 256 || c > 256) continue;
			double cover = (double)ceil_div(W, a) * a * (double)ceil_div(H, b) * b * (double)ceil_div(N, c) * c;
			double cost = cover / ((double)W * H * N) - 1e-6 * a;   // tie-break: prefer wide rows
			if (cost < best) { best = cost; tw = a; th = b; tn = c; }
		}
}

// ================================================================ forward / dgrad kernel
struct IgemmParams {
	// pixel space of the OUTPUT of this GEMM (== input pixel space shifted by the taps; stride 1)
	int W, H, N;
	int tw, th, tn;              // M-tile rectangle
	int tiles_w, tiles_h, tiles_n, tiles_m, tiles_nn, num_tiles;
	int f_h, f_w, off_h, off_w;  // taps and the (negative) padding offset of tap (0,0)
	int kc_blocks;               // channel blocks of BK per tap
	int n_real, n_pad;           // real / padded output channels
	int mode;                    // 0 = forward epilogue, 1 = dgrad epilogue
	int length;
	float bias_value;
	const float* bias_w;
	void* out;
	const void* prev_out;
	cb200_activ activ;           // forward: this layer; dgrad: the previous layer
	uint32_t idesc;
	// halo-reuse variant (conv_halo_kernel)
	int halo_w;                  // pixels per halo row = tw + f_w - 1
	int a_stage_bytes, a_stages; // one halo tile (rounded up to 1024 B), ring depth
	int b_blk_bytes;             // one (tap, channel block) of the resident filter bank
	// 2-CTA cluster variant of conv_igemm_kernel: the two CTAs of a cluster work on two M tiles of the same N tile and
	// each fetches half of the filter block, multicast to both
	int cluster, pairs_m;
	// strided convolutions.  Forward: output pixel (x, y) reads input pixel (x * stride + tap) - the TMA box traverses
	// the input with that stride.  Data gradient of a filter that tiles the input (f == stride, no padding): one launch
	// per tap; GEMM pixel (x, y) of dy is written to dx pixel (x * out_s + out_ox, y * out_s + out_oy) of the
	// (out_W, out_H) map, and the filter tap it uses is w_tap0 of the w_taps taps in the weight tensor.
	int stride;                  // 0 / 1 = dense
	int out_s, out_ox, out_oy, out_W, out_H;   // out_s == 0: output pixel grid == GEMM pixel grid
	int w_tap0, w_taps;          // w_taps == 0: the weight tensor holds f_h * f_w taps and all of them are used
	// halo kernel, forward, n_pad == BN <= 64 (optional, CB200_HALO_TMA_STORE=1): the output tile leaves through a swizzled
	// shared-memory staging tile and one TMA store per warp (one pixel row per thread makes a 128-bit warp store touch 32
	// different lines: the layer-2 forward kernel has its LSU data pipe 73-79 % busy with them,
	// profiles/r1_first_halo_full_digest.txt) - measured slower there, see halo_plan
	int tma_out;
	// group-norm statistics of the layer's output, accumulated by the forward epilogue (GN variant of epilogue_loop):
	// FP64 sums [N][gn_groups][2] = (sum y, sum y^2) per (sample, group of gn_gs output channels) of the values AS STORED
	// (rounded to the 16-bit type) - what norm_stats_kernel would read back from HBM.  gn_gs is a multiple of 8.
	double* gn_ws;
	int gn_gs, gn_groups;
	// every CTA takes a CONTIGUOUS run of tiles instead of every gridDim.x-th one (set with gn_ws: a warp then stays on one
	// sample for many tiles and hands its sums to memory once per sample, not once per tile)
	int contig;
	// ceil(2^64 / d) for d = tiles_m (pairs_m in pair / cluster order), tiles_w, tiles_h, tiles_w * tiles_h: the epilogue threads
	// split a tile index with multiply-high instead of four run-time integer divisions per tile (each a dependent chain of
	// ~20 instructions through the reciprocal unit, on the critical path of a tile's epilogue)
	unsigned long long mg_m, mg_w, mg_h, mg_wh;
	int wide_store;              // 256-bit epilogue stores (CB200_WIDE_STORE=0: off)
	int bias_once;               // single N tile: bias row staged once per CTA instead of per tile (CB200_BIAS_ONCE=0: off)
};
// n / d through the magic number M = ceil(2^64 / d) (exact for 32-bit n and d: n * (M - 2^64 / d) < 2^64 / d); M == 0: d == 1
static inline unsigned long long fastdiv_magic(int d) { return d <= 1 ? 0ull : (~0ull) / (unsigned long long)d + 1ull; }
__device__ __forceinline__ int fastdiv(int n, unsigned long long M) { return M == 0ull ? n : (int)__umul64hi((unsigned long long)(unsigned)n, M); }

// the tiles of this CTA: first, first + step, ... (count of them).  Strided: blockIdx.x, + gridDim.x, ...; contiguous: a
// run of consecutive tiles (pair / cluster order: consecutive PAIRS, this CTA's tile of each).  Same count either way.
struct TileRun { int first, step, count; };
__device__ __forceinline__ TileRun tile_run(const IgemmParams& p) {
	TileRun r;
	if (!p.contig) {
		r.first = blockIdx.x; r.step = gridDim.x;
		r.count = r.first < p.num_tiles ? (p.num_tiles - r.first + r.step - 1) / r.step : 0;
	} else if (p.cluster) {
		const int nc = gridDim.x >> 1, c = blockIdx.x >> 1, np = p.num_tiles >> 1;
		const int q = np / nc, rem = np - q * nc;
		r.count = q + (c < rem ? 1 : 0);
		r.first = 2 * (c * q + (c < rem ? c : rem)) + (int)(blockIdx.x & 1);
		r.step = 2;
	} else {
		const int nc = gridDim.x, c = blockIdx.x;
		const int q = p.num_tiles / nc, rem = p.num_tiles - q * nc;
		r.count = q + (c < rem ? 1 : 0);
		r.first = c * q + (c < rem ? c : rem);
		r.step = 1;
	}
	return r;
}

// tile index -> (M tile, N tile).  Cluster mode enumerates pairs: tile = 2*pair + rank, so that with an even grid the
// two CTAs of a cluster always hold the two tiles of one pair (a pair past the last M tile gets a dummy, all-OOB tile).
__device__ __forceinline__ void decode_tile(const IgemmParams& p, int tile, int& mt, int& nt) {
	if (p.cluster) { const int q = tile >> 1; nt = fastdiv(q, p.mg_m); mt = 2 * (q - nt * p.pairs_m) + (tile & 1); }
	else { nt = fastdiv(tile, p.mg_m); mt = tile - nt * p.tiles_m; }
}

// ---------------------------------------------------------------- shared epilogue of the forward / dgrad kernels
// Eight epilogue warps in two groups of four (one warp per TMEM lane quadrant): group g drains the accumulators of the
// CTA's tiles g, g+2, g+4, ... so two tiles are in their epilogue at any time while the MMA warp runs up to ACC_STAGES
// tiles ahead.  tcgen05.ld -> bias + activation (forward) or the previous layer's derivative (dgrad) -> cast -> store.
// PAIR (cta_group::2 kernels): only the leader CTA's MMA warp waits for drained accumulators, so the epilogue warps of
// both CTAs arrive on the LEADER's barriers (tempty0 is then a shared::cluster address).
// ---- group-norm statistics in the epilogue (GN = true) ----
// A thread owns one pixel row of the tile; per 32-column chunk it sums y and y^2 of the values it has just rounded for
// the store, per group of gn_gs channels (NV = 2 * max(1, 32 / gs) sums: 8 / 4 / 2 for gs = 8 / 16 / >= 32).  A butterfly
// over the low lane bits then leaves every lane with ONE of the NV sums, added over the lanes of its sample (the 32 rows
// of a warp are min(32, tw * th) consecutive pixels of each of one or more samples), and the lane adds it to a register
// that lives across tiles - one per chunk.  Only when the warp moves to other samples or output channels (and at the
// end) do the registers go to memory, as FP64 atomics: with contiguous tile runs that is a few times per CTA.
// Returns the lane's sum of value index bitrev(lane & (NV - 1)): even index = sum y, odd = sum y^2 of group index >> 1.
__device__ __forceinline__ float gn_butterfly(float (&s)[8], int nv, int seg, int lane) {
	if (nv == 8) {
		const bool up = lane & 1;
#pragma unroll
		for (int i = 0; i < 4; i++) { const float send = up ? s[i] : s[i + 4], keep = up ? s[i + 4] : s[i]; s[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1); }
	}
	if (nv >= 4) {
		const int o = nv == 8 ? 2 : 1;
		const bool up = lane & o;
#pragma unroll
		for (int i = 0; i < 2; i++) { const float send = up ? s[i] : s[i + 2], keep = up ? s[i + 2] : s[i]; s[i] = keep + __shfl_xor_sync(0xffffffffu, send, o); }
	}
	{
		const int o = nv >> 1;
		const bool up = lane & o;
		const float send = up ? s[0] : s[1], keep = up ? s[1] : s[0];
		s[0] = keep + __shfl_xor_sync(0xffffffffu, send, o);
	}
	float v = s[0];
	for (int o = nv; o < seg; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

// one 256-bit store (STG.256, sm_100): two adjacent 8-channel packets of a pixel row.  With one pixel row per thread every
// store instruction of a warp touches 32 different lines, and the LSU data pipe - 72 % busy in the layer-2 forward kernel -
// is what the early layers' epilogues wait for: half as many store instructions per row.
__device__ __forceinline__ void st256(void* p, const uint4& a, const uint4& b) {
	asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
	             "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
template <typename T> __device__ __forceinline__ uint4 pack8(const float (&in)[8]);
template <> __device__ __forceinline__ uint4 pack8<__half>(const float (&in)[8]) {
	uint4 r;
	__half2* h = reinterpret_cast<__half2*>(&r);
#pragma unroll
	for (int i = 0; i < 4; i++) h[i] = __floats2half2_rn(in[2 * i], in[2 * i + 1]);
	return r;
}
template <> __device__ __forceinline__ uint4 pack8<__nv_bfloat16>(const float (&in)[8]) {
	uint4 r;
	__nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
	for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(in[2 * i], in[2 * i + 1]);
	return r;
}

template <typename T, int BN, int ACC_STAGES, int NGROUPS = 2, bool PAIR = false, bool TMA_OUT = false, bool GN = false>
__device__ __forceinline__ void epilogue_loop(const IgemmParams& p, uint32_t tmem_base, uint32_t tfull0, uint32_t tempty0,
                                              float* bias_rows, int warp, int lane, int first_warp,
                                              const CUtensorMap* tmap_out = nullptr, uint32_t stg = 0, uint8_t* stg_ptr = nullptr,
                                              float* gn_smem = nullptr /* GN: [BN / 32][NGROUPS * 128] floats */) {
	static_assert(NGROUPS <= ACC_STAGES, "an accumulator stage belongs to one epilogue group at a time");
	const int ew = warp - first_warp;                // 0..7
	const int quad = warp & 3;                       // TMEM lane quadrant this warp may access
	const int grp = ew >> 2;                         // epilogue group
	const int gtid = (ew & 3) * 32 + lane;           // 0..127 inside the group
	const int row = quad * 32 + lane;                // row of the 128-pixel tile
	T* __restrict__ out = reinterpret_cast<T*>(p.out);
	const T* __restrict__ prev = reinterpret_cast<const T*>(p.prev_out);
	// everything the inner loop needs, in registers (the parameter block lives in constant memory)
	const int act = p.activ.type, mode = p.mode, n_real = p.n_real, n_pad = p.n_pad, length = p.length;
	const float leak = p.activ.leak, sat = p.activ.saturation, beta = p.activ.beta, bias_value = p.bias_value;
	const float* __restrict__ bias_w = p.bias_w;
	const int tw = p.tw, th = p.th, tn = p.tn, PW = p.W, PH = p.H, PN = p.N, tiles_w = p.tiles_w, tiles_h = p.tiles_h;
	const int out_s = p.out_s, out_ox = p.out_ox, out_oy = p.out_oy, OW = p.out_W, OH = p.out_H;
	const bool mask_tail = act == CB200_RELU || act == CB200_LOGISTIC || act == CB200_SOFTMAX;
	const bool hook = mode == 1 && prev != nullptr && act != CB200_LINEAR;
	// 0 <= leak <= 1, sat >= 0: z <= 0 ? z*leak : (z > sat ? hi : z) == min(max(z, z*leak), hi) value for value
	const bool relu_minmax = leak >= 0.0f && leak <= 1.0f && sat >= 0.0f;
	const float sat_c = sat - sat * leak;
	// rows of a multiple of 16 channels leave in 32-byte stores (packets come in valid pairs: BN is a multiple of 16 too, and
	// a strided output grid keeps rows 32-byte aligned since n_pad * 2 is a multiple of 32)
	// Measured per layer shape (batch 128, same box A/B): rows of >= 128 channels gain (1x1 64 -> 128 data gradient at 112 px
	// 212 -> 142 us, 3x3 64 -> 128 forward 238 -> 226 us), 32-channel rows do not care, 64-channel rows lose 3 % (layer-2
	// forward 455 -> 471 us): on from 128 channels.
	const bool wide_store = (n_pad & 15) == 0 && n_pad >= 128 && p.wide_store != 0;
	float* bs = bias_rows + grp * 256;
	// group-norm statistics (GN): one running sum per lane and 32-column chunk, see gn_butterfly.  They live in shared
	// memory, one private column per thread (conflict-free), so that the chunk loop can stay rolled: fully unrolled for
	// register accumulators, the epilogue of a 256-column tile is ~80 KB of code that eight warps walk at different places -
	// measured 2.2 x slower on the 3x3 128 -> 256 layer (instruction fetch), where the rolled loop costs nothing.
	constexpr int GN_CHUNKS = (BN + 31) / 32, GN_STRIDE = NGROUPS * 128;
	float* gn_acc = GN ? gn_smem + (ew * 32 + lane) : nullptr;
	if (GN) {
#pragma unroll 1
		for (int i = 0; i < GN_CHUNKS; i++) gn_acc[i * GN_STRIDE] = 0.0f;
	}
	int gn_tni = -1, gn_nt = -1;
	const int gn_gs = GN ? p.gn_gs : 8;
	const int gn_nv = gn_gs >= 32 ? 2 : (gn_gs == 16 ? 4 : 8);
	const int gn_seg = tw * th < 32 ? tw * th : 32;
	auto gn_flush = [&]() {
		// lane -> (sample, value index): the lanes of a sample whose bits above the value index are zero hold its sums
		const int vi = gn_nv == 8 ? (((lane & 1) << 2) | (lane & 2) | ((lane >> 2) & 1)) : (gn_nv == 4 ? (((lane & 1) << 1) | ((lane >> 1) & 1)) : (lane & 1));
		const int pn_l = gn_tni * tn + row / (tw * th);
		const bool writer = ((lane & (gn_seg - 1)) & ~(gn_nv - 1)) == 0 && pn_l < PN;
#pragma unroll 1
		for (int c = 0; c < GN_CHUNKS; c++) {
			const int g = (gn_nt * BN + c * 32) / gn_gs + (gn_gs < 32 ? (vi >> 1) : 0);
			const float a = gn_acc[c * GN_STRIDE];
			if (writer && g < p.gn_groups && a != 0.0f)
				atomicAdd(p.gn_ws + ((size_t)pn_l * p.gn_groups + g) * 2 + (vi & 1), (double)a);
			gn_acc[c * GN_STRIDE] = 0.0f;
		}
	};
	const TileRun run = tile_run(p);
	const bool bias_once = p.tiles_nn == 1 && p.bias_once != 0;
	bool bias_staged = false;
	int it = 0;
	for (int tile = run.first; it < run.count; tile += run.step, it++) {
		if ((it % NGROUPS) != grp) continue;
		const int acc = it % ACC_STAGES;
		const uint32_t acc_phase = (uint32_t)(it / ACC_STAGES) & 1u;
		int mt, nt;
		decode_tile(p, tile, mt, nt);
		const int tni = fastdiv(mt, p.mg_wh), mrem = mt - tni * (tiles_w * tiles_h);
		const int thi = fastdiv(mrem, p.mg_w), twi = mrem - thi * tiles_w;
		if (GN && (tni != gn_tni || nt != gn_nt)) {
			if (gn_tni >= 0) gn_flush();
			gn_tni = tni; gn_nt = nt;
		}
		const int px = twi * tw + (row % tw);
		const int py = thi * th + (row / tw) % th;
		const int pn = tni * tn + row / (tw * th);
		const bool row_ok = px < PW && py < PH && pn < PN;
		const size_t pix = ((size_t)pn * OH + (py * out_s + out_oy)) * OW + (px * out_s + out_ox);
		const bool dead = mask_tail && pn >= length;
		if (TMA_OUT) { if (lane == 0) bulk_wait_read0(); __syncwarp(); }  // this warp's previous rows have left the staging buffer
		// bias row (bias_value * W[f][bias column]) of the tile's output channels, staged in shared memory by the group: once
		// for the whole run when there is a single N tile (two group barriers less in every tile's latency chain - the early
		// layers are bound by that chain), else per tile
		if (!bias_once || !bias_staged) {
			asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");      // previous tile's readers are done
			if (mode == 0)
				for (int c = gtid; c < BN; c += 128) { const int ch = nt * BN + c; bs[c] = ch < n_real ? bias_value * __ldg(bias_w + ch) : 0.0f; }
			asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
			bias_staged = true;
		}

		mbar_wait(tfull0 + 8u * acc, acc_phase);
		tc_fence_after();
		const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
		for (int c0 = 0; c0 < BN; c0 += 32) {
			float gsum[8];
			if (GN) {
#pragma unroll
				for (int i = 0; i < 8; i++) gsum[i] = 0.0f;
			}
			uint32_t r[32];
			if (BN - c0 >= 32) tmem_ld_32x32(t_row + c0, r);
			else { uint32_t h[16]; tmem_ld_32x16(t_row + c0, h);
#pragma unroll
				for (int j = 0; j < 16; j++) { r[j] = h[j]; r[16 + j] = 0; } }
			tmem_ld_wait();
			const int col0 = nt * BN + c0;
			uint4 keep = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
			for (int v = 0; v < 4; v++) {
				const int col = col0 + v * 8;
				if (!row_ok || col >= n_pad || c0 + v * 8 >= BN) continue;
				float o[8];
#pragma unroll
				for (int j = 0; j < 8; j++) o[j] = __uint_as_float(r[v * 8 + j]);
				if (dead) {
#pragma unroll
					for (int j = 0; j < 8; j++) o[j] = 0.0f;
				} else if (mode == 0) {
					const float4 b0 = *reinterpret_cast<const float4*>(bs + c0 + v * 8), b1 = *reinterpret_cast<const float4*>(bs + c0 + v * 8 + 4);
					// bias and activation on register pairs (sm100_ptx.cuh: FADD2, leaky_sat_f32x2): 28 instead of 48 issue
					// slots per 8 values in the warps whose chain tcgen05.ld -> activation -> store bounds the thin layers
					add_f32x2(o[0], o[1], b0.x, b0.y); add_f32x2(o[2], o[3], b0.z, b0.w);
					add_f32x2(o[4], o[5], b1.x, b1.y); add_f32x2(o[6], o[7], b1.z, b1.w);
					if (act == CB200_RELU) {
						if (relu_minmax) {
#pragma unroll
							for (int j = 0; j < 8; j += 2) leaky_sat_f32x2(o[j], o[j + 1], leak, sat_c);
						} else {
#pragma unroll
							for (int j = 0; j < 8; j++) {
								const float z = o[j];
								const float hi = sat + (z - sat) * leak;
								o[j] = z <= 0.0f ? z * leak : (z > sat ? hi : z);
							}
						}
					} else if (act == CB200_LOGISTIC) {
#pragma unroll
						for (int j = 0; j < 8; j++) o[j] = 1.0f / (1.0f + expf(fminf(-beta * o[j], sat)));
					}
					if (col + 8 > n_real) {
#pragma unroll
						for (int j = 0; j < 8; j++) if (col + j >= n_real) o[j] = 0.0f;
					}
				} else {
					if (hook) {
						float pv[8];
						load8<T>(prev + pix * n_pad + col, pv);
						if (act == CB200_RELU) {
#pragma unroll
							for (int j = 0; j < 8; j++) o[j] = (pv[j] <= 0.0f || pv[j] > sat) ? o[j] * leak : o[j];
						} else {
#pragma unroll
							for (int j = 0; j < 8; j++) o[j] = o[j] * beta * pv[j] * (1.0f - pv[j]);
						}
					}
					if (col + 8 > n_real) {
#pragma unroll
						for (int j = 0; j < 8; j++) if (col + j >= n_real) o[j] = 0.0f;
					}
				}
				if (GN) {
					// round once, store those bits, and sum what was stored
					Raw8<T> pk;
					pk.a = pack8<T>(o);
					if (!wide_store) *reinterpret_cast<uint4*>(out + pix * n_pad + col) = pk.a;
					else if ((v & 1) == 0) keep = pk.a;
					else st256(out + pix * n_pad + col - 8, keep, pk.a);
					float q[8];
					unpack8(pk, q);
#pragma unroll
					for (int j = 0; j < 8; j++) { gsum[2 * v] += q[j]; gsum[2 * v + 1] = fmaf(q[j], q[j], gsum[2 * v + 1]); }
				} else if (TMA_OUT) {
					// staging tile [128 rows][BN channels], 16-byte chunks XOR-swizzled like the TMA map of the output (64B / 128B)
					const int sw_x = BN == 32 ? ((row >> 1) & 3) : (row & 7);
					const int chunk = (((c0 >> 3) + v) ^ sw_x) & (BN / 8 - 1);
					store8<T>(reinterpret_cast<T*>(stg_ptr + row * (BN * 2) + chunk * 16), o);
				} else if (wide_store) {
					const uint4 pk = pack8<T>(o);
					if ((v & 1) == 0) keep = pk;
					else st256(out + pix * n_pad + col - 8, keep, pk);
				} else
				store8<T>(out + pix * n_pad + col, o);
			}
			__syncwarp();        // reconverge before the next warp-collective tcgen05.ld
			if (GN) {
				// 8-column sums -> group sums of this chunk (gs = 8: as they are; 16: pairs; >= 32: the whole chunk)
				if (gn_gs >= 16) { gsum[0] += gsum[2]; gsum[1] += gsum[3]; gsum[2] = gsum[4] + gsum[6]; gsum[3] = gsum[5] + gsum[7]; }
				if (gn_gs >= 32) { gsum[0] += gsum[2]; gsum[1] += gsum[3]; }
				gn_acc[(c0 >> 5) * GN_STRIDE] += gn_butterfly(gsum, gn_nv, gn_seg, lane);
			}
		}
		tc_fence_before();
		__syncwarp();
		if (lane == 0) { if (PAIR) mbar_arrive_cluster(tempty0 + 8u * acc); else mbar_arrive(tempty0 + 8u * acc); }
		if (TMA_OUT) {
			// each warp stores its own 32 rows = (32 / tw) pixel rows of the tile rectangle (box of the output map): no group
			// barrier in the epilogue's latency chain (a per-group store measured 13 % slower than plain global stores)
			fence_proxy_async();                                             // generic-proxy writes -> visible to the TMA unit
			__syncwarp();
			if (lane == 0) {                                                 // rows outside the tensor are clipped
				tma_store_4d(tmap_out, stg + (uint32_t)(quad * 32 * BN * 2), 0, twi * tw, thi * th + quad * (32 / tw), tni * tn);
				bulk_commit();
			}
		}
	}
	if (TMA_OUT && lane == 0) bulk_wait0();
	if (GN && gn_tni >= 0) gn_flush();
}

template <int BN, int BK>
struct IgemmCfg {
	static constexpr int A_BYTES = 128 * BK * 2;
	static constexpr int B_BYTES = BN * BK * 2;
	static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
	static constexpr int STAGES_RAW = (196 * 1024) / STAGE_BYTES;
	static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
	static constexpr int GN_BYTES = ((BN + 31) / 32) * 256 * 4;                 // group-norm sums of the epilogue threads (GN epilogue)
	static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 4096 /*bias rows*/ + GN_BYTES;
	static constexpr int ACC_STAGES = BN <= 128 ? 4 : 2;                      // accumulator ring in TMEM (512 columns)
	// (measured: four groups for BN <= 128 do not help this kernel - its small-N launches are issue-bound, not
	//  epilogue-latency-bound - and cost registers)
	static constexpr int EPI_GROUPS = 2;
	static constexpr int THREADS = (2 + 4 * EPI_GROUPS) * 32;
	static constexpr int ACC_COLS = ACC_STAGES * BN;
	static constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128 : ACC_COLS <= 256 ? 256 : 512;
	static constexpr uint32_t LAYOUT = BK == 64 ? 2u : (BK == 32 ? 4u : 6u);   // 128B / 64B / 32B swizzle
	static constexpr uint32_t SBO = 8 * BK * 2;                                 // 8 rows of one swizzle atom
};

template <typename T, int BN, int BK>
__global__ void __launch_bounds__(IgemmCfg<BN, BK>::THREADS, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const IgemmParams p) {
	using Cfg = IgemmCfg<BN, BK>;
	extern __shared__ uint8_t smem_raw[];
	const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
	const uint32_t bar_base = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
	// barrier slots (8 B each): full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then the tmem address slot
	auto full_bar = [&](int s) { return bar_base + 8u * s; };
	auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
	auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + s); };
	auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + 4 + s); };
	const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::STAGES + 8);
	const uint32_t bias_smem = bar_base + 256u;          // 2 x 256 floats, one row per accumulator stage
	uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

	if (threadIdx.x == 0) {
		prefetch_tensormap(&tmap_a);
		prefetch_tensormap(&tmap_b);
		// cluster mode: a stage is free again when BOTH CTAs have consumed it (each writes half of B into both)
		for (int s = 0; s < Cfg::STAGES; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), p.cluster ? 2 : 1); }
		for (int s = 0; s < Cfg::ACC_STAGES; s++) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 4); }
		fence_barrier_init();
	}
	if (warp == 1) { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
	tc_fence_before();
	__syncthreads();
	if (p.cluster) cluster_sync();        // the peer's barriers exist before anything is multicast to them
	tc_fence_after();
	const uint32_t tmem_base = *tmem_slot_ptr;
	const uint32_t rank = p.cluster ? cluster_ctarank() : 0u;

	const int taps = p.f_h * p.f_w;
	const int k_iters = taps * p.kc_blocks;

	if (warp == 0) {
		// ===================== TMA producer =====================
		{      // the whole warp, converged: elect.sync inside the asm picks the issuing lane (see conv_halo_kernel's MMA warp)
			int stage = 0; uint32_t phase = 0;
			const TileRun run = tile_run(p);
			for (int ti = 0, tile = run.first; ti < run.count; ti++, tile += run.step) {
				int mt, nt;
				decode_tile(p, tile, mt, nt);
				const int twi = mt % p.tiles_w, thi = (mt / p.tiles_w) % p.tiles_h, tni = mt / (p.tiles_w * p.tiles_h);
				const int w0 = twi * p.tw, h0 = thi * p.th, n0 = tni * p.tn;
				for (int tap = 0; tap < taps; tap++) {
					const int ky = tap / p.f_w, kx = tap - ky * p.f_w;
					for (int cb = 0; cb < p.kc_blocks; cb++) {
						mbar_wait(empty_bar(stage), phase ^ 1u);
						const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES, sb = sa + Cfg::A_BYTES;
						mbar_arrive_expect_tx_warp(full_bar(stage), Cfg::STAGE_BYTES);
						tma_load_4d_warp(sa, &tmap_a, full_bar(stage), cb * BK, w0 * p.stride + kx + p.off_w, h0 * p.stride + ky + p.off_h, n0);
						if (p.cluster) {   // this CTA's half of the filter block, to both CTAs (tmap_b boxes are BN/2 rows here)
							if (lane == 0)
								tma_load_3d_multicast(sb + rank * (Cfg::B_BYTES / 2), &tmap_b, full_bar(stage), cb * BK, tap + p.w_tap0,
								                      nt * BN + (int)rank * (BN / 2), (uint16_t)3);
							__syncwarp();
						} else
							tma_load_3d_warp(sb, &tmap_b, full_bar(stage), cb * BK, tap + p.w_tap0, nt * BN);
						if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
					}
				}
			}
		}
	} else if (warp == 1) {
		// ===================== MMA issuer =====================
		{      // the whole warp, converged: elect.sync inside the asm picks the issuing lane (see conv_halo_kernel)
			int stage = 0; uint32_t phase = 0;
			int acc = 0; uint32_t acc_phase = 0;
			const uint64_t desc_proto = make_smem_desc(0, 16, Cfg::SBO, Cfg::LAYOUT);
			const uint32_t idesc = p.idesc;
			const bool cluster = p.cluster != 0;
			const TileRun run = tile_run(p);
			for (int ti = 0; ti < run.count; ti++) {
				mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
				tc_fence_after();
				const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
				for (int k = 0; k < k_iters; k++) {
					mbar_wait(full_bar(stage), phase);
					tc_fence_after();
					// (descriptor = prototype + start address >> 4: the issuing thread is alone, keep its path short)
					const uint64_t da = desc_proto + ((smem_base + (uint32_t)(stage * Cfg::STAGE_BYTES)) >> 4);
					const uint64_t db = da + (Cfg::A_BYTES >> 4);
#pragma unroll
					for (int kk = 0; kk < BK / 16; kk++)
						mma_f16_ss_warp(d_tmem, da + 2 * kk, db + 2 * kk, idesc, (k | kk) != 0 ? 1u : 0u);
					if (cluster) mma_commit_multicast_warp(empty_bar(stage), (uint16_t)3);   // frees the slot in both CTAs
					else mma_commit_warp(empty_bar(stage));          // frees the smem slot when these MMAs retire
					if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
				}
				mma_commit_warp(tfull_bar(acc));                // accumulator complete -> epilogue
				if (++acc == Cfg::ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
			}
		}
	} else {
		// ===================== epilogue warps =====================
		float* bias_rows = reinterpret_cast<float*>(smem_raw + (bias_smem - smem_u32(smem_raw)));
		if (p.gn_ws != nullptr)
			epilogue_loop<T, BN, Cfg::ACC_STAGES, Cfg::EPI_GROUPS, false, false, true>(p, tmem_base, tfull_bar(0), tempty_bar(0), bias_rows, warp, lane, 2,
				nullptr, 0, nullptr, bias_rows + 1024);
		else
			epilogue_loop<T, BN, Cfg::ACC_STAGES, Cfg::EPI_GROUPS>(p, tmem_base, tfull_bar(0), tempty_bar(0), bias_rows, warp, lane, 2);
	}

	tc_fence_before();
	__syncthreads();
	if (p.cluster) cluster_sync();        // neither CTA leaves while the other may still signal its barriers
	if (warp == 1) { __syncwarp(); tc_fence_after(); tmem_dealloc(tmem_base, Cfg::TMEM_COLS); }
}

// ---------------------------------------------------------------- CTA-pair variant (cta_group::2) for the wide-N layers
// With one SM per tile every K = 16 step of a 128 x 256 MMA reads 4 KB of A + 8 KB of B from shared memory while TMA
// refills the same 12 KB: 180 B/clk at full tensor rate against the SM's 128 B/clk (the ~70 % ceiling measured on the
// N >= 256 layers).  Here two SMs run ONE MMA of M = 256: each CTA holds its own M tile of A and HALF of the
// filter block (its N half), 8 KB read + 8 KB written per step and SM.  The pair takes two M tiles of the same N tile
// (decode_tile's pair order).  Roles per CTA: warp 0 TMA producer (its A tile and B half; the bytes complete on the
// LEADER's full barrier), warp 1 allocates TMEM in both CTAs and - in the leader only - issues the MMAs and commits to
// the barriers of both CTAs, 8 epilogue warps drain the CTA's own 128 accumulator rows and report to the leader.
// Operand / accumulator placement as checked on the hardware by scripts/exp/cta_pair_probe.cu.
template <int BN, int BK>
struct PairCfg {
	static constexpr int A_BYTES = 128 * BK * 2;
	static constexpr int B_BYTES = (BN / 2) * BK * 2;                          // this CTA's half of the filter block
	static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
	static constexpr int STAGES_RAW = (196 * 1024) / STAGE_BYTES;
	static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
	static constexpr int GN_BYTES = ((BN + 31) / 32) * 256 * 4;
	static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + 4096 + GN_BYTES;
	static constexpr int ACC_STAGES = BN <= 128 ? 4 : 2;
	static constexpr int EPI_GROUPS = 2;
	static constexpr int THREADS = (2 + 4 * EPI_GROUPS) * 32;
	static constexpr int TMEM_COLS = ACC_STAGES * BN <= 256 ? 256 : 512;
	static constexpr uint32_t LAYOUT = BK == 64 ? 2u : (BK == 32 ? 4u : 6u);
	static constexpr uint32_t SBO = 8 * BK * 2;
};

template <typename T, int BN, int BK>
__global__ void __launch_bounds__(PairCfg<BN, BK>::THREADS, 1)
conv_igemm_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const IgemmParams p) {
	using Cfg = PairCfg<BN, BK>;
	extern __shared__ uint8_t smem_raw[];
	const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
	const uint32_t bar_base = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
	auto full_bar = [&](int s) { return bar_base + 8u * s; };
	auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
	auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + s); };
	auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + 4 + s); };
	const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::STAGES + 8);
	const uint32_t bias_smem = bar_base + 256u;
	uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

	if (threadIdx.x == 0) {
		prefetch_tensormap(&tmap_a);
		prefetch_tensormap(&tmap_b);
		// full: the leader's producer arrives once (with the byte count of both CTAs' loads); empty / tfull: one commit of
		// the leader's MMA thread, multicast to both CTAs; tempty (used in the leader only): 4 epilogue warps of each CTA
		for (int s = 0; s < Cfg::STAGES; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
		for (int s = 0; s < Cfg::ACC_STAGES; s++) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 8); }
		fence_barrier_init();
	}
	if (warp == 1) { tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish_pair(); }
	tc_fence_before();
	__syncthreads();
	cluster_sync();                       // the peer's barriers exist before anything is signalled on them
	tc_fence_after();
	const uint32_t tmem_base = *tmem_slot_ptr;
	const uint32_t rank = cluster_ctarank();

	const int taps = p.f_h * p.f_w;
	const int k_iters = taps * p.kc_blocks;

	if (warp == 0) {
		// ===================== TMA producer (both CTAs) =====================
		{      // the whole warp, converged (see above)
			int stage = 0; uint32_t phase = 0;
			const uint32_t lead_full0 = mapa_rank(full_bar(0), 0);
			const TileRun run = tile_run(p);
			for (int ti = 0, tile = run.first; ti < run.count; ti++, tile += run.step) {
				int mt, nt;
				decode_tile(p, tile, mt, nt);
				const int twi = mt % p.tiles_w, thi = (mt / p.tiles_w) % p.tiles_h, tni = mt / (p.tiles_w * p.tiles_h);
				const int w0 = twi * p.tw, h0 = thi * p.th, n0 = tni * p.tn;
				for (int tap = 0; tap < taps; tap++) {
					const int ky = tap / p.f_w, kx = tap - ky * p.f_w;
					for (int cb = 0; cb < p.kc_blocks; cb++) {
						mbar_wait(empty_bar(stage), phase ^ 1u);
						const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES, sb = sa + Cfg::A_BYTES;
						if (rank == 0) mbar_arrive_expect_tx_warp(full_bar(stage), 2 * Cfg::STAGE_BYTES);
						const uint32_t lead_full = lead_full0 + 8u * stage;
						tma_load_4d_pair_warp(sa, &tmap_a, lead_full, cb * BK, w0 * p.stride + kx + p.off_w, h0 * p.stride + ky + p.off_h, n0);
						tma_load_3d_pair_warp(sb, &tmap_b, lead_full, cb * BK, tap + p.w_tap0, nt * BN + (int)rank * (BN / 2));
						if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
					}
				}
			}
		}
	} else if (warp == 1) {
		// ===================== MMA issuer (leader CTA only) =====================
		if (rank == 0) {      // the whole warp, converged: elect.sync inside the asm picks the issuing lane (see conv_halo_kernel)
			int stage = 0; uint32_t phase = 0;
			int acc = 0; uint32_t acc_phase = 0;
			const uint64_t desc_proto = make_smem_desc(0, 16, Cfg::SBO, Cfg::LAYOUT);
			const uint32_t idesc = p.idesc;                  // M = 256
			const TileRun run = tile_run(p);
			for (int ti = 0; ti < run.count; ti++) {
				mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
				tc_fence_after();
				const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
				for (int k = 0; k < k_iters; k++) {
					mbar_wait(full_bar(stage), phase);
					tc_fence_after();
					const uint64_t da = desc_proto + ((smem_base + (uint32_t)(stage * Cfg::STAGE_BYTES)) >> 4);
					const uint64_t db = da + (Cfg::A_BYTES >> 4);
#pragma unroll
					for (int kk = 0; kk < BK / 16; kk++)
						mma_f16_ss_pair_warp(d_tmem, da + 2 * kk, db + 2 * kk, idesc, (k | kk) != 0 ? 1u : 0u);
					mma_commit_pair_warp(empty_bar(stage), (uint16_t)3);      // frees the slot in both CTAs
					if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
				}
				mma_commit_pair_warp(tfull_bar(acc), (uint16_t)3);           // both CTAs' accumulator halves are complete
				if (++acc == Cfg::ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
			}
		}
	} else {
		// ===================== epilogue warps (both CTAs: their own 128 rows) =====================
		float* bias_rows = reinterpret_cast<float*>(smem_raw + (bias_smem - smem_u32(smem_raw)));
		if (p.gn_ws != nullptr)
			epilogue_loop<T, BN, Cfg::ACC_STAGES, Cfg::EPI_GROUPS, true, false, true>(p, tmem_base, tfull_bar(0), mapa_rank(tempty_bar(0), 0), bias_rows, warp, lane, 2,
				nullptr, 0, nullptr, bias_rows + 1024);
		else
			epilogue_loop<T, BN, Cfg::ACC_STAGES, Cfg::EPI_GROUPS, true>(p, tmem_base, tfull_bar(0), mapa_rank(tempty_bar(0), 0), bias_rows, warp, lane, 2);
	}

	tc_fence_before();
	__syncthreads();
	cluster_sync();                       // neither CTA leaves while the other may still signal its barriers / read its smem
	if (warp == 1) { __syncwarp(); tc_fence_after(); tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS); }
}

template <typename T, int BN, int BK>
static int launch_igemm_pair(const CUtensorMap& ma, const CUtensorMap& mb, const IgemmParams& p, cudaStream_t st) {
	using Cfg = PairCfg<BN, BK>;
	static bool configured = false;
	static int max_clusters = -1;
	auto kern = conv_igemm_pair_kernel<T, BN, BK>;
	if (!configured) {
		if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) != cudaSuccess) {
			set_error("cudaFuncSetAttribute(smem=%d) failed", Cfg::SMEM_BYTES); return CB200_ERR_CUDA;
		}
		configured = true;
	}
	cudaLaunchConfig_t cfg;
	memset(&cfg, 0, sizeof(cfg));
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
	cfg.blockDim = dim3(Cfg::THREADS); cfg.dynamicSmemBytes = Cfg::SMEM_BYTES; cfg.stream = st; cfg.attrs = attr; cfg.numAttrs = 1;
	if (max_clusters < 0) {
		cfg.gridDim = dim3((unsigned)(g_num_sms & ~1));
		int n = 0;
		if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) n = g_num_sms / 2 - 2;
		max_clusters = n;
	}
	int grid = 2 * max_clusters;
	if (grid > p.num_tiles) grid = p.num_tiles;              // (num_tiles is even in pair order)
	cfg.gridDim = dim3((unsigned)grid);
	if (cudaLaunchKernelEx(&cfg, kern, ma, mb, p) != cudaSuccess) { set_error("cluster launch of conv_igemm_pair_kernel failed: %s", cudaGetErrorString(cudaGetLastError())); return CB200_ERR_CUDA; }
	g_launches++;
	return CB200_OK;
}

template <typename T, int BN, int BK>
static int launch_igemm(const CUtensorMap& ma, const CUtensorMap& mb, const IgemmParams& p, cudaStream_t st) {
	using Cfg = IgemmCfg<BN, BK>;
	static bool configured = false;
	auto kern = conv_igemm_kernel<T, BN, BK>;
	if (!configured) {
		if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) != cudaSuccess) {
			set_error("cudaFuncSetAttribute(smem=%d) failed", Cfg::SMEM_BYTES); return CB200_ERR_CUDA;
		}
		configured = true;
	}
	int grid = p.num_tiles < g_num_sms ? p.num_tiles : g_num_sms;
	if (p.cluster) {
		// clusters of two CTAs (one per SM): as many as the device can keep resident at once, so that the persistent
		// tile loop of every cluster starts together
		static int max_clusters = -1;
		cudaLaunchConfig_t cfg;
		memset(&cfg, 0, sizeof(cfg));
		cudaLaunchAttribute attr[1];
		attr[0].id = cudaLaunchAttributeClusterDimension;
		attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
		cfg.blockDim = dim3(Cfg::THREADS); cfg.dynamicSmemBytes = Cfg::SMEM_BYTES; cfg.stream = st; cfg.attrs = attr; cfg.numAttrs = 1;
		if (max_clusters < 0) {
			cfg.gridDim = dim3((unsigned)(g_num_sms & ~1));
			int n = 0;
			if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) n = g_num_sms / 2 - 2;
			max_clusters = n;
		}
		grid = 2 * max_clusters;
		if (grid > p.num_tiles) grid = p.num_tiles;          // (num_tiles is even in cluster mode)
		cfg.gridDim = dim3((unsigned)grid);
		if (cudaLaunchKernelEx(&cfg, kern, ma, mb, p) != cudaSuccess) { set_error("cluster launch of conv_igemm_kernel failed: %s", cudaGetErrorString(cudaGetLastError())); return CB200_ERR_CUDA; }
		g_launches++;
		return CB200_OK;
	}
	kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(ma, mb, p);
	CB_LAUNCH_CHECK();
	return CB200_OK;
}

template <typename T>
static int dispatch_igemm(int bn, int bk, const CUtensorMap& ma, const CUtensorMap& mb, const IgemmParams& p, cudaStream_t st) {
#define CASE(BN_, BK_) if (bn == BN_ && bk == BK_) return launch_igemm<T, BN_, BK_>(ma, mb, p, st)
	CASE(256, 64); CASE(128, 64); CASE(64, 64); CASE(32, 64); CASE(16, 64);
	CASE(256, 32); CASE(128, 32); CASE(64, 32); CASE(32, 32); CASE(16, 32);
	CASE(256, 16); CASE(128, 16); CASE(64, 16); CASE(32, 16); CASE(16, 16);
#undef CASE
	set_error("conv_tc: no kernel instance for BN=%d BK=%d", bn, bk);
	return CB200_ERR_UNSUPPORTED;
}

// channel block of the K loop: the last block of a tap may run past the tensor (TMA zero-fills it)
static int pick_bk(int cp) { return cp >= 64 ? 64 : (cp >= 32 ? 32 : (cp >= 16 ? 16 : 0)); }
static int pick_bn(int n_pad) { return n_pad > 128 ? 256 : n_pad > 64 ? 128 : n_pad > 32 ? 64 : n_pad > 16 ? 32 : 16; }
CUtensorMapSwizzle swizzle_for(int bk) { return bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : bk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B; }

// whole-map filters (dense layers behind conv / pool, common.cuh: conv_whole_map) as the 1x1 convolution they are in the
// channels-last layout: f_h*f_w*Cp input "channels" on a 1x1 map.  Operand buffers need no change: w_fwd / grad rows
// [f][tap][c] ARE [f][tap*Cp + c], and w_bwd is kept in (tap, c) row order for these layers (conv.cu: wbwd_row).
// Without this a 32x32x12 -> 3072 dense layer would run 1024 taps of K = 16 instead of 256 K blocks of 64.
static cb200_conv_desc tc_view(const cb200_conv_desc* d) {
	cb200_conv_desc v = *d;
	if (conv_whole_map(d)) {
		v.in_c = d->f_h * d->f_w * round8(d->in_c);
		v.in_h = 1; v.in_w = 1; v.f_h = 1; v.f_w = 1;
	}
	return v;
}

static bool tc_common_ok(const cb200_conv_desc* d) {
	if (d->dtype != CB200_FP16 && d->dtype != CB200_BF16) return false;
	// strided layers: the same kernels with a TMA traversal stride (square strides; 128-pixel rows times the stride must
	// fit a TMA box, so 2 only - what upstream's networks use in place of pooling)
	if (d->stride_h != d->stride_w || d->stride_w < 1 || d->stride_w > 2) return false;
	if (d->pad_h > d->f_h - 1 || d->pad_w > d->f_w - 1) return false;
	return true;
}
bool conv_tc_fwd_supported(const cb200_conv_desc* d_in) {
	const cb200_conv_desc v = tc_view(d_in), *d = &v;
	return tc_common_ok(d) && pick_bk(round8(d->in_c)) != 0;
}
bool conv_tc_dgrad_supported(const cb200_conv_desc* d_in) {
	const cb200_conv_desc v = tc_view(d_in), *d = &v;
	if (!(tc_common_ok(d) && pick_bk(round8(d->out_c)) != 0 && round8(d->in_c) >= 16)) return false;
	if (d->stride_w == 1) return true;
	// strided: only filters that tile the input exactly (every input pixel has one tap and one output pixel)
	return d->f_h == d->stride_h && d->f_w == d->stride_w && d->pad_h == 0 && d->pad_w == 0 &&
	       d->in_h == d->stride_h * d->out_h && d->in_w == d->stride_w * d->out_w;
}

// ================================================================ halo-reuse forward / dgrad kernel
// For filters larger than 1x1 on LARGE maps with FEW channels the per-tap kernel above is bound by the L2 -> SM path:
// every tap re-fetches its own shifted copy of the 128-pixel tile (9x the activation bytes for 3x3) and, per tile, the
// whole filter bank.  This variant fetches, per channel block, ONE halo tile [(16+f_h-1) rows][(8+f_w-1) px][BK ch] with a
// single TMA box and runs all taps from it: tap (ky,kx) is the same shared-memory tile read through a K-major descriptor
// whose start address is moved by (ky*halo_w + kx) pixel rows and whose stride between 8-row groups is one halo row
// (M tile = 16 image rows x 8 pixels, so 8-row group g is image row g).  The swizzle XOR is a function of the absolute
// shared-memory address, so such unaligned starts address exactly the bytes TMA wrote (scripts/exp/halo_desc_probe.cu
// checks this on the hardware for the 128B and 64B swizzles).  The filter bank of the layer (all taps, all channel
// blocks, <= ~150 KB) is loaded once per CTA and stays resident.  L2 -> SM traffic per tile: 1.4x the tile's activation
// bytes instead of 9x + the filters.
template <int BN>
struct HaloCfg {
	static constexpr int ACC_STAGES = 4;                 // BN <= 128
	static constexpr int ACC_COLS = ACC_STAGES * BN;
	static constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128 : ACC_COLS <= 256 ? 256 : 512;
	static constexpr int MAX_A_STAGES = 6;
};
constexpr int HALO_TW = 8, HALO_TH = 16;
constexpr int HALO_EPI_GROUPS = 4;
constexpr int HALO_THREADS = (3 + 4 * HALO_EPI_GROUPS) * 32;    // A producer, B loader, MMA issuer, 4 epilogue groups of 4 warps

template <typename T, int BN, int BK, int FS>
__global__ void __launch_bounds__(HALO_THREADS, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_out, const IgemmParams p) {
	using Cfg = HaloCfg<BN>;
	constexpr uint32_t ROWB = BK * 2;
	constexpr uint32_t LAYOUT = BK == 64 ? 2u : (BK == 32 ? 4u : 6u);
	extern __shared__ uint8_t smem_raw[];
	const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
	constexpr int taps = FS * FS, HALO_W = HALO_TW + FS - 1, HALO_H = HALO_TH + FS - 1;
	const uint32_t b_smem = smem_base + (uint32_t)(p.a_stages * p.a_stage_bytes);
	const uint32_t bar_base = b_smem + (uint32_t)(taps * p.kc_blocks * p.b_blk_bytes);
	auto full_bar = [&](int s) { return bar_base + 8u * s; };
	auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::MAX_A_STAGES + s); };
	auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::MAX_A_STAGES + s); };
	auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::MAX_A_STAGES + 4 + s); };
	const uint32_t bfull_bar = bar_base + 8u * (2 * Cfg::MAX_A_STAGES + 8);
	const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::MAX_A_STAGES + 9);
	const uint32_t bias_smem = bar_base + 256u;
	uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

	if (threadIdx.x == 0) {
		prefetch_tensormap(&tmap_a);
		prefetch_tensormap(&tmap_b);
		for (int s = 0; s < p.a_stages; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
		for (int s = 0; s < Cfg::ACC_STAGES; s++) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 4); }
		mbar_init(bfull_bar, 1);
		fence_barrier_init();
	}
	if (warp == 2) { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem_base = *tmem_slot_ptr;
	constexpr uint32_t a_tx = (uint32_t)(HALO_H * HALO_W) * ROWB;

	if (warp == 0) {
		// ===================== halo producer: one TMA box per (tile, channel block) =====================
		{      // the whole warp, converged (see the MMA warp below)
			int stage = 0; uint32_t phase = 0;
			const TileRun run = tile_run(p);
			for (int ti = 0, tile = run.first; ti < run.count; ti++, tile += run.step) {
				const int twi = tile % p.tiles_w, thi = (tile / p.tiles_w) % p.tiles_h, tni = tile / (p.tiles_w * p.tiles_h);
				for (int cb = 0; cb < p.kc_blocks; cb++) {
					mbar_wait(empty_bar(stage), phase ^ 1u);
					mbar_arrive_expect_tx_warp(full_bar(stage), a_tx);
					tma_load_4d_warp(smem_base + (uint32_t)(stage * p.a_stage_bytes), &tmap_a, full_bar(stage), cb * BK,
					                 twi * HALO_TW + p.off_w, thi * HALO_TH + p.off_h, tni);
					if (++stage == p.a_stages) { stage = 0; phase ^= 1u; }
				}
			}
		}
	} else if (warp == 1) {
		// ===================== filter bank: loaded once, resident for the CTA's life =====================
		if (lane == 0) {
			mbar_arrive_expect_tx(bfull_bar, (uint32_t)(taps * p.kc_blocks * p.b_blk_bytes));
			for (int tap = 0; tap < taps; tap++)
				for (int cb = 0; cb < p.kc_blocks; cb++)
					tma_load_3d(b_smem + (uint32_t)((tap * p.kc_blocks + cb) * p.b_blk_bytes), &tmap_b, bfull_bar, cb * BK, tap, 0);
		}
	} else if (warp == 2) {
		// ===================== MMA issuer =====================
		// The WHOLE warp walks this loop converged, with identical values, and elect.sync inside the asm picks the issuing lane
		// (sm100_ptx.cuh: mma_f16_ss_warp): written as `if (lane == 0) { loop }` every operand lives in a per-thread register and the compiler wraps each
		// UTCHMMA in an elect / branch "waterfall" (ELECT, 2 x PLOP3, BRA.U.ANY, R2UR: ~70 clocks per issue in the ncu source
		// view) - with 36 small MMAs per tile the issuing thread, not the tensor pipe, bounded the 224 px layers.
		{
			mbar_wait(bfull_bar, 0);
			int stage = 0; uint32_t phase = 0;
			int acc = 0; uint32_t acc_phase = 0;
			// The issuing thread is alone: every instruction between two MMAs is on the critical path of a tile whose MMAs
			// are short (N <= 128).  Descriptors are therefore one 64-bit add away from a per-stage base: the start-address
			// field holds addr >> 4 in the low 14 bits and shared memory is < 256 KB, so adding (byte offset >> 4) never
			// carries out of the field; tap offsets are compile-time constants.
			const uint64_t da_proto = make_smem_desc(0, 16, (uint32_t)HALO_W * ROWB, LAYOUT);
			const uint64_t db_proto = make_smem_desc(0, 16, 8 * ROWB, LAYOUT);
			const uint32_t b_tap_step = (uint32_t)(p.kc_blocks * p.b_blk_bytes) >> 4;
			const uint32_t idesc = p.idesc;
			const int kc_blocks = p.kc_blocks, a_stages = p.a_stages, a_stage_bytes = p.a_stage_bytes, b_blk_bytes = p.b_blk_bytes;
			const TileRun run = tile_run(p);
			for (int ti = 0; ti < run.count; ti++) {
				mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
				tc_fence_after();
				const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
				for (int cb = 0; cb < kc_blocks; cb++) {
					mbar_wait(full_bar(stage), phase);
					tc_fence_after();
					const uint64_t da0 = da_proto + ((smem_base + (uint32_t)(stage * a_stage_bytes)) >> 4);
					uint64_t db_t = db_proto + ((b_smem + (uint32_t)(cb * b_blk_bytes)) >> 4);
#pragma unroll
					for (int tap = 0; tap < taps; tap++) {
						const uint64_t da_t = da0 + (((uint32_t)((tap / FS) * HALO_W + (tap % FS)) * ROWB) >> 4);
#pragma unroll
						for (int kk = 0; kk < BK / 16; kk++)
							mma_f16_ss_warp(d_tmem, da_t + 2 * kk, db_t + 2 * kk, idesc, (tap | kk) != 0 ? 1u : (cb != 0 ? 1u : 0u));
						db_t += b_tap_step;
					}
					mma_commit_warp(empty_bar(stage));
					if (++stage == a_stages) { stage = 0; phase ^= 1u; }
				}
				mma_commit_warp(tfull_bar(acc));
				if (++acc == Cfg::ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
			}
		}
	} else {
		float* bias_rows = reinterpret_cast<float*>(smem_raw + (bias_smem - smem_u32(smem_raw)));
		bool done = false;
		if constexpr (BN <= 64) {
			if (p.tma_out) {
				// staging tiles behind the bias rows: bar_base is 1024-byte aligned, 256 B of barriers + 4 KB of bias rows -> + 5 KB
				const uint32_t stg = bar_base + 5120u + (uint32_t)((warp - 3) >> 2) * (128 * BN * 2);
				epilogue_loop<T, BN, Cfg::ACC_STAGES, HALO_EPI_GROUPS, false, true>(p, tmem_base, tfull_bar(0), tempty_bar(0), bias_rows, warp, lane, 3,
					&tmap_out, stg, smem_raw + (stg - smem_u32(smem_raw)));
				done = true;
			}
		}
		if (!done) epilogue_loop<T, BN, Cfg::ACC_STAGES, HALO_EPI_GROUPS>(p, tmem_base, tfull_bar(0), tempty_bar(0), bias_rows, warp, lane, 3);
	}

	tc_fence_before();
	__syncthreads();
	if (warp == 2) { __syncwarp(); tc_fence_after(); tmem_dealloc(tmem_base, Cfg::TMEM_COLS); }
}

constexpr int HALO_SMEM_MAX = 225 * 1024;

template <typename T, int BN, int BK, int FS>
static int launch_halo(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo, const IgemmParams& p, int smem_bytes, cudaStream_t st) {
	static bool configured = false;
	auto kern = conv_halo_kernel<T, BN, BK, FS>;
	if (!configured) {
		if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, HALO_SMEM_MAX) != cudaSuccess) {
			set_error("cudaFuncSetAttribute(smem=%d) failed", HALO_SMEM_MAX); return CB200_ERR_CUDA;
		}
		configured = true;
	}
	const int grid = p.num_tiles < g_num_sms ? p.num_tiles : g_num_sms;
	kern<<<grid, HALO_THREADS, smem_bytes, st>>>(ma, mb, mo, p);
	CB_LAUNCH_CHECK();
	return CB200_OK;
}

template <typename T>
static int dispatch_halo(int bn, int bk, int fs, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo, const IgemmParams& p, int smem, cudaStream_t st) {
#define CASE(BN_, BK_) if (bn == BN_ && bk == BK_) return fs == 3 ? launch_halo<T, BN_, BK_, 3>(ma, mb, mo, p, smem, st) : launch_halo<T, BN_, BK_, 5>(ma, mb, mo, p, smem, st)
	CASE(128, 64); CASE(64, 64); CASE(32, 64);
	CASE(128, 32); CASE(64, 32); CASE(32, 32);
#undef CASE
	set_error("conv_tc: no halo kernel instance for BN=%d BK=%d", bn, bk);
	return CB200_ERR_UNSUPPORTED;
}

extern const char* g_last_conv_impl;
int g_enable_cluster = 0;   // cb200_force_simt bit 2: 2-CTA multicast variant of conv_igemm_kernel (off by default, see run_igemm)
int g_enable_pair = 1;      // cta_group::2 kernel for the wide-N layers (conv_igemm_pair_kernel); cb200_force_simt bit 3 on / bit 4 off, env CB200_CTA_PAIR=0
int g_enable_pair_wgrad = 0;   // conv_wgrad_pair_kernel: off by default (measured 4.03 vs 3.92 ms per step for the family), same bits, env CB200_WGRAD_PAIR=1
int g_disable_halo = 0;     // test hook (cb200_force_simt bit 1): route everything through the per-tap kernel

// Decide whether the layer goes to the halo kernel and, if so, fill its tiling; returns the dynamic smem size or 0.
static int halo_plan(int cin_p, int n_pad, int f_h, int f_w, int out_h, int out_w, int bk, int bn, IgemmParams& p) {
	if (g_disable_halo || f_h != f_w || (f_h != 3 && f_h != 5)) return 0;
	if (bk < 32 || bn > 128 || bn < 32 || n_pad > bn) return 0;
	if (out_w < HALO_TW || out_h < HALO_TH) return 0;
	const double cover = (double)ceil_div(out_w, HALO_TW) * HALO_TW * ceil_div(out_h, HALO_TH) * HALO_TH / ((double)out_w * out_h);
	if (cover > 1.15) return 0;
	const int kcb = ceil_div(cin_p, bk);
	const int b_blk = bn * bk * 2;
	const int b_bytes = f_h * f_w * kcb * b_blk;
	const int halo_w = HALO_TW + f_w - 1, halo_h = HALO_TH + f_h - 1;
	const int a_stage = (halo_h * halo_w * bk * 2 + 1023) & ~1023;
	int fixed = 1024 /*align*/ + 256 /*barriers*/ + HALO_EPI_GROUPS * 1024 /*bias rows*/;
	// forward with exactly one swizzle span of filters per pixel: output through staging tiles + TMA stores, if they fit.
	// OFF by default: measured on the layer-2 forward of Darknet19 (batch 128) the LSU data pipe drops from 79 % to 29 %
	// busy but the launch gets SLOWER, 410 -> 451 us (467 with one store per epilogue group instead of one per warp): this
	// kernel is bound by the latency of a tile's epilogue chain, which the staging + proxy fence lengthen, and its TMA
	// unit already carries the halo loads.  (The first-layer kernel, conv_first.cu, gains 9 % from the same idea.)
	const char* halo_tma = getenv("CB200_HALO_TMA_STORE");
	p.tma_out = 0;
	if (halo_tma != nullptr && halo_tma[0] == '1' && p.mode == 0 && n_pad == bn && bn <= 64 && p.gn_ws == nullptr) {
		const int fixed_out = 1024 + 5120 + HALO_EPI_GROUPS * 128 * bn * 2;
		if ((HALO_SMEM_MAX - fixed_out - b_bytes) / a_stage >= 3) { p.tma_out = 1; fixed = fixed_out; }
	}
	int stages = (HALO_SMEM_MAX - fixed - b_bytes) / a_stage;
	if (stages > HaloCfg<16>::MAX_A_STAGES) stages = HaloCfg<16>::MAX_A_STAGES;
	if (stages < 2) return 0;
	p.halo_w = halo_w; p.a_stage_bytes = a_stage; p.a_stages = stages; p.b_blk_bytes = b_blk;
	return fixed + b_bytes + stages * a_stage;
}

// GEMM over: input tensor `src` with cin_p channels on an (in_h, in_w) grid, weights wmat[rows=n_real][taps][cin_p],
// output pixel grid (out_h, out_w), tap (0,0) reads input pixel (oy + off_h, ox + off_w).
// group-norm statistics in the epilogue: group sizes the chunk arithmetic of epilogue_loop<GN> covers, and every lane group
// of a sample wide enough for the butterfly (gn_butterfly: min(32, tw * th) >= number of sums per 32-column chunk)
// mode 1 (default): only where the sums are hidden behind the tensor pipe - K loops of at least 16 blocks (3x3 filters on
// >= 128 channels): measured at batch 128 on the Darknet19 shapes (profiles/r2_gn_epilogue_stats.txt) the epilogue sums
// cost 9-37 us per launch there against 17-47 us for the statistics pass they replace, while on the 1x1 layers they
// cost what they save (17-60 us against 17-49 us); mode 2: wherever the arithmetic allows (tests); mode 0: never
int g_gn_epilogue_mode = -1;
static bool gn_stats_ok(const IgemmParams& p, int tw, int th, int k_iters) {
	if (p.gn_ws == nullptr || p.mode != 0) return false;
	if (g_gn_epilogue_mode < 0) { const char* e = getenv("CB200_GN_EPILOGUE_STATS"); g_gn_epilogue_mode = e != nullptr && e[0] != '\0' ? atoi(e) : 1; }
	if (g_gn_epilogue_mode == 0 || (g_gn_epilogue_mode == 1 && k_iters < 16)) return false;
	const int gs = p.gn_gs;
	if (!(gs == 8 || gs == 16 || (gs >= 32 && gs % 32 == 0))) return false;
	const int nv = gs >= 32 ? 2 : (gs == 16 ? 4 : 8);
	const int seg = tw * th < 32 ? tw * th : 32;
	return seg >= nv;
}

static int run_igemm(int dtype, const void* src, int cin_p, int in_h, int in_w, int batch,
                     const void* wmat, int n_real, int f_h, int f_w, int off_h, int off_w,
                     int out_h, int out_w, IgemmParams p, cudaStream_t st, int* gn_fused = nullptr) {
	const int bk = pick_bk(cin_p);
	const int n_pad = round8(n_real);
	const int bn = pick_bn(n_pad);
	CUtensorMap ma, mb;
	if (gn_fused) *gn_fused = 0;
	static const int diag = getenv("CB200_TILE_DIAG") ? atoi(getenv("CB200_TILE_DIAG")) : 0;   // 1: contiguous tile runs everywhere (measurement only)
	if (diag & 1) p.contig = 1;
	static const bool no_wide_store = getenv("CB200_WIDE_STORE") != nullptr && getenv("CB200_WIDE_STORE")[0] == '0';
	p.wide_store = no_wide_store ? 0 : 1;
	static const bool no_bias_once = getenv("CB200_BIAS_ONCE") != nullptr && getenv("CB200_BIAS_ONCE")[0] == '0';
	p.bias_once = no_bias_once ? 0 : 1;
	if (p.stride < 1) p.stride = 1;
	if (p.out_s < 1) { p.out_s = 1; p.out_ox = 0; p.out_oy = 0; p.out_W = out_w; p.out_H = out_h; }
	const int w_taps = p.w_taps > 0 ? p.w_taps : f_h * f_w;
	const bool dense = p.stride == 1 && p.out_s == 1 && p.w_taps == 0;
	if (dense) {
		IgemmParams ph = p;
		const int smem = halo_plan(cin_p, n_pad, f_h, f_w, out_h, out_w, bk, bn, ph);
		if (smem > 0) {
			int rc = make_act_map(&ma, src, dtype, cin_p, in_w, in_h, batch, bk, ph.halo_w, HALO_TH + f_h - 1, 1, swizzle_for(bk), 1);
			if (rc) return rc;
			rc = make_w_map(&mb, wmat, dtype, cin_p, f_h * f_w, n_real, bk, bn, swizzle_for(bk));
			if (rc) return rc;
			ph.W = out_w; ph.H = out_h; ph.N = batch;
			ph.tw = HALO_TW; ph.th = HALO_TH; ph.tn = 1;
			ph.tiles_w = ceil_div(out_w, HALO_TW); ph.tiles_h = ceil_div(out_h, HALO_TH); ph.tiles_n = batch;
			ph.tiles_m = ph.tiles_w * ph.tiles_h * ph.tiles_n;
			ph.tiles_nn = 1;
			ph.num_tiles = ph.tiles_m;
			ph.f_h = f_h; ph.f_w = f_w; ph.off_h = off_h; ph.off_w = off_w;
			ph.kc_blocks = ceil_div(cin_p, bk);
			ph.n_real = n_real; ph.n_pad = n_pad;
			ph.idesc = make_idesc_f16(dtype == CB200_BF16, 128, bn, 0, 0);
			ph.mg_m = fastdiv_magic(ph.tiles_m); ph.mg_w = fastdiv_magic(ph.tiles_w); ph.mg_h = fastdiv_magic(ph.tiles_h);
			ph.mg_wh = fastdiv_magic(ph.tiles_w * ph.tiles_h);
			// No epilogue statistics in this kernel (measured, Darknet19 layers 2 / 3 / 5 at batch 128,
			// profiles/r2_gn_epilogue_stats.txt): its launches are bound by the epilogue's latency chain and by HBM, so sums
			// in the epilogue are paid in full, and the contiguous tile runs that keep a warp on one sample (strided runs
			// would mean an FP64 atomic burst every 2-3 tiles) alone cost 70 us of DRAM locality per launch (148 far-apart
			// streams of 512-byte row pieces: 414 -> 488 us, 222 -> 294 us) - more than half of the 146 / 79 us pass saved.
			ph.gn_ws = nullptr;
			g_last_conv_impl = "tcgen05-halo";
			CUtensorMap mo = mb;
			if (ph.tma_out) {
				rc = make_act_map(&mo, ph.out, dtype, n_pad, out_w, out_h, batch, bn, HALO_TW, 32 / HALO_TW, 1,     // one warp's rows
				                  bn == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, 1);
				if (rc) return rc;
			}
			if (dtype == CB200_FP16) return dispatch_halo<__half>(bn, bk, f_h, ma, mb, mo, ph, smem, st);
			return dispatch_halo<__nv_bfloat16>(bn, bk, f_h, ma, mb, mo, ph, smem, st);
		}
	}
	int tw, th, tn;
	choose_rect(out_w, out_h, batch, 128, tw, th, tn);
	int rc = make_act_map(&ma, src, dtype, cin_p, in_w, in_h, batch, bk, tw, th, tn, swizzle_for(bk), p.stride);
	if (rc) return rc;
	p.W = out_w; p.H = out_h; p.N = batch;
	p.tw = tw; p.th = th; p.tn = tn;
	p.tiles_w = ceil_div(out_w, tw); p.tiles_h = ceil_div(out_h, th); p.tiles_n = ceil_div(batch, tn);
	p.tiles_m = p.tiles_w * p.tiles_h * p.tiles_n;
	p.tiles_nn = ceil_div(n_pad, bn);
	p.num_tiles = p.tiles_m * p.tiles_nn;
	// Optional 2-CTA cluster variant (cb200_force_simt bit 2): two CTAs take two M tiles of the same N tile and each
	// fetches half of the filter block, TMA-multicast to both - a third less L2 output traffic.  Measured on the wide-N
	// layers of Darknet19: no gain (71.5 % vs 70.5 % tensor-pipe activity).  Those launches are bound by SHARED-MEMORY
	// bandwidth, not by L2: per K=16 step the MMA reads 4 KB of A + 8 KB of B while TMA refills 12 KB, 180 B/clk at full
	// tensor rate against 128 B/clk per SM, i.e. a 71 % ceiling that multicast does not move (each SM still receives and
	// reads the whole B block).  Lifting it took cta_group::2 MMAs (B split between the two SMs): conv_igemm_pair_kernel below.  The path
	// stays as a tested option.
	p.cluster = (g_enable_cluster && dense && bn >= 128 && bk == 64 && p.tiles_m >= 4 && p.num_tiles >= 2 * g_num_sms) ? 1 : 0;
	// CTA-pair kernel (cta_group::2: one M = 256 MMA over two SMs, each holding half of the filter block): N tiles of 256
	if (g_enable_pair && dense && bn == 256 && bk == 64 && p.tiles_m >= 4 && p.num_tiles >= 2 * g_num_sms) p.cluster = 2;
	if (p.cluster) {
		p.pairs_m = ceil_div(p.tiles_m, 2);
		p.num_tiles = 2 * p.pairs_m * p.tiles_nn;
	}
	rc = make_w_map(&mb, wmat, dtype, cin_p, w_taps, n_real, bk, p.cluster ? bn / 2 : bn, swizzle_for(bk));
	if (rc) return rc;
	p.f_h = f_h; p.f_w = f_w; p.off_h = off_h; p.off_w = off_w;
	p.kc_blocks = ceil_div(cin_p, bk);
	p.n_real = n_real; p.n_pad = n_pad;
	p.idesc = make_idesc_f16(dtype == CB200_BF16, p.cluster == 2 ? 256 : 128, bn, 0, 0);
	p.mg_m = fastdiv_magic(p.cluster ? p.pairs_m : p.tiles_m); p.mg_w = fastdiv_magic(p.tiles_w); p.mg_h = fastdiv_magic(p.tiles_h);
	p.mg_wh = fastdiv_magic(p.tiles_w * p.tiles_h);
	if (gn_stats_ok(p, tw, th, f_h * f_w * p.kc_blocks)) {
		p.contig = 1;
		if (cudaMemsetAsync(p.gn_ws, 0, sizeof(double) * 2 * (size_t)batch * p.gn_groups, st) != cudaSuccess) { set_error("cudaMemsetAsync(group-norm sums) failed"); return CB200_ERR_CUDA; }
		if (gn_fused) *gn_fused = 1;
	} else p.gn_ws = nullptr;
	if (p.cluster == 2) {
		g_last_conv_impl = "tcgen05-pair";
		if (dtype == CB200_FP16) return launch_igemm_pair<__half, 256, 64>(ma, mb, p, st);
		return launch_igemm_pair<__nv_bfloat16, 256, 64>(ma, mb, p, st);
	}
	if (dtype == CB200_FP16) return dispatch_igemm<__half>(bn, bk, ma, mb, p, st);
	return dispatch_igemm<__nv_bfloat16>(bn, bk, ma, mb, p, st);
}

// gn / gn_ws: the group-norm layer that follows and its FP64 workspace; *gn_fused = 1 when the epilogue has left the
// layer's (sum, sum of squares) there (cb200_conv_forward_stats), 0 when the caller still has to run the statistics pass
int conv_forward_tc(const cb200_conv_desc* d_in, const cb200_conv_weights* w, const void* x, void* y, cudaStream_t st,
                    const cb200_norm_desc* gn, void* gn_ws, int* gn_fused) {
	const cb200_conv_desc v = tc_view(d_in), *d = &v;
	IgemmParams p;
	memset(&p, 0, sizeof(p));
	p.mode = 0; p.length = d->length; p.bias_value = d->bias_value; p.bias_w = w->bias_w;
	p.out = y; p.prev_out = nullptr; p.activ = d->activ;
	p.stride = d->stride_w;
	if (gn != nullptr && gn_ws != nullptr && gn->c == d->out_c && gn->batch == d->batch && gn->h == d->out_h && gn->w == d->out_w) {
		p.gn_ws = (double*)gn_ws; p.gn_gs = gn->group_size; p.gn_groups = gn->nb_group;
	}
	return run_igemm(d->dtype, x, round8(d->in_c), d->in_h, d->in_w, d->batch, w->w_fwd, d->out_c,
	                 d->f_h, d->f_w, -d->pad_h, -d->pad_w, d->out_h, d->out_w, p, st, gn_fused);
}

int conv_dgrad_tc(const cb200_conv_desc* d_in, const cb200_conv_weights* w, const void* dy, void* dx,
                  const cb200_activ* prev_activ, const void* prev_out, cudaStream_t st) {
	const cb200_conv_desc v = tc_view(d_in), *d = &v;
	IgemmParams p;
	memset(&p, 0, sizeof(p));
	p.mode = 1; p.length = d->length; p.bias_value = 0.0f; p.bias_w = nullptr;
	p.out = dx; p.prev_out = prev_out;
	p.activ.type = CB200_LINEAR;
	if (prev_activ && prev_out) p.activ = *prev_activ;
	if (d->stride_w > 1) {
		// the filter tiles the input (f == stride, no padding: conv_tc_dgrad_supported): input pixel (s*oy + ky, s*ox + kx)
		// receives dy(oy, ox) through tap (ky, kx) only - one 1x1 GEMM per tap, scattered with stride s into dx
		const int s = d->stride_w, taps = d->f_h * d->f_w;
		for (int ky = 0; ky < d->f_h; ky++)
			for (int kx = 0; kx < d->f_w; kx++) {
				IgemmParams q = p;
				q.out_s = s; q.out_oy = ky; q.out_ox = kx; q.out_W = d->in_w; q.out_H = d->in_h;
				q.w_taps = taps; q.w_tap0 = taps - 1 - (ky * d->f_w + kx);        // w_bwd keeps the taps in rotated order
				int rc = run_igemm(d->dtype, dy, round8(d->out_c), d->out_h, d->out_w, d->batch, w->w_bwd, d->in_c,
				                   1, 1, 0, 0, d->out_h, d->out_w, q, st);
				if (rc) return rc;
			}
		return CB200_OK;
	}
	// full correlation of dy with the rotated filters: padding f-1-p
	return run_igemm(d->dtype, dy, round8(d->out_c), d->out_h, d->out_w, d->batch, w->w_bwd, d->in_c,
	                 d->f_h, d->f_w, -(d->f_h - 1 - d->pad_h), -(d->f_w - 1 - d->pad_w), d->in_h, d->in_w, p, st);
}

// ================================================================ weight-gradient kernel
// One CTA owns MF blocks of 128 output channels x up to TGMAX filter taps x BNC input channels: per 64-pixel K step it
// fetches the dy slabs once and one shifted x tile per tap, and issues MF * taps accumulating MMAs into separate TMEM
// column ranges.  Sharing dy across taps (and x across the MF output blocks) is what keeps the L2 -> SM traffic of the
// early, HBM-bound layers near one pass over dy instead of one pass per tap.
struct WgradParams {
	int W, H, N;                 // OUTPUT pixel grid of the layer (dy) ; x is read at pixel + tap offset
	int tw, th, tn;              // pixel rectangle of one K step (64 pixels)
	int tiles_w, tiles_h, tiles_n, pix_tiles;
	int f_h, f_w, off_h, off_w;
	int stride;                  // x pixel = dy pixel * stride + tap offset
	int f_groups, c_tiles, tap_groups, tg;   // job grid; tg = taps per group (<= TGMAX)
	int splits, tiles_per_split, stages;
	int out_c, in_cp;
	float* grad;                 // [out_c][taps][in_cp]
	uint32_t idesc, tmem_cols;
};

template <int BNC, int SLAB_C>
struct WgradCfg {
	static constexpr int KPIX = 64;                                 // pixels per pipeline stage
	static constexpr int A_SLAB_BYTES = KPIX * 128;                 // dy: slabs of [64 pix][64 ch]
	static constexpr int A_BYTES = 2 * A_SLAB_BYTES;                // one block of 128 output channels
	static constexpr int B_SLABS = BNC / SLAB_C;
	static constexpr int B_ROW_BYTES = SLAB_C * 2;
	static constexpr int B_SLAB_BYTES = KPIX * B_ROW_BYTES;
	static constexpr int B_BYTES = B_SLABS * B_SLAB_BYTES;          // one tap
	static constexpr uint32_t B_LAYOUT = SLAB_C == 64 ? 2u : (SLAB_C == 32 ? 4u : 6u);
	static constexpr int SMEM_DATA = 196 * 1024;
	static constexpr int SMEM_BYTES = SMEM_DATA + 1024 + 256;
};

// grid.x = f_groups * c_tiles * tap_groups, grid.y = splits
template <int BNC, int SLAB_C, int MF, int TGMAX>
__global__ void __launch_bounds__(192, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x, const WgradParams p) {
	using Cfg = WgradCfg<BNC, SLAB_C>;
	constexpr int MAX_STAGES = 8;
	extern __shared__ uint8_t smem_raw[];
	const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
	const uint32_t bar_base = smem_base + Cfg::SMEM_DATA;
	auto full_bar = [&](int s) { return bar_base + 8u * s; };
	auto empty_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
	const uint32_t done_bar = bar_base + 8u * (2 * MAX_STAGES);
	const uint32_t tmem_slot = bar_base + 8u * (2 * MAX_STAGES + 1);
	uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

	const int taps = p.f_h * p.f_w;
	int job = blockIdx.x;
	const int tgi = job % p.tap_groups; job /= p.tap_groups;
	const int ct = job % p.c_tiles;
	const int fg = job / p.c_tiles;
	const int tap0 = tgi * p.tg;
	const int ntap = min(p.tg, taps - tap0);                          // taps of this CTA (>= 1)
	const uint32_t stage_bytes = (uint32_t)(MF * Cfg::A_BYTES + p.tg * Cfg::B_BYTES);
	const uint32_t tx_bytes = (uint32_t)(MF * Cfg::A_BYTES + ntap * Cfg::B_BYTES);
	const int stages = p.stages;

	if (threadIdx.x == 0) {
		prefetch_tensormap(&tmap_dy);
		prefetch_tensormap(&tmap_x);
		for (int s = 0; s < stages; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
		mbar_init(done_bar, 1);
		fence_barrier_init();
	}
	if (warp == 1) { tmem_alloc(tmem_slot, p.tmem_cols); tmem_relinquish(); }
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem_base = *tmem_slot_ptr;

	const int t_begin = blockIdx.y * p.tiles_per_split;
	int t_end = t_begin + p.tiles_per_split;
	if (t_end > p.pix_tiles) t_end = p.pix_tiles;
	const int n_steps = t_end - t_begin;

	if (warp == 0) {
		{      // the whole warp, converged: elect.sync inside the asm picks the issuing lane
			int stage = 0; uint32_t phase = 0;
			for (int t = t_begin; t < t_end; t++) {
				const int twi = t % p.tiles_w, thi = (t / p.tiles_w) % p.tiles_h, tni = t / (p.tiles_w * p.tiles_h);
				const int w0 = twi * p.tw, h0 = thi * p.th, n0 = tni * p.tn;
				mbar_wait(empty_bar(stage), phase ^ 1u);
				const uint32_t sa = smem_base + stage * stage_bytes, sb = sa + MF * Cfg::A_BYTES;
				mbar_arrive_expect_tx_warp(full_bar(stage), tx_bytes);
#pragma unroll
				for (int mf = 0; mf < MF; mf++) {
					const int ch0 = (fg * MF + mf) * 128;
					tma_load_4d_warp(sa + mf * Cfg::A_BYTES, &tmap_dy, full_bar(stage), ch0, w0, h0, n0);
					tma_load_4d_warp(sa + mf * Cfg::A_BYTES + Cfg::A_SLAB_BYTES, &tmap_dy, full_bar(stage), ch0 + 64, w0, h0, n0);
				}
				for (int ti = 0; ti < ntap; ti++) {
					const int tap = tap0 + ti;
					const int ky = tap / p.f_w, kx = tap - ky * p.f_w;
#pragma unroll
					for (int sl = 0; sl < Cfg::B_SLABS; sl++)
						tma_load_4d_warp(sb + ti * Cfg::B_BYTES + sl * Cfg::B_SLAB_BYTES, &tmap_x, full_bar(stage),
						            ct * BNC + sl * SLAB_C, w0 * p.stride + kx + p.off_w, h0 * p.stride + ky + p.off_h, n0);
				}
				if (++stage == stages) { stage = 0; phase ^= 1u; }
			}
		}
	} else if (warp == 1) {
		{      // the whole warp, converged: elect.sync inside the asm picks the issuing lane (see conv_halo_kernel)
			int stage = 0; uint32_t phase = 0;
			const uint64_t da_proto = make_smem_desc(0, Cfg::A_SLAB_BYTES, 1024, 2);
			const uint64_t db_proto = make_smem_desc(0, Cfg::B_SLAB_BYTES, 8 * Cfg::B_ROW_BYTES, Cfg::B_LAYOUT);
			const uint32_t idesc = p.idesc;
			const int tg = p.tg;
			for (int k = 0; k < n_steps; k++) {
				mbar_wait(full_bar(stage), phase);
				tc_fence_after();
				// MN-major: 8 pixel rows per K group -> SBO; channel slabs -> LBO; 16 pixel rows per MMA.  Descriptors are a
				// 64-bit add away from the per-stage base (single issuing thread: keep its instruction path short).
				const uint32_t sa = smem_base + stage * stage_bytes;
				const uint64_t da_s = da_proto + (sa >> 4);
				const uint64_t db_s = db_proto + ((sa + MF * Cfg::A_BYTES) >> 4);
				const uint32_t accum = k != 0 ? 1u : 0u;
#pragma unroll
				for (int mf = 0; mf < MF; mf++) {
					uint64_t db_t = db_s;
					uint32_t d_tmem = tmem_base + (uint32_t)(mf * tg * BNC);
					for (int ti = 0; ti < ntap; ti++) {
#pragma unroll
						for (int kk = 0; kk < Cfg::KPIX / 16; kk++)
							mma_f16_ss_warp(d_tmem, da_s + ((mf * Cfg::A_BYTES + kk * 2048) >> 4), db_t + ((kk * 16 * Cfg::B_ROW_BYTES) >> 4), idesc,
							           kk != 0 ? 1u : accum);
						db_t += Cfg::B_BYTES >> 4;
						d_tmem += BNC;
					}
				}
				mma_commit_warp(empty_bar(stage));
				if (++stage == stages) { stage = 0; phase ^= 1u; }
			}
			mma_commit_warp(done_bar);
		}
	} else if (n_steps > 0) {
		const int quad = warp & 3;
		mbar_wait(done_bar, 0);
		tc_fence_after();
		for (int mf = 0; mf < MF; mf++) {
			const int f = (fg * MF + mf) * 128 + quad * 32 + lane;
			for (int ti = 0; ti < ntap; ti++) {
				const int tap = tap0 + ti;
				const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((mf * p.tg + ti) * BNC);
#pragma unroll 1
				for (int c0 = 0; c0 < BNC; c0 += 32) {
					uint32_t r[32];
					if (BNC - c0 >= 32) tmem_ld_32x32(t_row + c0, r);
					else { uint32_t h[16]; tmem_ld_32x16(t_row + c0, h);
#pragma unroll
						for (int j = 0; j < 16; j++) { r[j] = h[j]; r[16 + j] = 0; } }
					tmem_ld_wait();
					if (f < p.out_c) {
						float* dst = p.grad + ((size_t)f * taps + tap) * p.in_cp + ct * BNC + c0;
#pragma unroll
						for (int j = 0; j < 32; j += 4) {
							if (c0 + j < BNC && ct * BNC + c0 + j < p.in_cp)      // in_cp is a multiple of 8: whole quads are in or out
								asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(r[j])),
								             "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])), "f"(__uint_as_float(r[j + 3])) : "memory");
						}
					}
					__syncwarp();
				}
			}
		}
	}
	tc_fence_before();
	__syncthreads();
	if (warp == 1) { __syncwarp(); tc_fence_after(); tmem_dealloc(tmem_base, p.tmem_cols); }
}

// ---------------------------------------------------------------- weight gradient on a CTA pair (cta_group::2)
// The wide layers (>= 256 channels on both sides) ran two 128 x 256 x 16 MMAs per 16-pixel step on ONE SM (MF = 2 blocks
// of output channels sharing the x tile): 24 KB of shared-memory operand reads + 16 KB of TMA writes per 262 tensor
// clocks = 153 B/clk against the SM's 128 B/clk (measured 60 % tensor activity).  Here the two blocks of output channels
// sit on two SMs and run ONE M = 256 MMA: each CTA fetches its own dy block and HALF of the x tile's channels
// (MN-major operands: pixels are the contraction index), up to two taps sharing the dy slabs: 16 KB read + 12 KB written
// per 262 clocks and SM.  Accumulators: CTA r's TMEM lanes = output channels of block 2 fg + r, 256 columns per tap.
// MEASURED (Darknet19-448, batch 128): no gain - the family takes 4.03 ms per step with it, 3.92 without (both with the
// wave-aware splits below, 4.40 before them): these launches are bound by the operand bytes pulled through L2 and by
// whole waves of 0.1-0.3 ms CTAs, not by shared-memory bandwidth, and five tap groups of (2,2,2,2,1) taps balance worse
// than nine of one.  Kept bit-exact and tested, off by default (cb200_force_simt bit 3 / CB200_WGRAD_PAIR=1).
// grid.x = 2 * (f_groups * c_tiles * tap_groups), cluster (2, 1, 1); grid.y = splits.
struct WgradPairCfg {
	static constexpr int KPIX = 64, BNC = 256;
	static constexpr int SLAB_BYTES = KPIX * 128;                   // [64 pix][64 ch], 128B swizzle
	static constexpr int A_BYTES = 2 * SLAB_BYTES;                  // this CTA's 128 output channels of dy
	static constexpr int B_BYTES = 2 * SLAB_BYTES;                  // this CTA's 128 of the tile's 256 input channels, one tap
	static constexpr int SMEM_DATA = 192 * 1024;
	static constexpr int SMEM_BYTES = SMEM_DATA + 1024 + 256;
};

__global__ void __launch_bounds__(192, 1)
conv_wgrad_pair_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x, const WgradParams p) {
	using Cfg = WgradPairCfg;
	constexpr int MAX_STAGES = 8;
	extern __shared__ uint8_t smem_raw[];
	const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
	const uint32_t bar_base = smem_base + Cfg::SMEM_DATA;
	auto full_bar = [&](int s) { return bar_base + 8u * s; };
	auto empty_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
	const uint32_t done_bar = bar_base + 8u * (2 * MAX_STAGES);
	const uint32_t tmem_slot = bar_base + 8u * (2 * MAX_STAGES + 1);
	uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t rank = cluster_ctarank();

	const int taps = p.f_h * p.f_w;
	int job = blockIdx.x >> 1;
	const int tgi = job % p.tap_groups; job /= p.tap_groups;
	const int ct = job % p.c_tiles;
	const int fg = job / p.c_tiles;
	const int tap0 = tgi * p.tg;
	const int ntap = min(p.tg, taps - tap0);
	const uint32_t stage_bytes = (uint32_t)(Cfg::A_BYTES + p.tg * Cfg::B_BYTES);
	const uint32_t tx_bytes = 2u * (uint32_t)(Cfg::A_BYTES + ntap * Cfg::B_BYTES);      // both CTAs' loads
	const int stages = p.stages;

	if (threadIdx.x == 0) {
		prefetch_tensormap(&tmap_dy);
		prefetch_tensormap(&tmap_x);
		for (int s = 0; s < stages; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
		mbar_init(done_bar, 1);
		fence_barrier_init();
	}
	if (warp == 1) { tmem_alloc_pair(tmem_slot, p.tmem_cols); tmem_relinquish_pair(); }
	tc_fence_before();
	__syncthreads();
	cluster_sync();
	tc_fence_after();
	const uint32_t tmem_base = *tmem_slot_ptr;

	const int t_begin = blockIdx.y * p.tiles_per_split;
	int t_end = t_begin + p.tiles_per_split;
	if (t_end > p.pix_tiles) t_end = p.pix_tiles;
	const int n_steps = t_end - t_begin;

	if (warp == 0) {
		{      // the whole warp, converged: elect.sync inside the asm picks the issuing lane
			int stage = 0; uint32_t phase = 0;
			const uint32_t lead_full0 = mapa_rank(full_bar(0), 0);
			const int ch0 = (fg * 2 + (int)rank) * 128;                    // this CTA's output channels
			const int xc0 = ct * Cfg::BNC + (int)rank * 128;               // this CTA's half of the input-channel tile
			for (int t = t_begin; t < t_end; t++) {
				const int twi = t % p.tiles_w, thi = (t / p.tiles_w) % p.tiles_h, tni = t / (p.tiles_w * p.tiles_h);
				const int w0 = twi * p.tw, h0 = thi * p.th, n0 = tni * p.tn;
				mbar_wait(empty_bar(stage), phase ^ 1u);
				const uint32_t sa = smem_base + stage * stage_bytes, sb = sa + Cfg::A_BYTES;
				if (rank == 0) mbar_arrive_expect_tx_warp(full_bar(stage), tx_bytes);
				const uint32_t lead_full = lead_full0 + 8u * stage;
				tma_load_4d_pair_warp(sa, &tmap_dy, lead_full, ch0, w0, h0, n0);
				tma_load_4d_pair_warp(sa + Cfg::SLAB_BYTES, &tmap_dy, lead_full, ch0 + 64, w0, h0, n0);
				for (int ti = 0; ti < ntap; ti++) {
					const int tap = tap0 + ti;
					const int ky = tap / p.f_w, kx = tap - ky * p.f_w;
#pragma unroll
					for (int sl = 0; sl < 2; sl++)
						tma_load_4d_pair_warp(sb + ti * Cfg::B_BYTES + sl * Cfg::SLAB_BYTES, &tmap_x, lead_full,
						                 xc0 + sl * 64, w0 * p.stride + kx + p.off_w, h0 * p.stride + ky + p.off_h, n0);
				}
				if (++stage == stages) { stage = 0; phase ^= 1u; }
			}
		}
	} else if (warp == 1) {
		if (rank == 0) {      // the whole warp, converged: elect.sync inside the asm picks the issuing lane (see conv_halo_kernel)
			int stage = 0; uint32_t phase = 0;
			const uint64_t d_proto = make_smem_desc(0, Cfg::SLAB_BYTES, 1024, 2);      // MN-major: channel slabs -> LBO, 8 pixel rows -> SBO
			const uint32_t idesc = p.idesc;                                            // M = 256, N = 256
			for (int k = 0; k < n_steps; k++) {
				mbar_wait(full_bar(stage), phase);
				tc_fence_after();
				const uint32_t sa = smem_base + stage * stage_bytes;
				const uint64_t da_s = d_proto + (sa >> 4);
				uint64_t db_t = d_proto + ((sa + Cfg::A_BYTES) >> 4);
				const uint32_t accum = k != 0 ? 1u : 0u;
				uint32_t d_tmem = tmem_base;
				for (int ti = 0; ti < ntap; ti++) {
#pragma unroll
					for (int kk = 0; kk < Cfg::KPIX / 16; kk++)
						mma_f16_ss_pair_warp(d_tmem, da_s + ((kk * 2048) >> 4), db_t + ((kk * 2048) >> 4), idesc, kk != 0 ? 1u : accum);
					db_t += Cfg::B_BYTES >> 4;
					d_tmem += Cfg::BNC;
				}
				mma_commit_pair_warp(empty_bar(stage), (uint16_t)3);
				if (++stage == stages) { stage = 0; phase ^= 1u; }
			}
			mma_commit_pair_warp(done_bar, (uint16_t)3);
		}
	} else if (n_steps > 0) {
		const int quad = warp & 3;
		mbar_wait(done_bar, 0);
		tc_fence_after();
		const int f = (fg * 2 + (int)rank) * 128 + quad * 32 + lane;
		for (int ti = 0; ti < ntap; ti++) {
			const int tap = tap0 + ti;
			const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(ti * Cfg::BNC);
#pragma unroll 1
			for (int c0 = 0; c0 < Cfg::BNC; c0 += 32) {
				uint32_t r[32];
				tmem_ld_32x32(t_row + c0, r);
				tmem_ld_wait();
				if (f < p.out_c) {
					float* dst = p.grad + ((size_t)f * taps + tap) * p.in_cp + ct * Cfg::BNC + c0;
#pragma unroll
					for (int j = 0; j < 32; j += 4) {
						if (ct * Cfg::BNC + c0 + j < p.in_cp)
							asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(r[j])),
							             "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])), "f"(__uint_as_float(r[j + 3])) : "memory");
					}
				}
				__syncwarp();
			}
		}
	}
	tc_fence_before();
	__syncthreads();
	cluster_sync();
	if (warp == 1) { __syncwarp(); tc_fence_after(); tmem_dealloc_pair(tmem_base, p.tmem_cols); }
}

// ---------------------------------------------------------------- weight gradient of thin 3x3 layers, operands swapped
// For layers with few channels on large maps (Darknet19 layers 2, 3, 5: 32 -> 64 channels at 224 x 224, 64 -> 128 at
// 112 x 112) the kernel above is bound by shared-memory bandwidth, not by the tensor pipe: with M = output channels it
// issues one 128 x C x 16 MMA per tap and 16-pixel step, each re-reading the dy slab (for 64 filters half of its 128 rows
// are padding): 9 MMAs and 45 KB of shared-memory reads per 16 pixels for C = 32 (measured 24 % tensor activity).
// Here the roles are swapped: M = (filter tap, input channel) - the x tiles of consecutive taps sit one after the other
// in shared memory, so 128 / C of them are the channel slabs of ONE MN-major A operand (LBO = tile size; a slab past
// the CTA's last tap is whatever follows and is ignored) - and N = the real output channels.  C = 32: 3 MMAs of
// 128 x 64 x 16 and 18 KB of reads per 16 pixels (914 -> 414 us on layer 2).  C = 64: the 9 x 64 x 128 FP32 accumulators
// exceed TMEM, so taps 0-4 and 5-8 go to two groups of CTAs sized 3 : 2 like their MMA counts.
struct WswapParams {
	WgradParams g;               // pixel tiling, taps, gradient pointer (tiles_per_split / splits unused)
	int ctas0, tps0, tps1;       // CTAs [0, ctas0) take taps [0, taps0) with tps0 pixel tiles each, the rest taps [taps0, 9) with tps1
	int taps0;
};

template <int C, int NF>
struct WswapCfg {
	static constexpr int KPIX = 64;
	static constexpr int SLOTS = 128 / C;                     // taps per MMA
	static constexpr int MAXT = C == 32 ? 9 : 5;              // taps per CTA
	static constexpr int NACC = (MAXT + SLOTS - 1) / SLOTS;   // accumulators of NF columns
	static constexpr int DY_SLAB = KPIX * 128;                // [64 px][64 ch], 128B swizzle
	static constexpr int DY_BYTES = (NF / 64) * DY_SLAB;
	static constexpr int X_ROW = C * 2;
	static constexpr int X_TILE = KPIX * X_ROW;               // one tap: [64 px][C ch]
	static constexpr int STAGE_BYTES = DY_BYTES + MAXT * X_TILE;
	static constexpr int STAGES = C == 32 ? 4 : 3;
	static constexpr int SMEM_DATA = STAGES * STAGE_BYTES + SLOTS * X_TILE;   // + the ignored slabs behind the last tile
	static constexpr int SMEM_BYTES = ((SMEM_DATA + 1023) & ~1023) + 1024 + 256;
	static constexpr uint32_t X_LAYOUT = C == 64 ? 2u : 4u;
	static constexpr uint32_t TMEM_COLS = NACC * NF <= 256 ? 256 : 512;
	static_assert((C == 32 && NF == 64) || (C == 64 && NF == 128), "instantiated shapes");
	static_assert(STAGE_BYTES % 1024 == 0, "stage bases stay aligned for the 128B swizzle");
};

template <int C, int NF>
__global__ void __launch_bounds__(192, 1)
conv_wgrad_swap_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x, const WswapParams q) {
	using Cfg = WswapCfg<C, NF>;
	const WgradParams& p = q.g;
	extern __shared__ uint8_t smem_raw[];
	const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
	const uint32_t bar_base = smem_base + ((Cfg::SMEM_DATA + 1023) & ~1023);
	auto full_bar = [&](int s) { return bar_base + 8u * s; };
	auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
	const uint32_t done_bar = bar_base + 8u * (2 * Cfg::STAGES);
	const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::STAGES + 1);
	uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

	if (threadIdx.x == 0) {
		prefetch_tensormap(&tmap_dy);
		prefetch_tensormap(&tmap_x);
		for (int s = 0; s < Cfg::STAGES; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
		mbar_init(done_bar, 1);
		fence_barrier_init();
	}
	if (warp == 1) { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem_base = *tmem_slot_ptr;

	// this CTA's taps and pixel tiles
	const bool first = (int)blockIdx.x < q.ctas0;
	const int tap0 = C == 32 ? 0 : (first ? 0 : q.taps0);                       // (C = 32: one group of nine taps, all
	const int ntap = C == 32 ? 9 : (first ? q.taps0 : 9 - q.taps0);             //  loop bounds below fold to constants)
	const int tps = first ? q.tps0 : q.tps1;
	const int t_begin = (first ? (int)blockIdx.x : (int)blockIdx.x - q.ctas0) * tps;
	int t_end = t_begin + tps;
	if (t_end > p.pix_tiles) t_end = p.pix_tiles;
	const int n_steps = t_end > t_begin ? t_end - t_begin : 0;
	const int n_mma = (ntap + Cfg::SLOTS - 1) / Cfg::SLOTS;

	if (warp == 0) {
		{      // the whole warp, converged: elect.sync inside the asm picks the issuing lane
			int stage = 0; uint32_t phase = 0;
			const uint32_t tx_bytes = (uint32_t)(Cfg::DY_BYTES + ntap * Cfg::X_TILE);
			int tap_dx[Cfg::MAXT], tap_dy[Cfg::MAXT];       // pixel offset of this CTA's taps (registers once unrolled)
#pragma unroll
			for (int s = 0; s < Cfg::MAXT; s++) { const int tap = tap0 + s; tap_dy[s] = tap / 3 + p.off_h; tap_dx[s] = tap % 3 + p.off_w; }
			for (int t = t_begin; t < t_end; t++) {
				const int twi = t % p.tiles_w, thi = (t / p.tiles_w) % p.tiles_h, tni = t / (p.tiles_w * p.tiles_h);
				const int w0 = twi * p.tw, h0 = thi * p.th, n0 = tni * p.tn;
				mbar_wait(empty_bar(stage), phase ^ 1u);
				const uint32_t sd = smem_base + stage * Cfg::STAGE_BYTES, sx = sd + Cfg::DY_BYTES;
				mbar_arrive_expect_tx_warp(full_bar(stage), tx_bytes);
#pragma unroll
				for (int sl = 0; sl < NF / 64; sl++) tma_load_4d_warp(sd + sl * Cfg::DY_SLAB, &tmap_dy, full_bar(stage), sl * 64, w0, h0, n0);
#pragma unroll
				for (int s = 0; s < Cfg::MAXT; s++)
					if (s < ntap) tma_load_4d_warp(sx + s * Cfg::X_TILE, &tmap_x, full_bar(stage), 0, w0 + tap_dx[s], h0 + tap_dy[s], n0);
				if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
			}
		}
	} else if (warp == 1) {
		{      // the whole warp, converged: elect.sync inside the asm picks the issuing lane (see conv_halo_kernel)
			int stage = 0; uint32_t phase = 0;
			// both operands MN-major (pixels are the contraction index): 8 pixel rows per K group -> SBO, channel slabs -> LBO
			const uint64_t da_proto = make_smem_desc(0, Cfg::X_TILE, 8 * Cfg::X_ROW, Cfg::X_LAYOUT);
			const uint64_t db_proto = make_smem_desc(0, Cfg::DY_SLAB, 1024, 2);
			const uint32_t idesc = p.idesc;
			for (int k = 0; k < n_steps; k++) {
				mbar_wait(full_bar(stage), phase);
				tc_fence_after();
				const uint32_t sd = smem_base + stage * Cfg::STAGE_BYTES;
				const uint64_t db_s = db_proto + (sd >> 4);
				const uint64_t da_s = da_proto + ((sd + Cfg::DY_BYTES) >> 4);
				// (fully unrolled: the issuing thread is alone, every descriptor offset below is a compile-time constant)
#pragma unroll
				for (int kk = 0; kk < Cfg::KPIX / 16; kk++)
#pragma unroll
					for (int i = 0; i < Cfg::NACC; i++)
						if (i < n_mma)
							mma_f16_ss_warp(tmem_base + (uint32_t)(i * NF), da_s + ((i * Cfg::SLOTS * Cfg::X_TILE + kk * 16 * Cfg::X_ROW) >> 4),
							           db_s + ((kk * 16 * 128) >> 4), idesc, (k | kk) != 0 ? 1u : 0u);
				mma_commit_warp(empty_bar(stage));
				if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
			}
			mma_commit_warp(done_bar);
		}
	} else if (n_steps > 0) {
		const int quad = warp & 3;                // accumulator row quad*32 + lane = (tap slot row / C, channel row % C)
		mbar_wait(done_bar, 0);
		tc_fence_after();
		const int row = quad * 32 + lane, slot_in = row / C, c = row % C;
		for (int i = 0; i < n_mma; i++) {
			const int s = i * Cfg::SLOTS + slot_in;
#pragma unroll 1
			for (int c0 = 0; c0 < NF; c0 += 32) {
				uint32_t r[32];
				tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(i * NF + c0), r);
				tmem_ld_wait();
				if (s < ntap && c < p.in_cp) {
					float* dst = p.grad + (size_t)(tap0 + s) * p.in_cp + c;
#pragma unroll
					for (int j = 0; j < 32; j++)
						if (c0 + j < p.out_c) atomicAdd(dst + (size_t)(c0 + j) * 9 * p.in_cp, __uint_as_float(r[j]));
				}
				__syncwarp();
			}
		}
	}
	tc_fence_before();
	__syncthreads();
	if (warp == 1) { __syncwarp(); tc_fence_after(); tmem_dealloc(tmem_base, Cfg::TMEM_COLS); }
}

static bool wgrad_swap_ok(const cb200_conv_desc* d) {
	const int in_cp = round8(d->in_c), out_cp = round8(d->out_c);
	if (!(d->f_h == 3 && d->f_w == 3 && d->stride_h == 1 && d->stride_w == 1)) return false;
	if ((long long)d->batch * d->out_h * d->out_w < 64LL * 4 * g_num_sms) return false;      // enough pixel tiles to fill the machine
	return (in_cp >= 16 && in_cp <= 32 && out_cp >= 16 && out_cp <= 64) || (in_cp > 32 && in_cp <= 64 && out_cp > 64 && out_cp <= 128);
}

template <int C, int NF>
static int launch_wgrad_swap(const CUtensorMap& mdy, const CUtensorMap& mx, const WswapParams& q, int ctas, cudaStream_t st) {
	using Cfg = WswapCfg<C, NF>;
	static bool configured = false;
	auto kern = conv_wgrad_swap_kernel<C, NF>;
	if (!configured) {
		if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) != cudaSuccess) {
			set_error("cudaFuncSetAttribute(smem=%d) failed", Cfg::SMEM_BYTES); return CB200_ERR_CUDA;
		}
		configured = true;
	}
	kern<<<ctas, 192, Cfg::SMEM_BYTES, st>>>(mdy, mx, q);
	CB_LAUNCH_CHECK();
	return CB200_OK;
}

static int conv_wgrad_swap(const cb200_conv_desc* d, const cb200_conv_weights* w, const void* x, const void* dy, cudaStream_t st) {
	const int in_cp = round8(d->in_c), out_cp = round8(d->out_c);
	const bool wide = in_cp > 32;                 // C = 64, NF = 128, two tap groups
	WswapParams q;
	memset(&q, 0, sizeof(q));
	WgradParams& p = q.g;
	choose_rect(d->out_w, d->out_h, d->batch, 64, p.tw, p.th, p.tn);
	CUtensorMap mdy, mx;
	int rc = make_act_map(&mdy, dy, d->dtype, out_cp, d->out_w, d->out_h, d->batch, 64, p.tw, p.th, p.tn, CU_TENSOR_MAP_SWIZZLE_128B, 1);
	if (rc) return rc;
	rc = make_act_map(&mx, x, d->dtype, in_cp, d->in_w, d->in_h, d->batch, wide ? 64 : 32, p.tw, p.th, p.tn,
	                  wide ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, 1);
	if (rc) return rc;
	p.stride = 1;
	p.W = d->out_w; p.H = d->out_h; p.N = d->batch;
	p.tiles_w = ceil_div(p.W, p.tw); p.tiles_h = ceil_div(p.H, p.th); p.tiles_n = ceil_div(p.N, p.tn);
	p.pix_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
	p.f_h = 3; p.f_w = 3; p.off_h = -d->pad_h; p.off_w = -d->pad_w;
	p.out_c = d->out_c; p.in_cp = in_cp;
	p.grad = w->grad;
	p.idesc = make_idesc_f16(d->dtype == CB200_BF16, 128, wide ? 128 : 64, 1, 1);
	int ctas;
	if (!wide) {
		q.taps0 = 9;
		q.ctas0 = g_num_sms < p.pix_tiles ? g_num_sms : p.pix_tiles;
		q.tps0 = ceil_div(p.pix_tiles, q.ctas0);
		q.ctas0 = ceil_div(p.pix_tiles, q.tps0);
		q.tps1 = 0;
		ctas = q.ctas0;
	} else {
		// taps 0-4 (3 MMAs per step) and 5-8 (2 MMAs): CTAs in the same proportion, so both groups finish together
		q.taps0 = 5;
		int c0 = (g_num_sms * 3 + 2) / 5, c1 = g_num_sms - c0;
		if (c0 > p.pix_tiles) c0 = p.pix_tiles;
		if (c1 > p.pix_tiles) c1 = p.pix_tiles;
		q.tps0 = ceil_div(p.pix_tiles, c0); q.ctas0 = ceil_div(p.pix_tiles, q.tps0);
		q.tps1 = ceil_div(p.pix_tiles, c1);
		ctas = q.ctas0 + ceil_div(p.pix_tiles, q.tps1);
	}
	if (cudaMemsetAsync(w->grad, 0, sizeof(float) * (size_t)d->out_c * 9 * in_cp, st) != cudaSuccess) { set_error("wgrad memset failed"); return CB200_ERR_CUDA; }
	if (wide) return launch_wgrad_swap<64, 128>(mdy, mx, q, ctas, st);
	return launch_wgrad_swap<32, 64>(mdy, mx, q, ctas, st);
}

bool conv_tc_wgrad_supported(const cb200_conv_desc* d_in) {
	const cb200_conv_desc v = tc_view(d_in), *d = &v;
	if (!tc_common_ok(d)) return false;
	const int in_cp = round8(d->in_c), out_cp = round8(d->out_c);
	// (a 64-channel dy slab on a tensor with fewer channels relies on TMA zero-filling the missing ones; so does an x slab
	//  of 32 / 64 channels on a tensor with 24 / 40 / 48 / 56 - the epilogue only adds the real columns)
	return out_cp >= 16 && in_cp >= 16;
}

template <int BNC, int SLAB_C, int MF, int TGMAX>
static int launch_wgrad(const CUtensorMap& mdy, const CUtensorMap& mx, const WgradParams& p, dim3 grid, cudaStream_t st) {
	using Cfg = WgradCfg<BNC, SLAB_C>;
	static bool configured = false;
	auto kern = conv_wgrad_kernel<BNC, SLAB_C, MF, TGMAX>;
	if (!configured) {
		if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) != cudaSuccess) {
			set_error("cudaFuncSetAttribute(smem=%d) failed", Cfg::SMEM_BYTES); return CB200_ERR_CUDA;
		}
		configured = true;
	}
	kern<<<grid, 192, Cfg::SMEM_BYTES, st>>>(mdy, mx, p);
	CB_LAUNCH_CHECK();
	return CB200_OK;
}

int conv_wgrad_tc(const cb200_conv_desc* d_in, const cb200_conv_weights* w, const void* x, const void* dy, cudaStream_t st) {
	const cb200_conv_desc v = tc_view(d_in), *d = &v;
	static const bool no_swap = getenv("CB200_NO_WGRAD_SWAP") != nullptr;
	if (!no_swap && wgrad_swap_ok(d)) return conv_wgrad_swap(d, w, x, dy, st);
	const int in_cp = round8(d->in_c), out_cp = round8(d->out_c);
	const int taps = d->f_h * d->f_w;
	const int slab_c = in_cp > 32 ? 64 : (in_cp > 16 ? 32 : 16);
	const int bnc = in_cp > 128 ? 256 : (in_cp > 64 ? 128 : slab_c);
	// blocks of 128 output channels per CTA and taps per CTA, bounded by the 512 TMEM columns
	const int mf = (out_cp > 128 && bnc >= 128) ? 2 : 1;
	int tg_cap = 512 / (mf * bnc);
	if (bnc == 256 && mf == 1) tg_cap = 2;
	if (bnc == 128 && mf == 1) tg_cap = 3;
	if (bnc == 64) tg_cap = 5;
	if (bnc <= 32) tg_cap = 9;
	// wide layers: the two blocks of output channels on a CTA pair (conv_wgrad_pair_kernel), up to two taps per pair
	const bool pair = g_enable_pair_wgrad && bnc == 256 && mf == 2 && d->stride_w == 1;
	if (pair) tg_cap = 2;
	const int tg = taps < tg_cap ? taps : tg_cap;
	WgradParams p;
	memset(&p, 0, sizeof(p));
	choose_rect(d->out_w, d->out_h, d->batch, 64, p.tw, p.th, p.tn);
	CUtensorMap mdy, mx;
	int rc = make_act_map(&mdy, dy, d->dtype, out_cp, d->out_w, d->out_h, d->batch, 64, p.tw, p.th, p.tn, CU_TENSOR_MAP_SWIZZLE_128B, 1);
	if (rc) return rc;
	rc = make_act_map(&mx, x, d->dtype, in_cp, d->in_w, d->in_h, d->batch, slab_c, p.tw, p.th, p.tn, swizzle_for(slab_c), d->stride_w);
	if (rc) return rc;
	p.stride = d->stride_w;
	p.W = d->out_w; p.H = d->out_h; p.N = d->batch;
	p.tiles_w = ceil_div(p.W, p.tw); p.tiles_h = ceil_div(p.H, p.th); p.tiles_n = ceil_div(p.N, p.tn);
	p.pix_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
	p.f_h = d->f_h; p.f_w = d->f_w; p.off_h = -d->pad_h; p.off_w = -d->pad_w;
	p.f_groups = ceil_div(out_cp, 128 * mf); p.c_tiles = ceil_div(in_cp, bnc);
	p.tg = tg; p.tap_groups = ceil_div(taps, tg);
	const int stage_bytes = pair ? WgradPairCfg::A_BYTES + tg * WgradPairCfg::B_BYTES : mf * 16384 + tg * 64 * bnc * 2;
	p.stages = ((pair ? 192 : 196) * 1024) / stage_bytes;
	if (p.stages > 8) p.stages = 8;
	const int acc_cols = pair ? tg * bnc : mf * tg * bnc;
	p.tmem_cols = acc_cols <= 32 ? 32 : acc_cols <= 64 ? 64 : acc_cols <= 128 ? 128 : acc_cols <= 256 ? 256 : 512;
	const int jobs = p.f_groups * p.c_tiles * p.tap_groups;
	// Pixel-range splits per job.  These launches last 0.1-0.3 ms with one CTA (pair) per SM (pair of SMs), so whole waves
	// matter: take the split count that minimises waves x (steps per CTA x cost of a 64-pixel step + epilogue), with the
	// step bound by the tensor pipe (131 clk per 128 x 256 x 16 MMA) or by the operand bytes it pulls through L2 (~50 B/clk
	// per SM), and the FP32 red.add epilogue at ~30 B/clk.
	const int max_splits = ceil_div(p.pix_tiles, 8);
	int splits = 1;
	{
		const int slots = pair ? g_num_sms / 2 : g_num_sms;
		const double mma_clk = 4.0 * 131.0 * bnc / 256.0 * tg * (pair ? 1 : mf);
		const double step_clk = fmax(mma_clk, (double)stage_bytes / 50.0);
		const double epi_clk = (pair ? 1 : mf) * tg * 128.0 * bnc * 4.0 / 30.0 + 2000.0;
		double best = 1e30;
		for (int sp = 1; sp <= max_splits; sp++) {
			const int tps = ceil_div(p.pix_tiles, sp);
			const int real = ceil_div(p.pix_tiles, tps);
			if (real != sp) continue;
			const double waves = (double)ceil_div(jobs * sp, slots);
			const double cost = waves * (tps * step_clk + epi_clk);
			if (cost < best) { best = cost; splits = sp; }
		}
	}
	p.tiles_per_split = ceil_div(p.pix_tiles, splits);
	splits = ceil_div(p.pix_tiles, p.tiles_per_split);
	p.splits = splits;
	p.out_c = d->out_c; p.in_cp = in_cp;
	p.grad = w->grad;
	p.idesc = make_idesc_f16(d->dtype == CB200_BF16, pair ? 256 : 128, bnc, 1, 1);
	if (p.stages < 2 || acc_cols > 512) { set_error("conv_wgrad_tc: bad tiling (stages %d, tmem %d)", p.stages, acc_cols); return CB200_ERR_UNSUPPORTED; }
	if (cudaMemsetAsync(w->grad, 0, sizeof(float) * (size_t)d->out_c * taps * in_cp, st) != cudaSuccess) {
		set_error("wgrad memset failed"); return CB200_ERR_CUDA;
	}
	dim3 grid((unsigned)jobs, (unsigned)splits);
	if (pair) {
		static bool configured = false;
		if (!configured) {
			if (cudaFuncSetAttribute(conv_wgrad_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WgradPairCfg::SMEM_BYTES) != cudaSuccess) {
				set_error("cudaFuncSetAttribute(smem=%d) failed", WgradPairCfg::SMEM_BYTES); return CB200_ERR_CUDA;
			}
			configured = true;
		}
		cudaLaunchConfig_t cfg;
		memset(&cfg, 0, sizeof(cfg));
		cudaLaunchAttribute attr[1];
		attr[0].id = cudaLaunchAttributeClusterDimension;
		attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
		cfg.gridDim = dim3((unsigned)(2 * jobs), (unsigned)splits); cfg.blockDim = dim3(192);
		cfg.dynamicSmemBytes = WgradPairCfg::SMEM_BYTES; cfg.stream = st; cfg.attrs = attr; cfg.numAttrs = 1;
		if (cudaLaunchKernelEx(&cfg, conv_wgrad_pair_kernel, mdy, mx, p) != cudaSuccess) {
			set_error("cluster launch of conv_wgrad_pair_kernel failed: %s", cudaGetErrorString(cudaGetLastError())); return CB200_ERR_CUDA;
		}
		g_launches++;
		g_last_conv_impl = "tcgen05-pair";
		return CB200_OK;
	}
	if (bnc == 256 && mf == 2) return launch_wgrad<256, 64, 2, 1>(mdy, mx, p, grid, st);
	if (bnc == 256) return launch_wgrad<256, 64, 1, 2>(mdy, mx, p, grid, st);
	if (bnc == 128 && mf == 2) return launch_wgrad<128, 64, 2, 2>(mdy, mx, p, grid, st);
	if (bnc == 128) return launch_wgrad<128, 64, 1, 3>(mdy, mx, p, grid, st);
	if (bnc == 64) return launch_wgrad<64, 64, 1, 5>(mdy, mx, p, grid, st);
	if (bnc == 32) return launch_wgrad<32, 32, 1, 9>(mdy, mx, p, grid, st);
	return launch_wgrad<16, 16, 1, 9>(mdy, mx, p, grid, st);
}

}  // namespace cb200
