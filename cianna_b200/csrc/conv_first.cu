// conv_first.cu - the network's first convolution, straight from the dataset batch, on the tensor cores.
//
// A first layer has 1-3 input channels: its GEMM K is tiny (3x3x3+1 = 28) and both passes are bound by memory traffic.
// Upstream unrolls the receptive fields to memory (im2col_kernel, src/cuda/cuda_conv_layer.cu:36-103: 28 values written
// and read back per output pixel) and so did this core's first version (cb200_import_input_patches: 1.6 GB written + read
// per pass at batch 128, 448 px).  Here the unrolled rows only ever exist in SHARED memory: builder warps read the planar
// dataset rows [B][C*H*W+1] (L1/L2 resident: every input value is used by f_h*f_w neighbouring pixels), assemble the GEMM
// operand tile in the canonical swizzled layout the tensor core expects - byte for byte what a TMA load of the
// materialised patch rows would have produced - and hand it to the MMA warp through an mbarrier (generic-proxy writes
// are made visible to the tensor core's async proxy with fence.proxy.async).  HBM traffic per pass drops to "read the
// images once, write (forward) or read (weight gradient) the layer's output once".
//
//   forward : D[128 px][BN filters] = patch[128 px][KP] * W[filters][KP]^T     (K-major A built on chip, B by TMA once)
//             16 builder warps (four groups taking tiles in turn), 1 MMA warp, 8 epilogue warps (two groups), TMEM accumulator ring
//   wgrad   : G[filters][KP] += dy[px][filters]^T * patch[px][KP]               (both MN-major; dy by TMA, patches built)
//             persistent CTAs over contiguous pixel ranges, FP32 red.add of the per-CTA partial at the end
// Column order of a patch row = upstream's filter column order c*taps + tap, then the bias input, zero padded to KP.
#include <cuda.h>
#include <cstdlib>
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace cb200 {

using namespace ptx;

int make_act_map(CUtensorMap* m, const void* base, int dtype, int cp, int w, int h, int n, int bc, int bw, int bh, int bn, CUtensorMapSwizzle sw, int pix_stride);
int make_w_map(CUtensorMap* m, const void* base, int dtype, int cp, int taps, int rows, int bc, int brow, CUtensorMapSwizzle sw);
void choose_rect(int W, int H, int N, int npix, int& tw, int& th, int& tn);
CUtensorMapSwizzle swizzle_for(int bk);

struct FirstParams {
	const void* src;             // dataset batch, [N][c*h*w + 1] values of the compute type
	int c, h, w, s_h, s_w, p_h, p_w;
	int W, H, N;                 // output pixel grid
	int tw, th, tn, tiles_w, tiles_h, tiles_n, num_tiles;
	int n_real, n_pad, length;
	float bias_value;
	cb200_activ activ;
	void* out;                   // forward: layer output [N][H][W][n_pad]
	float* grad;                 // wgrad: [n_real][KP]
	int tiles_per_cta;           // contiguous run of tiles owned by one CTA
	uint32_t idesc;
	int tma_store;               // forward: the output tile leaves through shared memory + one TMA store per tile
	int prefetch;                // builders prefetch the next tile's receptive fields into L1 (CB200_FIRST_PREFETCH=0: off)
	// weight gradient with <= 32 filters: dy arrives as ONE [64 px][32 filters] box (64-byte rows, 64B swizzle, 4 KB) instead
	// of two 64-filter boxes of which three quarters are zero fill (16 KB written per step), and the M = 128 MMA reads
	// that 32-row atom four times (leading byte offset 0: accumulator rows 32..127 repeat rows 0..31 and are never read)
	int wg_narrow;
	// forward: group-norm sums of the layer output out of the epilogue (conv_first_forward): FP64 workspace
	// [N][gn_groups][2] = (sum, sum of squares) of the group-norm layer that follows, its group size and group count
	double* gn_ws;
	int gn_groups;
	uint32_t gn_map;             // group of columns 4i..4i+3 in nibble i (i = 0..7)
};

template <int KP> struct PatchCfg {
	static constexpr int ROW_BYTES = KP * 2;
	static constexpr uint32_t LAYOUT = KP == 64 ? 2u : (KP == 32 ? 4u : 6u);           // 128B / 64B / 32B swizzle
	static constexpr int SW_SHIFT = KP == 64 ? 0 : (KP == 32 ? 1 : 2);                  // row bits that feed the XOR
	static constexpr int SW_MASK = KP / 8 - 1;
	static constexpr uint32_t SBO = 8 * ROW_BYTES;
};

// Assemble the patch row of output pixel (py, px) of image `img` and store it as row `row` of a tile at smem address
// `tile` (16-byte chunks XOR-swizzled exactly like CU_TENSOR_MAP_SWIZZLE_{128,64,32}B does for rows of KP*2 bytes).
// nxt != nullptr: the receptive field of the pixel this thread builds NEXT (same position in the group's next tile) is
// prefetched into L1 right behind this row's own loads.  The builders are bound by the latency of those loads, not by
// their number (ncu source view of the forward kernel: 80 % of the builder warps' samples wait on the first use of a
// loaded value, while half of the epilogue warps' samples wait for accumulators), and a row's lines come from L2 /
// HBM the first time they are touched; a prefetch needs no register.
template <typename T, int C, int FH, int FW, int KP>
__device__ __forceinline__ void build_patch_row(const T* __restrict__ img, bool valid, int h, int w, int iy0, int ix0,
                                                unsigned short bias_bits, uint32_t tile, int row,
                                                const T* __restrict__ nxt = nullptr, int niy0 = 0, int nix0 = 0) {
	using PC = PatchCfg<KP>;
	constexpr int TAPS = FH * FW, KREAL = C * TAPS;
	static_assert(KREAL + 1 <= KP, "patch row does not fit");
	const unsigned short* __restrict__ src = reinterpret_cast<const unsigned short*>(img);
	uint32_t packed[KP / 2];
#pragma unroll
	for (int i = 0; i < KP / 2; i++) packed[i] = 0u;
	if (valid && iy0 >= 0 && iy0 + FH <= h && ix0 >= 0 && ix0 + FW <= w) {
		// interior pixel (all but the image border): every tap is inside, plain loads at fixed offsets from one pointer
		// per (channel, filter row) - the builders are instruction-bound, so this path carries no predicates
		// (32-bit element offsets inside one image - c*h*w < 2^31 - keep the nine line pointers at one IMAD.WIDE each)
		// ALL loads first, then the packing: with load and use interleaved in the source the compiler kept only a few
		// loads in flight (ncu: 16 long-scoreboard stall cycles per issued instruction in the weight-gradient kernel, i.e.
		// several L2 round trips per patch row instead of one)
		const int off0 = iy0 * w + ix0;
		const int chs = h * w;
		unsigned short v[KREAL];
#pragma unroll
		for (int ch = 0; ch < C; ch++) {
#pragma unroll
			for (int ky = 0; ky < FH; ky++) {
				const unsigned short* __restrict__ line = src + (off0 + ch * chs + ky * w);
#pragma unroll
				for (int kx = 0; kx < FW; kx++) {
					unsigned short t;
					asm volatile("ld.global.nc.u16 %0, [%1];" : "=h"(t) : "l"(line + kx));
					v[(ch * FH + ky) * FW + kx] = t;
				}
			}
		}
		if (nxt != nullptr && niy0 >= 0 && niy0 + FH <= h && nix0 >= 0 && nix0 + FW <= w) {
			const unsigned short* __restrict__ nsrc = reinterpret_cast<const unsigned short*>(nxt) + (niy0 * w + nix0);
#pragma unroll
			for (int ch = 0; ch < C; ch++)
#pragma unroll
				for (int ky = 0; ky < FH; ky++)
					asm volatile("prefetch.global.L1 [%0];" ::"l"(nsrc + (ch * chs + ky * w)));
		}
#pragma unroll
		for (int k = 0; k < KREAL; k++) packed[k >> 1] |= (uint32_t)v[k] << (16 * (k & 1));
	} else {
		bool col_ok[FW];
#pragma unroll
		for (int kx = 0; kx < FW; kx++) col_ok[kx] = valid && (ix0 + kx) >= 0 && (ix0 + kx) < w;
#pragma unroll
		for (int ch = 0; ch < C; ch++) {
#pragma unroll
			for (int ky = 0; ky < FH; ky++) {
				const int iy = iy0 + ky;
				const bool row_ok = iy >= 0 && iy < h;
				const unsigned short* __restrict__ line = src + ((size_t)ch * h + (row_ok ? iy : 0)) * w + ix0;
#pragma unroll
				for (int kx = 0; kx < FW; kx++) {
					const int k = (ch * FH + ky) * FW + kx;
					const uint32_t v = (row_ok && col_ok[kx]) ? (uint32_t)__ldg(line + kx) : 0u;
					packed[k >> 1] |= v << (16 * (k & 1));
				}
			}
		}
	}
	if (valid) packed[KREAL >> 1] |= (uint32_t)bias_bits << (16 * (KREAL & 1));
	const uint32_t base = tile + (uint32_t)row * PC::ROW_BYTES;
	const int x = (row >> PC::SW_SHIFT) & PC::SW_MASK;
#pragma unroll
	for (int j = 0; j < KP / 8; j++) {
		const uint32_t dst = base + (uint32_t)((j ^ x) << 4);
		asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(packed[4 * j]), "r"(packed[4 * j + 1]),
		             "r"(packed[4 * j + 2]), "r"(packed[4 * j + 3]) : "memory");
	}
}

template <typename T> __device__ __forceinline__ unsigned short bits_of(float v);
template <> __device__ __forceinline__ unsigned short bits_of<__half>(float v) { return __half_as_ushort(__float2half_rn(v)); }
template <> __device__ __forceinline__ unsigned short bits_of<__nv_bfloat16>(float v) { return __bfloat16_as_ushort(__float2bfloat16_rn(v)); }

// ================================================================ forward
// ncu (stall sampling) on the first version showed the builders waiting for free stages 30 % of the time and the
// epilogue groups waiting for accumulators: the limiter is the LATENCY of one tile's epilogue (tcgen05.ld -> activation ->
// stores is a ~2000-cycle dependent chain for a warp sharing its scheduler with six others), so tiles in flight in the
// epilogue matter more than builder warps: 2 builder groups, 4 epilogue groups (= accumulator stages)
#ifndef CB200_FWD_BUILD_GROUPS
#define CB200_FWD_BUILD_GROUPS 2
#define CB200_FWD_EPI_GROUPS 4
#endif
constexpr int FWD_BUILD_GROUPS = CB200_FWD_BUILD_GROUPS, FWD_BUILD_WARPS = 4 * FWD_BUILD_GROUPS, FWD_EPI_GROUPS = CB200_FWD_EPI_GROUPS;
constexpr int FWD_THREADS = (FWD_BUILD_WARPS + 1 + 4 * FWD_EPI_GROUPS) * 32;
template <int KP, int BN> struct FirstFwdCfg {
	static constexpr int A_BYTES = 128 * KP * 2;
	static constexpr int STAGES = 8;
	static constexpr int B_BYTES = BN * KP * 2;
	static constexpr int ACC_STAGES = 4;
	static constexpr int TMEM_COLS = ACC_STAGES * BN <= 128 ? 128 : 256;
	static constexpr int OUT_TILE_BYTES = 128 * BN * 2;                 // staging tile of one epilogue group (TMA store)
	static constexpr int SMEM_BYTES = STAGES * A_BYTES + ((B_BYTES + 1023) & ~1023) + 1024 + 1024 + FWD_EPI_GROUPS * OUT_TILE_BYTES;
};

template <typename T, int C, int FH, int FW, int KP, int BN, bool GN>
__global__ void __launch_bounds__(FWD_THREADS, 1)
conv_first_fwd_kernel(const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_out, const FirstParams p) {
	using Cfg = FirstFwdCfg<KP, BN>;
	using PC = PatchCfg<KP>;
	extern __shared__ uint8_t smem_raw[];
	const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
	const uint32_t b_smem = smem_base + Cfg::STAGES * Cfg::A_BYTES;
	const uint32_t bar_base = b_smem + ((Cfg::B_BYTES + 1023) & ~1023);
	auto full_bar = [&](int s) { return bar_base + 8u * s; };
	auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
	auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + s); };
	auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + 4 + s); };
	const uint32_t bfull_bar = bar_base + 8u * (2 * Cfg::STAGES + 8);
	const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::STAGES + 9);
	const uint32_t out_smem = bar_base + 1024u;          // FWD_EPI_GROUPS staging tiles (1024-byte aligned like everything above)
	uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	constexpr int MMA_WARP = FWD_BUILD_WARPS;

	if (threadIdx.x == 0) {
		prefetch_tensormap(&tmap_b);
		if (p.tma_store) prefetch_tensormap(&tmap_out);
		for (int s = 0; s < Cfg::STAGES; s++) { mbar_init(full_bar(s), 4); mbar_init(empty_bar(s), 1); }
		for (int s = 0; s < Cfg::ACC_STAGES; s++) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 4); }
		mbar_init(bfull_bar, 1);
		fence_barrier_init();
	}
	if (warp == MMA_WARP) { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem_base = *tmem_slot_ptr;
	// a CTA owns a contiguous run of tiles: neighbouring tiles share input rows, which keeps them in L1
	const int tile0 = blockIdx.x * p.tiles_per_cta;
	const int n_tiles = min(p.tiles_per_cta, p.num_tiles - tile0);

	if (warp < FWD_BUILD_WARPS) {
		// ===================== builders: group g fills the stages of the tiles g, g+4, ... of the run =====================
		const int grp = warp >> 2;
		const int row = (warp & 3) * 32 + lane;
		const T* __restrict__ src = reinterpret_cast<const T*>(p.src);
		const size_t img_stride = (size_t)p.c * p.h * p.w + 1;
		const unsigned short bias_bits = bits_of<T>(p.bias_value);
		const int rx = row % p.tw, ry = (row / p.tw) % p.th, rn = row / (p.tw * p.th);
		// tile coordinates advance incrementally (one division at the start instead of three per tile)
		const int tiles_w = p.tiles_w, tiles_h = p.tiles_h;
		const bool prefetch_next = p.prefetch != 0;
		int twi = (tile0 + grp) % tiles_w, thi = ((tile0 + grp) / tiles_w) % tiles_h, tni = (tile0 + grp) / (tiles_w * tiles_h);
		for (int it = grp; it < n_tiles; it += FWD_BUILD_GROUPS) {
			const int stage = it % Cfg::STAGES;
			const uint32_t phase = (uint32_t)(it / Cfg::STAGES) & 1u;
			const int px = twi * p.tw + rx, py = thi * p.th + ry, pn = tni * p.tn + rn;
			twi += FWD_BUILD_GROUPS;
			while (twi >= tiles_w) { twi -= tiles_w; if (++thi == tiles_h) { thi = 0; tni++; } }
			const bool valid = px < p.W && py < p.H && pn < p.N;
			// this thread's pixel in the group's next tile (twi / thi / tni already point there)
			const int npx = twi * p.tw + rx, npy = thi * p.th + ry, npn = tni * p.tn + rn;
			const bool nvalid = prefetch_next && it + FWD_BUILD_GROUPS < n_tiles && npx < p.W && npy < p.H && npn < p.N;
			mbar_wait(empty_bar(stage), phase ^ 1u);
			build_patch_row<T, C, FH, FW, KP>(src + (size_t)(valid ? pn : 0) * img_stride, valid, p.h, p.w,
				py * p.s_h - p.p_h, px * p.s_w - p.p_w, bias_bits, smem_base + stage * Cfg::A_BYTES, row,
				nvalid ? src + (size_t)npn * img_stride : nullptr, npy * p.s_h - p.p_h, npx * p.s_w - p.p_w);
			fence_proxy_async();          // each writer publishes its row to the async proxy ...
			__syncwarp();
			if (lane == 0) mbar_arrive(full_bar(stage));     // ... then one arrival per warp
		}
	} else if (warp == MMA_WARP) {
		// ===================== MMA issuer (also fetches the filters once) =====================
		if (lane == 0) {
			mbar_arrive_expect_tx(bfull_bar, Cfg::B_BYTES);
			tma_load_3d(b_smem, &tmap_b, bfull_bar, 0, 0, 0);
		}
		__syncwarp();
		{      // the whole warp, converged: elect.sync inside the asm picks the issuing lane (conv_tc.cu: conv_halo_kernel)
			mbar_wait(bfull_bar, 0);
			for (int it = 0; it < n_tiles; it++) {
				const int stage = it % Cfg::STAGES, acc = it % Cfg::ACC_STAGES;
				const uint32_t phase = (uint32_t)(it / Cfg::STAGES) & 1u, acc_phase = (uint32_t)(it / Cfg::ACC_STAGES) & 1u;
				mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
				mbar_wait(full_bar(stage), phase);
				tc_fence_after();
				const uint32_t sa = smem_base + stage * Cfg::A_BYTES;
				const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
#pragma unroll
				for (int kk = 0; kk < KP / 16; kk++) {
					const uint64_t da = make_smem_desc(sa + kk * 32, 16, PC::SBO, PC::LAYOUT);
					const uint64_t db = make_smem_desc(b_smem + kk * 32, 16, PC::SBO, PC::LAYOUT);
					mma_f16_ss_warp(d_tmem, da, db, p.idesc, kk != 0 ? 1u : 0u);
				}
				mma_commit_warp(empty_bar(stage));
				mma_commit_warp(tfull_bar(acc));
			}
		}
	} else {
		// ===================== epilogue: activation, cast, store (the bias is a GEMM column) =====================
		const int quad = warp & 3;
		const int egrp = (warp - MMA_WARP - 1) >> 2;
		const int row = quad * 32 + lane;
		T* __restrict__ out = reinterpret_cast<T*>(p.out);
		const int act = p.activ.type, n_real = p.n_real, n_pad = p.n_pad;
		const float leak = p.activ.leak, sat = p.activ.saturation, beta = p.activ.beta;
		const bool mask_tail = act == CB200_RELU || act == CB200_LOGISTIC || act == CB200_SOFTMAX;
		// 0 <= leak <= 1, sat >= 0: z <= 0 ? z*leak : (z > sat ? hi : z) == min(max(z, z*leak), hi) value for value (hi >= z
		// below the saturation, hi < z above it) - two FMNMX instead of two compares and two selects per element: this
		// epilogue is bound by the ALU pipe (profiles/r1_conv_first_fwd_full_raw.csv: 65 % busy)
		const bool relu_minmax = leak >= 0.0f && leak <= 1.0f && sat >= 0.0f;
		const float sat_c = sat - sat * leak;
		const int rx = row % p.tw, ry = (row / p.tw) % p.th, rn = row / (p.tw * p.th);
		if (p.tma_store) {
			// Output through shared memory: with one pixel row per thread a 128-bit global store of a warp touches 32
			// different lines (16 of them for 32 filters) - the epilogue stores were 60 % of this kernel's LSU wavefronts,
			// its busiest unit (profiles/r1_first_halo_full_digest.txt).  Here each thread writes its packed row into the
			// group's staging tile (the TMA swizzle keeps the 128-bit shared stores conflict-free), and one thread hands the
			// tile to the TMA unit, which also clips rows outside the tensor.
			const int gtid = (warp - MMA_WARP - 1 - 4 * egrp) * 32 + lane;          // 0..127 inside the group
			const uint32_t stg = out_smem + (uint32_t)egrp * Cfg::OUT_TILE_BYTES;
			uint8_t* stg_ptr = smem_raw + (stg - smem_u32(smem_raw));
			constexpr int CHUNKS = BN / 8;                                         // 16-byte chunks per row: 4 (64B swizzle) or 8 (128B)
			const int sw_x = BN == 32 ? ((row >> 1) & 3) : (row & 7);
			// tile coordinates advance incrementally (three divisions by run-time values per tile were 15 % of this warp's
			// instructions, and the kernel issues at 67 % of its slots once the builders prefetch)
			const int tiles_w = p.tiles_w, tiles_h = p.tiles_h;
			int twi = (tile0 + egrp) % tiles_w, thi = ((tile0 + egrp) / tiles_w) % tiles_h, tni = (tile0 + egrp) / (tiles_w * tiles_h);
			// Group-norm sums of the layer that follows (GN; every warp's 32 rows in ONE sample, BN = 32): the statistics pass over
			// this layer's output is the largest of the network (1.6 GB at batch 128).  Every thread keeps (sum, sum of
			// squares) of its pixel's columns 4i..4i+3, i = 0..7, as FP32 pairs fed by FADD2 / FFMA2 (one instruction per
			// value); a warp folds them every 32 tiles or when the sample changes: butterfly over the lanes, then one FP64
			// atomic per sum into the workspace of norm_stats_kernel (any group size that is a multiple of 4).  The values
			// are the activated FP32 ones before the rounding to 16 bit (the sums differ from those of the stored tensor by
			// the mean rounding error: < 1e-6 relative in FP16, < 1e-5 in BF16).
			float gsum[GN ? BN / 4 : 1], gsq[GN ? BN / 4 : 1];
#pragma unroll
			for (int i = 0; i < (GN ? BN / 4 : 1); i++) { gsum[i] = 0.0f; gsq[i] = 0.0f; }
			int gn_pn = -1, gn_cnt = 0;
			auto gn_flush = [&]() {
#pragma unroll
				for (int i = 0; i < (GN ? BN / 4 : 1); i++) {
					float a = gsum[i], b = gsq[i];
#pragma unroll
					for (int m = 16; m > 0; m >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, m); b += __shfl_xor_sync(0xffffffffu, b, m); }
					const int g = (int)((p.gn_map >> (4 * i)) & 15u);
					if ((lane >> 1) == i && g < p.gn_groups) {
						const float v = (lane & 1) ? b : a;
						if (v != 0.0f) atomicAdd(p.gn_ws + ((size_t)gn_pn * p.gn_groups + g) * 2 + (lane & 1), (double)v);
					}
					gsum[i] = 0.0f; gsq[i] = 0.0f;
				}
			};
			for (int it = egrp; GN || it < n_tiles; it += FWD_EPI_GROUPS) {
				const int acc = it % Cfg::ACC_STAGES;
				const uint32_t acc_phase = (uint32_t)(it / Cfg::ACC_STAGES) & 1u;
				const int c_twi = twi, c_thi = thi, c_tni = tni;
				twi += FWD_EPI_GROUPS;
				while (twi >= tiles_w) { twi -= tiles_w; if (++thi == tiles_h) { thi = 0; tni++; } }
				const int pn = c_tni * p.tn + rn;
				const bool dead = mask_tail && pn >= p.length;
				bool gn_row = false;
				if (GN) {      // (one flush site: the pass after the last tile only folds what is left)
					const bool past = it >= n_tiles;
					if (past || pn != gn_pn || gn_cnt == 32) {       // (pn is the same for the 32 rows of a warp: tw * th % 32 == 0)
						if (gn_pn >= 0 && gn_pn < p.N) gn_flush();
						gn_pn = pn; gn_cnt = 0;
					}
					if (past) break;
					gn_cnt++;
					gn_row = c_twi * p.tw + rx < p.W && c_thi * p.th + ry < p.H && pn < p.N;
				}
				mbar_wait(tfull_bar(acc), acc_phase);
				tc_fence_after();
				if (gtid == 0) bulk_wait_read0();                                    // the previous tile of this group has left the buffer
				asm volatile("bar.sync %0, 128;" ::"r"(1 + egrp) : "memory");
				const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
				for (int c0 = 0; c0 < BN; c0 += 32) {
					uint32_t r[32];
					tmem_ld_32x32(t_row + c0, r);
					tmem_ld_wait();
#pragma unroll
					for (int v = 0; v < 4; v++) {
						const int col = c0 + v * 8;
						float o[8];
#pragma unroll
						for (int j = 0; j < 8; j++) o[j] = __uint_as_float(r[v * 8 + j]);
						if (dead) {
#pragma unroll
							for (int j = 0; j < 8; j++) o[j] = 0.0f;
						} else if (act == CB200_RELU) {
							if (relu_minmax) {
								// packed pairs (sm100_ptx.cuh: leaky_sat_f32x2): 3 instead of 5 issue slots per value in the warps
								// that bound this kernel
#pragma unroll
								for (int j = 0; j < 8; j += 2) leaky_sat_f32x2(o[j], o[j + 1], leak, sat_c);
							} else {
#pragma unroll
								for (int j = 0; j < 8; j++) { const float z = o[j]; const float hi = sat + (z - sat) * leak; o[j] = z <= 0.0f ? z * leak : (z > sat ? hi : z); }
							}
						} else if (act == CB200_LOGISTIC) {
#pragma unroll
							for (int j = 0; j < 8; j++) o[j] = 1.0f / (1.0f + expf(fminf(-beta * o[j], sat)));
						}
						if (col + 8 > n_real) {
#pragma unroll
							for (int j = 0; j < 8; j++) if (col + j >= n_real) o[j] = 0.0f;
						}
						if (GN && gn_row) {
							static_assert(!GN || BN == 32, "group-norm sums: one 32-column pass per tile");
							const int h = 2 * v;                               // BN = 32: one pass of the c0 loop, columns 8v..8v+7
#pragma unroll
							for (int j = 0; j < 4; j++) {
								add_f32x2(gsum[GN ? h : 0], gsum[GN ? h + 1 : 0], o[j], o[j + 4]);
								sq_acc_f32x2(gsq[GN ? h : 0], gsq[GN ? h + 1 : 0], o[j], o[j + 4]);
							}
						}
						const int chunk = ((col >> 3) ^ sw_x) & (CHUNKS - 1);
						store8<T>(reinterpret_cast<T*>(stg_ptr + row * (BN * 2) + chunk * 16), o);
					}
					__syncwarp();
				}
				tc_fence_before();
				__syncwarp();
				if (lane == 0) mbar_arrive(tempty_bar(acc));
				fence_proxy_async();                                                 // generic-proxy writes -> visible to the TMA unit
				asm volatile("bar.sync %0, 128;" ::"r"(1 + egrp) : "memory");
				if (gtid == 0) {
					tma_store_4d(&tmap_out, stg, 0, c_twi * p.tw, c_thi * p.th, c_tni * p.tn);
					bulk_commit();
				}
			}
			if (gtid == 0) bulk_wait0();
		} else
		{
		const int tiles_w = p.tiles_w, tiles_h = p.tiles_h;
		int twi = (tile0 + egrp) % tiles_w, thi = ((tile0 + egrp) / tiles_w) % tiles_h, tni = (tile0 + egrp) / (tiles_w * tiles_h);
		for (int it = egrp; it < n_tiles; it += FWD_EPI_GROUPS) {
			const int acc = it % Cfg::ACC_STAGES;
			const uint32_t acc_phase = (uint32_t)(it / Cfg::ACC_STAGES) & 1u;
			const int px = twi * p.tw + rx, py = thi * p.th + ry, pn = tni * p.tn + rn;
			twi += FWD_EPI_GROUPS;
			while (twi >= tiles_w) { twi -= tiles_w; if (++thi == tiles_h) { thi = 0; tni++; } }
			const bool row_ok = px < p.W && py < p.H && pn < p.N;
			const size_t pix = ((size_t)pn * p.H + py) * p.W + px;
			const bool dead = mask_tail && pn >= p.length;
			mbar_wait(tfull_bar(acc), acc_phase);
			tc_fence_after();
			const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
			for (int c0 = 0; c0 < BN; c0 += 32) {
				uint32_t r[32];
				tmem_ld_32x32(t_row + c0, r);
				tmem_ld_wait();
#pragma unroll
				for (int v = 0; v < 4; v++) {
					const int col = c0 + v * 8;
					if (!row_ok || col >= n_pad) continue;
					float o[8];
#pragma unroll
					for (int j = 0; j < 8; j++) o[j] = __uint_as_float(r[v * 8 + j]);
					if (dead) {
#pragma unroll
						for (int j = 0; j < 8; j++) o[j] = 0.0f;
					} else if (act == CB200_RELU) {
						if (relu_minmax) {
#pragma unroll
							for (int j = 0; j < 8; j++) {
								const float z = o[j];
								const float hi = sat + (z - sat) * leak;
								o[j] = fminf(fmaxf(z, z * leak), hi);
							}
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
						for (int j = 0; j < 8; j++) o[j] = dead ? 0.0f : 1.0f / (1.0f + expf(fminf(-beta * o[j], sat)));
					}
					if (col + 8 > n_real) {
#pragma unroll
						for (int j = 0; j < 8; j++) if (col + j >= n_real) o[j] = 0.0f;
					}
					store8<T>(out + pix * n_pad + col, o);
				}
				__syncwarp();
			}
			tc_fence_before();
			__syncwarp();
			if (lane == 0) mbar_arrive(tempty_bar(acc));
		}
		}
	}
	tc_fence_before();
	__syncthreads();
	if (warp == MMA_WARP) { __syncwarp(); tc_fence_after(); tmem_dealloc(tmem_base, Cfg::TMEM_COLS); }
}

// ================================================================ weight gradient
constexpr int WG_BUILD_WARPS = 16, WG_THREADS = (2 + 4 + WG_BUILD_WARPS) * 32;
template <int KP> struct FirstWgCfg {
	static constexpr int KPIX = 64;
	static constexpr int A_SLAB_BYTES = KPIX * 128;                 // dy: [64 px][64 filters], 128B swizzle
	static constexpr int A_BYTES = 2 * A_SLAB_BYTES;                // M = 128 filter rows (rows past the tensor are zero-filled)
	static constexpr int B_BYTES = KPIX * KP * 2;
	static constexpr int STAGE_BYTES = A_BYTES + ((B_BYTES + 1023) & ~1023);
	static constexpr int STAGES = 8;
	static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
	static constexpr int TMEM_COLS = 32 > KP ? 32 : KP;
};

template <typename T, int C, int FH, int FW, int KP>
__global__ void __launch_bounds__(WG_THREADS, 1)
conv_first_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_dy, const FirstParams p) {
	using Cfg = FirstWgCfg<KP>;
	using PC = PatchCfg<KP>;
	extern __shared__ uint8_t smem_raw[];
	const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
	const uint32_t bar_base = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
	auto full_bar = [&](int s) { return bar_base + 8u * s; };
	auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
	const uint32_t done_bar = bar_base + 8u * (2 * Cfg::STAGES);
	const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::STAGES + 1);
	uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

	if (threadIdx.x == 0) {
		prefetch_tensormap(&tmap_dy);
		for (int s = 0; s < Cfg::STAGES; s++) { mbar_init(full_bar(s), 1 + Cfg::KPIX / 32); mbar_init(empty_bar(s), 1); }
		mbar_init(done_bar, 1);
		fence_barrier_init();
	}
	if (warp == 1) { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem_base = *tmem_slot_ptr;

	const int t_begin = blockIdx.x * p.tiles_per_cta;
	int t_end = t_begin + p.tiles_per_cta;
	if (t_end > p.num_tiles) t_end = p.num_tiles;
	const int n_steps = t_end > t_begin ? t_end - t_begin : 0;

	if (warp == 0) {
		{      // the whole warp, converged: elect.sync inside the asm picks the issuing lane
			for (int k = 0; k < n_steps; k++) {
				const int t = t_begin + k;
				const int stage = k % Cfg::STAGES;
				const uint32_t phase = (uint32_t)(k / Cfg::STAGES) & 1u;
				const int twi = t % p.tiles_w, thi = (t / p.tiles_w) % p.tiles_h, tni = t / (p.tiles_w * p.tiles_h);
				mbar_wait(empty_bar(stage), phase ^ 1u);
				const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
				mbar_arrive_expect_tx_warp(full_bar(stage), p.wg_narrow ? Cfg::KPIX * 64 : Cfg::A_BYTES);
				tma_load_4d_warp(sa, &tmap_dy, full_bar(stage), 0, twi * p.tw, thi * p.th, tni * p.tn);
				if (!p.wg_narrow) tma_load_4d_warp(sa + Cfg::A_SLAB_BYTES, &tmap_dy, full_bar(stage), 64, twi * p.tw, thi * p.th, tni * p.tn);
			}
		}
	} else if (warp == 1) {
		{      // the whole warp, converged (see above)
			const uint64_t da_proto = p.wg_narrow ? make_smem_desc(0, 0, 512, 4) : make_smem_desc(0, Cfg::A_SLAB_BYTES, 1024, 2);
			const uint32_t a_kstep = (p.wg_narrow ? 16 * 64 : 16 * 128) >> 4;      // 16 pixels further down the slab
			const uint64_t db_proto = make_smem_desc(0, Cfg::B_BYTES, PC::SBO, PC::LAYOUT);
			const uint32_t idesc = p.idesc;
			for (int k = 0; k < n_steps; k++) {
				const int stage = k % Cfg::STAGES;
				const uint32_t phase = (uint32_t)(k / Cfg::STAGES) & 1u;
				mbar_wait(full_bar(stage), phase);
				tc_fence_after();
				const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
				const uint64_t da = da_proto + (sa >> 4), db = db_proto + ((sa + Cfg::A_BYTES) >> 4);
#pragma unroll
				for (int kk = 0; kk < Cfg::KPIX / 16; kk++)
					mma_f16_ss_warp(tmem_base, da + kk * a_kstep, db + ((kk * 16 * PC::ROW_BYTES) >> 4), idesc, (k | kk) != 0 ? 1u : 0u);
				mma_commit_warp(empty_bar(stage));
			}
			mma_commit_warp(done_bar);
		}
	} else if (warp < 6) {
		if (n_steps > 0) {
			const int quad = warp & 3;
			mbar_wait(done_bar, 0);
			tc_fence_after();
			const int f = quad * 32 + lane;
			const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
			for (int c0 = 0; c0 < KP; c0 += 32) {
				uint32_t r[32];
				if (KP - c0 >= 32) tmem_ld_32x32(t_row + c0, r);
				else { uint32_t hh[16]; tmem_ld_32x16(t_row + c0, hh);
#pragma unroll
					for (int j = 0; j < 16; j++) { r[j] = hh[j]; r[16 + j] = 0; } }
				tmem_ld_wait();
				if (f < p.n_real) {
					float* dst = p.grad + (size_t)f * KP + c0;
#pragma unroll
					for (int j = 0; j < 32; j += 4)
						if (c0 + j < KP)
							asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(r[j])),
							             "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])), "f"(__uint_as_float(r[j + 3])) : "memory");
				}
				__syncwarp();
			}
		}
	} else {
		// ===================== builders: warp pair q fills the patch tile of steps q, q+8, ... =====================
		const int bw = warp - 6;
		const int pair = bw >> 1;
		const int row = (bw & 1) * 32 + lane;
		const T* __restrict__ src = reinterpret_cast<const T*>(p.src);
		const size_t img_stride = (size_t)p.c * p.h * p.w + 1;
		const unsigned short bias_bits = bits_of<T>(p.bias_value);
		const int rx = row % p.tw, ry = (row / p.tw) % p.th, rn = row / (p.tw * p.th);
		const int tiles_w = p.tiles_w, tiles_h = p.tiles_h;
		int twi = (t_begin + pair) % tiles_w, thi = ((t_begin + pair) / tiles_w) % tiles_h, tni = (t_begin + pair) / (tiles_w * tiles_h);
		for (int k = pair; k < n_steps; k += WG_BUILD_WARPS / 2) {
			const int stage = k % Cfg::STAGES;
			const uint32_t phase = (uint32_t)(k / Cfg::STAGES) & 1u;
			const int px = twi * p.tw + rx, py = thi * p.th + ry, pn = tni * p.tn + rn;
			twi += WG_BUILD_WARPS / 2;
			while (twi >= tiles_w) { twi -= tiles_w; if (++thi == tiles_h) { thi = 0; tni++; } }
			const bool valid = px < p.W && py < p.H && pn < p.N;
			const int npx = twi * p.tw + rx, npy = thi * p.th + ry, npn = tni * p.tn + rn;
			// (measured: eight steps are in flight here, the prefetch buys nothing - 683 us with, 673 us without)
			const bool nvalid = p.prefetch > 1 && k + WG_BUILD_WARPS / 2 < n_steps && npx < p.W && npy < p.H && npn < p.N;
			mbar_wait(empty_bar(stage), phase ^ 1u);
			build_patch_row<T, C, FH, FW, KP>(src + (size_t)(valid ? pn : 0) * img_stride, valid, p.h, p.w,
				py * p.s_h - p.p_h, px * p.s_w - p.p_w, bias_bits, smem_base + stage * Cfg::STAGE_BYTES + Cfg::A_BYTES, row,
				nvalid ? src + (size_t)npn * img_stride : nullptr, npy * p.s_h - p.p_h, npx * p.s_w - p.p_w);
			fence_proxy_async();          // each writer publishes its row to the async proxy ...
			__syncwarp();
			if (lane == 0) mbar_arrive(full_bar(stage));     // ... then one arrival per warp
		}
	}
	tc_fence_before();
	__syncthreads();
	if (warp == 1) { __syncwarp(); tc_fence_after(); tmem_dealloc(tmem_base, Cfg::TMEM_COLS); }
}

// ================================================================ host side
static int kp_of(const cb200_conv_desc* d) { return cb200_patch_width(d->in_c, d->f_h, d->f_w); }

bool conv_first_supported(const cb200_conv_desc* d) {
	if (d->dtype != CB200_FP16 && d->dtype != CB200_BF16) return false;
	const bool shape = (d->in_c == 3 && d->f_h == 3 && d->f_w == 3) || (d->in_c == 1 && d->f_h == 3 && d->f_w == 3) ||
	                   (d->in_c == 1 && d->f_h == 5 && d->f_w == 5) || (d->in_c == 2 && d->f_h == 3 && d->f_w == 3);
	return shape && round8(d->out_c) <= 64 && d->out_c >= 8;
}

static void fill_params(const cb200_conv_desc* d, const void* src, int npix, FirstParams& p) {
	memset(&p, 0, sizeof(p));
	p.src = src;
	p.c = d->in_c; p.h = d->in_h; p.w = d->in_w; p.s_h = d->stride_h; p.s_w = d->stride_w; p.p_h = d->pad_h; p.p_w = d->pad_w;
	p.W = d->out_w; p.H = d->out_h; p.N = d->batch;
	choose_rect(p.W, p.H, p.N, npix, p.tw, p.th, p.tn);
	p.tiles_w = ceil_div(p.W, p.tw); p.tiles_h = ceil_div(p.H, p.th); p.tiles_n = ceil_div(p.N, p.tn);
	p.num_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
	p.n_real = d->out_c; p.n_pad = round8(d->out_c); p.length = d->length;
	p.bias_value = d->bias_value; p.activ = d->activ;
	static const bool no_prefetch = getenv("CB200_FIRST_PREFETCH") != nullptr && getenv("CB200_FIRST_PREFETCH")[0] == '0';
	p.prefetch = no_prefetch ? 0 : 1;
}

template <typename T, int C, int FH, int FW, int KP, int BN, bool GN = false>
static int launch_first_fwd(const CUtensorMap& mb, const CUtensorMap& mo, const FirstParams& p, int grid, cudaStream_t st) {
	using Cfg = FirstFwdCfg<KP, BN>;
	static bool configured = false;
	auto kern = conv_first_fwd_kernel<T, C, FH, FW, KP, BN, GN>;
	if (!configured) {
		if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) != cudaSuccess) {
			set_error("cudaFuncSetAttribute(smem=%d) failed", Cfg::SMEM_BYTES); return CB200_ERR_CUDA;
		}
		configured = true;
	}
	kern<<<grid, FWD_THREADS, Cfg::SMEM_BYTES, st>>>(mb, mo, p);
	CB_LAUNCH_CHECK();
	return CB200_OK;
}

template <typename T, int C, int FH, int FW, int KP>
static int launch_first_wgrad(const CUtensorMap& mdy, const FirstParams& p, int grid, cudaStream_t st) {
	using Cfg = FirstWgCfg<KP>;
	static bool configured = false;
	auto kern = conv_first_wgrad_kernel<T, C, FH, FW, KP>;
	if (!configured) {
		if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) != cudaSuccess) {
			set_error("cudaFuncSetAttribute(smem=%d) failed", Cfg::SMEM_BYTES); return CB200_ERR_CUDA;
		}
		configured = true;
	}
	kern<<<grid, WG_THREADS, Cfg::SMEM_BYTES, st>>>(mdy, p);
	CB_LAUNCH_CHECK();
	return CB200_OK;
}

#define FIRST_SHAPES(X)  X(3, 3, 3, 32) X(1, 3, 3, 16) X(1, 5, 5, 32) X(2, 3, 3, 32)

template <typename T>
static int first_fwd_typed(const cb200_conv_desc* d, const CUtensorMap& mb, const CUtensorMap& mo, const FirstParams& p, int grid, cudaStream_t st) {
	const int bn = p.n_pad > 32 ? 64 : 32;
#define X(C_, FH_, FW_, KP_) \
	if (d->in_c == C_ && d->f_h == FH_ && d->f_w == FW_) \
		return bn == 64 ? launch_first_fwd<T, C_, FH_, FW_, KP_, 64>(mb, mo, p, grid, st) \
		     : (p.gn_ws != nullptr ? launch_first_fwd<T, C_, FH_, FW_, KP_, 32, true>(mb, mo, p, grid, st) \
		                           : launch_first_fwd<T, C_, FH_, FW_, KP_, 32>(mb, mo, p, grid, st));
	FIRST_SHAPES(X)
#undef X
	set_error("conv_first: no kernel instance"); return CB200_ERR_UNSUPPORTED;
}
template <typename T>
static int first_wgrad_typed(const cb200_conv_desc* d, const CUtensorMap& mdy, const FirstParams& p, int grid, cudaStream_t st) {
#define X(C_, FH_, FW_, KP_) \
	if (d->in_c == C_ && d->f_h == FH_ && d->f_w == FW_) return launch_first_wgrad<T, C_, FH_, FW_, KP_>(mdy, p, grid, st);
	FIRST_SHAPES(X)
#undef X
	set_error("conv_first: no kernel instance"); return CB200_ERR_UNSUPPORTED;
}

// gn / gn_ws: the group-norm layer that follows and its FP64 workspace; *gn_fused = 1 when the epilogue has left the sums
// there (conv_tc.cu: conv_forward_tc has the same contract).  CB200_FIRST_GN_STATS=0 or cb200_set_gn_epilogue_stats(0): off.
extern int g_gn_epilogue_mode;
int conv_first_forward(const cb200_conv_desc* d, const cb200_conv_weights* w, const void* x_raw, void* y, cudaStream_t st,
                       const cb200_norm_desc* gn, void* gn_ws, int* gn_fused) {
	if (gn_fused) *gn_fused = 0;
	const int kp = kp_of(d);
	FirstParams p;
	fill_params(d, x_raw, 128, p);
	p.out = y;
	const int bn = p.n_pad > 32 ? 64 : 32;
	p.idesc = make_idesc_f16(d->dtype == CB200_BF16, 128, bn, 0, 0);
	CUtensorMap mb;
	int rc = make_w_map(&mb, w->w_fwd, d->dtype, kp, 1, d->out_c, kp, bn, swizzle_for(kp));
	if (rc) return rc;
	// output tile through shared memory + TMA store when a row of the tile is exactly one swizzle span (32 / 64 filters)
	static const bool no_tma_store = getenv("CB200_NO_TMA_STORE") != nullptr;
	CUtensorMap mo = mb;
	p.tma_store = (!no_tma_store && p.n_pad == bn) ? 1 : 0;
	if (p.tma_store) {
		rc = make_act_map(&mo, y, d->dtype, p.n_pad, p.W, p.H, p.N, bn, p.tw, p.th, p.tn, bn == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, 1);
		if (rc) return rc;
	}
	int grid = p.num_tiles < g_num_sms ? p.num_tiles : g_num_sms;
	p.tiles_per_cta = ceil_div(p.num_tiles, grid);
	grid = ceil_div(p.num_tiles, p.tiles_per_cta);
	static const bool no_gn = getenv("CB200_FIRST_GN_STATS") != nullptr && getenv("CB200_FIRST_GN_STATS")[0] == '0';
	if (g_gn_epilogue_mode < 0) { const char* e = getenv("CB200_GN_EPILOGUE_STATS"); g_gn_epilogue_mode = e != nullptr && e[0] != '\0' ? atoi(e) : 1; }
	if (gn_fused != nullptr && gn != nullptr && gn_ws != nullptr && !no_gn && g_gn_epilogue_mode != 0 && p.tma_store && bn == 32 && (p.tw * p.th) % 32 == 0 &&
	    gn->c == d->out_c && gn->batch == d->batch && gn->h == d->out_h && gn->w == d->out_w && gn->group_size >= 4 && gn->group_size % 4 == 0) {
		p.gn_ws = (double*)gn_ws; p.gn_groups = gn->nb_group;
		for (int i = 0; i < 8; i++) p.gn_map |= (uint32_t)((4 * i) / gn->group_size) << (4 * i);
		if (cudaMemsetAsync(gn_ws, 0, sizeof(double) * 2 * (size_t)d->batch * gn->nb_group, st) != cudaSuccess) { set_error("cudaMemsetAsync(group-norm sums) failed"); return CB200_ERR_CUDA; }
	}
	rc = d->dtype == CB200_FP16 ? first_fwd_typed<__half>(d, mb, mo, p, grid, st) : first_fwd_typed<__nv_bfloat16>(d, mb, mo, p, grid, st);
	if (rc == CB200_OK && p.gn_ws != nullptr) *gn_fused = 1;
	return rc;
}

int conv_first_wgrad(const cb200_conv_desc* d, const cb200_conv_weights* w, const void* x_raw, const void* dy, cudaStream_t st) {
	const int kp = kp_of(d);
	FirstParams p;
	fill_params(d, x_raw, 64, p);
	p.grad = w->grad;
	p.idesc = make_idesc_f16(d->dtype == CB200_BF16, 128, kp, 1, 1);
	CUtensorMap mdy;
	static const bool no_narrow = getenv("CB200_FIRST_WG_NARROW") != nullptr && getenv("CB200_FIRST_WG_NARROW")[0] == '0';
	p.wg_narrow = (p.n_pad <= 32 && !no_narrow) ? 1 : 0;
	int rc = p.wg_narrow ? make_act_map(&mdy, dy, d->dtype, p.n_pad, d->out_w, d->out_h, d->batch, 32, p.tw, p.th, p.tn, CU_TENSOR_MAP_SWIZZLE_64B, 1)
	                     : make_act_map(&mdy, dy, d->dtype, p.n_pad, d->out_w, d->out_h, d->batch, 64, p.tw, p.th, p.tn, CU_TENSOR_MAP_SWIZZLE_128B, 1);
	if (rc) return rc;
	int grid = p.num_tiles < g_num_sms ? p.num_tiles : g_num_sms;
	p.tiles_per_cta = ceil_div(p.num_tiles, grid);
	grid = ceil_div(p.num_tiles, p.tiles_per_cta);
	if (cudaMemsetAsync(w->grad, 0, sizeof(float) * (size_t)d->out_c * kp, st) != cudaSuccess) { set_error("wgrad memset failed"); return CB200_ERR_CUDA; }
	if (d->dtype == CB200_FP16) return first_wgrad_typed<__half>(d, mdy, p, grid, st);
	return first_wgrad_typed<__nv_bfloat16>(d, mdy, p, grid, st);
}

}  // namespace cb200
