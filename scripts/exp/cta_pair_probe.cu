// cta_pair_probe.cu - hardware probe (not part of the product): one 256 x 256 x K GEMM tile on a CTA PAIR
// (tcgen05.mma.cta_group::2, M = 256 across two SMs), the building block the wide-N convolution kernels need to get
// under the shared-memory bandwidth ceiling (with M = 128 per SM every K = 16 step reads 12 KB of operands
// while TMA refills 12 KB; with a pair each SM holds its own 128 rows of A and HALF of B: 8 KB + 8 KB per step).
// What it pins down on the hardware, against a CPU product of small integers (exact in FP16 / FP32):
//   - both CTAs' TMA loads (cp.async.bulk.tensor ... .cta_group::2) completing on the LEADER's mbarrier (mapa address);
//   - operand placement: CTA r holds rows [128 r, 128 r + 128) of A and rows [128 r, 128 r + 128) of B (N half), at the
//     same shared-memory offsets in both CTAs; one thread of the leader issues the MMAs with its local descriptors;
//   - tcgen05.commit.cta_group::2 ... multicast::cluster arriving on the barrier at the same offset in both CTAs;
//   - accumulator placement: CTA r's TMEM lanes 0..127 = rows 128 r + lane of D, columns = N;
//   - tcgen05.alloc.cta_group::2 issued by one warp of EACH CTA.
// mode 0: K-major A and B (forward / data gradient); mode 1: MN-major A and B (weight gradient: contraction over pixels).
// Every wait is bounded (a timed-out wait reports its code instead of hanging the GPU).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o cta_pair_probe cta_pair_probe.cu -lcuda
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../cianna_b200/csrc/sm100_ptx.cuh"
using namespace cb200::ptx;

constexpr int M = 256, N = 256, KB = 64;          // tile; K block = 64 halves = one 128-byte swizzle row
constexpr int A_BYTES = 128 * KB * 2, B_BYTES = 128 * KB * 2;

__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity) {
	for (int i = 0; i < 4000000; i++) {
		uint32_t ok;
		asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
		             : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
		if (ok) return true;
	}
	return false;
}
// TMA load whose completion bytes go to an mbarrier that may live in the peer CTA of the pair
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1) {
	asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
	             ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
// mode 0: A[M][K], B[N][K] (K contiguous); mode 1: A[K][M], B[K][N] (M / N contiguous)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, float* out, int* status, int kblocks, int mode) {
	extern __shared__ uint8_t smem_raw[];
	const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
	const uint32_t rank = cluster_ctarank();
	// stage s (one per K block, nothing is reused): A at base + s * 32 KB, B half behind it
	auto a_smem = [&](int s) { return base + (uint32_t)s * (A_BYTES + B_BYTES); };
	auto b_smem = [&](int s) { return a_smem(s) + A_BYTES; };
	const uint32_t bar_base = base + 4 * (A_BYTES + B_BYTES);
	auto full_bar = [&](int s) { return bar_base + 8u * s; };
	const uint32_t done_bar = bar_base + 64, slot = bar_base + 128;
	uint32_t* slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (threadIdx.x == 0) {
		for (int s = 0; s < 4; s++) mbar_init(full_bar(s), 1);
		mbar_init(done_bar, 1);
		fence_barrier_init();
	}
	if (warp == 1) { tmem_alloc_pair(slot, 256); tmem_relinquish_pair(); }
	tc_fence_before();
	__syncthreads();
	cluster_sync();                                    // both CTAs' barriers exist before anyone signals them
	tc_fence_after();
	const uint32_t tmem = *slot_ptr;
	if (threadIdx.x == 0) status[8 + rank] = (int)tmem;

	if (warp == 0 && lane == 0) {
		// producer of this CTA: its 128 rows of A and its 128 rows of B per K block, all completing on the leader's barrier
		for (int s = 0; s < kblocks; s++) {
			if (rank == 0) mbar_arrive_expect_tx(full_bar(s), 2 * (A_BYTES + B_BYTES));
			const uint32_t bar = mapa_rank(full_bar(s), 0);
			if (mode == 0) {
				tma_load_2d_pair(a_smem(s), &tmap_a, bar, s * KB, (int)rank * 128);
				tma_load_2d_pair(b_smem(s), &tmap_b, bar, s * KB, (int)rank * 128);
			} else {
				// MN-major: box = 64 M (or N) elements x 64 K rows; two boxes per operand half
				for (int h = 0; h < 2; h++) {
					tma_load_2d_pair(a_smem(s) + h * 8192, &tmap_a, bar, (int)rank * 128 + h * 64, s * KB);
					tma_load_2d_pair(b_smem(s) + h * 8192, &tmap_b, bar, (int)rank * 128 + h * 64, s * KB);
				}
			}
		}
	} else if (warp == 1 && lane == 0 && rank == 0) {
		// MMA issuer: the leader only
		const uint32_t idesc = make_idesc_f16(0, M, N, mode, mode);
		bool ok = true;
		for (int s = 0; s < kblocks && ok; s++) {
			ok = mbar_wait_bounded(full_bar(s), 0);
			if (!ok) { status[0] = 100 + s; break; }
			tc_fence_after();
			for (int kk = 0; kk < KB / 16; kk++) {
				uint64_t da, db;
				if (mode == 0) {
					da = make_smem_desc(a_smem(s) + kk * 32, 16, 1024, 2);
					db = make_smem_desc(b_smem(s) + kk * 32, 16, 1024, 2);
				} else {
					// MN-major, 128B swizzle: 64-element slabs 8192 B apart (LBO), 8 K rows = 1024 B (SBO); K step of 16 rows = 2048 B
					da = make_smem_desc(a_smem(s) + kk * 2048, 8192, 1024, 2);
					db = make_smem_desc(b_smem(s) + kk * 2048, 8192, 1024, 2);
				}
				mma_f16_ss_pair(tmem, da, db, idesc, (s | kk) != 0 ? 1u : 0u);
			}
		}
		mma_commit_pair(done_bar, 0x3);
	}
	if (warp >= 2) {
		// epilogue: 4 warps, each its TMEM lane quadrant; rows of D = 128 * rank + lane index
		const bool ok = mbar_wait_bounded(done_bar, 0);
		if (!ok && lane == 0) status[1 + rank] = 200 + warp;
		tc_fence_after();
		const int quad = warp & 3;
		if (ok) {
			for (int c0 = 0; c0 < N; c0 += 32) {
				uint32_t r[32];
				tmem_ld_32x32(tmem + ((uint32_t)(quad * 32) << 16) + c0, r);
				tmem_ld_wait();
				for (int j = 0; j < 32; j++) out[(size_t)(rank * 128 + quad * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
			}
		}
		tc_fence_before();
	}
	__syncthreads();
	cluster_sync();
	if (warp == 1) { tc_fence_after(); tmem_dealloc_pair(tmem, 256); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map(EncodeTiledFn enc, CUtensorMap* m, void* base, int inner, int outer, int box_inner, int box_outer) {
	cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
	cuuint64_t strides[1] = {(cuuint64_t)inner * 2};
	cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer}, estr[2] = {1, 1};
	CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
	                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
	return 0;
}

static int run(EncodeTiledFn enc, int mode, int kblocks) {
	const int K = kblocks * KB;
	__half* ha = (__half*)malloc(sizeof(__half) * M * K);
	__half* hb = (__half*)malloc(sizeof(__half) * N * K);
	float* ref = (float*)malloc(sizeof(float) * M * N);
	// logical A[m][k], B[n][k]; small integers so every product and sum is exact
	auto av = [&](int m, int k) { return (float)((m * 5 + k * 3) % 7 - 3); };
	auto bv = [&](int n, int k) { return (float)((n * 3 + k * 7) % 5 - 2); };
	for (int m = 0; m < M; m++) for (int k = 0; k < K; k++) ha[mode == 0 ? m * K + k : k * M + m] = __float2half(av(m, k));
	for (int n = 0; n < N; n++) for (int k = 0; k < K; k++) hb[mode == 0 ? n * K + k : k * N + n] = __float2half(bv(n, k));
	for (int m = 0; m < M; m++) for (int n = 0; n < N; n++) {
		float s = 0.0f;
		for (int k = 0; k < K; k++) s += av(m, k) * bv(n, k);
		ref[m * N + n] = s;
	}
	__half *da, *db; float* dout; int* dstatus;
	cudaMalloc(&da, sizeof(__half) * M * K); cudaMalloc(&db, sizeof(__half) * N * K);
	cudaMalloc(&dout, sizeof(float) * M * N); cudaMalloc(&dstatus, sizeof(int) * 16);
	cudaMemcpy(da, ha, sizeof(__half) * M * K, cudaMemcpyHostToDevice);
	cudaMemcpy(db, hb, sizeof(__half) * N * K, cudaMemcpyHostToDevice);
	cudaMemset(dout, 0xff, sizeof(float) * M * N);
	cudaMemset(dstatus, 0, sizeof(int) * 16);
	CUtensorMap ma, mb;
	if (mode == 0) {
		if (make_map(enc, &ma, da, K, M, KB, 128) || make_map(enc, &mb, db, K, N, KB, 128)) return 1;
	} else {
		if (make_map(enc, &ma, da, M, K, 64, KB) || make_map(enc, &mb, db, N, K, 64, KB)) return 1;
	}
	const int smem = 4 * (A_BYTES + B_BYTES) + 1024 + 256;
	cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	pair_kernel<<<2, 192, smem>>>(ma, mb, dout, dstatus, kblocks, mode);
	cudaError_t e = cudaDeviceSynchronize();
	if (e != cudaSuccess) { printf("mode %d: kernel error %s\n", mode, cudaGetErrorString(e)); return 1; }
	float* hout = (float*)malloc(sizeof(float) * M * N);
	int hs[16];
	cudaMemcpy(hout, dout, sizeof(float) * M * N, cudaMemcpyDeviceToHost);
	cudaMemcpy(hs, dstatus, sizeof(hs), cudaMemcpyDeviceToHost);
	int bad = 0, bad_q[4] = {0, 0, 0, 0};
	for (int m = 0; m < M; m++) for (int n = 0; n < N; n++)
		if (hout[m * N + n] != ref[m * N + n]) { bad++; bad_q[(m / 128) * 2 + n / 128]++; }
	printf("mode %d (%s) K=%d: %s  wrong %d / %d  [per quadrant m<128,n<128: %d | m<128,n>=128: %d | m>=128,n<128: %d | m>=128,n>=128: %d]  "
	       "status mma=%d epi0=%d epi1=%d tmem=%x/%x\n", mode, mode == 0 ? "K-major" : "MN-major", K, bad ? "MISMATCH" : "exact", bad, M * N,
	       bad_q[0], bad_q[1], bad_q[2], bad_q[3], hs[0], hs[1], hs[2], hs[8], hs[9]);
	if (bad) {
		int shown = 0;
		for (int m = 0; m < M && shown < 6; m += 37) for (int n = 0; n < N && shown < 6; n += 53)
			if (hout[m * N + n] != ref[m * N + n]) { printf("   D[%d][%d] = %g, want %g\n", m, n, hout[m * N + n], ref[m * N + n]); shown++; }
	}
	return bad != 0;
}

int main() {
	cudaSetDevice(0);
	cudaFree(0);
	void* fn = nullptr;
	cudaDriverEntryPointQueryResult q;
	if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { printf("no encode fn\n"); return 1; }
	int rc = 0;
	rc |= run((EncodeTiledFn)fn, 0, 1);
	rc |= run((EncodeTiledFn)fn, 0, 4);
	rc |= run((EncodeTiledFn)fn, 1, 1);
	rc |= run((EncodeTiledFn)fn, 1, 4);
	return rc;
}
