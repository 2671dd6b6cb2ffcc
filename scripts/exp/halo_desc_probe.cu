// halo_desc_probe.cu - hardware probe (not part of the product): can a tcgen05 K-major shared-memory descriptor start at
// a row that is NOT aligned to the swizzle atom, with a stride between 8-row groups that is not a multiple of the atom?
// That is what reusing one TMA-loaded halo tile [18 rows][10 px][C] for the nine taps of a 3x3 convolution needs
// (M tile = 16 image rows x 8 px: group g = image row g, start = (ky*10 + kx) rows into the tile, SBO = 10 rows).
// For every tap it runs D = A_shifted * I and compares D with the expected shifted pixels, once with base_offset = 0 and
// once with base_offset = (start >> 7) & 7.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o halo_desc_probe halo_desc_probe.cu -lcuda
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../cianna_b200/csrc/sm100_ptx.cuh"
using namespace cb200::ptx;

constexpr int TW = 8, TH = 16, HW = TW + 2, HH = TH + 2;

__device__ __forceinline__ uint64_t desc_bo(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t base_off) {
	return make_smem_desc(saddr, lbo, sbo, layout) | ((uint64_t)(base_off & 7) << 49);
}

template <int BK>
__global__ void __launch_bounds__(192, 1) probe_kernel(const __grid_constant__ CUtensorMap tmap, float* out, int mode) {
	constexpr int ROWB = BK * 2;
	constexpr uint32_t LAYOUT = BK == 64 ? 2u : 4u;
	extern __shared__ uint8_t smem_raw[];
	const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
	const uint32_t a_smem = base;                           // halo tile: HH*HW rows of ROWB bytes
	const uint32_t b_smem = base + 32768;                   // identity [BK][BK], K-major, same swizzle
	const uint32_t bar = base + 49152, bar2 = bar + 8, slot = bar + 16;
	uint32_t* slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); fence_barrier_init(); }
	if (warp == 1) { tmem_alloc(slot, 64); tmem_relinquish(); }
	// identity B: row n, element k = (n == k); 16B chunk j holds k = 8j..8j+7, chunk index XOR-swizzled with the row
	for (int i = threadIdx.x; i < BK * (BK / 8); i += blockDim.x) {
		const int n = i / (BK / 8), j = i % (BK / 8);
		uint32_t w[4] = {0, 0, 0, 0};
		if (n / 8 == j) { const int e = n % 8; w[e >> 1] = 0x3C00u << (16 * (e & 1)); }   // half 1.0
		const int x = BK == 64 ? (n & 7) : ((n >> 1) & 3);
		const uint32_t dst = b_smem + n * ROWB + ((j ^ x) << 4);
		asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
	}
	fence_proxy_async();
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem = *slot_ptr;
	if (threadIdx.x == 0) {
		mbar_arrive_expect_tx(bar, HH * HW * ROWB);
		tma_load_3d(a_smem, &tmap, bar, 0, 0, 0);
	}
	const uint32_t idesc = make_idesc_f16(0, 128, BK, 0, 0);
	for (int tap = 0; tap < 9; tap++) {
		if (warp == 1 && lane == 0) {
			if (tap == 0) mbar_wait(bar, 0);
			tc_fence_after();
			const int ky = tap / 3, kx = tap % 3;
			const uint32_t start = a_smem + (ky * HW + kx) * ROWB;
			const uint32_t bo = mode == 1 ? ((start >> 7) & 7) : 0;
			for (int kk = 0; kk < BK / 16; kk++) {
				const uint64_t da = desc_bo(start + kk * 32, 16, HW * ROWB, LAYOUT, bo);
				const uint64_t db = make_smem_desc(b_smem + kk * 32, 16, 8 * ROWB, LAYOUT);
				mma_f16_ss(tmem, da, db, idesc, kk != 0 ? 1u : 0u);
			}
			mma_commit(bar2);
		}
		if (warp >= 2) {
			mbar_wait(bar2, tap & 1);
			tc_fence_after();
			const int quad = warp & 3;
			for (int c0 = 0; c0 < BK; c0 += 32) {
				uint32_t r[32];
				tmem_ld_32x32(tmem + ((uint32_t)(quad * 32) << 16) + c0, r);
				tmem_ld_wait();
				for (int j = 0; j < 32; j++) out[(tap * 128 + quad * 32 + lane) * BK + c0 + j] = __uint_as_float(r[j]);
			}
			tc_fence_before();
		}
		__syncthreads();
	}
	tc_fence_before();
	__syncthreads();
	if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int BK> static int run(EncodeTiledFn enc) {
	const int H = 24, W = 16;
	__half* hx = (__half*)malloc(sizeof(__half) * H * W * BK);
	for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) for (int c = 0; c < BK; c++)
		hx[(y * W + x) * BK + c] = __float2half((float)(((y * W + x) * 7 + c * 3) % 17 - 8));
	__half* dx; float* dout;
	cudaMalloc(&dx, sizeof(__half) * H * W * BK);
	cudaMemcpy(dx, hx, sizeof(__half) * H * W * BK, cudaMemcpyHostToDevice);
	cudaMalloc(&dout, sizeof(float) * 9 * 128 * BK);
	CUtensorMap m;
	cuuint64_t dims[3] = {(cuuint64_t)BK, (cuuint64_t)W, (cuuint64_t)H};
	cuuint64_t strides[2] = {(cuuint64_t)BK * 2, (cuuint64_t)W * BK * 2};
	cuuint32_t box[3] = {(cuuint32_t)BK, HW, HH}, estr[3] = {1, 1, 1};
	CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, dx, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
	                 BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
	cudaFuncSetAttribute(probe_kernel<BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
	float* hout = (float*)malloc(sizeof(float) * 9 * 128 * BK);
	for (int mode = 0; mode < 2; mode++) {
		cudaMemset(dout, 0xff, sizeof(float) * 9 * 128 * BK);
		probe_kernel<BK><<<1, 192, 64 * 1024>>>(m, dout, mode);
		cudaError_t e = cudaDeviceSynchronize();
		if (e != cudaSuccess) { printf("BK=%d mode=%d: kernel error %s\n", BK, mode, cudaGetErrorString(e)); return 1; }
		cudaMemcpy(hout, dout, sizeof(float) * 9 * 128 * BK, cudaMemcpyDeviceToHost);
		for (int tap = 0; tap < 9; tap++) {
			int bad = 0;
			for (int mrow = 0; mrow < 128; mrow++) for (int c = 0; c < BK; c++) {
				const int y = mrow / 8 + tap / 3, x = mrow % 8 + tap % 3;
				const float want = (float)(((y * W + x) * 7 + c * 3) % 17 - 8);
				if (hout[(tap * 128 + mrow) * BK + c] != want) bad++;
			}
			printf("BK=%d base_offset_mode=%d tap(%d,%d): %s (%d / %d wrong)\n", BK, mode, tap / 3, tap % 3, bad ? "MISMATCH" : "exact", bad, 128 * BK);
		}
	}
	return 0;
}

int main() {
	cudaSetDevice(0);
	cudaFree(0);
	void* fn = nullptr;
	cudaDriverEntryPointQueryResult q;
	if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { printf("no encode fn\n"); return 1; }
	int rc = run<64>((EncodeTiledFn)fn);
	rc |= run<32>((EncodeTiledFn)fn);
	return rc;
}
