// sm100_ptx.cuh - thin inline-PTX wrappers for the Blackwell (sm_100a) primitives the implicit-GEMM
// kernels are built from: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace cb200 { namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
	asm volatile(
		"{\n\t"
		".reg .pred P1;\n\t"
		"LAB_WAIT:\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
		"@P1 bra DONE;\n\t"
		"bra LAB_WAIT;\n\t"
		"DONE:\n\t"
		"}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* m) {
	asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
	             ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
	asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
	             ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
	asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
	             ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// The same for a producer warp that walks its loop converged (all 32 lanes, identical operands): elect.sync inside the asm
// picks the lane that issues (see mma_f16_ss_warp for why).
#define CB200_ELECT_ASM(body) "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q " body "\n\t}"
__device__ __forceinline__ void mbar_arrive_expect_tx_warp(uint32_t bar, uint32_t bytes) {
	asm volatile(CB200_ELECT_ASM("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;") ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d_warp(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
	asm volatile(CB200_ELECT_ASM("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];")
	             ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_4d_warp(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
	asm volatile(CB200_ELECT_ASM("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];")
	             ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// TMA store of a shared-memory box (written by generic-proxy stores + fence.proxy.async) to global memory; elements of
// the box outside the tensor are not written.  Bulk-group completion: commit, then wait_group(.read) before the buffer
// is reused (.read: the source has been read) or before the data must be visible (no .read).
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3) {
	asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
	             ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------- thread-block clusters
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
	asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// one TMA box written to the same CTA-relative shared-memory offset of every CTA in `cta_mask`, signalling the mbarrier
// at the same offset in each of them
__device__ __forceinline__ void tma_load_3d_multicast(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, uint16_t cta_mask) {
	asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;"
	             ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "h"(cta_mask) : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
// whole-warp instructions (.sync.aligned): call with all 32 lanes converged
__device__ __forceinline__ void tmem_alloc(uint32_t smem_slot, uint32_t ncols) {
	asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_slot), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
	asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], 16-bit inputs, FP32 accumulate; issued by ONE thread
__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
	asm volatile(
		"{\n\t"
		".reg .pred p;\n\t"
		"setp.ne.b32 p, %4, 0;\n\t"
		"tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
		"}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// The same two instructions for a warp that walks its issue loop CONVERGED (all 32 lanes, identical operands): one lane,
// chosen by elect.sync inside the asm, issues.  Called under `if (lane == 0)` the plain forms above make the compiler
// wrap every UTCHMMA in an elect / branch waterfall (its operands are per-thread registers that may differ between
// lanes); here the operands are provably uniform and the instruction is just predicated.
__device__ __forceinline__ void mma_f16_ss_warp(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
	asm volatile(
		"{\n\t"
		".reg .pred p, q;\n\t"
		"elect.sync _|q, 0xffffffff;\n\t"
		"setp.ne.b32 p, %4, 0;\n\t"
		"@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
		"}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit_warp(uint32_t bar) {
	asm volatile(
		"{\n\t"
		".reg .pred q;\n\t"
		"elect.sync _|q, 0xffffffff;\n\t"
		"@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
		"}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mma_commit_multicast_warp(uint32_t bar, uint16_t cta_mask) {
	asm volatile(
		"{\n\t"
		".reg .pred q;\n\t"
		"elect.sync _|q, 0xffffffff;\n\t"
		"@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t"
		"}" ::"r"(bar), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void mma_f16_ss_pair_warp(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
	asm volatile(
		"{\n\t"
		".reg .pred p, q;\n\t"
		"elect.sync _|q, 0xffffffff;\n\t"
		"setp.ne.b32 p, %4, 0;\n\t"
		"@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
		"}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit_pair_warp(uint32_t bar, uint16_t cta_mask) {
	asm volatile(
		"{\n\t"
		".reg .pred q;\n\t"
		"elect.sync _|q, 0xffffffff;\n\t"
		"@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t"
		"}" ::"r"(bar), "h"(cta_mask) : "memory");
}
// arrive on an mbarrier once all tcgen05.mma issued so far by this thread have completed
__device__ __forceinline__ void mma_commit(uint32_t bar) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// the same, arriving on the mbarrier at this CTA-relative offset in every CTA of `cta_mask`
__device__ __forceinline__ void mma_commit_multicast(uint32_t bar, uint16_t cta_mask) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(cta_mask) : "memory");
}
// 32 lanes x 32 consecutive columns of 32-bit: thread i of the warp receives lane (base+i), columns c..c+31
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
	asm volatile(
		"tcgen05.ld.sync.aligned.32x32b.x32.b32 "
		"{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
		"%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
		: "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
		  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
		  "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
		  "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
		: "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
	asm volatile(
		"tcgen05.ld.sync.aligned.32x32b.x16.b32 "
		"{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
		: "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
		  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
		: "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- CTA pairs (cta_group::2): one MMA of M = 256 over two SMs
// Checked on the hardware by scripts/exp/cta_pair_probe.cu: CTA r of the pair holds rows [128 r, 128 r + 128) of A and the
// r-th half of B's N rows at the SAME shared-memory offsets; one thread of the leader (rank 0) issues the MMAs with its
// local descriptors; CTA r's TMEM lanes hold rows 128 r + lane of D.
// shared::cluster address of `local` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t local, uint32_t rank) {
	uint32_t r;
	asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
	return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
	asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
// TMA loads into this CTA's shared memory whose bytes complete on an mbarrier of either CTA of the pair
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1, int c2) {
	asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
	             ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1, int c2, int c3) {
	asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
	             ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair_warp(uint32_t dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1, int c2) {
	asm volatile(CB200_ELECT_ASM("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];")
	             ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair_warp(uint32_t dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1, int c2, int c3) {
	asm volatile(CB200_ELECT_ASM("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];")
	             ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// one warp of EACH CTA of the pair
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_slot, uint32_t ncols) {
	asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_slot), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
	asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma_f16_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
	asm volatile(
		"{\n\t"
		".reg .pred p;\n\t"
		"setp.ne.b32 p, %4, 0;\n\t"
		"tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
		"}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on the mbarrier at this CTA-relative offset in every CTA of `cta_mask` once the pair's MMAs issued so far completed
__device__ __forceinline__ void mma_commit_pair(uint32_t bar, uint16_t cta_mask) {
	asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(cta_mask) : "memory");
}

// ---------------------------------------------------------------- descriptors
// shared-memory matrix descriptor (sm_100 format): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout type [61,64): 0 none, 2 = 128B swizzle, 4 = 64B, 6 = 32B.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
	uint64_t d = 0;
	d |= (uint64_t)((saddr >> 4) & 0x3fff);
	d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
	d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
	d |= (uint64_t)1 << 46;
	d |= (uint64_t)(layout_type & 7) << 61;
	return d;
}
// instruction descriptor for kind::f16: c_format F32 (bit 4), a/b format (0 = F16, 1 = BF16) at bits 7/10,
// a_major bit 15, b_major bit 16 (0 = K-major, 1 = MN-major), N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ inline uint32_t make_idesc_f16(int is_bf16, int m, int n, int a_mn_major, int b_mn_major) {
	uint32_t d = 0;
	d |= 1u << 4;
	d |= (uint32_t)(is_bf16 ? 1 : 0) << 7;
	d |= (uint32_t)(is_bf16 ? 1 : 0) << 10;
	d |= (uint32_t)(a_mn_major ? 1 : 0) << 15;
	d |= (uint32_t)(b_mn_major ? 1 : 0) << 16;
	d |= (uint32_t)(n >> 3) << 17;
	d |= (uint32_t)(m >> 4) << 24;
	return d;
}

// ---------------------------------------------------------------- packed FP32 pairs (FADD2 / FMUL2 / FFMA2)
// Two FP32 operations per issue slot on a 64-bit register pair (sm_100: add / mul / fma .f32x2, each half rounded like
// the scalar instruction).  The epilogues that are bound by issue slots use them on neighbouring accumulator columns
// (tcgen05.ld leaves those in a register pair) and for running sums.
__device__ __forceinline__ void add_f32x2(float& d0, float& d1, float a0, float a1) {            // d += a
	asm("{\n\t.reg .b64 a, d;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 d, {%0, %1};\n\tadd.rn.f32x2 d, a, d;\n\tmov.b64 {%0, %1}, d;\n\t}"
	    : "+f"(d0), "+f"(d1) : "f"(a0), "f"(a1));
}
__device__ __forceinline__ void sq_acc_f32x2(float& d0, float& d1, float a0, float a1) {         // d += a * a
	asm("{\n\t.reg .b64 a, d;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 d, {%0, %1};\n\tfma.rn.f32x2 d, a, a, d;\n\tmov.b64 {%0, %1}, d;\n\t}"
	    : "+f"(d0), "+f"(d1) : "f"(a0), "f"(a1));
}
__device__ __forceinline__ void fma_f32x2(float& d0, float& d1, float a0, float a1, float s, float c) {   // d = a * s + c
	asm("{\n\t.reg .b64 a, b, c, d;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %4};\n\tmov.b64 c, {%5, %5};\n\tfma.rn.f32x2 d, a, b, c;\n\tmov.b64 {%0, %1}, d;\n\t}"
	    : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(s), "f"(c));
}
__device__ __forceinline__ void mul_f32x2(float& d0, float& d1, float a0, float a1, float s) {            // d = a * s
	asm("{\n\t.reg .b64 a, b, d;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %4};\n\tmul.rn.f32x2 d, a, b;\n\tmov.b64 {%0, %1}, d;\n\t}"
	    : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(s));
}
// leaky ReLU with saturation on two neighbouring values: z <= 0 ? z * leak : (z > sat ? sat + (z - sat) * leak : z) as
// min(max(z, z * leak), z * leak + sat_c), sat_c = sat - sat * leak, for 0 <= leak <= 1, sat >= 0: 3 issue slots per value
// instead of 5.  Below the saturation the result is z or z * leak exactly as in the scalar form; above it the
// rearranged branch can differ in the last FP32 bit (invisible after the rounding to 16 bit).
__device__ __forceinline__ void leaky_sat_f32x2(float& z0, float& z1, float leak, float sat_c) {
	float t0, t1, h0, h1;
	mul_f32x2(t0, t1, z0, z1, leak);
	fma_f32x2(h0, h1, z0, z1, leak, sat_c);
	z0 = fminf(fmaxf(z0, t0), h0);
	z1 = fminf(fmaxf(z1, t1), h1);
}

}}  // namespace cb200::ptx
