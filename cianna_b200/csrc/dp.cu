// dp.cu - data-parallel gradient exchange: one process per GPU, NCCL all-reduce over NVLink/NVSwitch
// on a dedicated communication stream so that the exchange of layer l overlaps the backward pass of
// layers l-1 ... 0.  The reference has no multi-GPU path at all (src/cuda/cuda_main.cu:1066 is a
// placeholder for device selection); this is new functionality behind the same training loop.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2) so that the library loads on machines
// where only single-GPU use is wanted, and so that a process that already mapped an NCCL
// (e.g. through torch.distributed) shares it instead of loading a second copy.
#include <dlfcn.h>
#include <nccl.h>
#include "common.cuh"

namespace cb200 {
struct NcclApi {
	void* handle = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static ncclComm_t g_comm = nullptr;
static int g_world = 1, g_rank = 0;
static cudaStream_t g_comm_stream = nullptr;
static cudaEvent_t g_ev_ready = nullptr, g_ev_done = nullptr;

static int load_nccl() {
	if (g_nccl.handle) return CB200_OK;
	const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
	for (int i = 0; names[i] && !g_nccl.handle; i++) g_nccl.handle = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
	if (!g_nccl.handle) { set_error("cannot dlopen libnccl.so.2: %s", dlerror()); return CB200_ERR_NCCL; }
	g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(g_nccl.handle, "ncclGetUniqueId");
	g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(g_nccl.handle, "ncclCommInitRank");
	g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(g_nccl.handle, "ncclAllReduce");
	g_nccl.Broadcast = (decltype(g_nccl.Broadcast))dlsym(g_nccl.handle, "ncclBroadcast");
	g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(g_nccl.handle, "ncclCommDestroy");
	g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(g_nccl.handle, "ncclGetErrorString");
	if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.Broadcast || !g_nccl.CommDestroy) {
		set_error("libnccl is missing required symbols"); return CB200_ERR_NCCL;
	}
	return CB200_OK;
}
#define CB_NCCL(call)                                                                          \
	do { ncclResult_t r__ = (call); if (r__ != ncclSuccess) {                                   \
	     set_error("%s -> %s", #call, g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "nccl error"); \
	     return CB200_ERR_NCCL; } } while (0)
}  // namespace cb200
using namespace cb200;

extern "C" {

int cb200_dp_unique_id(void* id128) {
	int rc = load_nccl(); if (rc) return rc;
	static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
	ncclUniqueId id;
	CB_NCCL(g_nccl.GetUniqueId(&id));
	memcpy(id128, &id, sizeof(id));
	return CB200_OK;
}

int cb200_dp_init(const void* id128, int rank, int world) {
	CB_REQUIRE_DEVICE();
	CB_ARG(world >= 1 && rank >= 0 && rank < world);
	g_world = world; g_rank = rank;
	if (world == 1) return CB200_OK;
	int rc = load_nccl(); if (rc) return rc;
	ncclUniqueId id;
	memcpy(&id, id128, sizeof(id));
	CB_NCCL(g_nccl.CommInitRank(&g_comm, world, id, rank));
	{
		// the exchange is short and everything after the backward sweep waits for it: its CTAs go first whenever an SM frees
		// up (the weight-gradient stream has the lowest priority, the compute stream the highest - the same as this one)
		int lo = 0, hi = 0;
		CB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
		CB_CUDA(cudaStreamCreateWithPriority(&g_comm_stream, cudaStreamNonBlocking, hi));
	}
	CB_CUDA(cudaEventCreateWithFlags(&g_ev_ready, cudaEventDisableTiming));
	CB_CUDA(cudaEventCreateWithFlags(&g_ev_done, cudaEventDisableTiming));
	return CB200_OK;
}

int cb200_dp_world(void) { return g_world; }

int cb200_dp_allreduce(float* buf, size_t n, void* after_stream) {
	CB_REQUIRE_DEVICE();
	if (g_world == 1 || n == 0) return CB200_OK;
	CB_ARG(g_comm != nullptr);
	// the communication stream picks up once the producer stream has reached this point
	CB_CUDA(cudaEventRecord(g_ev_ready, as_stream(after_stream)));
	CB_CUDA(cudaStreamWaitEvent(g_comm_stream, g_ev_ready, 0));
	CB_NCCL(g_nccl.AllReduce(buf, buf, n, ncclFloat, ncclSum, g_comm, g_comm_stream));
	return CB200_OK;
}

int cb200_dp_after(void* stream) {
	CB_REQUIRE_DEVICE();
	if (g_world == 1) return CB200_OK;
	CB_ARG(g_comm != nullptr);
	CB_CUDA(cudaEventRecord(g_ev_ready, as_stream(stream)));
	CB_CUDA(cudaStreamWaitEvent(g_comm_stream, g_ev_ready, 0));
	return CB200_OK;
}

int cb200_dp_broadcast(void* buf, size_t bytes, int root, void* stream) {
	CB_REQUIRE_DEVICE();
	if (g_world == 1 || bytes == 0) return CB200_OK;
	CB_ARG(g_comm != nullptr && root >= 0 && root < g_world);
	CB_NCCL(g_nccl.Broadcast(buf, buf, bytes, ncclChar, root, g_comm, as_stream(stream)));
	return CB200_OK;
}

int cb200_dp_rank(void) { return g_rank; }

int cb200_dp_join(void* stream) {
	CB_REQUIRE_DEVICE();
	if (g_world == 1) return CB200_OK;
	CB_CUDA(cudaEventRecord(g_ev_done, g_comm_stream));
	CB_CUDA(cudaStreamWaitEvent(as_stream(stream), g_ev_done, 0));
	return CB200_OK;
}

int cb200_dp_finalize(void) {
	if (g_comm) { g_nccl.CommDestroy(g_comm); g_comm = nullptr; }
	g_world = 1; g_rank = 0;
	return CB200_OK;
}

}  // extern "C"
