// rtb_api.cu — implementation of the C ABI declared in include/rtb.h: context, scene upload,
// launch wrappers.  Device code: rtb_kernels.cuh.  Build: see raytracingrenderer_b200/build.py
// (nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo).
#include "rtb_accel.hpp"
#include "rtb_cwbvh.hpp"
#include "rtb_wavefront.cuh"
#include "rtb_gpu_build.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <dlfcn.h>
#include <chrono>
#include <map>
#include <new>
#include <string>
#include <thread>
#include <vector>

// run __VA_ARGS__ with `TR` bound to the compile-time value of a run-time rtb_traversal
#define RTB_TRAV_SWITCH(trav, ...)                         \
	switch (trav)                                          \
	{                                                      \
	case RTB_TRAV_EXACT:                                   \
	{                                                      \
		constexpr int TR = RTB_TRAV_EXACT;                 \
		__VA_ARGS__;                                       \
	}                                                      \
	break;                                                 \
	case RTB_TRAV_FAST:                                    \
	{                                                      \
		constexpr int TR = RTB_TRAV_FAST;                  \
		__VA_ARGS__;                                       \
	}                                                      \
	break;                                                 \
	case RTB_TRAV_CW:                                      \
	{                                                      \
		constexpr int TR = RTB_TRAV_CW;                    \
		__VA_ARGS__;                                       \
	}                                                      \
	break;                                                 \
	case RTB_TRAV_Q16:                                     \
	{                                                      \
		constexpr int TR = RTB_TRAV_Q16;                   \
		__VA_ARGS__;                                       \
	}                                                      \
	break;                                                 \
	default:                                               \
	{                                                      \
		constexpr int TR = RTB_TRAV_WIDE;                  \
		__VA_ARGS__;                                       \
	}                                                      \
	break;                                                 \
	}

namespace
{
thread_local std::string g_createError;

struct EventPair
{
	cudaEvent_t a, b;
};
struct StageEvents
{
	cudaEvent_t e[4]; // before extend, after extend, after shade, after shadow
};
} // namespace

#define RTB_MAX_POOLS 8

// One wavefront render in flight (renderWavefront): everything needed to enqueue further iterations of it later.
struct WfRun
{
	bool active = false; // iterations are enqueued but nobody has checked yet that the pool drained (rtb_render returned early)
	WfArgs A[RTB_MAX_POOLS];
	rtb_params P;
	int K = 0, ti = 0;
	uint32_t perPool = 0, nSlotsAlloc = 0, bound = 0, it = 0, vertices = 0, batch = 0;
	unsigned gridExtend = 0, gridSlots = 0, gridSort = 0;
	bool shadows = false, sortSh = false, sortEx = false, cwKernels = false, cwShadow = false, drained = false;
	unsigned long long totalJobs = 0, key = 0;
};

struct rtb_ctx
{
	int device = 0;
	cudaStream_t stream = nullptr;
	std::string error;
	rtb_params params;
	bool haveScene = false;
	DevScene S;
	std::vector<void*> sceneAllocs;
	float* film = nullptr;
	float* filmFiltered = nullptr;
	uint8_t* tone = nullptr;
	unsigned long long* counters = nullptr; // 8 x u64
	uint32_t width = 0, height = 0;
	uint32_t spp = 0;
	uint64_t launches = 0;
	double renderMs = 0.0;
	std::vector<EventPair> pending;
	std::vector<StageEvents> pendingStages;
	double stageMs[3] = {0, 0, 0};
	uint64_t timedIterations = 0;
	std::vector<cudaEvent_t> eventPool;
	uint32_t fastDepth = 0;
	long long* accum = nullptr; // fixed-point film sums (master copy; `film` is derived)
	bool filmDirty = false;     // accum changed since the last resolve
	// wavefront pool (allocated on first use)
	void* wfState = nullptr;
	size_t wfStateBytes = 0;
	WfCtrl* wfCtrl = nullptr;
	uint32_t wfCtrlEntries = 0;
	WfGlobal* wfGlobal = nullptr;
	uint32_t* wfTiles = nullptr;
	long long* accumScratch = nullptr; // rtb_render_adaptive: per-phase sums
	uint32_t* wfTilesAdaptive = nullptr; // 32 sub-tiles (8x4 pixels) per 32x32 tile, 0xFFFFFFFF outside the image
	unsigned long long* adaptJobBase = nullptr;
	uint32_t* adaptSamples = nullptr;
	float* adaptVariance = nullptr;
	float4* wfPrimary = nullptr; // per-pixel primary hits (params.primary_reuse), rebuilt by every render call
	int primaryPasses = 2; // profiles/r01_primary_reuse.txt
	uint32_t wfTileCount = 0;
	int wfTilePart[3] = {-1, -1, -1}; // partition, rank, world the tile list was built for
	unsigned long long* hostProbe = nullptr; // pinned: {nextJob, alive}
	int smCount = 0;
	int travBlocksPerSM[5] = {0, 0, 0, 0, 0}; // persistent extend kernel, per rtb_traversal
	// RTB_TRAV_CW: shared-memory staging of the top of the tree (per block), persistent kernels' resident blocks
	uint32_t cwStageNodes = 0, cwStageLeaves = 0, cwSmemBytes = 0;
	int cwBlocksPerSM[2] = {0, 0}; // closest hit, any hit
	int cwStageKB = 32;            // RTB_CW_STAGE_KB
	int shadowPersistent = -1;     // RTB_SHADOW_PERSISTENT: the persistent any-hit kernel (1), one thread per queued ray (0), per scene (-1)
	bool shadowPersistentAuto = false;
	uint32_t chunkOverride = 0; // RTB_CHUNK
	int cwShadowPersistent = -1;   // RTB_CW_SHADOW: 1 = persistent any-hit kernel, 0 = one thread per queued ray, -1 = per scene
	bool haveWide = false, haveCw = false, haveQ16 = false; // re-encodings of the FAST tree, built on first use
	uint32_t poolSlots = 8u << 20; // profiles/r01_pool_sweep.txt: per-launch ramp/tail amortise up to ~8 M slots
	bool simpleExtend = false;
	int pools = 2; // sub-pools advancing concurrently on their own streams (profiles/r01_pool_sweep.txt)
	cudaStream_t poolStreams[RTB_MAX_POOLS] = {};
	cudaEvent_t poolDone[RTB_MAX_POOLS] = {};
	// the shadow stage of iteration i runs beside the extend stage of iteration i+1 of the same sub-pool
	cudaStream_t shadowStreams[RTB_MAX_POOLS] = {};
	cudaEvent_t evShaded[RTB_MAX_POOLS] = {}, evShadowed[RTB_MAX_POOLS] = {};
	bool shadowAsync = true;
	// ray binning between the stages (k_sort_*): -1 = decide per scene (rtb_upload_scene), 0 / 1 = forced (RTB_SORT_SHADOW / RTB_SORT_EXTEND)
	int sortShadow = -1, sortExtend = -1;
	bool sortShadowAuto = false, sortExtendAuto = false;
	uint32_t* wfPerm = nullptr; // [shadow queue entries + slots]
	size_t wfPermEntries = 0;
	uint32_t* wfSortWork = nullptr; // per sub-pool: hist[2][4096], cursor[2][4096]
	unsigned lightGrid = 0;
	void* vpls = nullptr; // rtb_render_ir: n_paths segments of RTB_VPL_SEGMENT VPLs
	uint32_t* vplCounts = nullptr;
	uint32_t vplPaths = 0;
	cudaEvent_t evFork = nullptr;
	uint64_t wfIterations = 0, wfHostSyncs = 0, wfAsyncRenders = 0;
	// rtb_render is asynchronous once a configuration has been rendered before: the iteration count it needed is remembered
	// (iterHint) and the next identical call enqueues that many (+6 % + 3) at once and returns; whoever touches the context next
	// (wfSettle) checks that the pool drained and enqueues the rest if it did not.
	WfRun run;
	std::map<unsigned long long, uint32_t> iterHint;
	cudaEvent_t evProbe = nullptr;
	int asyncRender = 1; // RTB_ASYNC_RENDER
	// ---- device group (rtb_create_multi): this context is device 0 of the group and owns the film that is read
	// out; `peers` are complete single-device contexts on the other GPUs (same scene, their own slot pools and
	// accumulators).  See the "device groups" section below.
	std::vector<rtb_ctx*> peers;
	std::vector<char> peerDirect;   // peers[k]'s accumulators are readable from this device (P2P over NVLink)
	rtb_params userParams;          // group: what the caller set (members get a composed partition per render call)
	cudaEvent_t evGroup = nullptr;  // member: "my work so far is done" for the group's read-out
	bool accumDirty = false;        // member: accumulators changed since the last gather
	int reduceMode = 0;             // 0 = P2P kernel where peer access exists, NCCL otherwise; 1 = force NCCL; 2 = staged copies
	void* ncclLib = nullptr;
	void* ncclComms[8] = {};
	long long* gatherScratch = nullptr;
	uint64_t gathers = 0, gatherP2P = 0, gatherNccl = 0;
};

namespace
{
int fail(rtb_ctx* ctx, int code, const char* fmt, ...)
{
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	if (ctx) ctx->error = buf;
	else g_createError = buf;
	return code;
}

#define CK(call)                                                                                        \
	do                                                                                                  \
	{                                                                                                   \
		cudaError_t e_ = (call);                                                                        \
		if (e_ != cudaSuccess)                                                                          \
			return fail(ctx, e_ == cudaErrorMemoryAllocation ? RTB_ERR_OOM : RTB_ERR_CUDA, "%s: %s", #call, \
			            cudaGetErrorString(e_));                                                        \
	} while (0)

int bind(rtb_ctx* ctx)
{
	CK(cudaSetDevice(ctx->device));
	return RTB_OK;
}

void freeScene(rtb_ctx* ctx)
{
	for (void* p : ctx->sceneAllocs) cudaFree(p);
	ctx->sceneAllocs.clear();
	if (ctx->film) cudaFree(ctx->film);
	if (ctx->filmFiltered) cudaFree(ctx->filmFiltered);
	if (ctx->tone) cudaFree(ctx->tone);
	if (ctx->wfState) cudaFree(ctx->wfState);
	if (ctx->wfCtrl) cudaFree(ctx->wfCtrl);
	if (ctx->wfGlobal) cudaFree(ctx->wfGlobal);
	if (ctx->wfTiles) cudaFree(ctx->wfTiles);
	if (ctx->wfPrimary) cudaFree(ctx->wfPrimary);
	if (ctx->wfPerm) cudaFree(ctx->wfPerm);
	if (ctx->wfSortWork) cudaFree(ctx->wfSortWork);
	ctx->wfPerm = nullptr, ctx->wfPermEntries = 0, ctx->wfSortWork = nullptr;
	if (ctx->vpls) cudaFree(ctx->vpls);
	if (ctx->vplCounts) cudaFree(ctx->vplCounts);
	ctx->vpls = nullptr, ctx->vplCounts = nullptr, ctx->vplPaths = 0;
	if (ctx->accumScratch) cudaFree(ctx->accumScratch);
	if (ctx->wfTilesAdaptive) cudaFree(ctx->wfTilesAdaptive);
	if (ctx->adaptJobBase) cudaFree(ctx->adaptJobBase);
	if (ctx->adaptSamples) cudaFree(ctx->adaptSamples);
	if (ctx->adaptVariance) cudaFree(ctx->adaptVariance);
	ctx->accumScratch = nullptr, ctx->wfTilesAdaptive = nullptr, ctx->adaptJobBase = nullptr, ctx->adaptSamples = nullptr;
	ctx->adaptVariance = nullptr;
	if (ctx->accum) cudaFree(ctx->accum);
	ctx->wfState = nullptr, ctx->wfStateBytes = 0;
	ctx->wfCtrl = nullptr, ctx->wfCtrlEntries = 0;
	ctx->wfGlobal = nullptr, ctx->wfTiles = nullptr, ctx->wfTileCount = 0;
	ctx->wfPrimary = nullptr;
	ctx->wfTilePart[0] = ctx->wfTilePart[1] = ctx->wfTilePart[2] = -1;
	ctx->accum = nullptr;
	ctx->film = ctx->filmFiltered = nullptr;
	ctx->tone = nullptr;
	ctx->haveScene = false;
}

template <class T>
int uploadArray(rtb_ctx* ctx, const T* host, size_t n, const T** dev)
{
	*dev = nullptr;
	if (n == 0) return RTB_OK;
	void* p = nullptr;
	CK(cudaMalloc(&p, n * sizeof(T)));
	ctx->sceneAllocs.push_back(p);
	CK(cudaMemcpyAsync(p, host, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
	*dev = (const T*)p;
	return RTB_OK;
}

// scratch device buffer for the batched entry points
struct Scratch
{
	std::vector<void*> ptrs;
	~Scratch()
	{
		for (void* p : ptrs) cudaFree(p);
	}
	template <class T>
	cudaError_t in(const T* host, size_t n, T** dev, cudaStream_t s)
	{
		*dev = nullptr;
		if (!host || n == 0) return cudaSuccess;
		void* p = nullptr;
		cudaError_t e = cudaMalloc(&p, n * sizeof(T));
		if (e != cudaSuccess) return e;
		ptrs.push_back(p);
		*dev = (T*)p;
		return cudaMemcpyAsync(p, host, n * sizeof(T), cudaMemcpyHostToDevice, s);
	}
	template <class T>
	cudaError_t out(const T* hostWanted, size_t n, T** dev)
	{
		*dev = nullptr;
		if (!hostWanted || n == 0) return cudaSuccess;
		void* p = nullptr;
		cudaError_t e = cudaMalloc(&p, n * sizeof(T));
		if (e != cudaSuccess) return e;
		ptrs.push_back(p);
		*dev = (T*)p;
		return cudaSuccess;
	}
};

cudaEvent_t getEvent(rtb_ctx* ctx)
{
	if (!ctx->eventPool.empty())
	{
		cudaEvent_t e = ctx->eventPool.back();
		ctx->eventPool.pop_back();
		return e;
	}
	cudaEvent_t e = nullptr;
	cudaEventCreate(&e);
	return e;
}

void resolveTimings(rtb_ctx* ctx)
{
	for (EventPair& p : ctx->pending)
	{
		float ms = 0.0f;
		if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess)
			ctx->renderMs += ms;
		ctx->eventPool.push_back(p.a);
		ctx->eventPool.push_back(p.b);
	}
	ctx->pending.clear();
	for (StageEvents& s : ctx->pendingStages)
	{
		if (cudaEventSynchronize(s.e[3]) == cudaSuccess)
		{
			for (int k = 0; k < 3; k++)
			{
				float ms = 0.0f;
				if (cudaEventElapsedTime(&ms, s.e[k], s.e[k + 1]) == cudaSuccess) ctx->stageMs[k] += ms;
			}
			ctx->timedIterations++;
		}
		for (int k = 0; k < 4; k++) ctx->eventPool.push_back(s.e[k]);
	}
	ctx->pendingStages.clear();
}

int checkTrav(rtb_ctx* ctx, int traversal)
{
	if (traversal < RTB_TRAV_EXACT || traversal > RTB_TRAV_Q16) return fail(ctx, RTB_ERR_ARG, "bad traversal %d", traversal);
	return RTB_OK;
}
} // namespace

template <int TRAV>
static void launchRender(rtb_ctx* ctx, const RenderArgs& A, dim3 grid, dim3 block)
{
	switch (ctx->params.integrator)
	{
	case RTB_INT_DIRECT: k_render<TRAV, RTB_INT_DIRECT><<<grid, block, 0, ctx->stream>>>(ctx->S, A); break;
	case RTB_INT_ALBEDO: k_render<TRAV, RTB_INT_ALBEDO><<<grid, block, 0, ctx->stream>>>(ctx->S, A); break;
	case RTB_INT_NORMALS: k_render<TRAV, RTB_INT_NORMALS><<<grid, block, 0, ctx->stream>>>(ctx->S, A); break;
	case RTB_INT_PATH_MIS: k_render<TRAV, RTB_INT_PATH_MIS><<<grid, block, 0, ctx->stream>>>(ctx->S, A); break;
	default: k_render<TRAV, RTB_INT_PATH><<<grid, block, 0, ctx->stream>>>(ctx->S, A); break;
	}
}


static int wfSettle(rtb_ctx* ctx);

static int renderMegakernel(rtb_ctx* ctx, uint32_t spp_begin, uint32_t spp_count)
{
	if (int rc = wfSettle(ctx)) return rc;
	RenderArgs A;
	A.accum = ctx->accum;
	A.counters = ctx->counters;
	A.spp_begin = spp_begin, A.spp_count = spp_count;
	A.width = ctx->width, A.height = ctx->height;
	A.P = ctx->params;
	uint32_t warps = ((ctx->width + 7) / 8) * ((ctx->height + 3) / 4);
	dim3 block(64), grid((warps + 1) / 2);
	EventPair ev = {getEvent(ctx), getEvent(ctx)};
	cudaEventRecord(ev.a, ctx->stream);
	RTB_TRAV_SWITCH(ctx->params.traversal, launchRender<TR>(ctx, A, grid, block));
	cudaEventRecord(ev.b, ctx->stream);
	ctx->pending.push_back(ev);
	ctx->launches++;
	CK(cudaGetLastError());
	ctx->filmDirty = true;
	return RTB_OK;
}

// float film <- fixed-point sums (only when they changed)
static int resolveFilm(rtb_ctx* ctx)
{
	if (!ctx->filmDirty) return RTB_OK;
	uint32_t n = ctx->width * ctx->height * 3;
	k_wf_resolve<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->accum, ctx->film, n);
	ctx->launches++;
	CK(cudaGetLastError());
	ctx->filmDirty = false;
	return RTB_OK;
}

// Wavefront schedule (rtb_wavefront.cuh).  The slot pool is split into K independent sub-pools,
// each advancing on its own stream (all draw jobs from one counter): a stage kernel ends with a
// long, thinly occupied tail (its slowest rays), which the other sub-pools' kernels fill.  The
// number of iterations depends on the paths, so launches are enqueued in batches sized from the
// measured job rate; between batches the host reads {jobs claimed, slots alive} (a few dozen
// bytes + a sync, a handful of times per call).  Work is forked from / joined to ctx->stream.
struct AdaptivePlan
{
	unsigned long long totalJobs; // sum over tiles of 1024 * samples
	uint32_t nTiles32;
};

template <int TRAV>
static int32_t travRootHost(const DevScene& S)
{
	return TRAV == RTB_TRAV_WIDE ? S.wide_root : TRAV == RTB_TRAV_Q16 ? S.q16_root : S.fast_root;
}

// enqueues iterations [R.it, end) of the run on its sub-pools' streams
static void wfLaunch(rtb_ctx* ctx, WfRun& R, uint32_t end)
{
	ctx->wfIterations += (uint64_t)(end - R.it) * R.K;
	for (; R.it < end; R.it++)
	{
			for (int k = 0; k < R.K; k++)
			{
				cudaStream_t st = ctx->poolStreams[k];
				// stage timing: sub-pool 0, every 8th iteration (other sub-pools run concurrently)
				bool timed = (k == 0) && (R.it % 8u) == 4u;
				StageEvents se;
				if (timed)
				{
					for (int e = 0; e < 4; e++) se.e[e] = getEvent(ctx);
					cudaEventRecord(se.e[0], st);
				}
				// long rays (the scenes that also take the persistent any-hit kernel): smaller chunks shorten the tail of a
				// persistent launch (+1.5 ... 2 %), short rays prefer fewer atomics (profiles/r02_refill_chunk_sweep.txt)
				R.A[k].chunk = ctx->chunkOverride ? ctx->chunkOverride : (ctx->shadowPersistentAuto ? 64u : (uint32_t)WF_CHUNK);
				if (R.sortEx && !ctx->simpleExtend && R.ti != RTB_TRAV_EXACT)
				{
					k_sort_count<1><<<R.gridSort, 256, 0, st>>>(ctx->S, R.A[k], R.it);
					k_sort_scan<1><<<1, 1024, 0, st>>>(R.A[k], R.it);
					k_sort_scatter<1><<<R.gridSort, 256, 0, st>>>(ctx->S, R.A[k], R.it);
					ctx->launches += 3;
				}
				if (ctx->simpleExtend)
				{
					RTB_TRAV_SWITCH(R.ti, k_wf_extend_simple<TR><<<(R.perPool + 127) / 128, 128, 0, st>>>(ctx->S, R.A[k], R.it));
				}
				else if (R.cwKernels)
				{
					unsigned g = (unsigned)(ctx->smCount * ctx->cwBlocksPerSM[0]);
					unsigned need = (R.perPool + WF_CW_THREADS - 1) / WF_CW_THREADS;
					k_wf_trace_cw<false><<<g < need ? g : need, WF_CW_THREADS, ctx->cwSmemBytes, st>>>(ctx->S, R.A[k], R.it, ctx->cwStageNodes, ctx->cwStageLeaves);
				}
				else
				{
					RTB_TRAV_SWITCH(R.ti, k_wf_extend<TR><<<R.gridExtend, 128, 0, st>>>(ctx->S, R.A[k], R.it));
				}
				if (timed) cudaEventRecord(se.e[1], st);
				// the shade stage refills the shadow queue: the previous iteration's shadow stage must be done with it
				if (R.shadows && ctx->shadowAsync) cudaStreamWaitEvent(st, ctx->evShadowed[k], 0);
#define RTB_SHADE_LAUNCH(INTEG)                                                                      \
	do                                                                                               \
	{                                                                                                \
		if (R.A[k].primary) k_wf_shade<INTEG, true><<<R.gridSlots, 128, 0, st>>>(ctx->S, R.A[k], R.it);      \
		else k_wf_shade<INTEG, false><<<R.gridSlots, 128, 0, st>>>(ctx->S, R.A[k], R.it);                  \
	} while (0)
				switch (R.P.integrator)
				{
				case RTB_INT_DIRECT: RTB_SHADE_LAUNCH(RTB_INT_DIRECT); break;
				case RTB_INT_ALBEDO: RTB_SHADE_LAUNCH(RTB_INT_ALBEDO); break;
				case RTB_INT_NORMALS: RTB_SHADE_LAUNCH(RTB_INT_NORMALS); break;
				case RTB_INT_PATH_MIS: RTB_SHADE_LAUNCH(RTB_INT_PATH_MIS); break;
				default: RTB_SHADE_LAUNCH(RTB_INT_PATH); break;
				}
#undef RTB_SHADE_LAUNCH
				ctx->launches += 2;
				if (timed) cudaEventRecord(se.e[2], st);
				cudaStream_t sst = st;
				if (R.shadows)
				{
					if (ctx->shadowAsync)
					{
						sst = ctx->shadowStreams[k];
						cudaEventRecord(ctx->evShaded[k], st);
						cudaStreamWaitEvent(sst, ctx->evShaded[k], 0);
					}
					if (R.sortSh)
					{
						k_sort_count<0><<<R.gridSort, 256, 0, sst>>>(ctx->S, R.A[k], R.it);
						k_sort_scan<0><<<1, 1024, 0, sst>>>(R.A[k], R.it);
						k_sort_scatter<0><<<R.gridSort, 256, 0, sst>>>(ctx->S, R.A[k], R.it);
						ctx->launches += 3;
					}
					const bool persistShadow = !R.cwShadow && R.P.integrator != RTB_INT_PATH_MIS && R.ti != RTB_TRAV_EXACT && R.ti != RTB_TRAV_CW &&
					                           (ctx->shadowPersistent < 0 ? ctx->shadowPersistentAuto : ctx->shadowPersistent != 0);
					if (persistShadow)
					{
						RTB_TRAV_SWITCH(R.ti, if (travRootHost<TR>(ctx->S) >= 0) k_wf_shadow_persist<TR><<<R.gridExtend, 128, 0, sst>>>(ctx->S, R.A[k], R.it);
						                else k_wf_shadow<TR><<<R.gridSlots, 128, 0, sst>>>(ctx->S, R.A[k], R.it));
					}
					else if (R.cwShadow)
					{
						unsigned g = (unsigned)(ctx->smCount * ctx->cwBlocksPerSM[1]);
						k_wf_trace_cw<true><<<g, WF_CW_THREADS, ctx->cwSmemBytes, sst>>>(ctx->S, R.A[k], R.it, ctx->cwStageNodes, ctx->cwStageLeaves);
					}
					else if (R.P.integrator == RTB_INT_PATH_MIS)
					{
						RTB_TRAV_SWITCH(R.ti, k_wf_mis<TR><<<R.gridSlots, 128, 0, sst>>>(ctx->S, R.A[k], R.it));
					}
					else
					{
						RTB_TRAV_SWITCH(R.ti, k_wf_shadow<TR><<<R.gridSlots, 128, 0, sst>>>(ctx->S, R.A[k], R.it));
					}
					ctx->launches++;
					if (ctx->shadowAsync) cudaEventRecord(ctx->evShadowed[k], sst);
				}
				if (timed)
				{
					cudaEventRecord(se.e[3], sst);
					ctx->pendingStages.push_back(se);
				}
			}
	}
}

// device -> host: {shadow rays, slots alive} of the last enqueued iteration of every sub-pool, jobs claimed, work counters
static int wfProbeIssue(rtb_ctx* ctx, WfRun& R)
{
	for (int k = 0; k < R.K; k++)
		CK(cudaMemcpyAsync(&ctx->hostProbe[1 + k], &R.A[k].ctrl[R.it - 1], 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->poolStreams[k]));
	CK(cudaMemcpyAsync(&ctx->hostProbe[0], &ctx->wfGlobal->nextJob, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->poolStreams[0]));
	CK(cudaMemcpyAsync(&ctx->hostProbe[1 + RTB_MAX_POOLS], ctx->counters, RTB_COUNTER_STRIPES * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
	                   ctx->poolStreams[0]));
	return RTB_OK;
}

// after the probe's copies have completed: drained?  otherwise the size of the next batch
static void wfProbeRead(rtb_ctx* ctx, WfRun& R)
{
	ctx->wfHostSyncs++;
	{
		// Which any-hit kernel suits this scene: long shadow rays (bathroom: 51 box tests per ray with the optimised
		// tree, 65 without; the soups: 130+) gain 6 ... 12 % from the persistent kernel, short ones (coffee: 33 ... 38, the
		// small scenes: 9 ... 14) lose 5 ... 12 % (profiles/r02_persistent_shadow.txt, r02_tree_opt.txt).  The threshold
		// sits midway between the two measured groups.  Decided from the work counters the probe brings along anyway; the
		// film does not depend on the choice.
		unsigned long long rays = 0, boxes = 0;
		for (int r = 0; r < RTB_COUNTER_STRIPES; r++)
			rays += ctx->hostProbe[1 + RTB_MAX_POOLS + r * 8 + 2], boxes += ctx->hostProbe[1 + RTB_MAX_POOLS + r * 8 + 5];
		if (rays > 100000ull) ctx->shadowPersistentAuto = boxes > 42ull * rays;
	}
	unsigned long long claimed = ctx->hostProbe[0];
	uint32_t alive = 0;
	for (int k = 0; k < R.K; k++) alive += (uint32_t)(ctx->hostProbe[1 + k] >> 32); // WfCtrl{nShadow, alive}
	if (alive == 0)
	{
		R.drained = true;
		return;
	}
	// predict the remaining iterations from the job rate seen so far
	unsigned long long started = claimed < R.totalJobs ? claimed : R.totalJobs;
	double perIter = (started > R.nSlotsAlloc) ? (double)(started - R.nSlotsAlloc) / (double)R.it : 0.0;
	double remaining = (double)(R.totalJobs - started);
	double predict = (perIter > 0.0) ? remaining / perIter : (double)R.batch * 2.0;
	uint32_t next = (uint32_t)(predict * 0.9);
	if (remaining == 0.0) next = R.vertices; // tail: the last paths finish within `vertices` iterations
	if (next < 4) next = 4;
	if (next > 4096) next = 4096;
	R.batch = next;
}

// the sub-pools' streams rejoin ctx->stream; the render's device time runs from `ev.a` to here
static void wfJoin(rtb_ctx* ctx, WfRun& R, EventPair ev)
{
	for (int k = 0; k < R.K; k++)
	{
		if (R.shadows && ctx->shadowAsync) cudaStreamWaitEvent(ctx->poolStreams[k], ctx->evShadowed[k], 0);
		cudaEventRecord(ctx->poolDone[k], ctx->poolStreams[k]);
		cudaStreamWaitEvent(ctx->stream, ctx->poolDone[k], 0);
	}
	cudaEventRecord(ev.b, ctx->stream);
	ctx->pending.push_back(ev);
}

// batches + probes until the pool has drained (the synchronous part of a render)
static int wfRunToEnd(rtb_ctx* ctx, WfRun& R)
{
	while (!R.drained && R.it < R.bound)
	{
		uint32_t end = R.it + R.batch;
		if (end > R.bound) end = R.bound;
		wfLaunch(ctx, R, end);
		if (int rc = wfProbeIssue(ctx, R)) return rc;
		for (int k = 0; k < R.K; k++) CK(cudaStreamSynchronize(ctx->poolStreams[k]));
		wfProbeRead(ctx, R);
	}
	if (R.drained)
	{
		// how many iterations the sub-pools really needed: the first whose shade stage left no slot alive
		uint32_t lookBack = R.it < 4096u ? R.it : 4096u, needed = R.it;
		std::vector<WfCtrl> tail((size_t)lookBack);
		uint32_t worst = 0;
		for (int k = 0; k < R.K; k++)
		{
			CK(cudaMemcpy(tail.data(), &R.A[k].ctrl[R.it - lookBack], (size_t)lookBack * sizeof(WfCtrl), cudaMemcpyDeviceToHost));
			uint32_t first = R.it;
			for (uint32_t j = 0; j < lookBack; j++)
				if (tail[j].alive == 0)
				{
					first = R.it - lookBack + j + 1;
					break;
				}
			if (first > worst) worst = first;
		}
		if (worst) needed = worst;
		ctx->iterHint[R.key] = needed;
	}
	return RTB_OK;
}

// Completes a render that rtb_render left in flight (WfRun::active): waits for its probe, and if the remembered iteration
// count fell short, enqueues the rest.  Called by every entry point before it touches the context.
static int wfSettle(rtb_ctx* ctx)
{
	WfRun& R = ctx->run;
	if (!R.active) return RTB_OK;
	R.active = false;
	if (int rc = bind(ctx)) return rc;
	CK(cudaEventSynchronize(ctx->evProbe));
	wfProbeRead(ctx, R);
	if (!R.drained)
	{
		EventPair ev = {getEvent(ctx), getEvent(ctx)};
		cudaEventRecord(ev.a, ctx->stream);
		cudaEventRecord(ctx->evFork, ctx->stream);
		for (int k = 0; k < R.K; k++) cudaStreamWaitEvent(ctx->poolStreams[k], ctx->evFork, 0);
		int rc = wfRunToEnd(ctx, R);
		wfJoin(ctx, R, ev);
		if (rc) return rc;
		CK(cudaGetLastError());
		if (!R.drained) return fail(ctx, RTB_ERR_STATE, "wavefront did not drain within its iteration bound (%u)", R.bound);
	}
	return RTB_OK;
}

// the context and, for a device group, every member
static int wfSettleAll(rtb_ctx* ctx)
{
	for (rtb_ctx* m : ctx->peers)
		if (int rc = wfSettle(m))
		{
			ctx->error = m->error;
			return rc;
		}
	int rc = wfSettle(ctx);
	if (!ctx->peers.empty()) cudaSetDevice(ctx->device);
	return rc;
}

static int renderWavefront(rtb_ctx* ctx, uint32_t spp_begin, uint32_t spp_count, const AdaptivePlan* plan = nullptr)
{
	if (int rc = wfSettle(ctx)) return rc; // the previous render of this context may still be in flight
	const rtb_params& P = ctx->params;
	uint32_t sFirst = spp_begin, sStep = 1, sCount = spp_count;
	if (!plan && P.partition == RTB_PART_SPP && P.part_world > 1)
	{
		uint32_t w = (uint32_t)P.part_world, r = (uint32_t)P.part_rank;
		sFirst = spp_begin + (r + w - (spp_begin % w)) % w;
		sStep = w;
		uint64_t end = (uint64_t)spp_begin + spp_count;
		sCount = (sFirst < end) ? (uint32_t)((end - 1 - sFirst) / w + 1) : 0u;
	}
	if (sCount == 0) return RTB_OK;
	// ---- owned 8x4 tiles (TILE_SIZE 32 tiles dealt round-robin, Renderer.h:18)
	int part[3] = {P.partition == RTB_PART_TILE ? 1 : 0, P.partition == RTB_PART_TILE ? P.part_rank : 0,
	               P.partition == RTB_PART_TILE ? P.part_world : 1};
	if (!ctx->wfTiles || memcmp(part, ctx->wfTilePart, sizeof(part)) != 0)
	{
		uint32_t tilesX = (ctx->width + 7) / 8, tilesY = (ctx->height + 3) / 4, t32x = (ctx->width + 31) / 32;
		if (tilesX > 0xFFFFu || tilesY > 0xFFFEu) return fail(ctx, RTB_ERR_ARG, "film larger than 524280 x 262136 pixels");
		std::vector<uint32_t> tiles;
		tiles.reserve((size_t)tilesX * tilesY);
		for (uint32_t ty = 0; ty < tilesY; ty++)
			for (uint32_t tx = 0; tx < tilesX; tx++)
			{
				uint32_t tile32 = ((ty * 4) >> 5) * t32x + ((tx * 8) >> 5);
				if (part[0] && part[2] > 1 && (int)(tile32 % (uint32_t)part[2]) != part[1]) continue;
				tiles.push_back((ty << 16) | tx);
			}
		CK(cudaStreamSynchronize(ctx->stream));
		if (ctx->wfTiles) cudaFree(ctx->wfTiles);
		ctx->wfTiles = nullptr;
		ctx->wfTileCount = (uint32_t)tiles.size();
		if (!tiles.empty())
		{
			CK(cudaMalloc((void**)&ctx->wfTiles, tiles.size() * sizeof(uint32_t)));
			CK(cudaMemcpy(ctx->wfTiles, tiles.data(), tiles.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
		}
		memcpy(ctx->wfTilePart, part, sizeof(part));
	}
	if (ctx->wfTileCount == 0) return RTB_OK;
	unsigned long long totalJobs = plan ? plan->totalJobs : (unsigned long long)ctx->wfTileCount * 32ull * sCount;
	if (totalJobs == 0) return RTB_OK;
	// ---- pools
	if (!ctx->poolStreams[0])
	{
		if (const char* e = getenv("RTB_POOLS"))
		{
			int v = atoi(e);
			if (v >= 1 && v <= RTB_MAX_POOLS) ctx->pools = v;
		}
		for (int k = 0; k < RTB_MAX_POOLS; k++)
		{
			CK(cudaStreamCreateWithFlags(&ctx->poolStreams[k], cudaStreamNonBlocking));
			CK(cudaEventCreateWithFlags(&ctx->poolDone[k], cudaEventDisableTiming));
			CK(cudaStreamCreateWithFlags(&ctx->shadowStreams[k], cudaStreamNonBlocking));
			CK(cudaEventCreateWithFlags(&ctx->evShaded[k], cudaEventDisableTiming));
			CK(cudaEventCreateWithFlags(&ctx->evShadowed[k], cudaEventDisableTiming));
		}
		if (const char* e = getenv("RTB_SHADOW_ASYNC")) ctx->shadowAsync = atoi(e) != 0;
		if (const char* e = getenv("RTB_ASYNC_RENDER")) ctx->asyncRender = atoi(e);
		CK(cudaEventCreateWithFlags(&ctx->evFork, cudaEventDisableTiming));
	}
	uint32_t nSlotsAll = ctx->poolSlots;
	if ((unsigned long long)nSlotsAll > totalJobs) nSlotsAll = (uint32_t)((totalJobs + 31ull) & ~31ull);
	int K = ctx->pools;
	while (K > 1 && nSlotsAll / (uint32_t)K < (1u << 19)) K--; // sub-pools of at least 512 k slots
	uint32_t perPool = ((nSlotsAll + (uint32_t)K - 1) / (uint32_t)K + 31u) & ~31u;
	uint32_t nSlotsAlloc = perPool * (uint32_t)K;
	// shadow queue: one entry per slot and launch, WF_SHADOW_PER_SLOT with the multi-pass shade stage
	const size_t F = P.primary_reuse ? WF_SHADOW_PER_SLOT : 1u;
	const size_t RQ = (P.integrator == RTB_INT_PATH_MIS) ? 6u : 3u; // float4 per queue record
	size_t need = (size_t)nSlotsAlloc * (4 + RQ * F) * sizeof(float4);
	if (need > ctx->wfStateBytes)
	{
		CK(cudaStreamSynchronize(ctx->stream));
		if (ctx->wfState) cudaFree(ctx->wfState);
		ctx->wfState = nullptr, ctx->wfStateBytes = 0;
		CK(cudaMalloc(&ctx->wfState, need));
		ctx->wfStateBytes = need;
	}
	if (const char* e = getenv("RTB_SORT_SHADOW")) ctx->sortShadow = atoi(e);
	if (const char* e = getenv("RTB_SORT_EXTEND")) ctx->sortExtend = atoi(e);
	const bool sortSh = ctx->sortShadow < 0 ? ctx->sortShadowAuto : ctx->sortShadow != 0;
	const bool sortEx = ctx->sortExtend < 0 ? ctx->sortExtendAuto : ctx->sortExtend != 0;
	if (sortSh || sortEx)
	{
		size_t entries = (size_t)nSlotsAlloc * (F + 1);
		if (entries > ctx->wfPermEntries)
		{
			CK(cudaStreamSynchronize(ctx->stream));
			if (ctx->wfPerm) cudaFree(ctx->wfPerm);
			ctx->wfPerm = nullptr, ctx->wfPermEntries = 0;
			CK(cudaMalloc((void**)&ctx->wfPerm, entries * sizeof(uint32_t)));
			ctx->wfPermEntries = entries;
		}
		if (!ctx->wfSortWork) CK(cudaMalloc((void**)&ctx->wfSortWork, (size_t)RTB_MAX_POOLS * 4 * WF_SORT_BUCKETS * sizeof(uint32_t)));
		CK(cudaMemsetAsync(ctx->wfSortWork, 0, (size_t)RTB_MAX_POOLS * 4 * WF_SORT_BUCKETS * sizeof(uint32_t), ctx->stream));
	}
	if (!ctx->wfGlobal) CK(cudaMalloc((void**)&ctx->wfGlobal, sizeof(WfGlobal)));
	if (!ctx->hostProbe) CK(cudaMallocHost((void**)&ctx->hostProbe, (1 + RTB_MAX_POOLS + RTB_COUNTER_STRIPES * 8) * sizeof(unsigned long long)));
	uint32_t vertices = (P.integrator == RTB_INT_PATH || P.integrator == RTB_INT_PATH_MIS) ? (uint32_t)P.max_depth + 2u : 1u;
	// list-scheduling bound on the iterations of a sub-pool: total work / slots + longest job
	unsigned long long bound64 = (totalJobs * vertices + perPool - 1) / perPool + vertices + 1;
	if (bound64 > (1ull << 22)) return fail(ctx, RTB_ERR_ARG, "too many samples per call; split the render");
	uint32_t bound = (uint32_t)bound64;
	if ((size_t)bound * K > ctx->wfCtrlEntries)
	{
		CK(cudaStreamSynchronize(ctx->stream));
		if (ctx->wfCtrl) cudaFree(ctx->wfCtrl);
		ctx->wfCtrl = nullptr, ctx->wfCtrlEntries = 0;
		CK(cudaMalloc((void**)&ctx->wfCtrl, (size_t)bound * K * sizeof(WfCtrl)));
		ctx->wfCtrlEntries = (uint32_t)((size_t)bound * K);
	}
	CK(cudaMemsetAsync(ctx->wfCtrl, 0, (size_t)bound * K * sizeof(WfCtrl), ctx->stream));
	CK(cudaMemsetAsync(ctx->wfGlobal, 0, sizeof(WfGlobal), ctx->stream));
	if (ctx->smCount <= 0)
	{
		cudaDeviceProp prop;
		CK(cudaGetDeviceProperties(&prop, ctx->device));
		ctx->smCount = prop.multiProcessorCount;
		CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->travBlocksPerSM[0], k_wf_extend<RTB_TRAV_EXACT>, 128, 0));
		CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->travBlocksPerSM[1], k_wf_extend<RTB_TRAV_FAST>, 128, 0));
		CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->travBlocksPerSM[2], k_wf_extend<RTB_TRAV_WIDE>, 128, 0));
		ctx->travBlocksPerSM[3] = ctx->travBlocksPerSM[1];
		CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->travBlocksPerSM[4], k_wf_extend<RTB_TRAV_Q16>, 128, 0));
		if (const char* e = getenv("RTB_CW_STAGE_KB"))
		{
			int v = atoi(e);
			if (v >= 0 && v <= 200) ctx->cwStageKB = v;
		}
		if (const char* e = getenv("RTB_CW_SHADOW")) ctx->cwShadowPersistent = atoi(e);
		if (const char* e = getenv("RTB_SHADOW_PERSISTENT")) ctx->shadowPersistent = atoi(e);
		if (const char* e = getenv("RTB_CHUNK"))
		{
			int v = atoi(e);
			if (v >= 32 && v <= 4096) ctx->chunkOverride = (uint32_t)v & ~31u;
		}
		if (const char* e = getenv("RTB_SIMPLE_EXTEND")) ctx->simpleExtend = atoi(e) != 0;
		if (const char* e = getenv("RTB_PRIMARY_PASSES"))
		{
			int v = atoi(e);
			if (v >= 1 && v <= 64) ctx->primaryPasses = v;
		}
	}
	int ti = P.traversal;
	const bool cwKernels = ti == RTB_TRAV_CW && ctx->S.cw_valid && !ctx->simpleExtend;
	if (cwKernels && ctx->cwBlocksPerSM[0] == 0)
	{
		// top of the tree in shared memory: as many breadth-first nodes as the per-block budget holds, then the exact
		// leaf boxes if ALL of them fit in what is left (small scenes live in shared memory entirely)
		uint32_t budget = (uint32_t)ctx->cwStageKB * 1024u;
		uint32_t sn = ctx->S.n_cwnodes < budget / 80u ? ctx->S.n_cwnodes : budget / 80u;
		uint32_t left = budget - sn * 80u;
		uint32_t sl = (sn == ctx->S.n_cwnodes && (size_t)ctx->S.n_cwleaves * 32u <= left) ? ctx->S.n_cwleaves : 0u;
		ctx->cwStageNodes = sn, ctx->cwStageLeaves = sl, ctx->cwSmemBytes = sn * 80u + sl * 32u;
		// the attribute belongs to the function, not to this context: always the largest budget any context may ask for
		CK(cudaFuncSetAttribute(k_wf_trace_cw<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
		CK(cudaFuncSetAttribute(k_wf_trace_cw<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
		CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->cwBlocksPerSM[0], k_wf_trace_cw<false>, WF_CW_THREADS, ctx->cwSmemBytes));
		CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->cwBlocksPerSM[1], k_wf_trace_cw<true>, WF_CW_THREADS, ctx->cwSmemBytes));
		if (ctx->cwBlocksPerSM[0] < 1 || ctx->cwBlocksPerSM[1] < 1) return fail(ctx, RTB_ERR_STATE, "RTB_TRAV_CW: %u bytes of shared memory per block do not fit", ctx->cwSmemBytes);
	}
	const bool cwShadow = cwKernels && (ctx->cwShadowPersistent < 0 ? true : ctx->cwShadowPersistent != 0) && P.integrator != RTB_INT_PATH_MIS;
	if (P.primary_reuse && !ctx->wfPrimary) CK(cudaMalloc((void**)&ctx->wfPrimary, (size_t)ctx->width * ctx->height * sizeof(float4)));
	WfRun& R = ctx->run;
	WfArgs* A = R.A;
	float4* base = (float4*)ctx->wfState;
	for (int k = 0; k < K; k++)
	{
		WfArgs& a = A[k];
		size_t off = (size_t)k * perPool, n = nSlotsAlloc;
		a.rayO = base + off, a.rayD = base + n + off, a.hit = base + n * 2 + off, a.thr = base + n * 3 + off;
		a.shO = base + n * 4 + off * F, a.shD = base + n * (4 + F) + off * F, a.shC = base + n * (4 + 2 * F) + off * F;
		a.misA = base + n * (4 + 3 * F) + off * F, a.misB = base + n * (4 + 4 * F) + off * F, a.misC = base + n * (4 + 5 * F) + off * F;
		a.ctrl = ctx->wfCtrl + (size_t)k * bound;
		a.glob = ctx->wfGlobal;
		a.tileList = plan ? ctx->wfTilesAdaptive : ctx->wfTiles;
		a.tileJobBase = plan ? ctx->adaptJobBase : nullptr;
		a.nTiles32 = plan ? plan->nTiles32 : 0u;
		a.shPerm = sortSh ? ctx->wfPerm + off * F : nullptr;
		a.exPerm = sortEx ? ctx->wfPerm + (size_t)nSlotsAlloc * F + off : nullptr;
		a.sortHist = ctx->wfSortWork ? ctx->wfSortWork + (size_t)k * 4 * WF_SORT_BUCKETS : nullptr;
		a.sortCursor = ctx->wfSortWork ? ctx->wfSortWork + (size_t)k * 4 * WF_SORT_BUCKETS + 2 * WF_SORT_BUCKETS : nullptr;
		a.primary = P.primary_reuse ? ctx->wfPrimary : nullptr;
		a.primaryPasses = (uint32_t)ctx->primaryPasses;
		a.counters = ctx->counters;
		a.accum = ctx->accum;
		a.chunk = WF_CHUNK;
		a.nSlots = perPool, a.nTiles = plan ? plan->nTiles32 * 32u : ctx->wfTileCount;
		a.width = ctx->width, a.height = ctx->height;
		a.sFirst = sFirst, a.sStep = sStep, a.sCount = sCount;
		a.totalJobs = totalJobs;
		a.P = P;
	}
	// persistent traversal kernel: one resident wave; grid-stride kernels: enough blocks to fill the machine
	int extendBlocks = ctx->travBlocksPerSM[ti] > 0 ? ctx->travBlocksPerSM[ti] : 1;
	if (const char* e = getenv("RTB_EXTEND_BLOCKS_PER_SM"))
	{
		int v = atoi(e);
		if (v >= 1 && v < extendBlocks) extendBlocks = v;
	}
	unsigned gridExtend = (unsigned)(ctx->smCount * extendBlocks);
	unsigned maxBlocks = (unsigned)ctx->smCount * 16u;
	unsigned gridSort = (unsigned)(((size_t)perPool * F + WF_SORT_CHUNK - 1) / WF_SORT_CHUNK);
	if (gridSort > (unsigned)ctx->smCount * 4u) gridSort = (unsigned)ctx->smCount * 4u;
	unsigned gridSlots = (perPool + 127) / 128;
	if (gridSlots > maxBlocks) gridSlots = maxBlocks;
	if (gridExtend > (perPool + 127) / 128) gridExtend = (perPool + 127) / 128;
	bool shadows = (P.integrator == RTB_INT_PATH || P.integrator == RTB_INT_DIRECT || P.integrator == RTB_INT_PATH_MIS);
	R.P = P, R.K = K, R.ti = ti, R.perPool = perPool, R.nSlotsAlloc = nSlotsAlloc, R.bound = bound, R.vertices = vertices;
	R.gridExtend = gridExtend, R.gridSlots = gridSlots, R.gridSort = gridSort;
	R.shadows = shadows, R.sortSh = sortSh, R.sortEx = sortEx, R.cwKernels = cwKernels, R.cwShadow = cwShadow, R.totalJobs = totalJobs;
	EventPair ev = {getEvent(ctx), getEvent(ctx)};
	cudaEventRecord(ev.a, ctx->stream);
	if (P.primary_reuse)
	{
		unsigned warps = ((ctx->width + 7) / 8) * ((ctx->height + 3) / 4);
		unsigned grid = (warps + 3) / 4;
		if (grid > maxBlocks) grid = maxBlocks;
		RTB_TRAV_SWITCH(ti, k_wf_primary<TR><<<grid, 128, 0, ctx->stream>>>(ctx->S, ctx->wfPrimary, ctx->width, ctx->height, P.epsilon,
		                                                                   P.cull_rel, ctx->counters));
		ctx->launches++;
	}
	// fork
	cudaEventRecord(ctx->evFork, ctx->stream);
	for (int k = 0; k < K; k++)
	{
		cudaStreamWaitEvent(ctx->poolStreams[k], ctx->evFork, 0);
		k_wf_init<<<(perPool + 255) / 256, 256, 0, ctx->poolStreams[k]>>>(ctx->S, A[k]);
		ctx->launches++;
	}
	R.it = 0;
	R.batch = R.vertices * 4 < 16 ? 16 : R.vertices * 4;
	R.drained = false;
	// A configuration that has been rendered before is enqueued in one go and NOT waited for: the remembered iteration
	// count + 6 % + 3 (launches past the drain exit at once); wfSettle checks later and enqueues the rest if need be.
	{
		unsigned long long h = 1469598103934665603ull;
		const unsigned long long parts[10] = {R.totalJobs, R.perPool, (unsigned long long)R.K, (unsigned long long)P.integrator, (unsigned long long)P.max_depth,
		                                       (unsigned long long)P.traversal, (unsigned long long)P.primary_reuse, (unsigned long long)P.sampling,
		                                       (unsigned long long)(plan ? 1 : 0), (unsigned long long)ctx->S.n_tris};
		for (unsigned long long v : parts) h = (h ^ v) * 1099511628211ull;
		R.key = h;
	}
	auto hint = ctx->iterHint.find(R.key);
	if (ctx->asyncRender && !plan && hint != ctx->iterHint.end())
	{
		if (!ctx->evProbe) CK(cudaEventCreateWithFlags(&ctx->evProbe, cudaEventDisableTiming));
		uint32_t n = hint->second + hint->second / 16u + 3u;
		if (ctx->asyncRender == 2) n = hint->second / 2u + 1u; // RTB_ASYNC_RENDER=2 (tests): deliberately too few, wfSettle must enqueue the rest
		if (n > R.bound) n = R.bound;
		wfLaunch(ctx, R, n);
		if (int rc = wfProbeIssue(ctx, R)) return rc;
		wfJoin(ctx, R, ev);
		CK(cudaEventRecord(ctx->evProbe, ctx->stream)); // ctx->stream now waits for every sub-pool stream, hence for the probe's copies
		CK(cudaGetLastError());
		R.active = true;
		ctx->wfAsyncRenders++;
		ctx->filmDirty = true;
		return RTB_OK;
	}
	int rc = wfRunToEnd(ctx, R);
	wfJoin(ctx, R, ev);
	if (rc) return rc;
	CK(cudaGetLastError());
	if (!R.drained) return fail(ctx, RTB_ERR_STATE, "wavefront did not drain within its iteration bound (%u)", R.bound);
	ctx->filmDirty = true;
	return RTB_OK;
}

static int filteredFilm(rtb_ctx* ctx, const float** src)
{
	if (int rc = resolveFilm(ctx)) return rc;
	*src = ctx->film;
	if (ctx->params.filter != RTB_FILTER_GAUSSIAN) return RTB_OK;
	size_t npx = (size_t)ctx->width * ctx->height;
	if (!ctx->filmFiltered) CK(cudaMalloc((void**)&ctx->filmFiltered, npx * 3 * sizeof(float)));
	int size = (int)ceilf(ctx->params.filter_radius);
	if (size < 0) size = 0;
	if (size > 2) size = 2; // filterWeights[25], RTBase/Imaging.h:212
	k_gaussian<<<(unsigned)((npx + 255) / 256), 256, 0, ctx->stream>>>(ctx->film, ctx->filmFiltered, (int)ctx->width, (int)ctx->height,
	                                                                      size, ctx->params.filter_radius, ctx->params.filter_alpha);
	ctx->launches++;
	CK(cudaGetLastError());
	*src = ctx->filmFiltered;
	return RTB_OK;
}


// ======================================================================================
// Device groups (rtb_create_multi).  RayTracer::init sizes the renderer to every processor of the machine
// (RTBase/Renderer.h:52-55) and pathTracerTileBased fans the tiles out over them (:836-853); here the
// "processors" are the GPUs of one box.  A group context is an ordinary context on its first device plus one
// complete single-device context per further GPU (scene replicated, own slot pool, own fixed-point film).
//   rtb_render      every member renders its slice of the call's samples (spp slice: perfect balance; tile slice
//                   when the call has fewer samples than devices or the caller asked for tiles), each driven by
//                   its own host thread, like the reference's per-call worker threads.
//   film read-out   the members' accumulators are SUMMED into device 0's: an exact int64 sum, so the film is
//                   bit-identical to a single-GPU render.  Where device 0 can read a peer's memory (NVLink P2P)
//                   ONE kernel on device 0 does reduce + int64 -> float conversion in a single pass over peer
//                   memory (k_film_gather); otherwise ncclReduce over a communicator created with
//                   ncclCommInitAll (libnccl.so.2 is loaded on demand); last resort: staged peer copies.
// Nothing else is exchanged between the GPUs.
// ======================================================================================
static size_t groupSize(const rtb_ctx* g) { return 1 + g->peers.size(); }

template <class F>
static int forEachMember(rtb_ctx* g, F fn)
{
	size_t n = groupSize(g);
	if (n == 1) return fn(g, 0);
	std::vector<int> rc(n, 0);
	std::vector<std::thread> th;
	th.reserve(n - 1);
	for (size_t d = 1; d < n; d++) th.emplace_back([&rc, &fn, g, d]() { rc[d] = fn(g->peers[d - 1], (int)d); });
	rc[0] = fn(g, 0);
	for (std::thread& t : th) t.join();
	for (size_t d = 1; d < n; d++)
		if (rc[d])
		{
			g->error = "device " + std::to_string(g->peers[d - 1]->device) + ": " + g->peers[d - 1]->error;
			return rc[d];
		}
	return rc[0];
}

// The partition member d of the group renders: the caller's own partition (rank r of w processes, e.g. one
// process per node) refined n-fold.  s = r (mod w) and (s - r) / w = d (mod n)  <=>  s = r + w d (mod w n).
static rtb_params composedParams(const rtb_ctx* g, int d, uint32_t units)
{
	rtb_params P = g->userParams;
	int n = (int)groupSize(g);
	if (n == 1) return P;
	int w = (P.partition != RTB_PART_NONE && P.part_world > 1) ? P.part_world : 1;
	int r = (w > 1) ? P.part_rank : 0;
	int mode = (P.partition == RTB_PART_TILE) ? RTB_PART_TILE : RTB_PART_SPP;
	if (P.partition == RTB_PART_NONE && units < (uint32_t)n) mode = RTB_PART_TILE;
	P.partition = mode;
	P.part_world = w * n;
	P.part_rank = r + w * d;
	return P;
}

// A device group shards the PASSES of the light-driven estimators (a pass is a unit whose splats land anywhere on
// the film / one VPL set for the whole image): member d takes a contiguous share of [pass_begin, pass_begin + count).
template <class F>
static int splitPasses(rtb_ctx* g, uint32_t pass_begin, uint32_t pass_count, F one)
{
	uint32_t n = (uint32_t)groupSize(g), base = pass_count / n, rem = pass_count % n;
	uint32_t share0 = base + (rem > 0 ? 1u : 0u);
	int rc = forEachMember(g, [&](rtb_ctx* m, int d) {
		uint32_t cnt = base + ((uint32_t)d < rem ? 1u : 0u);
		uint32_t off = (uint32_t)d * base + ((uint32_t)d < rem ? (uint32_t)d : rem);
		if (cnt == 0) return (int)RTB_OK;
		m->accumDirty = true;
		return one(m, pass_begin + off, cnt);
	});
	if (rc) return rc;
	g->spp += pass_count - share0; // Film::SPP counts every pass of the call, not only device 0's share
	return RTB_OK;
}

namespace
{
struct NcclApi
{
	int (*CommInitAll)(void**, int, const int*) = nullptr;
	int (*CommDestroy)(void*) = nullptr;
	int (*Reduce)(const void*, void*, size_t, int, int, int, void*, cudaStream_t) = nullptr;
	int (*GroupStart)() = nullptr;
	int (*GroupEnd)() = nullptr;
	const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
const int kNcclInt64 = 4, kNcclSum = 0; // ncclDataType_t / ncclRedOp_t (stable across NCCL 2.x)
} // namespace

static int ncclInit(rtb_ctx* g)
{
	if (g->ncclComms[0]) return RTB_OK;
	if (!g_nccl.Reduce)
	{
		void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
		if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
		if (!h) return fail(g, RTB_ERR_STATE, "device group: no peer access between the GPUs and libnccl.so.2 cannot be loaded (%s)", dlerror());
		g->ncclLib = h;
		g_nccl.CommInitAll = (int (*)(void**, int, const int*))dlsym(h, "ncclCommInitAll");
		g_nccl.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
		g_nccl.Reduce = (int (*)(const void*, void*, size_t, int, int, int, void*, cudaStream_t))dlsym(h, "ncclReduce");
		g_nccl.GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
		g_nccl.GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
		g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
		if (!g_nccl.CommInitAll || !g_nccl.CommDestroy || !g_nccl.Reduce || !g_nccl.GroupStart || !g_nccl.GroupEnd)
		{
			g_nccl = NcclApi();
			return fail(g, RTB_ERR_STATE, "libnccl.so.2 lacks ncclCommInitAll / ncclReduce");
		}
	}
	int devs[8];
	int n = (int)groupSize(g);
	devs[0] = g->device;
	for (int d = 1; d < n; d++) devs[d] = g->peers[d - 1]->device;
	int rc = g_nccl.CommInitAll(g->ncclComms, n, devs);
	if (rc != 0)
	{
		memset(g->ncclComms, 0, sizeof(g->ncclComms));
		return fail(g, RTB_ERR_CUDA, "ncclCommInitAll: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error");
	}
	return RTB_OK;
}

// Sum the members' accumulators into device 0's (exact), zero the members'.  Leaves the float film resolved when
// the P2P kernel did the work.
static int gatherAccum(rtb_ctx* g)
{
	rtb_ctx* ctx = g; // for CK
	if (int rc = wfSettleAll(g)) return rc;
	size_t n = groupSize(g);
	if (n == 1) return RTB_OK;
	bool any = false;
	for (rtb_ctx* p : g->peers) any = any || p->accumDirty;
	if (!any) return RTB_OK;
	const uint32_t count = g->width * g->height * 3u;
	const size_t bytes = (size_t)count * sizeof(long long);
	// device 0's stream waits for every member's work
	for (rtb_ctx* p : g->peers)
	{
		CK(cudaSetDevice(p->device));
		CK(cudaEventRecord(p->evGroup, p->stream));
	}
	CK(cudaSetDevice(g->device));
	for (rtb_ctx* p : g->peers) CK(cudaStreamWaitEvent(g->stream, p->evGroup, 0));
	bool allDirect = true;
	for (char c : g->peerDirect) allDirect = allDirect && c;
	int mode = g->reduceMode;
	if (mode == 0 && !allDirect) mode = 1;
	g->gathers++;
	if (mode == 0)
	{
		GatherSrc src;
		src.n = 0;
		for (rtb_ctx* p : g->peers)
			if (p->accumDirty) src.p[src.n++] = p->accum;
		unsigned grid = (count + 255u) / 256u;
		k_film_gather<<<grid, 256, 0, g->stream>>>(g->accum, src, count, g->film);
		g->launches++;
		CK(cudaGetLastError());
		g->filmDirty = false; // the kernel wrote the float film too
		g->gatherP2P++;
	}
	else if (mode == 1)
	{
		if (int rc = ncclInit(g)) return rc;
		// the root's in-place reduce runs on its stream; every member enqueues its part on its own stream
		int rc = g_nccl.GroupStart();
		for (size_t d = 0; d < n && rc == 0; d++)
		{
			rtb_ctx* m = d == 0 ? g : g->peers[d - 1];
			cudaSetDevice(m->device);
			rc = g_nccl.Reduce(m->accum, d == 0 ? g->accum : nullptr, count, kNcclInt64, kNcclSum, 0, g->ncclComms[d], m->stream);
		}
		int rc2 = g_nccl.GroupEnd();
		cudaSetDevice(g->device);
		if (rc != 0 || rc2 != 0) return fail(g, RTB_ERR_CUDA, "ncclReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc ? rc : rc2) : "error");
		g->filmDirty = true;
		g->gatherNccl++;
	}
	else
	{
		if (!g->gatherScratch) CK(cudaMalloc((void**)&g->gatherScratch, bytes));
		for (rtb_ctx* p : g->peers)
		{
			if (!p->accumDirty) continue;
			CK(cudaMemcpyPeerAsync(g->gatherScratch, g->device, p->accum, p->device, bytes, g->stream));
			GatherSrc src;
			src.n = 1, src.p[0] = g->gatherScratch;
			k_film_gather<<<(count + 255u) / 256u, 256, 0, g->stream>>>(g->accum, src, count, g->film);
			g->launches++;
		}
		CK(cudaGetLastError());
		g->filmDirty = false;
	}
	// the members' sums now live in device 0's accumulators: zero theirs (after the gather has read them)
	CK(cudaEventRecord(g->evGroup, g->stream));
	for (rtb_ctx* p : g->peers)
	{
		if (!p->accumDirty) continue;
		CK(cudaSetDevice(p->device));
		CK(cudaStreamWaitEvent(p->stream, g->evGroup, 0));
		CK(cudaMemsetAsync(p->accum, 0, bytes, p->stream));
		p->accumDirty = false;
	}
	CK(cudaSetDevice(g->device));
	return RTB_OK;
}

// RTB_UPLOAD_TIMING=1: rtb_upload_scene prints where its time goes (stderr), one line per stage
struct UploadClock
{
	bool on;
	std::chrono::steady_clock::time_point t;
	UploadClock() : on(getenv("RTB_UPLOAD_TIMING") && atoi(getenv("RTB_UPLOAD_TIMING")) != 0), t(std::chrono::steady_clock::now()) {}
	void lap(const char* what)
	{
		if (!on) return;
		auto n = std::chrono::steady_clock::now();
		fprintf(stderr, "[rtb upload] %-48s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
		t = n;
	}
};

// rtb_upload_scene's device-side re-encodings (the host hands over the caller's arrays untouched):
// EXACT tree node pair [bmin, a][bmax, b] -> [bmin, skip link][bmax, leaf code] (rtb_accel::exactNodePair)
__global__ void k_exact_nodes(float4* nodes, const uint32_t* __restrict__ skip, uint32_t n)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const int32_t a = __float_as_int(nodes[2 * (size_t)i].w), b = __float_as_int(nodes[2 * (size_t)i + 1].w);
	uint32_t leaf = 0xFFFFFFFFu;
	if (a < 0) leaf = ((uint32_t)(~a) << 2) | (uint32_t)b;
	nodes[2 * (size_t)i].w = __uint_as_float(skip[i]);
	nodes[2 * (size_t)i + 1].w = __uint_as_float(leaf);
}
// rtb_tri_isect [v0,d][v1,1/A][v2,material][n,A] -> [n,d][v0,1/A][v1,material][v2,A]: the plane test reads one
// 16-byte row, the edge tests the other three
__global__ void k_pack_tris(float4* tris, uint32_t n)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	float4* t = tris + 4 * (size_t)i;
	const float4 r0 = t[0], r1 = t[1], r2 = t[2], r3 = t[3];
	t[0] = make_float4(r3.x, r3.y, r3.z, r0.w);
	t[1] = make_float4(r0.x, r0.y, r0.z, r1.w);
	t[2] = make_float4(r1.x, r1.y, r1.z, r2.w);
	t[3] = make_float4(r2.x, r2.y, r2.z, r3.w);
}

__global__ void k_pad_texels(const float* __restrict__ rgb, float4* out, size_t n)
{
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = make_float4(rgb[3 * (size_t)i], rgb[3 * (size_t)i + 1], rgb[3 * (size_t)i + 2], 0.0f);
}

// Host side of rtb_upload_scene: validate the description and build everything derived from it (skip links, the
// accelerated trees, the env sampling tables) ONCE; a device group uploads the same prepared scene to every member.
struct PreparedScene
{
	const rtb_scene_desc* sc = nullptr;
	std::vector<uint32_t> skip;              // EXACT tree: first node after each node's subtree (k_exact_nodes)
	rtb_accel::FastTree fast;
	bool gpuBuild = false;                   // the FAST tree is built on each device from `leaves` (rtb_gpu_build.cuh)
	std::vector<rtb_accel::RefLeaf> leaves;
	std::vector<float> marginal, cond;
	int envW = 0, envH = 0;
};

static int prepareScene(rtb_ctx* ctx, const rtb_scene_desc* sc, PreparedScene& ps)
{
	if (sc->n_tris && (!sc->tri_isect || !sc->tri_shade)) return fail(ctx, RTB_ERR_ARG, "triangle arrays missing");
	if (sc->n_ref_nodes && !sc->ref_nodes) return fail(ctx, RTB_ERR_ARG, "ref_nodes missing");
	if (sc->n_tris && !sc->n_ref_nodes) return fail(ctx, RTB_ERR_ARG, "triangles without a BVH");
	if (sc->n_tris && (!sc->materials || !sc->n_materials)) return fail(ctx, RTB_ERR_ARG, "materials missing");
	if (!(sc->camera.width >= 1.0f) || !(sc->camera.height >= 1.0f) || sc->camera.width > 65536.0f || sc->camera.height > 65536.0f)
		return fail(ctx, RTB_ERR_ARG, "bad film size %g x %g", sc->camera.width, sc->camera.height);
	// the film kernels (resolve, import, tonemap, Gaussian, adaptive merge) index width*height*3 elements in 32 bits
	if ((uint64_t)sc->camera.width * (uint64_t)sc->camera.height * 3ull >= (1ull << 32))
		return fail(ctx, RTB_ERR_ARG, "film %g x %g too large: width*height*3 must stay below 2^32", sc->camera.width, sc->camera.height);
	for (uint32_t i = 0; i < sc->n_tris; i++)
		if (sc->tri_isect[i].material >= sc->n_materials) return fail(ctx, RTB_ERR_ARG, "triangle %u: material %u out of range", i, sc->tri_isect[i].material);
	for (uint32_t i = 0; i < sc->n_materials; i++)
	{
		const rtb_material& m = sc->materials[i];
		if (m.type > RTB_BSDF_PLASTIC) return fail(ctx, RTB_ERR_ARG, "material %u: unknown bsdf type %u", i, m.type);
		if (m.tex < 0 || (uint32_t)m.tex >= sc->n_textures) return fail(ctx, RTB_ERR_ARG, "material %u: texture %d out of range", i, m.tex);
	}
	if (sc->n_texels && !sc->texels) return fail(ctx, RTB_ERR_ARG, "texels missing");
	if (sc->n_texels > (1ull << 36)) return fail(ctx, RTB_ERR_ARG, "texel pool too large");
	for (uint32_t i = 0; i < sc->n_textures; i++)
	{
		const rtb_texture& t = sc->textures[i];
		if (t.width < 1 || t.height < 1 || (uint64_t)t.offset + (uint64_t)t.width * (uint64_t)t.height > sc->n_texels)
			return fail(ctx, RTB_ERR_ARG, "texture %u out of the texel pool", i);
	}
	for (uint32_t i = 0; i < sc->n_lights; i++)
	{
		const rtb_light& l = sc->lights[i];
		if (l.type > RTB_LIGHT_ENVMAP) return fail(ctx, RTB_ERR_ARG, "light %u: unknown type", i);
		if (l.type == RTB_LIGHT_AREA && l.triangle >= sc->n_tris) return fail(ctx, RTB_ERR_ARG, "light %u: triangle out of range", i);
		if (l.type == RTB_LIGHT_ENVMAP && (l.tex < 0 || (uint32_t)l.tex >= sc->n_textures)) return fail(ctx, RTB_ERR_ARG, "light %u: env texture out of range", i);
	}
	if (sc->background_type == RTB_LIGHT_ENVMAP && (sc->background_tex < 0 || (uint32_t)sc->background_tex >= sc->n_textures))
		return fail(ctx, RTB_ERR_ARG, "background env texture out of range");

	// host-side acceleration data
	std::vector<rtb_accel::RefLeaf>& leaves = ps.leaves;
	const char* err = nullptr;
	UploadClock clk;
	if (!rtb_accel::buildSkipLinks(sc->ref_nodes, sc->n_ref_nodes, sc->n_tris, ps.skip, leaves, &err)) return fail(ctx, RTB_ERR_ARG, "%s", err);
	clk.lap("skip links + leaves");
	rtb_accel::FastTree& fast = ps.fast;
	// The tree can be built on the device instead (a linear BVH over the same leaves: four kernels and a sort, where the
	// host's binned-SAH recursion takes seconds on the largest soups).
	{
		// Opt-in (profiles/r02_gpu_builder.txt): the linear BVH is built in milliseconds (the host SAH tree of 16 M triangles
		// takes 0.9 s of a 1.6 s upload, profiles/r02_upload_timing.txt) but costs 8 % (soups) ... 25 % (coffee) of the
		// render rate, and the metric is the render rate.  RTB_GPU_BUILD=1, or RTB_GPU_BUILD_MIN_LEAVES=n for "from n leaves on".
		size_t minLeaves = (size_t)-1;
		if (const char* e = getenv("RTB_GPU_BUILD_MIN_LEAVES")) minLeaves = (size_t)atoll(e);
		ps.gpuBuild = leaves.size() >= minLeaves;
		if (const char* e = getenv("RTB_GPU_BUILD")) ps.gpuBuild = atoi(e) != 0;
		if (leaves.size() < 2) ps.gpuBuild = false;
	}
	if (!ps.gpuBuild)
	{
		rtb_accel::FastBuilder fb(leaves);
		fb.build(fast);
		clk.lap("binned-SAH tree");
		// Insertion-based re-optimisation of the tree (rtb_accel::FastOptimizer) + a child order for any-hit rays
		// (rtb_accel::orderForAnyHit).  Default: scenes of 32 K ... 400 K leaves get two passes over their 5 % largest
		// nodes and the p / C order (coffee +5 %, bathroom +15 % Msamples/s for 0.05 / 0.2 s of upload); smaller scenes
		// measured -1 % and uniform soups +1 % for seconds of build, so they keep the builder's tree
		// (profiles/r02_tree_opt.txt).  RTB_TREE_OPT=<passes>[:fraction] and RTB_TREE_ORDER=0..4 override; the optimised
		// tree is kept only if it still fits the traversal stacks.
		const size_t nLeaves = leaves.size();
		int passes = (nLeaves >= 32768 && nLeaves <= 400000) ? 2 : 0;
		float fraction = 0.05f;
		if (const char* e = getenv("RTB_TREE_OPT"))
		{
			passes = atoi(e);
			fraction = 1.0f;
			if (const char* c = strchr(e, ':')) fraction = (float)atof(c + 1);
			if (passes > 64) passes = 64;
		}
		int order = passes > 0 ? 3 : 0;
		if (const char* e = getenv("RTB_TREE_ORDER")) order = atoi(e);
		if (passes > 0 && fast.root >= 0 && fast.nodes.size() >= 8)
		{
			rtb_accel::FastOptimizer opt;
			opt.load(fast);
			for (int k = 0; k < passes; k++) opt.pass(fraction);
			rtb_accel::FastTree better;
			if (opt.store(better) + 2 <= RTB_STACK) fast.nodes.swap(better.nodes), fast.root = better.root, fast.maxDepth = better.maxDepth;
			clk.lap("tree optimisation");
		}
		if (order >= 1 && order <= 4 && fast.root >= 0)
		{
			rtb_accel::orderForAnyHit(fast, sc->tri_isect, order);
			clk.lap("child order for any-hit rays");
		}
	}
	// stack need: one pending sibling per level.  (The WIDE / CW / Q16 re-encodings of this tree are built on first use,
	// ensureTraversal: the upload of a 16 M-triangle scene should not pay for trees nobody selected.)
	if (!ps.gpuBuild && fast.maxDepth + 2 > RTB_STACK) return fail(ctx, RTB_ERR_STATE, "accelerated tree too deep (%u levels)", fast.maxDepth);

	for (uint32_t i = 0; i < sc->n_lights; i++)
	{
		if (sc->lights[i].type == RTB_LIGHT_ENVMAP)
		{
			const rtb_texture& t = sc->textures[sc->lights[i].tex];
			rtb_accel::buildEnvTables(sc->texels + (size_t)t.offset * 3, t.width, t.height, ps.marginal, ps.cond);
			ps.envW = t.width, ps.envH = t.height;
			break;
		}
	}
	ps.sc = sc;
	return RTB_OK;
}

static int uploadPrepared(rtb_ctx* ctx, const PreparedScene& ps)
{
	const rtb_scene_desc* sc = ps.sc;
	const rtb_accel::FastTree& fast = ps.fast;
	UploadClock clk;
	if (int rc = bind(ctx)) return rc;
	CK(cudaStreamSynchronize(ctx->stream));
	freeScene(ctx);
	DevScene& S = ctx->S;
	memset(&S, 0, sizeof(S));
	S.cam = sc->camera;
	int rc;
	const rtb_accel::F4* dx = nullptr;
	const rtb_accel::F4* df = nullptr;
	// EXACT tree: the reference nodes go up as they are; the device swaps the child words for (skip link, leaf code)
	{
		const rtb_ref_node* dn = nullptr;
		if ((rc = uploadArray(ctx, sc->ref_nodes, sc->n_ref_nodes, &dn))) return rc;
		if (sc->n_ref_nodes)
		{
			Scratch tmp; // the skip links are only needed by the re-encoding kernel; freed on every path out
			uint32_t* dskip = nullptr;
			CK(tmp.in(ps.skip.data(), (size_t)sc->n_ref_nodes, &dskip, ctx->stream));
			k_exact_nodes<<<(sc->n_ref_nodes + 255) / 256, 256, 0, ctx->stream>>>((float4*)dn, dskip, sc->n_ref_nodes);
			CK(cudaGetLastError());
			CK(cudaStreamSynchronize(ctx->stream));
		}
		dx = (const rtb_accel::F4*)dn;
	}
	clk.lap("reference nodes -> device");
	uint32_t fastDepth = fast.maxDepth;
	bool built = false;
	if (ps.gpuBuild)
	{
		const uint32_t nl = (uint32_t)ps.leaves.size();
		void* p = nullptr;
		CK(cudaMalloc(&p, (size_t)(nl - 1) * 4 * sizeof(float4)));
		ctx->sceneAllocs.push_back(p);
		CK(rtb_gpu_build::build(ps.leaves.data(), nl, sc->ref_nodes[0].bmin, sc->ref_nodes[0].bmax, (float4*)p, &fastDepth, ctx->stream));
		if (fastDepth + 2 <= RTB_STACK)
		{
			df = (const rtb_accel::F4*)p;
			S.n_fnodes = nl - 1;
			S.fast_root = 0;
			built = true;
		}
	}
	if (!built)
	{
		// host tree (the default for scenes it builds in a fraction of a second; also if the linear BVH came out too deep)
		rtb_accel::FastTree local;
		const rtb_accel::FastTree* ft = &fast;
		if (ps.gpuBuild)
		{
			rtb_accel::FastBuilder fb(ps.leaves);
			fb.build(local);
			if (local.maxDepth + 2 > RTB_STACK) return fail(ctx, RTB_ERR_STATE, "accelerated tree too deep (%u levels)", local.maxDepth);
			ft = &local;
		}
		if ((rc = uploadArray(ctx, ft->nodes.data(), ft->nodes.size(), &df))) return rc;
		CK(cudaStreamSynchronize(ctx->stream)); // `local` dies with this block
		S.n_fnodes = (uint32_t)(ft->nodes.size() / 4);
		S.fast_root = ft->root;
		fastDepth = ft->maxDepth;
	}
	S.xnodes = (const float4*)dx;
	S.fnodes = (const float4*)df;
	S.n_xnodes = sc->n_ref_nodes;
	ctx->fastDepth = fastDepth;
	S.wide_root = S.q16_root = S.fast_root < 0 ? S.fast_root : 0; // a leaf root needs no tree; otherwise set by ensureTraversal
	ctx->haveWide = ctx->haveCw = ctx->haveQ16 = false;
	ctx->cwBlocksPerSM[0] = ctx->cwBlocksPerSM[1] = 0; // staging is sized per scene
	const rtb_accel::F4* dti = nullptr;
	const rtb_tri_shade* dts = nullptr;
	clk.lap("accelerated tree -> device");
	// triangle records: copied as they are, re-packed in place on the device (layout: triTest, rtb_dev_scene.cuh)
	{
		const rtb_tri_isect* draw = nullptr;
		if ((rc = uploadArray(ctx, sc->tri_isect, sc->n_tris, &draw))) return rc;
		if (sc->n_tris)
		{
			k_pack_tris<<<(sc->n_tris + 255) / 256, 256, 0, ctx->stream>>>((float4*)draw, sc->n_tris);
			CK(cudaGetLastError());
		}
		dti = (const rtb_accel::F4*)draw;
	}
	if ((rc = uploadArray(ctx, sc->tri_shade, sc->n_tris, &dts))) return rc;
	S.tri = (const float4*)dti;
	S.tsh = (const float4*)dts;
	S.n_tris = sc->n_tris;
	if ((rc = uploadArray(ctx, sc->materials, sc->n_materials, &S.mats))) return rc;
	if ((rc = uploadArray(ctx, sc->textures, sc->n_textures, &S.texs))) return rc;
	if (sc->n_texels)
	{
		// texels: 12-byte RGB from the caller, 16-byte records on the device
		Scratch tmp;
		float* raw = nullptr;
		void* padded = nullptr;
		CK(tmp.in(sc->texels, (size_t)sc->n_texels * 3, &raw, ctx->stream));
		CK(cudaMalloc(&padded, (size_t)sc->n_texels * sizeof(float4)));
		ctx->sceneAllocs.push_back(padded);
		k_pad_texels<<<(unsigned)(((size_t)sc->n_texels + 255) / 256), 256, 0, ctx->stream>>>(raw, (float4*)padded, (size_t)sc->n_texels);
		CK(cudaGetLastError());
		CK(cudaStreamSynchronize(ctx->stream));
		S.texels = (const float4*)padded;
	}
	if ((rc = uploadArray(ctx, sc->lights, sc->n_lights, &S.lights))) return rc;
	S.n_mats = sc->n_materials, S.n_texs = sc->n_textures, S.n_lights = sc->n_lights;
	S.bg_type = sc->background_type;
	memcpy(S.bg_colour, sc->background_colour, sizeof(S.bg_colour));
	S.bg_tex = sc->background_tex;
	S.area_lights_only = sc->n_lights > 0 ? 1u : 0u;
	for (uint32_t i = 0; i < sc->n_lights; i++)
		if (sc->lights[i].type != RTB_LIGHT_AREA) S.area_lights_only = 0u;
	for (int k = 0; k < 3; k++)
	{
		float lo = sc->n_ref_nodes ? sc->ref_nodes[0].bmin[k] : 0.0f, hi = sc->n_ref_nodes ? sc->ref_nodes[0].bmax[k] : 1.0f;
		float ext = hi - lo;
		S.bmin[k] = lo;
		S.bscale[k] = (ext > 0.0f && ext < FLT_MAX) ? 16.0f / ext : 0.0f;
	}
	// env sampling tables when the environment map is a light
	if (!ps.marginal.empty())
	{
		S.env_w = ps.envW, S.env_h = ps.envH;
		if ((rc = uploadArray(ctx, ps.marginal.data(), ps.marginal.size(), &S.env_marginal))) return rc;
		if ((rc = uploadArray(ctx, ps.cond.data(), ps.cond.size(), &S.env_cond))) return rc;
	}
	ctx->width = (uint32_t)sc->camera.width;
	ctx->height = (uint32_t)sc->camera.height;
	size_t npx = (size_t)ctx->width * ctx->height;
	CK(cudaMalloc((void**)&ctx->film, npx * 3 * sizeof(float)));
	CK(cudaMemsetAsync(ctx->film, 0, npx * 3 * sizeof(float), ctx->stream));
	CK(cudaMalloc((void**)&ctx->accum, npx * 3 * sizeof(long long)));
	CK(cudaMemsetAsync(ctx->accum, 0, npx * 3 * sizeof(long long), ctx->stream));
	ctx->filmDirty = false;
	if (const char* e = getenv("RTB_POOL_SLOTS"))
	{
		long v = atol(e);
		if (v >= 1024 && v <= (64l << 20)) ctx->poolSlots = (uint32_t)v & ~31u;
	}
	CK(cudaMemsetAsync(ctx->counters, 0, RTB_COUNTER_STRIPES * 8 * sizeof(unsigned long long), ctx->stream));
	ctx->spp = 0;
	ctx->renderMs = 0.0;
	// the host vectors above die at return: finish the async copies first
	CK(cudaStreamSynchronize(ctx->stream));
	clk.lap("triangles, materials, textures, film -> device");
	ctx->haveScene = true;
	ctx->accumDirty = false;
	ctx->iterHint.clear(); // iteration counts remembered for the previous scene
	return RTB_OK;
}



// The WIDE, CW and Q16 trees are re-encodings of the FAST tree; they are built the first time a call selects them (from
// the FAST nodes already on the device) instead of at every upload.
static int ensureTraversal(rtb_ctx* ctx, int trav)
{
	if (trav != RTB_TRAV_WIDE && trav != RTB_TRAV_CW && trav != RTB_TRAV_Q16) return RTB_OK;
	bool& have = trav == RTB_TRAV_WIDE ? ctx->haveWide : trav == RTB_TRAV_CW ? ctx->haveCw : ctx->haveQ16;
	if (have) return RTB_OK;
	DevScene& S = ctx->S;
	have = true;
	if (S.fast_root < 0 || S.n_fnodes == 0) return RTB_OK; // single-leaf / empty scene: the traversals use the reference tree
	if (int rc = bind(ctx)) return rc;
	rtb_accel::FastTree fast;
	fast.nodes.resize((size_t)S.n_fnodes * 4);
	fast.root = S.fast_root;
	fast.maxDepth = ctx->fastDepth;
	CK(cudaStreamSynchronize(ctx->stream));
	CK(cudaMemcpy(fast.nodes.data(), S.fnodes, fast.nodes.size() * sizeof(rtb_accel::F4), cudaMemcpyDeviceToHost));
	int rc;
	if (trav == RTB_TRAV_WIDE)
	{
		rtb_accel::WideTree wide;
		rtb_accel::WideBuilder wb(fast);
		wb.build(wide);
		if (3 * wide.maxDepth + 6 > RTB_STACK)
		{
			have = false;
			return fail(ctx, RTB_ERR_STATE, "RTB_TRAV_WIDE: tree too deep (%u levels)", wide.maxDepth);
		}
		const rtb_accel::F4* dw = nullptr;
		if ((rc = uploadArray(ctx, wide.nodes.data(), wide.nodes.size(), &dw))) return rc;
		CK(cudaStreamSynchronize(ctx->stream)); // `wide` dies at return
		S.wnodes = (const float4*)dw;
		S.n_wnodes = (uint32_t)(wide.nodes.size() / 8);
		S.wide_root = wide.root;
	}
	else if (trav == RTB_TRAV_CW)
	{
		rtb_accel::CwTree cw;
		rtb_accel::CwBuilder cb(fast);
		cb.build(cw);
		if (cw.valid && cw.maxDepth + 2 <= RTB_CW_STACK) // deeper: RTB_TRAV_CW walks the FAST tree
		{
			const rtb_accel::F4 *dcn = nullptr, *dcl = nullptr;
			if ((rc = uploadArray(ctx, cw.nodes.data(), cw.nodes.size(), &dcn))) return rc;
			if ((rc = uploadArray(ctx, cw.leaves.data(), cw.leaves.size(), &dcl))) return rc;
			CK(cudaStreamSynchronize(ctx->stream));
			S.cwnodes = (const float4*)dcn, S.cwleaves = (const float4*)dcl;
			S.n_cwnodes = (uint32_t)(cw.nodes.size() / 5), S.n_cwleaves = (uint32_t)(cw.leaves.size() / 2);
			S.cw_valid = 1u, S.cw_depth = cw.maxDepth;
		}
	}
	else
	{
		rtb_accel::Q16Tree q16;
		rtb_accel::buildQ16(fast, q16);
		const rtb_accel::F4 *dqn = nullptr, *dql = nullptr;
		if ((rc = uploadArray(ctx, q16.nodes.data(), q16.nodes.size(), &dqn))) return rc;
		if ((rc = uploadArray(ctx, q16.leaves.data(), q16.leaves.size(), &dql))) return rc;
		CK(cudaStreamSynchronize(ctx->stream));
		S.qnodes = (const float4*)dqn, S.qleaves = (const float4*)dql;
		S.n_qnodes = (uint32_t)(q16.nodes.size() / 2);
		S.q16_root = q16.root;
		memcpy(S.qmin, q16.qmin, sizeof(S.qmin));
		memcpy(S.qstep, q16.qstep, sizeof(S.qstep));
	}
	return RTB_OK;
}

extern "C" {

int rtb_abi_version(void) { return RTB_ABI_VERSION; }

void rtb_default_params(rtb_params* p)
{
	if (!p) return;
	memset(p, 0, sizeof(*p));
	p->max_depth = 4;     // MAX_DEPTH, RTBase/Renderer.h:20
	p->epsilon = 1e-4f;   // EPSILON, RTBase/Geometry.h:60
	p->rr_cap = 0.9f;     // RTBase/Renderer.h:353
	p->integrator = RTB_INT_PATH;
	p->sampling = RTB_SAMPLING_STRICT;
	p->traversal = RTB_TRAV_FAST; // profiles/r01_wide_tree.txt: the 4-wide tree measured ~5 % slower
	p->filter = RTB_FILTER_BOX; // RTBase/Renderer.h:50
	p->filter_radius = 2.0f;    // RTBase/Renderer.h:51
	p->filter_alpha = 0.1f;
	p->seed = 1;                // MTRandom(seed = 1), RTBase/Sampling.h:18
	p->partition = RTB_PART_NONE;
	p->part_rank = 0;
	p->part_world = 1;
	p->cull_rel = 1e-5f;
	p->scheduler = RTB_SCHED_WAVEFRONT;
	p->primary_reuse = 1;
}

static int createOne(int device, rtb_ctx** out)
{
	rtb_ctx* ctx = nullptr;
	if (!out) return fail(nullptr, RTB_ERR_ARG, "rtb_create: out is NULL");
	*out = nullptr;
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n == 0)
		return fail(nullptr, RTB_ERR_NODEV, "no CUDA device (%s); librtb200 has no CPU fallback",
		            e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
	if (device < 0 || device >= n) return fail(nullptr, RTB_ERR_NODEV, "device %d out of range [0,%d)", device, n);
	ctx = new (std::nothrow) rtb_ctx();
	if (!ctx) return fail(nullptr, RTB_ERR_OOM, "out of host memory");
	ctx->device = device;
	rtb_default_params(&ctx->params);
	ctx->userParams = ctx->params;
	memset(&ctx->S, 0, sizeof(ctx->S));
	if (cudaSetDevice(device) != cudaSuccess || cudaMalloc((void**)&ctx->counters, RTB_COUNTER_STRIPES * 8 * sizeof(unsigned long long)) != cudaSuccess ||
	    cudaMemset(ctx->counters, 0, RTB_COUNTER_STRIPES * 8 * sizeof(unsigned long long)) != cudaSuccess)
	{
		int rc = fail(nullptr, RTB_ERR_CUDA, "context creation on device %d failed: %s", device,
		              cudaGetErrorString(cudaGetLastError()));
		delete ctx;
		return rc;
	}
	if (cudaEventCreateWithFlags(&ctx->evGroup, cudaEventDisableTiming) != cudaSuccess) ctx->evGroup = nullptr;
	*out = ctx;
	return RTB_OK;
}

int rtb_create(int device, rtb_ctx** out) { return createOne(device, out); }

int rtb_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
	return n;
}

int rtb_create_multi(const int* devices, int n, rtb_ctx** out)
{
	if (!out) return fail(nullptr, RTB_ERR_ARG, "rtb_create_multi: out is NULL");
	*out = nullptr;
	int all[8];
	if (!devices)
	{
		// every visible device (RayTracer::init: numProcs = dwNumberOfProcessors, Renderer.h:52-55)
		int have = rtb_device_count();
		if (n <= 0 || n > have) n = have;
		if (n > 8) n = 8;
		for (int i = 0; i < n; i++) all[i] = i;
		devices = all;
	}
	if (n < 1 || n > 8) return fail(nullptr, n < 1 ? RTB_ERR_NODEV : RTB_ERR_ARG, "rtb_create_multi: %d devices (1..8 supported; no CUDA device means no renderer: there is no CPU fallback)", n);
	for (int i = 0; i < n; i++)
		for (int j = 0; j < i; j++)
			if (devices[i] == devices[j]) return fail(nullptr, RTB_ERR_ARG, "rtb_create_multi: device %d listed twice", devices[i]);
	rtb_ctx* g = nullptr;
	if (int rc = createOne(devices[0], &g)) return rc;
	for (int i = 1; i < n; i++)
	{
		rtb_ctx* p = nullptr;
		int rc = createOne(devices[i], &p);
		if (rc)
		{
			rtb_destroy(g);
			return rc;
		}
		g->peers.push_back(p);
		int can = 0;
		bool direct = false;
		if (cudaDeviceCanAccessPeer(&can, g->device, p->device) == cudaSuccess && can)
		{
			cudaSetDevice(g->device);
			cudaError_t e = cudaDeviceEnablePeerAccess(p->device, 0);
			direct = (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled);
			cudaGetLastError();
		}
		g->peerDirect.push_back(direct ? 1 : 0);
	}
	if (const char* e = getenv("RTB_GROUP_REDUCE"))
	{
		if (!strcmp(e, "nccl")) g->reduceMode = 1;
		else if (!strcmp(e, "staged")) g->reduceMode = 2;
	}
	cudaSetDevice(g->device);
	*out = g;
	return RTB_OK;
}

int rtb_group_size(const rtb_ctx* ctx) { return ctx ? (int)groupSize(ctx) : 0; }

int rtb_group_info(const rtb_ctx* ctx, int* devices, int* p2p, uint64_t* gathers_p2p, uint64_t* gathers_nccl)
{
	if (!ctx) return RTB_ERR_ARG;
	size_t n = groupSize(ctx);
	for (size_t d = 0; d < n; d++)
	{
		if (devices) devices[d] = d == 0 ? ctx->device : ctx->peers[d - 1]->device;
		if (p2p) p2p[d] = d == 0 ? 1 : (int)ctx->peerDirect[d - 1];
	}
	if (gathers_p2p) *gathers_p2p = ctx->gatherP2P;
	if (gathers_nccl) *gathers_nccl = ctx->gatherNccl;
	return RTB_OK;
}

void rtb_destroy(rtb_ctx* ctx)
{
	if (!ctx) return;
	wfSettle(ctx);
	if (ctx->evProbe) cudaEventDestroy(ctx->evProbe);
	for (int d = 0; d < 8; d++)
		if (ctx->ncclComms[d] && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->ncclComms[d]);
	for (rtb_ctx* p : ctx->peers) rtb_destroy(p);
	ctx->peers.clear();
	cudaSetDevice(ctx->device);
	if (ctx->gatherScratch) cudaFree(ctx->gatherScratch);
	if (ctx->evGroup) cudaEventDestroy(ctx->evGroup);
	cudaStreamSynchronize(ctx->stream);
	resolveTimings(ctx);
	for (cudaEvent_t e : ctx->eventPool) cudaEventDestroy(e);
	freeScene(ctx);
	for (int k = 0; k < RTB_MAX_POOLS; k++)
	{
		if (ctx->poolStreams[k]) cudaStreamDestroy(ctx->poolStreams[k]);
		if (ctx->poolDone[k]) cudaEventDestroy(ctx->poolDone[k]);
		if (ctx->shadowStreams[k]) cudaStreamDestroy(ctx->shadowStreams[k]);
		if (ctx->evShaded[k]) cudaEventDestroy(ctx->evShaded[k]);
		if (ctx->evShadowed[k]) cudaEventDestroy(ctx->evShadowed[k]);
	}
	if (ctx->evFork) cudaEventDestroy(ctx->evFork);
	if (ctx->counters) cudaFree(ctx->counters);
	if (ctx->hostProbe) cudaFreeHost(ctx->hostProbe);
	delete ctx;
}

const char* rtb_last_error(const rtb_ctx* ctx) { return ctx ? ctx->error.c_str() : g_createError.c_str(); }

int rtb_set_stream(rtb_ctx* ctx, void* cuda_stream)
{
	if (!ctx) return RTB_ERR_ARG;
	ctx->stream = (cudaStream_t)cuda_stream;
	return RTB_OK;
}

int rtb_synchronize(rtb_ctx* ctx)
{
	if (!ctx) return RTB_ERR_ARG;
	if (int rc = wfSettleAll(ctx)) return rc;
	for (rtb_ctx* p : ctx->peers)
	{
		CK(cudaSetDevice(p->device));
		CK(cudaStreamSynchronize(p->stream));
	}
	if (int rc = bind(ctx)) return rc;
	CK(cudaStreamSynchronize(ctx->stream));
	return RTB_OK;
}

int rtb_set_params(rtb_ctx* ctx, const rtb_params* p)
{
	if (!ctx || !p) return fail(ctx, RTB_ERR_ARG, "rtb_set_params: NULL argument");
	if (p->integrator < RTB_INT_PATH || p->integrator > RTB_INT_PATH_MIS) return fail(ctx, RTB_ERR_ARG, "bad integrator %d", p->integrator);
	if (p->sampling != RTB_SAMPLING_STRICT && p->sampling != RTB_SAMPLING_IMPORTANCE) return fail(ctx, RTB_ERR_ARG, "bad sampling %d", p->sampling);
	if (int rc = checkTrav(ctx, p->traversal)) return rc;
	if (p->filter != RTB_FILTER_BOX && p->filter != RTB_FILTER_GAUSSIAN) return fail(ctx, RTB_ERR_ARG, "bad filter %d", p->filter);
	if (p->partition < RTB_PART_NONE || p->partition > RTB_PART_TILE) return fail(ctx, RTB_ERR_ARG, "bad partition %d", p->partition);
	if (p->partition != RTB_PART_NONE && (p->part_world < 1 || p->part_rank < 0 || p->part_rank >= p->part_world))
		return fail(ctx, RTB_ERR_ARG, "bad partition rank %d of %d", p->part_rank, p->part_world);
	if (p->scheduler != RTB_SCHED_WAVEFRONT && p->scheduler != RTB_SCHED_MEGAKERNEL) return fail(ctx, RTB_ERR_ARG, "bad scheduler %d", p->scheduler);
	if (p->max_depth > 200) return fail(ctx, RTB_ERR_ARG, "max_depth too large");
	if (p->max_depth < 0 || !(p->epsilon >= 0.0f)) return fail(ctx, RTB_ERR_ARG, "bad max_depth/epsilon");
	// a zero-initialised rtb_params (instead of rtb_default_params) would end every path at depth 0 (rr_cap 0) and
	// remove the cull slack that FAST / WIDE hit-ID parity rests on (cull_rel 0): wrong images with RTB_OK
	if (!(p->rr_cap > 0.0f && p->rr_cap <= 1.0f)) return fail(ctx, RTB_ERR_ARG, "rr_cap %g outside (0, 1]: start from rtb_default_params()", p->rr_cap);
	if (!(p->cull_rel >= 1e-7f && p->cull_rel <= 1e-2f))
		return fail(ctx, RTB_ERR_ARG, "cull_rel %g outside [1e-7, 1e-2]: start from rtb_default_params()", p->cull_rel);
	if (!(p->filter_radius >= 0.0f && p->filter_radius <= 16.0f) || !(fabsf(p->filter_alpha) <= 1e6f))
		return fail(ctx, RTB_ERR_ARG, "bad Gaussian filter parameters (radius %g, alpha %g)", p->filter_radius, p->filter_alpha);
	if (p->primary_reuse != 0 && p->primary_reuse != 1) return fail(ctx, RTB_ERR_ARG, "primary_reuse must be 0 or 1");
	if (int rc = wfSettleAll(ctx)) return rc;
	ctx->params = *p;
	ctx->userParams = *p;
	for (rtb_ctx* m : ctx->peers) m->params = *p, m->userParams = *p;
	return RTB_OK;
}

int rtb_get_params(const rtb_ctx* ctx, rtb_params* p)
{
	if (!ctx || !p) return RTB_ERR_ARG;
	*p = ctx->userParams;
	return RTB_OK;
}

int rtb_upload_scene(rtb_ctx* ctx, const rtb_scene_desc* sc)
{
	if (!ctx || !sc) return fail(ctx, RTB_ERR_ARG, "rtb_upload_scene: NULL argument");
	if (int rc = wfSettleAll(ctx)) return rc;
	PreparedScene ps;
	if (int rc = prepareScene(ctx, sc, ps)) return rc;
	return forEachMember(ctx, [&ps](rtb_ctx* m, int) { return uploadPrepared(m, ps); });
}

int rtb_update_camera(rtb_ctx* ctx, const rtb_camera* cam)
{
	if (!ctx || !cam) return fail(ctx, RTB_ERR_ARG, "rtb_update_camera: NULL argument");
	if (!ctx->haveScene) return fail(ctx, RTB_ERR_STATE, "no scene uploaded");
	if ((uint32_t)cam->width != ctx->width || (uint32_t)cam->height != ctx->height)
		return fail(ctx, RTB_ERR_ARG, "camera film size differs from the uploaded scene's");
	if (int rc = wfSettleAll(ctx)) return rc; // iterations enqueued later must not see another camera
	ctx->S.cam = *cam;
	for (rtb_ctx* m : ctx->peers) m->S.cam = *cam;
	return RTB_OK;
}

static int clearOne(rtb_ctx* ctx)
{
	if (!ctx->haveScene) return fail(ctx, RTB_ERR_STATE, "no scene uploaded");
	if (int rc = wfSettle(ctx)) return rc;
	if (int rc = bind(ctx)) return rc;
	ctx->accumDirty = false;
	CK(cudaMemsetAsync(ctx->film, 0, (size_t)ctx->width * ctx->height * 3 * sizeof(float), ctx->stream));
	CK(cudaMemsetAsync(ctx->accum, 0, (size_t)ctx->width * ctx->height * 3 * sizeof(long long), ctx->stream));
	ctx->filmDirty = false;
	CK(cudaMemsetAsync(ctx->counters, 0, RTB_COUNTER_STRIPES * 8 * sizeof(unsigned long long), ctx->stream));
	resolveTimings(ctx);
	ctx->spp = 0;
	ctx->renderMs = 0.0;
	ctx->stageMs[0] = ctx->stageMs[1] = ctx->stageMs[2] = 0.0;
	ctx->timedIterations = 0, ctx->wfIterations = 0, ctx->wfHostSyncs = 0;
	return RTB_OK;
}

int rtb_clear(rtb_ctx* ctx)
{
	if (!ctx) return RTB_ERR_ARG;
	return forEachMember(ctx, [](rtb_ctx* m, int) { return clearOne(m); });
}

int rtb_render(rtb_ctx* ctx, uint32_t spp_begin, uint32_t spp_count)
{
	if (!ctx) return RTB_ERR_ARG;
	if (!ctx->haveScene) return fail(ctx, RTB_ERR_STATE, "rtb_render before rtb_upload_scene");
	if (spp_count == 0) return RTB_OK;
	if ((uint64_t)spp_begin + spp_count > 0xFFFFFFFFull) return fail(ctx, RTB_ERR_ARG, "sample index overflow");
	// every member of a device group renders its slice on its own host thread (the reference spawns its worker
	// threads per render() call too, Renderer.h:842-849); a plain context is a group of one
	int rc = forEachMember(ctx, [ctx, spp_begin, spp_count](rtb_ctx* m, int d) {
		if (int rc = bind(m)) return rc;
		if (int rc = ensureTraversal(m, ctx->userParams.traversal)) return rc;
		m->params = composedParams(ctx, d, spp_count);
		bool mega = m->params.scheduler == RTB_SCHED_MEGAKERNEL;
		int rc = mega ? renderMegakernel(m, spp_begin, spp_count) : renderWavefront(m, spp_begin, spp_count);
		m->params = ctx->userParams;
		m->accumDirty = true;
		return rc;
	});
	if (rc) return rc;
	ctx->spp += spp_count;
	return RTB_OK;
}

// pass_count x RayTracer::lightTracer() (Renderer.h:220-231).
static int lightOne(rtb_ctx* ctx, uint32_t pass_begin, uint32_t pass_count)
{
	if (int rc = wfSettle(ctx)) return rc;
	if (int rc = bind(ctx)) return rc;
	if (int rc = ensureTraversal(ctx, ctx->params.traversal)) return rc;
	rtb_camera_ext ce;
	if (!rtb_camera_derive(&ctx->S.cam, &ce)) return fail(ctx, RTB_ERR_ARG, "camera matrices are singular");
	RenderArgs A;
	A.accum = ctx->accum;
	A.counters = ctx->counters;
	A.spp_begin = pass_begin, A.spp_count = pass_count;
	A.width = ctx->width, A.height = ctx->height;
	A.P = ctx->params;
	if (!ctx->lightGrid)
	{
		cudaDeviceProp prop;
		CK(cudaGetDeviceProperties(&prop, ctx->device));
		ctx->lightGrid = (unsigned)prop.multiProcessorCount * 16u;
	}
	unsigned long long total = (unsigned long long)ctx->width * ctx->height * pass_count;
	unsigned grid = ctx->lightGrid;
	if ((unsigned long long)grid * 128ull > total) grid = (unsigned)((total + 127ull) / 128ull);
	EventPair ev = {getEvent(ctx), getEvent(ctx)};
	cudaEventRecord(ev.a, ctx->stream);
	RTB_TRAV_SWITCH(ctx->params.traversal, k_light_trace<TR><<<grid, 128, 0, ctx->stream>>>(ctx->S, A, ce));
	cudaEventRecord(ev.b, ctx->stream);
	ctx->pending.push_back(ev);
	ctx->launches++;
	CK(cudaGetLastError());
	ctx->filmDirty = true;
	ctx->spp += pass_count; // Film::incrementSPP once per render() (Renderer.h:878)
	return RTB_OK;
}

int rtb_render_light(rtb_ctx* ctx, uint32_t pass_begin, uint32_t pass_count)
{
	if (!ctx) return RTB_ERR_ARG;
	if (!ctx->haveScene) return fail(ctx, RTB_ERR_STATE, "rtb_render_light before rtb_upload_scene");
	if (pass_count == 0) return RTB_OK;
	if ((uint64_t)pass_begin + pass_count > 0xFFFFFFFFull) return fail(ctx, RTB_ERR_ARG, "pass index overflow");
	if (ctx->userParams.partition != RTB_PART_NONE && ctx->userParams.part_world > 1)
		return fail(ctx, RTB_ERR_ARG, "rtb_render_light: light paths land anywhere on the film; shard the passes over devices instead");
	return splitPasses(ctx, pass_begin, pass_count, [](rtb_ctx* m, uint32_t b, uint32_t c) { return lightOne(m, b, c); });
}

// pass_count x RayTracer::instantRadiosity() (Renderer.h:102-123).
static int irOne(rtb_ctx* ctx, uint32_t pass_begin, uint32_t pass_count, uint32_t n_paths)
{
	if (int rc = wfSettle(ctx)) return rc;
	if (int rc = bind(ctx)) return rc;
	if (int rc = ensureTraversal(ctx, ctx->params.traversal)) return rc;
	if (ctx->vplPaths < n_paths)
	{
		CK(cudaStreamSynchronize(ctx->stream));
		if (ctx->vpls) cudaFree(ctx->vpls);
		if (ctx->vplCounts) cudaFree(ctx->vplCounts);
		ctx->vpls = nullptr, ctx->vplCounts = nullptr, ctx->vplPaths = 0;
		CK(cudaMalloc((void**)&ctx->vpls, (size_t)n_paths * RTB_VPL_SEGMENT * sizeof(VplD)));
		CK(cudaMalloc((void**)&ctx->vplCounts, (size_t)n_paths * sizeof(uint32_t)));
		ctx->vplPaths = n_paths;
	}
	RenderArgs A;
	A.accum = ctx->accum;
	A.counters = ctx->counters;
	A.spp_begin = pass_begin, A.spp_count = pass_count;
	A.width = ctx->width, A.height = ctx->height;
	A.P = ctx->params;
	uint32_t warps = ((ctx->width + 7) / 8) * ((ctx->height + 3) / 4);
	EventPair ev = {getEvent(ctx), getEvent(ctx)};
	cudaEventRecord(ev.a, ctx->stream);
	for (uint32_t pass = pass_begin; pass < pass_begin + pass_count; pass++)
	{
		RTB_TRAV_SWITCH(ctx->params.traversal,
		                k_ir_vpls<TR><<<(n_paths + 63) / 64, 64, 0, ctx->stream>>>(ctx->S, A, pass, n_paths, (VplD*)ctx->vpls, ctx->vplCounts));
		RTB_TRAV_SWITCH(ctx->params.traversal,
		                k_ir_gather<TR><<<(warps + 3) / 4, 128, 0, ctx->stream>>>(ctx->S, A, n_paths, (const VplD*)ctx->vpls, ctx->vplCounts));
		ctx->launches += 2;
	}
	cudaEventRecord(ev.b, ctx->stream);
	ctx->pending.push_back(ev);
	CK(cudaGetLastError());
	ctx->filmDirty = true;
	ctx->spp += pass_count; // Film::incrementSPP once per render() (Renderer.h:878)
	return RTB_OK;
}

int rtb_render_ir(rtb_ctx* ctx, uint32_t pass_begin, uint32_t pass_count, uint32_t n_paths)
{
	if (!ctx) return RTB_ERR_ARG;
	if (!ctx->haveScene) return fail(ctx, RTB_ERR_STATE, "rtb_render_ir before rtb_upload_scene");
	if (pass_count == 0) return RTB_OK;
	if (n_paths < 1 || n_paths > 65536) return fail(ctx, RTB_ERR_ARG, "rtb_render_ir: 1 <= n_paths <= 65536 light paths per pass");
	if ((uint64_t)pass_begin + pass_count > 0xFFFFFFFFull) return fail(ctx, RTB_ERR_ARG, "pass index overflow");
	if (ctx->userParams.partition != RTB_PART_NONE && ctx->userParams.part_world > 1)
		return fail(ctx, RTB_ERR_ARG, "rtb_render_ir: shard the passes over devices (a pass is one VPL set for the whole image)");
	return splitPasses(ctx, pass_begin, pass_count, [n_paths](rtb_ctx* m, uint32_t b, uint32_t c) { return irOne(m, b, c, n_paths); });
}

// One member's share of RayTracer::adaptiveRender: member d of n owns the 32x32 tiles t with t % n == d (the tile
// partition of rtb_params); tiles it does not own get no jobs in its plans and add nothing to its film.
static int adaptiveAlloc(rtb_ctx* ctx)
{
	const uint32_t W = ctx->width, H = ctx->height, t32x = (W + 31) / 32, t32y = (H + 31) / 32, nT = t32x * t32y;
	if (ctx->accumScratch) return RTB_OK;
	CK(cudaMalloc((void**)&ctx->accumScratch, (size_t)W * H * 3 * sizeof(long long)));
	CK(cudaMalloc((void**)&ctx->adaptJobBase, (size_t)(nT + 1) * sizeof(unsigned long long)));
	CK(cudaMalloc((void**)&ctx->adaptSamples, (size_t)nT * sizeof(uint32_t)));
	CK(cudaMalloc((void**)&ctx->adaptVariance, (size_t)nT * sizeof(float)));
	std::vector<uint32_t> sub((size_t)nT * 32);
	uint32_t tilesX = (W + 7) / 8, tilesY = (H + 3) / 4;
	for (uint32_t t = 0; t < nT; t++)
		for (uint32_t k = 0; k < 32; k++)
		{
			uint32_t tx = (t % t32x) * 4 + (k & 3u), ty = (t / t32x) * 8 + (k >> 2);
			sub[(size_t)t * 32 + k] = (tx < tilesX && ty < tilesY) ? ((ty << 16) | tx) : 0xFFFFFFFFu;
		}
	CK(cudaMalloc((void**)&ctx->wfTilesAdaptive, sub.size() * sizeof(uint32_t)));
	CK(cudaMemcpy(ctx->wfTilesAdaptive, sub.data(), sub.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
	return RTB_OK;
}

// renders counts[t] samples (sample indices from sampleFirst on) per pixel of every OWNED tile into the scratch sums
static int adaptiveRenderPlan(rtb_ctx* ctx, int d, int n, uint32_t sampleFirst, uint32_t maxCount, const std::vector<uint32_t>& counts)
{
	const uint32_t nT = (uint32_t)counts.size();
	std::vector<unsigned long long> base(nT + 1);
	base[0] = 0;
	for (uint32_t t = 0; t < nT; t++) base[t + 1] = base[t] + (((int)(t % (uint32_t)n) == d) ? 1024ull * counts[t] : 0ull);
	CK(cudaMemsetAsync(ctx->accumScratch, 0, (size_t)ctx->width * ctx->height * 3 * sizeof(long long), ctx->stream));
	CK(cudaMemcpyAsync(ctx->adaptJobBase, base.data(), base.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream)); // `base` is pageable host memory
	AdaptivePlan plan = {base[nT], nT};
	long long* film = ctx->accum;
	ctx->accum = ctx->accumScratch;
	int rc = renderWavefront(ctx, sampleFirst, maxCount, &plan);
	ctx->accum = film;
	return rc;
}

// RayTracer::adaptiveRender (Renderer.h:679-749) on the wavefront schedule; on a device group every member steers and
// samples its own tiles, and the tile variances meet on the host (they go through it anyway: the plan is host arithmetic).
int rtb_render_adaptive(rtb_ctx* ctx, uint32_t init_samples, uint32_t min_samples, uint32_t max_samples, uint32_t* tile_samples,
                        float* tile_variance)
{
	if (!ctx) return RTB_ERR_ARG;
	if (!ctx->haveScene) return fail(ctx, RTB_ERR_STATE, "rtb_render_adaptive before rtb_upload_scene");
	const rtb_params P = ctx->userParams;
	if (init_samples < 1 || min_samples < 1 || max_samples < min_samples || max_samples > (1u << 20))
		return fail(ctx, RTB_ERR_ARG, "rtb_render_adaptive: need 1 <= init, 1 <= min <= max <= 2^20");
	if (P.scheduler != RTB_SCHED_WAVEFRONT) return fail(ctx, RTB_ERR_ARG, "rtb_render_adaptive runs on the wavefront schedule");
	if (P.partition != RTB_PART_NONE && P.part_world > 1)
		return fail(ctx, RTB_ERR_ARG, "rtb_render_adaptive: the tile sample counts come from the whole image; use one context (a device group) per image");
	// Every call draws from its own range of sample indices, like the reference's ever-advancing MTRandom makes
	// successive adaptiveRender() calls independent: call number k (= Film::SPP before the call, one per call)
	// uses [k * (init + max), (k + 1) * (init + max)).
	const uint64_t base64 = (uint64_t)ctx->spp * ((uint64_t)init_samples + max_samples);
	if (base64 + init_samples + max_samples > 0xFFFFFFFFull) return fail(ctx, RTB_ERR_ARG, "rtb_render_adaptive: sample index overflow (clear the film)");
	const uint32_t sampleBase = (uint32_t)base64;
	const uint32_t W = ctx->width, H = ctx->height, t32x = (W + 31) / 32, t32y = (H + 31) / 32, nT = t32x * t32y;
	const int n = (int)groupSize(ctx);
	// ---- phase 1 (adaptiveSampling): init_samples per pixel into the scratch sums, tile variances
	std::vector<std::vector<float>> varOf((size_t)n, std::vector<float>(nT, 0.0f));
	const std::vector<uint32_t> initCounts(nT, init_samples);
	int rc = forEachMember(ctx, [&](rtb_ctx* m, int d) {
		if (int rc = bind(m)) return rc;
		if (int rc = ensureTraversal(m, P.traversal)) return rc;
		if (int rc = adaptiveAlloc(m)) return rc;
		if (int rc = adaptiveRenderPlan(m, d, n, sampleBase, init_samples, initCounts)) return rc;
		rtb_ctx* ctx = m; // for CK
		k_tile_variance<<<nT, 256, 0, m->stream>>>(m->accumScratch, W, H, init_samples, m->adaptVariance);
		m->launches++;
		CK(cudaMemcpyAsync(varOf[(size_t)d].data(), m->adaptVariance, nT * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
		CK(cudaStreamSynchronize(m->stream));
		return (int)RTB_OK;
	});
	if (rc) return rc;
	std::vector<float> var(nT);
	for (uint32_t t = 0; t < nT; t++) var[t] = varOf[(size_t)(t % (uint32_t)n)][t];
	// ---- weights and sample counts exactly like adaptiveRender / sampleTileWithWeight (float arithmetic)
	float total = 0.0f;
	for (float v : var) total += v;
	std::vector<uint32_t> samples(nT);
	for (uint32_t t = 0; t < nT; t++)
	{
		float w = (total > 0.0f) ? var[t] / total : 0.0f;
		w = sqrtf(w);
		int sample = (int)(w * (float)max_samples);
		samples[t] = (uint32_t)((sample > (int)min_samples) ? sample : (int)min_samples);
	}
	// ---- phase 2 (sampleTileWithWeight): fresh samples (indices init_samples...), film += their mean
	rc = forEachMember(ctx, [&](rtb_ctx* m, int d) {
		if (int rc = bind(m)) return rc;
		if (int rc = adaptiveRenderPlan(m, d, n, sampleBase + init_samples, max_samples, samples)) return rc;
		rtb_ctx* ctx = m; // for CK
		CK(cudaMemcpyAsync(m->adaptSamples, samples.data(), nT * sizeof(uint32_t), cudaMemcpyHostToDevice, m->stream));
		k_adaptive_merge<<<(W * H + 255) / 256, 256, 0, m->stream>>>(m->accumScratch, m->accum, W, H, m->adaptSamples);
		m->launches++;
		CK(cudaGetLastError());
		CK(cudaStreamSynchronize(m->stream)); // `samples` is pageable host memory
		m->filmDirty = true;
		m->accumDirty = true;
		return (int)RTB_OK;
	});
	if (rc) return rc;
	if (n > 1) cudaSetDevice(ctx->device);
	ctx->spp += 1; // Film::incrementSPP once per render() (Renderer.h:878)
	if (tile_samples) memcpy(tile_samples, samples.data(), nT * sizeof(uint32_t));
	if (tile_variance) memcpy(tile_variance, var.data(), nT * sizeof(float));
	return RTB_OK;
}

int rtb_read_film(rtb_ctx* ctx, float* rgb_sum, uint32_t* spp)
{
	if (!ctx) return RTB_ERR_ARG;
	if (!ctx->haveScene) return fail(ctx, RTB_ERR_STATE, "no scene uploaded");
	if (int rc = bind(ctx)) return rc;
	if (int rc = gatherAccum(ctx)) return rc;
	if (rgb_sum)
	{
		const float* src = nullptr;
		if (int rc = filteredFilm(ctx, &src)) return rc;
		CK(cudaMemcpyAsync(rgb_sum, src, (size_t)ctx->width * ctx->height * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
	}
	CK(cudaStreamSynchronize(ctx->stream));
	if (spp) *spp = ctx->spp;
	return RTB_OK;
}

// The inverse of rtb_read_film: replaces the film sums (what RayTracer::denoise does to film->film after
// its filter ran, Renderer.h:784-790).
int rtb_write_film(rtb_ctx* ctx, const float* rgb_sum)
{
	if (!ctx || !rgb_sum) return fail(ctx, RTB_ERR_ARG, "rtb_write_film: NULL argument");
	if (!ctx->haveScene) return fail(ctx, RTB_ERR_STATE, "no scene uploaded");
	if (int rc = wfSettleAll(ctx)) return rc;
	if (int rc = bind(ctx)) return rc;
	// a device group: the written film replaces the sum of ALL members' accumulators
	for (rtb_ctx* m : ctx->peers)
	{
		CK(cudaSetDevice(m->device));
		CK(cudaMemsetAsync(m->accum, 0, (size_t)m->width * m->height * 3 * sizeof(long long), m->stream));
		m->accumDirty = false;
	}
	if (int rc = bind(ctx)) return rc;
	uint32_t n = ctx->width * ctx->height * 3;
	CK(cudaMemcpyAsync(ctx->film, rgb_sum, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
	k_film_import<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->film, ctx->accum, n);
	ctx->launches++;
	CK(cudaGetLastError());
	CK(cudaStreamSynchronize(ctx->stream)); // rgb_sum may be pageable
	ctx->filmDirty = true;                  // the float film is re-derived from the fixed-point sums
	return RTB_OK;
}

int rtb_film_device_ptr(rtb_ctx* ctx, void** dptr, uint64_t* n_floats)
{
	if (!ctx || !dptr) return RTB_ERR_ARG;
	if (!ctx->haveScene) return fail(ctx, RTB_ERR_STATE, "no scene uploaded");
	if (int rc = bind(ctx)) return rc;
	if (int rc = gatherAccum(ctx)) return rc;
	if (int rc = resolveFilm(ctx)) return rc;
	*dptr = ctx->film;
	if (n_floats) *n_floats = (uint64_t)ctx->width * ctx->height * 3;
	return RTB_OK;
}

int rtb_accum_device_ptr(rtb_ctx* ctx, void** dptr, uint64_t* n_int64)
{
	if (!ctx || !dptr) return RTB_ERR_ARG;
	if (!ctx->haveScene) return fail(ctx, RTB_ERR_STATE, "no scene uploaded");
	if (int rc = wfSettleAll(ctx)) return rc; // the caller is about to reduce these sums: the render must be complete
	if (!ctx->peers.empty())
	{
		if (int rc = bind(ctx)) return rc;
		if (int rc = gatherAccum(ctx)) return rc; // a group hands out the sum of its members
	}
	*dptr = ctx->accum;
	if (n_int64) *n_int64 = (uint64_t)ctx->width * ctx->height * 3;
	ctx->filmDirty = true; // the caller may reduce into it
	return RTB_OK;
}

int rtb_set_spp(rtb_ctx* ctx, uint32_t spp)
{
	if (!ctx) return RTB_ERR_ARG;
	ctx->spp = spp;
	return RTB_OK;
}

int rtb_film_size(const rtb_ctx* ctx, uint32_t* width, uint32_t* height)
{
	if (!ctx) return RTB_ERR_ARG;
	if (width) *width = ctx->width;
	if (height) *height = ctx->height;
	return RTB_OK;
}

int rtb_tonemap(rtb_ctx* ctx, uint8_t* rgb8, float exposure)
{
	if (!ctx || !rgb8) return fail(ctx, RTB_ERR_ARG, "rtb_tonemap: NULL argument");
	if (!ctx->haveScene) return fail(ctx, RTB_ERR_STATE, "no scene uploaded");
	if (int rc = bind(ctx)) return rc;
	if (int rc = gatherAccum(ctx)) return rc;
	size_t n = (size_t)ctx->width * ctx->height * 3;
	if (!ctx->tone) CK(cudaMalloc((void**)&ctx->tone, n));
	const float* src = nullptr;
	if (int rc = filteredFilm(ctx, &src)) return rc;
	k_tonemap<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(src, ctx->width * ctx->height, (float)ctx->spp, exposure, ctx->tone);
	ctx->launches++;
	CK(cudaGetLastError());
	CK(cudaMemcpyAsync(rgb8, ctx->tone, n, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return RTB_OK;
}

static int statsOne(rtb_ctx* ctx, rtb_stats* out);

int rtb_get_stats(rtb_ctx* ctx, rtb_stats* out)
{
	if (!ctx || !out) return RTB_ERR_ARG;
	if (int rc = statsOne(ctx, out)) return rc;
	// a device group: work counters add up, the render time is the slowest member's
	for (rtb_ctx* m : ctx->peers)
	{
		rtb_stats s;
		if (int rc = statsOne(m, &s))
		{
			ctx->error = m->error;
			return rc;
		}
		out->samples += s.samples, out->closest_rays += s.closest_rays, out->shadow_rays += s.shadow_rays;
		out->kernel_launches += s.kernel_launches;
		out->box_tests += s.box_tests, out->tri_tests += s.tri_tests;
		out->shadow_box_tests += s.shadow_box_tests, out->shadow_tri_tests += s.shadow_tri_tests;
		out->iterations += s.iterations, out->host_syncs += s.host_syncs;
		if (s.render_ms > out->render_ms) out->render_ms = s.render_ms;
	}
	if (!ctx->peers.empty()) cudaSetDevice(ctx->device);
	return RTB_OK;
}

static int statsOne(rtb_ctx* ctx, rtb_stats* out)
{
	if (int rc = wfSettle(ctx)) return rc;
	if (int rc = bind(ctx)) return rc;
	unsigned long long rows[RTB_COUNTER_STRIPES * 8], c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
	CK(cudaMemcpyAsync(rows, ctx->counters, sizeof(rows), cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	for (int r = 0; r < RTB_COUNTER_STRIPES; r++)
		for (int k = 0; k < 8; k++) c[k] += rows[r * 8 + k];
	resolveTimings(ctx);
	memset(out, 0, sizeof(*out));
	out->samples = c[0], out->closest_rays = c[1], out->shadow_rays = c[2];
	out->box_tests = c[3], out->tri_tests = c[4];
	out->shadow_box_tests = c[5], out->shadow_tri_tests = c[6];
	out->kernel_launches = ctx->launches;
	out->render_ms = ctx->renderMs;
	out->extend_ms = ctx->stageMs[0], out->shade_ms = ctx->stageMs[1], out->shadow_ms = ctx->stageMs[2];
	out->timed_iterations = ctx->timedIterations;
	out->iterations = ctx->wfIterations;
	out->host_syncs = ctx->wfHostSyncs;
	return RTB_OK;
}

// --------------------------------------------------------------------------------------
// parity entry points
// --------------------------------------------------------------------------------------
int rtb_primary_hits(rtb_ctx* ctx, int traversal, uint32_t* ids, float* t, rtb_ray* rays)
{
	if (!ctx) return RTB_ERR_ARG;
	if (!ctx->haveScene) return fail(ctx, RTB_ERR_STATE, "no scene uploaded");
	if (int rc = checkTrav(ctx, traversal)) return rc;
	if (int rc = wfSettleAll(ctx)) return rc;
	if (int rc = bind(ctx)) return rc;
	if (int rc = ensureTraversal(ctx, traversal)) return rc;
	size_t n = (size_t)ctx->width * ctx->height;
	Scratch sc;
	uint32_t* dIds;
	float* dT;
	rtb_ray* dR;
	CK(sc.out(ids, n, &dIds));
	CK(sc.out(t, n, &dT));
	CK(sc.out(rays, n, &dR));
	unsigned grid = (unsigned)((n + 127) / 128);
	RTB_TRAV_SWITCH(traversal, k_primary<TR><<<grid, 128, 0, ctx->stream>>>(ctx->S, ctx->params.epsilon, ctx->params.cull_rel, ctx->width, ctx->height, dIds, dT, dR, ctx->counters));
	ctx->launches++;
	CK(cudaGetLastError());
	if (ids) CK(cudaMemcpyAsync(ids, dIds, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
	if (t) CK(cudaMemcpyAsync(t, dT, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
	if (rays) CK(cudaMemcpyAsync(rays, dR, n * sizeof(rtb_ray), cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return RTB_OK;
}

int rtb_trace(rtb_ctx* ctx, int traversal, int any_hit, const rtb_ray* rays, uint64_t n, rtb_hit* hits)
{
	if (!ctx) return RTB_ERR_ARG;
	if (!ctx->haveScene) return fail(ctx, RTB_ERR_STATE, "no scene uploaded");
	if (int rc = checkTrav(ctx, traversal)) return rc;
	if (n == 0) return RTB_OK;
	if (!rays || !hits) return fail(ctx, RTB_ERR_ARG, "rtb_trace: NULL buffer");
	if (int rc = wfSettleAll(ctx)) return rc;
	if (int rc = bind(ctx)) return rc;
	if (int rc = ensureTraversal(ctx, traversal)) return rc;
	Scratch sc;
	rtb_ray* dR;
	rtb_hit* dH;
	CK(sc.in(rays, n, &dR, ctx->stream));
	CK(sc.out(hits, n, &dH));
	unsigned grid = (unsigned)((n + 127) / 128);
	RTB_TRAV_SWITCH(traversal, k_trace<TR><<<grid, 128, 0, ctx->stream>>>(ctx->S, ctx->params.epsilon, ctx->params.cull_rel, any_hit, dR, n, dH, ctx->counters));
	ctx->launches++;
	CK(cudaGetLastError());
	CK(cudaMemcpyAsync(hits, dH, n * sizeof(rtb_hit), cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return RTB_OK;
}

int rtb_visible(rtb_ctx* ctx, int traversal, const float* p1p2, uint64_t n, uint8_t* out)
{
	if (!ctx) return RTB_ERR_ARG;
	if (!ctx->haveScene) return fail(ctx, RTB_ERR_STATE, "no scene uploaded");
	if (int rc = checkTrav(ctx, traversal)) return rc;
	if (n == 0) return RTB_OK;
	if (!p1p2 || !out) return fail(ctx, RTB_ERR_ARG, "rtb_visible: NULL buffer");
	if (int rc = wfSettleAll(ctx)) return rc;
	if (int rc = bind(ctx)) return rc;
	if (int rc = ensureTraversal(ctx, traversal)) return rc;
	Scratch sc;
	float* dP;
	uint8_t* dO;
	CK(sc.in(p1p2, n * 6, &dP, ctx->stream));
	CK(sc.out(out, n, &dO));
	unsigned grid = (unsigned)((n + 127) / 128);
	RTB_TRAV_SWITCH(traversal, k_visible<TR><<<grid, 128, 0, ctx->stream>>>(ctx->S, ctx->params.epsilon, ctx->params.cull_rel, dP, n, dO, ctx->counters));
	ctx->launches++;
	CK(cudaGetLastError());
	CK(cudaMemcpyAsync(out, dO, n, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return RTB_OK;
}

int rtb_shading_data(rtb_ctx* ctx, const rtb_ray* rays, const rtb_hit* hits, uint64_t n, rtb_shading* out)
{
	if (!ctx) return RTB_ERR_ARG;
	if (!ctx->haveScene) return fail(ctx, RTB_ERR_STATE, "no scene uploaded");
	if (n == 0) return RTB_OK;
	if (!rays || !hits || !out) return fail(ctx, RTB_ERR_ARG, "rtb_shading_data: NULL buffer");
	if (int rc = wfSettleAll(ctx)) return rc;
	if (int rc = bind(ctx)) return rc;
	Scratch sc;
	rtb_ray* dR;
	rtb_hit* dH;
	rtb_shading* dO;
	CK(sc.in(rays, n, &dR, ctx->stream));
	CK(sc.in(hits, n, &dH, ctx->stream));
	CK(sc.out(out, n, &dO));
	k_shading<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(ctx->S, dR, dH, n, dO);
	ctx->launches++;
	CK(cudaGetLastError());
	CK(cudaMemcpyAsync(out, dO, n * sizeof(rtb_shading), cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return RTB_OK;
}

int rtb_eval_bsdf(rtb_ctx* ctx, const rtb_shading* sd, const float* wi, const float* u, uint64_t n, float* eval,
                  float* pdf, float* s_wi, float* s_f, float* s_pdf)
{
	if (!ctx) return RTB_ERR_ARG;
	if (!ctx->haveScene) return fail(ctx, RTB_ERR_STATE, "no scene uploaded");
	if (n == 0) return RTB_OK;
	if (!sd || !wi || !u) return fail(ctx, RTB_ERR_ARG, "rtb_eval_bsdf: NULL input");
	for (uint64_t i = 0; i < n; i++)
		if (sd[i].material < 0 || (uint32_t)sd[i].material >= ctx->S.n_mats) return fail(ctx, RTB_ERR_ARG, "rtb_eval_bsdf: record %llu has no material", (unsigned long long)i);
	if (int rc = wfSettleAll(ctx)) return rc;
	if (int rc = bind(ctx)) return rc;
	Scratch sc;
	rtb_shading* dS;
	float *dWi, *dU, *dE, *dP, *dSw, *dSf, *dSp;
	CK(sc.in(sd, n, &dS, ctx->stream));
	CK(sc.in(wi, n * 3, &dWi, ctx->stream));
	CK(sc.in(u, n * 3, &dU, ctx->stream));
	CK(sc.out(eval, n * 3, &dE));
	CK(sc.out(pdf, n, &dP));
	CK(sc.out(s_wi, n * 3, &dSw));
	CK(sc.out(s_f, n * 3, &dSf));
	CK(sc.out(s_pdf, n, &dSp));
	k_eval_bsdf<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(ctx->S, dS, dWi, dU, n, dE, dP, dSw, dSf, dSp);
	ctx->launches++;
	CK(cudaGetLastError());
	if (eval) CK(cudaMemcpyAsync(eval, dE, n * 12, cudaMemcpyDeviceToHost, ctx->stream));
	if (pdf) CK(cudaMemcpyAsync(pdf, dP, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
	if (s_wi) CK(cudaMemcpyAsync(s_wi, dSw, n * 12, cudaMemcpyDeviceToHost, ctx->stream));
	if (s_f) CK(cudaMemcpyAsync(s_f, dSf, n * 12, cudaMemcpyDeviceToHost, ctx->stream));
	if (s_pdf) CK(cudaMemcpyAsync(s_pdf, dSp, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return RTB_OK;
}

int rtb_eval_light(rtb_ctx* ctx, const int32_t* light, const float* wi, const float* u, uint64_t n, float* p_or_wi,
                   float* emitted, float* pdf, float* eval)
{
	if (!ctx) return RTB_ERR_ARG;
	if (!ctx->haveScene) return fail(ctx, RTB_ERR_STATE, "no scene uploaded");
	if (n == 0) return RTB_OK;
	if (!light || !wi || !u) return fail(ctx, RTB_ERR_ARG, "rtb_eval_light: NULL input");
	for (uint64_t i = 0; i < n; i++)
		if (light[i] < 0 || (uint32_t)light[i] >= ctx->S.n_lights) return fail(ctx, RTB_ERR_ARG, "rtb_eval_light: light index %d out of range", light[i]);
	if (int rc = wfSettleAll(ctx)) return rc;
	if (int rc = bind(ctx)) return rc;
	Scratch sc;
	int32_t* dL;
	float *dWi, *dU, *dP, *dE, *dPd, *dEv;
	CK(sc.in(light, n, &dL, ctx->stream));
	CK(sc.in(wi, n * 3, &dWi, ctx->stream));
	CK(sc.in(u, n * 2, &dU, ctx->stream));
	CK(sc.out(p_or_wi, n * 3, &dP));
	CK(sc.out(emitted, n * 3, &dE));
	CK(sc.out(pdf, n, &dPd));
	CK(sc.out(eval, n * 3, &dEv));
	k_eval_light<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(ctx->S, ctx->params, dL, dWi, dU, n, dP, dE, dPd, dEv);
	ctx->launches++;
	CK(cudaGetLastError());
	if (p_or_wi) CK(cudaMemcpyAsync(p_or_wi, dP, n * 12, cudaMemcpyDeviceToHost, ctx->stream));
	if (emitted) CK(cudaMemcpyAsync(emitted, dE, n * 12, cudaMemcpyDeviceToHost, ctx->stream));
	if (pdf) CK(cudaMemcpyAsync(pdf, dPd, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
	if (eval) CK(cudaMemcpyAsync(eval, dEv, n * 12, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return RTB_OK;
}

int rtb_rng_draws(rtb_ctx* ctx, uint32_t pixel, uint32_t sample, uint32_t n, float* out)
{
	if (!ctx) return RTB_ERR_ARG;
	if (n == 0) return RTB_OK;
	if (!out) return fail(ctx, RTB_ERR_ARG, "rtb_rng_draws: NULL buffer");
	if (int rc = bind(ctx)) return rc;
	Scratch sc;
	float* d;
	CK(sc.out(out, n, &d));
	uint32_t blocks = (n + 3) / 4;
	k_rng<<<(blocks + 127) / 128, 128, 0, ctx->stream>>>(ctx->params.seed, pixel, sample, n, d);
	ctx->launches++;
	CK(cudaGetLastError());
	CK(cudaMemcpyAsync(out, d, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return RTB_OK;
}

} // extern "C"
