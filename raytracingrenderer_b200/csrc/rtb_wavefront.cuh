// rtb_wavefront.cuh — the production schedule of the path-tracing loop: wavefront stages over a
// resident pool of path SLOTS with dynamic job assignment.
//
// Why (profiles/r01_v1_megakernel_summary.md, profiles/r01_v2_wavefront_fixedslots_summary.txt):
// one thread running whole paths keeps 7.9 of 32 lanes busy; a pool whose slots are tied to
// pixels drains unevenly (25 % average occupancy) and scans dead slots.  Here
//
//   JOB   = one pixel sample: job j -> (sample ordinal j / Q, pixel q = j % Q) where Q = 32 x the
//           number of 8x4 pixel tiles this rank owns (consecutive jobs = neighbouring pixels of
//           one sample: coherent primary rays).
//   SLOT  = 64 B of path state (ray, hit, throughput, job).  A slot whose path ends takes the
//           next unclaimed job at once, so the pool stays full until the render runs out of jobs.
//
// One ITERATION advances every live slot by one path vertex:
//   k_wf_extend  Scene::traverse (Scene.h:107-130) for every slot's ray.
//   k_wf_shade   pathTrace's body for one vertex (Renderer.h:328-392): miss / emitter /
//                computeDirect's light sample with visibility deferred to a compact queue of
//                shadow rays / Russian roulette / BSDF sample / regeneration.
//   k_wf_shadow  Scene::visible (Scene.h:161-169) for the queued segments; adds the already
//                weighted NEE contribution when unoccluded.  (RTB_INT_PATH_MIS: k_wf_mis instead, which
//                also traces computeDirectMIS's BSDF-strategy probe ray of the same queue record.)
// Variations on the same pool: a per-pixel primary-hit table (k_wf_primary, WF_PREHIT slots, a second
// shade pass per launch), per-tile job plans for RayTracer::adaptiveRender (WfArgs::tileJobBase,
// k_tile_variance, k_adaptive_merge).
// Film::splat with BoxFilter (Imaging.h:139-154, 209-232) is an atomic add into per-pixel
// 64-bit FIXED-POINT sums (2^-32 units): integer addition is associative, so the film is
// bit-reproducible whatever the scheduling, and tile- or spp-partitioned renders compose to
// exactly the single-GPU film.  k_wf_resolve converts the sums to the float film.
#pragma once
#include "rtb_kernels.cuh"
#include "rtb_dev_cw.cuh"

// resident 128-thread blocks per SM the stage kernels are compiled for (register budget)
#ifndef WF_VOTE
#define WF_VOTE 1
#endif
// k_wf_shade can fetch the next round's slot state ahead: 0 off (default), 1 prefetch.global.L2, 2 cp.async into shared memory.
// Measured (profiles/r02_shade_prefetch.txt): no gain in either form - with the stages of two sub-pools overlapping, a
// stalled shade warp is covered by the other kernels; the whole step is bound by issued instructions, not by latency.
#ifndef WF_SHADE_PREFETCH
#define WF_SHADE_PREFETCH 0
#endif
#ifndef WF_SHADE_MIN_BLOCKS
#define WF_SHADE_MIN_BLOCKS 6
#endif
#ifndef WF_EXTEND_MIN_BLOCKS
#define WF_EXTEND_MIN_BLOCKS 8
#endif

#define WF_SHADOW_PER_SLOT 2u /* shadow-queue entries per slot when the shade stage runs several passes */
#define WF_PREHIT 0x400u /* hit[slot] already holds the vertex (primary-hit table): k_wf_extend skips the slot */
#define WF_ALIVE 0x200u
#define WF_CANHIT 0x100u
#define WF_DEPTH_MASK 0xFFu

struct WfCtrl
{
	unsigned int nShadow;    // shadow rays queued by k_wf_shade of this iteration
	unsigned int alive;      // slots alive after k_wf_shade of this iteration
	unsigned int extendHead; // next unclaimed chunk of slots (k_wf_extend)
	unsigned int shadowHead; // next unclaimed chunk of shadow rays (k_wf_shadow)
	unsigned int nExtend;    // rays in the binned extend order of this iteration (k_sort_scan)
	unsigned int pad_[3];
};

struct WfGlobal
{
	unsigned long long nextJob;
};

struct WfArgs
{
	float4* rayO;  // o.xyz, bits(q)            q = pixel ordinal within the owned tile list
	float4* rayD;  // d.xyz, bits(flags | depth)
	float4* hit;   // bits(id), t, alpha, beta
	float4* thr;   // throughput.xyz, bits(sample ordinal)
	float4* shO;   // shadow queue: origin, maxT
	float4* shD;   //               direction, bits(film pixel index)
	float4* shC;   //               contribution if unoccluded, bits(MIS flags)
	float4* misA;  // RTB_INT_PATH_MIS only: shading point x, pdf of the BSDF-strategy direction
	float4* misB;  //                        BSDF-strategy direction, pdf * pmf of the sampled light
	float4* misC;  //                        T * f * max(dot(wi, sN), 0) / pdf_bsdf (what an emitter's Le is multiplied by)
	WfCtrl* ctrl;  // [iterations], zeroed before the render
	WfGlobal* glob;
	const uint32_t* tileList; // owned 8x4-pixel tiles: (tile row << 16) | tile column
	const unsigned long long* tileJobBase; // adaptive plan: first job of every 32x32 tile (+ total), or NULL
	uint32_t nTiles32;                     //                number of 32x32 tiles
	// ray binning (k_sort_*): a permutation of the shadow queue / of the live slots in bucket order, or NULL
	uint32_t* shPerm;
	uint32_t* exPerm;
	uint32_t* sortHist;   // [2][WF_SORT_BUCKETS]: shadow, extend
	uint32_t* sortCursor; // [2][WF_SORT_BUCKETS]
	const float4* primary;    // per-pixel primary hit (bits(id), t, alpha, beta), or NULL: trace every camera ray
	uint32_t primaryPasses;   // fresh primary vertices a shade thread may take on per launch
	unsigned long long* counters;
	long long* accum; // [width*height*3] fixed-point film sums
	uint32_t chunk; // rays a warp of a persistent kernel claims per atomic (WF_CHUNK, or 64 on scenes with long rays)
	uint32_t nSlots, nTiles;
	uint32_t width, height;
	uint32_t sFirst, sStep, sCount; // sample ordinal n -> global sample index sFirst + n * sStep
	unsigned long long totalJobs;
	rtb_params P;
};

__global__ void __launch_bounds__(256) k_wf_resolve(const long long* __restrict__ accum, float* film, uint32_t n)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	film[i] = (float)((double)accum[i] * (1.0 / 4294967296.0));
}

// Film read-out of a device group (rtb_create_multi): device 0 adds the other GPUs' fixed-point sums to its own —
// reading them straight from peer memory over NVLink — and converts the total to the float film in the same pass:
// reduction + finalisation in ONE kernel, no staging buffer, exact (integer sums: the result is the single-GPU film).
struct GatherSrc
{
	const long long* p[8];
	uint32_t n;
};
__global__ void __launch_bounds__(256) k_film_gather(long long* __restrict__ accum, const GatherSrc src, uint32_t n, float* __restrict__ film)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	long long v = accum[i];
	for (uint32_t k = 0; k < src.n; k++) v += src.p[k][i];
	accum[i] = v;
	film[i] = (float)((double)v * (1.0 / 4294967296.0));
}

// the inverse of k_wf_resolve: host-provided float sums become the fixed-point master copy (rtb_write_film)
__global__ void __launch_bounds__(256) k_film_import(const float* __restrict__ film, long long* accum, uint32_t n)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	accum[i] = toFixed(film[i]);
}

// ---------------------------------------------------------------------------------------
// jobs
// ---------------------------------------------------------------------------------------
RTB_DEV bool wfJobPixel(const WfArgs& A, uint32_t q, uint32_t& px, uint32_t& py)
{
	// tile list entry = (tile row << 16) | tile column of an 8x4-pixel tile (no division per vertex)
	uint32_t tile = __ldg(A.tileList + (q >> 5)), lane = q & 31u;
	if (tile == 0xFFFFFFFFu) return false; // adaptive plan: sub-tile outside the image
	px = (tile & 0xFFFFu) * 8u + (lane & 7u);
	py = (tile >> 16) * 4u + (lane >> 3);
	return px < A.width && py < A.height;
}

// job j -> (sample ordinal n, pixel ordinal q); false when q is a padding pixel of an edge tile.
RTB_DEV bool wfDecodeJob(const WfArgs& A, unsigned long long j, uint32_t& n, uint32_t& q)
{
	if (A.tileJobBase)
	{
		// adaptive plan (RayTracer::sampleTileWithWeight, Renderer.h:645-677): every 32x32 tile has
		// its own sample count; its jobs are 1024 pixel slots x count, sample-major
		uint32_t lo = 0, hi = A.nTiles32;
		while (hi - lo > 1u)
		{
			uint32_t mid = (lo + hi) >> 1;
			if (__ldg(A.tileJobBase + mid) <= j) lo = mid;
			else hi = mid;
		}
		unsigned long long l = j - __ldg(A.tileJobBase + lo);
		n = (uint32_t)(l >> 10);
		q = lo * 1024u + (uint32_t)(l & 1023ull);
	}
	else
	{
		uint32_t Q = A.nTiles * 32u;
		n = (uint32_t)(j / Q), q = (uint32_t)(j % Q);
	}
	uint32_t px, py;
	return wfJobPixel(A, q, px, py);
}

// The first vertex of job (n, q) as slot state.  Every sample of a pixel has the SAME camera ray
// (pixel centres only, Renderer.h:806-807, and generateRay draws no random numbers), so with a
// primary-hit table the closest hit is looked up instead of traced.
RTB_DEV void wfFirstVertex(const DevScene& S, const WfArgs& A, uint32_t n, uint32_t q, float4& ro, float4& rd, float4& hh, float4& tq)
{
	uint32_t px, py;
	wfJobPixel(A, q, px, py);
	RayD r = generateRay(S.cam, (float)px + 0.5f, (float)py + 0.5f);
	ro = make_float4(r.o.x, r.o.y, r.o.z, __uint_as_float(q));
	rd = make_float4(r.d.x, r.d.y, r.d.z, __uint_as_float(WF_ALIVE | WF_CANHIT | (A.primary ? WF_PREHIT : 0u)));
	tq = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(n));
	if (A.primary) hh = __ldg(A.primary + (size_t)py * A.width + px);
}

RTB_DEV void wfWriteJob(const DevScene& S, const WfArgs& A, uint32_t slot, uint32_t n, uint32_t q)
{
	float4 ro, rd, hh, tq;
	wfFirstVertex(S, A, n, q, ro, rd, hh, tq);
	A.rayO[slot] = ro;
	A.rayD[slot] = rd;
	A.thr[slot] = tq;
	if (A.primary) A.hit[slot] = hh;
}

// Warp-cooperative: every lane with `need` claims jobs until it holds a valid one or the render
// has no jobs left.  One atomic per warp and round.  Returns true if the lane got a job (n, q).
RTB_DEV bool wfClaimJob(const WfArgs& A, bool need, uint32_t& n, uint32_t& q)
{
	bool got = false;
	for (;;)
	{
		unsigned mask = __ballot_sync(0xFFFFFFFFu, need);
		if (!mask) break;
		int leader = __ffs(mask) - 1;
		unsigned long long base = 0;
		if ((int)(threadIdx.x & 31u) == leader) base = atomicAdd(&A.glob->nextJob, (unsigned long long)__popc(mask));
		base = __shfl_sync(0xFFFFFFFFu, base, leader);
		if (base >= A.totalJobs) break;
		if (need)
		{
			unsigned long long j = base + __popc(mask & ((1u << (threadIdx.x & 31u)) - 1u));
			if (j < A.totalJobs)
			{
				if (wfDecodeJob(A, j, n, q))
				{
					need = false;
					got = true;
				}
			}
			else
				need = false;
		}
	}
	return got;
}


// ---------------------------------------------------------------------------------------
// RayTracer::adaptiveSampling (Renderer.h:583-641): per 32x32 tile, the variance of the pixels'
// initial estimates around the tile mean: ((sum_r + sum_g + sum_b) / 3) / (n - 1).  One block per
// tile; sums in double from the exact fixed-point accumulators.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_tile_variance(const long long* __restrict__ accum, uint32_t width, uint32_t height, uint32_t initSamples,
                                                       float* __restrict__ variance)
{
	__shared__ double sRed[3][8];
	__shared__ double sMean[3];
	uint32_t t32x = (width + 31u) >> 5;
	uint32_t x0 = (blockIdx.x % t32x) * 32u, y0 = (blockIdx.x / t32x) * 32u;
	uint32_t w = min(32u, width - x0), h = min(32u, height - y0), n = w * h;
	const double scale = 1.0 / (4294967296.0 * (double)initSamples);
	double v[4][3];
	double s0 = 0, s1 = 0, s2 = 0;
	for (int k = 0; k < 4; k++)
	{
		uint32_t i = threadIdx.x + k * 256u;
		uint32_t lx = i & 31u, ly = i >> 5;
		bool in = lx < w && ly < h;
		for (int c = 0; c < 3; c++) v[k][c] = in ? (double)accum[((size_t)(y0 + ly) * width + x0 + lx) * 3 + c] * scale : 0.0;
		s0 += v[k][0], s1 += v[k][1], s2 += v[k][2];
	}
	auto blockSum = [&](double a, double b, double c, double* out) {
		for (int o = 16; o > 0; o >>= 1)
		{
			a += __shfl_xor_sync(0xFFFFFFFFu, a, o);
			b += __shfl_xor_sync(0xFFFFFFFFu, b, o);
			c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
		}
		__syncthreads();
		if ((threadIdx.x & 31u) == 0) sRed[0][threadIdx.x >> 5] = a, sRed[1][threadIdx.x >> 5] = b, sRed[2][threadIdx.x >> 5] = c;
		__syncthreads();
		for (int ch = 0; ch < 3; ch++)
		{
			double t = 0;
			for (int k = 0; k < 8; k++) t += sRed[ch][k];
			out[ch] = t;
		}
	};
	double tot[3];
	blockSum(s0, s1, s2, tot);
	if (threadIdx.x == 0)
		for (int c = 0; c < 3; c++) sMean[c] = tot[c] / (double)n;
	__syncthreads();
	double d0 = 0, d1 = 0, d2 = 0;
	for (int k = 0; k < 4; k++)
	{
		uint32_t i = threadIdx.x + k * 256u;
		if ((i & 31u) < w && (i >> 5) < h)
		{
			double a = v[k][0] - sMean[0], b = v[k][1] - sMean[1], c = v[k][2] - sMean[2];
			d0 += a * a, d1 += b * b, d2 += c * c;
		}
	}
	blockSum(d0, d1, d2, tot);
	if (threadIdx.x == 0) variance[blockIdx.x] = (float)(((tot[0] + tot[1] + tot[2]) / 3.0) / (double)(n - 1u));
}

// film += (sum of the tile's `count` samples) / count  (sampleTileWithWeight's col / sample, then splat)
__global__ void __launch_bounds__(256) k_adaptive_merge(const long long* __restrict__ scratch, long long* __restrict__ accum, uint32_t width,
                                                        uint32_t height, const uint32_t* __restrict__ tileSamples)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= width * height) return;
	uint32_t x = i % width, y = i / width, t32x = (width + 31u) >> 5;
	double inv = 1.0 / (double)tileSamples[(y >> 5) * t32x + (x >> 5)];
	for (int c = 0; c < 3; c++) accum[(size_t)i * 3 + c] += __double2ll_rn((double)scratch[(size_t)i * 3 + c] * inv);
}

// The primary-hit table: Scene::traverse of the ONE camera ray of every pixel (8x4 pixels per warp),
// computed at the start of each render call when params.primary_reuse is set.
template <int TRAV>
__global__ void __launch_bounds__(128) k_wf_primary(const __grid_constant__ DevScene S, float4* table, uint32_t width, uint32_t height,
                                                    float epsilon, float cullRel, unsigned long long* counters)
{
	Tally tl = {0, 0, 0, 0, 0, 0, 0};
	uint32_t tilesX = (width + 7u) >> 3, tilesY = (height + 3u) >> 2;
	uint32_t nWarps = tilesX * tilesY;
	for (uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < nWarps; w += (gridDim.x * blockDim.x) >> 5)
	{
		uint32_t l = threadIdx.x & 31u;
		uint32_t px = (w % tilesX) * 8u + (l & 7u), py = (w / tilesX) * 4u + (l >> 3);
		if (px >= width || py >= height) continue;
		RayD r = generateRay(S.cam, (float)px + 0.5f, (float)py + 0.5f);
		HitD h;
		tl.closest++;
		closestHit<TRAV>(S, r, epsilon, cullRel, h, tl.box, tl.tri);
		table[(size_t)py * width + px] = make_float4(__uint_as_float(h.id), h.t, h.alpha, h.beta);
	}
	flushTally(tl, counters);
}

// ---------------------------------------------------------------------------------------
// Ray binning between the stages ("warp-level ray compaction" taken one step further: not only are dead
// lanes squeezed out, the survivors are put next to rays that will walk the same part of the tree).
// profiles/r01_v8_final_summary.md: one thread per queued shadow ray keeps 4.8 of 32 lanes busy, because the
// rays of a warp share nothing — neighbouring slots hold unrelated paths once the pool has been refilled a few
// times.  A counting sort by a 12-bit key (Morton cell of the origin in a 16^3 grid over the scene box for
// scenes lit by a few emitter triangles — origin cell + common target = same way through the tree; 8^3 cells +
// direction octant otherwise) makes a warp's rays start together and head the same way.  Three small kernels:
// block-local shared-memory histograms merged into a global one; a 4096-entry scan; a scatter that re-histograms
// each 4096-ray chunk in shared memory, reserves the chunk's ranges with ONE global atomic per non-empty bucket
// and writes ray indices (the rays themselves stay where they are: 4 B moved per ray, not 48).  The order inside
// a bucket is arbitrary — nothing downstream depends on it (the film is an integer sum).
// ---------------------------------------------------------------------------------------
#define WF_SORT_BUCKETS 4096u
#define WF_SORT_CHUNK 4096u /* rays per block and round of k_sort_scatter: 256 threads x 16 */

RTB_DEV uint32_t wfSpread3(uint32_t x) // 4 bits -> every third bit
{
	x = (x | (x << 8)) & 0x0300F00Fu;
	x = (x | (x << 4)) & 0x030C30C3u;
	x = (x | (x << 2)) & 0x09249249u;
	return x;
}
RTB_DEV uint32_t wfSortKey(const DevScene& S, float ox, float oy, float oz, float dx, float dy, float dz)
{
	int cx = (int)((ox - S.bmin[0]) * S.bscale[0]), cy = (int)((oy - S.bmin[1]) * S.bscale[1]), cz = (int)((oz - S.bmin[2]) * S.bscale[2]);
	cx = min(max(cx, 0), 15), cy = min(max(cy, 0), 15), cz = min(max(cz, 0), 15);
	if (S.area_lights_only) return wfSpread3((uint32_t)cx) | (wfSpread3((uint32_t)cy) << 1) | (wfSpread3((uint32_t)cz) << 2);
	uint32_t oct = (dx < 0.0f ? 1u : 0u) | (dy < 0.0f ? 2u : 0u) | (dz < 0.0f ? 4u : 0u);
	uint32_t m = wfSpread3((uint32_t)cx >> 1) | (wfSpread3((uint32_t)cy >> 1) << 1) | (wfSpread3((uint32_t)cz >> 1) << 2);
	return (oct << 9) | m;
}

// WHAT 0: the shadow queue of iteration `iter`; WHAT 1: the live, not yet intersected slots
template <int WHAT>
RTB_DEV bool wfSortItem(const DevScene& S, const WfArgs& A, uint32_t i, uint32_t& key)
{
	float4 o, d;
	if (WHAT == 0) o = A.shO[i], d = A.shD[i];
	else
	{
		d = A.rayD[i];
		if ((__float_as_uint(d.w) & (WF_ALIVE | WF_PREHIT)) != WF_ALIVE) return false;
		o = A.rayO[i];
	}
	key = wfSortKey(S, o.x, o.y, o.z, d.x, d.y, d.z);
	return true;
}

template <int WHAT>
__global__ void __launch_bounds__(256) k_sort_count(const __grid_constant__ DevScene S, const __grid_constant__ WfArgs A, uint32_t iter)
{
	__shared__ uint32_t h[WF_SORT_BUCKETS];
	if (WHAT == 1 && iter > 0 && A.ctrl[iter - 1].alive == 0) return;
	const uint32_t n = WHAT == 0 ? A.ctrl[iter].nShadow : A.nSlots;
	if (blockIdx.x * WF_SORT_CHUNK >= n) return;
	for (uint32_t k = threadIdx.x; k < WF_SORT_BUCKETS; k += 256u) h[k] = 0u;
	__syncthreads();
	for (uint32_t base = blockIdx.x * WF_SORT_CHUNK; base < n; base += gridDim.x * WF_SORT_CHUNK)
		for (uint32_t j = 0; j < WF_SORT_CHUNK / 256u; j++)
		{
			uint32_t i = base + j * 256u + threadIdx.x, key;
			if (i < n && wfSortItem<WHAT>(S, A, i, key)) atomicAdd(&h[key], 1u);
		}
	__syncthreads();
	uint32_t* hist = A.sortHist + WHAT * WF_SORT_BUCKETS;
	for (uint32_t k = threadIdx.x; k < WF_SORT_BUCKETS; k += 256u)
		if (h[k]) atomicAdd(&hist[k], h[k]);
}

// exclusive scan of the 4096 bucket counts -> cursors; clears the histogram for the next iteration
template <int WHAT>
__global__ void __launch_bounds__(1024) k_sort_scan(const __grid_constant__ WfArgs A, uint32_t iter)
{
	__shared__ uint32_t warpSum[32];
	uint32_t* hist = A.sortHist + WHAT * WF_SORT_BUCKETS;
	uint32_t* cursor = A.sortCursor + WHAT * WF_SORT_BUCKETS;
	uint32_t t = threadIdx.x, v[4], s = 0;
	for (int k = 0; k < 4; k++) v[k] = hist[t * 4 + k], s += v[k];
	uint32_t incl = s;
	for (int o = 1; o < 32; o <<= 1)
	{
		uint32_t x = __shfl_up_sync(0xFFFFFFFFu, incl, o);
		if ((t & 31u) >= (uint32_t)o) incl += x;
	}
	if ((t & 31u) == 31u) warpSum[t >> 5] = incl;
	__syncthreads();
	if (t < 32u)
	{
		uint32_t w = warpSum[t], wi = w;
		for (int o = 1; o < 32; o <<= 1)
		{
			uint32_t x = __shfl_up_sync(0xFFFFFFFFu, wi, o);
			if (t >= (uint32_t)o) wi += x;
		}
		warpSum[t] = wi - w;
		if (WHAT == 1 && t == 31u) A.ctrl[iter].nExtend = wi;
	}
	__syncthreads();
	uint32_t run = warpSum[t >> 5] + incl - s;
	for (int k = 0; k < 4; k++)
	{
		cursor[t * 4 + k] = run;
		run += v[k];
		hist[t * 4 + k] = 0u;
	}
}

template <int WHAT>
__global__ void __launch_bounds__(256) k_sort_scatter(const __grid_constant__ DevScene S, const __grid_constant__ WfArgs A, uint32_t iter)
{
	__shared__ uint32_t h[WF_SORT_BUCKETS];
	if (WHAT == 1 && iter > 0 && A.ctrl[iter - 1].alive == 0) return;
	const uint32_t n = WHAT == 0 ? A.ctrl[iter].nShadow : A.nSlots;
	uint32_t* cursor = A.sortCursor + WHAT * WF_SORT_BUCKETS;
	uint32_t* perm = WHAT == 0 ? A.shPerm : A.exPerm;
	for (uint32_t base = blockIdx.x * WF_SORT_CHUNK; base < n; base += gridDim.x * WF_SORT_CHUNK)
	{
		for (uint32_t k = threadIdx.x; k < WF_SORT_BUCKETS; k += 256u) h[k] = 0u;
		__syncthreads();
		uint32_t kr[WF_SORT_CHUNK / 256u]; // key << 16 | rank inside the chunk's bucket (rank < 4096)
#pragma unroll
		for (uint32_t j = 0; j < WF_SORT_CHUNK / 256u; j++)
		{
			uint32_t i = base + j * 256u + threadIdx.x, key;
			kr[j] = 0xFFFFFFFFu;
			if (i < n && wfSortItem<WHAT>(S, A, i, key)) kr[j] = (key << 16) | atomicAdd(&h[key], 1u);
		}
		__syncthreads();
		for (uint32_t k = threadIdx.x; k < WF_SORT_BUCKETS; k += 256u)
		{
			uint32_t c = h[k];
			if (c) h[k] = atomicAdd(&cursor[k], c);
		}
		__syncthreads();
#pragma unroll
		for (uint32_t j = 0; j < WF_SORT_CHUNK / 256u; j++)
			if (kr[j] != 0xFFFFFFFFu) perm[h[kr[j] >> 16] + (kr[j] & 0xFFFFu)] = base + j * 256u + threadIdx.x;
		__syncthreads();
	}
}

__global__ void __launch_bounds__(256) k_wf_init(const __grid_constant__ DevScene S, const __grid_constant__ WfArgs A)
{
	uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
	bool inRange = slot < A.nSlots;
	uint32_t n = 0, q = 0;
	bool live = wfClaimJob(A, inRange, n, q);
	if (live) wfWriteJob(S, A, slot, n, q);
	else if (inRange) A.rayD[slot] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(0u));
}

// ---------------------------------------------------------------------------------------
// Persistent traversal kernels.  profiles/r01_v3_wavefront_dynjobs_summary.txt: one thread per
// ray keeps 8.7 of 32 lanes busy (a warp lasts as long as its longest ray).  Here every warp
// owns a contiguous range of slots and keeps all lanes fed: each lane holds its CURRENT ray and
// a PREFETCHED next one (loads issued long before they are consumed, no atomics); a lane that
// finishes swaps the prefetched ray in.  Inside a warp, interior-node steps and leaf steps are
// scheduled by majority vote so that at least half of the busy lanes take part in every step.
// The per-ray decisions are those of closestAccel / visibleAccel (rtb_dev_scene.cuh).  (Parking a reached leaf and
// descending on — "speculative traversal" — was measured 3 % slower on every scene: the delayed
// t_best costs more box tests than the fuller leaf steps save, profiles/r01_v8_final_summary.md.)
// ---------------------------------------------------------------------------------------
#ifndef WF_REFILL_IDLE
#define WF_REFILL_IDLE 12 /* refill once this many lanes are idle (6..16 measured within 4 %) */
#endif
#ifndef WF_CHUNK
#define WF_CHUNK 128u /* slots a warp claims per atomic */
#endif
#ifndef WF_SHADE_REGROUP
// k_wf_shade: sort each block's 128 slots by vertex kind (surface hit / miss / dead) before shading, so that a warp runs
// one path.  Measured (profiles/r02_shade_regroup.txt): -1 ... -4 % on every scene — the permuted slot accesses touch
// four times the sectors per load and the two extra barriers cost more than the skipped path saves.  Off.
#define WF_SHADE_REGROUP 0
#endif
#ifndef WF_ISTEPS
#define WF_ISTEPS 3 /* interior steps per majority vote in k_wf_extend (profiles/r02_vote_granularity.txt: 1 -> 3 is +1.5 ... +5 %, 6+ loses again) */
#endif
#ifndef WF_SSTACK
#define WF_SSTACK 12 /* stack entries per thread kept in shared memory by k_wf_extend (8 B x 128 threads each) */
#endif

template <int TRAV>
__global__ void __launch_bounds__(128, WF_EXTEND_MIN_BLOCKS) k_wf_extend(const __grid_constant__ DevScene S, const __grid_constant__ WfArgs A, uint32_t iter)
{
	const rtb_params& P = A.P;
	if (iter > 0 && A.ctrl[iter - 1].alive == 0) return; // pool drained
	Tally tl = {0, 0, 0, 0, 0, 0, 0};
	if (TRAV == RTB_TRAV_EXACT || travRoot<TRAV>(S) < 0)
	{
		// parity path: the reference's own tree, one thread per slot
		for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < A.nSlots; slot += gridDim.x * blockDim.x)
		{
			float4 d = A.rayD[slot];
			if ((__float_as_uint(d.w) & (WF_ALIVE | WF_PREHIT)) != WF_ALIVE) continue;
			float4 o = A.rayO[slot];
			RayD r = mkRay(mk(o), mk(d));
			HitD h;
			tl.closest++;
			closestExact(S, r, P.epsilon, h, tl.box, tl.tri);
			A.hit[slot] = make_float4(__uint_as_float(h.id), h.t, h.alpha, h.beta);
		}
		flushTally(tl, A.counters);
		return;
	}
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t ltMask = (1u << lane) - 1u;
	// binned order (k_sort_*): only the live rays, neighbours start in the same cell and head the same way
	const uint32_t nItems = A.exPerm ? A.ctrl[iter].nExtend : A.nSlots;
	// work is claimed in chunks of WF_CHUNK consecutive slots (one atomic per chunk and warp)
	uint32_t cursor = 0, end = 0;
	bool exhausted = false;
#if WF_SSTACK > 0
	__shared__ float2 sStack[WF_SSTACK * 128];
	SharedStack<WF_SSTACK, 128> stk;
	stk.s = sStack + threadIdx.x;
#else
	LocalStack stk;
#endif
	LaneTrav<false> t;
	t.cur = RTB_TRAV_DONE_;
	t.sp = 0;
	uint32_t slot = 0, preSlot = 0;
	float4 preO = make_float4(0, 0, 0, 0), preD = make_float4(0, 0, 0, 0);
	bool have = false, pre = false;
	for (;;)
	{
		// ---- promote the prefetched ray of every idle lane
		if (!have && pre)
		{
			pre = false;
			if ((__float_as_uint(preD.w) & (WF_ALIVE | WF_PREHIT)) == WF_ALIVE)
			{
				t.r = mkRay(mk(preO), mk(preD));
				travSetBest<false>(t, FLT_MAX, P.cull_rel);
				t.bestId = RTB_MISS_ID, t.bestU = t.bestV = 0.0f;
				t.sp = 0;
				slot = preSlot;
				have = true;
				tl.closest++;
				if (travDegenerate<TRAV>(t.r))
				{
					// 0*inf = NaN rays take the reference's own tree
					HitD h;
					closestExact(S, t.r, P.epsilon, h, tl.box, tl.tri);
					A.hit[slot] = make_float4(__uint_as_float(h.id), h.t, h.alpha, h.beta);
					have = false;
				}
				else
				{
					t.cur = travRoot<TRAV>(S);
					travStart<TRAV, false>(S, t);
				}
			}
		}
		// ---- issue the next prefetches from the warp's chunk
		unsigned want = __ballot_sync(0xFFFFFFFFu, !pre);
		if (want && cursor >= end && !exhausted)
		{
			uint32_t c = 0;
			if (lane == 0) c = atomicAdd(&A.ctrl[iter].extendHead, (unsigned)A.chunk);
			c = __shfl_sync(0xFFFFFFFFu, c, 0);
			if (c >= nItems) exhausted = true;
			else
			{
				cursor = c;
				end = (c + A.chunk < nItems) ? c + A.chunk : nItems;
			}
		}
		if (want && cursor < end)
		{
			uint32_t idx = cursor + __popc(want & ltMask);
			if (!pre && idx < end)
			{
				const uint32_t sl = A.exPerm ? A.exPerm[idx] : idx;
				preO = A.rayO[sl];
				preD = A.rayD[sl];
				preSlot = sl;
				pre = true;
			}
			cursor += __popc(want);
		}
		unsigned busy = __ballot_sync(0xFFFFFFFFu, have);
		if (!busy)
		{
			if (!__ballot_sync(0xFFFFFFFFu, pre) && cursor >= end && exhausted) break;
			continue;
		}
		// ---- traverse until enough lanes are idle to make a refill worthwhile
		for (;;)
		{
#if WF_VOTE
			// majority vote: the step kind (interior / leaf) that more lanes wait for runs next
			unsigned mI = __ballot_sync(0xFFFFFFFFu, have && t.cur >= 0 && t.cur != RTB_TRAV_DONE_);
			unsigned mL = __ballot_sync(0xFFFFFFFFu, have && t.cur < 0);
			bool doInterior = __popc(mI) >= __popc(mL);
#else
			// while-while: interior steps until no lane has an interior node, then the leaves
			bool doInterior = __any_sync(0xFFFFFFFFu, have && t.cur >= 0 && t.cur != RTB_TRAV_DONE_);
#endif
			if (doInterior)
			{
				if (have && t.cur >= 0 && t.cur != RTB_TRAV_DONE_) stepInterior<TRAV, false>(S, t, stk, tl.box);
#if WF_ISTEPS > 1
				// further interior steps without a new vote: the lanes that reached a leaf wait (measured: WF_ISTEPS)
#pragma unroll
				for (int rep = 1; rep < WF_ISTEPS; rep++)
					if (have && t.cur >= 0 && t.cur != RTB_TRAV_DONE_) stepInterior<TRAV, false>(S, t, stk, tl.box);
#endif
			}
			else if (have && t.cur < 0)
			{
				int32_t ref;
				if (travLeafRef<TRAV, false>(S, t, ref, tl.box))
				{
					HitD h;
					h.id = t.bestId, h.t = t.bestT, h.alpha = t.bestU, h.beta = t.bestV;
					leafClosest(S, ref, t.r, P.epsilon, h, tl.tri);
					t.bestId = h.id, t.bestU = h.alpha, t.bestV = h.beta;
					travSetBest<false>(t, h.t, P.cull_rel);
				}
				lanePop<false>(t, stk);
			}
			if (have && t.cur == RTB_TRAV_DONE_)
			{
				A.hit[slot] = make_float4(__uint_as_float(t.bestId), t.bestT, t.bestU, t.bestV);
				have = false;
			}
			unsigned idle = __ballot_sync(0xFFFFFFFFu, !have);
			if (idle == 0xFFFFFFFFu) break;
			if (__popc(idle) >= WF_REFILL_IDLE && (__ballot_sync(0xFFFFFFFFu, pre) & idle)) break;
		}
	}
	flushTally(tl, A.counters);
}



// ---------------------------------------------------------------------------------------
// RTB_TRAV_CW: persistent traversal of the 8-wide compressed tree for BOTH ray kinds — the closest-hit rays of the
// slots (ANYHIT = false: Scene::traverse) and the queued shadow rays (ANYHIT = true: Scene::visible).  Same schedule
// as k_wf_extend (warps claim chunks, every lane keeps a prefetched next ray, majority vote between node steps and
// leaf steps), different per-ray machine (rtb_dev_cw.cuh).  At the start every block stages the top of the tree — the
// first `stageNodes` nodes of the breadth-first array, and the exact leaf boxes when they all fit — into its shared
// memory with one bulk copy (cp.async.bulk -> mbarrier): the levels every ray passes through are then read with
// LDS instead of competing for the L1 data pipe, the measured bound of the binary kernels on the heavy scenes.
// ---------------------------------------------------------------------------------------
#ifndef WF_CW_THREADS
#define WF_CW_THREADS 128
#endif
#ifndef WF_CW_MIN_BLOCKS
#define WF_CW_MIN_BLOCKS 5
#endif

template <bool ANYHIT>
__global__ void __launch_bounds__(WF_CW_THREADS, WF_CW_MIN_BLOCKS) k_wf_trace_cw(const __grid_constant__ DevScene S, const __grid_constant__ WfArgs A, uint32_t iter,
                                                                               uint32_t stageNodes, uint32_t stageLeaves)
{
	extern __shared__ float4 sStage[];
	__shared__ unsigned long long sBar;
	const rtb_params& P = A.P;
	if (!ANYHIT && iter > 0 && A.ctrl[iter - 1].alive == 0) return; // pool drained
	const uint32_t nItems = ANYHIT ? A.ctrl[iter].nShadow : (A.exPerm ? A.ctrl[iter].nExtend : A.nSlots);
	if (nItems == 0) return;
	CwView V = cwGlobalView(S);
	if (stageNodes | stageLeaves)
	{
		cwStageBulk(sStage, S.cwnodes, stageNodes * 80u, sStage + (size_t)stageNodes * 5, S.cwleaves, stageLeaves * 32u, &sBar);
		V.sNodes = sStage, V.nShared = stageNodes;
		if (stageLeaves) V.leaves = sStage + (size_t)stageNodes * 5;
	}
	Tally tl = {0, 0, 0, 0, 0, 0, 0};
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t ltMask = (1u << lane) - 1u;
	unsigned int* head = ANYHIT ? &A.ctrl[iter].shadowHead : &A.ctrl[iter].extendHead;
	const uint32_t* perm = ANYHIT ? A.shPerm : A.exPerm;
	const float4* srcO = ANYHIT ? A.shO : A.rayO;
	const float4* srcD = ANYHIT ? A.shD : A.rayD;
	uint32_t cursor = 0, end = 0;
	bool exhausted = false;
	uint2 stack[RTB_CW_STACK];
	LaneCw<ANYHIT> t;
	t.sp = 0;
	t.ng = t.lg = make_uint2(0u, 0u);
	uint32_t item = 0, preItem = 0;
	float4 preO = make_float4(0, 0, 0, 0), preD = make_float4(0, 0, 0, 0);
	uint32_t curW = 0; // any hit: film pixel of the current ray
	bool have = false, pre = false;
	for (;;)
	{
		// ---- promote the prefetched ray of every idle lane
		if (!have && pre)
		{
			pre = false;
			if (ANYHIT || (__float_as_uint(preD.w) & (WF_ALIVE | WF_PREHIT)) == WF_ALIVE)
			{
				RayD r = mkRay(mk(preO), mk(preD));
				item = preItem;
				if (ANYHIT) tl.shadow++;
				else tl.closest++;
				if (cwRayDegenerate(r))
				{
					// 0 * inf = NaN and axis-parallel rays take the reference's own tree
					if (ANYHIT)
					{
						if (visibleExact(S, r, P.epsilon, preO.w, tl.sbox, tl.stri)) filmAdd(A.accum, __float_as_uint(preD.w), mk(A.shC[item]));
					}
					else
					{
						HitD h;
						closestExact(S, r, P.epsilon, h, tl.box, tl.tri);
						A.hit[item] = make_float4(__uint_as_float(h.id), h.t, h.alpha, h.beta);
					}
				}
				else
				{
					cwStart<ANYHIT>(t, r, ANYHIT ? preO.w : FLT_MAX, P.cull_rel);
					curW = __float_as_uint(preD.w);
					have = true;
				}
			}
		}
		// ---- issue the next prefetches from the warp's chunk
		unsigned want = __ballot_sync(0xFFFFFFFFu, !pre);
		if (want && cursor >= end && !exhausted)
		{
			uint32_t c = 0;
			if (lane == 0) c = atomicAdd(head, (unsigned)A.chunk);
			c = __shfl_sync(0xFFFFFFFFu, c, 0);
			if (c >= nItems) exhausted = true;
			else
			{
				cursor = c;
				end = (c + A.chunk < nItems) ? c + A.chunk : nItems;
			}
		}
		if (want && cursor < end)
		{
			uint32_t idx = cursor + __popc(want & ltMask);
			if (!pre && idx < end)
			{
				const uint32_t it = perm ? perm[idx] : idx;
				preO = srcO[it];
				preD = srcD[it];
				preItem = it;
				pre = true;
			}
			cursor += __popc(want);
		}
		unsigned busy = __ballot_sync(0xFFFFFFFFu, have);
		if (!busy)
		{
			if (!__ballot_sync(0xFFFFFFFFu, pre) && cursor >= end && exhausted) break;
			continue;
		}
		// ---- traverse until enough lanes are idle to make a refill worthwhile
		for (;;)
		{
			const bool wantLeaf = have && (t.lg.y & 0xFFu) != 0u;
			const bool wantNode = have && !wantLeaf && (t.ng.y & 0xFF000000u) != 0u;
			unsigned mN = __ballot_sync(0xFFFFFFFFu, wantNode);
			unsigned mL = __ballot_sync(0xFFFFFFFFu, wantLeaf);
			bool occluded = false;
			if (__popc(mN) >= __popc(mL))
			{
				if (wantNode) cwNodeStep<ANYHIT>(V, t, stack, ANYHIT ? tl.sbox : tl.box);
			}
			else if (wantLeaf)
				occluded = cwLeafStep<ANYHIT>(S, V, t, P.epsilon, P.cull_rel, ANYHIT ? tl.sbox : tl.box, ANYHIT ? tl.stri : tl.tri);
			if (have)
			{
				cwPop<ANYHIT>(t, stack);
				if (occluded || cwIdle<ANYHIT>(t))
				{
					if (ANYHIT)
					{
						if (!occluded) filmAdd(A.accum, curW, mk(A.shC[item]));
					}
					else
						A.hit[item] = make_float4(__uint_as_float(t.bestId), t.bestT, t.bestU, t.bestV);
					have = false;
					t.sp = 0;
					t.ng = t.lg = make_uint2(0u, 0u);
				}
			}
			unsigned idle = __ballot_sync(0xFFFFFFFFu, !have);
			if (idle == 0xFFFFFFFFu) break;
			if (__popc(idle) >= WF_REFILL_IDLE && (__ballot_sync(0xFFFFFFFFu, pre) & idle)) break;
		}
	}
	flushTally(tl, A.counters);
}

// The same schedule for the queued shadow rays (Scene::visible, Scene.h:161-169): persistent warps, per-lane prefetch,
// majority vote, shared-memory short stack.  Round 1 measured a persistent any-hit kernel slower than one thread per
// ray on every scene (profiles/r01_shadow_stage.txt: short rays, the refill bookkeeping outweighs the regained
// lanes); on the heavy scenes the one-thread kernel sits at 4.6 of 32 lanes and 68 % of the issue slots
// (profiles/r02_base_bathroom_summary.md), so with the cheaper stack and three interior steps per vote it is selectable
// per scene (rtb_api.cu: shadowPersistent) — the numbers decide (profiles/r02_persistent_shadow.txt).
template <int TRAV>
__global__ void __launch_bounds__(128, WF_EXTEND_MIN_BLOCKS) k_wf_shadow_persist(const __grid_constant__ DevScene S, const __grid_constant__ WfArgs A, uint32_t iter)
{
	const rtb_params& P = A.P;
	const uint32_t nItems = A.ctrl[iter].nShadow;
	if (nItems == 0) return;
	Tally tl = {0, 0, 0, 0, 0, 0, 0};
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t ltMask = (1u << lane) - 1u;
	uint32_t cursor = 0, end = 0;
	bool exhausted = false;
#if WF_SSTACK > 0
	__shared__ float2 sStack[WF_SSTACK * 128];
	SharedStack<WF_SSTACK, 128> stk;
	stk.s = sStack + threadIdx.x;
#else
	LocalStack stk;
#endif
	LaneTrav<true> t;
	t.cur = RTB_TRAV_DONE_;
	t.sp = 0;
	uint32_t item = 0, preItem = 0, pixel = 0;
	float4 preO = make_float4(0, 0, 0, 0), preD = make_float4(0, 0, 0, 0);
	bool have = false, pre = false;
	for (;;)
	{
		if (!have && pre)
		{
			pre = false;
			t.r = mkRay(mk(preO), mk(preD));
			travSetBest<true>(t, preO.w, P.cull_rel);
			t.sp = 0;
			item = preItem, pixel = __float_as_uint(preD.w);
			tl.shadow++;
			if (travDegenerate<TRAV>(t.r))
			{
				if (visibleExact(S, t.r, P.epsilon, preO.w, tl.sbox, tl.stri)) filmAdd(A.accum, pixel, mk(A.shC[item]));
			}
			else
			{
				have = true;
				t.cur = travRoot<TRAV>(S);
				travStart<TRAV, true>(S, t);
			}
		}
		unsigned want = __ballot_sync(0xFFFFFFFFu, !pre);
		if (want && cursor >= end && !exhausted)
		{
			uint32_t c = 0;
			if (lane == 0) c = atomicAdd(&A.ctrl[iter].shadowHead, (unsigned)A.chunk);
			c = __shfl_sync(0xFFFFFFFFu, c, 0);
			if (c >= nItems) exhausted = true;
			else
			{
				cursor = c;
				end = (c + A.chunk < nItems) ? c + A.chunk : nItems;
			}
		}
		if (want && cursor < end)
		{
			uint32_t idx = cursor + __popc(want & ltMask);
			if (!pre && idx < end)
			{
				const uint32_t it = A.shPerm ? A.shPerm[idx] : idx;
				preO = A.shO[it];
				preD = A.shD[it];
				preItem = it;
				pre = true;
			}
			cursor += __popc(want);
		}
		unsigned busy = __ballot_sync(0xFFFFFFFFu, have);
		if (!busy)
		{
			if (!__ballot_sync(0xFFFFFFFFu, pre) && cursor >= end && exhausted) break;
			continue;
		}
		for (;;)
		{
			unsigned mI = __ballot_sync(0xFFFFFFFFu, have && t.cur >= 0 && t.cur != RTB_TRAV_DONE_);
			unsigned mL = __ballot_sync(0xFFFFFFFFu, have && t.cur < 0);
			bool occluded = false;
			if (__popc(mI) >= __popc(mL))
			{
#pragma unroll
				for (int rep = 0; rep < WF_ISTEPS; rep++)
					if (have && t.cur >= 0 && t.cur != RTB_TRAV_DONE_) stepInterior<TRAV, true>(S, t, stk, tl.sbox);
			}
			else if (have && t.cur < 0)
			{
				int32_t ref;
				if (travLeafRef<TRAV, true>(S, t, ref, tl.sbox) && leafOccludes(S, ref, t.r, P.epsilon, t.bestT, tl.stri)) occluded = true;
				else lanePop<true>(t, stk);
			}
			if (have && (occluded || t.cur == RTB_TRAV_DONE_))
			{
				if (!occluded) filmAdd(A.accum, pixel, mk(A.shC[item]));
				have = false;
				t.cur = RTB_TRAV_DONE_;
			}
			unsigned idle = __ballot_sync(0xFFFFFFFFu, !have);
			if (idle == 0xFFFFFFFFu) break;
			if (__popc(idle) >= WF_REFILL_IDLE && (__ballot_sync(0xFFFFFFFFu, pre) & idle)) break;
		}
	}
	flushTally(tl, A.counters);
}

// One thread per slot (the v3 extend stage), kept selectable for A/B measurements against the
// persistent kernel above (RTB_SIMPLE_EXTEND=1).
template <int TRAV>
__global__ void __launch_bounds__(128) k_wf_extend_simple(const __grid_constant__ DevScene S, const __grid_constant__ WfArgs A, uint32_t iter)
{
	const rtb_params& P = A.P;
	if (iter > 0 && A.ctrl[iter - 1].alive == 0) return;
	Tally tl = {0, 0, 0, 0, 0, 0, 0};
	for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < A.nSlots; slot += gridDim.x * blockDim.x)
	{
		float4 d = A.rayD[slot];
		if ((__float_as_uint(d.w) & (WF_ALIVE | WF_PREHIT)) != WF_ALIVE) continue;
		float4 o = A.rayO[slot];
		RayD r = mkRay(mk(o), mk(d));
		HitD h;
		tl.closest++;
		closestHit<TRAV>(S, r, P.epsilon, P.cull_rel, h, tl.box, tl.tri);
		A.hit[slot] = make_float4(__uint_as_float(h.id), h.t, h.alpha, h.beta);
	}
	flushTally(tl, A.counters);
}

// Shadow rays: one thread per queued ray.  (A persistent/prefetching variant like k_wf_extend was
// measured slower on EVERY scene, twice — profiles/r01_v4_persistent_traversal.txt and
// profiles/r01_shadow_stage.txt: materialball -12 %, coffee -22 %, cornell-box -17 %, bathroom -2 %,
// soups -6 %: any-hit rays end early, so the refill bookkeeping outweighs the regained lanes, and a
// kernel that owns every resident block cannot share the SMs with the other streams.  Round 2 tried the lightest form -
// this kernel with a per-lane grid-stride refill at leaf boundaries, no votes, chunks or shared memory: cornell-box -10 %,
// materialball -5 %, coffee -18 %, bathroom / soup 0.  The refill's scattered 48-byte reads and set-up run at a few lanes
// each time and cost more than the lanes they free.)
template <int TRAV>
__global__ void __launch_bounds__(128) k_wf_shadow(const __grid_constant__ DevScene S, const __grid_constant__ WfArgs A, uint32_t iter)
{
	const rtb_params& P = A.P;
	const uint32_t n = A.ctrl[iter].nShadow;
	Tally tl = {0, 0, 0, 0, 0, 0, 0};
	for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x)
	{
		const uint32_t i = A.shPerm ? A.shPerm[q] : q; // binned order (k_sort_*): a warp's rays start together and head the same way
		float4 o = A.shO[i], d = A.shD[i];
		RayD r = mkRay(mk(o), mk(d));
		float maxT = o.w;
		tl.shadow++;
		bool vis = anyVisible<TRAV>(S, r, P.epsilon, maxT, P.cull_rel, tl.sbox, tl.stri);
		if (vis) filmAdd(A.accum, __float_as_uint(d.w), mk(A.shC[i]));
	}
	flushTally(tl, A.counters);
}


// RTB_INT_PATH_MIS: resolves one queue record of computeDirectMIS (Renderer.h:474-557) per thread — the
// light strategy's visibility test, then (unless a visible non-area light ended the estimator, :516-527)
// the BSDF strategy's probe ray: closest hit, and if that is an emitter its radiance with the balance weight
// against the sampled light's pdf converted to solid angle at the hit (:535-552).
template <int TRAV>
__global__ void __launch_bounds__(128) k_wf_mis(const __grid_constant__ DevScene S, const __grid_constant__ WfArgs A, uint32_t iter)
{
	const rtb_params& P = A.P;
	const uint32_t n = A.ctrl[iter].nShadow;
	Tally tl = {0, 0, 0, 0, 0, 0, 0};
	for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x)
	{
		const uint32_t i = A.shPerm ? A.shPerm[q] : q;
		float4 o = A.shO[i], d = A.shD[i], c = A.shC[i];
		uint32_t flags = __float_as_uint(c.w), pixel = __float_as_uint(d.w);
		if (flags & 1u)
		{
			RayD r = mkRay(mk(o), mk(d));
			tl.shadow++;
			if (anyVisible<TRAV>(S, r, P.epsilon, o.w, P.cull_rel, tl.sbox, tl.stri))
			{
				filmAdd(A.accum, pixel, mk(c));
				if (flags & 2u) continue;
			}
		}
		float4 xa = A.misA[i], wb = A.misB[i];
		V3 x = mk(xa), wiB = mk(wb);
		RayD pr = mkRay(x + (wiB * P.epsilon), wiB);
		HitD h;
		closestHit<TRAV>(S, pr, P.epsilon, P.cull_rel, h, tl.box, tl.tri);
		tl.closest++;
		if (h.id == RTB_MISS_ID) continue;
		ShadeD sh;
		calcShading(S, h.id, h.t, h.alpha, h.beta, 1.0f - (h.alpha + h.beta), pr, sh);
		rtb_material mh = S.mats[sh.mat];
		if (!(mh.flags & RTB_MAT_LIGHT)) continue;
		V3 wi = sh.x - x;
		float dist2 = lengthSq(wi);
		wi = normalize(wi);
		float cl = selMax(0.0f, dot(mk(-wi.x, -wi.y, -wi.z), sh.sN));
		float pdfL = (cl > 0.0f) ? (wb.w * dist2 / cl) : 0.0f;
		float wgt = xa.w / (xa.w + pdfL);
		filmAdd(A.accum, pixel, (mk(A.misC[i]) * mk(mh.emission)) * wgt);
	}
	flushTally(tl, A.counters);
}

// k_wf_shade's software prefetch of one slot's state (rayD, rayO, hit, thr: 4 x 16 bytes) into the thread's staging row
__device__ __forceinline__ void shadePrefetchSlot(const WfArgs& A, uint32_t slot, float4 (*row)[128])
{
	if (slot < A.nSlots)
	{
		const float4* src[4] = {A.rayD + slot, A.rayO + slot, A.hit + slot, A.thr + slot};
#pragma unroll
		for (int k = 0; k < 4; k++)
		{
			uint32_t dst = (uint32_t)__cvta_generic_to_shared(&row[k][threadIdx.x]);
			asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src[k]) : "memory");
		}
	}
	asm volatile("cp.async.commit_group;" ::: "memory");
}

template <int INTEGRATOR, bool REUSE>
__global__ void __launch_bounds__(128, WF_SHADE_MIN_BLOCKS) k_wf_shade(const __grid_constant__ DevScene S, const __grid_constant__ WfArgs A, uint32_t iter)
{
	const rtb_params& P = A.P;
	if (iter > 0 && A.ctrl[iter - 1].alive == 0) return; // pool drained
	const uint32_t lane = threadIdx.x & 31u;
	uint32_t nAlive = 0, nDone = 0, phase = 0;
	constexpr int NREC = (INTEGRATOR == RTB_INT_PATH_MIS) ? 6 : 3;
	__shared__ float4 sStage[NREC][128]; // the thread's queue record waits here across the reservation barriers (12 / 24 registers)
	__shared__ uint32_t sCountS[2][4], sCountJ[2][4];
	__shared__ unsigned int sBaseS[2];
	__shared__ unsigned long long sBaseJ[2];
	__shared__ uint32_t sBaseN[2], sBaseQ[2]; // sBaseJ split into (sample ordinal, pixel ordinal): one 64-bit division per block
	// whole blocks stride together: the cooperative queue/job operations below need every lane
	uint32_t nRounds = (A.nSlots + gridDim.x * blockDim.x - 1) / (gridDim.x * blockDim.x);
#if WF_SHADE_REGROUP
	__shared__ uint8_t sPerm[128];
	__shared__ uint32_t sClsCnt[2][4];
#endif
#if WF_SHADE_PREFETCH
	// Most slots hold a live vertex in the steady state: fetch the NEXT round's slot state while this round is shaded.
	// Near the drain (few live slots) the plain path reads 16 bytes per dead slot instead of 64.
	const bool pre = iter > 0 && A.ctrl[iter - 1].alive > (A.nSlots >> 1);
#endif
#if WF_SHADE_PREFETCH == 2
	__shared__ float4 sPre[2][4][128];
	if (pre) shadePrefetchSlot(A, blockIdx.x * blockDim.x + threadIdx.x, sPre[0]);
#endif
	for (uint32_t round = 0; round < nRounds; round++)
	{
		uint32_t slot = round * gridDim.x * blockDim.x + blockIdx.x * blockDim.x + threadIdx.x;
#if WF_SHADE_REGROUP
		// Regroup the block's 128 slots so that a warp shades one KIND of vertex: surface hits first, then misses
		// (environment lookup + regeneration only), then dead slots.  Unsorted, nearly every warp holds both kinds and
		// executes both paths (profiles/r01_v8_final_summary.md: 20 of 32 lanes per instruction); any thread of the
		// block may shade any of its slots, the slot keeps its place in memory.
		{
			const uint32_t warp = threadIdx.x >> 5;
			uint32_t cls = 2u; // dead / out of range
			if (slot < A.nSlots && (__float_as_uint(A.rayD[slot].w) & WF_ALIVE)) cls = (__float_as_uint(A.hit[slot].x) == RTB_MISS_ID) ? 1u : 0u;
			const unsigned m0 = __ballot_sync(0xFFFFFFFFu, cls == 0u), m1 = __ballot_sync(0xFFFFFFFFu, cls == 1u);
			if (lane == 0) sClsCnt[0][warp] = __popc(m0), sClsCnt[1][warp] = __popc(m1);
			__syncthreads();
			uint32_t n0 = 0, n1 = 0, b0 = 0, b1 = 0;
			for (uint32_t w = 0; w < 4; w++)
			{
				if (w < warp) b0 += sClsCnt[0][w], b1 += sClsCnt[1][w];
				n0 += sClsCnt[0][w], n1 += sClsCnt[1][w];
			}
			const uint32_t below = (1u << lane) - 1u;
			uint32_t dst;
			if (cls == 0u) dst = b0 + __popc(m0 & below);
			else if (cls == 1u) dst = n0 + b1 + __popc(m1 & below);
			else dst = n0 + n1 + (warp * 32u - b0 - b1) + __popc(~(m0 | m1) & below);
			sPerm[dst] = (uint8_t)threadIdx.x;
			__syncthreads();
			slot = slot - threadIdx.x + sPerm[threadIdx.x];
		}
#endif
		bool inRange = slot < A.nSlots;
		bool haveVertex = false, slotLive = false;
		uint32_t queued = 0;
		float4 ro, rd, hh, tq;
#if WF_SHADE_PREFETCH == 2
		if (pre)
		{
			// the slot's 64 bytes of state were requested one round ago (cp.async into the block's staging rows): the
			// round starts on shared-memory latency instead of HBM latency, at no register cost
			asm volatile("cp.async.wait_group 0;" ::: "memory");
			if (inRange)
			{
				rd = sPre[round & 1u][0][threadIdx.x];
				if (__float_as_uint(rd.w) & WF_ALIVE)
				{
					ro = sPre[round & 1u][1][threadIdx.x], hh = sPre[round & 1u][2][threadIdx.x], tq = sPre[round & 1u][3][threadIdx.x];
					haveVertex = true;
				}
			}
			shadePrefetchSlot(A, slot + gridDim.x * blockDim.x, sPre[(round + 1u) & 1u]);
		}
		else
#endif
		if (inRange)
		{
#if WF_SHADE_PREFETCH == 1
			if (pre)
			{
				const uint32_t nx = slot + gridDim.x * blockDim.x;
				if (nx < A.nSlots)
				{
					asm volatile("prefetch.global.L2 [%0];" ::"l"(A.rayD + nx));
					asm volatile("prefetch.global.L2 [%0];" ::"l"(A.rayO + nx));
					asm volatile("prefetch.global.L2 [%0];" ::"l"(A.hit + nx));
					asm volatile("prefetch.global.L2 [%0];" ::"l"(A.thr + nx));
				}
			}
#endif
			rd = A.rayD[slot];
			if (__float_as_uint(rd.w) & WF_ALIVE)
			{
				ro = A.rayO[slot];
				hh = A.hit[slot];
				tq = A.thr[slot];
				haveVertex = true;
			}
		}
		// pass 0 shades the slot's vertex; with a primary-hit table, a thread whose path ended
		// shades the first vertex of its next job in the following pass instead of parking it in
		// the slot for a whole iteration (a primary miss then never touches slot memory)
		for (uint32_t pass = 0;; pass++)
		{
			bool done = false, haveShadow = false;
			if (haveVertex)
			{
				uint32_t flags = __float_as_uint(rd.w);
				uint32_t q = __float_as_uint(ro.w), n = __float_as_uint(tq.w);
				uint32_t depth = flags & WF_DEPTH_MASK;
				bool canHitLight = (flags & WF_CANHIT) != 0;
				uint32_t px, py;
				wfJobPixel(A, q, px, py);
				uint32_t pixel = py * A.width + px;
				uint32_t sample = A.sFirst + n * A.sStep;
				RayD ray = mkRay(mk(ro), mk(rd));
				V3 T = mk(tq);
				V3 add = mk(0.0f, 0.0f, 0.0f);
				done = true;
				uint32_t id = __float_as_uint(hh.x);
				if (id == RTB_MISS_ID)
				{
					if (INTEGRATOR == RTB_INT_PATH || INTEGRATOR == RTB_INT_PATH_MIS || INTEGRATOR == RTB_INT_ALBEDO) add = backgroundEval(S, ray.d);
				}
				else
				{
					ShadeD sd;
					calcShading(S, id, hh.y, hh.z, hh.w, 1.0f - (hh.z + hh.w), ray, sd);
					if (INTEGRATOR == RTB_INT_NORMALS)
					{
						add = mk(fabsf(sd.sN.x), fabsf(sd.sN.y), fabsf(sd.sN.z));
					}
					else
					{
						rtb_material m = S.mats[sd.mat];
						if (m.flags & RTB_MAT_LIGHT)
						{
							if (INTEGRATOR == RTB_INT_PATH || INTEGRATOR == RTB_INT_PATH_MIS)
							{
								if (canHitLight) add = T * mk(m.emission);
							}
							else
								add = mk(m.emission);
						}
						else if (INTEGRATOR == RTB_INT_ALBEDO)
						{
							add = bsdfEvaluate(S, m, sd, mk(0.0f, 1.0f, 0.0f));
						}
						else
						{
							float4 ua = rngBlock(P.seed, pixel, sample, 2u * depth);
							V3 p1, p2, contrib;
							if (INTEGRATOR == RTB_INT_PATH_MIS)
							{
								// computeDirectMIS (Renderer.h:474-557): the light strategy's segment and the BSDF
								// strategy's probe ray go into ONE queue record, resolved by k_wf_mis
								bool haveSeg, nonArea;
								float pdfAreaPmf;
								if (misLightSample(S, P, sd, m, ua.x, ua.y, ua.z, haveSeg, p2, contrib, nonArea, pdfAreaPmf))
								{
									float4 um = rngBlock(P.seed, pixel, sample, RTB_RNG_MIS_BLOCK + depth);
									V3 fB;
									float pdfB;
									V3 wiB = bsdfSample(S, m, sd, um.x, um.y, um.z, fB, pdfB);
									V3 payload = ((T * fB) * selMax(0.0f, dot(wiB, sd.sN))) / pdfB;
									V3 dir = mk(0.0f, 0.0f, 0.0f), o = sd.x, c = mk(0.0f, 0.0f, 0.0f);
									float maxT = 0.0f;
									if (haveSeg)
									{
										dir = p2 - sd.x;
										maxT = sqrtf(lengthSq(dir)) - (2.0f * P.epsilon);
										dir = normalize(dir);
										o = sd.x + (dir * P.epsilon);
										c = T * contrib;
									}
									sStage[0][threadIdx.x] = make_float4(o.x, o.y, o.z, maxT);
									sStage[1][threadIdx.x] = make_float4(dir.x, dir.y, dir.z, __uint_as_float(pixel));
									sStage[2][threadIdx.x] = make_float4(c.x, c.y, c.z, __uint_as_float((haveSeg ? 1u : 0u) | (nonArea ? 2u : 0u)));
									sStage[NREC - 3][threadIdx.x] = make_float4(sd.x.x, sd.x.y, sd.x.z, pdfB);
									sStage[NREC - 2][threadIdx.x] = make_float4(wiB.x, wiB.y, wiB.z, pdfAreaPmf);
									sStage[NREC - 1][threadIdx.x] = make_float4(payload.x, payload.y, payload.z, 0.0f);
									haveShadow = true;
								}
							}
							else if (directSample(S, P, sd, m, ua.x, ua.y, ua.z, p1, p2, contrib))
							{
								// Scene::visible's ray (Scene.h:161-169); traced by k_wf_shadow
								V3 dir = p2 - p1;
								float maxT = sqrtf(lengthSq(dir)) - (2.0f * P.epsilon);
								dir = normalize(dir);
								V3 o = p1 + (dir * P.epsilon);
								V3 c = (INTEGRATOR == RTB_INT_DIRECT) ? contrib : (T * contrib);
								sStage[0][threadIdx.x] = make_float4(o.x, o.y, o.z, maxT);
								sStage[1][threadIdx.x] = make_float4(dir.x, dir.y, dir.z, __uint_as_float(pixel));
								sStage[2][threadIdx.x] = make_float4(c.x, c.y, c.z, 0.0f);
								haveShadow = true;
							}
							if ((INTEGRATOR == RTB_INT_PATH || INTEGRATOR == RTB_INT_PATH_MIS) && !((int)depth > P.max_depth))
							{
								float rr = selMin(lum(T), P.rr_cap);
								if (ua.w < rr)
								{
									T = divFast(T, rr);
									float4 ub = rngBlock(P.seed, pixel, sample, 2u * depth + 1u);
									V3 f;
									float pdf;
									V3 wi = bsdfSample(S, m, sd, ub.x, ub.y, ub.z, f, pdf);
									bool spec = (m.flags & RTB_MAT_SPECULAR) != 0;
									if (spec) T = divFast(T * f, pdf);
									else T = divFast((T * f) * fabsf(dot(wi, sd.sN)), pdf);
									V3 o = sd.x + (wi * P.epsilon);
									A.rayO[slot] = make_float4(o.x, o.y, o.z, __uint_as_float(q));
									A.rayD[slot] = make_float4(wi.x, wi.y, wi.z,
									                           __uint_as_float(WF_ALIVE | (spec ? WF_CANHIT : 0u) | (depth + 1u)));
									A.thr[slot] = make_float4(T.x, T.y, T.z, __uint_as_float(n));
									done = false;
									slotLive = true;
								}
							}
						}
					}
				}
				filmAdd(A.accum, pixel, add);
			}
			haveVertex = false;
			// ---- one cooperative step reserves the block's shadow-queue entries AND its next jobs:
			// two ballots, two barriers, one atomic per counter and block (shared words are
			// double-buffered, so the next pass may write while stragglers still read)
			if (done) nDone++;
			const uint32_t warp = threadIdx.x >> 5, ph = REUSE ? ((phase++) & 1u) : (round & 1u);
			unsigned mS = __ballot_sync(0xFFFFFFFFu, haveShadow), mJ = __ballot_sync(0xFFFFFFFFu, done);
			if (lane == 0) sCountS[ph][warp] = __popc(mS), sCountJ[ph][warp] = __popc(mJ);
			__syncthreads();
			if (threadIdx.x == 0)
			{
				uint32_t tot = sCountS[ph][0] + sCountS[ph][1] + sCountS[ph][2] + sCountS[ph][3];
				sBaseS[ph] = tot ? atomicAdd(&A.ctrl[iter].nShadow, tot) : 0u;
			}
			if (threadIdx.x == 32)
			{
				uint32_t tot = sCountJ[ph][0] + sCountJ[ph][1] + sCountJ[ph][2] + sCountJ[ph][3];
				unsigned long long b = tot ? atomicAdd(&A.glob->nextJob, (unsigned long long)tot) : 0ull;
				sBaseJ[ph] = b;
				if (tot && !A.tileJobBase)
				{
					uint32_t Q = A.nTiles * 32u;
					sBaseN[ph] = (uint32_t)(b / Q), sBaseQ[ph] = (uint32_t)(b % Q);
				}
			}
			bool anyDone = __syncthreads_or(done) != 0;
			if (haveShadow)
			{
				uint32_t at = sBaseS[ph] + __popc(mS & ((1u << lane) - 1u));
				for (uint32_t w = 0; w < warp; w++) at += sCountS[ph][w];
				A.shO[at] = sStage[0][threadIdx.x], A.shD[at] = sStage[1][threadIdx.x], A.shC[at] = sStage[2][threadIdx.x];
				if (INTEGRATOR == RTB_INT_PATH_MIS)
					A.misA[at] = sStage[NREC - 3][threadIdx.x], A.misB[at] = sStage[NREC - 2][threadIdx.x], A.misC[at] = sStage[NREC - 1][threadIdx.x];
			}
			// ---- regeneration: finished paths take the next jobs of the render (the rare lanes that
			// drew a padding pixel of an edge tile retry warp-wise)
			bool fresh = false;
			uint32_t n = 0, q = 0;
			{
				bool retry = false;
				if (done)
				{
					uint32_t off = __popc(mJ & ((1u << lane) - 1u));
					for (uint32_t w = 0; w < warp; w++) off += sCountJ[ph][w];
					unsigned long long j = sBaseJ[ph] + off;
					if (j < A.totalJobs)
					{
						if (A.tileJobBase) fresh = wfDecodeJob(A, j, n, q);
						else
						{
							// (n, q) = (j / Q, j % Q) from the block's base: off < 128 <= Q
							uint32_t Q = A.nTiles * 32u, px, py;
							n = sBaseN[ph], q = sBaseQ[ph] + off;
							while (q >= Q) q -= Q, n++;
							fresh = wfJobPixel(A, q, px, py);
						}
						retry = !fresh;
					}
				}
				if (__any_sync(0xFFFFFFFFu, retry)) fresh = wfClaimJob(A, retry, n, q) || fresh;
			}
			if (done && !fresh) A.rayD[slot] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(0u));
			// REUSE = false compiles to the single-pass kernel (no loop-carried vertex state).  A thread that
			// has queued WF_SHADOW_PER_SLOT shadow rays in this launch parks its next job in the slot instead
			// of shading it now: that is the capacity of the shadow queue per slot (rtb_api.cu).
			if (haveShadow) queued++;
			const bool lastPass = !REUSE || pass + 1u >= A.primaryPasses;
			if (fresh && (lastPass || queued >= WF_SHADOW_PER_SLOT))
			{
				wfWriteJob(S, A, slot, n, q);
				slotLive = true;
				fresh = false;
			}
			if (lastPass) break;
			if (!anyDone) break; // no thread of the block can hold a fresh vertex
			if (fresh)
			{
				wfFirstVertex(S, A, n, q, ro, rd, hh, tq);
				haveVertex = true;
			}
		}
		if (slotLive) nAlive++;
	}
	// block-level totals: one atomic per block for `alive`, one striped atomic for the sample count
	for (int o = 16; o > 0; o >>= 1)
	{
		nAlive += __shfl_xor_sync(0xFFFFFFFFu, nAlive, o);
		nDone += __shfl_xor_sync(0xFFFFFFFFu, nDone, o);
	}
	__shared__ uint32_t sAlive[4], sDone[4];
	if (lane == 0) sAlive[threadIdx.x >> 5] = nAlive, sDone[threadIdx.x >> 5] = nDone;
	__syncthreads();
	if (threadIdx.x == 0)
	{
		uint32_t a = 0, d = 0;
		for (uint32_t i = 0; i < (blockDim.x >> 5); i++) a += sAlive[i], d += sDone[i];
		if (a) atomicAdd(&A.ctrl[iter].alive, a);
		if (d) atomicAdd(&A.counters[(size_t)(blockIdx.x % RTB_COUNTER_STRIPES) * 8], (unsigned long long)d);
	}
}
