// rtb_wavefront.cuh — the production schedule of the path-tracing loop: wavefront stages over a
// resident pool of path SLOTS (profiles/r01_v1_megakernel_summary.md: the one-thread-per-pixel
// megakernel keeps 7.9 of 32 lanes busy; stages run converged).
//
//   slot j  <->  pixel p = tileSwizzle(j mod Npix), stream k = j / Npix          (fixed mapping)
//   stream k of a pixel renders the local sample ordinals n = k, k+S, k+2S, ...  (S streams)
//
// One ITERATION advances every live slot by one path vertex:
//   k_wf_extend  persistent warps pull slots from the range [0, P) and run Scene::traverse
//                (Scene.h:107-130) for the slot's ray; a lane whose ray is finished refills from
//                the range at once, so warps stay full while rays differ in length.
//   k_wf_shade   one thread per slot: pathTrace's body for one vertex (Renderer.h:328-392):
//                miss / emitter / computeDirect's light sample (visibility deferred to a shadow
//                ray) / Russian roulette / BSDF sample; a finished path immediately starts the
//                stream's next sample (path regeneration) or retires the slot.
//   k_wf_shadow  persistent warps trace the deferred Scene::visible segments (Scene.h:161-169)
//                and add the already-weighted NEE contribution when unoccluded.
// Every radiance term is added to the SLOT's private accumulator in program order, so the film
// is bit-reproducible and independent of scheduling; k_wf_resolve adds the S accumulators of a
// pixel to the film in stream order: Film::splat with BoxFilter (Imaging.h:139-154, 209-232).
#pragma once
#include "rtb_kernels.cuh"

#define WF_ALIVE 0x200u
#define WF_CANHIT 0x100u
#define WF_DEPTH_MASK 0xFFu

struct WfCtrl
{
	unsigned int extendHead;  // next unclaimed slot of k_wf_extend
	unsigned int shadowHead;  // next unclaimed slot of k_wf_shadow
	unsigned int alive;       // slots still alive after k_wf_shade
	unsigned int pad_;
};

struct WfArgs
{
	float4* rayO;  // o.xyz, bits(sample ordinal n of the stream)
	float4* rayD;  // d.xyz, bits(flags | depth)
	float4* hit;   // bits(id), t, alpha, beta
	float4* thr;   // path throughput
	float4* acc;   // slot accumulator (sum over the slot's samples)
	float4* shO;   // shadow ray origin, maxT (< 0: none)
	float4* shD;   // shadow ray direction
	float4* shC;   // contribution if unoccluded (already multiplied by the throughput)
	WfCtrl* ctrl;  // [maxIterations + 1], zeroed before the batch
	unsigned long long* counters;
	float* film;
	uint32_t nSlots, nPix, streams;
	uint32_t width, height;
	uint32_t sFirst, sStep, sCount; // local sample ordinal n -> global sample index sFirst + n * sStep
	rtb_params P;
};

// slot -> pixel: consecutive 32 slots = an 8x4 pixel tile (coherent primary rays, coalesced
// state).  Returns false for the padding slots of a partial tile row/column.
RTB_DEV bool wfSlotPixel(const WfArgs& A, uint32_t slot, uint32_t& px, uint32_t& py, uint32_t& stream)
{
	uint32_t tilesX = (A.width + 7u) >> 3;
	stream = slot / A.nPix;
	uint32_t q = slot - stream * A.nPix;
	uint32_t tile = q >> 5, lane = q & 31u;
	px = (tile % tilesX) * 8u + (lane & 7u);
	py = (tile / tilesX) * 4u + (lane >> 3);
	return px < A.width && py < A.height;
}

RTB_DEV bool wfPixelOwned(const WfArgs& A, uint32_t px, uint32_t py)
{
	if (A.P.partition == RTB_PART_TILE && A.P.part_world > 1)
	{
		uint32_t t32x = (A.width + 31u) >> 5;
		uint32_t tile = (py >> 5) * t32x + (px >> 5);
		return (int)(tile % (uint32_t)A.P.part_world) == A.P.part_rank;
	}
	return true;
}

RTB_DEV void wfStartSample(const DevScene& S, const WfArgs& A, uint32_t slot, uint32_t px, uint32_t py, uint32_t n)
{
	RayD r = generateRay(S.cam, (float)px + 0.5f, (float)py + 0.5f);
	A.rayO[slot] = make_float4(r.o.x, r.o.y, r.o.z, __uint_as_float(n));
	A.rayD[slot] = make_float4(r.d.x, r.d.y, r.d.z, __uint_as_float(WF_ALIVE | WF_CANHIT));
	A.thr[slot] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
}

// nPix here is the PADDED pixel count (tiles * 32) so that the swizzle is a bijection
__global__ void __launch_bounds__(256) k_wf_init(const __grid_constant__ DevScene S, const __grid_constant__ WfArgs A)
{
	uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
	if (slot >= A.nSlots) return;
	uint32_t px, py, k;
	bool ok = wfSlotPixel(A, slot, px, py, k) && wfPixelOwned(A, px, py) && k < A.sCount;
	A.acc[slot] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	A.shO[slot] = make_float4(0.0f, 0.0f, 0.0f, -1.0f);
	if (ok) wfStartSample(S, A, slot, px, py, k);
	else A.rayD[slot] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(0u));
}

// ---------------------------------------------------------------------------------------
// FAST closest-hit as a resumable per-lane state machine (same decisions as closestFast in
// rtb_dev_scene.cuh: exact slab arithmetic on every box, near-first order, culling with the
// relative slack, lexicographic (t, ID) winner), driven by persistent warps.
// ---------------------------------------------------------------------------------------
struct Trav
{
	RayD r;
	HitD h;
	int32_t cur;   // >= 0 interior node, < 0 leaf reference, RTB_TRAV_DONE finished
	int sp;
};
#define RTB_TRAV_DONE 0x7FFFFFFF

template <bool ANYHIT>
RTB_DEV void travPop(Trav& t, const int32_t* stackNode, const float* stackT, float cullRel, float maxT)
{
	for (;;)
	{
		if (t.sp == 0)
		{
			t.cur = RTB_TRAV_DONE;
			return;
		}
		t.sp--;
		float te = stackT[t.sp];
		float lim = ANYHIT ? maxT : t.h.t;
		bool cull = ANYHIT ? ((te - fabsf(te) * cullRel) >= lim) : ((te - fabsf(te) * cullRel) > lim);
		if (cull) continue;
		t.cur = stackNode[t.sp];
		return;
	}
}

template <int TRAV>
__global__ void __launch_bounds__(128) k_wf_extend(const __grid_constant__ DevScene S, const __grid_constant__ WfArgs A, uint32_t iter)
{
	const rtb_params& P = A.P;
	WfCtrl* ctrl = A.ctrl + iter;
	const uint32_t lane = threadIdx.x & 31u;
	Tally tl = {0, 0, 0, 0, 0};
	if (iter > 0 && A.ctrl[iter - 1].alive == 0) return; // pool drained
	if (TRAV == RTB_TRAV_EXACT)
	{
		// parity path: plain per-lane exhaustive traversal of the reference tree
		for (;;)
		{
			uint32_t base = 0;
			if (lane == 0) base = atomicAdd(&ctrl->extendHead, 32u);
			base = __shfl_sync(0xFFFFFFFFu, base, 0);
			if (base >= A.nSlots) break;
			uint32_t slot = base + lane;
			if (slot < A.nSlots)
			{
				float4 d = A.rayD[slot];
				if (__float_as_uint(d.w) & WF_ALIVE)
				{
					float4 o = A.rayO[slot];
					RayD r = mkRay(mk(o), mk(d));
					HitD h;
					closestExact(S, r, P.epsilon, h, tl.box, tl.tri);
					tl.closest++;
					A.hit[slot] = make_float4(__uint_as_float(h.id), h.t, h.alpha, h.beta);
				}
			}
		}
		flushTally(tl, A.counters);
		return;
	}
	int32_t stackNode[RTB_STACK];
	float stackT[RTB_STACK];
	Trav t;
	t.cur = RTB_TRAV_DONE;
	t.sp = 0;
	uint32_t slot = 0xFFFFFFFFu;
	bool have = false;
	bool exhausted = false;
	for (;;)
	{
		// ---- refill idle lanes from the slot range
		unsigned idle = __ballot_sync(0xFFFFFFFFu, !have);
		if (idle && !exhausted)
		{
			uint32_t n = __popc(idle), base = 0;
			int leader = __ffs(idle) - 1;
			if ((int)lane == leader) base = atomicAdd(&ctrl->extendHead, n);
			base = __shfl_sync(0xFFFFFFFFu, base, leader);
			if (base >= A.nSlots) exhausted = true;
			if (!have)
			{
				uint32_t mine = base + __popc(idle & ((1u << lane) - 1u));
				if (mine < A.nSlots)
				{
					float4 d = A.rayD[mine];
					if (__float_as_uint(d.w) & WF_ALIVE)
					{
						float4 o = A.rayO[mine];
						t.r = mkRay(mk(o), mk(d));
						t.h.id = RTB_MISS_ID, t.h.t = FLT_MAX, t.h.alpha = t.h.beta = 0.0f;
						t.sp = 0;
						slot = mine;
						have = true;
						tl.closest++;
						if (rayIsDegenerate(t.r) || S.fast_root < 0)
						{
							// 0*inf = NaN rays (and single-leaf scenes) take the reference's own tree
							closestExact(S, t.r, P.epsilon, t.h, tl.box, tl.tri);
							t.cur = RTB_TRAV_DONE;
						}
						else
							t.cur = S.fast_root;
					}
				}
			}
		}
		unsigned active = __ballot_sync(0xFFFFFFFFu, have);
		if (!active)
		{
			if (exhausted) break;
			continue;
		}
		// ---- traverse until enough lanes have finished to make a refill worthwhile
		for (;;)
		{
			// interior nodes: both children per step
			while (have && t.cur >= 0 && t.cur != RTB_TRAV_DONE)
			{
				const float4* nd = S.fnodes + (size_t)t.cur * 4;
				float4 n0 = ldg4(nd), n1 = ldg4(nd + 1), nz = ldg4(nd + 2), ch = ldg4(nd + 3);
				float t0, t1;
				tl.box += 2;
				bool h0 = slabTest(n0.x, n0.z, nz.x, n0.y, n0.w, nz.y, t.r, t0);
				bool h1 = slabTest(n1.x, n1.z, nz.z, n1.y, n1.w, nz.w, t.r, t1);
				h0 = h0 && !((t0 - fabsf(t0) * P.cull_rel) > t.h.t);
				h1 = h1 && !((t1 - fabsf(t1) * P.cull_rel) > t.h.t);
				int32_t c0 = __float_as_int(ch.x), c1 = __float_as_int(ch.y);
				if (h0 && h1)
				{
					bool swap = t1 < t0;
					stackNode[t.sp] = swap ? c0 : c1;
					stackT[t.sp] = swap ? t0 : t1;
					t.sp++;
					t.cur = swap ? c1 : c0;
				}
				else if (h0) t.cur = c0;
				else if (h1) t.cur = c1;
				else travPop<false>(t, stackNode, stackT, P.cull_rel, 0.0f);
			}
			// leaves (<= 2 triangles; the box that admitted them was the exact leaf box)
			if (have && t.cur < 0)
			{
				leafClosest(S, t.cur, t.r, P.epsilon, t.h, tl.tri);
				travPop<false>(t, stackNode, stackT, P.cull_rel, 0.0f);
			}
			if (have && t.cur == RTB_TRAV_DONE)
			{
				A.hit[slot] = make_float4(__uint_as_float(t.h.id), t.h.t, t.h.alpha, t.h.beta);
				have = false;
			}
			unsigned still = __ballot_sync(0xFFFFFFFFu, have);
			if (!still) break;
			if (!exhausted && __popc(still) <= 20) break; // refill
		}
	}
	flushTally(tl, A.counters);
}

template <int TRAV>
__global__ void __launch_bounds__(128) k_wf_shadow(const __grid_constant__ DevScene S, const __grid_constant__ WfArgs A, uint32_t iter)
{
	const rtb_params& P = A.P;
	WfCtrl* ctrl = A.ctrl + iter;
	const uint32_t lane = threadIdx.x & 31u;
	Tally tl = {0, 0, 0, 0, 0};
	if (iter > 0 && A.ctrl[iter - 1].alive == 0) return; // pool drained (its last shadow rays ran in iter-1)
	int32_t stackNode[RTB_STACK];
	float stackT[RTB_STACK];
	Trav t;
	t.cur = RTB_TRAV_DONE;
	t.sp = 0;
	uint32_t slot = 0xFFFFFFFFu;
	float maxT = 0.0f;
	bool have = false, exhausted = false, occluded = false;
	for (;;)
	{
		unsigned idle = __ballot_sync(0xFFFFFFFFu, !have);
		if (idle && !exhausted)
		{
			uint32_t n = __popc(idle), base = 0;
			int leader = __ffs(idle) - 1;
			if ((int)lane == leader) base = atomicAdd(&ctrl->shadowHead, n);
			base = __shfl_sync(0xFFFFFFFFu, base, leader);
			if (base >= A.nSlots) exhausted = true;
			if (!have)
			{
				uint32_t mine = base + __popc(idle & ((1u << lane) - 1u));
				if (mine < A.nSlots)
				{
					float4 o = A.shO[mine];
					if (o.w >= 0.0f)
					{
						float4 d = A.shD[mine];
						t.r = mkRay(mk(o), mk(d));
						maxT = o.w;
						t.sp = 0;
						slot = mine;
						have = true;
						occluded = false;
						tl.shadow++;
						A.shO[mine] = make_float4(0.0f, 0.0f, 0.0f, -1.0f); // consumed
						if (TRAV == RTB_TRAV_EXACT || rayIsDegenerate(t.r) || S.fast_root < 0)
						{
							occluded = !visibleExact(S, t.r, P.epsilon, maxT, tl.box, tl.tri);
							t.cur = RTB_TRAV_DONE;
						}
						else
							t.cur = S.fast_root;
					}
				}
			}
		}
		unsigned active = __ballot_sync(0xFFFFFFFFu, have);
		if (!active)
		{
			if (exhausted) break;
			continue;
		}
		for (;;)
		{
			while (have && t.cur >= 0 && t.cur != RTB_TRAV_DONE)
			{
				const float4* nd = S.fnodes + (size_t)t.cur * 4;
				float4 n0 = ldg4(nd), n1 = ldg4(nd + 1), nz = ldg4(nd + 2), ch = ldg4(nd + 3);
				float t0, t1;
				tl.box += 2;
				bool h0 = slabTest(n0.x, n0.z, nz.x, n0.y, n0.w, nz.y, t.r, t0);
				bool h1 = slabTest(n1.x, n1.z, nz.z, n1.y, n1.w, nz.w, t.r, t1);
				h0 = h0 && !((t0 - fabsf(t0) * P.cull_rel) >= maxT);
				h1 = h1 && !((t1 - fabsf(t1) * P.cull_rel) >= maxT);
				int32_t c0 = __float_as_int(ch.x), c1 = __float_as_int(ch.y);
				if (h0 && h1)
				{
					stackNode[t.sp] = c1;
					stackT[t.sp] = t1;
					t.sp++;
					t.cur = c0;
				}
				else if (h0) t.cur = c0;
				else if (h1) t.cur = c1;
				else travPop<true>(t, stackNode, stackT, P.cull_rel, maxT);
			}
			if (have && t.cur < 0)
			{
				if (leafOccludes(S, t.cur, t.r, P.epsilon, maxT, tl.tri))
				{
					occluded = true;
					t.cur = RTB_TRAV_DONE;
				}
				else
					travPop<true>(t, stackNode, stackT, P.cull_rel, maxT);
			}
			if (have && t.cur == RTB_TRAV_DONE)
			{
				if (!occluded)
				{
					float4 c = A.shC[slot];
					float4 a = A.acc[slot];
					A.acc[slot] = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, 0.0f);
				}
				have = false;
			}
			unsigned still = __ballot_sync(0xFFFFFFFFu, have);
			if (!still) break;
			if (!exhausted && __popc(still) <= 20) break;
		}
	}
	flushTally(tl, A.counters);
}

// ---------------------------------------------------------------------------------------
// NEE with deferred visibility: RayTracer::computeDirect (Renderer.h:423-473) up to the
// scene->visible() call; returns false when no shadow ray is needed (contribution is zero).
// ---------------------------------------------------------------------------------------
RTB_DEV bool directSample(const DevScene& S, const rtb_params& P, const ShadeD& sd, const rtb_material& m, float uPick,
                          float r1, float r2, V3& p1, V3& p2, V3& contrib)
{
	if (m.flags & RTB_MAT_SPECULAR) return false;
	if (S.n_lights == 0) return false;
	float nl = (float)S.n_lights;
	float pmf = 1.0f / nl;
	int li = (int)(nl * uPick);
	if (li > (int)S.n_lights - 1) li = (int)S.n_lights - 1;
	rtb_light L = S.lights[li];
	if (L.type == RTB_LIGHT_AREA)
	{
		V3 p = trianglePoint(S, L.triangle, r1, r2);
		float pdf = 1.0f / L.area;
		V3 wi = p - sd.x;
		float l = lengthSq(wi);
		wi = normalize(wi);
		V3 nL = triangleGNormal(S, L.triangle);
		float G = (selMax(dot(wi, sd.sN), 0.0f) * selMax(-dot(wi, nL), 0.0f)) / l;
		if (!(G > 0.0f)) return false;
		contrib = ((bsdfEvaluate(S, m, sd, wi) * mk(L.emission)) * G) / (pmf * pdf);
		p1 = sd.x;
		p2 = p;
		return true;
	}
	V3 wi, emitted;
	float pdf;
	if (L.type == RTB_LIGHT_ENVMAP && P.sampling == RTB_SAMPLING_IMPORTANCE && S.env_marginal != nullptr)
	{
		int W = S.env_w, H = S.env_h;
		int lo = 0, hi = H;
		while (hi - lo > 1)
		{
			int mid = (lo + hi) >> 1;
			if (__ldg(S.env_marginal + mid) <= r1) lo = mid;
			else hi = mid;
		}
		int row = lo;
		float m0 = __ldg(S.env_marginal + row), m1 = __ldg(S.env_marginal + row + 1);
		float fr = (m1 > m0) ? (r1 - m0) / (m1 - m0) : 0.5f;
		const float* cd = S.env_cond + (size_t)row * (W + 1);
		lo = 0, hi = W;
		while (hi - lo > 1)
		{
			int mid = (lo + hi) >> 1;
			if (__ldg(cd + mid) <= r2) lo = mid;
			else hi = mid;
		}
		int col = lo;
		float c0 = __ldg(cd + col), c1 = __ldg(cd + col + 1);
		float fc = (c1 > c0) ? (r2 - c0) / (c1 - c0) : 0.5f;
		float v = ((float)row + fr) / (float)H;
		float u = ((float)col + fc) / (float)W;
		float theta = v * RTB_PI_F, phi = u * (2.0f * RTB_PI_F);
		float st, ct, sp, cp;
		sincosf(theta, &st, &ct);
		sincosf(phi, &sp, &cp);
		wi = mk(cp * st, ct, sp * st);
		float pmfTexel = (m1 - m0) * (c1 - c0);
		pdf = pmfTexel * ((float)W * (float)H) / (2.0f * RTB_PI_F * RTB_PI_F * fmaxf(st, 1e-8f));
		if (!(pdf > 0.0f)) return false;
		emitted = envLookup(S, L.tex, wi);
	}
	else
	{
		wi = uniformSampleSphere(r1, r2);
		pdf = (float)(1.0 / (4.0 * RTB_PI_D));
		emitted = (L.type == RTB_LIGHT_ENVMAP) ? envLookup(S, L.tex, wi) : mk(L.emission);
	}
	float G = selMax(dot(wi, sd.sN), 0.0f);
	if (!(G > 0.0f)) return false;
	contrib = ((bsdfEvaluate(S, m, sd, wi) * emitted) * G) / (pmf * pdf);
	p1 = sd.x;
	p2 = sd.x + (wi * 10000.0f);
	return true;
}

template <int INTEGRATOR>
__global__ void __launch_bounds__(128) k_wf_shade(const __grid_constant__ DevScene S, const __grid_constant__ WfArgs A, uint32_t iter)
{
	const rtb_params& P = A.P;
	uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t samplesDone = 0;
	bool aliveAfter = false;
	if (iter > 0 && A.ctrl[iter - 1].alive == 0) return; // pool drained
	if (slot < A.nSlots)
	{
		float4 rd = A.rayD[slot];
		uint32_t flags = __float_as_uint(rd.w);
		if (flags & WF_ALIVE)
		{
			float4 ro = A.rayO[slot];
			float4 hh = A.hit[slot];
			uint32_t n = __float_as_uint(ro.w);
			uint32_t depth = flags & WF_DEPTH_MASK;
			bool canHitLight = (flags & WF_CANHIT) != 0;
			uint32_t px, py, stream;
			wfSlotPixel(A, slot, px, py, stream);
			uint32_t pixel = py * A.width + px;
			uint32_t sample = A.sFirst + n * A.sStep;
			RayD ray = mkRay(mk(ro), mk(rd));
			float4 tq = A.thr[slot];
			V3 T = mk(tq);
			V3 add = mk(0.0f, 0.0f, 0.0f);
			bool done = true;
			uint32_t id = __float_as_uint(hh.x);
			if (id == RTB_MISS_ID)
			{
				if (INTEGRATOR == RTB_INT_PATH || INTEGRATOR == RTB_INT_ALBEDO) add = backgroundEval(S, ray.d);
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
						if (INTEGRATOR == RTB_INT_PATH)
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
						if (directSample(S, P, sd, m, ua.x, ua.y, ua.z, p1, p2, contrib))
						{
							// Scene::visible's ray (Scene.h:161-169); traced by k_wf_shadow
							V3 dir = p2 - p1;
							float maxT = sqrtf(lengthSq(dir)) - (2.0f * P.epsilon);
							dir = normalize(dir);
							V3 o = p1 + (dir * P.epsilon);
							V3 c = (INTEGRATOR == RTB_INT_DIRECT) ? contrib : (T * contrib);
							// maxT < 0: nothing can lie in (eps, maxT) -> visible (traverseVisible never rejects)
							A.shO[slot] = make_float4(o.x, o.y, o.z, fmaxf(maxT, 0.0f));
							A.shD[slot] = make_float4(dir.x, dir.y, dir.z, 0.0f);
							A.shC[slot] = make_float4(c.x, c.y, c.z, 0.0f);
						}
						if (INTEGRATOR == RTB_INT_PATH && !((int)depth > P.max_depth))
						{
							float rr = selMin(lum(T), P.rr_cap);
							if (ua.w < rr)
							{
								T = T / rr;
								float4 ub = rngBlock(P.seed, pixel, sample, 2u * depth + 1u);
								V3 f;
								float pdf;
								V3 wi = bsdfSample(S, m, sd, ub.x, ub.y, ub.z, f, pdf);
								bool spec = (m.flags & RTB_MAT_SPECULAR) != 0;
								if (spec) T = (T * f) / pdf;
								else T = ((T * f) * fabsf(dot(wi, sd.sN))) / pdf;
								V3 o = sd.x + (wi * P.epsilon);
								A.rayO[slot] = make_float4(o.x, o.y, o.z, __uint_as_float(n));
								A.rayD[slot] = make_float4(wi.x, wi.y, wi.z,
								                           __uint_as_float(WF_ALIVE | (spec ? WF_CANHIT : 0u) | (depth + 1u)));
								A.thr[slot] = make_float4(T.x, T.y, T.z, 0.0f);
								done = false;
								aliveAfter = true;
							}
						}
					}
				}
			}
			if (add.x != 0.0f || add.y != 0.0f || add.z != 0.0f)
			{
				float4 a = A.acc[slot];
				A.acc[slot] = make_float4(a.x + add.x, a.y + add.y, a.z + add.z, 0.0f);
			}
			if (done)
			{
				samplesDone = 1;
				uint32_t next = n + A.streams;
				if (next < A.sCount)
				{
					wfStartSample(S, A, slot, px, py, next);
					aliveAfter = true;
				}
				else
					A.rayD[slot] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(0u));
			}
		}
	}
	// tallies: samples finished, slots alive after this iteration
	unsigned aliveMask = __ballot_sync(0xFFFFFFFFu, aliveAfter);
	unsigned doneMask = __ballot_sync(0xFFFFFFFFu, samplesDone != 0);
	if ((threadIdx.x & 31u) == 0)
	{
		if (aliveMask) atomicAdd(&A.ctrl[iter].alive, (unsigned)__popc(aliveMask));
		if (doneMask) atomicAdd(&A.counters[0], (unsigned long long)__popc(doneMask));
	}
}

// film[pixel] += sum over the pixel's streams, in stream order
__global__ void __launch_bounds__(256) k_wf_resolve(const __grid_constant__ WfArgs A)
{
	uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= A.nPix) return;
	uint32_t px, py, k;
	if (!wfSlotPixel(A, q, px, py, k)) return;
	float r = 0.0f, g = 0.0f, b = 0.0f;
	for (uint32_t s = 0; s < A.streams; s++)
	{
		float4 a = A.acc[q + s * A.nPix];
		r += a.x, g += a.y, b += a.z;
	}
	float* f = A.film + ((size_t)py * A.width + px) * 3;
	f[0] += r, f[1] += g, f[2] += b;
}
