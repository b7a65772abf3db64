// rtb_kernels.cuh — the kernels of librtb200.so (sm_100a).  Launch wrappers live in
// rtb_api.cu.  Hot path: k_render (RayTracer::render -> renderTile -> pathTrace /
// computeDirect -> Scene::traverse / visible -> Film::splat, RTBase/Renderer.h:328-473,
// 795-885).  The rest are the batched parity entry points of include/rtb.h.
#pragma once
#include "rtb_dev_scene.cuh"

struct RenderArgs
{
	long long* accum;              // width*height*3 fixed-point running sums (Film::film, see filmAdd)
	unsigned long long* counters;  // [0] samples [1] closest rays [2] shadow rays [3] box [4] tri (closest) [5] box [6] tri (shadow)
	uint32_t spp_begin, spp_count;
	uint32_t width, height;
	rtb_params P;
};

struct Tally
{
	uint32_t samples, closest, shadow, box, tri, sbox, stri;
};

// Work counters are STRIPED: RTB_COUNTER_STRIPES rows of 8 x u64, a block adds to row
// blockIdx.x % stripes, the host sums the rows.  (One row made ~10^4 same-address atomics per
// launch serialise in the L2 atomic unit: a fixed cost of tens of microseconds per kernel.)
#define RTB_COUNTER_STRIPES 64
RTB_DEV void flushTally(const Tally& c, unsigned long long* counters)
{
	counters += (size_t)(blockIdx.x % RTB_COUNTER_STRIPES) * 8;
	uint32_t v[7] = {c.samples, c.closest, c.shadow, c.box, c.tri, c.sbox, c.stri};
#pragma unroll
	for (int k = 0; k < 7; k++)
	{
		uint32_t x = v[k];
		for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, o);
		if ((threadIdx.x & 31) == 0 && x) atomicAdd(&counters[k], (unsigned long long)x);
	}
}

// ---------------------------------------------------------------------------------------
// fixed-point film
// ---------------------------------------------------------------------------------------
#define RTB_FIX_SCALE 4294967296.0f /* 2^32 */
#define RTB_FIX_LIMIT 1073741824.0f /* 2^30: a single contribution is clamped to +-1e9 */

RTB_DEV long long toFixed(float c)
{
	// NaN contributes nothing (the reference would poison the pixel for good); +-inf clamps
	if (!(c == c)) return 0ll;
	c = fminf(fmaxf(c, -RTB_FIX_LIMIT), RTB_FIX_LIMIT);
	return __float2ll_rn(c * RTB_FIX_SCALE);
}
RTB_DEV void filmAdd(long long* accum, uint32_t pixel, V3 c)
{
	long long* a = accum + (size_t)pixel * 3;
	if (c.x != 0.0f) atomicAdd((unsigned long long*)a, (unsigned long long)toFixed(c.x));
	if (c.y != 0.0f) atomicAdd((unsigned long long*)a + 1, (unsigned long long)toFixed(c.y));
	if (c.z != 0.0f) atomicAdd((unsigned long long*)a + 2, (unsigned long long)toFixed(c.z));
}

// ---------------------------------------------------------------------------------------
// Direction sample of a non-area light (BackgroundColour / EnvironmentMap).
//   STRICT     = the reference: uniform sphere, pdf 1/4pi (Lights.h:99-100, 143-149).
//   IMPORTANCE = env-map luminance CDF (tables: rtb_accel.hpp buildEnvTables): row from the marginal
//                CDF with r1, column from that row's conditional CDF with r2, uniform inside the texel
//                cell; pdf = cell pmf x W H / (2 pi^2 sin theta).  Positive wherever the map can be
//                non-zero, so the estimator's expectation is the reference's.
// Shared by computeDirect and computeDirectMIS (both schedules) and rtb_eval_light.  Returns false when the
// density is not positive (nothing to add).
// ---------------------------------------------------------------------------------------
RTB_DEV bool sampleNonAreaLight(const DevScene& S, const rtb_params& P, const rtb_light& L, float r1, float r2, V3& wi, float& pdf,
                                V3& emitted)
{
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
		return true;
	}
	wi = uniformSampleSphere(r1, r2);
	pdf = 0.0795774683356285095f; // 1 / (4 pi)
	emitted = (L.type == RTB_LIGHT_ENVMAP) ? envLookup(S, L.tex, wi) : mk(L.emission);
	return true;
}

// ---------------------------------------------------------------------------------------
// RayTracer::computeDirect (RTBase/Renderer.h:423-473) with Scene::sampleLight (Scene.h:131-140),
// split at the scene->visible() call so that the wavefront schedule can defer the shadow ray:
// directSample returns false when no shadow ray is needed (the contribution is zero), else the
// segment p1 -> p2 to test and the contribution if it is unoccluded.  u = (light pick, r1, r2).
// (Queueing the env-map lookups of miss lanes and NEE lanes for one shared call site was measured
// 4 % SLOWER: the longer live ranges spill at 80 registers.)
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
		contrib = divFast((bsdfEvaluate(S, m, sd, wi) * mk(L.emission)) * G, pmf * pdf);
		p1 = sd.x;
		p2 = p;
		return true;
	}
	V3 wi, emitted;
	float pdf;
	if (!sampleNonAreaLight(S, P, L, r1, r2, wi, pdf, emitted)) return false;
	float G = selMax(dot(wi, sd.sN), 0.0f);
	if (!(G > 0.0f)) return false;
	contrib = divFast((bsdfEvaluate(S, m, sd, wi) * emitted) * G, pmf * pdf);
	p1 = sd.x;
	p2 = sd.x + (wi * 10000.0f);
	return true;
}

template <int TRAV>
RTB_DEV V3 computeDirect(const DevScene& S, const rtb_params& P, const ShadeD& sd, const rtb_material& m, float uPick,
                         float r1, float r2, Tally& tl)
{
	V3 p1, p2, contrib;
	if (!directSample(S, P, sd, m, uPick, r1, r2, p1, p2, contrib)) return mk(0.0f, 0.0f, 0.0f);
	tl.shadow++;
	if (sceneVisible<TRAV>(S, p1, p2, P.epsilon, P.cull_rel, tl.sbox, tl.stri)) return contrib;
	return mk(0.0f, 0.0f, 0.0f);
}

// ---------------------------------------------------------------------------------------
// RayTracer::computeDirectMIS (Renderer.h:474-557): light strategy + BSDF strategy combined with
// the balance heuristic (:408-410) after converting the light's area pdf to solid angle
// (:411-422).  The reference's quirks are kept: a visible
// non-area light returns un-weighted at once; the BSDF strategy uses the pdf of the light that
// was SAMPLED.  um = (bsdf r1, r2, r3) of the BSDF strategy.
// ---------------------------------------------------------------------------------------
template <int TRAV>
RTB_DEV V3 computeDirectMIS(const DevScene& S, const rtb_params& P, const ShadeD& sd, const rtb_material& m, float uPick,
                            float r1, float r2, float4 um, Tally& tl)
{
	V3 zero = mk(0.0f, 0.0f, 0.0f);
	if (m.flags & RTB_MAT_SPECULAR) return zero;
	if (S.n_lights == 0) return zero;
	float nl = (float)S.n_lights;
	float pmf = 1.0f / nl;
	int li = (int)(nl * uPick);
	if (li > (int)S.n_lights - 1) li = (int)S.n_lights - 1;
	rtb_light L = S.lights[li];
	V3 result = zero;
	float pdf;
	if (L.type == RTB_LIGHT_AREA)
	{
		V3 p = trianglePoint(S, L.triangle, r1, r2);
		pdf = 1.0f / L.area;
		V3 wi = p - sd.x;
		float l = lengthSq(wi);
		wi = normalize(wi);
		float cs = selMax(dot(wi, sd.sN), 0.0f);
		float cl = selMax(-dot(wi, triangleGNormal(S, L.triangle)), 0.0f);
		float G = cs * cl / l;
		if (G > 0.0f)
		{
			tl.shadow++;
			if (sceneVisible<TRAV>(S, sd.x, p, P.epsilon, P.cull_rel, tl.sbox, tl.stri))
			{
				float pdfB = bsdfPdf(m, sd, wi);
				float pdfL = (cl > 0.0f) ? ((pdf * pmf) * l / cl) : 0.0f;
				float w = pdfL / (pdfL + pdfB);
				result = result + ((((bsdfEvaluate(S, m, sd, wi) * mk(L.emission)) * G) * w) / (pmf * pdf));
			}
		}
	}
	else
	{
		V3 wi, emitted;
		bool ok = sampleNonAreaLight(S, P, L, r1, r2, wi, pdf, emitted);
		float G = ok ? selMax(dot(wi, sd.sN), 0.0f) : 0.0f;
		if (G > 0.0f)
		{
			tl.shadow++;
			if (sceneVisible<TRAV>(S, sd.x, sd.x + (wi * 10000.0f), P.epsilon, P.cull_rel, tl.sbox, tl.stri))
				return ((bsdfEvaluate(S, m, sd, wi) * emitted) * G) / (pmf * pdf);
		}
	}
	V3 f;
	float pdfBsdf;
	V3 wiB = bsdfSample(S, m, sd, um.x, um.y, um.z, f, pdfBsdf);
	RayD r = mkRay(sd.x + (wiB * P.epsilon), wiB);
	HitD h;
	closestHit<TRAV>(S, r, P.epsilon, P.cull_rel, h, tl.box, tl.tri);
	tl.closest++;
	if (h.id != RTB_MISS_ID)
	{
		ShadeD sh;
		calcShading(S, h.id, h.t, h.alpha, h.beta, 1.0f - (h.alpha + h.beta), r, sh);
		rtb_material mh = S.mats[sh.mat];
		if (mh.flags & RTB_MAT_LIGHT)
		{
			V3 wi = sh.x - sd.x;
			float dist2 = lengthSq(wi);
			wi = normalize(wi);
			float cl = selMax(0.0f, dot(mk(-wi.x, -wi.y, -wi.z), sh.sN));
			float pdfL = (cl > 0.0f) ? ((pdf * pmf) * dist2 / cl) : 0.0f;
			float w = pdfBsdf / (pdfBsdf + pdfL);
			result = result + ((((f * mk(mh.emission)) * selMax(0.0f, dot(wiB, sd.sN))) * w) / pdfBsdf);
		}
	}
	return result;
}

// The light strategy of computeDirectMIS (Renderer.h:483-529) up to the scene->visible() call, for the
// wavefront schedule.  Returns false for specular surfaces / scenes without lights (computeDirectMIS
// returns black there and traces nothing).  Otherwise: haveSeg = a visibility segment sd.x -> p2 is needed;
// contrib = what to add if it is unoccluded (balance weight applied for area lights); nonArea = a visible
// light ends the estimator without the BSDF strategy (:516-527); pdfAreaPmf = pdf * pmf of the sampled
// light, which the BSDF strategy converts to solid angle at whatever emitter it hits (:548).
RTB_DEV bool misLightSample(const DevScene& S, const rtb_params& P, const ShadeD& sd, const rtb_material& m, float uPick, float r1,
                            float r2, bool& haveSeg, V3& p2, V3& contrib, bool& nonArea, float& pdfAreaPmf)
{
	haveSeg = false, nonArea = false;
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
		pdfAreaPmf = pdf * pmf;
		V3 wi = p - sd.x;
		float l = lengthSq(wi);
		wi = normalize(wi);
		float cs = selMax(dot(wi, sd.sN), 0.0f);
		float cl = selMax(-dot(wi, triangleGNormal(S, L.triangle)), 0.0f);
		float G = cs * cl / l;
		if (G > 0.0f)
		{
			float pdfB = bsdfPdf(m, sd, wi);
			float pdfL = (cl > 0.0f) ? (pdfAreaPmf * l / cl) : 0.0f;
			float w = pdfL / (pdfL + pdfB);
			contrib = (((bsdfEvaluate(S, m, sd, wi) * mk(L.emission)) * G) * w) / (pmf * pdf);
			p2 = p;
			haveSeg = true;
		}
		return true;
	}
	nonArea = true;
	V3 wi, emitted;
	float pdf;
	bool ok = sampleNonAreaLight(S, P, L, r1, r2, wi, pdf, emitted);
	pdfAreaPmf = pdf * pmf;
	float G = ok ? selMax(dot(wi, sd.sN), 0.0f) : 0.0f;
	if (G > 0.0f)
	{
		contrib = ((bsdfEvaluate(S, m, sd, wi) * emitted) * G) / (pmf * pdf);
		p2 = sd.x + (wi * 10000.0f);
		haveSeg = true;
	}
	return true;
}

// ---------------------------------------------------------------------------------------
// k_render: one thread per pixel (a warp = an 8x4 pixel tile); each thread runs its pixel's
// samples back to back, so a lane whose path ends early immediately regenerates the next
// sample instead of idling ("path regeneration").  The pixel's colour is summed in
// registers and added to the film once: Film::splat with BoxFilter, size() == 0
// (RTBase/Imaging.h:139-154, 209-232) without atomics.
//
// INTEGRATOR: rtb_integrator.  pathTrace's recursion (Renderer.h:328-392) is unrolled into
// the loop below; SURVEY A.6 lists the estimator's quirks that are reproduced on purpose:
// the miss term is NOT weighted by the throughput, emitters count only after a specular
// bounce (or at depth 0), Russian roulette from depth 0 with p = min(Lum(T), 0.9).
// ---------------------------------------------------------------------------------------
template <int TRAV, int INTEGRATOR>
__global__ void __launch_bounds__(64) k_render(const __grid_constant__ DevScene S, const __grid_constant__ RenderArgs A)
{
	const rtb_params& P = A.P;
	uint32_t lane = threadIdx.x & 31u;
	uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	uint32_t tilesX = (A.width + 7u) >> 3, tilesY = (A.height + 3u) >> 2;
	Tally tl = {0, 0, 0, 0, 0, 0, 0};
	if (warp < tilesX * tilesY)
	{
		uint32_t px = (warp % tilesX) * 8u + (lane & 7u);
		uint32_t py = (warp / tilesX) * 4u + (lane >> 3);
		bool valid = px < A.width && py < A.height;
		uint32_t sBegin = A.spp_begin, sEnd = A.spp_begin + A.spp_count, sStep = 1;
		if (P.partition == RTB_PART_TILE && P.part_world > 1)
		{
			// TILE_SIZE = 32 (Renderer.h:18), tiles dealt round-robin in row-major order
			uint32_t t32x = (A.width + 31u) >> 5;
			uint32_t tile = (py >> 5) * t32x + (px >> 5);
			valid = valid && ((int)(tile % (uint32_t)P.part_world) == P.part_rank);
		}
		else if (P.partition == RTB_PART_SPP && P.part_world > 1)
		{
			uint32_t w = (uint32_t)P.part_world, r = (uint32_t)P.part_rank;
			sBegin += (r + w - (sBegin % w)) % w;
			sStep = w;
		}
		if (valid)
		{
			uint32_t pixel = py * A.width + px;
			RayD primary = generateRay(S.cam, (float)px + 0.5f, (float)py + 0.5f);
			long long accX = 0, accY = 0, accZ = 0; // fixed point: the same film whatever the schedule
			uint32_t s = sBegin, curS = 0;
			bool alive = false;
			RayD ray = primary;
			V3 T = mk(1.0f, 1.0f, 1.0f), Lo = mk(0.0f, 0.0f, 0.0f);
			int depth = 0;
			bool canHitLight = true;
			for (;;)
			{
				if (!alive)
				{
					if (s >= sEnd) break;
					curS = s;
					s += sStep;
					ray = primary;
					T = mk(1.0f, 1.0f, 1.0f);
					Lo = mk(0.0f, 0.0f, 0.0f);
					depth = 0;
					canHitLight = true;
					alive = true;
				}
				HitD h;
				closestHit<TRAV>(S, ray, P.epsilon, P.cull_rel, h, tl.box, tl.tri);
				tl.closest++;
				bool done = true;
				if (h.id == RTB_MISS_ID)
				{
					if (INTEGRATOR == RTB_INT_PATH || INTEGRATOR == RTB_INT_PATH_MIS || INTEGRATOR == RTB_INT_ALBEDO)
						Lo = Lo + backgroundEval(S, ray.d); // un-weighted (Renderer.h:390)
				}
				else
				{
					ShadeD sd;
					calcShading(S, h.id, h.t, h.alpha, h.beta, 1.0f - (h.alpha + h.beta), ray, sd);
					if (INTEGRATOR == RTB_INT_NORMALS)
					{
						Lo = mk(fabsf(sd.sN.x), fabsf(sd.sN.y), fabsf(sd.sN.z)); // Renderer.h:572-581
					}
					else
					{
						rtb_material m = S.mats[sd.mat];
						if (m.flags & RTB_MAT_LIGHT)
						{
							if (INTEGRATOR == RTB_INT_PATH || INTEGRATOR == RTB_INT_PATH_MIS)
							{
								if (canHitLight) Lo = Lo + (T * mk(m.emission));
							}
							else
								Lo = mk(m.emission);
						}
						else if (INTEGRATOR == RTB_INT_ALBEDO)
						{
							Lo = bsdfEvaluate(S, m, sd, mk(0.0f, 1.0f, 0.0f)); // Renderer.h:558-571
						}
						else
						{
							float4 ua = rngBlock(P.seed, pixel, curS, 2u * (uint32_t)depth);
							V3 direct;
							if (INTEGRATOR == RTB_INT_PATH_MIS)
								direct = computeDirectMIS<TRAV>(S, P, sd, m, ua.x, ua.y, ua.z,
								                                rngBlock(P.seed, pixel, curS, RTB_RNG_MIS_BLOCK + (uint32_t)depth), tl);
							else
								direct = computeDirect<TRAV>(S, P, sd, m, ua.x, ua.y, ua.z, tl);
							if (INTEGRATOR == RTB_INT_DIRECT)
							{
								Lo = direct; // Renderer.h:393-407
							}
							else
							{
								Lo = Lo + (T * direct);
								if (!(depth > P.max_depth))
								{
									float rr = selMin(lum(T), P.rr_cap); // Renderer.h:353
									if (ua.w < rr)
									{
										T = T / rr;
										float4 ub = rngBlock(P.seed, pixel, curS, 2u * (uint32_t)depth + 1u);
										V3 f;
										float pdf;
										V3 wi = bsdfSample(S, m, sd, ub.x, ub.y, ub.z, f, pdf);
										bool spec = (m.flags & RTB_MAT_SPECULAR) != 0;
										if (spec) T = (T * f) / pdf;
										else T = ((T * f) * fabsf(dot(wi, sd.sN))) / pdf;
										ray = mkRay(sd.x + (wi * P.epsilon), wi);
										canHitLight = spec;
										depth++;
										done = false;
									}
								}
							}
						}
					}
				}
				if (done)
				{
					accX += toFixed(Lo.x), accY += toFixed(Lo.y), accZ += toFixed(Lo.z);
					tl.samples++;
					alive = false;
				}
			}
			long long* f = A.accum + (size_t)pixel * 3; // this thread owns the pixel during the launch
			f[0] += accX;
			f[1] += accY;
			f[2] += accZ;
		}
	}
	flushTally(tl, A.counters);
}

// ---------------------------------------------------------------------------------------
// RayTracer::lightTracer (Renderer.h:220-326): one thread per light path, width*height paths per pass.
// lightTrace_init (:262-288): uniform light pick, only area lights emit; position on the triangle, cosine
// direction about its geometric normal (Lights.h:67-82).  The light vertex and every non-specular,
// non-emitting hit are connected to the camera (connectToCamera :233-260) and splatted with the box filter
// (Imaging.h:209-232) — an atomic add into the fixed-point film.  Russian roulette min(Lum(T), 0.9), no depth
// limit (:307-315).  RNG stream 1: block 0 = [pick, pos r1, pos r2, dir r1], block 1 = [dir r2],
// vertex k: block 2 + k = [roulette, bsdf r1, r2, r3].
// ---------------------------------------------------------------------------------------
RTB_DEV V3 mulPoint(const float* m, V3 v) // Core.h:302-309
{
	return mk((v.x * m[0] + v.y * m[1] + v.z * m[2]) + m[3], (v.x * m[4] + v.y * m[5] + v.z * m[6]) + m[7],
	          (v.x * m[8] + v.y * m[9] + v.z * m[10]) + m[11]);
}

template <int TRAV>
RTB_DEV void connectToCamera(const DevScene& S, const rtb_params& P, const rtb_camera_ext& ce, long long* accum, V3 p, V3 n, V3 col,
                             Tally& tl)
{
	// Camera::projectOntoCamera, Scene.h:55-69
	V3 pv = mulPoint(ce.world_to_cam, p), v1 = mulPoint(ce.proj, pv);
	const float* m = ce.proj;
	float w = (m[12] * pv.x) + (m[13] * pv.y) + (m[14] * pv.z) + m[15];
	w = 1.0f / w;
	v1 = v1 * w;
	float x = (v1.x + 1.0f) * 0.5f, y = (v1.y + 1.0f) * 0.5f;
	if (x < 0.0f || x > 1.0f || y < 0.0f || y > 1.0f) return;
	x = x * S.cam.width;
	y = 1.0f - y;
	y = y * S.cam.height;
	V3 origin = mk(S.cam.origin);
	V3 dir = origin - p;
	float dist2 = lengthSq(dir);
	dir = normalize(dir);
	float cs = dot(n, dir);
	float cc = dot(mk(ce.view_dir), -dir);
	if (cs < 0.0f || cc < 0.0f) return;
	float G = (cs * cc) / dist2;
	tl.shadow++;
	if (!sceneVisible<TRAV>(S, p, origin, P.epsilon, P.cull_rel, tl.sbox, tl.stri)) return;
	float We = 1.0f / (ce.afilm * ((cc * cc) * (cc * cc)));
	col = (col * We) * G;
	int px = (int)x, py = (int)y;
	if (px >= 0 && px < (int)S.cam.width && py >= 0 && py < (int)S.cam.height) filmAdd(accum, (uint32_t)py * (uint32_t)S.cam.width + (uint32_t)px, col);
}

// One flat loop per thread: an iteration starts a new light path in the lanes that have none and advances every live
// path by one vertex, so a lane whose path ended does not wait for the longest path of its warp before it takes the next
// one (round 1's path-per-iteration loop ran at 5 of 32 lanes per instruction).  Paths are keyed (path, pass) in the RNG
// and the film is an integer sum: the result does not depend on which lane traces which path, or when.
template <int TRAV>
__global__ void __launch_bounds__(128) k_light_trace(const __grid_constant__ DevScene S, const __grid_constant__ RenderArgs A,
                                                     const __grid_constant__ rtb_camera_ext ce)
{
	const rtb_params& P = A.P;
	Tally tl = {0, 0, 0, 0, 0, 0, 0};
	const unsigned long long perPass = (unsigned long long)A.width * A.height, total = perPass * A.spp_count;
	const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
	unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	bool live = false;
	uint32_t pass = 0, path = 0, k = 0;
	RayD r;
	V3 T = mk(1.0f, 1.0f, 1.0f), Le = mk(0.0f, 0.0f, 0.0f);
	while (live || j < total)
	{
		if (!live)
		{
			// lightTrace_init (Renderer.h:262-288)
			pass = A.spp_begin + (uint32_t)(j / perPass), path = (uint32_t)(j % perPass);
			j += stride;
			tl.samples++;
			if (S.n_lights == 0) continue;
			float4 u0 = rngBlock(P.seed, path, pass, 0u, 1u), u1 = rngBlock(P.seed, path, pass, 1u, 1u);
			float nl = (float)S.n_lights;
			float pmf = 1.0f / nl;
			int li = (int)(nl * u0.x);
			if (li > (int)S.n_lights - 1) li = (int)S.n_lights - 1;
			rtb_light L = S.lights[li];
			if (L.type != RTB_LIGHT_AREA) continue;
			V3 p = trianglePoint(S, L.triangle, u0.y, u0.z);
			float pdfPos = 1.0f / L.area;
			V3 wl = cosineSampleHemisphere(u0.w, u1.x);
			float pdfDir = (wl.z >= 0.0f) ? (wl.z * RTB_INV_PI_F) : 0.0f;
			V3 nL = triangleGNormal(S, L.triangle);
			V3 fu, fv, fw;
			frameFromVector(nL, fu, fv, fw);
			V3 wi = ((fu * wl.x) + (fv * wl.y)) + (fw * wl.z);
			float cosTheta = dot(nL, wi);
			Le = (dot(-wi, nL) < 0.0f) ? mk(L.emission) : mk(0.0f, 0.0f, 0.0f); // AreaLight::evaluate(-wi), Lights.h:41-48
			Le = (Le * cosTheta) / (pmf * pdfDir * pdfPos);
			connectToCamera<TRAV>(S, P, ce, A.accum, p, nL, Le, tl);
			r = mkRay(p, wi);
			T = mk(1.0f, 1.0f, 1.0f);
			k = 0;
			live = true;
		}
		// one vertex of lightTracePath (Renderer.h:290-326)
		live = false;
		HitD h;
		closestHit<TRAV>(S, r, P.epsilon, P.cull_rel, h, tl.box, tl.tri);
		tl.closest++;
		if (h.id == RTB_MISS_ID) continue;
		ShadeD sd;
		calcShading(S, h.id, h.t, h.alpha, h.beta, 1.0f - (h.alpha + h.beta), r, sd);
		rtb_material m = S.mats[sd.mat];
		if (m.flags & (RTB_MAT_LIGHT | RTB_MAT_SPECULAR)) continue;
		connectToCamera<TRAV>(S, P, ce, A.accum, sd.x, sd.sN, (T * bsdfEvaluate(S, m, sd, mk(0.0f, 1.0f, 0.0f))) * Le, tl);
		float4 uk = rngBlock(P.seed, path, pass, 2u + k, 1u);
		float rr = selMin(lum(T), P.rr_cap);
		if (!(uk.x < rr)) continue;
		T = T / rr;
		V3 f;
		float pdf;
		V3 wi2 = bsdfSample(S, m, sd, uk.y, uk.z, uk.w, f, pdf);
		T = ((T * f) * fabsf(dot(wi2, sd.sN))) / pdf;
		r = mkRay(sd.x + (wi2 * P.epsilon), wi2);
		k++;
		live = k < 100000u;
	}
	flushTally(tl, A.counters);
}

// ---------------------------------------------------------------------------------------
// RayTracer::instantRadiosity (Renderer.h:82-218).  k_ir_vpls = traceVPLs + VPLTracePath (:159-218): one
// thread per light path (MAX_VPL = 50 of them), each writing its VPLs into its own segment so that the list
// order (path, vertex) does not depend on scheduling.  k_ir_gather = renderBlockinstantRadiosity +
// computeVPLsContribution (:82-101, 124-158): one thread per pixel (a warp = an 8x4 tile) sums all VPLs, one
// visibility ray each — the lanes of a warp shoot at the same VPL.  RNG stream 2, light-tracer block layout.
// ---------------------------------------------------------------------------------------
struct VplD
{
	float4 x;  // position, -
	float4 n;  // normal, -
	float4 le; // Le, -
};
#define RTB_VPL_SEGMENT 256 /* VPLs one light path may store: survival <= 0.9 per vertex */

template <int TRAV>
__global__ void __launch_bounds__(64) k_ir_vpls(const __grid_constant__ DevScene S, const __grid_constant__ RenderArgs A, uint32_t pass,
                                                uint32_t nPaths, VplD* vpls, uint32_t* counts)
{
	const rtb_params& P = A.P;
	Tally tl = {0, 0, 0, 0, 0, 0, 0};
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nPaths; i += gridDim.x * blockDim.x)
	{
		VplD* out = vpls + (size_t)i * RTB_VPL_SEGMENT;
		uint32_t n = 0;
		counts[i] = 0;
		if (S.n_lights == 0) continue;
		float4 u0 = rngBlock(P.seed, i, pass, 0u, 2u), u1 = rngBlock(P.seed, i, pass, 1u, 2u);
		float nl = (float)S.n_lights;
		float pmf = 1.0f / nl;
		int li = (int)(nl * u0.x);
		if (li > (int)S.n_lights - 1) li = (int)S.n_lights - 1;
		rtb_light L = S.lights[li];
		if (L.type != RTB_LIGHT_AREA) continue;
		V3 p = trianglePoint(S, L.triangle, u0.y, u0.z);
		float pdfPos = 1.0f / L.area;
		V3 wl = cosineSampleHemisphere(u0.w, u1.x);
		V3 nL = triangleGNormal(S, L.triangle);
		V3 fu, fv, fw;
		frameFromVector(nL, fu, fv, fw);
		V3 wi = ((fu * wl.x) + (fv * wl.y)) + (fw * wl.z);
		V3 Lev = (dot(-wi, nL) < 0.0f) ? mk(L.emission) : mk(0.0f, 0.0f, 0.0f);
		float norm = pmf * pdfPos * (float)nPaths;
		V3 l0 = Lev / norm;
		out[n].x = make_float4(p.x, p.y, p.z, 0.0f), out[n].n = make_float4(nL.x, nL.y, nL.z, 0.0f), out[n].le = make_float4(l0.x, l0.y, l0.z, 0.0f);
		n++;
		V3 Le = (Lev * dot(wi, nL)) / norm;
		RayD r = mkRay(p, wi);
		V3 T = mk(1.0f, 1.0f, 1.0f);
		for (uint32_t kk = 0; kk < 100000u; kk++)
		{
			HitD h;
			closestHit<TRAV>(S, r, P.epsilon, P.cull_rel, h, tl.box, tl.tri);
			tl.closest++;
			if (h.id == RTB_MISS_ID) break;
			ShadeD sd;
			calcShading(S, h.id, h.t, h.alpha, h.beta, 1.0f - (h.alpha + h.beta), r, sd);
			rtb_material m = S.mats[sd.mat];
			if (!(m.flags & (RTB_MAT_LIGHT | RTB_MAT_SPECULAR)) && n < RTB_VPL_SEGMENT)
			{
				V3 le = ((T * Le) * bsdfEvaluate(S, m, sd, -r.d)) * fabsf(dot(-r.d, sd.sN));
				out[n].x = make_float4(sd.x.x, sd.x.y, sd.x.z, 0.0f), out[n].n = make_float4(sd.sN.x, sd.sN.y, sd.sN.z, 0.0f);
				out[n].le = make_float4(le.x, le.y, le.z, 0.0f);
				n++;
			}
			float4 uk = rngBlock(P.seed, i, pass, 2u + kk, 2u);
			float rr = selMin(lum(T), P.rr_cap);
			if (!(uk.x < rr)) break;
			T = T / rr;
			V3 f;
			float pdf;
			V3 wi2 = bsdfSample(S, m, sd, uk.y, uk.z, uk.w, f, pdf);
			T = ((T * f) * fabsf(dot(wi2, sd.sN))) / pdf;
			r = mkRay(sd.x + (wi2 * P.epsilon), wi2);
		}
		counts[i] = n;
	}
	flushTally(tl, A.counters);
}

template <int TRAV>
__global__ void __launch_bounds__(128) k_ir_gather(const __grid_constant__ DevScene S, const __grid_constant__ RenderArgs A, uint32_t nPaths,
                                                   const VplD* __restrict__ vpls, const uint32_t* __restrict__ counts)
{
	const rtb_params& P = A.P;
	Tally tl = {0, 0, 0, 0, 0, 0, 0};
	uint32_t tilesX = (A.width + 7u) >> 3, tilesY = (A.height + 3u) >> 2;
	uint32_t lane = threadIdx.x & 31u;
	for (uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < tilesX * tilesY; w += (gridDim.x * blockDim.x) >> 5)
	{
		uint32_t px = (w % tilesX) * 8u + (lane & 7u), py = (w / tilesX) * 4u + (lane >> 3);
		if (px >= A.width || py >= A.height) continue;
		RayD r = generateRay(S.cam, (float)px + 0.5f, (float)py + 0.5f);
		HitD h;
		closestHit<TRAV>(S, r, P.epsilon, P.cull_rel, h, tl.box, tl.tri);
		tl.closest++;
		tl.samples++;
		if (h.id == RTB_MISS_ID) continue;
		ShadeD sd;
		calcShading(S, h.id, h.t, h.alpha, h.beta, 1.0f - (h.alpha + h.beta), r, sd);
		rtb_material m = S.mats[sd.mat];
		if (m.flags & (RTB_MAT_LIGHT | RTB_MAT_SPECULAR)) continue;
		V3 f = bsdfEvaluate(S, m, sd, mk(0.0f, 1.0f, 0.0f));
		V3 col = mk(0.0f, 0.0f, 0.0f);
		for (uint32_t i = 0; i < nPaths; i++)
		{
			uint32_t cnt = __ldg(counts + i);
			const VplD* seg = vpls + (size_t)i * RTB_VPL_SEGMENT;
			for (uint32_t k = 0; k < cnt; k++)
			{
				V3 vx = mk(__ldg(&seg[k].x)), vn = mk(__ldg(&seg[k].n));
				V3 d = vx - sd.x;
				float dist2 = lengthSq(d);
				if (dist2 < 1e-4f) continue;
				d = normalize(d);
				float cv = dot(vn, -d), cx = dot(sd.sN, d);
				if (cv <= 0.0f || cx <= 0.0f) continue;
				float G = (cv * cx) / dist2;
				tl.shadow++;
				if (!sceneVisible<TRAV>(S, sd.x, vx, P.epsilon, P.cull_rel, tl.sbox, tl.stri)) continue;
				col = col + ((mk(__ldg(&seg[k].le)) * f) * G);
			}
		}
		long long* a = A.accum + ((size_t)py * A.width + px) * 3; // this thread owns the pixel during the launch
		a[0] += toFixed(col.x), a[1] += toFixed(col.y), a[2] += toFixed(col.z);
	}
	flushTally(tl, A.counters);
}

// ---------------------------------------------------------------------------------------
// Parity kernels
// ---------------------------------------------------------------------------------------
template <int TRAV>
__global__ void __launch_bounds__(128) k_primary(const __grid_constant__ DevScene S, float eps, float cull,
                                                 uint32_t width, uint32_t height, uint32_t* ids, float* ts,
                                                 rtb_ray* rays, unsigned long long* counters)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	Tally tl = {0, 0, 0, 0, 0, 0, 0};
	if (i < width * height)
	{
		uint32_t px = i % width, py = i / width;
		RayD r = generateRay(S.cam, (float)px + 0.5f, (float)py + 0.5f);
		HitD h;
		closestHit<TRAV>(S, r, eps, cull, h, tl.box, tl.tri);
		tl.closest++;
		if (ids) ids[i] = h.id;
		if (ts) ts[i] = h.t;
		if (rays)
		{
			rtb_ray o;
			o.o[0] = r.o.x, o.o[1] = r.o.y, o.o[2] = r.o.z, o.tmax = FLT_MAX;
			o.d[0] = r.d.x, o.d[1] = r.d.y, o.d[2] = r.d.z, o.pad_ = 0.0f;
			rays[i] = o;
		}
	}
	flushTally(tl, counters);
}

template <int TRAV>
__global__ void __launch_bounds__(128) k_trace(const __grid_constant__ DevScene S, float eps, float cull, int anyHit,
                                               const rtb_ray* __restrict__ rays, uint64_t n, rtb_hit* hits,
                                               unsigned long long* counters)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	Tally tl = {0, 0, 0, 0, 0, 0, 0};
	if (i < n)
	{
		rtb_ray q = rays[i];
		RayD r = mkRay(mk(q.o), mk(q.d));
		rtb_hit o;
		if (anyHit)
		{
			bool vis = anyVisible<TRAV>(S, r, eps, q.tmax, cull, tl.sbox, tl.stri);
			tl.shadow++;
			o.id = vis ? 0u : 1u;
			o.t = o.alpha = o.beta = o.gamma = 0.0f;
		}
		else
		{
			HitD h;
			closestHit<TRAV>(S, r, eps, cull, h, tl.box, tl.tri);
			tl.closest++;
			o.id = h.id, o.t = h.t;
			if (h.id == RTB_MISS_ID) o.alpha = o.beta = o.gamma = 0.0f;
			else o.alpha = h.alpha, o.beta = h.beta, o.gamma = 1.0f - (h.alpha + h.beta);
		}
		hits[i] = o;
	}
	flushTally(tl, counters);
}

template <int TRAV>
__global__ void __launch_bounds__(128) k_visible(const __grid_constant__ DevScene S, float eps, float cull,
                                                 const float* __restrict__ p1p2, uint64_t n, uint8_t* out,
                                                 unsigned long long* counters)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	Tally tl = {0, 0, 0, 0, 0, 0, 0};
	if (i < n)
	{
		const float* p = p1p2 + i * 6;
		bool vis = sceneVisible<TRAV>(S, mk(p), mk(p + 3), eps, cull, tl.sbox, tl.stri);
		tl.shadow++;
		out[i] = vis ? 1 : 0;
	}
	flushTally(tl, counters);
}

RTB_DEV void storeShading(const ShadeD& sd, rtb_shading& o)
{
	o.x[0] = sd.x.x, o.x[1] = sd.x.y, o.x[2] = sd.x.z;
	o.wo[0] = sd.wo.x, o.wo[1] = sd.wo.y, o.wo[2] = sd.wo.z;
	o.s_normal[0] = sd.sN.x, o.s_normal[1] = sd.sN.y, o.s_normal[2] = sd.sN.z;
	o.g_normal[0] = sd.gN.x, o.g_normal[1] = sd.gN.y, o.g_normal[2] = sd.gN.z;
	o.tu = sd.tu, o.tv = sd.tv;
	o.frame_u[0] = sd.fu.x, o.frame_u[1] = sd.fu.y, o.frame_u[2] = sd.fu.z;
	o.frame_v[0] = sd.fv.x, o.frame_v[1] = sd.fv.y, o.frame_v[2] = sd.fv.z;
	o.frame_w[0] = sd.fw.x, o.frame_w[1] = sd.fw.y, o.frame_w[2] = sd.fw.z;
	o.t = sd.t;
	o.material = sd.mat;
}
RTB_DEV void loadShading(const rtb_shading& o, ShadeD& sd)
{
	sd.x = mk(o.x), sd.wo = mk(o.wo), sd.sN = mk(o.s_normal), sd.gN = mk(o.g_normal);
	sd.tu = o.tu, sd.tv = o.tv;
	sd.fu = mk(o.frame_u), sd.fv = mk(o.frame_v), sd.fw = mk(o.frame_w);
	sd.t = o.t, sd.mat = o.material;
}

__global__ void __launch_bounds__(128) k_shading(const __grid_constant__ DevScene S, const rtb_ray* __restrict__ rays,
                                                 const rtb_hit* __restrict__ hits, uint64_t n, rtb_shading* out)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	rtb_ray q = rays[i];
	rtb_hit h = hits[i];
	RayD r = mkRay(mk(q.o), mk(q.d));
	rtb_shading o;
	memset(&o, 0, sizeof(o));
	if (h.t < FLT_MAX && h.id < S.n_tris)
	{
		ShadeD sd;
		calcShading(S, h.id, h.t, h.alpha, h.beta, h.gamma, r, sd);
		storeShading(sd, o);
	}
	else
	{
		// Scene.h:197-201: only wo and t are set on a miss
		V3 wo = -r.d;
		o.wo[0] = wo.x, o.wo[1] = wo.y, o.wo[2] = wo.z;
		o.t = h.t;
		o.material = -1;
	}
	out[i] = o;
}

__global__ void __launch_bounds__(128) k_eval_bsdf(const __grid_constant__ DevScene S,
                                                   const rtb_shading* __restrict__ sds, const float* __restrict__ wi,
                                                   const float* __restrict__ u, uint64_t n, float* eval, float* pdf,
                                                   float* sWi, float* sF, float* sPdf)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	ShadeD sd;
	loadShading(sds[i], sd);
	if (sd.mat < 0 || (uint32_t)sd.mat >= S.n_mats) return;
	rtb_material m = S.mats[sd.mat];
	V3 w = mk(wi + i * 3);
	if (eval)
	{
		V3 e = bsdfEvaluate(S, m, sd, w);
		eval[i * 3] = e.x, eval[i * 3 + 1] = e.y, eval[i * 3 + 2] = e.z;
	}
	if (pdf) pdf[i] = bsdfPdf(m, sd, w);
	if (sWi || sF || sPdf)
	{
		V3 f;
		float p;
		V3 d = bsdfSample(S, m, sd, u[i * 3], u[i * 3 + 1], u[i * 3 + 2], f, p);
		if (sWi) sWi[i * 3] = d.x, sWi[i * 3 + 1] = d.y, sWi[i * 3 + 2] = d.z;
		if (sF) sF[i * 3] = f.x, sF[i * 3 + 1] = f.y, sF[i * 3 + 2] = f.z;
		if (sPdf) sPdf[i] = p;
	}
}

// Light::sample / Light::evaluate (RTBase/Lights.h:35-48, 89-100, 143-166); the direction sample of an
// environment map follows rtb_params.sampling (sampleNonAreaLight).
__global__ void __launch_bounds__(128) k_eval_light(const __grid_constant__ DevScene S, const __grid_constant__ rtb_params P,
                                                    const int32_t* __restrict__ light, const float* __restrict__ wi,
                                                    const float* __restrict__ u, uint64_t n, float* pOrWi,
                                                    float* emitted, float* pdf, float* eval)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	int li = light[i];
	if (li < 0 || (uint32_t)li >= S.n_lights) return;
	rtb_light L = S.lights[li];
	V3 w = mk(wi + i * 3);
	V3 p, e, ev;
	float pd;
	if (L.type == RTB_LIGHT_AREA)
	{
		p = trianglePoint(S, L.triangle, u[i * 2], u[i * 2 + 1]);
		pd = 1.0f / L.area;
		e = mk(L.emission);
		ev = (dot(w, triangleGNormal(S, L.triangle)) < 0.0f) ? mk(L.emission) : mk(0.0f, 0.0f, 0.0f);
	}
	else
	{
		if (!sampleNonAreaLight(S, P, L, u[i * 2], u[i * 2 + 1], p, pd, e)) pd = 0.0f, e = mk(0.0f, 0.0f, 0.0f);
		ev = (L.type == RTB_LIGHT_ENVMAP) ? envLookup(S, L.tex, w) : mk(L.emission);
	}
	if (pOrWi) pOrWi[i * 3] = p.x, pOrWi[i * 3 + 1] = p.y, pOrWi[i * 3 + 2] = p.z;
	if (emitted) emitted[i * 3] = e.x, emitted[i * 3 + 1] = e.y, emitted[i * 3 + 2] = e.z;
	if (pdf) pdf[i] = pd;
	if (eval) eval[i * 3] = ev.x, eval[i * 3 + 1] = ev.y, eval[i * 3 + 2] = ev.z;
}

__global__ void k_rng(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t n, float* out)
{
	uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b * 4 >= n) return;
	float4 r = rngBlock(seed, pixel, sample, b);
	float v[4] = {r.x, r.y, r.z, r.w};
	for (int k = 0; k < 4; k++)
		if (b * 4 + k < n) out[b * 4 + k] = v[k];
}

// Film::tonemap (RTBase/Imaging.h:233-242): pixel = film * exposure / SPP;
// c8 = min(powf(max(c, 0), 1/2.2f) * 255, 255).
__global__ void __launch_bounds__(256) k_tonemap(const float* __restrict__ film, uint32_t nPixels, float spp,
                                                 float exposure, uint8_t* out)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nPixels * 3) return;
	float c = (film[i] * exposure) / spp;
	float v = powf(stdMax(c, 0.0f), 1.0f / 2.2f) * 255.0f;
	out[i] = (uint8_t)stdMin(v, 255.0f);
}

// GaussianFilter splat (RTBase/Imaging.h:155-187, 209-232) applied to the accumulated box
// film.  Samples sit at pixel centres and filter(j, i) receives integer tap offsets, so the
// splat of every sample of source pixel p is the same (2*size+1)^2 stencil normalised by
// the sum of p's in-bounds taps: linear in the per-pixel sums.
__global__ void __launch_bounds__(256) k_gaussian(const float* __restrict__ box, float* out, int width, int height,
                                                  int size, float radius, float alpha)
{
	int q = blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= width * height) return;
	int qx = q % width, qy = q / width;
	float g[5];
	// GaussianFilter::Gaussian (Imaging.h:176-180): exp() on floats is the float overload
	float er = expf(-alpha * (radius * radius));
	for (int d = -size; d <= size; d++) g[d + size] = expf(-alpha * ((float)d * (float)d)) - er;
	float r = 0.0f, gg = 0.0f, b = 0.0f;
	for (int i = -size; i <= size; i++)
	{
		int sy = qy - i; // source row whose tap offset i lands on qy
		if (sy < 0 || sy >= height) continue;
		for (int j = -size; j <= size; j++)
		{
			int sx = qx - j;
			if (sx < 0 || sx >= width) continue;
			// total weight of source (sx, sy): its in-bounds taps, in the reference's loop order
			float total = 0.0f;
			for (int ii = -size; ii <= size; ii++)
				for (int jj = -size; jj <= size; jj++)
				{
					int tx = sx + jj, ty = sy + ii;
					if (tx >= 0 && tx < width && ty >= 0 && ty < height) total += g[jj + size] * g[ii + size];
				}
			float w = g[j + size] * g[i + size];
			const float* s = box + ((size_t)sy * width + sx) * 3;
			r += (s[0] * w) / total;
			gg += (s[1] * w) / total;
			b += (s[2] * w) / total;
		}
	}
	out[(size_t)q * 3] = r, out[(size_t)q * 3 + 1] = gg, out[(size_t)q * 3 + 2] = b;
}
