// rtb_dev_cw.cuh — traversal of the CW tree (RTB_TRAV_CW; layout and the argument why it returns the reference's
// hits: rtb_cwbvh.hpp).  Scene::traverse / BVHNode::traverse (RTBase/Scene.h:107-130, Geometry.h:399-434) and
// Scene::visible / traverseVisible (Scene.h:161-169, Geometry.h:435-462) for non-degenerate rays.
//
// profiles/r02_base_bathroom_summary.md: on the heavy scenes the binary FAST tree is bound by the L1 data pipe
// (78 % of the LSU wavefront peak in k_wf_extend: four divergent LDG.128 per visited 64-byte node, ~36 nodes per
// ray, plus a 768-byte local-memory stack per thread).  The CW tree cuts the node fetches per ray ~3x (80-byte
// nodes holding eight children), the stack to one 8-byte entry per LEVEL (a node group: base index + hit bits,
// no per-entry t), and its top levels are read from SHARED MEMORY, where the kernels stage them with one bulk
// copy (cp.async.bulk + mbarrier) per block.
//
// Per visited node: decode (biased exponents -> grid steps, origin) once, then per child six FFMA on bytes
// turned into floats by one PRMT each (0x47000000 | q << 8 = 32768 + q), min/max, ONE multiply for the
// conservative slack and a compare; the eight hit bits are permuted by XOR with the ray's octant so that
// "highest set bit" = nearest child (no sorting network).  A hit LEAF slot only admits the leaf's exact box
// test (reference arithmetic, slabTestNoNaN) and then its triangles (leafClosest / leafOccludes, unchanged).
#pragma once
#include "rtb_dev_scene.cuh"

#define RTB_CW_STACK 40 /* node groups: one per level; rtb_upload_scene rejects deeper CW trees (falls back to FAST) */
// tmin * RTB_CW_SLACK <= tmax: relative slack 2^-18 between entry and exit (the quantised boxes are padded by a
// full grid step as well).  Covers the <= 2^-22 relative difference between the FMA form and the reference's
// (plane - o) * invDir, see rtb_cwbvh.hpp.
#define RTB_CW_SLACK 0.99999619f

struct CwView
{
	const float4* nodes;   // global, 5 x float4 per node
	const float4* sNodes;  // the first nShared nodes in shared memory (generic pointer), or nullptr
	const float4* leaves;  // 2 x float4 per leaf record (global, or shared when the whole array was staged)
	uint32_t nShared;
};

template <bool ANYHIT>
struct LaneCw
{
	RayD r;
	float cullT, bestT;
	uint32_t bestId;
	float bestU, bestV;
	uint2 ng; // node group: x = index of the first internal child, y = hits (octant-permuted) << 24 | imask
	uint2 lg; // leaf group: x = index of the first leaf record,   y = hits (octant-permuted) | lmask << 8
	uint32_t q; // bit k set: direction component k is positive
	int sp;
};

// The CW test multiplies by grid step * invDir and by 32768: reciprocals beyond 1e28 (direction components below
// 1e-28: axis-parallel rays) could overflow, so such rays take the reference's own tree like the 0 * inf ones.
RTB_DEV bool cwRayDegenerate(const RayD& r)
{
	return !(fabsf(r.inv.x) <= 1e28f) || !(fabsf(r.inv.y) <= 1e28f) || !(fabsf(r.inv.z) <= 1e28f) || rayHasNaN(r);
}

template <bool ANYHIT>
RTB_DEV void cwSetBest(LaneCw<ANYHIT>& t, float best, float cullRel)
{
	t.bestT = best;
	t.cullT = best + fabsf(best) * (2.0f * cullRel);
}

template <bool ANYHIT>
RTB_DEV void cwStart(LaneCw<ANYHIT>& t, const RayD& r, float best, float cullRel)
{
	t.r = r;
	cwSetBest<ANYHIT>(t, best, cullRel);
	t.bestId = RTB_MISS_ID, t.bestU = t.bestV = 0.0f;
	t.q = (r.d.x >= 0.0f ? 1u : 0u) | (r.d.y >= 0.0f ? 2u : 0u) | (r.d.z >= 0.0f ? 4u : 0u);
	// a virtual parent whose only child (slot 0, internal) is the root: node 0
	t.ng = make_uint2(0u, (1u << (24u + t.q)) | 1u);
	t.lg = make_uint2(0u, 0u);
	t.sp = 0;
}

template <bool ANYHIT>
RTB_DEV bool cwIdle(const LaneCw<ANYHIT>& t)
{
	return !(t.lg.y & 0xFFu) && !(t.ng.y & 0xFF000000u) && t.sp == 0;
}

// byte j of w as the float 32768 + byte
#define RTB_CW_F(w, j) __uint_as_float(__byte_perm((w), 0x47000000u, 0x7404u | ((j) << 4)))

// Visit the nearest unvisited internal child of the current node group: push the remaining siblings, fetch the
// child's 80 bytes, test its eight children.
template <bool ANYHIT>
RTB_DEV void cwNodeStep(const CwView& V, LaneCw<ANYHIT>& t, uint2* stack, uint32_t& nBox)
{
	const uint32_t pos = 31u - (uint32_t)__clz((int)(t.ng.y & 0xFF000000u));
	t.ng.y &= ~(1u << pos);
	if (t.ng.y & 0xFF000000u) stack[t.sp++] = t.ng;
	const uint32_t slot = (pos - 24u) ^ t.q;
	const uint32_t idx = t.ng.x + (uint32_t)__popc(t.ng.y & 0xFFu & ((1u << slot) - 1u));
	const float4* nd = (idx < V.nShared) ? V.sNodes + (size_t)idx * 5 : V.nodes + (size_t)idx * 5;
	const float4 n0 = nd[0], n1 = nd[1], n2 = nd[2], n3 = nd[3], n4 = nd[4];
	const uint32_t meta = __float_as_uint(n0.w);
	const uint32_t imask = meta >> 24, lmask = __float_as_uint(n1.z) & 0xFFu;
	// t = (p + q * step - o) * inv = (32768 + q) * a + b with a = step * inv, b = (p - o) * inv - 32768 a
	const float ax = __uint_as_float((meta & 0xFFu) << 23) * t.r.inv.x;
	const float ay = __uint_as_float(((meta >> 8) & 0xFFu) << 23) * t.r.inv.y;
	const float az = __uint_as_float(((meta >> 16) & 0xFFu) << 23) * t.r.inv.z;
	const float bx = fmaf(-32768.0f, ax, (n0.x - t.r.o.x) * t.r.inv.x);
	const float by = fmaf(-32768.0f, ay, (n0.y - t.r.o.y) * t.r.inv.y);
	const float bz = fmaf(-32768.0f, az, (n0.z - t.r.o.z) * t.r.inv.z);
	const bool px = (t.q & 1u) != 0u, py = (t.q & 2u) != 0u, pz = (t.q & 4u) != 0u;
	uint32_t hs = 0u;
#pragma unroll
	for (int half = 0; half < 2; half++)
	{
		const uint32_t lox = __float_as_uint(half ? n2.y : n2.x), loy = __float_as_uint(half ? n2.w : n2.z);
		const uint32_t loz = __float_as_uint(half ? n3.y : n3.x), hix = __float_as_uint(half ? n3.w : n3.z);
		const uint32_t hiy = __float_as_uint(half ? n4.y : n4.x), hiz = __float_as_uint(half ? n4.w : n4.z);
		const uint32_t nx = px ? lox : hix, fx = px ? hix : lox;
		const uint32_t ny = py ? loy : hiy, fy = py ? hiy : loy;
		const uint32_t nz = pz ? loz : hiz, fz = pz ? hiz : loz;
#pragma unroll
		for (int j = 0; j < 4; j++)
		{
			const float tnx = fmaf(RTB_CW_F(nx, j), ax, bx), tny = fmaf(RTB_CW_F(ny, j), ay, by), tnz = fmaf(RTB_CW_F(nz, j), az, bz);
			const float tfx = fmaf(RTB_CW_F(fx, j), ax, bx), tfy = fmaf(RTB_CW_F(fy, j), ay, by), tfz = fmaf(RTB_CW_F(fz, j), az, bz);
			const float tmin = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.0f));
			const float tmax = fminf(fminf(tfx, tfy), fminf(tfz, t.cullT));
			if (tmin * RTB_CW_SLACK <= tmax) hs |= 1u << (half * 4 + j);
		}
	}
	nBox += (uint32_t)__popc(imask | lmask);
	// internal hits in bits 0..7, leaf hits in bits 8..15; permute each byte: bit s -> bit s xor q
	uint32_t x = (hs & imask) | ((hs & lmask) << 8);
	if (t.q & 1u) x = ((x & 0x5555u) << 1) | ((x >> 1) & 0x5555u);
	if (t.q & 2u) x = ((x & 0x3333u) << 2) | ((x >> 2) & 0x3333u);
	if (t.q & 4u) x = ((x & 0x0F0Fu) << 4) | ((x >> 4) & 0x0F0Fu);
	t.ng = make_uint2(__float_as_uint(n1.x), ((x & 0xFFu) << 24) | imask);
	t.lg = make_uint2(__float_as_uint(n1.y), (x >> 8) | (lmask << 8));
}

// Test the nearest untested leaf of the current leaf group: the reference's exact leaf box (Geometry.h:173-184), then
// its triangles.  Any hit: returns true when the segment is occluded.
template <bool ANYHIT>
RTB_DEV bool cwLeafStep(const DevScene& S, const CwView& V, LaneCw<ANYHIT>& t, float eps, float cullRel, uint32_t& nBox, uint32_t& nTri)
{
	const uint32_t pos = 31u - (uint32_t)__clz((int)(t.lg.y & 0xFFu));
	t.lg.y &= ~(1u << pos);
	const uint32_t slot = pos ^ t.q;
	const uint32_t idx = t.lg.x + (uint32_t)__popc((t.lg.y >> 8) & 0xFFu & ((1u << slot) - 1u));
	const float4* lf = V.leaves + (size_t)idx * 2;
	const float4 A = lf[0], B = lf[1];
	float te;
	nBox++;
	if (!slabTestNoNaN(A.x, A.y, A.z, B.x, B.y, B.z, t.r, te)) return false;
	if (travCull<ANYHIT>(te, t.cullT)) return false;
	const int32_t ref = ~(int32_t)__float_as_uint(A.w);
	if (ANYHIT) return leafOccludes(S, ref, t.r, eps, t.bestT, nTri);
	HitD h;
	h.id = t.bestId, h.t = t.bestT, h.alpha = t.bestU, h.beta = t.bestV;
	leafClosest(S, ref, t.r, eps, h, nTri);
	t.bestId = h.id, t.bestU = h.alpha, t.bestV = h.beta;
	cwSetBest<ANYHIT>(t, h.t, cullRel);
	return false;
}

template <bool ANYHIT>
RTB_DEV void cwPop(LaneCw<ANYHIT>& t, const uint2* stack)
{
	if (!(t.lg.y & 0xFFu) && !(t.ng.y & 0xFF000000u) && t.sp > 0) t.ng = stack[--t.sp];
}

RTB_DEV CwView cwGlobalView(const DevScene& S)
{
	CwView V;
	V.nodes = S.cwnodes, V.sNodes = nullptr, V.leaves = S.cwleaves, V.nShared = 0u;
	return V;
}

// One thread runs one ray to the end (parity entry points, the simple shadow stage, megakernel).
RTB_DEV void closestCw(const DevScene& S, const CwView& V, const RayD& r, float eps, float cullRel, HitD& h, uint32_t& nBox, uint32_t& nTri)
{
	if (cwRayDegenerate(r))
	{
		closestExact(S, r, eps, h, nBox, nTri);
		return;
	}
	uint2 stack[RTB_CW_STACK];
	LaneCw<false> t;
	cwStart<false>(t, r, FLT_MAX, cullRel);
	for (;;)
	{
		if (t.lg.y & 0xFFu) cwLeafStep<false>(S, V, t, eps, cullRel, nBox, nTri);
		else if (t.ng.y & 0xFF000000u) cwNodeStep<false>(V, t, stack, nBox);
		else if (t.sp > 0) t.ng = stack[--t.sp];
		else break;
	}
	h.id = t.bestId, h.t = t.bestT, h.alpha = t.bestU, h.beta = t.bestV;
}

RTB_DEV bool visibleCw(const DevScene& S, const CwView& V, const RayD& r, float eps, float maxT, float cullRel, uint32_t& nBox, uint32_t& nTri)
{
	if (cwRayDegenerate(r)) return visibleExact(S, r, eps, maxT, nBox, nTri);
	uint2 stack[RTB_CW_STACK];
	LaneCw<true> t;
	cwStart<true>(t, r, maxT, cullRel);
	for (;;)
	{
		if (t.lg.y & 0xFFu)
		{
			if (cwLeafStep<true>(S, V, t, eps, cullRel, nBox, nTri)) return false;
		}
		else if (t.ng.y & 0xFF000000u) cwNodeStep<true>(V, t, stack, nBox);
		else if (t.sp > 0) t.ng = stack[--t.sp];
		else return true;
	}
}

RTB_DEV void closestCwGlobal(const DevScene& S, const RayD& r, float eps, float cullRel, HitD& h, uint32_t& nBox, uint32_t& nTri)
{
	closestCw(S, cwGlobalView(S), r, eps, cullRel, h, nBox, nTri);
}
RTB_DEV bool visibleCwGlobal(const DevScene& S, const RayD& r, float eps, float maxT, float cullRel, uint32_t& nBox, uint32_t& nTri)
{
	return visibleCw(S, cwGlobalView(S), r, eps, maxT, cullRel, nBox, nTri);
}

// ---------------------------------------------------------------------------------------
// Shared-memory staging of the top of the tree: ONE bulk copy (cp.async.bulk, the 1-D TMA path: SASS UBLKCP)
// per block, completion through an mbarrier.  `bytes` is a multiple of 16.
// ---------------------------------------------------------------------------------------
RTB_DEV void cwStageBulk(void* smemDst0, const void* gmemSrc0, uint32_t bytes0, void* smemDst1, const void* gmemSrc1, uint32_t bytes1,
                            unsigned long long* mbar)
{
	const uint32_t bar = (uint32_t)__cvta_generic_to_shared(mbar);
	if (threadIdx.x == 0)
	{
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	if (threadIdx.x == 0)
	{
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes0 + bytes1) : "memory");
		for (int seg = 0; seg < 2; seg++)
		{
			uint32_t dst = (uint32_t)__cvta_generic_to_shared(seg ? smemDst1 : smemDst0);
			const char* src = (const char*)(seg ? gmemSrc1 : gmemSrc0);
			uint32_t left = seg ? bytes1 : bytes0;
			while (left)
			{
				uint32_t n = left > 32768u ? 32768u : left;
				asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(n),
				             "r"(bar)
				             : "memory");
				dst += n, src += n, left -= n;
			}
		}
	}
	// every thread waits for phase 0 of the barrier
	uint32_t ok = 0;
	while (!ok)
	{
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar) : "memory");
	}
}
