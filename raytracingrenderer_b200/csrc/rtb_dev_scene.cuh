// rtb_dev_scene.cuh — device-side scene view, ray/box/triangle tests, traversal, shading
// data, textures, BSDFs, lights and the counter-based RNG.  Each function cites the
// reference code it restates; arithmetic order follows rtb_dev_math.cuh.
#pragma once
#include "../../include/rtb.h"
#include "rtb_dev_math.cuh"

#ifndef RTB_LDG256
#define RTB_LDG256 1
#endif
#ifndef RTB_TRI256
#define RTB_TRI256 0 /* triangle records: [n, d] alone decides most tests, so one LDG.128 + three more measured 1.3 ... 2.5 % faster than 2 x 256 (profiles/r02_ldg256.txt) */
#endif
#define RTB_INTERIOR 0xFFFFFFFFu
#define RTB_MISS_ID 0xFFFFFFFFu
#define RTB_PI_F 3.14159274101257324f   /* (float)M_PI */
#define RTB_PI_D 3.14159265358979323846 /* M_PI */
#define RTB_2PI_F 6.28318548202514648f  /* (float)(2 M_PI) */
#define RTB_INV_PI_F 0.318309873342514038f
#define RTB_INV_2PI_F 0.159154936671257019f

// Device view of an uploaded scene; passed to kernels by value (__grid_constant__).
struct DevScene
{
	rtb_camera cam;
	// EXACT tree: the reference's nodes in pre-order, 2 x float4 per node:
	//   [2i]   = bmin.xyz, bits(skip)   skip = index of the first node after i's subtree
	//   [2i+1] = bmax.xyz, bits(leaf)   leaf = RTB_INTERIOR, or (start << 2) | count
	const float4* xnodes;
	// FAST tree: binary tree over the reference's leaves, 4 x float4 per node:
	//   [4i]   = c0.min.x c0.max.x c0.min.y c0.max.y
	//   [4i+1] = c1.min.x c1.max.x c1.min.y c1.max.y
	//   [4i+2] = c0.min.z c0.max.z c1.min.z c1.max.z
	//   [4i+3] = bits(child0) bits(child1) - -     child >= 0: node index; < 0: ~((start<<2)|count)
	const float4* fnodes;
	// WIDE tree: FAST collapsed to 4 children, 8 x float4 per node (layout: rtb_accel.hpp WideBuilder)
	const float4* wnodes;
	// CW tree: FAST collapsed to 8 children with 8-bit quantised boxes, 5 x float4 per node, breadth-first; exact leaf
	// boxes as 2 x float4 per leaf record (layout: rtb_cwbvh.hpp).  cw_valid = 0: single-leaf / empty / too deep -> FAST
	const float4* cwnodes;
	const float4* cwleaves;
	uint32_t n_cwnodes, n_cwleaves, cw_valid, cw_depth;
	// Q16 tree: the FAST tree with 32-byte nodes — both child boxes quantised CONSERVATIVELY to 16 bits per plane on
	// ONE grid over the scene box (plane = qmin + q * qstep), 2 x float4 per node:
	//   [2i]   = c0: (min.x | max.x << 16) (min.y | max.y << 16) (min.z | max.z << 16) bits(child0)
	//   [2i+1] = c1: the same, bits(child1)        child >= 0: node index; < 0: ~(leaf record index)
	// qleaves: 2 x float4 per leaf record = exact min.xyz, bits(start << 2 | count) | exact max.xyz, 0
	const float4* qnodes;
	const float4* qleaves;
	int32_t q16_root;
	uint32_t n_qnodes;
	float qmin[3], qstep[3];
	const float4* tri;  // 4 x float4 per triangle: rtb_tri_isect re-packed for the intersection test (layout at triTest)
	const float4* tsh;  // 4 x float4 per triangle = rtb_tri_shade
	const rtb_material* mats;
	const rtb_texture* texs;
	const float4* texels;    // RGB texels padded to 16 bytes on upload (k_pad_texels): one LDG.128 per bilinear tap instead of three loads
	const rtb_light* lights;
	// env-map importance tables (RTB_SAMPLING_IMPORTANCE): rtb_accel.hpp buildEnvTables
	const float* env_marginal; // [H+1] cdf over rows
	const float* env_cond;     // [H*(W+1)] cdf over columns of each row
	uint32_t n_xnodes, n_fnodes, n_tris, n_lights, n_mats, n_texs;
	int32_t fast_root; // child reference of the FAST root (may itself be a leaf)
	int32_t wide_root;
	uint32_t n_wnodes;
	uint32_t bg_type;
	float bg_colour[3];
	int32_t bg_tex;
	int32_t env_w, env_h;
	// scene bounds (the reference tree's root box) for the ray-binning keys of rtb_wavefront.cuh:
	// cell = (p - bmin) * bscale, bscale = 16 / extent
	float bmin[3], bscale[3];
	uint32_t area_lights_only; // every light is an AreaLight: shadow rays end on a few triangles
};

struct HitD
{
	uint32_t id;
	float t, alpha, beta;
};

RTB_DEV float4 ldg4(const float4* p) { return __ldg(p); }
// 32 bytes in ONE load instruction (sm_100: LDG.E.256), p 32-byte aligned.  profiles/r02_v2_bathroom_summary.md: the
// traversal of the heavy scenes is bound by the L1 data pipe (87 % of the LSU wavefront peak) — every divergent lane costs
// one wavefront per load INSTRUCTION, so a 64-byte node fetched as 2 x 256 bits costs half of 4 x 128.
struct F8
{
	float4 a, b;
};
RTB_DEV F8 ldg8(const float4* p)
{
	F8 r;
	asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	    : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w)
	    : "l"(p));
	return r;
}

// ---------------------------------------------------------------------------------------
// AABB::rayAABB (RTBase/Geometry.h:173-184).  Returns the reference's accept decision;
// tEntry is t_entry_f (used only by the FAST traversal's ordering and culling).
// ---------------------------------------------------------------------------------------
RTB_DEV bool slabTest(float minx, float miny, float minz, float maxx, float maxy, float maxz, const RayD& r,
                      float& tEntry)
{
	float ax = (minx - r.o.x) * r.inv.x, ay = (miny - r.o.y) * r.inv.y, az = (minz - r.o.z) * r.inv.z;
	float bx = (maxx - r.o.x) * r.inv.x, by = (maxy - r.o.y) * r.inv.y, bz = (maxz - r.o.z) * r.inv.z;
	float ex = selMin(ax, bx), ey = selMin(ay, by), ez = selMin(az, bz);
	float xx = selMax(ax, bx), xy = selMax(ay, by), xz = selMax(az, bz);
	float te = stdMax(stdMax(ex, ey), ez);
	float tx = stdMin(stdMin(xx, xy), xz);
	tEntry = te;
	return !(tx < te || tx < 0.0f);
}

// ---------------------------------------------------------------------------------------
// Triangle::rayIntersect (RTBase/Geometry.h:89-105).
// ---------------------------------------------------------------------------------------
RTB_DEV bool triTest(const float4* __restrict__ tri, uint32_t id, const RayD& r, float& t, float& u, float& v)
{
	// device record (rtb_api.cu prepareScene): [n.xyz, d] [v0.xyz, inv_area] [v1.xyz, bits(material)] [v2.xyz, area] — the plane
	// test (which rejects most candidates) needs the first 16 bytes only
	const float4* q = tri + (size_t)id * 4;
#if RTB_TRI256
	const F8 h0 = ldg8(q);
	const float4 qn = h0.a, qv0 = h0.b;
#else
	const float4 qn = ldg4(q);
#endif
	V3 n = mk(qn);
	float denom = dot(n, r.d);
	if (denom == 0.0f) return false;
	t = (qn.w - dot(n, r.o)) / denom;
	if (t < 0.0f) return false;
#if RTB_TRI256
	const F8 h1 = ldg8(q + 2);
	const float4 q1 = h1.a, q2 = h1.b;
#else
	const float4 qv0 = ldg4(q + 1), q1 = ldg4(q + 2), q2 = ldg4(q + 3);
#endif
	V3 v0 = mk(qv0), v1 = mk(q1), v2 = mk(q2);
	V3 p = r.o + (r.d * t);
	V3 e1 = v2 - v1;
	V3 e2 = v0 - v2;
	float invArea = qv0.w;
	u = dot(cross(e1, p - v1), n) * invArea;
	if (u < 0.0f || u > 1.0f) return false;
	v = dot(cross(e2, p - v2), n) * invArea;
	if (v < 0.0f || (u + v) > 1.0f) return false;
	return true;
}

// ---------------------------------------------------------------------------------------
// EXACT traversal: BVHNode::traverse (RTBase/Geometry.h:399-427) on the reference's own
// tree, every node whose box passes, left before right, no culling; stack-free through the
// pre-order skip links.  Leaf acceptance `t < best && t > EPSILON` (:412).
// ---------------------------------------------------------------------------------------
// A NaN in the origin or the direction (the reference's glass / normalise paths can produce one) makes
// every triangle's plane distance t NaN: (d - n.o) / (n.d) has a NaN operand whichever component it is, and
// 0 * NaN = NaN.  Triangle::rayIntersect then "hits" (all its rejects are comparisons, false on NaN) but
// `t < best && t > EPSILON` (Geometry.h:412) is false: the reference's exhaustive walk returns a miss.
// Walking 10^5 boxes that all pass to learn that would cost one thread tens of milliseconds — and a
// persistent warp, hence the launch, with it (one such ray in bathroom sample 861 cost 0.35 s).
RTB_DEV bool rayHasNaN(const RayD& r)
{
	return isnan(r.o.x) || isnan(r.o.y) || isnan(r.o.z) || isnan(r.d.x) || isnan(r.d.y) || isnan(r.d.z);
}

RTB_DEV void closestExact(const DevScene& S, const RayD& r, float eps, HitD& h, uint32_t& nBox, uint32_t& nTri)
{
	h.id = RTB_MISS_ID;
	h.t = FLT_MAX;
	h.alpha = h.beta = 0.0f;
	if (rayHasNaN(r)) return;
	uint32_t i = 0;
	while (i < S.n_xnodes)
	{
		float4 A = ldg4(S.xnodes + 2 * (size_t)i);
		float4 B = ldg4(S.xnodes + 2 * (size_t)i + 1);
		float te;
		nBox++;
		if (!slabTest(A.x, A.y, A.z, B.x, B.y, B.z, r, te))
		{
			i = __float_as_uint(A.w);
			continue;
		}
		uint32_t leaf = __float_as_uint(B.w);
		if (leaf == RTB_INTERIOR)
		{
			i = i + 1;
			continue;
		}
		uint32_t start = leaf >> 2, count = leaf & 3u;
		for (uint32_t k = 0; k < count; k++)
		{
			float t, u, v;
			nTri++;
			if (triTest(S.tri, start + k, r, t, u, v))
			{
				if (t < h.t && t > eps)
				{
					h.t = t, h.id = start + k, h.alpha = u, h.beta = v;
				}
			}
		}
		i = __float_as_uint(A.w);
	}
}

// BVHNode::traverseVisible (RTBase/Geometry.h:435-462): true = nothing in (eps, maxT).
RTB_DEV bool visibleExact(const DevScene& S, const RayD& r, float eps, float maxT, uint32_t& nBox, uint32_t& nTri)
{
	uint32_t i = 0;
	while (i < S.n_xnodes)
	{
		float4 A = ldg4(S.xnodes + 2 * (size_t)i);
		float4 B = ldg4(S.xnodes + 2 * (size_t)i + 1);
		float te;
		nBox++;
		if (!slabTest(A.x, A.y, A.z, B.x, B.y, B.z, r, te))
		{
			i = __float_as_uint(A.w);
			continue;
		}
		uint32_t leaf = __float_as_uint(B.w);
		if (leaf == RTB_INTERIOR)
		{
			i = i + 1;
			continue;
		}
		uint32_t start = leaf >> 2, count = leaf & 3u;
		for (uint32_t k = 0; k < count; k++)
		{
			float t, u, v;
			nTri++;
			if (triTest(S.tri, start + k, r, t, u, v))
			{
				if (t >= maxT || t <= eps) continue;
				return false;
			}
		}
		i = __float_as_uint(A.w);
	}
	return true;
}

// ---------------------------------------------------------------------------------------
// FAST traversal: a different tree over the SAME leaves.  Legal because the
// reference's result is the lexicographic (t, ID) minimum over the triangles of every leaf
// whose exact box passes; a leaf box passing implies all its ancestors pass (box nesting +
// monotone rounding) unless a ray-direction component is 0/denormal (0*inf = NaN can
// reject an ancestor only) — such rays take the EXACT path.  Every box here, interior or
// leaf, is tested with the reference's exact slab arithmetic, children are visited
// near-first and popped nodes are dropped when t_entry - |t_entry|*rel > t_best.
// ---------------------------------------------------------------------------------------
RTB_DEV bool rayIsDegenerate(const RayD& r)
{
	// an infinite reciprocal (direction component +-0 or so small that 1/d overflows) or a NaN anywhere:
	// the accelerated trees assume finite slab arithmetic
	return !(fabsf(r.inv.x) <= FLT_MAX) || !(fabsf(r.inv.y) <= FLT_MAX) || !(fabsf(r.inv.z) <= FLT_MAX) || rayHasNaN(r);
}

RTB_DEV void leafClosest(const DevScene& S, int32_t ref, const RayD& r, float eps, HitD& h, uint32_t& nTri)
{
	uint32_t leaf = (uint32_t)(~ref);
	uint32_t start = leaf >> 2, count = leaf & 3u;
	for (uint32_t k = 0; k < count; k++)
	{
		float t, u, v;
		nTri++;
		if (triTest(S.tri, start + k, r, t, u, v))
		{
			uint32_t id = start + k;
			if (t > eps && (t < h.t || (t == h.t && id < h.id)))
			{
				h.t = t, h.id = id, h.alpha = u, h.beta = v;
			}
		}
	}
}

// Without NaN operands (no 0*inf: degenerate rays go to the EXACT tree) the selects of
// RTBase/Core.h:187-195 and fminf/fmaxf agree except for the sign of a zero, which no later
// comparison can see: same accept decisions as slabTest().
RTB_DEV bool slabTestNoNaN(float minx, float miny, float minz, float maxx, float maxy, float maxz, const RayD& r,
                           float& tEntry)
{
	float ax = (minx - r.o.x) * r.inv.x, ay = (miny - r.o.y) * r.inv.y, az = (minz - r.o.z) * r.inv.z;
	float bx = (maxx - r.o.x) * r.inv.x, by = (maxy - r.o.y) * r.inv.y, bz = (maxz - r.o.z) * r.inv.z;
	float te = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
	float tx = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
	tEntry = te;
	return !(tx < te || tx < 0.0f);
}

#define RTB_TRAV_DONE_ 0x7FFFFFFF
#define RTB_WIDE_EMPTY 0x7FFFFFFE

// Per-ray traversal state shared by the one-thread-per-ray loops below and the persistent
// warp kernels of rtb_wavefront.cuh.  `bestT` is the best t so far (closest hit) or maxT (any hit).
template <bool ANYHIT>
struct LaneTrav
{
	RayD r;
	float cullT; // cull threshold derived from bestT (travSetBest)
	float bestT;
	uint32_t bestId;
	float bestU, bestV;
	int32_t cur; // >= 0 interior node, < 0 leaf reference, RTB_TRAV_DONE_ finished
	int sp;
	// RTB_TRAV_Q16 only (dead registers otherwise): t_plane = (2^23 + q) * qa + qb per axis, and the PRMT selectors that
	// pick the NEAR plane's 16 bits of a (min | max << 16) word for this ray's direction signs
	float qa[3], qb[3];
	uint32_t qsel[3];
};

// Traversal stacks of (node, t_entry) pairs.  LocalStack: RTB_STACK entries of local memory per thread.  SharedStack
// (the persistent kernels): the first NS entries live in shared memory, entry-major ([entry][thread]: whatever their
// depths, the 32 lanes of a warp touch 32 different 8-byte words of 32 banks pairs - 2 wavefronts per access, where
// a divergent local-memory access costs up to 32 on the L1 data pipe, the measured bound of the extend stage on the
// heavy scenes); deeper entries spill to local memory.
#define RTB_STACK 96 /* rtb_upload_scene rejects trees that could need more */
struct LocalStack
{
	int32_t node[RTB_STACK];
	float t[RTB_STACK];
	RTB_DEV void put(int i, int32_t n, float te) { node[i] = n, t[i] = te; }
	RTB_DEV void get(int i, int32_t& n, float& te) const { n = node[i], te = t[i]; }
	// any-hit rays: maxT never changes, so an entry that was admitted when pushed is still admitted when popped
	RTB_DEV void putNode(int i, int32_t n) { node[i] = n; }
	RTB_DEV int32_t getNode(int i) const { return node[i]; }
};
template <int NS, int THREADS>
struct SharedStack
{
	float2* s; // this thread's column: entry i at s[i * THREADS]
	int32_t node[RTB_STACK - NS];
	float t[RTB_STACK - NS];
	RTB_DEV void put(int i, int32_t n, float te)
	{
		if (i < NS) s[i * THREADS] = make_float2(__int_as_float(n), te);
		else node[i - NS] = n, t[i - NS] = te;
	}
	RTB_DEV void get(int i, int32_t& n, float& te) const
	{
		if (i < NS)
		{
			float2 v = s[i * THREADS];
			n = __float_as_int(v.x), te = v.y;
		}
		else
			n = node[i - NS], te = t[i - NS];
	}
	RTB_DEV void putNode(int i, int32_t n) { put(i, n, 0.0f); }
	RTB_DEV int32_t getNode(int i) const
	{
		int32_t n;
		float te;
		get(i, n, te);
		return n;
	}
};

// Culling: a box entered beyond the best hit cannot matter, but t_entry (slab
// arithmetic) and the triangle's t (plane arithmetic) are rounded differently, so the comparison
// needs slack: a subtree is dropped only if t_entry > bestT * (1 + 2 rel) (closest hit) or
// t_entry >= maxT * (1 + 2 rel) (any hit, where t < maxT is required).  The threshold is kept in
// a register and refreshed when bestT changes: one compare per box.
template <bool ANYHIT>
RTB_DEV void travSetBest(LaneTrav<ANYHIT>& t, float best, float cullRel)
{
	t.bestT = best;
	t.cullT = best + fabsf(best) * (2.0f * cullRel);
}
template <bool ANYHIT>
RTB_DEV bool travCull(float te, float cullT)
{
	return ANYHIT ? (te >= cullT) : (te > cullT);
}

template <bool ANYHIT, class STK>
RTB_DEV void lanePop(LaneTrav<ANYHIT>& t, const STK& stk)
{
	t.cur = RTB_TRAV_DONE_;
	if (ANYHIT)
	{
		if (t.sp > 0) t.cur = stk.getNode(--t.sp);
		return;
	}
	while (t.sp > 0)
	{
		t.sp--;
		int32_t n;
		float te;
		stk.get(t.sp, n, te);
		if (travCull<ANYHIT>(te, t.cullT)) continue;
		t.cur = n;
		break;
	}
}

// One interior step of the binary FAST tree: both child boxes, near child next, far child pushed.
// any-hit rays at a node whose two children are both hit: 0 = stored order (child 0 first), 1 = nearer child first
#ifndef RTB_ANYHIT_NEAR_FIRST
#define RTB_ANYHIT_NEAR_FIRST 0
#endif
template <bool ANYHIT, class STK>
RTB_DEV void stepFast(const DevScene& S, LaneTrav<ANYHIT>& t, STK& stk, uint32_t& nBox)
{
	const float4* nd = S.fnodes + (size_t)t.cur * 4;
#if RTB_LDG256
	const F8 lo = ldg8(nd), hi = ldg8(nd + 2);
	const float4 n0 = lo.a, n1 = lo.b, nz = hi.a, ch = hi.b;
#else
	float4 n0 = ldg4(nd), n1 = ldg4(nd + 1), nz = ldg4(nd + 2), ch = ldg4(nd + 3);
#endif
	float t0, t1;
	nBox += 2;
	bool h0 = slabTestNoNaN(n0.x, n0.z, nz.x, n0.y, n0.w, nz.y, t.r, t0);
	bool h1 = slabTestNoNaN(n1.x, n1.z, nz.z, n1.y, n1.w, nz.w, t.r, t1);
	h0 = h0 && !travCull<ANYHIT>(t0, t.cullT);
	h1 = h1 && !travCull<ANYHIT>(t1, t.cullT);
	int32_t c0 = __float_as_int(ch.x), c1 = __float_as_int(ch.y);
	if (h0 && h1)
	{
		bool swap = (!ANYHIT || RTB_ANYHIT_NEAR_FIRST) && (t1 < t0);
		if (ANYHIT) stk.putNode(t.sp, swap ? c0 : c1);
		else stk.put(t.sp, swap ? c0 : c1, swap ? t0 : t1);
		t.sp++;
		t.cur = swap ? c1 : c0;
	}
	else if (h0) t.cur = c0;
	else if (h1) t.cur = c1;
	else lanePop<ANYHIT>(t, stk);
}

// One interior step of the Q16 tree: the FAST tree's topology in 32-byte nodes (two LDG.128 instead of four, half the
// cache footprint).  Interior boxes only have to be conservative: every plane is rounded outwards by two
// steps of a 2^16 grid over the scene box at build time, and the test is the FMA form t = (2^23 + q) * a + b with
// a = step * invDir and b = (qmin - o) * invDir - 2^23 a per RAY (not per node); PRMT turns 16 bits into the float
// 2^23 + q and picks near / far by the ray's direction signs at no cost.  Its error against the reference's
// (plane - o) * invDir — under 0.6 grid steps plus 2^-22 relative — is inside the padding and the relative slack
// RTB_Q16_SLACK on the entry distance.  The entry distance used for ordering and culling is that lower bound.
#define RTB_Q16_SLACK 0.99999619f /* 1 - 2^-18 */
#define RTB_Q16_F(w, sel) __uint_as_float(__byte_perm((w), 0x4B000000u, (sel)))
template <bool ANYHIT, class STK>
RTB_DEV void stepQ16(const DevScene& S, LaneTrav<ANYHIT>& t, STK& stk, uint32_t& nBox)
{
	const float4* nd = S.qnodes + (size_t)t.cur * 2;
#if RTB_LDG256
	const F8 nn = ldg8(nd);
	const float4 A = nn.a, B = nn.b;
#else
	const float4 A = ldg4(nd), B = ldg4(nd + 1);
#endif
	const uint32_t snx = t.qsel[0], sny = t.qsel[1], snz = t.qsel[2];
	const uint32_t sfx = snx ^ 0x0022u, sfy = sny ^ 0x0022u, sfz = snz ^ 0x0022u;
	const uint32_t a0 = __float_as_uint(A.x), a1 = __float_as_uint(A.y), a2 = __float_as_uint(A.z);
	const uint32_t b0 = __float_as_uint(B.x), b1 = __float_as_uint(B.y), b2 = __float_as_uint(B.z);
	nBox += 2;
	float n0 = fmaxf(fmaxf(fmaf(RTB_Q16_F(a0, snx), t.qa[0], t.qb[0]), fmaf(RTB_Q16_F(a1, sny), t.qa[1], t.qb[1])),
	                 fmaxf(fmaf(RTB_Q16_F(a2, snz), t.qa[2], t.qb[2]), 0.0f));
	float f0 = fminf(fminf(fmaf(RTB_Q16_F(a0, sfx), t.qa[0], t.qb[0]), fmaf(RTB_Q16_F(a1, sfy), t.qa[1], t.qb[1])),
	                 fminf(fmaf(RTB_Q16_F(a2, sfz), t.qa[2], t.qb[2]), t.cullT));
	float n1 = fmaxf(fmaxf(fmaf(RTB_Q16_F(b0, snx), t.qa[0], t.qb[0]), fmaf(RTB_Q16_F(b1, sny), t.qa[1], t.qb[1])),
	                 fmaxf(fmaf(RTB_Q16_F(b2, snz), t.qa[2], t.qb[2]), 0.0f));
	float f1 = fminf(fminf(fmaf(RTB_Q16_F(b0, sfx), t.qa[0], t.qb[0]), fmaf(RTB_Q16_F(b1, sfy), t.qa[1], t.qb[1])),
	                 fminf(fmaf(RTB_Q16_F(b2, sfz), t.qa[2], t.qb[2]), t.cullT));
	const float t0 = n0 * RTB_Q16_SLACK, t1 = n1 * RTB_Q16_SLACK;
	const bool h0 = t0 <= f0, h1 = t1 <= f1;
	const int32_t c0 = __float_as_int(A.w), c1 = __float_as_int(B.w);
	if (h0 && h1)
	{
		bool swap = !ANYHIT && (t1 < t0);
		if (ANYHIT) stk.putNode(t.sp, c1);
		else stk.put(t.sp, swap ? c0 : c1, swap ? t0 : t1);
		t.sp++;
		t.cur = swap ? c1 : c0;
	}
	else if (h0) t.cur = c0;
	else if (h1) t.cur = c1;
	else lanePop<ANYHIT>(t, stk);
}

// per-ray constants of stepQ16
template <bool ANYHIT>
RTB_DEV void q16Start(const DevScene& S, LaneTrav<ANYHIT>& t)
{
	const float inv[3] = {t.r.inv.x, t.r.inv.y, t.r.inv.z}, o[3] = {t.r.o.x, t.r.o.y, t.r.o.z};
#pragma unroll
	for (int k = 0; k < 3; k++)
	{
		t.qa[k] = S.qstep[k] * inv[k];
		t.qb[k] = fmaf(-8388608.0f, t.qa[k], (S.qmin[k] - o[k]) * inv[k]);
		t.qsel[k] = inv[k] >= 0.0f ? 0x7410u : 0x7432u;
	}
}

// One interior step of the 4-wide tree (rtb_accel.hpp WideBuilder): four child boxes from one
// 128-byte structure-of-arrays node.  Closest hit: children are visited near to far.  The order
// comes from sorting four keys = float bits of max(t_entry, 0) with the child slot in the two
// low mantissa bits (a 5-comparator network on unsigned integers); the t pushed on the stack is
// that key with the slot bits cleared: never larger than the true t_entry, so culling stays
// conservative.
RTB_DEV void cswap(uint32_t& a, uint32_t& b)
{
	uint32_t lo = min(a, b), hi = max(a, b);
	a = lo, b = hi;
}
template <bool ANYHIT, class STK>
RTB_DEV void stepWide(const DevScene& S, LaneTrav<ANYHIT>& t, STK& stk, uint32_t& nBox)
{
	const float4* nd = S.wnodes + (size_t)t.cur * 8;
#if RTB_LDG256
	const F8 w0 = ldg8(nd), w1 = ldg8(nd + 2), w2 = ldg8(nd + 4), w3 = ldg8(nd + 6);
	const float4 mnx = w0.a, mxx = w0.b, mny = w1.a, mxy = w1.b, mnz = w2.a, mxz = w2.b, rf = w3.a;
#else
	float4 mnx = ldg4(nd), mxx = ldg4(nd + 1), mny = ldg4(nd + 2), mxy = ldg4(nd + 3), mnz = ldg4(nd + 4), mxz = ldg4(nd + 5);
	float4 rf = ldg4(nd + 6);
#endif
	int32_t c0 = __float_as_int(rf.x), c1 = __float_as_int(rf.y), c2 = __float_as_int(rf.z), c3 = __float_as_int(rf.w);
	float e0, e1, e2, e3;
	bool h0 = slabTestNoNaN(mnx.x, mny.x, mnz.x, mxx.x, mxy.x, mxz.x, t.r, e0);
	bool h1 = slabTestNoNaN(mnx.y, mny.y, mnz.y, mxx.y, mxy.y, mxz.y, t.r, e1);
	bool h2 = slabTestNoNaN(mnx.z, mny.z, mnz.z, mxx.z, mxy.z, mxz.z, t.r, e2) && c2 != RTB_WIDE_EMPTY;
	bool h3 = slabTestNoNaN(mnx.w, mny.w, mnz.w, mxx.w, mxy.w, mxz.w, t.r, e3) && c3 != RTB_WIDE_EMPTY;
	nBox += 2u + (c2 != RTB_WIDE_EMPTY) + (c3 != RTB_WIDE_EMPTY);
	h0 = h0 && !travCull<ANYHIT>(e0, t.cullT);
	h1 = h1 && !travCull<ANYHIT>(e1, t.cullT);
	h2 = h2 && !travCull<ANYHIT>(e2, t.cullT);
	h3 = h3 && !travCull<ANYHIT>(e3, t.cullT);
	if (ANYHIT)
	{
		// order is irrelevant for the result; push every admitted child, continue with the last
		if (h0) stk.putNode(t.sp, c0), t.sp++;
		if (h1) stk.putNode(t.sp, c1), t.sp++;
		if (h2) stk.putNode(t.sp, c2), t.sp++;
		if (h3) stk.putNode(t.sp, c3), t.sp++;
		lanePop<ANYHIT>(t, stk);
		return;
	}
	uint32_t k0 = h0 ? ((__float_as_uint(fmaxf(e0, 0.0f)) & ~3u) | 0u) : 0xFFFFFFFFu;
	uint32_t k1 = h1 ? ((__float_as_uint(fmaxf(e1, 0.0f)) & ~3u) | 1u) : 0xFFFFFFFFu;
	uint32_t k2 = h2 ? ((__float_as_uint(fmaxf(e2, 0.0f)) & ~3u) | 2u) : 0xFFFFFFFFu;
	uint32_t k3 = h3 ? ((__float_as_uint(fmaxf(e3, 0.0f)) & ~3u) | 3u) : 0xFFFFFFFFu;
	cswap(k0, k1), cswap(k2, k3), cswap(k0, k2), cswap(k1, k3), cswap(k1, k2);
	// farthest first onto the stack, nearest becomes the next node
#define RTB_WIDE_REF(k) (((k) & 3u) == 0u ? c0 : ((k) & 3u) == 1u ? c1 : ((k) & 3u) == 2u ? c2 : c3)
	if (k3 != 0xFFFFFFFFu) stk.put(t.sp, RTB_WIDE_REF(k3), __uint_as_float(k3 & ~3u)), t.sp++;
	if (k2 != 0xFFFFFFFFu) stk.put(t.sp, RTB_WIDE_REF(k2), __uint_as_float(k2 & ~3u)), t.sp++;
	if (k1 != 0xFFFFFFFFu) stk.put(t.sp, RTB_WIDE_REF(k1), __uint_as_float(k1 & ~3u)), t.sp++;
	if (k0 != 0xFFFFFFFFu) t.cur = RTB_WIDE_REF(k0);
	else lanePop<ANYHIT>(t, stk);
#undef RTB_WIDE_REF
}

template <int TRAV, bool ANYHIT, class STK>
RTB_DEV void stepInterior(const DevScene& S, LaneTrav<ANYHIT>& t, STK& stk, uint32_t& nBox)
{
	if (TRAV == RTB_TRAV_WIDE) stepWide<ANYHIT>(S, t, stk, nBox);
	else if (TRAV == RTB_TRAV_Q16) stepQ16<ANYHIT>(S, t, stk, nBox);
	else stepFast<ANYHIT>(S, t, stk, nBox);
}
template <int TRAV>
RTB_DEV int32_t travRoot(const DevScene& S)
{
	return TRAV == RTB_TRAV_WIDE ? S.wide_root : TRAV == RTB_TRAV_Q16 ? S.q16_root : S.fast_root;
}
// what a ray needs besides LaneTrav's common fields before its first step
template <int TRAV, bool ANYHIT>
RTB_DEV void travStart(const DevScene& S, LaneTrav<ANYHIT>& t)
{
	if (TRAV == RTB_TRAV_Q16) q16Start<ANYHIT>(S, t);
}
// rays the accelerated trees do not take (they walk the reference's own tree): 0 * inf = NaN rays, and for
// the FMA-form tests reciprocals so large that step * invDir * 2^23 could overflow (axis-parallel to 1e-28)
template <int TRAV>
RTB_DEV bool travDegenerate(const RayD& r)
{
	if (TRAV == RTB_TRAV_Q16) return !(fabsf(r.inv.x) <= 1e28f) || !(fabsf(r.inv.y) <= 1e28f) || !(fabsf(r.inv.z) <= 1e28f) || rayHasNaN(r);
	return rayIsDegenerate(r);
}

RTB_DEV bool leafOccludes(const DevScene& S, int32_t ref, const RayD& r, float eps, float maxT, uint32_t& nTri)
{
	uint32_t leaf = (uint32_t)(~ref);
	uint32_t start = leaf >> 2, count = leaf & 3u;
	for (uint32_t k = 0; k < count; k++)
	{
		float t, u, v;
		nTri++;
		if (triTest(S.tri, start + k, r, t, u, v))
		{
			if (t >= maxT || t <= eps) continue;
			return true;
		}
	}
	return false;
}


// A reached leaf.  FAST / WIDE: `cur` = ~(start << 2 | count) and the box that admitted it WAS the reference's exact leaf
// box.  Q16: `cur` = ~(leaf record); the quantised slot box was only conservative, so the reference's exact leaf box
// (Geometry.h:173-184 arithmetic) is tested here before the leaf's triangles (condition 2).
template <int TRAV, bool ANYHIT>
RTB_DEV bool travLeafRef(const DevScene& S, const LaneTrav<ANYHIT>& t, int32_t& ref, uint32_t& nBox)
{
	if (TRAV != RTB_TRAV_Q16)
	{
		ref = t.cur;
		return true;
	}
	const float4* lf = S.qleaves + (size_t)(uint32_t)(~t.cur) * 2;
#if RTB_LDG256
	const F8 ll = ldg8(lf);
	const float4 A = ll.a, B = ll.b;
#else
	const float4 A = ldg4(lf), B = ldg4(lf + 1);
#endif
	float te;
	nBox++;
	if (!slabTestNoNaN(A.x, A.y, A.z, B.x, B.y, B.z, t.r, te)) return false;
	if (travCull<ANYHIT>(te, t.cullT)) return false;
	ref = ~(int32_t)__float_as_uint(A.w);
	return true;
}

// One thread runs one ray to the end (parity entry points, shadow stage, megakernel).
template <int TRAV>
RTB_DEV void closestAccel(const DevScene& S, const RayD& r, float eps, float cullRel, HitD& h, uint32_t& nBox,
                          uint32_t& nTri)
{
	if (travDegenerate<TRAV>(r) || travRoot<TRAV>(S) < 0)
	{
		// 0*inf = NaN rays and single-leaf scenes take the reference's own tree
		closestExact(S, r, eps, h, nBox, nTri);
		return;
	}
	LocalStack stk;
	LaneTrav<false> t;
	t.r = r;
	travSetBest<false>(t, FLT_MAX, cullRel);
	t.bestId = RTB_MISS_ID, t.bestU = t.bestV = 0.0f;
	t.sp = 0;
	t.cur = travRoot<TRAV>(S);
	travStart<TRAV, false>(S, t);
	for (;;)
	{
		while (t.cur >= 0 && t.cur != RTB_TRAV_DONE_) stepInterior<TRAV, false>(S, t, stk, nBox);
		if (t.cur == RTB_TRAV_DONE_) break;
		int32_t ref;
		if (travLeafRef<TRAV, false>(S, t, ref, nBox))
		{
			HitD b;
			b.id = t.bestId, b.t = t.bestT, b.alpha = t.bestU, b.beta = t.bestV;
			leafClosest(S, ref, t.r, eps, b, nTri);
			t.bestId = b.id, t.bestU = b.alpha, t.bestV = b.beta;
			travSetBest<false>(t, b.t, cullRel);
		}
		lanePop<false>(t, stk);
	}
	h.id = t.bestId, h.t = t.bestT, h.alpha = t.bestU, h.beta = t.bestV;
}

template <int TRAV>
RTB_DEV bool visibleAccel(const DevScene& S, const RayD& r, float eps, float maxT, float cullRel, uint32_t& nBox,
                          uint32_t& nTri)
{
	if (travDegenerate<TRAV>(r) || travRoot<TRAV>(S) < 0) return visibleExact(S, r, eps, maxT, nBox, nTri);
	LocalStack stk;
	LaneTrav<true> t;
	t.r = r;
	travSetBest<true>(t, maxT, cullRel);
	t.sp = 0;
	t.cur = travRoot<TRAV>(S);
	travStart<TRAV, true>(S, t);
	for (;;)
	{
		while (t.cur >= 0 && t.cur != RTB_TRAV_DONE_) stepInterior<TRAV, true>(S, t, stk, nBox);
		if (t.cur == RTB_TRAV_DONE_) return true;
		int32_t ref;
		if (travLeafRef<TRAV, true>(S, t, ref, nBox) && leafOccludes(S, ref, t.r, eps, maxT, nTri)) return false;
		lanePop<true>(t, stk);
	}
}

// RTB_TRAV_CW (rtb_dev_cw.cuh), nodes read from global memory; scenes without a CW tree take FAST
RTB_DEV void closestCwGlobal(const DevScene& S, const RayD& r, float eps, float cullRel, HitD& h, uint32_t& nBox, uint32_t& nTri);
RTB_DEV bool visibleCwGlobal(const DevScene& S, const RayD& r, float eps, float maxT, float cullRel, uint32_t& nBox, uint32_t& nTri);

template <int TRAV>
RTB_DEV void closestHit(const DevScene& S, const RayD& r, float eps, float cullRel, HitD& h, uint32_t& nBox,
                        uint32_t& nTri)
{
	if (TRAV == RTB_TRAV_EXACT) closestExact(S, r, eps, h, nBox, nTri);
	else if (TRAV == RTB_TRAV_CW)
	{
		if (S.cw_valid) closestCwGlobal(S, r, eps, cullRel, h, nBox, nTri);
		else closestAccel<RTB_TRAV_FAST>(S, r, eps, cullRel, h, nBox, nTri);
	}
	else closestAccel<TRAV>(S, r, eps, cullRel, h, nBox, nTri);
}

template <int TRAV>
RTB_DEV bool anyVisible(const DevScene& S, const RayD& r, float eps, float maxT, float cullRel, uint32_t& nBox,
                        uint32_t& nTri)
{
	if (TRAV == RTB_TRAV_EXACT) return visibleExact(S, r, eps, maxT, nBox, nTri);
	if (TRAV == RTB_TRAV_CW)
	{
		if (S.cw_valid) return visibleCwGlobal(S, r, eps, maxT, cullRel, nBox, nTri);
		return visibleAccel<RTB_TRAV_FAST>(S, r, eps, maxT, cullRel, nBox, nTri);
	}
	return visibleAccel<TRAV>(S, r, eps, maxT, cullRel, nBox, nTri);
}

// Scene::visible (RTBase/Scene.h:161-169)
template <int TRAV>
RTB_DEV bool sceneVisible(const DevScene& S, V3 p1, V3 p2, float eps, float cullRel, uint32_t& nBox, uint32_t& nTri)
{
	V3 dir = p2 - p1;
	float maxT = sqrtf(lengthSq(dir)) - (2.0f * eps);
	dir = normalize(dir);
	RayD r = mkRay(p1 + (dir * eps), dir);
	return anyVisible<TRAV>(S, r, eps, maxT, cullRel, nBox, nTri);
}

// ---------------------------------------------------------------------------------------
// Camera::generateRay (RTBase/Scene.h:43-54) with Matrix::mulPoint / mulVec
// (RTBase/Core.h:295-309).  Must be bit-exact: it feeds the hit-ID gate.
// ---------------------------------------------------------------------------------------
RTB_DEV RayD generateRay(const rtb_camera& c, float x, float y)
{
	float xprime = x / c.width;
	float yprime = 1.0f - (y / c.height);
	xprime = (xprime * 2.0f) - 1.0f;
	yprime = (yprime * 2.0f) - 1.0f;
	const float* m = c.inv_proj;
	V3 d = mk(((xprime * m[0] + yprime * m[1]) + 1.0f * m[2]) + m[3],
	          ((xprime * m[4] + yprime * m[5]) + 1.0f * m[6]) + m[7],
	          ((xprime * m[8] + yprime * m[9]) + 1.0f * m[10]) + m[11]);
	const float* k = c.cam_to_world;
	V3 w = mk((d.x * k[0] + d.y * k[1]) + d.z * k[2], (d.x * k[4] + d.y * k[5]) + d.z * k[6],
	          (d.x * k[8] + d.y * k[9]) + d.z * k[10]);
	w = normalize(w);
	return mkRay(mk(c.origin), w);
}

// ---------------------------------------------------------------------------------------
// ShadingData (RTBase/Materials.h:15-35) + Scene::calculateShadingData (Scene.h:174-203),
// Triangle::interpolateAttributes / gNormal (Geometry.h:106-112,127-130),
// Frame::fromVector (Core.h:513-527).
// ---------------------------------------------------------------------------------------
struct ShadeD
{
	V3 x, wo, sN, gN;
	float tu, tv;
	V3 fu, fv, fw;
	float t;
	int32_t mat;
};

RTB_DEV void frameFromVector(V3 n, V3& u, V3& v, V3& w)
{
	w = normalize(n);
	if (fabsf(w.x) > fabsf(w.y))
	{
		float l = 1.0f / sqrtf(w.x * w.x + w.z * w.z);
		u = mk(w.z * l, 0.0f, -w.x * l);
	}
	else
	{
		float l = 1.0f / sqrtf(w.y * w.y + w.z * w.z);
		u = mk(0.0f, w.z * l, -w.y * l);
	}
	v = cross(w, u);
}
RTB_DEV V3 toLocal(const ShadeD& s, V3 a) { return mk(dot(a, s.fu), dot(a, s.fv), dot(a, s.fw)); }
RTB_DEV V3 toWorld(const ShadeD& s, V3 a) { return ((s.fu * a.x) + (s.fv * a.y)) + (s.fw * a.z); }

RTB_DEV void calcShading(const DevScene& S, uint32_t id, float t, float alpha, float beta, float gamma,
                         const RayD& r, ShadeD& sd)
{
	const float4* qi = S.tri + (size_t)id * 4;
	const float4* qs = S.tsh + (size_t)id * 4;
	float4 i2 = ldg4(qi + 2), i3 = ldg4(qi); // device record: material in [2].w, n in [0] (triTest)
	float4 s0 = ldg4(qs), s1 = ldg4(qs + 1), s2 = ldg4(qs + 2), s3 = ldg4(qs + 3);
	sd.x = r.o + (r.d * t);
	sd.gN = mk(i3) * s3.w;
	V3 nrm = ((mk(s0) * alpha) + (mk(s1) * beta)) + (mk(s2) * gamma);
	sd.sN = normalize(nrm);
	sd.tu = (s0.w * alpha + s1.w * beta) + s2.w * gamma;
	sd.tv = (s3.x * alpha + s3.y * beta) + s3.z * gamma;
	sd.mat = (int32_t)__float_as_uint(i2.w);
	sd.wo = -r.d;
	uint32_t flags = S.mats[sd.mat].flags;
	if (flags & RTB_MAT_TWO_SIDED)
	{
		if (dot(sd.wo, sd.sN) < 0.0f) sd.sN = -sd.sN;
		if (dot(sd.wo, sd.gN) < 0.0f) sd.gN = -sd.gN;
	}
	frameFromVector(sd.sN, sd.fu, sd.fv, sd.fw);
	sd.t = t;
}

// ---------------------------------------------------------------------------------------
// Texture::sample (RTBase/Imaging.h:72-94): software bilinear with wrap, float texels.
// ---------------------------------------------------------------------------------------
// (The same treatment of the material / texture / light RECORDS - 128-bit loads instead of the compiler's per-field
// loads - measured 1.6 % slower, and caching the albedo lookup per vertex 0.4 % slower: profiles/r02_texel16.txt.)
RTB_DEV V3 texel(const float4* __restrict__ texels, uint32_t base, int idx)
{
	return mk(ldg4(texels + ((size_t)base + (size_t)idx)));
}
RTB_DEV V3 sampleTexture(const DevScene& S, int tex, float tu, float tv)
{
	rtb_texture T = S.texs[tex];
	float u = stdMax(0.0f, fabsf(tu)) * (float)T.width;
	float v = stdMax(0.0f, fabsf(tv)) * (float)T.height;
	float flu = floorf(u), flv = floorf(v);
	float fu = u - flu;
	float fv = v - flv;
	float w0 = (1.0f - fu) * (1.0f - fv);
	float w1 = fu * (1.0f - fv);
	float w2 = (1.0f - fu) * fv;
	float w3 = fu * fv;
	if (T.width == 1 && T.height == 1)
	{
		// constant-colour textures (most materials): all four taps are texel 0
		V3 a = texel(S.texels, T.offset, 0);
		return (((a * w0) + (a * w1)) + (a * w2)) + (a * w3);
	}
	int x = (int)flu;
	int y = (int)flv;
	// x % width, (x + 1) % width with x >= 0; coordinates inside [0, 1) need no division
	if (x >= T.width) x = x % T.width;
	if (y >= T.height) y = y % T.height;
	int x1 = (x + 1 == T.width) ? 0 : x + 1, y1 = (y + 1 == T.height) ? 0 : y + 1;
	V3 a = texel(S.texels, T.offset, y * T.width + x);
	V3 b = texel(S.texels, T.offset, y * T.width + x1);
	V3 c = texel(S.texels, T.offset, y1 * T.width + x);
	V3 d = texel(S.texels, T.offset, y1 * T.width + x1);
	return (((a * w0) + (b * w1)) + (c * w2)) + (d * w3);
}

// ---------------------------------------------------------------------------------------
// SamplingDistributions (RTBase/Sampling.h:29-70) + sphericalToWorld (Core.h:547-550).
// `2.0f * M_PI * r2` is a double product in the reference; the float product used here differs
// by <= 1 ulp of phi, far inside the 1e-5 evaluation tolerance.  The same holds for the other
// places where the reference's arithmetic silently promotes to double (x / M_PI): they are done
// in float with a reciprocal multiply (<= 1.5 ulp).
// ---------------------------------------------------------------------------------------
RTB_DEV V3 sphericalToWorld(float theta, float phi)
{
	float st, ct, sp, cp;
	sincosf(theta, &st, &ct);
	sincosf(phi, &sp, &cp);
	return mk(cp * st, sp * st, ct);
}
RTB_DEV V3 cosineSampleHemisphere(float r1, float r2)
{
	float theta = acosf(sqrtf(r1));
	float phi = RTB_2PI_F * r2;
	return sphericalToWorld(theta, phi);
}
RTB_DEV V3 uniformSampleSphere(float r1, float r2)
{
	float theta = acosf(1.0f - 2.0f * r1);
	float phi = RTB_2PI_F * r2;
	return sphericalToWorld(theta, phi);
}

// ---------------------------------------------------------------------------------------
// BSDFs in PARITY mode = what RTBase/Materials.h actually computes:
//   Diffuse :118, OrenNayar :369, Plastic :414, Conductor :203, Dielectric :320 -> Lambert
//   Mirror :158, Glass :252 (+ ShadingHelper::fresnelDielectric :55-77); Layered -> base.
// ---------------------------------------------------------------------------------------
RTB_DEV V3 bsdfAlbedo(const DevScene& S, const rtb_material& m, const ShadeD& sd)
{
	return sampleTexture(S, m.tex, sd.tu, sd.tv);
}

RTB_DEV V3 bsdfEvaluate(const DevScene& S, const rtb_material& m, const ShadeD& sd, V3 wi)
{
	(void)wi;
	if (m.type == RTB_BSDF_GLASS) return mk(0.0f, 0.0f, 0.0f);     // Materials.h:295-299
	if (m.type == RTB_BSDF_MIRROR) return bsdfAlbedo(S, m, sd);     // Materials.h:178-183 (sic)
	return bsdfAlbedo(S, m, sd) * RTB_INV_PI_F;                     // albedo / M_PI
}

RTB_DEV float bsdfPdf(const rtb_material& m, const ShadeD& sd, V3 wi)
{
	if (m.type == RTB_BSDF_GLASS || m.type == RTB_BSDF_MIRROR) return 0.0f;
	V3 l = toLocal(sd, wi);
	return (l.z >= 0.0f) ? l.z * RTB_INV_PI_F : 0.0f; // cosineHemispherePDF, Sampling.h:44-48
}

// ShadingHelper::fresnelDielectric (Materials.h:55-77), including its non-standard Fpe
// denominator.  iorInt/iorExt are the CALL's arguments (etaI, etaT at Materials.h:274).
RTB_DEV float fresnelDielectric(float cosTheta, float iorInt, float iorExt, V3& wt, V3 wol)
{
	float ior = iorInt / iorExt;
	float sinTheta_i = sqrtf(1.0f - (cosTheta * cosTheta));
	float sinTheta_t = ior * sinTheta_i;
	float ior2sin2 = (ior * ior) * (1.0f - (cosTheta * cosTheta));
	if (ior2sin2 > 1.0f) return 1.0f;
	float cosTheta_t = sqrtf(1.0f - (sinTheta_t * sinTheta_t));
	wt = mk(-ior * wol.x, -ior * wol.y, -cosTheta_t);
	float Fpa = (cosTheta - ior * cosTheta_t) / (cosTheta + ior * cosTheta_t);
	float Fpe = (ior * cosTheta - cosTheta_t) / (ior * cosTheta + ior * cosTheta_t);
	float average = ((Fpa * Fpa) + (Fpe * Fpe)) * 0.5f;
	// clamp() = std::max(0, std::min(1, x)) (Materials.h:8-11) returns 1 for a NaN (cos_i a hair above 1 makes
	// sqrtf(1 - cos^2) NaN).  nvcc fuses the select pattern into FMUL.SAT, which flushes NaN to 0: the GPU would
	// refract along a NaN direction where the reference reflects.
	if (average != average) return 1.0f;
	return stdMax(0.0f, stdMin(1.0f, average));
}

// BSDF::sample.  r1, r2: the cosine-hemisphere uniforms; r3: glass reflect/refract draw.
// `usedR3` tells the caller whether the reference would have consumed that draw.
RTB_DEV V3 bsdfSample(const DevScene& S, const rtb_material& m, const ShadeD& sd, float r1, float r2, float r3,
                      V3& f, float& pdf)
{
	if (m.type == RTB_BSDF_MIRROR) // Materials.h:167-177
	{
		V3 wol = toLocal(sd, sd.wo);
		V3 wi = mk(-wol.x, -wol.y, wol.z);
		pdf = 1.0f;
		f = bsdfAlbedo(S, m, sd);
		return toWorld(sd, wi);
	}
	if (m.type == RTB_BSDF_GLASS) // Materials.h:265-294
	{
		V3 wol = toLocal(sd, sd.wo);
		float cosTheta_i = fabsf(wol.z);
		bool enter = wol.z > 0.0f;
		float etaI = enter ? m.ext_ior : m.int_ior;
		float etaT = enter ? m.int_ior : m.ext_ior;
		V3 wt = mk(0.0f, 0.0f, 0.0f); // Vec3() default (w = 1 is unused)
		float R = fresnelDielectric(cosTheta_i, etaI, etaT, wt, wol);
		if (!enter) wt.z = -wt.z;
		bool reflect = (R == 1.0f) || (r3 < R);
		V3 a = bsdfAlbedo(S, m, sd);
		V3 wi;
		if (reflect)
		{
			wi = mk(-wol.x, -wol.y, wol.z);
			pdf = R;
			f = a * R;
		}
		else
		{
			wi = wt;
			pdf = 1.0f - R;
			f = a * (1.0f - R);
		}
		return toWorld(sd, wi);
	}
	// Lambert-like stubs.  DiffuseBSDF clamps the pdf (Materials.h:130), the others use
	// pdf = wi.z / M_PI un-clamped (:222, :339, :384, :437).
	V3 wl = cosineSampleHemisphere(r1, r2);
	float p = wl.z * RTB_INV_PI_F;
	if (m.type == RTB_BSDF_DIFFUSE && !(wl.z >= 0.0f)) p = 0.0f;
	pdf = p;
	f = bsdfAlbedo(S, m, sd) * RTB_INV_PI_F;
	return toWorld(sd, wl);
}

// ---------------------------------------------------------------------------------------
// Lights (RTBase/Lights.h).  EnvironmentMap::evaluate :158-165; BackgroundColour :84-133;
// AreaLight :30-82 with Triangle::sample (Geometry.h:114-126).
// ---------------------------------------------------------------------------------------
RTB_DEV V3 envLookup(const DevScene& S, int tex, V3 wi)
{
	float u = atan2f(wi.z, wi.x);
	u = (u < 0.0f) ? u + RTB_2PI_F : u;
	u = u * RTB_INV_2PI_F;
	float v = acosf(wi.y) * RTB_INV_PI_F;
	return sampleTexture(S, tex, u, v);
}
// Scene::background->evaluate(dir)
RTB_DEV V3 backgroundEval(const DevScene& S, V3 wi)
{
	if (S.bg_type == RTB_LIGHT_ENVMAP) return envLookup(S, S.bg_tex, wi);
	return mk(S.bg_colour);
}
RTB_DEV V3 trianglePoint(const DevScene& S, uint32_t id, float r1, float r2)
{
	const float4* q = S.tri + (size_t)id * 4;
	V3 v0 = mk(ldg4(q + 1)), v1 = mk(ldg4(q + 2)), v2 = mk(ldg4(q + 3));
	float sr = sqrtf(r1);
	float alpha = 1.0f - sr;
	float beta = r2 * sr;
	float gamma = 1.0f - (alpha + beta);
	return ((v0 * alpha) + (v1 * beta)) + (v2 * gamma);
}
RTB_DEV V3 triangleGNormal(const DevScene& S, uint32_t id)
{
	float4 n = ldg4(S.tri + (size_t)id * 4);
	float gs = ldg4(S.tsh + (size_t)id * 4 + 3).w;
	return mk(n) * gs;
}

// ---------------------------------------------------------------------------------------
// Counter-based RNG (replaces MTRandom, RTBase/Sampling.h:13-26): Philox-4x32-10
// (Salmon et al. 2011).  counter = (pixel, sample, block, 0), key = (seed, "RTB2").
// One block = 4 uniforms; a path vertex at depth k uses blocks 2k and 2k+1:
//   block 2k   : [0] light pick  [1],[2] light sample  [3] Russian roulette
//   block 2k+1 : [0],[1] BSDF (r1, r2)  [2] glass reflect/refract  [3] unused
// Uniforms lie strictly inside (0,1): ((x >> 9) + 0.5) * 2^-23.
// ---------------------------------------------------------------------------------------
RTB_DEV uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1)
{
#pragma unroll
	for (int i = 0; i < 10; i++)
	{
		uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
		uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
		c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
		k0 += 0x9E3779B9u;
		k1 += 0xBB67AE85u;
	}
	return c;
}
RTB_DEV float u01(uint32_t x) { return ((float)(x >> 9) + 0.5f) * 1.1920928955078125e-7f; }
// stream 0 = camera paths (pixel, sample, block); stream 1 = light paths (path, pass, block)
RTB_DEV float4 rngBlock(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t block, uint32_t stream = 0u)
{
	uint4 r = philox4x32_10(make_uint4(pixel, sample, block, stream), seed, 0x52544232u);
	return make_float4(u01(r.x), u01(r.y), u01(r.z), u01(r.w));
}
