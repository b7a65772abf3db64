// Test infrastructure: builds the product's host-side trees (csrc/rtb_accel.hpp, csrc/rtb_cwbvh.hpp) from a flat scene's
// reference BVH and checks the properties the device traversals rely on:
//   FAST  every reference leaf is a primitive exactly once; every child box is EXACTLY the union of its leaves' boxes
//   WIDE  same leaves, same boxes, 2..4 children per node
//   CW    every quantised child box CONTAINS the exact box it stands for with >= 1 grid step of margin per side; every
//         reference leaf has exactly one leaf record carrying its exact box and triangle range; child / leaf base indices
//         and masks are consistent; depth fits the device stack
//   Q16   every quantised child box contains the exact FAST child box with >= 2 steps of margin; leaf records as above
// Compiled by tests/test_accel_cpu.py with g++ (the builders are plain host C++).  Returns 0 or a negative error code and
// a message.
#include "../../raytracingrenderer_b200/csrc/rtb_accel.hpp"
#include "../../raytracingrenderer_b200/csrc/rtb_cwbvh.hpp"

#include <cstdio>
#include <map>
#include <string>

using namespace rtb_accel;

static std::string g_msg;
static uint32_t fbits(float f)
{
	uint32_t u;
	memcpy(&u, &f, 4);
	return u;
}
struct BoxD
{
	double mn[3], mx[3];
};
static int failf(int code, const char* fmt, double a = 0, double b = 0, double c = 0)
{
	char buf[256];
	snprintf(buf, sizeof(buf), fmt, a, b, c);
	g_msg = buf;
	return code;
}

static int g_optPasses = 0;
static double g_optFraction = 1.0, g_sah[2];
// passes > 0: accel_check optimises the FAST tree first (rtb_accel::FastOptimizer); sah = summed interior surface area
// over the root's, before and after
extern "C" void accel_set_optimise(int passes, double fraction) { g_optPasses = passes, g_optFraction = fraction; }
static int g_orderMode = 0;
static const rtb_tri_isect* g_orderTris = nullptr;
extern "C" void accel_set_order(int mode, const rtb_tri_isect* tris) { g_orderMode = mode, g_orderTris = tris; }
extern "C" void accel_get_sah(double* out) { out[0] = g_sah[0], out[1] = g_sah[1]; }
extern "C" const char* accel_check_message() { return g_msg.c_str(); }

// stats: [0] fast nodes [1] fast depth [2] wide nodes [3] cw nodes [4] cw depth [5] cw leaves [6] q16 leaves
//        [7] mean CW child box volume inflation [8] mean Q16 box inflation
extern "C" int accel_check(const rtb_ref_node* nodes, uint32_t n, uint32_t nTris, double* stats)
{
	g_msg.clear();
	std::vector<F4> xnodes;
	std::vector<RefLeaf> leaves;
	const char* err = nullptr;
	if (!buildExact(nodes, n, nTris, xnodes, leaves, &err)) return failf(-1, err);
	FastTree fast;
	{
		FastBuilder fb(leaves);
		fb.build(fast);
	}
	g_sah[0] = g_sah[1] = 0;
	if (g_optPasses > 0 && fast.root >= 0 && fast.nodes.size() >= 8)
	{
		// the checks below then run on the re-optimised tree (RTB_TREE_OPT): same leaves, exact unions, legal depth
		FastOptimizer opt;
		opt.load(fast);
		g_sah[0] = opt.cost();
		for (int k = 0; k < g_optPasses; k++) opt.pass((float)g_optFraction);
		g_sah[1] = opt.cost();
		FastTree better;
		opt.store(better);
		fast = better;
	}
	if (g_orderMode > 0) orderForAnyHit(fast, g_orderTris, g_orderMode); // swaps children only: every check below must still hold
	std::map<uint32_t, const RefLeaf*> byKey; // (start << 2 | count) -> leaf
	for (const RefLeaf& L : leaves) byKey[(L.start << 2) | L.count] = &L;
	if (byKey.size() != leaves.size()) return failf(-2, "duplicate reference leaves");
	size_t nf = fast.nodes.size() / 4;
	stats[0] = (double)nf, stats[1] = fast.maxDepth;
	if (fast.root < 0 || nf == 0)
	{
		for (int k = 2; k < 9; k++) stats[k] = 0;
		return 0; // single-leaf / empty scene: the device uses the reference tree
	}
	// ---- FAST: exact subtree boxes, each leaf once
	std::vector<BoxD> sub(nf);
	std::vector<int> seen(nf, 0);
	std::map<uint32_t, int> leafSeen;
	struct Rec
	{
		static bool run(const FastTree& T, int32_t node, std::vector<BoxD>& sub, std::map<uint32_t, int>& leafSeen,
		                const std::map<uint32_t, const RefLeaf*>& byKey, BoxD& out, int depth)
		{
			const F4* nd = &T.nodes[(size_t)node * 4];
			float mn[2][3] = {{nd[0].x, nd[0].z, nd[2].x}, {nd[1].x, nd[1].z, nd[2].z}};
			float mx[2][3] = {{nd[0].y, nd[0].w, nd[2].y}, {nd[1].y, nd[1].w, nd[2].w}};
			int32_t ref[2] = {(int32_t)fbits(nd[3].x), (int32_t)fbits(nd[3].y)};
			for (int k = 0; k < 3; k++) out.mn[k] = 1e300, out.mx[k] = -1e300;
			for (int c = 0; c < 2; c++)
			{
				BoxD b;
				if (ref[c] < 0)
				{
					uint32_t key = (uint32_t)(~ref[c]);
					auto it = byKey.find(key);
					if (it == byKey.end()) return false;
					leafSeen[key]++;
					for (int k = 0; k < 3; k++) b.mn[k] = it->second->bmin[k], b.mx[k] = it->second->bmax[k];
				}
				else if (depth > 200 || !run(T, ref[c], sub, leafSeen, byKey, b, depth + 1))
					return false;
				for (int k = 0; k < 3; k++)
				{
					if ((double)mn[c][k] != b.mn[k] || (double)mx[c][k] != b.mx[k]) return false; // child box == exact union
					out.mn[k] = std::min(out.mn[k], b.mn[k]), out.mx[k] = std::max(out.mx[k], b.mx[k]);
				}
			}
			sub[node] = out;
			return true;
		}
	};
	BoxD rootBox;
	if (!Rec::run(fast, fast.root, sub, leafSeen, byKey, rootBox, 0)) return failf(-3, "FAST: a child box is not the exact union of its leaves' boxes (or a bad reference)");
	if (leafSeen.size() != leaves.size()) return failf(-4, "FAST: %g of %g reference leaves reachable", (double)leafSeen.size(), (double)leaves.size());
	for (auto& kv : leafSeen)
		if (kv.second != 1) return failf(-5, "FAST: a reference leaf appears %g times", kv.second);
	// ---- WIDE
	{
		WideTree wide;
		WideBuilder wb(fast);
		wb.build(wide);
		stats[2] = (double)(wide.nodes.size() / 8);
		std::map<uint32_t, int> wl;
		for (size_t i = 0; i < wide.nodes.size() / 8; i++)
		{
			const F4* nd = &wide.nodes[i * 8];
			float refs[4] = {nd[6].x, nd[6].y, nd[6].z, nd[6].w};
			int live = 0;
			for (int c = 0; c < 4; c++)
			{
				int32_t r = (int32_t)fbits(refs[c]);
				if (r == RTB_WIDE_EMPTY) continue;
				live++;
				if (r < 0) wl[(uint32_t)(~r)]++;
			}
			if (live < 2) return failf(-6, "WIDE: node %g has %g children", (double)i, live);
		}
		if (wl.size() != leaves.size()) return failf(-7, "WIDE: %g of %g leaves", (double)wl.size(), (double)leaves.size());
	}
	// ---- CW
	double cwInfl = 0;
	size_t cwCount = 0;
	{
		CwTree cw;
		CwBuilder cb(fast);
		cb.build(cw);
		if (!cw.valid) return failf(-8, "CW: not built");
		size_t nn = cw.nodes.size() / 5, nl = cw.leaves.size() / 2;
		stats[3] = (double)nn, stats[4] = cw.maxDepth, stats[5] = (double)nl;
		if (nl != leaves.size()) return failf(-9, "CW: %g leaf records for %g reference leaves", (double)nl, (double)leaves.size());
		std::map<uint32_t, int> cl;
		for (size_t i = 0; i < nl; i++)
		{
			const F4* lf = &cw.leaves[i * 2];
			uint32_t key = fbits(lf[0].w);
			auto it = byKey.find(key);
			if (it == byKey.end()) return failf(-10, "CW: leaf record %g has an unknown triangle range", (double)i);
			const RefLeaf* L = it->second;
			if (lf[0].x != L->bmin[0] || lf[0].y != L->bmin[1] || lf[0].z != L->bmin[2] || lf[1].x != L->bmax[0] || lf[1].y != L->bmax[1] || lf[1].z != L->bmax[2])
				return failf(-11, "CW: leaf record %g does not carry the exact reference leaf box", (double)i);
			cl[key]++;
		}
		if (cl.size() != leaves.size()) return failf(-12, "CW: duplicate leaf records");
		// exact subtree box of every CW node, bottom-up (children have larger indices: breadth-first)
		std::vector<BoxD> ex(nn);
		std::vector<char> isChild(nn, 0);
		for (size_t ii = nn; ii-- > 0;)
		{
			const F4* nd = &cw.nodes[ii * 5];
			uint32_t meta = fbits(nd[0].w), imask = meta >> 24, lmask = fbits(nd[1].z) & 0xFFu;
			uint32_t childBase = fbits(nd[1].x), leafBase = fbits(nd[1].y);
			if (imask & lmask) return failf(-13, "CW: node %g has a slot that is both node and leaf", (double)ii);
			double e[3] = {std::ldexp(1.0, (int)(meta & 0xFF) - 127), std::ldexp(1.0, (int)((meta >> 8) & 0xFF) - 127), std::ldexp(1.0, (int)((meta >> 16) & 0xFF) - 127)};
			double p[3] = {nd[0].x, nd[0].y, nd[0].z};
			uint32_t words[12] = {fbits(nd[2].x), fbits(nd[2].y), fbits(nd[2].z), fbits(nd[2].w), fbits(nd[3].x), fbits(nd[3].y),
			                      fbits(nd[3].z), fbits(nd[3].w), fbits(nd[4].x), fbits(nd[4].y), fbits(nd[4].z), fbits(nd[4].w)};
			auto q = [&](int plane, int slot) { return (double)((words[plane * 2 + slot / 4] >> (8 * (slot & 3))) & 0xFF); };
			BoxD me;
			for (int k = 0; k < 3; k++) me.mn[k] = 1e300, me.mx[k] = -1e300;
			uint32_t ci = 0, li = 0;
			for (int s = 0; s < 8; s++)
			{
				BoxD b;
				if (imask >> s & 1)
				{
					uint32_t c = childBase + ci++;
					if (c <= ii || c >= nn) return failf(-14, "CW: node %g child index %g out of order", (double)ii, (double)c);
					isChild[c]++;
					b = ex[c];
				}
				else if (lmask >> s & 1)
				{
					uint32_t l = leafBase + li++;
					if (l >= nl) return failf(-15, "CW: leaf index out of range");
					const F4* lf = &cw.leaves[(size_t)l * 2];
					b.mn[0] = lf[0].x, b.mn[1] = lf[0].y, b.mn[2] = lf[0].z, b.mx[0] = lf[1].x, b.mx[1] = lf[1].y, b.mx[2] = lf[1].z;
				}
				else
					continue;
				double vq = 1, ve = 1;
				for (int k = 0; k < 3; k++)
				{
					double lo = p[k] + q(k, s) * e[k], hi = p[k] + q(3 + k, s) * e[k];
					// conservative with at least one full step of margin per side
					if (!(lo <= b.mn[k] - e[k] * 0.999999) || !(hi >= b.mx[k] + e[k] * 0.999999))
						return failf(-16, "CW: node %g slot %g axis %g: quantised box does not contain the exact box with one step of margin", (double)ii, s, k);
					vq *= (hi - lo), ve *= std::max(b.mx[k] - b.mn[k], e[k]);
					me.mn[k] = std::min(me.mn[k], b.mn[k]), me.mx[k] = std::max(me.mx[k], b.mx[k]);
				}
				cwInfl += vq / ve, cwCount++;
			}
			ex[ii] = me;
		}
		for (size_t i = 1; i < nn; i++)
			if (isChild[i] != 1) return failf(-17, "CW: node %g has %g parents", (double)i, isChild[i]);
	}
	stats[7] = cwCount ? cwInfl / (double)cwCount : 0;
	// ---- Q16
	{
		Q16Tree q16;
		buildQ16(fast, q16);
		stats[6] = (double)(q16.leaves.size() / 2);
		if (q16.nodes.size() / 2 != nf) return failf(-18, "Q16: node count differs from FAST");
		if (q16.leaves.size() / 2 != leaves.size()) return failf(-19, "Q16: %g leaf records", (double)(q16.leaves.size() / 2));
		double infl = 0;
		for (size_t i = 0; i < nf; i++)
		{
			const F4* fn = &fast.nodes[i * 4];
			float mn[2][3] = {{fn[0].x, fn[0].z, fn[2].x}, {fn[1].x, fn[1].z, fn[2].z}};
			float mx[2][3] = {{fn[0].y, fn[0].w, fn[2].y}, {fn[1].y, fn[1].w, fn[2].w}};
			int32_t fref[2] = {(int32_t)fbits(fn[3].x), (int32_t)fbits(fn[3].y)};
			for (int c = 0; c < 2; c++)
			{
				const F4& qn = q16.nodes[i * 2 + c];
				uint32_t w[3] = {fbits(qn.x), fbits(qn.y), fbits(qn.z)};
				int32_t qref = (int32_t)fbits(qn.w);
				if ((fref[c] >= 0) != (qref >= 0) || (fref[c] >= 0 && qref != fref[c])) return failf(-20, "Q16: node %g child reference differs from FAST", (double)i);
				if (qref < 0)
				{
					const F4* lf = &q16.leaves[(size_t)(uint32_t)(~qref) * 2];
					if (fbits(lf[0].w) != (uint32_t)(~fref[c])) return failf(-21, "Q16: leaf record of node %g has the wrong triangle range", (double)i);
					if (lf[0].x != mn[c][0] || lf[0].y != mn[c][1] || lf[0].z != mn[c][2] || lf[1].x != mx[c][0] || lf[1].y != mx[c][1] || lf[1].z != mx[c][2])
						return failf(-22, "Q16: leaf record of node %g does not carry the exact box", (double)i);
				}
				for (int k = 0; k < 3; k++)
				{
					double step = q16.qstep[k], lo = (double)q16.qmin[k] + (double)(w[k] & 0xFFFF) * step, hi = (double)q16.qmin[k] + (double)(w[k] >> 16) * step;
					if (!(lo <= (double)mn[c][k] - 1.999999 * step) || !(hi >= (double)mx[c][k] + 1.999999 * step))
						return failf(-23, "Q16: node %g child %g axis %g: quantised box lacks two steps of margin", (double)i, c, k);
					infl += (hi - lo) / std::max((double)mx[c][k] - (double)mn[c][k], step);
				}
			}
		}
		stats[8] = infl / (double)(nf * 6);
	}
	return 0;
}

// skip links of the EXACT traversal: out[i] = first node index after i's subtree (buildSkipLinks' one forward pass);
// returns the number of leaves, or -1 with the builder's message if the node array is rejected
extern "C" int accel_skip_links(const rtb_ref_node* nodes, uint32_t n, uint32_t nTris, uint32_t* out)
{
	g_msg.clear();
	std::vector<uint32_t> skip;
	std::vector<RefLeaf> leaves;
	const char* err = nullptr;
	if (!buildSkipLinks(nodes, n, nTris, skip, leaves, &err)) return failf(-1, err);
	for (uint32_t i = 0; i < n; i++) out[i] = skip[i];
	return (int)leaves.size();
}
