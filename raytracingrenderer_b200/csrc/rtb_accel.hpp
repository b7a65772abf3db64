// rtb_accel.hpp — host-side builders run by rtb_upload_scene:
//   * skip links for the stack-free EXACT traversal of the reference's tree,
//   * the FAST tree: a binned-SAH binary BVH whose primitives are the reference's LEAVES
//     (<= 2 triangles each, RTBase/Geometry.h:240,337) with their exact AABBs, so that the
//     box test that admits a leaf is the reference's own leaf test,
//   * the environment-map sampling tables.
#pragma once
#include "../../include/rtb.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace rtb_accel
{

struct F4
{
	float x, y, z, w;
};
inline float bitsToFloat(uint32_t u)
{
	float f;
	memcpy(&f, &u, 4);
	return f;
}

// xnodes[2i] = bmin, bits(skip); xnodes[2i+1] = bmax, bits(leaf).  Also checks that the
// array really is a pre-order tree (left child = i + 1) and collects the leaves.
struct RefLeaf
{
	float bmin[3], bmax[3];
	uint32_t start, count;
};

inline bool buildExact(const rtb_ref_node* nodes, uint32_t n, uint32_t nTris, std::vector<F4>& xnodes,
                       std::vector<RefLeaf>& leaves, const char** err)
{
	xnodes.resize((size_t)n * 2);
	leaves.clear();
	if (n == 0) return true;
	std::vector<uint32_t> skip(n, 0);
	// iterative post-order: skip[i] = first index after i's subtree
	struct Item
	{
		uint32_t node, end;
	};
	std::vector<Item> stack;
	stack.push_back({0u, n});
	while (!stack.empty())
	{
		Item it = stack.back();
		stack.pop_back();
		const rtb_ref_node& nd = nodes[it.node];
		skip[it.node] = it.end;
		if (nd.a < 0) continue;
		if ((uint32_t)nd.a != it.node + 1 || (uint32_t)nd.b <= it.node + 1 || (uint32_t)nd.b >= it.end)
		{
			*err = "ref_nodes is not a pre-order tree";
			return false;
		}
		stack.push_back({(uint32_t)nd.a, (uint32_t)nd.b});
		stack.push_back({(uint32_t)nd.b, it.end});
	}
	for (uint32_t i = 0; i < n; i++)
	{
		const rtb_ref_node& nd = nodes[i];
		uint32_t leaf = 0xFFFFFFFFu;
		if (nd.a < 0)
		{
			uint32_t start = (uint32_t)(~nd.a), count = (uint32_t)nd.b;
			if (count > 3u || start >= (1u << 30) || (uint64_t)start + count > nTris)
			{
				*err = "ref leaf out of range (count > 3 or start + count > n_tris)";
				return false;
			}
			leaf = (start << 2) | count;
			RefLeaf L;
			memcpy(L.bmin, nd.bmin, 12);
			memcpy(L.bmax, nd.bmax, 12);
			L.start = start, L.count = count;
			if (count) leaves.push_back(L);
		}
		xnodes[(size_t)i * 2] = {nd.bmin[0], nd.bmin[1], nd.bmin[2], bitsToFloat(skip[i])};
		xnodes[(size_t)i * 2 + 1] = {nd.bmax[0], nd.bmax[1], nd.bmax[2], bitsToFloat(leaf)};
	}
	return true;
}

// ---------------------------------------------------------------------------------------
// FAST tree
// ---------------------------------------------------------------------------------------
struct Box
{
	float mn[3], mx[3];
	void reset()
	{
		mn[0] = mn[1] = mn[2] = FLT_MAX;
		mx[0] = mx[1] = mx[2] = -FLT_MAX;
	}
	void grow(const float* a, const float* b)
	{
		for (int k = 0; k < 3; k++)
		{
			if (a[k] < mn[k]) mn[k] = a[k];
			if (b[k] > mx[k]) mx[k] = b[k];
		}
	}
	void grow(const Box& o) { grow(o.mn, o.mx); }
	float area() const
	{
		float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
		if (dx < 0 || dy < 0 || dz < 0) return 0.0f;
		return 2.0f * (dx * dy + dy * dz + dz * dx);
	}
};

struct FastTree
{
	std::vector<F4> nodes; // 4 x F4 per node (layout in rtb_dev_scene.cuh)
	int32_t root = 0;      // child reference of the root
	uint32_t maxDepth = 0;
};

class FastBuilder
{
public:
	FastBuilder(const std::vector<RefLeaf>& leaves) : L(leaves) {}

	void build(FastTree& out)
	{
		out.nodes.clear();
		out.maxDepth = 0;
		uint32_t n = (uint32_t)L.size();
		if (n == 0)
		{
			out.root = ~0; // empty leaf (start 0, count 0)
			return;
		}
		idx.resize(n);
		cen.resize((size_t)n * 3);
		for (uint32_t i = 0; i < n; i++)
		{
			idx[i] = i;
			for (int k = 0; k < 3; k++) cen[(size_t)i * 3 + k] = 0.5f * (L[i].bmin[k] + L[i].bmax[k]);
		}
		if (n == 1)
		{
			out.root = leafRef(0);
			return;
		}
		out.nodes.reserve((size_t)(n - 1) * 4);
		unsigned hw = std::thread::hardware_concurrency();
		int par = 0; // top levels of the recursion that build the two halves concurrently
		while ((1u << par) < hw && par < 5) par++;
		if (n < 100000) par = 0;
		Box b;
		uint32_t depthSeen = 0;
		out.root = recurse(out.nodes, 0, n, 0, b, par, depthSeen);
		out.maxDepth = depthSeen;
	}

private:
	static const int BINS = 32;
	static const uint32_t MAX_SAH_DEPTH = 40; // deeper than this: median splits (bounded stack on device)
	const std::vector<RefLeaf>& L;
	std::vector<uint32_t> idx;
	std::vector<float> cen;

	int32_t leafRef(uint32_t prim) const { return ~(int32_t)((L[prim].start << 2) | L[prim].count); }

	static void writeNode(F4* nd, const Box& b0, const Box& b1, int32_t c0, int32_t c1)
	{
		nd[0] = {b0.mn[0], b0.mx[0], b0.mn[1], b0.mx[1]};
		nd[1] = {b1.mn[0], b1.mx[0], b1.mn[1], b1.mx[1]};
		nd[2] = {b0.mn[2], b0.mx[2], b1.mn[2], b1.mx[2]};
		nd[3] = {bitsToFloat((uint32_t)c0), bitsToFloat((uint32_t)c1), 0.0f, 0.0f};
	}
	static uint32_t floatBits(float f)
	{
		uint32_t u;
		memcpy(&u, &f, 4);
		return u;
	}

	// Builds the subtree of idx[lo, hi) into `nodes` (indices relative to that vector); returns the
	// child reference and the exact box of the subtree (= union of leaf boxes).  With par > 0 the
	// two halves are built concurrently into private vectors (disjoint slices of idx) and spliced
	// in afterwards: the tree does not depend on the thread count.
	int32_t recurse(std::vector<F4>& nodes, uint32_t lo, uint32_t hi, uint32_t depth, Box& box, int par, uint32_t& depthSeen)
	{
		if (depth > depthSeen) depthSeen = depth;
		uint32_t n = hi - lo;
		box.reset();
		if (n == 1)
		{
			box.grow(L[idx[lo]].bmin, L[idx[lo]].bmax);
			return leafRef(idx[lo]);
		}
		uint32_t mid = split(lo, hi, depth);
		size_t self = nodes.size() / 4;
		nodes.resize(nodes.size() + 4);
		Box b0, b1;
		int32_t c0, c1;
		if (par > 0 && n > 50000)
		{
			std::vector<F4> left, right;
			uint32_t dl = 0, dr = 0;
			std::thread th([&]() { c0 = recurse(left, lo, mid, depth + 1, b0, par - 1, dl); });
			c1 = recurse(right, mid, hi, depth + 1, b1, par - 1, dr);
			th.join();
			if (dl > depthSeen) depthSeen = dl;
			if (dr > depthSeen) depthSeen = dr;
			auto splice = [&](std::vector<F4>& sub, int32_t& ref) {
				int32_t off = (int32_t)(nodes.size() / 4);
				for (size_t i = 0; i < sub.size(); i += 4)
				{
					int32_t a = (int32_t)floatBits(sub[i + 3].x), b = (int32_t)floatBits(sub[i + 3].y);
					if (a >= 0) sub[i + 3].x = bitsToFloat((uint32_t)(a + off));
					if (b >= 0) sub[i + 3].y = bitsToFloat((uint32_t)(b + off));
				}
				if (ref >= 0) ref += off;
				nodes.insert(nodes.end(), sub.begin(), sub.end());
				std::vector<F4>().swap(sub);
			};
			splice(left, c0);
			splice(right, c1);
		}
		else
		{
			c0 = recurse(nodes, lo, mid, depth + 1, b0, 0, depthSeen);
			c1 = recurse(nodes, mid, hi, depth + 1, b1, 0, depthSeen);
		}
		writeNode(&nodes[self * 4], b0, b1, c0, c1);
		box.grow(b0);
		box.grow(b1);
		return (int32_t)self;
	}

	uint32_t split(uint32_t lo, uint32_t hi, uint32_t depth)
	{
		uint32_t n = hi - lo;
		if (n == 2) return lo + 1;
		// centroid bounds
		float cmn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, cmx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
		for (uint32_t i = lo; i < hi; i++)
		{
			const float* c = &cen[(size_t)idx[i] * 3];
			for (int k = 0; k < 3; k++)
			{
				if (c[k] < cmn[k]) cmn[k] = c[k];
				if (c[k] > cmx[k]) cmx[k] = c[k];
			}
		}
		int bestAxis = -1, bestBin = -1;
		float bestCost = FLT_MAX;
		if (depth < MAX_SAH_DEPTH)
		{
			for (int ax = 0; ax < 3; ax++)
			{
				float ext = cmx[ax] - cmn[ax];
				if (!(ext > 0.0f)) continue;
				float scale = (float)BINS / ext;
				Box bb[BINS];
				uint32_t cnt[BINS];
				for (int b = 0; b < BINS; b++)
				{
					bb[b].reset();
					cnt[b] = 0;
				}
				for (uint32_t i = lo; i < hi; i++)
				{
					uint32_t p = idx[i];
					int b = (int)((cen[(size_t)p * 3 + ax] - cmn[ax]) * scale);
					if (b >= BINS) b = BINS - 1;
					if (b < 0) b = 0;
					bb[b].grow(L[p].bmin, L[p].bmax);
					cnt[b] += L[p].count; // cost weight: triangles in the leaf
				}
				float rightArea[BINS];
				uint32_t rightCnt[BINS];
				Box acc;
				acc.reset();
				uint32_t c = 0;
				for (int b = BINS - 1; b > 0; b--)
				{
					acc.grow(bb[b]);
					c += cnt[b];
					rightArea[b] = acc.area();
					rightCnt[b] = c;
				}
				acc.reset();
				c = 0;
				for (int b = 0; b < BINS - 1; b++)
				{
					acc.grow(bb[b]);
					c += cnt[b];
					if (c == 0 || rightCnt[b + 1] == 0) continue;
					float cost = acc.area() * (float)c + rightArea[b + 1] * (float)rightCnt[b + 1];
					if (cost < bestCost)
					{
						bestCost = cost;
						bestAxis = ax;
						bestBin = b;
					}
				}
			}
		}
		if (bestAxis >= 0)
		{
			float ext = cmx[bestAxis] - cmn[bestAxis];
			float scale = (float)BINS / ext;
			float base = cmn[bestAxis];
			int ax = bestAxis, bin = bestBin;
			uint32_t* first = &idx[lo];
			uint32_t* last = &idx[hi];
			uint32_t* m = std::partition(first, last, [&](uint32_t p) {
				int b = (int)((cen[(size_t)p * 3 + ax] - base) * scale);
				if (b >= BINS) b = BINS - 1;
				if (b < 0) b = 0;
				return b <= bin;
			});
			uint32_t mid = (uint32_t)(m - &idx[0]);
			if (mid > lo && mid < hi) return mid;
		}
		// fallback: median split along the widest centroid axis (or by index if all equal)
		int ax = 0;
		float e0 = cmx[0] - cmn[0], e1 = cmx[1] - cmn[1], e2 = cmx[2] - cmn[2];
		if (e1 > e0 && e1 >= e2) ax = 1;
		else if (e2 > e0 && e2 > e1) ax = 2;
		uint32_t mid = lo + n / 2;
		std::nth_element(&idx[lo], &idx[mid], &idx[0] + hi, [&](uint32_t a, uint32_t b) {
			float ca = cen[(size_t)a * 3 + ax], cb = cen[(size_t)b * 3 + ax];
			return ca < cb || (ca == cb && a < b);
		});
		return mid;
	}
};


// ---------------------------------------------------------------------------------------
// WIDE tree: the FAST binary tree collapsed to 4 children per node (the child with the largest
// surface area is replaced by its own two children until there are four).  Same boxes, same
// leaves; half the dependent node fetches per ray.  128-byte nodes, structure of arrays:
//   [0] min.x[4] [1] max.x[4] [2] min.y[4] [3] max.y[4] [4] min.z[4] [5] max.z[4]
//   [6] child references [4] (>= 0 node, < 0 leaf, RTB_WIDE_EMPTY unused slot) [7] unused
// ---------------------------------------------------------------------------------------
#define RTB_WIDE_EMPTY 0x7FFFFFFE

struct WideTree
{
	std::vector<F4> nodes; // 8 x F4 per node
	int32_t root = 0;
	uint32_t maxDepth = 0;
};

class WideBuilder
{
public:
	WideBuilder(const FastTree& t) : B(t) {}
	void build(WideTree& out)
	{
		out.nodes.clear();
		out.maxDepth = 0;
		W = &out;
		if (B.root < 0 || B.nodes.empty())
		{
			out.root = B.root; // single leaf / empty scene
			return;
		}
		out.nodes.reserve(B.nodes.size() * 2 / 3 + 8);
		out.root = emit(B.root, 0);
	}

private:
	const FastTree& B;
	WideTree* W = nullptr;
	struct Cand
	{
		int32_t ref;
		float mn[3], mx[3];
		float area() const
		{
			float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
			return 2.0f * (dx * dy + dy * dz + dz * dx);
		}
	};
	static uint32_t bits(float f)
	{
		uint32_t u;
		memcpy(&u, &f, 4);
		return u;
	}
	void children(int32_t node, Cand& c0, Cand& c1) const
	{
		const F4* nd = &B.nodes[(size_t)node * 4];
		c0.ref = (int32_t)bits(nd[3].x), c1.ref = (int32_t)bits(nd[3].y);
		c0.mn[0] = nd[0].x, c0.mx[0] = nd[0].y, c0.mn[1] = nd[0].z, c0.mx[1] = nd[0].w;
		c1.mn[0] = nd[1].x, c1.mx[0] = nd[1].y, c1.mn[1] = nd[1].z, c1.mx[1] = nd[1].w;
		c0.mn[2] = nd[2].x, c0.mx[2] = nd[2].y, c1.mn[2] = nd[2].z, c1.mx[2] = nd[2].w;
	}
	int32_t emit(int32_t node, uint32_t depth)
	{
		if (depth > W->maxDepth) W->maxDepth = depth;
		Cand c[4];
		int n = 2;
		children(node, c[0], c[1]);
		while (n < 4)
		{
			int best = -1;
			float bestArea = -1.0f;
			for (int i = 0; i < n; i++)
				if (c[i].ref >= 0 && c[i].area() > bestArea) bestArea = c[i].area(), best = i;
			if (best < 0) break;
			Cand a, b;
			children(c[best].ref, a, b);
			c[best] = a;
			c[n++] = b;
		}
		size_t self = W->nodes.size() / 8;
		W->nodes.resize(W->nodes.size() + 8);
		int32_t refs[4];
		for (int i = 0; i < 4; i++)
		{
			if (i >= n) refs[i] = RTB_WIDE_EMPTY;
			else refs[i] = (c[i].ref >= 0) ? emit(c[i].ref, depth + 1) : c[i].ref;
		}
		F4* nd = &W->nodes[self * 8];
		float v[6][4];
		for (int i = 0; i < 4; i++)
			for (int k = 0; k < 3; k++)
			{
				v[k * 2][i] = (i < n) ? c[i].mn[k] : FLT_MAX;
				v[k * 2 + 1][i] = (i < n) ? c[i].mx[k] : -FLT_MAX;
			}
		for (int r = 0; r < 6; r++) nd[r] = {v[r][0], v[r][1], v[r][2], v[r][3]};
		nd[6] = {bitsToFloat((uint32_t)refs[0]), bitsToFloat((uint32_t)refs[1]), bitsToFloat((uint32_t)refs[2]), bitsToFloat((uint32_t)refs[3])};
		nd[7] = {0, 0, 0, 0};
		return (int32_t)self;
	}
};

// ---------------------------------------------------------------------------------------
// Q16 tree (RTB_TRAV_Q16): the FAST tree, node for node, in 32 bytes: both child boxes on ONE 16-bit grid over the scene
// box (plane = qmin + q * qstep), rounded OUTWARDS by two steps — interior boxes only have to be conservative (SURVEY
// A.3); the exact box of every reference leaf moves to a 32-byte leaf record that the traversal tests with the
// reference's arithmetic before the leaf's triangles.  Layout: rtb_dev_scene.cuh (DevScene::qnodes / qleaves).
// ---------------------------------------------------------------------------------------
struct Q16Tree
{
	std::vector<F4> nodes;  // 2 x F4 per node, same indices as FastTree::nodes
	std::vector<F4> leaves; // 2 x F4 per leaf record
	int32_t root = 0;       // node index, or ~(leaf record) / ~0 for single-leaf / empty scenes
	float qmin[3] = {0, 0, 0}, qstep[3] = {1, 1, 1};
};

inline void buildQ16(const FastTree& B, Q16Tree& out)
{
	out.nodes.clear();
	out.leaves.clear();
	out.root = B.root;
	size_t n = B.nodes.size() / 4;
	if (B.root < 0 || n == 0) return; // the traversal falls back to the reference tree (travRoot < 0)
	auto fbits = [](float f) {
		uint32_t u;
		memcpy(&u, &f, 4);
		return u;
	};
	// scene box = union of the root's two child boxes
	double lo[3], hi[3];
	{
		const F4* r = &B.nodes[(size_t)B.root * 4];
		lo[0] = std::min(r[0].x, r[1].x), hi[0] = std::max(r[0].y, r[1].y);
		lo[1] = std::min(r[0].z, r[1].z), hi[1] = std::max(r[0].w, r[1].w);
		lo[2] = std::min(r[2].x, r[2].z), hi[2] = std::max(r[2].y, r[2].w);
	}
	for (int k = 0; k < 3; k++)
	{
		double ext = hi[k] - lo[k], mag = std::max(std::fabs(lo[k]), std::fabs(hi[k]));
		double step = std::max(std::max(ext / 65500.0, mag * 9.5367431640625e-07 /* 2^-20 */), 1e-30);
		out.qstep[k] = (float)step;
		if ((double)out.qstep[k] < step) out.qstep[k] = nextafterf(out.qstep[k], FLT_MAX);
		double m = lo[k] - 8.0 * (double)out.qstep[k];
		out.qmin[k] = (float)m;
		if ((double)out.qmin[k] > m) out.qmin[k] = nextafterf(out.qmin[k], -FLT_MAX);
	}
	auto qlo = [&](float v, int k) {
		double q = std::floor(((double)v - (double)out.qmin[k]) / (double)out.qstep[k]) - 2.0;
		return (uint32_t)std::min(std::max(q, 0.0), 65535.0);
	};
	auto qhi = [&](float v, int k) {
		double q = std::ceil(((double)v - (double)out.qmin[k]) / (double)out.qstep[k]) + 2.0;
		return (uint32_t)std::min(std::max(q, 0.0), 65535.0);
	};
	out.nodes.resize(n * 2);
	// leaf records in node order (a serial pass assigns their indices, the quantisation itself runs on all cores)
	std::vector<uint32_t> leafBase(n + 1, 0);
	for (size_t i = 0; i < n; i++)
	{
		const F4* nd = &B.nodes[i * 4];
		leafBase[i + 1] = leafBase[i] + ((int32_t)fbits(nd[3].x) < 0 ? 1u : 0u) + ((int32_t)fbits(nd[3].y) < 0 ? 1u : 0u);
	}
	out.leaves.resize((size_t)leafBase[n] * 2);
	unsigned hw = std::thread::hardware_concurrency();
	unsigned nth = n > 100000 ? std::min(std::max(hw, 1u), 32u) : 1u;
	auto work = [&](size_t a, size_t b) {
		for (size_t i = a; i < b; i++)
		{
			const F4* nd = &B.nodes[i * 4];
			float mn[2][3] = {{nd[0].x, nd[0].z, nd[2].x}, {nd[1].x, nd[1].z, nd[2].z}};
			float mx[2][3] = {{nd[0].y, nd[0].w, nd[2].y}, {nd[1].y, nd[1].w, nd[2].w}};
			int32_t ref[2] = {(int32_t)fbits(nd[3].x), (int32_t)fbits(nd[3].y)};
			uint32_t nl = leafBase[i];
			for (int c = 0; c < 2; c++)
			{
				uint32_t w[3];
				for (int k = 0; k < 3; k++) w[k] = qlo(mn[c][k], k) | (qhi(mx[c][k], k) << 16);
				int32_t r = ref[c];
				if (r < 0)
				{
					F4* lf = &out.leaves[(size_t)nl * 2];
					lf[0] = {mn[c][0], mn[c][1], mn[c][2], bitsToFloat((uint32_t)(~r))};
					lf[1] = {mx[c][0], mx[c][1], mx[c][2], 0.0f};
					r = ~(int32_t)nl;
					nl++;
				}
				out.nodes[i * 2 + c] = {bitsToFloat(w[0]), bitsToFloat(w[1]), bitsToFloat(w[2]), bitsToFloat((uint32_t)r)};
			}
		}
	};
	if (nth <= 1) work(0, n);
	else
	{
		std::vector<std::thread> th;
		for (unsigned t = 0; t < nth; t++) th.emplace_back(work, n * t / nth, n * (t + 1) / nth);
		for (std::thread& t : th) t.join();
	}
}

// ---------------------------------------------------------------------------------------
// Environment-map sampling tables (RTB_SAMPLING_IMPORTANCE).  EnvironmentMap::evaluate
// (RTBase/Lights.h:158-165) maps a direction to u = phi/2pi, v = theta/pi and
// Texture::sample (Imaging.h:72-94) blends texels floor(u W), floor(u W)+1 (no half-texel
// shift), so the radiance inside cell (x, y) = [x/W,(x+1)/W) x [y/H,(y+1)/H) depends on
// texels (x..x+1, y..y+1).  Cell weight = max luminance of those four texels x sin(theta)
// at the cell centre, plus a floor so that the pdf is positive wherever radiance can be:
// the estimator's expectation is unchanged for ANY such positive density.
// ---------------------------------------------------------------------------------------
inline void buildEnvTables(const float* texels, int W, int H, std::vector<float>& marginal, std::vector<float>& cond)
{
	std::vector<double> w((size_t)W * H);
	double sum = 0.0;
	auto lumAt = [&](int x, int y) {
		const float* p = texels + ((size_t)(y % H) * W + (size_t)(x % W)) * 3;
		return 0.2126 * p[0] + 0.7152 * p[1] + 0.0722 * p[2];
	};
	for (int y = 0; y < H; y++)
	{
		double st = sin(((double)y + 0.5) / (double)H * 3.14159265358979323846);
		for (int x = 0; x < W; x++)
		{
			double l = std::max(std::max(lumAt(x, y), lumAt(x + 1, y)), std::max(lumAt(x, y + 1), lumAt(x + 1, y + 1)));
			if (!(l > 0.0)) l = 0.0;
			w[(size_t)y * W + x] = l * st;
			sum += l * st;
		}
	}
	double mean = sum / ((double)W * H);
	double floorW = (mean > 0.0) ? 0.05 * mean : 1.0;
	marginal.assign((size_t)H + 1, 0.0f);
	cond.assign((size_t)H * (W + 1), 0.0f);
	std::vector<double> rowSum(H);
	double total = 0.0;
	for (int y = 0; y < H; y++)
	{
		double st = sin(((double)y + 0.5) / (double)H * 3.14159265358979323846);
		double rs = 0.0;
		for (int x = 0; x < W; x++)
		{
			w[(size_t)y * W + x] += floorW * st;
			rs += w[(size_t)y * W + x];
		}
		rowSum[y] = rs;
		total += rs;
	}
	double acc = 0.0;
	for (int y = 0; y < H; y++)
	{
		marginal[y] = (float)(acc / total);
		acc += rowSum[y];
		double ca = 0.0;
		float* cd = &cond[(size_t)y * (W + 1)];
		for (int x = 0; x < W; x++)
		{
			cd[x] = (float)(ca / rowSum[y]);
			ca += w[(size_t)y * W + x];
		}
		cd[W] = 1.0f;
	}
	marginal[H] = 1.0f;
	// make the float CDFs non-decreasing (rounding)
	for (int y = 1; y <= H; y++)
		if (marginal[y] < marginal[y - 1]) marginal[y] = marginal[y - 1];
	for (int y = 0; y < H; y++)
	{
		float* cd = &cond[(size_t)y * (W + 1)];
		for (int x = 1; x <= W; x++)
			if (cd[x] < cd[x - 1]) cd[x] = cd[x - 1];
	}
}

} // namespace rtb_accel
