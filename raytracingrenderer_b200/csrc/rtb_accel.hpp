// rtb_accel.hpp — host-side builders run by rtb_upload_scene:
//   * skip links for the stack-free EXACT traversal of the reference's tree,
//   * the FAST tree: a binned-SAH binary BVH whose primitives are the reference's LEAVES
//     (<= 2 triangles each, RTBase/Geometry.h:240,337) with their exact AABBs, so that the
//     box test that admits a leaf is the reference's own leaf test,
//   * the environment-map sampling tables.
#pragma once
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
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

// One forward pass over the pre-order node array: skip[i] = first index after i's subtree (a parent precedes its
// children, so skip[i] is final when i is reached), the leaves in index order, and the range checks.
inline bool buildSkipLinks(const rtb_ref_node* nodes, uint32_t n, uint32_t nTris, std::vector<uint32_t>& skip,
                           std::vector<RefLeaf>& leaves, const char** err)
{
	skip.assign(n, 0);
	leaves.clear();
	if (n == 0) return true;
	leaves.reserve((size_t)n / 2 + 1);
	skip[0] = n;
	for (uint32_t i = 0; i < n; i++)
	{
		const rtb_ref_node& nd = nodes[i];
		const uint32_t end = skip[i];
		if (end == 0)
		{
			*err = "ref_nodes is not a pre-order tree"; // a node no parent points to
			return false;
		}
		if (nd.a < 0)
		{
			uint32_t start = (uint32_t)(~nd.a), count = (uint32_t)nd.b;
			if (count > 3u || start >= (1u << 30) || (uint64_t)start + count > nTris)
			{
				*err = "ref leaf out of range (count > 3 or start + count > n_tris)";
				return false;
			}
			if (end != i + 1)
			{
				*err = "ref_nodes is not a pre-order tree";
				return false;
			}
			if (count)
			{
				RefLeaf L;
				memcpy(L.bmin, nd.bmin, 12);
				memcpy(L.bmax, nd.bmax, 12);
				L.start = start, L.count = count;
				leaves.push_back(L);
			}
			continue;
		}
		if ((uint32_t)nd.a != i + 1 || (uint32_t)nd.b <= i + 1 || (uint32_t)nd.b >= end)
		{
			*err = "ref_nodes is not a pre-order tree";
			return false;
		}
		skip[(uint32_t)nd.a] = (uint32_t)nd.b;
		skip[(uint32_t)nd.b] = end;
	}
	return true;
}

// The EXACT traversal's node pair [bmin, skip][bmax, leaf] (the device assembles the same pair from ref_nodes + skip:
// k_exact_nodes, rtb_api.cu)
inline void exactNodePair(const rtb_ref_node& nd, uint32_t skip, F4* out)
{
	uint32_t leaf = 0xFFFFFFFFu;
	if (nd.a < 0) leaf = ((uint32_t)(~nd.a) << 2) | (uint32_t)nd.b;
	out[0] = {nd.bmin[0], nd.bmin[1], nd.bmin[2], bitsToFloat(skip)};
	out[1] = {nd.bmax[0], nd.bmax[1], nd.bmax[2], bitsToFloat(leaf)};
}

inline bool buildExact(const rtb_ref_node* nodes, uint32_t n, uint32_t nTris, std::vector<F4>& xnodes,
                       std::vector<RefLeaf>& leaves, const char** err)
{
	std::vector<uint32_t> skip;
	xnodes.clear();
	if (!buildSkipLinks(nodes, n, nTris, skip, leaves, err)) return false;
	xnodes.resize((size_t)n * 2);
	for (uint32_t i = 0; i < n; i++) exactNodePair(nodes[i], skip[i], &xnodes[(size_t)i * 2]);
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
		unsigned hw = std::thread::hardware_concurrency();
		int par = 0; // top levels of the recursion that build the two halves concurrently
		while ((1u << par) < hw && par < 5) par++;
		if (n < 16384) par = 0;
		if (n == 1)
		{
			out.root = leafRef(L[0]);
			return;
		}
		// The builder permutes the 32-byte leaf records themselves, not indices into them: a subtree's leaves are
		// contiguous in memory, so the lower levels (most of the work) run out of the caches.
		items.resize(n);
		parallelFor(n >= PAR_MIN ? (1u << par) : 1u, 0, n, [&](unsigned, uint32_t a, uint32_t b) { memcpy(&items[a], &L[a], (size_t)(b - a) * sizeof(RefLeaf)); });
		// a binary tree over n leaves has n - 1 nodes and the subtree of idx[lo, hi) is laid out in pre-order from a
		// known base: every thread writes its nodes in place, nothing is spliced afterwards
		out.nodes.resize((size_t)(n - 1) * 4);
		if (n >= PAR_MIN) tmp.resize(n);
		Box b;
		uint32_t depthSeen = 0;
		out.root = recurse(out.nodes.data(), 0, 0, n, 0, b, par, depthSeen);
		out.maxDepth = depthSeen;
		std::vector<RefLeaf>().swap(tmp);
		std::vector<RefLeaf>().swap(items);
	}

private:
	static const int BINS = 32;
	static const uint32_t MAX_SAH_DEPTH = 40; // deeper than this: median splits (bounded stack on device)
	// From this many leaves on a node's split is itself computed by several threads (binning per chunk, merged; stable
	// partition through `tmp`).  The threshold depends on the node only: the tree is the same for any thread count.
	static const uint32_t PAR_MIN = 1u << 18;
	static const uint32_t SMALL = 32; // nodes of up to this many leaves: candidates from the sorted leaves, no bin arrays
	const std::vector<RefLeaf>& L;
	std::vector<RefLeaf> items, tmp;

	static int32_t leafRef(const RefLeaf& l) { return ~(int32_t)((l.start << 2) | l.count); }
	static float centroid(const RefLeaf& l, int k) { return 0.5f * (l.bmin[k] + l.bmax[k]); }

	template <class F>
	static void parallelFor(unsigned threads, uint32_t lo, uint32_t hi, F fn)
	{
		if (threads <= 1 || hi - lo < 2 * threads)
		{
			fn(0u, lo, hi);
			return;
		}
		std::vector<std::thread> th;
		uint64_t n = hi - lo;
		for (unsigned t = 1; t < threads; t++)
			th.emplace_back([=]() { fn(t, lo + (uint32_t)(n * t / threads), lo + (uint32_t)(n * (t + 1) / threads)); });
		fn(0u, lo, lo + (uint32_t)(n / threads));
		for (auto& x : th) x.join();
	}

	static void writeNode(F4* nd, const Box& b0, const Box& b1, int32_t c0, int32_t c1)
	{
		nd[0] = {b0.mn[0], b0.mx[0], b0.mn[1], b0.mx[1]};
		nd[1] = {b1.mn[0], b1.mx[0], b1.mn[1], b1.mx[1]};
		nd[2] = {b0.mn[2], b0.mx[2], b1.mn[2], b1.mx[2]};
		nd[3] = {bitsToFloat((uint32_t)c0), bitsToFloat((uint32_t)c1), 0.0f, 0.0f};
	}

	// Builds the subtree of items[lo, hi) into nodes[base ...] (pre-order: the node, its left subtree, its right
	// subtree); returns the child reference and the exact box of the subtree (= union of leaf boxes).  With par > 0
	// the two halves are built concurrently (disjoint slices of items and of the node array).
	int32_t recurse(F4* nodes, uint32_t base, uint32_t lo, uint32_t hi, uint32_t depth, Box& box, int par, uint32_t& depthSeen)
	{
		if (depth > depthSeen) depthSeen = depth;
		uint32_t n = hi - lo;
		box.reset();
		if (n == 1)
		{
			box.grow(items[lo].bmin, items[lo].bmax);
			return leafRef(items[lo]);
		}
		uint32_t mid = split(lo, hi, depth, par > 0 ? (1u << par) : 1u);
		const uint32_t self = base, baseL = base + 1, baseR = base + (mid - lo);
		Box b0, b1;
		int32_t c0, c1;
		if (par > 0 && n > 8192)
		{
			uint32_t dl = 0, dr = 0;
			std::thread th([&]() { c0 = recurse(nodes, baseL, lo, mid, depth + 1, b0, par - 1, dl); });
			c1 = recurse(nodes, baseR, mid, hi, depth + 1, b1, par - 1, dr);
			th.join();
			if (dl > depthSeen) depthSeen = dl;
			if (dr > depthSeen) depthSeen = dr;
		}
		else
		{
			c0 = recurse(nodes, baseL, lo, mid, depth + 1, b0, 0, depthSeen);
			c1 = recurse(nodes, baseR, mid, hi, depth + 1, b1, 0, depthSeen);
		}
		writeNode(&nodes[(size_t)self * 4], b0, b1, c0, c1);
		box.grow(b0);
		box.grow(b1);
		return (int32_t)self;
	}

	struct Bins
	{
		Box bb[3][BINS];
		uint32_t cnt[3][BINS];
		void reset()
		{
			for (int a = 0; a < 3; a++)
				for (int b = 0; b < BINS; b++) bb[a][b].reset(), cnt[a][b] = 0;
		}
	};
	static int binOf(float c, float base, float scale)
	{
		int b = (int)((c - base) * scale);
		if (b >= BINS) b = BINS - 1;
		if (b < 0) b = 0;
		return b;
	}

	uint32_t split(uint32_t lo, uint32_t hi, uint32_t depth, unsigned threads)
	{
		uint32_t n = hi - lo;
		if (n == 2) return lo + 1;
		const bool wide = n >= PAR_MIN; // chunked passes (results do not depend on `threads`)
		if (!wide) threads = 1;
		// centroid bounds
		float cmn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, cmx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
		if (threads <= 1)
		{
			for (uint32_t i = lo; i < hi; i++)
				for (int k = 0; k < 3; k++)
				{
					float c = centroid(items[i], k);
					if (c < cmn[k]) cmn[k] = c;
					if (c > cmx[k]) cmx[k] = c;
				}
		}
		else
		{
			std::vector<Box> part(threads);
			parallelFor(threads, lo, hi, [&](unsigned t, uint32_t a, uint32_t b) {
				Box bx;
				bx.reset();
				for (uint32_t i = a; i < b; i++)
				{
					const float c[3] = {centroid(items[i], 0), centroid(items[i], 1), centroid(items[i], 2)};
					bx.grow(c, c);
				}
				part[t] = bx;
			});
			for (unsigned t = 0; t < threads; t++)
				for (int k = 0; k < 3; k++)
				{
					if (part[t].mn[k] < cmn[k]) cmn[k] = part[t].mn[k];
					if (part[t].mx[k] > cmx[k]) cmx[k] = part[t].mx[k];
				}
		}
		int bestAxis = -1, bestBin = -1;
		float bestCost = FLT_MAX;
		if (depth < MAX_SAH_DEPTH && n <= SMALL)
		{
			// Half of all nodes hold a handful of leaves: resetting and sweeping 3 x 32 bins for them is most of the
			// build.  Same candidates, same costs, same order of evaluation from the leaves sorted by bin: a split
			// after an empty bin costs what the split after the previous occupied bin costs and never wins (strict <).
			for (int ax = 0; ax < 3; ax++)
			{
				float ext = cmx[ax] - cmn[ax];
				if (!(ext > 0.0f)) continue;
				float scale = (float)BINS / ext;
				uint8_t bin[SMALL], ord[SMALL];
				for (uint32_t i = 0; i < n; i++)
				{
					bin[i] = (uint8_t)binOf(centroid(items[lo + i], ax), cmn[ax], scale);
					uint32_t j = i;
					while (j > 0 && bin[ord[j - 1]] > bin[i]) ord[j] = ord[j - 1], j--;
					ord[j] = (uint8_t)i;
				}
				float rightArea[SMALL];
				uint32_t rightCnt[SMALL];
				Box acc;
				acc.reset();
				uint32_t c = 0;
				for (uint32_t k = n; k-- > 1;)
				{
					const RefLeaf& lf = items[lo + ord[k]];
					acc.grow(lf.bmin, lf.bmax);
					c += lf.count;
					rightArea[k] = acc.area();
					rightCnt[k] = c;
				}
				acc.reset();
				c = 0;
				for (uint32_t k = 0; k + 1 < n; k++)
				{
					const RefLeaf& lf = items[lo + ord[k]];
					acc.grow(lf.bmin, lf.bmax);
					c += lf.count;
					int b = bin[ord[k]];
					if (b == bin[ord[k + 1]] || b >= BINS - 1) continue; // not the last leaf of its bin / no split after the last bin
					if (c == 0 || rightCnt[k + 1] == 0) continue;
					float cost = acc.area() * (float)c + rightArea[k + 1] * (float)rightCnt[k + 1];
					if (cost < bestCost)
					{
						bestCost = cost;
						bestAxis = ax;
						bestBin = b;
					}
				}
			}
		}
		else if (depth < MAX_SAH_DEPTH)
		{
			// one pass bins all three axes (each leaf record is fetched once)
			float scale[3];
			bool use[3];
			for (int ax = 0; ax < 3; ax++)
			{
				float ext = cmx[ax] - cmn[ax];
				use[ax] = ext > 0.0f;
				scale[ax] = use[ax] ? (float)BINS / ext : 0.0f;
			}
			Bins one; // the common case (one thread) stays off the heap
			std::vector<Bins> more(threads > 1 ? threads : 0);
			Bins* part = threads > 1 ? more.data() : &one;
			parallelFor(threads, lo, hi, [&](unsigned t, uint32_t a, uint32_t b) {
				Bins& B = part[t];
				B.reset();
#if defined(__SSE2__)
				// same arithmetic four lanes at a time (min/max operand order = the scalar selects'; lane 3 is ignored)
				__m128 lo4[3][BINS], hi4[3][BINS];
				for (int ax = 0; ax < 3; ax++)
					for (int k = 0; k < BINS; k++) lo4[ax][k] = _mm_set1_ps(FLT_MAX), hi4[ax][k] = _mm_set1_ps(-FLT_MAX);
				const __m128 base = _mm_setr_ps(cmn[0], cmn[1], cmn[2], 0.0f), sc = _mm_setr_ps(scale[0], scale[1], scale[2], 0.0f);
				const __m128 half = _mm_set1_ps(0.5f);
				for (uint32_t i = a; i < b; i++)
				{
					const RefLeaf& lf = items[i];
					const __m128 mn = _mm_loadu_ps(lf.bmin);                                  // bmin.xyz, bmax.x
					const __m128 mx = _mm_loadu_ps(lf.bmax);                                  // bmax.xyz, start
					const __m128 cen = _mm_mul_ps(half, _mm_add_ps(mn, mx));
					alignas(16) int bin[4];
					_mm_store_si128((__m128i*)bin, _mm_cvttps_epi32(_mm_mul_ps(_mm_sub_ps(cen, base), sc)));
					for (int ax = 0; ax < 3; ax++)
					{
						if (!use[ax]) continue;
						int k = bin[ax];
						if (k >= BINS) k = BINS - 1;
						if (k < 0) k = 0;
						lo4[ax][k] = _mm_min_ps(mn, lo4[ax][k]);
						hi4[ax][k] = _mm_max_ps(mx, hi4[ax][k]);
						B.cnt[ax][k] += lf.count;
					}
				}
				for (int ax = 0; ax < 3; ax++)
					for (int k = 0; k < BINS; k++)
					{
						alignas(16) float l[4], h[4];
						_mm_store_ps(l, lo4[ax][k]), _mm_store_ps(h, hi4[ax][k]);
						for (int c = 0; c < 3; c++) B.bb[ax][k].mn[c] = l[c], B.bb[ax][k].mx[c] = h[c];
					}
#else
				for (uint32_t i = a; i < b; i++)
				{
					const RefLeaf& lf = items[i];
					for (int ax = 0; ax < 3; ax++)
					{
						if (!use[ax]) continue;
						int bin = binOf(centroid(lf, ax), cmn[ax], scale[ax]);
						B.bb[ax][bin].grow(lf.bmin, lf.bmax);
						B.cnt[ax][bin] += lf.count; // cost weight: triangles in the leaf
					}
				}
#endif
			});
			Bins& B = part[0];
			for (unsigned t = 1; t < threads; t++)
				for (int ax = 0; ax < 3; ax++)
					for (int b = 0; b < BINS; b++) B.bb[ax][b].grow(part[t].bb[ax][b]), B.cnt[ax][b] += part[t].cnt[ax][b];
			for (int ax = 0; ax < 3; ax++)
			{
				if (!use[ax]) continue;
				const Box* bb = B.bb[ax];
				const uint32_t* cnt = B.cnt[ax];
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
			auto goesLeft = [&](const RefLeaf& l) { return binOf(centroid(l, ax), base, scale) <= bin; };
			uint32_t mid;
			if (wide)
			{
				// stable partition through `tmp`: per-chunk counts, then every chunk scatters to its own offsets
				std::vector<uint32_t> nl(threads + 1, 0);
				parallelFor(threads, lo, hi, [&](unsigned t, uint32_t a, uint32_t b) {
					uint32_t c = 0;
					for (uint32_t i = a; i < b; i++) c += goesLeft(items[i]) ? 1u : 0u;
					nl[t + 1] = c;
				});
				for (unsigned t = 0; t < threads; t++) nl[t + 1] += nl[t];
				mid = lo + nl[threads];
				parallelFor(threads, lo, hi, [&](unsigned t, uint32_t a, uint32_t b) {
					uint32_t l = lo + nl[t], r = mid + (a - lo) - nl[t];
					for (uint32_t i = a; i < b; i++)
					{
						const RefLeaf& p = items[i];
						if (goesLeft(p)) tmp[l++] = p;
						else tmp[r++] = p;
					}
				});
				parallelFor(threads, lo, hi, [&](unsigned, uint32_t a, uint32_t b) { memcpy(&items[a], &tmp[a], (size_t)(b - a) * sizeof(RefLeaf)); });
			}
			else
			{
				RefLeaf* m = std::partition(&items[lo], &items[0] + hi, goesLeft);
				mid = (uint32_t)(m - &items[0]);
			}
			if (mid > lo && mid < hi) return mid;
		}
		// fallback: median split along the widest centroid axis (or by index if all equal)
		int ax = 0;
		float e0 = cmx[0] - cmn[0], e1 = cmx[1] - cmn[1], e2 = cmx[2] - cmn[2];
		if (e1 > e0 && e1 >= e2) ax = 1;
		else if (e2 > e0 && e2 > e1) ax = 2;
		uint32_t mid = lo + n / 2;
		std::nth_element(&items[lo], &items[mid], &items[0] + hi, [&](const RefLeaf& a, const RefLeaf& b) {
			float ca = centroid(a, ax), cb = centroid(b, ax);
			return ca < cb || (ca == cb && a.start < b.start); // leaves come in ascending triangle order
		});
		return mid;
	}
};


// ---------------------------------------------------------------------------------------
// Insertion-based optimisation of the FAST tree (after Bittner, Hapala, Havran: "Fast Insertion-Based Optimization of
// Bounding Volume Hierarchies", 2013).  One pass visits the interior nodes by decreasing surface area; a node is taken
// out together with its parent, and its two subtrees are re-inserted where they increase the summed surface area of the
// tree least (branch-and-bound search from the root).  Leaves, their exact boxes and the rule "a node's box is the exact
// union of its children's" are untouched, so the tree stays legal for hit-ID parity; only the topology above
// the leaves changes.  Opt-in (RTB_TREE_OPT=<passes>): profiles/r02_tree_opt.txt.
// ---------------------------------------------------------------------------------------
class FastOptimizer
{
public:
	struct ONode
	{
		float mn[3], mx[3];
		int32_t parent, c[2]; // c[0] < 0: a leaf (leafRef holds the reference)
		int32_t leafRef;
		float area;
	};
	std::vector<ONode> N;
	int32_t root = -1;
	uint32_t nInterior = 0;

	static float areaOf(const float* mn, const float* mx)
	{
		float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
		return 2.0f * (dx * dy + dy * dz + dz * dx);
	}
	static uint32_t bits(float f)
	{
		uint32_t u;
		memcpy(&u, &f, 4);
		return u;
	}
	void load(const FastTree& T)
	{
		nInterior = (uint32_t)(T.nodes.size() / 4);
		N.assign((size_t)nInterior * 2 + 1, ONode());
		uint32_t nextLeaf = nInterior;
		for (uint32_t i = 0; i < nInterior; i++)
		{
			const F4* nd = &T.nodes[(size_t)i * 4];
			int32_t ch[2] = {(int32_t)bits(nd[3].x), (int32_t)bits(nd[3].y)};
			float cmn[2][3] = {{nd[0].x, nd[0].z, nd[2].x}, {nd[1].x, nd[1].z, nd[2].z}};
			float cmx[2][3] = {{nd[0].y, nd[0].w, nd[2].y}, {nd[1].y, nd[1].w, nd[2].w}};
			for (int k = 0; k < 2; k++)
			{
				int32_t id = ch[k] >= 0 ? ch[k] : (int32_t)nextLeaf++;
				ONode& c = N[id];
				memcpy(c.mn, cmn[k], 12), memcpy(c.mx, cmx[k], 12);
				c.area = areaOf(c.mn, c.mx);
				c.parent = (int32_t)i;
				if (ch[k] < 0) c.c[0] = c.c[1] = -1, c.leafRef = ch[k];
				N[i].c[k] = id;
			}
		}
		root = T.root;
		N[root].parent = -1;
		refitOne(root);
	}
	void refitOne(int32_t i)
	{
		ONode& n = N[i];
		const ONode &a = N[n.c[0]], &b = N[n.c[1]];
		for (int k = 0; k < 3; k++) n.mn[k] = a.mn[k] < b.mn[k] ? a.mn[k] : b.mn[k], n.mx[k] = a.mx[k] > b.mx[k] ? a.mx[k] : b.mx[k];
		n.area = areaOf(n.mn, n.mx);
	}
	void refitUp(int32_t i)
	{
		while (i >= 0)
		{
			refitOne(i);
			i = N[i].parent;
		}
	}
	double cost() const
	{
		double s = 0;
		for (uint32_t i = 0; i < nInterior; i++) s += N[i].area;
		return s / N[root].area;
	}
	static float unionArea(const ONode& a, const ONode& b)
	{
		float mn[3], mx[3];
		for (int k = 0; k < 3; k++) mn[k] = a.mn[k] < b.mn[k] ? a.mn[k] : b.mn[k], mx[k] = a.mx[k] > b.mx[k] ? a.mx[k] : b.mx[k];
		return areaOf(mn, mx);
	}
	struct QE
	{
		float ci;
		int32_t node;
		bool operator<(const QE& o) const { return ci > o.ci; }
	};
	std::vector<QE> heap;
	int32_t findBest(int32_t x)
	{
		const ONode& X = N[x];
		float best = FLT_MAX;
		int32_t bestNode = root;
		heap.clear();
		heap.push_back({0.0f, root});
		while (!heap.empty())
		{
			std::pop_heap(heap.begin(), heap.end());
			QE e = heap.back();
			heap.pop_back();
			if (e.ci + X.area >= best) break;
			const ONode& Y = N[e.node];
			float direct = unionArea(Y, X);
			float total = e.ci + direct;
			if (total < best) best = total, bestNode = e.node;
			if (Y.c[0] >= 0)
			{
				float ci = e.ci + (direct - Y.area);
				if (ci + X.area < best)
				{
					heap.push_back({ci, Y.c[0]});
					std::push_heap(heap.begin(), heap.end());
					heap.push_back({ci, Y.c[1]});
					std::push_heap(heap.begin(), heap.end());
				}
			}
		}
		return bestNode;
	}
	void insertAt(int32_t x, int32_t y, int32_t fresh)
	{
		// `fresh` becomes the parent of (y, x) where y was
		int32_t p = N[y].parent;
		ONode& F = N[fresh];
		F.parent = p;
		F.c[0] = y, F.c[1] = x;
		N[y].parent = fresh, N[x].parent = fresh;
		if (p < 0) root = fresh;
		else N[p].c[N[p].c[0] == y ? 0 : 1] = fresh;
		refitUp(fresh);
	}
	bool process(int32_t n)
	{
		if (n == root) return false;
		int32_t p = N[n].parent;
		if (p < 0) return false;
		int32_t s = N[p].c[N[p].c[0] == n ? 1 : 0];
		int32_t g = N[p].parent;
		int32_t l = N[n].c[0], r = N[n].c[1];
		// unlink p and n: s takes p's place
		N[s].parent = g;
		if (g < 0) root = s;
		else
		{
			N[g].c[N[g].c[0] == p ? 0 : 1] = s;
			refitUp(g);
		}
		if (N[l].area < N[r].area) std::swap(l, r);
		insertAt(l, findBest(l), n);
		insertAt(r, findBest(r), p);
		return true;
	}
	// One pass over the `fraction` of the interior nodes with the largest boxes.  The re-insertions are greedy, not
	// monotone: a pass that ends with a larger summed area than it started with is undone (returns false).
	bool pass(float fraction)
	{
		const std::vector<ONode> before = N;
		const int32_t rootBefore = root;
		const double costBefore = cost();
		std::vector<std::pair<float, int32_t>> order;
		order.reserve(nInterior);
		for (uint32_t i = 0; i < nInterior; i++) order.push_back({-N[i].area, (int32_t)i});
		std::sort(order.begin(), order.end());
		if (!(fraction > 0.0f)) return true;
		size_t k = fraction >= 1.0f ? order.size() : (size_t)((double)order.size() * fraction);
		if (k > order.size()) k = order.size();
		for (size_t i = 0; i < k; i++) process(order[i].second);
		if (cost() <= costBefore) return true;
		N = before, root = rootBefore;
		return false;
	}
	// pre-order re-encoding
	uint32_t store(FastTree& T)
	{
		T.nodes.assign((size_t)nInterior * 4, F4{0, 0, 0, 0});
		uint32_t next = 0, maxDepth = 0;
		if (N[root].c[0] < 0)
		{
			T.root = N[root].leafRef;
			return 0;
		}
		T.root = 0;
		// slots in pre-order (node, left subtree, right subtree), like FastBuilder's layout
		std::vector<int32_t> slotOf(N.size(), -1);
		{
			std::vector<int32_t> s2;
			s2.push_back(root);
			next = 0;
			while (!s2.empty())
			{
				int32_t n = s2.back();
				s2.pop_back();
				slotOf[n] = (int32_t)next++;
				if (N[N[n].c[1]].c[0] >= 0) s2.push_back(N[n].c[1]);
				if (N[N[n].c[0]].c[0] >= 0) s2.push_back(N[n].c[0]);
			}
		}
		std::vector<std::pair<int32_t, uint32_t>> s3;
		s3.push_back({root, 0u});
		while (!s3.empty())
		{
			auto [n, d] = s3.back();
			s3.pop_back();
			if (d > maxDepth) maxDepth = d;
			const ONode &a = N[N[n].c[0]], &b = N[N[n].c[1]];
			int32_t r0 = a.c[0] < 0 ? a.leafRef : slotOf[N[n].c[0]], r1 = b.c[0] < 0 ? b.leafRef : slotOf[N[n].c[1]];
			F4* nd = &T.nodes[(size_t)slotOf[n] * 4];
			nd[0] = {a.mn[0], a.mx[0], a.mn[1], a.mx[1]};
			nd[1] = {b.mn[0], b.mx[0], b.mn[1], b.mx[1]};
			nd[2] = {a.mn[2], a.mx[2], b.mn[2], b.mx[2]};
			nd[3] = {bitsToFloat((uint32_t)r0), bitsToFloat((uint32_t)r1), 0.0f, 0.0f};
			if (a.c[0] >= 0) s3.push_back({N[n].c[0], d + 1});
			else if (d + 1 > maxDepth) maxDepth = d + 1;
			if (b.c[0] >= 0) s3.push_back({N[n].c[1], d + 1});
			else if (d + 1 > maxDepth) maxDepth = d + 1;
		}
		T.maxDepth = maxDepth;
		return maxDepth;
	}
};

// ---------------------------------------------------------------------------------------
// Child order of the FAST tree for ANY-HIT rays: a shadow ray that hits both children of a node enters child 0 first and
// stops at the first occluder, so the order decides how much of the tree an OCCLUDED ray walks (measured: +-25 % box
// tests per shadow ray, profiles/r02_tree_opt.txt).  Modes: 1 larger box first, 2 smaller box first; 3 / 4 put first
// the child with the larger p / C — the classic order for a sequential search that stops at the first success — with
//   p = min(1, 2 S / A): chance that a line crossing the child's box (surface area A) meets one of its triangles (summed
//       area S), by Cauchy-Crofton, and
//   C = expected box + triangle tests of walking the child (mode 3: the surface-area recursion
//       C = 2 + sum_k A_k / A * C_k, leaves = their triangle count; mode 4: 2 * log2(leaves) + 2).
// Swapping children does not change any box or leaf: parity is untouched.
// ---------------------------------------------------------------------------------------
inline void orderForAnyHit(FastTree& T, const rtb_tri_isect* tris, int mode)
{
	const size_t nf = T.nodes.size() / 4;
	if (nf == 0 || T.root < 0) return;
	struct Sub
	{
		float S, C; // triangle area, expected cost
		uint32_t leaves;
	};
	auto bitsOf = [](float f) { uint32_t u; memcpy(&u, &f, 4); return u; };
	auto areaOf = [](float x0, float x1, float y0, float y1, float z0, float z1) {
		float dx = x1 - x0, dy = y1 - y0, dz = z1 - z0;
		return 2.0f * (dx * dy + dy * dz + dz * dx);
	};
	std::vector<Sub> sub(nf);
	// children follow their parent in the array (pre-order): a backward sweep sees every child before its parent
	for (size_t i = nf; i-- > 0;)
	{
		F4* nd = &T.nodes[i * 4];
		const int32_t ref[2] = {(int32_t)bitsOf(nd[3].x), (int32_t)bitsOf(nd[3].y)};
		const float A[2] = {areaOf(nd[0].x, nd[0].y, nd[0].z, nd[0].w, nd[2].x, nd[2].y), areaOf(nd[1].x, nd[1].y, nd[1].z, nd[1].w, nd[2].z, nd[2].w)};
		Sub c[2];
		for (int k = 0; k < 2; k++)
		{
			if (ref[k] < 0)
			{
				uint32_t leaf = (uint32_t)(~ref[k]), start = leaf >> 2, count = leaf & 3u;
				c[k].S = 0.0f;
				for (uint32_t t = 0; t < count; t++) c[k].S += tris ? fabsf(tris[start + t].area) : 0.0f;
				c[k].C = (float)count;
				c[k].leaves = 1;
			}
			else c[k] = sub[(size_t)ref[k]];
		}
		float mn[3] = {std::min(nd[0].x, nd[1].x), std::min(nd[0].z, nd[1].z), std::min(nd[2].x, nd[2].z)};
		float mx[3] = {std::max(nd[0].y, nd[1].y), std::max(nd[0].w, nd[1].w), std::max(nd[2].y, nd[2].w)};
		const float An = areaOf(mn[0], mx[0], mn[1], mx[1], mn[2], mx[2]);
		sub[i].S = c[0].S + c[1].S;
		sub[i].leaves = c[0].leaves + c[1].leaves;
		sub[i].C = 2.0f + (An > 0.0f ? (A[0] / An) * c[0].C + (A[1] / An) * c[1].C : c[0].C + c[1].C);
		bool swap = false;
		if (mode == 1) swap = A[1] > A[0];
		else if (mode == 2) swap = A[1] < A[0];
		else
		{
			float key[2];
			for (int k = 0; k < 2; k++)
			{
				float p = A[k] > 0.0f ? std::min(1.0f, 2.0f * c[k].S / A[k]) : 1.0f;
				float cost = mode == 3 ? c[k].C : 2.0f * log2f((float)c[k].leaves) + 2.0f;
				key[k] = p / cost;
			}
			swap = key[1] > key[0];
		}
		if (swap)
		{
			std::swap(nd[0], nd[1]);
			std::swap(nd[2].x, nd[2].z), std::swap(nd[2].y, nd[2].w);
			std::swap(nd[3].x, nd[3].y);
		}
	}
}

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
