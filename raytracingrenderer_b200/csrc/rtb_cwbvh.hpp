// rtb_cwbvh.hpp — host-side builder of the CW tree (RTB_TRAV_CW): the FAST binary tree collapsed to EIGHT
// children per node, child boxes quantised to 8 bits per plane on a per-node power-of-two grid (the compressed
// wide BVH of Ylitie, Karras & Laine, HPG 2017, adapted to this library's exactness contract).
//
// Why it is legal: the reference's answer is the (t, ID) minimum over the triangles of every LEAF
// whose exact box passes the reference's slab test.  Interior structure is free as long as no such leaf is
// culled, so interior (and leaf-slot) boxes only have to be CONSERVATIVE: every quantised box contains the
// exact box it stands for, padded by one full grid step per side, and the device test adds a relative slack
// (rtb_dev_cw.cuh) — together they dominate every rounding difference between the quantised FMA test and the
// reference's (min - o) * invDir arithmetic.  The exact box of a reference leaf is kept (32 B) and tested with the
// reference's own arithmetic before the leaf's triangles.
//
// Node = 80 bytes = 5 x float4, breadth-first so that the top of the tree is a prefix of the array (it is staged
// into shared memory by the traversal kernels) and the internal children of a node are consecutive:
//   [0] p.x p.y p.z | bits(ex | ey << 8 | ez << 16 | imask << 24)      grid origin, biased exponents of the grid
//                                                                       steps, which slots hold internal nodes
//   [1] bits(childBase) bits(leafBase) bits(lmask) -                   first internal child / first leaf record,
//                                                                       which slots hold reference leaves
//   [2] qlo.x[0..3] qlo.x[4..7] qlo.y[0..3] qlo.y[4..7]                one byte per slot
//   [3] qlo.z[0..3] qlo.z[4..7] qhi.x[0..3] qhi.x[4..7]
//   [4] qhi.y[0..3] qhi.y[4..7] qhi.z[0..3] qhi.z[4..7]
// plane = p + q * 2^(e - 127).  Child k of a node (k-th set bit of imask, ascending slot) is node childBase + k;
// leaf k (k-th set bit of lmask) is leaf record leafBase + k.
// Slot assignment: the child that lies farthest along the diagonal direction D_s = (s&1 ? + : -, s&2 ? + : -,
// s&4 ? + : -) gets slot s, so a ray whose direction signs are D_q visits slots in DESCENDING (s xor q): near
// to far without sorting.
// Leaf record = 2 x float4: exact min.xyz, bits(start << 2 | count) | exact max.xyz, 0.
#pragma once
#include "rtb_accel.hpp"

#include <deque>

namespace rtb_accel
{

struct CwTree
{
	std::vector<F4> nodes;  // 5 x F4 per node
	std::vector<F4> leaves; // 2 x F4 per leaf record
	uint32_t maxDepth = 0;
	bool valid = false; // false: the scene is a single leaf / empty (the traversal falls back to the reference tree)
};

class CwBuilder
{
public:
	explicit CwBuilder(const FastTree& t) : B(t) {}

	void build(CwTree& out)
	{
		out.nodes.clear();
		out.leaves.clear();
		out.maxDepth = 0;
		out.valid = false;
		if (B.root < 0 || B.nodes.empty()) return;
		struct Item
		{
			int32_t fastNode;
			uint32_t cwIndex, depth;
		};
		std::deque<Item> queue;
		out.nodes.resize(5);
		queue.push_back({B.root, 0u, 0u});
		while (!queue.empty())
		{
			Item it = queue.front();
			queue.pop_front();
			if (it.depth > out.maxDepth) out.maxDepth = it.depth;
			Cand c[8];
			int n = collapse(it.fastNode, c);
			int slotOf[8], childAt[8];
			assignSlots(c, n, slotOf);
			for (int s = 0; s < 8; s++) childAt[s] = -1;
			for (int i = 0; i < n; i++) childAt[slotOf[i]] = i;
			uint32_t imask = 0, lmask = 0;
			uint32_t childBase = (uint32_t)(out.nodes.size() / 5), leafBase = (uint32_t)(out.leaves.size() / 2);
			for (int s = 0; s < 8; s++)
			{
				if (childAt[s] < 0) continue;
				const Cand& k = c[childAt[s]];
				if (k.ref >= 0)
				{
					imask |= 1u << s;
					uint32_t idx = (uint32_t)(out.nodes.size() / 5);
					out.nodes.resize(out.nodes.size() + 5);
					queue.push_back({k.ref, idx, it.depth + 1});
				}
				else
				{
					lmask |= 1u << s;
					out.leaves.push_back({k.mn[0], k.mn[1], k.mn[2], bitsToFloat((uint32_t)(~k.ref))});
					out.leaves.push_back({k.mx[0], k.mx[1], k.mx[2], 0.0f});
				}
			}
			writeNode(&out.nodes[(size_t)it.cwIndex * 5], c, childAt, imask, lmask, childBase, leafBase);
		}
		out.valid = true;
	}

private:
	const FastTree& B;
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
	// the binary node's two children, the largest internal one replaced by its own children until there are eight
	int collapse(int32_t node, Cand* c) const
	{
		int n = 2;
		children(node, c[0], c[1]);
		while (n < 8)
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
		return n;
	}
	// greedy assignment: repeatedly the (child, slot) pair with the largest projection of the child's centre (relative to
	// the node's) on the slot's diagonal
	static void assignSlots(const Cand* c, int n, int* slotOf)
	{
		float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
		for (int i = 0; i < n; i++)
			for (int k = 0; k < 3; k++)
			{
				if (c[i].mn[k] < mn[k]) mn[k] = c[i].mn[k];
				if (c[i].mx[k] > mx[k]) mx[k] = c[i].mx[k];
			}
		float cost[8][8];
		for (int i = 0; i < n; i++)
			for (int s = 0; s < 8; s++)
			{
				float v = 0.0f;
				for (int k = 0; k < 3; k++)
				{
					float d = 0.5f * (c[i].mn[k] + c[i].mx[k]) - 0.5f * (mn[k] + mx[k]);
					v += ((s >> k) & 1) ? d : -d;
				}
				cost[i][s] = v;
			}
		bool childDone[8] = {}, slotDone[8] = {};
		for (int round = 0; round < n; round++)
		{
			int bi = -1, bs = -1;
			float bv = -FLT_MAX;
			for (int i = 0; i < n; i++)
			{
				if (childDone[i]) continue;
				for (int s = 0; s < 8; s++)
					if (!slotDone[s] && (bi < 0 || cost[i][s] > bv)) bv = cost[i][s], bi = i, bs = s;
			}
			childDone[bi] = true, slotDone[bs] = true;
			slotOf[bi] = bs;
		}
	}
	static float floatDown(double v)
	{
		float f = (float)v;
		if ((double)f > v) f = nextafterf(f, -FLT_MAX);
		return f;
	}
	static void writeNode(F4* nd, const Cand* c, const int* childAt, uint32_t imask, uint32_t lmask, uint32_t childBase, uint32_t leafBase)
	{
		float p[3];
		uint32_t eb[3];
		uint8_t q[6][8];
		memset(q, 0, sizeof(q));
		for (int k = 0; k < 3; k++)
		{
			double lo = DBL_MAX, hi = -DBL_MAX;
			for (int s = 0; s < 8; s++)
				if (childAt[s] >= 0)
				{
					lo = std::min(lo, (double)c[childAt[s]].mn[k]);
					hi = std::max(hi, (double)c[childAt[s]].mx[k]);
				}
			// grid step 2^ex: the extent in at most 252 steps, and never finer than 2^-20 of the coordinates' magnitude (so that
			// the one-step padding means something for flat nodes and p - step is a float below lo)
			double mag = std::max(std::fabs(lo), std::fabs(hi));
			int ex = -126;
			if (hi > lo) ex = std::max(ex, (int)std::ceil(std::log2((hi - lo) / 252.0)));
			if (mag > 0.0) ex = std::max(ex, (int)std::ceil(std::log2(mag)) - 20);
			if (ex > 127) ex = 127;
			for (;; ex++)
			{
				double e = std::ldexp(1.0, ex);
				float pf = floatDown(lo - e);
				bool ok = true;
				for (int s = 0; s < 8 && ok; s++)
					if (childAt[s] >= 0)
					{
						double ql = std::floor(((double)c[childAt[s]].mn[k] - (double)pf) / e) - 1.0;
						double qh = std::ceil(((double)c[childAt[s]].mx[k] - (double)pf) / e) + 1.0;
						if (ql < 0.0 || qh > 255.0) ok = false;
					}
				if (ok || ex >= 127)
				{
					p[k] = pf;
					eb[k] = (uint32_t)(ex + 127);
					for (int s = 0; s < 8; s++)
						if (childAt[s] >= 0)
						{
							double ql = std::floor(((double)c[childAt[s]].mn[k] - (double)pf) / e) - 1.0;
							double qh = std::ceil(((double)c[childAt[s]].mx[k] - (double)pf) / e) + 1.0;
							q[k][s] = (uint8_t)std::min(std::max(ql, 0.0), 255.0);
							q[3 + k][s] = (uint8_t)std::min(std::max(qh, 0.0), 255.0);
						}
					break;
				}
			}
		}
		auto word = [&](int plane, int half) {
			const uint8_t* b = &q[plane][half * 4];
			return bitsToFloat((uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24));
		};
		nd[0] = {p[0], p[1], p[2], bitsToFloat(eb[0] | (eb[1] << 8) | (eb[2] << 16) | (imask << 24))};
		nd[1] = {bitsToFloat(childBase), bitsToFloat(leafBase), bitsToFloat(lmask), 0.0f};
		nd[2] = {word(0, 0), word(0, 1), word(1, 0), word(1, 1)};
		nd[3] = {word(2, 0), word(2, 1), word(3, 0), word(3, 1)};
		nd[4] = {word(4, 0), word(4, 1), word(5, 0), word(5, 1)};
	}
};

} // namespace rtb_accel
