// rtb_gpu_build.cuh — the FAST tree built ON THE DEVICE for large scenes ("a fast GPU/parallel BVH builder
// producing a different tree over the same triangle order").  A linear BVH (Lauterbach et al. 2009 / Karras 2012) over
// the reference's LEAVES: 63-bit Morton codes of the exact leaf boxes' centres, one radix sort, every internal node
// found independently from the sorted codes, boxes propagated bottom-up — four kernels and a sort instead of the host's
// recursive binned-SAH build (4 s for the 8.4 M leaves of the 16 M-triangle soup inside rtb_upload_scene).
//
// Legal for hit-ID parity for the same reason the host tree is: the primitives are the reference's leaves with
// their exact boxes, and every child box written here is the exact union (float min / max: no rounding) of the leaf boxes
// below it.  The tree is worse than the SAH tree (median splits of the Morton order: +15 % box tests per ray on the soups,
// +47 % on bathroom, profiles/r02_gpu_builder.txt), and the metric is the render rate, so the host builder stays the
// default; RTB_GPU_BUILD=1 (or RTB_GPU_BUILD_MIN_LEAVES=n) selects this one — time to first image.  Output: the FastTree node layout of
// rtb_dev_scene.cuh (4 x float4 per node: both child boxes + child references), root = node 0.
// The sort is cub::DeviceRadixSort (CUDA toolkit) — set-up plumbing, not the hot path.
#pragma once
#include "rtb_accel.hpp"

#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>

namespace rtb_gpu_build
{

struct LeafRec // == rtb_accel::RefLeaf
{
	float bmin[3], bmax[3];
	uint32_t start, count;
};

__device__ __forceinline__ unsigned long long spread21(unsigned long long x) // 21 bits -> every third bit
{
	x &= 0x1FFFFFull;
	x = (x | (x << 32)) & 0x1F00000000FFFFull;
	x = (x | (x << 16)) & 0x1F0000FF0000FFull;
	x = (x | (x << 8)) & 0x100F00F00F00F00Full;
	x = (x | (x << 4)) & 0x10C30C30C30C30C3ull;
	x = (x | (x << 2)) & 0x1249249249249249ull;
	return x;
}

__global__ void __launch_bounds__(256) k_lbvh_keys(const LeafRec* __restrict__ leaves, uint32_t n, float3 lo, float3 scale, unsigned long long* keys,
                                                   uint32_t* vals)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	LeafRec L = leaves[i];
	float cx = (0.5f * (L.bmin[0] + L.bmax[0]) - lo.x) * scale.x, cy = (0.5f * (L.bmin[1] + L.bmax[1]) - lo.y) * scale.y,
	      cz = (0.5f * (L.bmin[2] + L.bmax[2]) - lo.z) * scale.z;
	unsigned long long x = (unsigned long long)fminf(fmaxf(cx, 0.0f), 2097151.0f), y = (unsigned long long)fminf(fmaxf(cy, 0.0f), 2097151.0f),
	                   z = (unsigned long long)fminf(fmaxf(cz, 0.0f), 2097151.0f);
	keys[i] = spread21(x) | (spread21(y) << 1) | (spread21(z) << 2);
	vals[i] = i;
}

// common-prefix length of sorted positions i and j (ties broken by position: Karras 2012, section 4)
__device__ __forceinline__ int lbvhDelta(const unsigned long long* __restrict__ keys, int n, int i, int j)
{
	if (j < 0 || j >= n) return -1;
	unsigned long long a = keys[i], b = keys[j];
	if (a == b) return 64 + __clz(i ^ j);
	return __clzll((long long)(a ^ b));
}

// node i in [0, n-1): its two children.  child reference >= 0: internal node; < 0: ~(sorted leaf position)
__global__ void __launch_bounds__(256) k_lbvh_topology(const unsigned long long* __restrict__ keys, int n, int2* children, int* parent /* [2n-1]: internal, then leaves */)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n - 1) return;
	int d = (lbvhDelta(keys, n, i, i + 1) - lbvhDelta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
	int dmin = lbvhDelta(keys, n, i, i - d);
	int lmax = 2;
	while (lbvhDelta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
	int l = 0;
	for (int t = lmax >> 1; t >= 1; t >>= 1)
		if (lbvhDelta(keys, n, i, i + (l + t) * d) > dmin) l += t;
	int j = i + l * d;
	int dnode = lbvhDelta(keys, n, i, j);
	int s = 0;
	for (int t = (l + 1) >> 1;; t = (t + 1) >> 1)
	{
		if (lbvhDelta(keys, n, i, i + (s + t) * d) > dnode) s += t;
		if (t == 1) break;
	}
	int gamma = i + s * d + min(d, 0);
	int lo = min(i, j), hi = max(i, j);
	int c0 = (lo == gamma) ? ~gamma : gamma;
	int c1 = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
	children[i] = make_int2(c0, c1);
	parent[c0 >= 0 ? c0 : (n - 1) + (~c0)] = i;
	parent[c1 >= 0 ? c1 : (n - 1) + (~c1)] = i;
	if (i == 0) parent[0] = -1;
}

struct BoxF
{
	float mn[3], mx[3];
};

// bottom-up: the second thread to arrive at a node has both children's boxes; writes the FastTree node and climbs on
__global__ void __launch_bounds__(256) k_lbvh_boxes(const LeafRec* __restrict__ leaves, const uint32_t* __restrict__ order, int n, const int2* __restrict__ children,
                                                    const int* __restrict__ parent, unsigned int* flags, BoxF* boxes /* [n-1] */, unsigned int* height /* [n-1] */,
                                                    float4* fnodes, unsigned int* maxHeight)
{
	int leafPos = blockIdx.x * blockDim.x + threadIdx.x;
	if (leafPos >= n) return;
	int node = parent[(n - 1) + leafPos];
	while (node >= 0)
	{
		if (atomicAdd(&flags[node], 1u) == 0u) return; // first arrival: the sibling's subtree is not finished
		__threadfence();
		int2 ch = children[node];
		BoxF b[2];
		int32_t ref[2];
		unsigned int h = 0;
#pragma unroll
		for (int c = 0; c < 2; c++)
		{
			int k = c ? ch.y : ch.x;
			if (k < 0)
			{
				LeafRec L = leaves[order[~k]];
				for (int a = 0; a < 3; a++) b[c].mn[a] = L.bmin[a], b[c].mx[a] = L.bmax[a];
				ref[c] = ~(int32_t)((L.start << 2) | L.count);
			}
			else
			{
				const volatile float* vb = (const volatile float*)&boxes[k];
				for (int a = 0; a < 3; a++) b[c].mn[a] = vb[a], b[c].mx[a] = vb[3 + a];
				ref[c] = k;
				unsigned int hk = ((volatile unsigned int*)height)[k];
				h = hk > h ? hk : h;
			}
		}
		BoxF u;
		for (int a = 0; a < 3; a++) u.mn[a] = fminf(b[0].mn[a], b[1].mn[a]), u.mx[a] = fmaxf(b[0].mx[a], b[1].mx[a]);
		boxes[node] = u;
		height[node] = h + 1u;
		float4* nd = fnodes + (size_t)node * 4;
		nd[0] = make_float4(b[0].mn[0], b[0].mx[0], b[0].mn[1], b[0].mx[1]);
		nd[1] = make_float4(b[1].mn[0], b[1].mx[0], b[1].mn[1], b[1].mx[1]);
		nd[2] = make_float4(b[0].mn[2], b[0].mx[2], b[1].mn[2], b[1].mx[2]);
		nd[3] = make_float4(__int_as_float(ref[0]), __int_as_float(ref[1]), 0.0f, 0.0f);
		__threadfence();
		if (node == 0) *maxHeight = h + 1u;
		node = parent[node];
	}
}

// Builds the tree into `fnodes` (device, (n-1) x 4 float4, allocated by the caller).  Returns cudaSuccess and the tree's
// depth (levels below the root).  n >= 2.
inline cudaError_t build(const rtb_accel::RefLeaf* hostLeaves, uint32_t n, const float sceneMin[3], const float sceneMax[3], float4* fnodes, uint32_t* depthOut,
                         cudaStream_t stream)
{
	static_assert(sizeof(LeafRec) == sizeof(rtb_accel::RefLeaf), "leaf record layout");
	LeafRec* dLeaves = nullptr;
	unsigned long long *keysA = nullptr, *keysB = nullptr;
	uint32_t *valsA = nullptr, *valsB = nullptr;
	int2* children = nullptr;
	int* parent = nullptr;
	unsigned int *flags = nullptr, *height = nullptr, *maxH = nullptr;
	BoxF* boxes = nullptr;
	void* tmp = nullptr;
	size_t tmpBytes = 0;
	cudaError_t e = cudaSuccess;
#define RTB_GB(call)                 \
	do                               \
	{                                \
		e = (call);                  \
		if (e != cudaSuccess) goto done; \
	} while (0)
	RTB_GB(cudaMalloc((void**)&dLeaves, (size_t)n * sizeof(LeafRec)));
	RTB_GB(cudaMalloc((void**)&keysA, (size_t)n * 8));
	RTB_GB(cudaMalloc((void**)&keysB, (size_t)n * 8));
	RTB_GB(cudaMalloc((void**)&valsA, (size_t)n * 4));
	RTB_GB(cudaMalloc((void**)&valsB, (size_t)n * 4));
	RTB_GB(cudaMalloc((void**)&children, (size_t)(n - 1) * sizeof(int2)));
	RTB_GB(cudaMalloc((void**)&parent, (size_t)(2 * (size_t)n - 1) * sizeof(int)));
	RTB_GB(cudaMalloc((void**)&flags, (size_t)(n - 1) * 4));
	RTB_GB(cudaMalloc((void**)&height, (size_t)(n - 1) * 4));
	RTB_GB(cudaMalloc((void**)&boxes, (size_t)(n - 1) * sizeof(BoxF)));
	RTB_GB(cudaMalloc((void**)&maxH, 4));
	RTB_GB(cudaMemcpyAsync(dLeaves, hostLeaves, (size_t)n * sizeof(LeafRec), cudaMemcpyHostToDevice, stream));
	RTB_GB(cudaMemsetAsync(flags, 0, (size_t)(n - 1) * 4, stream));
	RTB_GB(cudaMemsetAsync(maxH, 0, 4, stream));
	{
		float3 lo = make_float3(sceneMin[0], sceneMin[1], sceneMin[2]);
		float ex = sceneMax[0] - sceneMin[0], ey = sceneMax[1] - sceneMin[1], ez = sceneMax[2] - sceneMin[2];
		float3 sc = make_float3(ex > 0 ? 2097151.0f / ex : 0.0f, ey > 0 ? 2097151.0f / ey : 0.0f, ez > 0 ? 2097151.0f / ez : 0.0f);
		unsigned grid = (n + 255u) / 256u;
		k_lbvh_keys<<<grid, 256, 0, stream>>>(dLeaves, n, lo, sc, keysA, valsA);
		RTB_GB(cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, keysA, keysB, valsA, valsB, (int)n, 0, 63, stream));
		RTB_GB(cudaMalloc(&tmp, tmpBytes));
		RTB_GB(cub::DeviceRadixSort::SortPairs(tmp, tmpBytes, keysA, keysB, valsA, valsB, (int)n, 0, 63, stream));
		k_lbvh_topology<<<(n - 1 + 255u) / 256u, 256, 0, stream>>>(keysB, (int)n, children, parent);
		k_lbvh_boxes<<<grid, 256, 0, stream>>>(dLeaves, valsB, (int)n, children, parent, flags, boxes, height, fnodes, maxH);
		RTB_GB(cudaGetLastError());
		unsigned int h = 0;
		RTB_GB(cudaMemcpyAsync(&h, maxH, 4, cudaMemcpyDeviceToHost, stream));
		RTB_GB(cudaStreamSynchronize(stream));
		*depthOut = h; // root at level 0 has height h: the deepest leaf reference sits h levels below
	}
done:
#undef RTB_GB
	cudaFree(dLeaves), cudaFree(keysA), cudaFree(keysB), cudaFree(valsA), cudaFree(valsB), cudaFree(children), cudaFree(parent), cudaFree(flags);
	cudaFree(height), cudaFree(boxes), cudaFree(maxH), cudaFree(tmp);
	return e;
}

} // namespace rtb_gpu_build
