// rtb_dev_math.cuh — float3 arithmetic with the REFERENCE's operation order.
//
// Everything that feeds a hit decision (ray generation, slab test, triangle test) has to
// round exactly like the g++ -O2 -ffp-contract=off build of RTBase/Core.h, so this file is
// compiled with -fmad=false (no FMA contraction) and without --use_fast_math, and every
// helper spells out the association the reference uses:
//   Dot   = ((x*x') + (y*y')) + (z*z')                      RTBase/Core.h:176-179
//   Cross = (y*z' - z*y', z*x' - x*z', x*y' - y*x')         RTBase/Core.h:170-173
//   normalize: l = 1.0f / sqrtf(x*x + y*y + z*z); v * l     RTBase/Core.h:161-165
//   Min/Max(Vec3): a<b ? a : b  /  a>b ? a : b (NaN -> b)   RTBase/Core.h:187-195
//   std::max(a,b) = (a<b) ? b : a ; std::min(a,b) = (b<a) ? b : a
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#define RTB_DEV __device__ __forceinline__

struct V3
{
	float x, y, z;
};

RTB_DEV V3 mk(float x, float y, float z)
{
	V3 v;
	v.x = x, v.y = y, v.z = z;
	return v;
}
RTB_DEV V3 mk(const float* p) { return mk(p[0], p[1], p[2]); }
RTB_DEV V3 mk(float4 f) { return mk(f.x, f.y, f.z); }
RTB_DEV V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
RTB_DEV V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
RTB_DEV V3 operator*(V3 a, V3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
RTB_DEV V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
RTB_DEV V3 operator/(V3 a, float s) { return mk(a.x / s, a.y / s, a.z / s); }
// shading only (1e-5 gate): one correctly rounded reciprocal instead of three divisions
RTB_DEV V3 divFast(V3 a, float s)
{
	float r = __frcp_rn(s);
	return mk(a.x * r, a.y * r, a.z * r);
}
RTB_DEV V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
RTB_DEV float dot(V3 a, V3 b) { return ((a.x * b.x) + (a.y * b.y)) + (a.z * b.z); }
RTB_DEV V3 cross(V3 a, V3 b)
{
	return mk((a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x));
}
RTB_DEV float lengthSq(V3 a) { return ((a.x * a.x) + (a.y * a.y)) + (a.z * a.z); }
RTB_DEV V3 normalize(V3 a)
{
	float l = 1.0f / sqrtf(((a.x * a.x) + (a.y * a.y)) + (a.z * a.z));
	return mk(a.x * l, a.y * l, a.z * l);
}
// The selects of RTBase/Core.h:187-195 and libstdc++'s std::min/std::max; NOT fminf/fmaxf,
// which differ when an operand is NaN (0 * inf in the slab test).
RTB_DEV float selMin(float a, float b) { return a < b ? a : b; }   // Min(Vec3,Vec3), Windows min()
RTB_DEV float selMax(float a, float b) { return a > b ? a : b; }   // Max(Vec3,Vec3), Windows max()
RTB_DEV float stdMax(float a, float b) { return (a < b) ? b : a; } // std::max
RTB_DEV float stdMin(float a, float b) { return (b < a) ? b : a; } // std::min
RTB_DEV float lum(V3 c) { return ((0.2126f * c.x) + (0.7152f * c.y)) + (0.0722f * c.z); } // Core.h:89-92

struct RayD
{
	V3 o, d, inv;
};
RTB_DEV RayD mkRay(V3 o, V3 d) // Ray::init, RTBase/Geometry.h:21-26
{
	RayD r;
	r.o = o;
	r.d = d;
	r.inv = mk(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
	return r;
}
