// rtb_scene.hpp — stand-alone host scene API with the reference's class names and field meaning
// (RTBase/Core.h, Geometry.h, Imaging.h, Materials.h, Lights.h, Scene.h, SceneLoader.h,
// GEMLoader.h), so that code written against RTBase's `Scene` / `loadScene` compiles against
// this header and host/Renderer.h-style code can flatten it with host/rtb_flatten.hpp.
//
// It is a DATA MODEL + LOADER + BVH BUILDER only.  The hot-path methods of the reference
// (Scene::traverse/visible, BSDF::sample/evaluate, Light::sample ...) are deliberately absent:
// that work runs on the GPU behind include/rtb.h and there is no CPU fallback.
//
// Bit-compatibility (tests/test_host_cpu.py compares the flattened output of this loader with the
// flattened output of the reference's own loadScene, byte for byte):
//   * every matrix / vertex transform spells out the reference's operation order;
//   * the BVH builder makes the reference's decisions (RTBase/Geometry.h:325-392): leaves of
//     <= 2 triangles, longest axis with the same tie rule, std::sort by centroid (the same
//     libstdc++ algorithm sees the same comparison results, so it produces the same permutation
//     although only 8-byte keys move), full SAH sweep with a strict '<';
//   * PNG (all colour types and bit depths, tRNS, Adam7), JPEG (host/rtb_jpeg.hpp: baseline +
//     progressive) and Radiance .hdr decode to the values stb_image produces.  A file that exists but cannot be decoded is
//     an error, never a silent default.
// Build with -ffp-contract=off.
#pragma once

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#include <zlib.h>

#include "rtb_jpeg.hpp"

#ifndef EPSILON
#define EPSILON 1e-4f /* RTBase/Geometry.h:60 */
#endif
#define MAXNODE_TRIANGLES 2 /* RTBase/Geometry.h:240 */

// ------------------------------------------------------------------------------------------
// Core.h
// ------------------------------------------------------------------------------------------
class Colour
{
public:
	float r = 0, g = 0, b = 0;
	Colour() {}
	Colour(float _r, float _g, float _b) : r(_r), g(_g), b(_b) {}
	float Lum() const { return ((0.2126f * r) + (0.7152f * g) + (0.0722f * b)); } // Core.h:89-92
};

class Vec3
{
public:
	union
	{
		struct
		{
			float x, y, z, w;
		};
		float coords[4];
	};
	Vec3() : x(0), y(0), z(0), w(1.0f) {}
	Vec3(float _x, float _y, float _z) : x(_x), y(_y), z(_z), w(1.0f) {}
	Vec3 operator+(const Vec3 v) const { return Vec3(x + v.x, y + v.y, z + v.z); }
	Vec3 operator-(const Vec3 v) const { return Vec3(x - v.x, y - v.y, z - v.z); }
	Vec3 operator*(const float v) const { return Vec3(x * v, y * v, z * v); }
	Vec3 operator/(const float v) const { return Vec3(x / v, y / v, z / v); }
	Vec3 operator-() const { return Vec3(-x, -y, -z); }
	float length() const { return sqrtf((x * x) + (y * y) + (z * z)); }
	Vec3 normalize() const // Core.h:161-165: one reciprocal, three products
	{
		float l = 1.0f / sqrtf((x * x) + (y * y) + (z * z));
		return Vec3(x * l, y * l, z * l);
	}
	float dot(Vec3 v) const { return ((x * v.x) + (y * v.y) + (z * v.z)); }
	Vec3 cross(Vec3 v) const { return Vec3((y * v.z) - (z * v.y), (z * v.x) - (x * v.z), (x * v.y) - (y * v.x)); }
};
static inline float Dot(const Vec3 a, const Vec3 b) { return ((a.x * b.x) + (a.y * b.y) + (a.z * b.z)); }
static inline Vec3 Max(Vec3 a, Vec3 b) { return Vec3(a.x > b.x ? a.x : b.x, a.y > b.y ? a.y : b.y, a.z > b.z ? a.z : b.z); }
static inline Vec3 Min(Vec3 a, Vec3 b) { return Vec3(a.x < b.x ? a.x : b.x, a.y < b.y ? a.y : b.y, a.z < b.z ? a.z : b.z); }

struct Vertex
{
	Vec3 p;
	Vec3 normal;
	float u;
	float v;
};

class Matrix
{
public:
	union
	{
		float a[4][4];
		float m[16];
	};
	Matrix() { identity(); }
	void identity()
	{
		memset(m, 0, sizeof(m));
		m[0] = m[5] = m[10] = m[15] = 1.0f;
	}
	Matrix transpose() const
	{
		Matrix t;
		for (int i = 0; i < 4; i++)
			for (int j = 0; j < 4; j++) t.a[i][j] = a[j][i];
		return t;
	}
	Vec3 mulVec(const Vec3& v) const // Core.h:295-301
	{
		return Vec3((v.x * m[0] + v.y * m[1] + v.z * m[2]), (v.x * m[4] + v.y * m[5] + v.z * m[6]),
		            (v.x * m[8] + v.y * m[9] + v.z * m[10]));
	}
	Vec3 mulPoint(const Vec3& v) const // Core.h:302-309
	{
		return Vec3((v.x * m[0] + v.y * m[1] + v.z * m[2]) + m[3], (v.x * m[4] + v.y * m[5] + v.z * m[6]) + m[7],
		            (v.x * m[8] + v.y * m[9] + v.z * m[10]) + m[11]);
	}
	Vec3 mulPointAndPerspectiveDivide(const Vec3& v) const // Core.h:310-320
	{
		Vec3 v1 = mulPoint(v);
		float w = (m[12] * v.x) + (m[13] * v.y) + (m[14] * v.z) + m[15];
		w = 1.0f / w;
		return (v1 * w);
	}
	// Cofactor inverse with the term order of the classic MESA/GLU routine the reference uses
	// (Core.h:326-438): out[i] = sum of six signed triple products, accumulated left to right.
	Matrix invert() const
	{
		// rows: output index, then six (sign, a, b, c) terms
		static const signed char T[16][1 + 6 * 4] = {
			{0, +1, 5, 10, 15, -1, 5, 11, 14, -1, 9, 6, 15, +1, 9, 7, 14, +1, 13, 6, 11, -1, 13, 7, 10},
			{4, -1, 4, 10, 15, +1, 4, 11, 14, +1, 8, 6, 15, -1, 8, 7, 14, -1, 12, 6, 11, +1, 12, 7, 10},
			{8, +1, 4, 9, 15, -1, 4, 11, 13, -1, 8, 5, 15, +1, 8, 7, 13, +1, 12, 5, 11, -1, 12, 7, 9},
			{12, -1, 4, 9, 14, +1, 4, 10, 13, +1, 8, 5, 14, -1, 8, 6, 13, -1, 12, 5, 10, +1, 12, 6, 9},
			{1, -1, 1, 10, 15, +1, 1, 11, 14, +1, 9, 2, 15, -1, 9, 3, 14, -1, 13, 2, 11, +1, 13, 3, 10},
			{5, +1, 0, 10, 15, -1, 0, 11, 14, -1, 8, 2, 15, +1, 8, 3, 14, +1, 12, 2, 11, -1, 12, 3, 10},
			{9, -1, 0, 9, 15, +1, 0, 11, 13, +1, 8, 1, 15, -1, 8, 3, 13, -1, 12, 1, 11, +1, 12, 3, 9},
			{13, +1, 0, 9, 14, -1, 0, 10, 13, -1, 8, 1, 14, +1, 8, 2, 13, +1, 12, 1, 10, -1, 12, 2, 9},
			{2, +1, 1, 6, 15, -1, 1, 7, 14, -1, 5, 2, 15, +1, 5, 3, 14, +1, 13, 2, 7, -1, 13, 3, 6},
			{6, -1, 0, 6, 15, +1, 0, 7, 14, +1, 4, 2, 15, -1, 4, 3, 14, -1, 12, 2, 7, +1, 12, 3, 6},
			{10, +1, 0, 5, 15, -1, 0, 7, 13, -1, 4, 1, 15, +1, 4, 3, 13, +1, 12, 1, 7, -1, 12, 3, 5},
			{14, -1, 0, 5, 14, +1, 0, 6, 13, +1, 4, 1, 14, -1, 4, 2, 13, -1, 12, 1, 6, +1, 12, 2, 5},
			{3, -1, 1, 6, 11, +1, 1, 7, 10, +1, 5, 2, 11, -1, 5, 3, 10, -1, 9, 2, 7, +1, 9, 3, 6},
			{7, +1, 0, 6, 11, -1, 0, 7, 10, -1, 4, 2, 11, +1, 4, 3, 10, +1, 8, 2, 7, -1, 8, 3, 6},
			{11, -1, 0, 5, 11, +1, 0, 7, 9, +1, 4, 1, 11, -1, 4, 3, 9, -1, 8, 1, 7, +1, 8, 3, 5},
			{15, +1, 0, 5, 10, -1, 0, 6, 9, -1, 4, 1, 10, +1, 4, 2, 9, +1, 8, 1, 6, -1, 8, 2, 5},
		};
		Matrix inv;
		for (int r = 0; r < 16; r++)
		{
			float acc = 0.0f;
			for (int k = 0; k < 6; k++)
			{
				const signed char* t = &T[r][1 + k * 4];
				float first = t[0] < 0 ? -m[t[1]] : m[t[1]];
				float term = first * m[t[2]] * m[t[3]];
				// "-x*y*z + ..." starts from the negated product; later terms are added or subtracted
				acc = (k == 0) ? term : acc + term;
			}
			inv.m[T[r][0]] = acc;
		}
		float det = m[0] * inv.m[0] + m[1] * inv.m[4] + m[2] * inv.m[8] + m[3] * inv.m[12];
		if (det == 0)
		{
			inv.identity(); // unreachable for the matrices a scene file can produce; keep a sane value
			return inv;
		}
		det = 1.0f / det;
		for (int i = 0; i < 16; i++) inv.m[i] = inv.m[i] * det;
		return inv;
	}
	static Matrix lookAt(const Vec3& from, const Vec3& to, const Vec3& up) // Core.h:439-459
	{
		Matrix mat;
		Vec3 dir = (from - to).normalize();
		Vec3 left = up.cross(dir).normalize();
		Vec3 newUp = dir.cross(left);
		mat.a[0][0] = left.x, mat.a[0][1] = left.y, mat.a[0][2] = left.z;
		mat.a[1][0] = newUp.x, mat.a[1][1] = newUp.y, mat.a[1][2] = newUp.z;
		mat.a[2][0] = dir.x, mat.a[2][1] = dir.y, mat.a[2][2] = dir.z;
		mat.a[0][3] = -from.dot(left);
		mat.a[1][3] = -from.dot(newUp);
		mat.a[2][3] = -from.dot(dir);
		mat.a[3][3] = 1;
		return mat;
	}
	static Matrix perspective(const float n, const float f, float aspect, const float fov) // Core.h:460-471
	{
		Matrix pers;
		memset(pers.m, 0, sizeof(pers.m));
		float t = 1.0f / (tanf(fov * 0.5f * 3.141592654f / 180.0f));
		pers.a[0][0] = t / aspect;
		pers.a[1][1] = t;
		pers.a[2][2] = -f / (f - n);
		pers.a[2][3] = -(f * n) / (f - n);
		pers.a[3][2] = -1.0f;
		return pers;
	}
};

// ------------------------------------------------------------------------------------------
// Imaging.h: Texture (decoded here without stb: PNG via zlib, Radiance RGBE)
// ------------------------------------------------------------------------------------------
namespace rtb_img
{
inline std::vector<unsigned char> readFile(const std::string& path)
{
	std::ifstream f(path, std::ios::binary);
	if (!f) return {};
	return std::vector<unsigned char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
inline uint32_t be32(const unsigned char* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

// 8-bit, non-interlaced PNG of colour type 0 (grey), 2 (RGB), 4 (grey+alpha), 6 (RGBA).
// PNG -> 8-bit interleaved pixels with the channel count stb_image reports for req_comp = 0: every colour type
// (grey, RGB, palette, grey+alpha, RGBA), bit depths 1/2/4/8/16, tRNS, Adam7 interlacing.  The conventions
// that PNG leaves to the reader are stb_image's, because the reference decodes with it: sub-byte grey is
// scaled to 0..255 (x255, x85, x17), 16-bit samples keep their high byte, a tRNS colour adds an alpha channel
// of 0 / 255, a palette expands to RGB (RGBA with tRNS).
inline bool pngUnfilter(const unsigned char* src, size_t srcLen, size_t& used, int x, int y, int pixelBits, std::vector<unsigned char>& rows)
{
	size_t rowBytes = ((size_t)x * pixelBits + 7) >> 3;
	int bpp = pixelBits >= 8 ? pixelBits / 8 : 1; // filter unit in bytes
	if (used + (rowBytes + 1) * (size_t)y > srcLen) return false;
	rows.assign(rowBytes * (size_t)y, 0);
	for (int j = 0; j < y; j++)
	{
		const unsigned char* in = src + used;
		unsigned char* dst = &rows[rowBytes * (size_t)j];
		const unsigned char* up = j ? dst - rowBytes : nullptr;
		int filter = in[0];
		in++;
		used += rowBytes + 1;
		for (size_t i = 0; i < rowBytes; i++)
		{
			int a = i >= (size_t)bpp ? dst[i - bpp] : 0;
			int bb = up ? up[i] : 0;
			int c = (up && i >= (size_t)bpp) ? up[i - bpp] : 0;
			int v = in[i];
			switch (filter)
			{
			case 0: break;
			case 1: v += a; break;
			case 2: v += bb; break;
			case 3: v += (a + bb) >> 1; break;
			case 4:
			{
				int pp = a + bb - c, pa = abs(pp - a), pb = abs(pp - bb), pc = abs(pp - c);
				v += (pa <= pb && pa <= pc) ? a : (pb <= pc ? bb : c);
				break;
			}
			default: return false;
			}
			dst[i] = (unsigned char)v;
		}
	}
	return true;
}

inline bool decodePNG(const std::vector<unsigned char>& file, int& w, int& h, int& channels, std::vector<unsigned char>& out)
{
	static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
	if (file.size() < 8 || memcmp(file.data(), sig, 8) != 0) return false;
	size_t pos = 8;
	std::vector<unsigned char> idat;
	int depth = 0, colourType = 0, interlace = 0;
	unsigned char palette[256][4];
	int palLen = 0;
	bool hasTrans = false, palTrans = false;
	unsigned tc[3] = {0, 0, 0};
	w = h = 0;
	while (pos + 12 <= file.size())
	{
		uint32_t len = be32(&file[pos]);
		const unsigned char* type = &file[pos + 4];
		const unsigned char* data = &file[pos + 8];
		if (pos + 12 + (size_t)len > file.size()) return false;
		if (!memcmp(type, "IHDR", 4))
		{
			if (len != 13) return false;
			w = (int)be32(data), h = (int)be32(data + 4);
			depth = data[8], colourType = data[9], interlace = data[12];
			if (data[10] != 0 || data[11] != 0 || interlace > 1) return false;
		}
		else if (!memcmp(type, "PLTE", 4))
		{
			if (len > 256 * 3 || len % 3) return false;
			palLen = (int)(len / 3);
			for (int i = 0; i < palLen; i++)
				palette[i][0] = data[i * 3], palette[i][1] = data[i * 3 + 1], palette[i][2] = data[i * 3 + 2], palette[i][3] = 255;
		}
		else if (!memcmp(type, "tRNS", 4))
		{
			if (colourType == 3)
			{
				if (palLen == 0 || (int)len > palLen) return false;
				for (uint32_t i = 0; i < len; i++) palette[i][3] = data[i];
				palTrans = true;
			}
			else
			{
				int n = (colourType & 2) ? 3 : 1;
				if ((colourType & 4) || (int)len != n * 2) return false;
				for (int k = 0; k < n; k++) tc[k] = ((unsigned)data[k * 2] << 8) | data[k * 2 + 1];
				hasTrans = true;
			}
		}
		else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
		else if (!memcmp(type, "IEND", 4)) break;
		pos += 12 + (size_t)len;
	}
	if (w <= 0 || h <= 0) return false;
	if (depth != 1 && depth != 2 && depth != 4 && depth != 8 && depth != 16) return false;
	int imgN;
	switch (colourType)
	{
	case 0: imgN = 1; break;
	case 2: imgN = 3; break;
	case 3: imgN = 1; break;
	case 4: imgN = 2; break;
	case 6: imgN = 4; break;
	default: return false;
	}
	if (colourType == 3 && (depth == 16 || palLen == 0)) return false;
	if ((colourType == 2 || colourType == 4 || colourType == 6) && depth < 8) return false;
	// inflate
	const int pixelBits = imgN * depth;
	size_t rawCap = 0;
	static const int xo[7] = {0, 4, 0, 2, 0, 1, 0}, yo[7] = {0, 0, 4, 0, 2, 0, 1}, xs[7] = {8, 8, 4, 4, 2, 2, 1}, ys[7] = {8, 8, 8, 4, 4, 2, 2};
	if (!interlace) rawCap = ((((size_t)w * pixelBits + 7) >> 3) + 1) * (size_t)h;
	else
		for (int p = 0; p < 7; p++)
		{
			int x = (w - xo[p] + xs[p] - 1) / xs[p], y = (h - yo[p] + ys[p] - 1) / ys[p];
			if (x > 0 && y > 0) rawCap += ((((size_t)x * pixelBits + 7) >> 3) + 1) * (size_t)y;
		}
	std::vector<unsigned char> raw(rawCap);
	uLongf rawLen = (uLongf)raw.size();
	if (idat.empty() || uncompress(raw.data(), &rawLen, idat.data(), (uLong)idat.size()) != Z_OK || rawLen != raw.size()) return false;
	// samples as 16-bit values (so that one code path serves all depths), imgN per pixel
	std::vector<uint16_t> px((size_t)w * h * imgN);
	static const int depthScale[9] = {0, 0xff, 0x55, 0, 0x11, 0, 0, 0, 0x01};
	const int scale = (colourType == 0 && depth < 8) ? depthScale[depth] : 1;
	size_t used = 0;
	std::vector<unsigned char> rows;
	for (int p = 0; p < (interlace ? 7 : 1); p++)
	{
		int x = interlace ? (w - xo[p] + xs[p] - 1) / xs[p] : w, y = interlace ? (h - yo[p] + ys[p] - 1) / ys[p] : h;
		if (x <= 0 || y <= 0) continue;
		if (!pngUnfilter(raw.data(), raw.size(), used, x, y, pixelBits, rows)) return false;
		size_t rowBytes = ((size_t)x * pixelBits + 7) >> 3;
		for (int j = 0; j < y; j++)
			for (int i = 0; i < x; i++)
			{
				int ox = interlace ? i * xs[p] + xo[p] : i, oy = interlace ? j * ys[p] + yo[p] : j;
				uint16_t* d = &px[((size_t)oy * w + ox) * imgN];
				const unsigned char* r = &rows[rowBytes * (size_t)j];
				for (int k = 0; k < imgN; k++)
				{
					size_t sidx = (size_t)i * imgN + k;
					if (depth == 16) d[k] = (uint16_t)((r[sidx * 2] << 8) | r[sidx * 2 + 1]);
					else if (depth == 8) d[k] = r[sidx];
					else
					{
						size_t bit = sidx * depth;
						int v = (r[bit >> 3] >> (8 - depth - (bit & 7))) & ((1 << depth) - 1);
						d[k] = (uint16_t)(v * scale);
					}
				}
			}
	}
	// to 8-bit output
	const size_t n = (size_t)w * h;
	if (colourType == 3)
	{
		channels = palTrans ? 4 : 3;
		out.assign(n * channels, 0);
		for (size_t i = 0; i < n; i++)
		{
			// stb_image reads palette[index] without a range check; its table has 256 zero-initialised-or-stale
			// entries, so an out-of-range index is an invalid file: reject it
			if ((int)px[i] >= palLen) return false;
			for (int k = 0; k < channels; k++) out[i * channels + k] = palette[px[i]][k];
		}
		return true;
	}
	channels = imgN + (hasTrans ? 1 : 0);
	out.assign(n * channels, 0);
	unsigned key[3] = {tc[0], tc[1], tc[2]};
	if (hasTrans && depth < 16)
		for (int k = 0; k < 3; k++) key[k] = ((tc[k] & 255u) * (unsigned)(depth < 8 ? depthScale[depth] : 1)) & 255u; // stb: (uc)(v & 255) * scale
	for (size_t i = 0; i < n; i++)
	{
		const uint16_t* sp = &px[i * imgN];
		unsigned char* d = &out[i * channels];
		for (int k = 0; k < imgN; k++) d[k] = depth == 16 ? (unsigned char)(sp[k] >> 8) : (unsigned char)sp[k];
		if (hasTrans)
		{
			bool same = true;
			for (int k = 0; k < imgN; k++) same = same && sp[k] == key[k];
			d[imgN] = same ? 0 : 255;
		}
	}
	return true;
}

// Radiance .hdr -> float RGB, value = mantissa * 2^(e - 136) (the conversion stbi_loadf applies).
inline bool decodeHDR(const std::vector<unsigned char>& file, int& w, int& h, std::vector<float>& out)
{
	size_t pos = 0;
	auto line = [&](std::string& s) {
		s.clear();
		while (pos < file.size() && file[pos] != '\n') s.push_back((char)file[pos++]);
		if (pos < file.size()) pos++;
		return true;
	};
	std::string s;
	line(s);
	if (s != "#?RADIANCE" && s != "#?RGBE") return false;
	bool fmt = false;
	for (;;)
	{
		line(s);
		if (s.empty()) break;
		if (s == "FORMAT=32-bit_rle_rgbe") fmt = true;
		if (pos >= file.size()) return false;
	}
	if (!fmt) return false;
	line(s);
	if (sscanf(s.c_str(), "-Y %d +X %d", &h, &w) != 2 || w <= 0 || h <= 0) return false;
	out.assign((size_t)w * h * 3, 0.0f);
	std::vector<unsigned char> scan((size_t)w * 4);
	auto convert = [&](const unsigned char* rgbe, float* dst) {
		if (rgbe[3] != 0)
		{
			float f1 = (float)ldexp(1.0f, rgbe[3] - (int)(128 + 8));
			dst[0] = rgbe[0] * f1, dst[1] = rgbe[1] * f1, dst[2] = rgbe[2] * f1;
		}
		else
			dst[0] = dst[1] = dst[2] = 0.0f;
	};
	bool flat = (w < 8 || w >= 32768);
	for (int y = 0; y < h; y++)
	{
		if (!flat && pos + 4 <= file.size())
		{
			const unsigned char* p = &file[pos];
			if (!(p[0] == 2 && p[1] == 2 && !(p[2] & 0x80)))
			{
				if (y != 0) return false;
				flat = true; // not run-length encoded
			}
		}
		if (flat)
		{
			if (pos + (size_t)w * 4 > file.size()) return false;
			for (int x = 0; x < w; x++) convert(&file[pos + (size_t)x * 4], &out[((size_t)y * w + x) * 3]);
			pos += (size_t)w * 4;
			continue;
		}
		if (pos + 4 > file.size()) return false; // truncated before a scanline header
		if ((((int)file[pos + 2] << 8) | file[pos + 3]) != w) return false;
		pos += 4;
		for (int c = 0; c < 4; c++)
		{
			int x = 0;
			while (x < w)
			{
				if (pos >= file.size()) return false;
				int n = file[pos++];
				if (n > 128)
				{
					n -= 128;
					if (x + n > w || pos >= file.size()) return false;
					unsigned char v = file[pos++];
					for (int i = 0; i < n; i++) scan[(size_t)(x++) * 4 + c] = v;
				}
				else
				{
					if (n == 0 || x + n > w || pos + n > file.size()) return false;
					for (int i = 0; i < n; i++) scan[(size_t)(x++) * 4 + c] = file[pos++];
				}
			}
		}
		for (int x = 0; x < w; x++) convert(&scan[(size_t)x * 4], &out[((size_t)y * w + x) * 3]);
	}
	return true;
}
} // namespace rtb_img

class Texture
{
public:
	Colour* texels = NULL;
	float* alpha = NULL;
	int width = 0, height = 0, channels = 0;
	void loadDefault() // Imaging.h:24-31
	{
		width = height = 1;
		channels = 3;
		texels = new Colour[1];
		texels[0] = Colour(1.0f, 1.0f, 1.0f);
	}
	// Imaging.h:32-71.  A file that does not exist becomes the 1x1 white default exactly like the
	// reference (stb leaves width/height 0); a file that exists but cannot be decoded is an error.
	void load(std::string filename)
	{
		alpha = NULL;
		std::vector<unsigned char> file = rtb_img::readFile(filename);
		if (file.empty())
		{
			loadDefault();
			return;
		}
		if (filename.find(".hdr") != std::string::npos)
		{
			std::vector<float> px;
			if (!rtb_img::decodeHDR(file, width, height, px)) throw std::runtime_error("cannot decode " + filename);
			channels = 3;
			texels = new Colour[(size_t)width * height];
			for (size_t i = 0; i < (size_t)width * height; i++) texels[i] = Colour(px[i * 3], px[i * 3 + 1], px[i * 3 + 2]);
			return;
		}
		std::vector<unsigned char> px;
		bool isJPEG = file.size() > 2 && file[0] == 0xFF && file[1] == 0xD8;
		if (!(isJPEG ? rtb_img::decodeJPEG(file, width, height, channels, px) : rtb_img::decodePNG(file, width, height, channels, px)))
			throw std::runtime_error("cannot decode " + filename + " (PNG, Huffman JPEG and Radiance .hdr are supported)");
		if (channels < 3) throw std::runtime_error(filename + ": grey textures are read out of bounds by the reference (Imaging.h:60)");
		texels = new Colour[(size_t)width * height];
		for (size_t i = 0; i < (size_t)width * height; i++)
			texels[i] = Colour(px[i * channels] / 255.0f, px[i * channels + 1] / 255.0f, px[i * channels + 2] / 255.0f);
		if (channels == 4)
		{
			alpha = new float[(size_t)width * height];
			for (size_t i = 0; i < (size_t)width * height; i++) alpha[i] = px[i * channels + 3] / 255.0f;
		}
	}
};

class ImageFilter
{
public:
	virtual float filter(const float x, const float y) const = 0;
	virtual int size() const = 0;
	virtual ~ImageFilter() {}
};
class BoxFilter : public ImageFilter // Imaging.h:139-154
{
public:
	float filter(float x, float y) const { return (fabsf(x) <= 1.f && fabsf(y) <= 1.f) ? 1.0f : 0.0f; }
	int size() const { return 0; }
};
class GaussianFilter : public ImageFilter // Imaging.h:155-187
{
public:
	float radius, alpha;
	GaussianFilter(float r = 2.0, float a = 1.0) : radius(r), alpha(a) {}
	float Gaussian(float d) const { return expf(-alpha * (d * d)) - expf(-alpha * (radius * radius)); }
	float filter(float x, float y) const { return Gaussian(x) * Gaussian(y); }
	int size() const { return static_cast<int>(std::ceil(radius)); }
};

// Film: host copy of the running sums (Imaging.h:201-272); splat/tonemap run on the device.
class Film
{
public:
	Colour* film = NULL;
	unsigned int width = 0, height = 0;
	int SPP = 0;
	ImageFilter* filter = NULL;
	void init(int _width, int _height, ImageFilter* _filter)
	{
		width = _width, height = _height;
		film = new Colour[(size_t)width * height];
		clear();
		filter = _filter;
	}
	void clear()
	{
		memset((void*)film, 0, (size_t)width * height * sizeof(Colour));
		SPP = 0;
	}
	void incrementSPP() { SPP++; }
	void save(std::string filename); // Imaging.h:262-271: film / SPP as Radiance .hdr (rtb_standalone.hpp)
};

// ------------------------------------------------------------------------------------------
// Geometry.h
// ------------------------------------------------------------------------------------------
class Triangle
{
public:
	Vertex vertices[3];
	Vec3 e1, e2, n;
	float area = 0, d = 0;
	unsigned int materialIndex = 0;
	void init(Vertex v0, Vertex v1, Vertex v2, unsigned int _materialIndex) // Geometry.h:72-83
	{
		materialIndex = _materialIndex;
		vertices[0] = v0, vertices[1] = v1, vertices[2] = v2;
		e1 = vertices[2].p - vertices[1].p;
		e2 = vertices[0].p - vertices[2].p;
		n = e1.cross(e2).normalize();
		area = e1.cross(e2).length() * 0.5f;
		d = Dot(n, vertices[0].p);
	}
	Vec3 centre() const { return (vertices[0].p + vertices[1].p + vertices[2].p) / 3.0f; }
	Vec3 gNormal() const { return (n * (Dot(vertices[0].normal, n) > 0 ? 1.0f : -1.0f)); }
};

class AABB
{
public:
	Vec3 max, min;
	AABB() { reset(); }
	void reset()
	{
		max = Vec3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
		min = Vec3(FLT_MAX, FLT_MAX, FLT_MAX);
	}
	void extend(const Vec3 p)
	{
		max = Max(max, p);
		min = Min(min, p);
	}
	void extend(const AABB& box)
	{
		extend(box.min);
		extend(box.max);
	}
	float area() const // Geometry.h:187-191
	{
		Vec3 size = max - min;
		return ((size.x * size.y) + (size.y * size.z) + (size.x * size.z)) * 2.0f;
	}
};

class BVHNode
{
public:
	AABB bounds;
	BVHNode* r = NULL;
	BVHNode* l = NULL;
	int startIndex = 0, endIndex = 0;

	// Same tree and triangle order as the reference builder (Geometry.h:325-398); works on
	// (centroid key, index) pairs and permutes the 180-byte triangles once at the end.
	void build(std::vector<Triangle>& inputTriangles, std::vector<Triangle>& outputTriangles)
	{
		size_t n = inputTriangles.size();
		Work w;
		w.tris = &inputTriangles;
		w.order.resize(n);
		w.cx.resize(n), w.cy.resize(n), w.cz.resize(n);
		w.bmin.resize(n), w.bmax.resize(n);
		for (size_t i = 0; i < n; i++)
		{
			w.order[i] = (uint32_t)i;
			Vec3 c = inputTriangles[i].centre();
			w.cx[i] = c.x, w.cy[i] = c.y, w.cz[i] = c.z;
			AABB b;
			b.extend(inputTriangles[i].vertices[0].p);
			b.extend(inputTriangles[i].vertices[1].p);
			b.extend(inputTriangles[i].vertices[2].p);
			w.bmin[i] = b.min, w.bmax[i] = b.max;
		}
		unsigned hw = std::thread::hardware_concurrency();
		// par = log2(threads): the top `par` levels also split their own sort / sweep over 2^(par - level)
		// threads; below that every subtree of more than 50 k triangles still gets its own thread (SAH
		// splits are unbalanced: many more tasks than cores keeps them busy)
		int par = 0;
		while ((1u << par) < hw && par < 6) par++;
		if (n < 200000) par = -1;
		if (const char* e = getenv("RTB_HOST_BUILD_SERIAL"))
			if (atoi(e) != 0) par = -1;
		buildRecursive(w, 0, (int)n, par);
		std::vector<Triangle> sorted(n);
		for (size_t i = 0; i < n; i++) sorted[i] = inputTriangles[w.order[i]];
		outputTriangles.swap(sorted);
	}

	// test hook: does sortLikeStdSort give std::sort's permutation on these keys (ids = positions)?
	static bool sortSelfTest(const float* k, uint32_t n, int par)
	{
		std::vector<Key> a(n), b(n);
		for (uint32_t i = 0; i < n; i++) a[i] = b[i] = {k[i], i};
		std::sort(a.begin(), a.end(), keyLess);
		sortLikeStdSort(b.data(), b.data() + n, par);
		for (uint32_t i = 0; i < n; i++)
			if (a[i].id != b[i].id) return false;
		return true;
	}

private:
	struct Work
	{
		std::vector<Triangle>* tris;
		std::vector<uint32_t> order;
		std::vector<float> cx, cy, cz;
		std::vector<Vec3> bmin, bmax;
	};
	struct Key
	{
		float k;
		uint32_t id;
	};
	static bool keyLess(const Key& a, const Key& b) { return a.k < b.k; }
	// AABB without a constructor: the same selects (Core.h:187-195) and the same area expression
	struct BB
	{
		float lo[3], hi[3];
		void extend(const BB& o)
		{
			for (int a = 0; a < 3; a++)
			{
				hi[a] = hi[a] > o.lo[a] ? hi[a] : o.lo[a];
				lo[a] = lo[a] < o.lo[a] ? lo[a] : o.lo[a];
				hi[a] = hi[a] > o.hi[a] ? hi[a] : o.hi[a];
				lo[a] = lo[a] < o.hi[a] ? lo[a] : o.hi[a];
			}
		}
		float area() const // Geometry.h:187-191
		{
			float sx = hi[0] - lo[0], sy = hi[1] - lo[1], sz = hi[2] - lo[2];
			return ((sx * sy) + (sy * sz) + (sx * sz)) * 2.0f;
		}
	};
	template <class F>
	static void parallelFor(int n, F f)
	{
		std::vector<std::thread> th;
		for (int c = 1; c < n; c++) th.emplace_back([c, &f]() { f(c); });
		f(0);
		for (auto& t : th) t.join();
	}
	// std::sort's permutation (ties included) with its independent sub-ranges sorted concurrently.
	// libstdc++'s std::sort is: introsort loop (median-of-three to *first, unguarded Hoare partition,
	// recurse on the right part, iterate on the left, depth limit 2*floor(log2 n) -> heapsort) down to
	// ranges of <= 16 elements, then ONE insertion-sort pass over the whole array.  After the loop every
	// element of an earlier range is <= every element of a later one, so that pass never moves an element
	// out of its range: sorting each small range by insertion where the loop leaves it is the same
	// permutation.  The partition steps are restated here (they are the algorithm); the heapsort fallback
	// calls the library's own make_heap / sort_heap.  tests/test_host_cpu.py checks it against std::sort.
	static void insertionSort(Key* first, Key* last)
	{
		if (first == last) return;
		for (Key* i = first + 1; i != last; ++i)
		{
			Key val = *i;
			Key* j = i;
			while (j != first && keyLess(val, *(j - 1)))
			{
				*j = *(j - 1);
				--j;
			}
			*j = val;
		}
	}
	static Key* partitionPivot(Key* first, Key* last)
	{
		Key* mid = first + (last - first) / 2;
		Key *a = first + 1, *b = mid, *c = last - 1;
		if (keyLess(*a, *b))
		{
			if (keyLess(*b, *c)) std::iter_swap(first, b);
			else if (keyLess(*a, *c)) std::iter_swap(first, c);
			else std::iter_swap(first, a);
		}
		else if (keyLess(*a, *c)) std::iter_swap(first, a);
		else if (keyLess(*b, *c)) std::iter_swap(first, c);
		else std::iter_swap(first, b);
		Key* lo = first + 1;
		Key* hi = last;
		for (;;)
		{
			while (keyLess(*lo, *first)) ++lo;
			--hi;
			while (keyLess(*first, *hi)) --hi;
			if (!(lo < hi)) return lo;
			std::iter_swap(lo, hi);
			++lo;
		}
	}
	static void introsortLoop(Key* first, Key* last, long depthLimit, int par)
	{
		std::vector<std::thread> spawned;
		while (last - first > 16)
		{
			if (depthLimit == 0)
			{
				std::make_heap(first, last, keyLess);
				std::sort_heap(first, last, keyLess);
				first = last; // nothing left for the insertion sort below
				break;
			}
			--depthLimit;
			Key* cut = partitionPivot(first, last);
			if (par > 0 && last - cut > 20000)
			{
				--par;
				spawned.emplace_back([cut, last, depthLimit, par]() { introsortLoop(cut, last, depthLimit, par); });
			}
			else
				introsortLoop(cut, last, depthLimit, 0);
			last = cut;
		}
		insertionSort(first, last);
		for (auto& t : spawned) t.join();
	}
	static void sortLikeStdSort(Key* first, Key* last, int par)
	{
		if (first == last) return;
		long lg = 0;
		for (size_t n = (size_t)(last - first); n > 1; n >>= 1) lg++;
		introsortLoop(first, last, lg * 2, par + 2);
	}
	// Sub-ranges are independent (disjoint slices of `order`, read-only keys and boxes), so the two
	// recursive calls may run concurrently without changing any decision.
	void buildRecursive(Work& w, int start, int end, int par)
	{
		bounds.reset();
		for (int i = start; i < end; i++)
		{
			// union of the three vertices == union of the triangle's box corners
			bounds.extend(w.bmin[w.order[i]]);
			bounds.extend(w.bmax[w.order[i]]);
		}
		int numTri = end - start;
		if (numTri <= MAXNODE_TRIANGLES)
		{
			startIndex = start;
			endIndex = end;
			return;
		}
		int axis = 0;
		Vec3 size = bounds.max - bounds.min;
		if (size.y >= size.x && size.y >= size.z) axis = 1;
		else if (size.z >= size.x && size.z >= size.y) axis = 2;
		const std::vector<float>& key = axis == 0 ? w.cx : (axis == 1 ? w.cy : w.cz);
		// scratch: the sort keys and the prefix / suffix boxes of the SAH sweep.  Millions of nodes hold a
		// handful of triangles: those use the stack; plain structs, never value-initialised.
		const int SMALL = 96;
		Key keysS[SMALL];
		BB leftS[SMALL], rightS[SMALL];
		std::unique_ptr<Key[]> keysH;
		std::unique_ptr<BB[]> leftH, rightH;
		Key* keys = keysS;
		BB *left = leftS, *right = rightS;
		if (numTri > SMALL)
		{
			keysH.reset(new Key[(size_t)numTri]);
			leftH.reset(new BB[(size_t)numTri]);
			rightH.reset(new BB[(size_t)numTri]);
			keys = keysH.get(), left = leftH.get(), right = rightH.get();
		}
		for (int i = 0; i < numTri; i++) keys[i] = {key[w.order[start + i]], w.order[start + i]};
		// big nodes near the root: the node's own sort and sweep are spread over the threads that have no
		// subtree of their own yet (2^par of them); the results are those of the serial code
		const int team = (par > 0 && numTri > 400000) ? (1 << par) : 1;
		const bool fork = par >= 0 && numTri > 50000;
		if (team > 1) sortLikeStdSort(keys, keys + numTri, par);
		else std::sort(keys, keys + numTri, keyLess);
		for (int i = 0; i < numTri; i++) w.order[start + i] = keys[i].id;
		auto triBox = [&](int i) {
			BB b;
			const Vec3 &lo = w.bmin[keys[i].id], &hi = w.bmax[keys[i].id];
			b.lo[0] = lo.x, b.lo[1] = lo.y, b.lo[2] = lo.z, b.hi[0] = hi.x, b.hi[1] = hi.y, b.hi[2] = hi.z;
			return b;
		};
		auto scanChunk = [&](int a, int b) {
			left[a] = triBox(a);
			for (int i = a + 1; i < b; i++)
			{
				left[i] = left[i - 1];
				left[i].extend(triBox(i));
			}
			right[b - 1] = triBox(b - 1);
			for (int i = b - 2; i >= a; i--)
			{
				right[i] = right[i + 1];
				right[i].extend(triBox(i));
			}
		};
		auto bestSplit = [&](int a, int b, float& cm, int& si) {
			cm = FLT_MAX, si = 0;
			for (int i = a; i < b; i++)
			{
				float num_left = (float)i, num_right = (float)(numTri - i);
				float cost = left[i - 1].area() * num_left + right[i].area() * num_right;
				if (cost < cm) cm = cost, si = i;
			}
		};
		float cost_min = FLT_MAX;
		int split_index = 0;
		if (team > 1)
		{
			// chunked scans: box unions are exact min/max, so any grouping gives the serial boxes
			const int T = team, chunk = (numTri + T - 1) / T;
			std::vector<BB> headL((size_t)T), headR((size_t)T);
			parallelFor(T, [&](int c) {
				int a = c * chunk, b = std::min(numTri, a + chunk);
				if (a < b) scanChunk(a, b);
			});
			for (int c = 0; c < T; c++)
			{
				int a = c * chunk, b = std::min(numTri, a + chunk);
				if (a >= b) continue;
				headL[c] = left[b - 1], headR[c] = right[a];
			}
			for (int c = 1; c < T; c++)
				if (c * chunk < numTri) headL[c].extend(headL[c - 1]);
			for (int c = T - 2; c >= 0; c--)
				if ((c + 1) * chunk < numTri) headR[c].extend(headR[c + 1]);
			parallelFor(T, [&](int c) {
				int a = c * chunk, b = std::min(numTri, a + chunk);
				if (a >= b) return;
				if (c > 0)
					for (int i = a; i < b; i++) left[i].extend(headL[c - 1]);
				if (c + 1 < T && (c + 1) * chunk < numTri)
					for (int i = a; i < b; i++) right[i].extend(headR[c + 1]);
			});
			std::vector<float> bestCost((size_t)T, FLT_MAX);
			std::vector<int> bestIdx((size_t)T, 0);
			parallelFor(T, [&](int c) {
				int a = std::max(1, c * chunk), b = std::min(numTri, c * chunk + chunk);
				if (a < b) bestSplit(a, b, bestCost[c], bestIdx[c]);
			});
			for (int c = 0; c < T; c++) // first index of the minimum, like the serial strict '<'
				if (bestCost[c] < cost_min) cost_min = bestCost[c], split_index = bestIdx[c];
		}
		else
		{
			scanChunk(0, numTri);
			bestSplit(1, numTri, cost_min, split_index);
		}
		int mid = start + split_index;
		l = new BVHNode();
		r = new BVHNode();
		keysH.reset(), leftH.reset(), rightH.reset(); // release the sweep scratch before descending
		if (fork)
		{
			const int next = par > 0 ? par - 1 : 0;
			BVHNode* lp = l;
			std::thread th([&w, lp, start, mid, next]() { lp->buildRecursive(w, start, mid, next); });
			r->buildRecursive(w, mid, end, next);
			th.join();
		}
		else
		{
			l->buildRecursive(w, start, mid, -1);
			r->buildRecursive(w, mid, end, -1);
		}
	}
};

// ------------------------------------------------------------------------------------------
// Materials.h (data + the three virtual classifications the flattener needs)
// ------------------------------------------------------------------------------------------
class BSDF
{
public:
	Colour emission;
	virtual bool isPureSpecular() = 0;
	virtual bool isTwoSided() = 0;
	bool isLight() { return emission.Lum() > 0 ? true : false; }
	void addLight(Colour _emission) { emission = _emission; }
	virtual ~BSDF() {}
};
#define RTB_SIMPLE_BSDF(NAME, SPECULAR, TWOSIDED)            \
	class NAME : public BSDF                                 \
	{                                                        \
	public:                                                  \
		Texture* albedo = NULL;                              \
		float intIOR = 1.33f, extIOR = 1.0f, alpha = 0, sigma = 0; \
		Colour eta, k;                                       \
		bool isPureSpecular() { return SPECULAR; }           \
		bool isTwoSided() { return TWOSIDED; }               \
	}
RTB_SIMPLE_BSDF(DiffuseBSDF, false, true);     // Materials.h:118
RTB_SIMPLE_BSDF(MirrorBSDF, true, true);       // :158
RTB_SIMPLE_BSDF(ConductorBSDF, false, true);   // :203
RTB_SIMPLE_BSDF(GlassBSDF, true, false);       // :252
RTB_SIMPLE_BSDF(DielectricBSDF, false, false); // :320
RTB_SIMPLE_BSDF(OrenNayarBSDF, false, true);   // :369
RTB_SIMPLE_BSDF(PlasticBSDF, false, true);     // :414
class LayeredBSDF : public BSDF                // :467
{
public:
	BSDF* base = NULL;
	Colour sigmaa;
	float thickness = 0, intIOR = 1.33f, extIOR = 1.0f;
	bool isPureSpecular() { return base->isPureSpecular(); }
	bool isTwoSided() { return true; }
};

// ------------------------------------------------------------------------------------------
// Lights.h
// ------------------------------------------------------------------------------------------
class Light
{
public:
	virtual bool isArea() = 0;
	virtual float totalIntegratedPower() = 0;
	virtual ~Light() {}
};
class AreaLight : public Light
{
public:
	Triangle* triangle = NULL;
	Colour emission;
	bool isArea() { return true; }
	float totalIntegratedPower() { return (triangle->area * emission.Lum()); }
};
class BackgroundColour : public Light
{
public:
	Colour emission;
	BackgroundColour(Colour _emission) : emission(_emission) {}
	bool isArea() { return false; }
	float totalIntegratedPower() { return emission.Lum() * 4.0f * M_PI; }
};
class EnvironmentMap : public Light
{
public:
	Texture* env;
	EnvironmentMap(Texture* _env) : env(_env) {}
	bool isArea() { return false; }
	float totalIntegratedPower() // Lights.h:177-190
	{
		float total = 0;
		for (int i = 0; i < env->height; i++)
		{
			float st = sinf(((float)i / (float)env->height) * M_PI);
			for (int n = 0; n < env->width; n++) total += (env->texels[(i * env->width) + n].Lum() * st);
		}
		total = total / (float)(env->width * env->height);
		return total * 4.0f * M_PI;
	}
};

// ------------------------------------------------------------------------------------------
// Scene.h
// ------------------------------------------------------------------------------------------
class Camera
{
public:
	Matrix projectionMatrix, inverseProjectionMatrix, camera, cameraToView;
	float width = 0, height = 0;
	Vec3 origin, viewDirection;
	float Afilm = 0;
	void init(Matrix ProjectionMatrix, int screenwidth, int screenheight) // Scene.h:22-32
	{
		projectionMatrix = ProjectionMatrix;
		inverseProjectionMatrix = ProjectionMatrix.invert();
		width = (float)screenwidth;
		height = (float)screenheight;
		float Wlens = (2.0f / ProjectionMatrix.a[1][1]);
		float aspect = ProjectionMatrix.a[0][0] / ProjectionMatrix.a[1][1];
		Afilm = Wlens * (Wlens * aspect);
	}
	void updateView(Matrix V) // Scene.h:33-41
	{
		camera = V;
		cameraToView = V.invert();
		origin = camera.mulPoint(Vec3(0, 0, 0));
		viewDirection = inverseProjectionMatrix.mulPointAndPerspectiveDivide(Vec3(0, 0, 1));
		viewDirection = camera.mulVec(viewDirection);
		viewDirection = viewDirection.normalize();
	}
};

class Scene
{
public:
	std::vector<Triangle> triangles;
	std::vector<BSDF*> materials;
	std::vector<Light*> lights;
	Light* background = NULL;
	BVHNode* bvh = NULL;
	Camera camera;
	AABB bounds;
	void init(std::vector<Triangle> meshTriangles, std::vector<BSDF*> meshMaterials, Light* _background) // Scene.h:142-160
	{
		for (size_t i = 0; i < meshTriangles.size(); i++)
		{
			triangles.push_back(meshTriangles[i]);
			bounds.extend(meshTriangles[i].vertices[0].p);
			bounds.extend(meshTriangles[i].vertices[1].p);
			bounds.extend(meshTriangles[i].vertices[2].p);
		}
		for (size_t i = 0; i < meshMaterials.size(); i++) materials.push_back(meshMaterials[i]);
		background = _background;
		if (background->totalIntegratedPower() > 0) lights.push_back(background);
	}
	void build() // Scene.h:82-106
	{
		std::vector<Triangle> input;
		input.swap(triangles);
		bvh = new BVHNode();
		bvh->build(input, triangles);
		for (size_t i = 0; i < triangles.size(); i++)
		{
			if (materials[triangles[i].materialIndex]->isLight())
			{
				AreaLight* light = new AreaLight();
				light->triangle = &triangles[i];
				light->emission = materials[triangles[i].materialIndex]->emission;
				lights.push_back(light);
			}
		}
	}
};

// ------------------------------------------------------------------------------------------
// GEMLoader.h: scene.json (strings for scalars, arrays of instance objects) and .gem meshes
// ------------------------------------------------------------------------------------------
namespace GEMLoader
{
class GEMProperty
{
public:
	std::string name, value;
	GEMProperty() {}
	GEMProperty(std::string n) : name(n) {}
	std::string getValue(std::string = "") { return value; }
	float getValue(float _default)
	{
		try { return std::stof(value); }
		catch (...) { return _default; }
	}
	int getValue(int _default)
	{
		try { return std::stoi(value); }
		catch (...) { return _default; }
	}
	void getValuesAsVector3(float& x, float& y, float& z, char sep = ' ', float _default = 0)
	{
		std::vector<float> v;
		std::stringstream ss(value);
		std::string word;
		while (std::getline(ss, word, sep))
		{
			try { v.push_back(std::stof(word)); }
			catch (...) { v.push_back(_default); }
		}
		while (v.size() < 3) v.push_back(_default);
		x = v[0], y = v[1], z = v[2];
	}
};
class GEMMaterial
{
public:
	std::vector<GEMProperty> properties;
	GEMProperty find(std::string name)
	{
		for (auto& p : properties)
			if (p.name == name) return p;
		return GEMProperty(name);
	}
};
struct GEMMatrix
{
	float m[16];
};
class GEMInstance
{
public:
	GEMMatrix w;
	std::string meshFilename;
	GEMMaterial material;
};
struct GEMStaticVertex // 44 bytes on disk (GEMLoader.h static vertex)
{
	float position[3], normal[3], tangent[3], u, v;
};
struct GEMMesh
{
	std::vector<GEMStaticVertex> verticesStatic;
	std::vector<unsigned int> indices;
};

// Minimal JSON value: what scene.json uses (objects, arrays, strings, numbers, true/false/null).
struct Json
{
	enum Type { Null, Bool, Number, String, Array, Object } type = Null;
	bool b = false;
	float num = 0;
	std::string str;
	std::vector<Json> arr;
	std::map<std::string, Json> obj; // key-sorted iteration like the reference's std::map
	std::string asStr() const // GEMLoader.h:459-480
	{
		switch (type)
		{
		case Bool: return std::to_string(b);
		case Number: return std::to_string(num);
		case String: return str;
		default: return "";
		}
	}
};
// Strict enough to reject a damaged scene.json with a position instead of loading an empty scene
// (the reference's own ad-hoc parser, GEMLoader.h:380-712, reads garbage silently).
class JsonParser
{
	const std::string& s;
	size_t pos = 0;
	void ws()
	{
		while (pos < s.size() && isspace((unsigned char)s[pos])) pos++;
	}
	char peek() const { return pos < s.size() ? s[pos] : 0; }
	[[noreturn]] void fail(const char* what) const
	{
		throw std::runtime_error("JSON: " + std::string(what) + " at byte " + std::to_string(pos));
	}
	void expect(char c, const char* what)
	{
		ws();
		if (peek() != c) fail(what);
		pos++;
	}
	void literal(const char* word)
	{
		size_t n = strlen(word);
		if (s.compare(pos, n, word) != 0) fail("unknown literal");
		pos += n;
	}
	std::string string()
	{
		std::string out;
		pos++; // opening quote
		for (;;)
		{
			if (pos >= s.size()) fail("unterminated string");
			char c = s[pos++];
			if (c == '"') break;
			if (c == '\\')
			{
				if (pos >= s.size()) fail("unterminated escape");
				char e = s[pos++];
				switch (e)
				{
				case 'n': out.push_back('\n'); break;
				case 't': out.push_back('\t'); break;
				case 'r': out.push_back('\r'); break;
				case 'b': out.push_back('\b'); break;
				case 'f': out.push_back('\f'); break;
				case 'u':
				{
					if (pos + 4 > s.size()) fail("short \\u escape");
					unsigned v = (unsigned)strtoul(s.substr(pos, 4).c_str(), nullptr, 16);
					pos += 4;
					out.push_back(v < 128 ? (char)v : '?'); // file names in scene.json are ASCII
					break;
				}
				default: out.push_back(e); break; // \" \\ \/
				}
			}
			else
				out.push_back(c);
		}
		return out;
	}

public:
	JsonParser(const std::string& text) : s(text) {}
	Json document()
	{
		Json j = value(0);
		ws();
		if (pos != s.size()) fail("trailing characters");
		return j;
	}
	Json value(int depth = 0)
	{
		if (depth > 64) fail("nesting too deep");
		ws();
		Json j;
		char c = peek();
		if (c == '{')
		{
			j.type = Json::Object;
			pos++;
			ws();
			if (peek() == '}')
			{
				pos++;
				return j;
			}
			for (;;)
			{
				ws();
				if (peek() != '"') fail("expected a member name");
				std::string key = string();
				expect(':', "expected ':'");
				j.obj[key] = value(depth + 1);
				ws();
				char d = peek();
				if (d != ',' && d != '}') fail("expected ',' or '}'");
				pos++;
				if (d == '}') break;
			}
		}
		else if (c == '[')
		{
			j.type = Json::Array;
			pos++;
			ws();
			if (peek() == ']')
			{
				pos++;
				return j;
			}
			for (;;)
			{
				j.arr.push_back(value(depth + 1));
				ws();
				char d = peek();
				if (d != ',' && d != ']') fail("expected ',' or ']'");
				pos++;
				if (d == ']') break;
			}
		}
		else if (c == '"')
		{
			j.type = Json::String;
			j.str = string();
		}
		else if (c == 't' || c == 'f')
		{
			j.type = Json::Bool;
			j.b = (c == 't');
			literal(j.b ? "true" : "false");
		}
		else if (c == 'n')
		{
			literal("null");
		}
		else if (c == '-' || isdigit((unsigned char)c))
		{
			size_t start = pos;
			if (peek() == '-') pos++;
			if (!isdigit((unsigned char)peek())) fail("malformed number");
			while (isdigit((unsigned char)peek())) pos++;
			if (peek() == '.')
			{
				pos++;
				while (isdigit((unsigned char)peek())) pos++;
			}
			if (peek() == 'e' || peek() == 'E')
			{
				pos++;
				if (peek() == '+' || peek() == '-') pos++;
				while (isdigit((unsigned char)peek())) pos++;
			}
			j.type = Json::Number;
			j.num = std::stof(s.substr(start, pos - start));
		}
		else
			fail(c ? "unexpected character" : "unexpected end of input");
		return j;
	}
};

class GEMScene
{
public:
	std::vector<GEMInstance> instances;
	std::vector<GEMProperty> sceneProperties;
	void load(std::string filename) // GEMLoader.h:714-737
	{
		std::ifstream file(filename);
		if (!file) throw std::runtime_error("cannot open " + filename);
		std::stringstream buffer;
		buffer << file.rdbuf();
		std::string content = buffer.str();
		Json data = JsonParser(content).document();
		if (data.type != Json::Object) throw std::runtime_error(filename + ": the top level of scene.json must be an object");
		for (const auto& item : data.obj)
		{
			if (item.second.type != Json::Array)
			{
				GEMProperty p(item.first);
				p.value = item.second.asStr();
				sceneProperties.push_back(p);
				continue;
			}
			for (const Json& inst : item.second.arr)
			{
				GEMInstance gi;
				memset(&gi.w, 0, sizeof(gi.w));
				for (const auto& kv : inst.obj)
				{
					if (kv.first == "filename") gi.meshFilename = kv.second.asStr();
					else if (kv.first == "world")
					{
						for (int i = 0; i < 16 && i < (int)kv.second.arr.size(); i++) gi.w.m[i] = kv.second.arr[i].num;
					}
					else
					{
						GEMProperty p(kv.first);
						p.value = kv.second.asStr();
						gi.material.properties.push_back(p);
					}
				}
				instances.push_back(gi);
			}
		}
	}
	GEMProperty findProperty(std::string name)
	{
		for (auto& p : sceneProperties)
			if (p.name == name) return p;
		return GEMProperty(name);
	}
};

class GEMModelLoader
{
public:
	// GEMLoader.h:344-365: u32 magic 0xF1EF0001, u32 isAnimated, u32 meshes; per mesh: u32 props x
	// {i32 len, bytes, i32 len, bytes}; u32 vertices x 44 B; u32 indices x u32.
	void load(std::string filename, std::vector<GEMMesh>& meshes)
	{
		std::vector<unsigned char> f = rtb_img::readFile(filename);
		size_t pos = 0;
		auto u32 = [&]() {
			if (pos + 4 > f.size()) throw std::runtime_error(filename + ": truncated");
			uint32_t v;
			memcpy(&v, &f[pos], 4);
			pos += 4;
			return v;
		};
		if (f.size() < 12 || u32() != 4058972161u)
		{
			// the reference prints this and exit(0)s
			throw std::runtime_error(filename + " is not a GE Model File");
		}
		uint32_t isAnimated = u32();
		uint32_t n = u32();
		if (isAnimated) throw std::runtime_error(filename + ": animated meshes are not part of the path-tracing path");
		for (uint32_t i = 0; i < n; i++)
		{
			GEMMesh mesh;
			uint32_t props = u32();
			for (uint32_t p = 0; p < props * 2; p++)
			{
				uint32_t len = u32();
				pos += len;
			}
			uint32_t nv = u32();
			if (pos + (size_t)nv * 44 > f.size()) throw std::runtime_error(filename + ": truncated");
			mesh.verticesStatic.resize(nv);
			memcpy(mesh.verticesStatic.data(), &f[pos], (size_t)nv * 44);
			pos += (size_t)nv * 44;
			uint32_t ni = u32();
			if (pos + (size_t)ni * 4 > f.size()) throw std::runtime_error(filename + ": truncated");
			mesh.indices.resize(ni);
			memcpy(mesh.indices.data(), &f[pos], (size_t)ni * 4);
			pos += (size_t)ni * 4;
			meshes.push_back(mesh);
		}
	}
};
} // namespace GEMLoader

// ------------------------------------------------------------------------------------------
// SceneLoader.h
// ------------------------------------------------------------------------------------------
static_assert(sizeof(GEMLoader::GEMStaticVertex) == 44, "static .gem vertex is 44 bytes");

inline Texture* loadTexture(std::string filename, std::map<std::string, Texture*>& textureManager) // SceneLoader.h:92-102
{
	auto it = textureManager.find(filename);
	if (it != textureManager.end()) return it->second;
	Texture* t = new Texture();
	t->load(filename);
	textureManager.insert({filename, t});
	return t;
}

inline void loadInstance(std::string sceneName, std::vector<Triangle>& meshTriangles, std::vector<BSDF*>& meshMaterials,
                         GEMLoader::GEMInstance& instance, std::map<std::string, Texture*>& textureManager) // SceneLoader.h:104-235
{
	GEMLoader::GEMModelLoader loader;
	std::vector<GEMLoader::GEMMesh> meshes;
	loader.load(sceneName + "/" + instance.meshFilename, meshes);
	GEMLoader::GEMMaterial& mp = instance.material;
	std::string kind = mp.find("bsdf").getValue("");
	auto albedo = [&]() { return loadTexture(sceneName + "/" + mp.find("reflectance").getValue(""), textureManager); };
	auto alphaOf = [&](float roughness) { return 1.62142f * sqrtf(roughness); };
	BSDF* material = NULL;
	// the reference tests the seven names one after the other (:111-172)
	if (kind == "diffuse")
	{
		DiffuseBSDF* b = new DiffuseBSDF();
		b->albedo = albedo();
		material = b;
	}
	else if (kind == "orennayar")
	{
		OrenNayarBSDF* b = new OrenNayarBSDF();
		b->albedo = albedo();
		b->sigma = mp.find("alpha").getValue(1.0f);
		material = b;
	}
	else if (kind == "glass" || kind == "dielectric")
	{
		Texture* tex = albedo();
		float intIOR = mp.find("intIOR").getValue(1.33f), extIOR = mp.find("extIOR").getValue(1.0f);
		float roughness = mp.find("roughness").getValue(1.0f);
		if (kind == "glass" || roughness < 0.001f)
		{
			GlassBSDF* b = new GlassBSDF();
			b->albedo = tex, b->intIOR = intIOR, b->extIOR = extIOR;
			material = b;
		}
		else
		{
			DielectricBSDF* b = new DielectricBSDF();
			b->albedo = tex, b->intIOR = intIOR, b->extIOR = extIOR, b->alpha = alphaOf(roughness);
			material = b;
		}
	}
	else if (kind == "mirror")
	{
		MirrorBSDF* b = new MirrorBSDF();
		b->albedo = albedo();
		material = b;
	}
	else if (kind == "plastic")
	{
		PlasticBSDF* b = new PlasticBSDF();
		b->albedo = albedo();
		b->intIOR = mp.find("intIOR").getValue(1.33f), b->extIOR = mp.find("extIOR").getValue(1.0f);
		b->alpha = alphaOf(mp.find("roughness").getValue(1.0f));
		material = b;
	}
	else if (kind == "conductor")
	{
		ConductorBSDF* b = new ConductorBSDF();
		b->albedo = albedo();
		mp.find("eta").getValuesAsVector3(b->eta.r, b->eta.g, b->eta.b);
		mp.find("k").getValuesAsVector3(b->k.r, b->k.g, b->k.b);
		b->alpha = alphaOf(mp.find("roughness").getValue(1.0f));
		material = b;
	}
	if (material) meshMaterials.push_back(material);
	if (material && mp.find("emission").getValue("") != "")
	{
		Colour e;
		mp.find("emission").getValuesAsVector3(e.r, e.g, e.b);
		material->addLight(e);
	}
	if (material && mp.find("coatingThickness").getValue(0) > 0)
	{
		// :179-188 — built AFTER the base was registered, so the layered object is never the one
		// in Scene::materials; kept for fidelity of the object graph
		LayeredBSDF* lay = new LayeredBSDF();
		lay->base = material;
		mp.find("coatingSigmaA").getValuesAsVector3(lay->sigmaa.r, lay->sigmaa.g, lay->sigmaa.b);
		lay->intIOR = mp.find("coatingIntIOR").getValue(1.33f), lay->extIOR = mp.find("coatingExtIOR").getValue(1.0f);
		lay->thickness = mp.find("coatingThickness").getValue(0.0f);
		material = lay;
	}
	if (material == NULL)
	{
		fprintf(stderr, "Error in loading\n"); // :189-194: flag it, keep loading the rest
		return;
	}
	int materialIndex = (int)meshMaterials.size() - 1;
	Matrix transform;
	memcpy(transform.m, instance.w.m, 16 * sizeof(float));
	Matrix vecTransform = transform.invert().transpose();
	std::vector<Vertex> vertices;
	std::vector<unsigned int> indices;
	for (size_t i = 0; i < meshes.size(); i++)
	{
		for (const GEMLoader::GEMStaticVertex& gv : meshes[i].verticesStatic)
		{
			Vertex v;
			v.p = transform.mulPoint(Vec3(gv.position[0], gv.position[1], gv.position[2]));
			v.normal = vecTransform.mulVec(Vec3(gv.normal[0], gv.normal[1], gv.normal[2])).normalize();
			v.u = gv.u, v.v = gv.v;
			vertices.push_back(v);
		}
		int offset = (int)indices.size(); // sic (:221): the offset counts INDICES, as in the reference
		for (unsigned int idx : meshes[i].indices) indices.push_back(offset + idx);
	}
	// a damaged file (or the index-counting offset above on a multi-mesh model) can point past the vertex array:
	// the reference reads out of bounds there; this loader reports it
	for (size_t i = 0; i < indices.size(); i++)
		if (indices[i] >= vertices.size())
			throw std::runtime_error("mesh index " + std::to_string(i) + " = " + std::to_string(indices[i]) + " is out of range (" +
			                         std::to_string(vertices.size()) + " vertices)");
	for (size_t i = 0; i + 2 < indices.size(); i += 3)
	{
		Triangle t;
		t.init(vertices[indices[i]], vertices[indices[i + 1]], vertices[indices[i + 2]], materialIndex);
		if (t.area > 0) meshTriangles.push_back(t);
	}
}

inline Scene* loadScene(std::string sceneName) // SceneLoader.h:237-291
{
	Scene* scene = new Scene();
	GEMLoader::GEMScene gemscene;
	gemscene.load(sceneName + "/scene.json");
	int width = gemscene.findProperty("width").getValue(1920);
	int height = gemscene.findProperty("height").getValue(1080);
	float fov = gemscene.findProperty("fov").getValue(45.0f);
	Matrix P = Matrix::perspective(0.001f, 10000.0f, (float)width / (float)height, fov);
	Vec3 from, to, up;
	gemscene.findProperty("from").getValuesAsVector3(from.x, from.y, from.z);
	gemscene.findProperty("to").getValuesAsVector3(to.x, to.y, to.z);
	gemscene.findProperty("up").getValuesAsVector3(up.x, up.y, up.z);
	Matrix V = Matrix::lookAt(from, to, up).invert();
	if (gemscene.findProperty("flipX").getValue(0) == 1) P.a[0][0] = -P.a[0][0];
	scene->camera.init(P, width, height);
	scene->camera.updateView(V);
	std::vector<Triangle> meshTriangles;
	std::vector<BSDF*> meshMaterials;
	std::map<std::string, Texture*> textureManager;
	for (size_t i = 0; i < gemscene.instances.size(); i++)
		loadInstance(sceneName, meshTriangles, meshMaterials, gemscene.instances[i], textureManager);
	Light* background;
	std::string env = gemscene.findProperty("envmap").getValue("");
	if (env != "") background = new EnvironmentMap(loadTexture(sceneName + "/" + env, textureManager));
	else background = new BackgroundColour(Colour(0.0f, 0.0f, 0.0f));
	scene->init(meshTriangles, meshMaterials, background);
	scene->build();
	return scene;
}
