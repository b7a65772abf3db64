// rtb_standalone.hpp — what host/Renderer.h and a Main.cpp-style program need besides the scene API of
// rtb_scene.hpp when NO reference header is on the include path: the window stand-in
// (GamesEngineeringBase::Window as used at RTBase/Renderer.h:35,45,77,897), MTRandom's name
// (Sampling.h:13-26; the GPU path draws from a counter-based generator), Film::save (Imaging.h:262-271:
// film / SPP written as Radiance .hdr) and stbi_write_png (8-bit RGB through zlib).
#pragma once
#include "rtb_scene.hpp"

#include <random>

namespace GamesEngineeringBase
{
// Headless: a back buffer with the original's accessors.
class Window
{
public:
	void create(unsigned int w, unsigned int h, const std::string& = "", float = 1.0f)
	{
		width = w, height = h;
		buffer.assign((size_t)w * h * 3, 0);
	}
	void draw(unsigned int x, unsigned int y, unsigned char r, unsigned char g, unsigned char b)
	{
		if (x >= width || y >= height) return;
		unsigned char* p = &buffer[((size_t)y * width + x) * 3];
		p[0] = r, p[1] = g, p[2] = b;
	}
	unsigned char* getBackBuffer() { return buffer.data(); }
	unsigned int getWidth() const { return width; }
	unsigned int getHeight() const { return height; }
	void checkInput() {}
	void clear() {}
	void present() {}
	bool keyPressed(int) const { return false; }

private:
	unsigned int width = 0, height = 0;
	std::vector<unsigned char> buffer;
};
} // namespace GamesEngineeringBase

class Sampler
{
public:
	virtual float next() = 0;
	virtual ~Sampler() {}
};
class MTRandom : public Sampler // Sampling.h:13-26
{
public:
	std::mt19937 generator;
	std::uniform_real_distribution<float> dist;
	MTRandom(unsigned int seed = 1) : dist(0.0f, 1.0f) { generator.seed(seed); }
	float next() { return dist(generator); }
};

namespace rtb_img
{
// Radiance RGBE, flat (un-compressed) scanlines, top to bottom: "-Y h +X w".
inline bool writeHDR(const std::string& path, int w, int h, const float* rgb)
{
	FILE* f = fopen(path.c_str(), "wb");
	if (!f) return false;
	fprintf(f, "#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y %d +X %d\n", h, w);
	std::vector<unsigned char> row((size_t)w * 4);
	for (int y = 0; y < h; y++)
	{
		for (int x = 0; x < w; x++)
		{
			const float* p = rgb + ((size_t)y * w + x) * 3;
			float m = std::max(p[0], std::max(p[1], p[2]));
			unsigned char* o = &row[(size_t)x * 4];
			if (!(m > 1e-32f)) o[0] = o[1] = o[2] = o[3] = 0;
			else
			{
				int e;
				float s = frexpf(m, &e) * 256.0f / m;
				o[0] = (unsigned char)(p[0] * s), o[1] = (unsigned char)(p[1] * s), o[2] = (unsigned char)(p[2] * s);
				o[3] = (unsigned char)(e + 128);
			}
		}
		fwrite(row.data(), 1, row.size(), f);
	}
	fclose(f);
	return true;
}
inline void pngChunk(FILE* f, const char* type, const unsigned char* data, size_t n)
{
	unsigned char len[4] = {(unsigned char)(n >> 24), (unsigned char)(n >> 16), (unsigned char)(n >> 8), (unsigned char)n};
	fwrite(len, 1, 4, f);
	fwrite(type, 1, 4, f);
	if (n) fwrite(data, 1, n, f);
	uLong crc = crc32(0L, (const Bytef*)type, 4);
	if (n) crc = crc32(crc, data, (uInt)n);
	unsigned char c[4] = {(unsigned char)(crc >> 24), (unsigned char)(crc >> 16), (unsigned char)(crc >> 8), (unsigned char)crc};
	fwrite(c, 1, 4, f);
}
inline bool writePNG(const std::string& path, int w, int h, int comp, const unsigned char* data, int stride)
{
	if (comp != 3 && comp != 4 && comp != 1) return false;
	FILE* f = fopen(path.c_str(), "wb");
	if (!f) return false;
	static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
	fwrite(sig, 1, 8, f);
	unsigned char ihdr[13] = {(unsigned char)(w >> 24), (unsigned char)(w >> 16), (unsigned char)(w >> 8), (unsigned char)w,
	                          (unsigned char)(h >> 24), (unsigned char)(h >> 16), (unsigned char)(h >> 8), (unsigned char)h,
	                          8, (unsigned char)(comp == 3 ? 2 : (comp == 4 ? 6 : 0)), 0, 0, 0};
	pngChunk(f, "IHDR", ihdr, 13);
	std::vector<unsigned char> raw((size_t)h * ((size_t)w * comp + 1));
	for (int y = 0; y < h; y++)
	{
		raw[(size_t)y * ((size_t)w * comp + 1)] = 0; // filter: none
		memcpy(&raw[(size_t)y * ((size_t)w * comp + 1) + 1], data + (size_t)y * stride, (size_t)w * comp);
	}
	uLongf bound = compressBound((uLong)raw.size());
	std::vector<unsigned char> z(bound);
	if (compress2(z.data(), &bound, raw.data(), (uLong)raw.size(), 6) != Z_OK)
	{
		fclose(f);
		return false;
	}
	pngChunk(f, "IDAT", z.data(), bound);
	pngChunk(f, "IEND", nullptr, 0);
	fclose(f);
	return true;
}
} // namespace rtb_img

inline void Film::save(std::string filename) // Imaging.h:262-271
{
	std::vector<float> mean((size_t)width * height * 3);
	float inv = SPP > 0 ? 1.0f / (float)SPP : 0.0f;
	for (size_t i = 0; i < (size_t)width * height; i++)
		mean[i * 3] = film[i].r * inv, mean[i * 3 + 1] = film[i].g * inv, mean[i * 3 + 2] = film[i].b * inv;
	rtb_img::writeHDR(filename, (int)width, (int)height, mean.data());
}

// the call Renderer.h:897 makes
inline int stbi_write_png(const char* filename, int w, int h, int comp, const void* data, int stride_in_bytes)
{
	return rtb_img::writePNG(filename, w, h, comp, (const unsigned char*)data, stride_in_bytes) ? 1 : 0;
}
