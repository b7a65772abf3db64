// rtb_host.cpp — C entry points of the stand-alone host layer (librtb200_host.so):
// scene loading / BVH building / flattening without any reference code, used by bench.py,
// the examples and the parity tests of the loader (tests/test_host_cpu.py).
// Build: g++ -std=c++17 -O2 -ffp-contract=off -shared -fPIC rtb_host.cpp -lz
#include "rtb_standalone.hpp"

#include "rtb_flatten.hpp"

#include <chrono>
#include <random>

namespace
{
thread_local std::string g_error;
}

extern "C" {

const char* rtbh_last_error() { return g_error.c_str(); }

// loadScene(dir) -> flatten -> <out>.rtbs.  0 on success.
int rtbh_load_scene_rtbs(const char* dir, const char* out_path)
{
	try
	{
		Scene* scene = loadScene(dir);
		rtb::FlatScene flat = rtb::flatten(*scene);
		if (!flat.save(out_path))
		{
			g_error = std::string("cannot write ") + out_path;
			return -2;
		}
		return 0;
	}
	catch (const std::exception& e)
	{
		g_error = e.what();
		return -1;
	}
}

int rtbh_matrix_invert(const float* m, float* out)
{
	Matrix a;
	memcpy(a.m, m, sizeof(a.m));
	Matrix b = a.invert();
	memcpy(out, b.m, sizeof(b.m));
	return 0;
}

// from, to, up, fov, width, height -> inverse projection, camera-to-world, origin (35 floats)
int rtbh_camera(const float* from, const float* to, const float* up, float fov, int width, int height, float* out)
{
	Matrix P = Matrix::perspective(0.001f, 10000.0f, (float)width / (float)height, fov);
	Matrix V = Matrix::lookAt(Vec3(from[0], from[1], from[2]), Vec3(to[0], to[1], to[2]), Vec3(up[0], up[1], up[2])).invert();
	Camera c;
	c.init(P, width, height);
	c.updateView(V);
	memcpy(out, c.inverseProjectionMatrix.m, 64);
	memcpy(out + 16, c.camera.m, 64);
	out[32] = c.origin.x, out[33] = c.origin.y, out[34] = c.origin.z;
	return 0;
}

// Synthetic random-triangle soup (SURVEY 8d cfg 5): N triangles, std::mt19937(0xB200 + log2 N),
// centre ~ U[0,1]^3, vertices = centre + U[-s/2, s/2]^3 with s = N^(-1/3), vertex normals =
// geometric normal, one DiffuseBSDF (1x1 texture 0.7), BackgroundColour(1,1,1), camera at
// (0.5,0.5,3) looking at (0.5,0.5,0.5), fov 25, width x height.  Writes <out>.rtbs.
int rtbh_build_soup(uint32_t n_tris, int width, int height, const char* out_path, double* build_seconds)
{
	try
	{
		uint32_t lg = 0;
		while ((1u << (lg + 1)) <= n_tris) lg++;
		std::mt19937 gen(0xB200u + lg);
		std::uniform_real_distribution<float> U(0.0f, 1.0f);
		float s = powf((float)n_tris, -1.0f / 3.0f);
		std::vector<Triangle> tris;
		tris.reserve(n_tris);
		for (uint32_t i = 0; i < n_tris; i++)
		{
			Vec3 c(U(gen), U(gen), U(gen));
			Vertex v[3];
			for (int k = 0; k < 3; k++)
			{
				v[k].p = Vec3(c.x + (U(gen) - 0.5f) * s, c.y + (U(gen) - 0.5f) * s, c.z + (U(gen) - 0.5f) * s);
				v[k].u = v[k].v = 0.0f;
			}
			Vec3 n = (v[2].p - v[1].p).cross(v[0].p - v[2].p);
			float len = n.length();
			if (!(len > 0)) continue;
			n = n.normalize();
			for (int k = 0; k < 3; k++) v[k].normal = n;
			Triangle t;
			t.init(v[0], v[1], v[2], 0);
			if (t.area > 0) tris.push_back(t);
		}
		Texture* tex = new Texture();
		tex->width = tex->height = 1, tex->channels = 3;
		tex->texels = new Colour[1];
		tex->texels[0] = Colour(0.7f, 0.7f, 0.7f);
		DiffuseBSDF* mat = new DiffuseBSDF();
		mat->albedo = tex;
		std::vector<BSDF*> mats{mat};
		Scene* scene = new Scene();
		Matrix P = Matrix::perspective(0.001f, 10000.0f, (float)width / (float)height, 25.0f);
		Matrix V = Matrix::lookAt(Vec3(0.5f, 0.5f, 3.0f), Vec3(0.5f, 0.5f, 0.5f), Vec3(0, 1, 0)).invert();
		scene->camera.init(P, width, height);
		scene->camera.updateView(V);
		scene->init(tris, mats, new BackgroundColour(Colour(1.0f, 1.0f, 1.0f)));
		tris.clear();
		tris.shrink_to_fit();
		auto t0 = std::chrono::steady_clock::now();
		scene->build();
		if (build_seconds) *build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		rtb::FlatScene flat = rtb::flatten(*scene);
		if (!flat.save(out_path))
		{
			g_error = std::string("cannot write ") + out_path;
			return -2;
		}
		return 0;
	}
	catch (const std::exception& e)
	{
		g_error = e.what();
		return -1;
	}
}

// Decode a PNG / JPEG file with the loader's own decoders (what Texture::load does before the float
// conversion).  out may be NULL to query the size.  Returns 0, -1 (cannot read / decode), -2 (cap too small).
int rtbh_decode_image(const char* path, int* w, int* h, int* channels, unsigned char* out, uint64_t cap)
{
	std::vector<unsigned char> file = rtb_img::readFile(path);
	if (file.empty()) return -1;
	std::vector<unsigned char> px;
	bool isJPEG = file.size() > 2 && file[0] == 0xFF && file[1] == 0xD8;
	bool ok = isJPEG ? rtb_img::decodeJPEG(file, *w, *h, *channels, px) : rtb_img::decodePNG(file, *w, *h, *channels, px);
	if (!ok) return -1;
	if (out)
	{
		if (px.size() > cap) return -2;
		memcpy(out, px.data(), px.size());
	}
	return 0;
}

// Radiance .hdr with the loader's own decoder -> float RGB (what Texture::load stores).  out may be NULL.
int rtbh_decode_hdr(const char* path, int* w, int* h, float* out, uint64_t cap_floats)
{
	std::vector<unsigned char> file = rtb_img::readFile(path);
	if (file.empty()) return -1;
	std::vector<float> px;
	if (!rtb_img::decodeHDR(file, *w, *h, px)) return -1;
	if (out)
	{
		if (px.size() > cap_floats) return -2;
		memcpy(out, px.data(), px.size() * sizeof(float));
	}
	return 0;
}

// The stand-alone program's image writers (Film::save's Radiance .hdr, savePNG's 8-bit PNG), callable for tests.
int rtbh_write_hdr(const char* path, int w, int h, const float* rgb)
{
	return rtb_img::writeHDR(path, w, h, rgb) ? 0 : -1;
}
int rtbh_write_png(const char* path, int w, int h, int channels, const unsigned char* data)
{
	return rtb_img::writePNG(path, w, h, channels, data, w * channels) ? 0 : -1;
}

// Test hook: 1 if the builder's parallel sort reproduces std::sort's permutation (ties included) on `keys`.
int rtbh_sort_selftest(const float* keys, uint32_t n, int par)
{
	return BVHNode::sortSelfTest(keys, n, par) ? 1 : 0;
}
} // extern "C"
