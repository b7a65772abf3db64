// TEST INFRASTRUCTURE (oracle/): C exports around the UNMODIFIED reference renderer.
//
// This translation unit is compiled by oracle/build_ref.py with
//     g++ -std=c++17 -O2 -ffp-contract=off -shared -fPIC -I oracle/_ref/shadow ...
// where oracle/_ref/shadow/ holds symlinks to /root/reference/RTBase/*.h plus the two stub
// headers of oracle/stubs/.  No reference source is copied or edited; the result
// (oracle/_ref/librtref.so) is the CPU oracle every parity test compares against and the
// `cpu_baseline.kind = "reference"` arm of bench.py.  Only tests/, __graft_entry__.smoke()
// and bench.py's CPU-baseline legs may load it; the product (librtb200.so) never does.
//
// It also instantiates the product's duck-typed flattener (host/rtb_flatten.hpp) on the
// reference's own `Scene`, which is exactly what a drop-in user of the reference does.
#include "GamesEngineeringBase.h"

#include "GEMLoader.h"
#include "Renderer.h"
#include "SceneLoader.h"

#include "rtb_flatten.hpp"

#include <chrono>

namespace
{
// Sampler that replays a fixed list of uniforms (for BSDF::sample / Light::sample vectors).
class ReplaySampler : public Sampler
{
public:
	float v[8];
	int n = 0, pos = 0;
	float next() override
	{
		float r = (pos < n) ? v[pos] : 0.5f;
		pos++;
		return r;
	}
};

struct RefScene
{
	Scene* scene = nullptr;
	RayTracer* rt = nullptr; // lazily created, film owner for AOVs/renders
	GamesEngineeringBase::Window canvas;
	int rtThreads = -1;
};

// `f(sampler.next(), sampler.next())` (Materials.h:129, Lights.h:75,126): the order in which
// the two draws are evaluated is unspecified in C++.  Detect what THIS build does so that the
// exported vectors can be stated in terms of (r1, r2) = the arguments actually received.
bool cosineArgsSwapped()
{
	static int cached = -1;
	if (cached < 0)
	{
		Texture tex;
		tex.loadDefault();
		tex.alpha = NULL;
		DiffuseBSDF bsdf(&tex);
		ShadingData sd(Vec3(0, 0, 0), Vec3(0, 0, 1));
		sd.frame.fromVector(Vec3(0, 0, 1));
		ReplaySampler s;
		s.v[0] = 0.25f, s.v[1] = 0.81f, s.n = 2;
		Colour c;
		float pdf;
		Vec3 wi = bsdf.sample(sd, s, c, pdf);
		// cos(theta) = sqrt(r1): 0.5 if r1 was the first draw, 0.9 if it was the second
		cached = (fabsf(wi.z - 0.5f) < fabsf(wi.z - 0.9f)) ? 0 : 1;
		tex.texels = NULL; // not owned by a `new[]` we want freed twice
	}
	return cached == 1;
}
bool sphereArgsSwapped()
{
	static int cached = -1;
	if (cached < 0)
	{
		BackgroundColour bg(Colour(1, 1, 1));
		ShadingData sd(Vec3(0, 0, 0), Vec3(0, 0, 1));
		ReplaySampler s;
		s.v[0] = 0.25f, s.v[1] = 0.9f, s.n = 2;
		Colour c;
		float pdf;
		Vec3 wi = bg.sample(sd, s, c, pdf);
		// cos(theta) = 1 - 2 r1: 0.5 if r1 was the first draw, -0.8 if the second
		cached = (fabsf(wi.z - 0.5f) < fabsf(wi.z + 0.8f)) ? 0 : 1;
	}
	return cached == 1;
}

void ensureRT(RefScene* h, int threads)
{
	if (h->rt && h->rtThreads == threads) return;
	rtb_ref_thread_override() = threads;
	h->canvas.create((unsigned int)h->scene->camera.width, (unsigned int)h->scene->camera.height, "oracle", 1.0f);
	h->rt = new RayTracer();
	h->rt->init(h->scene, &h->canvas);
	h->rtThreads = threads;
	rtb_ref_thread_override() = 0;
}

Ray makeRay(const rtb_ray& r)
{
	return Ray(Vec3(r.o[0], r.o[1], r.o[2]), Vec3(r.d[0], r.d[1], r.d[2]));
}

ShadingData toShading(Scene* scene, const rtb_shading& s)
{
	ShadingData sd = {};
	sd.x = Vec3(s.x[0], s.x[1], s.x[2]);
	sd.wo = Vec3(s.wo[0], s.wo[1], s.wo[2]);
	sd.sNormal = Vec3(s.s_normal[0], s.s_normal[1], s.s_normal[2]);
	sd.gNormal = Vec3(s.g_normal[0], s.g_normal[1], s.g_normal[2]);
	sd.tu = s.tu;
	sd.tv = s.tv;
	sd.frame.u = Vec3(s.frame_u[0], s.frame_u[1], s.frame_u[2]);
	sd.frame.v = Vec3(s.frame_v[0], s.frame_v[1], s.frame_v[2]);
	sd.frame.w = Vec3(s.frame_w[0], s.frame_w[1], s.frame_w[2]);
	sd.t = s.t;
	sd.bsdf = (s.material >= 0 && (size_t)s.material < scene->materials.size()) ? scene->materials[s.material] : NULL;
	return sd;
}

void fromShading(Scene* scene, const ShadingData& sd, rtb_shading& o)
{
	memset(&o, 0, sizeof(o));
	auto put = [](float* d, const Vec3& v) { d[0] = v.x, d[1] = v.y, d[2] = v.z; };
	put(o.x, sd.x);
	put(o.wo, sd.wo);
	put(o.s_normal, sd.sNormal);
	put(o.g_normal, sd.gNormal);
	o.tu = sd.tu;
	o.tv = sd.tv;
	put(o.frame_u, sd.frame.u);
	put(o.frame_v, sd.frame.v);
	put(o.frame_w, sd.frame.w);
	o.t = sd.t;
	o.material = -1;
	if (sd.bsdf)
	{
		for (size_t i = 0; i < scene->materials.size(); i++)
		{
			if (scene->materials[i] == sd.bsdf)
			{
				o.material = (int32_t)i;
				break;
			}
		}
	}
}
} // namespace

extern "C" {

// loadScene (SceneLoader.h:237).  `dir` is the scene directory (contains scene.json).
void* ref_load_scene(const char* dir)
{
	RefScene* h = new RefScene();
	h->scene = loadScene(std::string(dir));
	return h;
}

int ref_info(void* handle, int* out /* w,h,tris,materials,lights,threads_default */)
{
	RefScene* h = (RefScene*)handle;
	out[0] = (int)h->scene->camera.width;
	out[1] = (int)h->scene->camera.height;
	out[2] = (int)h->scene->triangles.size();
	out[3] = (int)h->scene->materials.size();
	out[4] = (int)h->scene->lights.size();
	SYSTEM_INFO si;
	GetSystemInfo(&si);
	out[5] = (int)si.dwNumberOfProcessors;
	return 0;
}

int ref_arg_order(int* out /* cosineSwapped, sphereSwapped */)
{
	out[0] = cosineArgsSwapped() ? 1 : 0;
	out[1] = sphereArgsSwapped() ? 1 : 0;
	return 0;
}

// Product flattener instantiated on the reference's Scene -> .rtbs file.
int ref_flatten(void* handle, const char* path)
{
	RefScene* h = (RefScene*)handle;
	rtb::FlatScene flat = rtb::flatten(*h->scene);
	return flat.save(path) ? 0 : -1;
}

// renderTile's ray generation + Scene::traverse for every pixel (Renderer.h:806-808).
int ref_primary_hits(void* handle, uint32_t* ids, float* t, rtb_ray* rays)
{
	RefScene* h = (RefScene*)handle;
	Scene* s = h->scene;
	int W = (int)s->camera.width, H = (int)s->camera.height;
	for (int y = 0; y < H; y++)
	{
		for (int x = 0; x < W; x++)
		{
			float px = x + 0.5f, py = y + 0.5f;
			Ray ray = s->camera.generateRay(px, py);
			IntersectionData it = s->traverse(ray);
			size_t i = (size_t)y * W + x;
			bool hit = it.t < FLT_MAX;
			if (ids) ids[i] = hit ? it.ID : 0xFFFFFFFFu;
			if (t) t[i] = it.t;
			if (rays)
			{
				rays[i].o[0] = ray.o.x, rays[i].o[1] = ray.o.y, rays[i].o[2] = ray.o.z;
				rays[i].d[0] = ray.dir.x, rays[i].d[1] = ray.dir.y, rays[i].d[2] = ray.dir.z;
				rays[i].tmax = FLT_MAX;
				rays[i].pad_ = 0;
			}
		}
	}
	return 0;
}

// Scene::traverse (any_hit=0) / BVHNode::traverseVisible(ray, tris, ray.tmax) (any_hit=1).
int ref_trace(void* handle, int any_hit, const rtb_ray* rays, uint64_t n, rtb_hit* hits)
{
	RefScene* h = (RefScene*)handle;
	Scene* s = h->scene;
	for (uint64_t i = 0; i < n; i++)
	{
		Ray ray = makeRay(rays[i]);
		if (any_hit)
		{
			bool vis = s->bvh->traverseVisible(ray, s->triangles, rays[i].tmax);
			hits[i].id = vis ? 0u : 1u;
			hits[i].t = hits[i].alpha = hits[i].beta = hits[i].gamma = 0;
		}
		else
		{
			IntersectionData it = s->traverse(ray);
			if (it.t < FLT_MAX)
			{
				hits[i].id = it.ID, hits[i].t = it.t;
				hits[i].alpha = it.alpha, hits[i].beta = it.beta, hits[i].gamma = it.gamma;
			}
			else
			{
				hits[i].id = 0xFFFFFFFFu, hits[i].t = FLT_MAX;
				hits[i].alpha = hits[i].beta = hits[i].gamma = 0;
			}
		}
	}
	return 0;
}

// Scene::visible (Scene.h:161-169).
int ref_visible(void* handle, const float* p1p2, uint64_t n, uint8_t* out)
{
	RefScene* h = (RefScene*)handle;
	for (uint64_t i = 0; i < n; i++)
	{
		const float* p = p1p2 + i * 6;
		out[i] = h->scene->visible(Vec3(p[0], p[1], p[2]), Vec3(p[3], p[4], p[5])) ? 1 : 0;
	}
	return 0;
}

// Scene::calculateShadingData (Scene.h:174-203).
int ref_shading_data(void* handle, const rtb_ray* rays, const rtb_hit* hits, uint64_t n, rtb_shading* out)
{
	RefScene* h = (RefScene*)handle;
	for (uint64_t i = 0; i < n; i++)
	{
		Ray ray = makeRay(rays[i]);
		IntersectionData it;
		it.ID = hits[i].id, it.t = hits[i].t;
		it.alpha = hits[i].alpha, it.beta = hits[i].beta, it.gamma = hits[i].gamma;
		ShadingData sd = h->scene->calculateShadingData(it, ray);
		fromShading(h->scene, sd, out[i]);
	}
	return 0;
}

// BSDF::evaluate / PDF / sample.  u = n*3 uniforms (r1, r2 of the hemisphere sample as the
// callee receives them; u[2] = glass reflect/refract draw).
int ref_eval_bsdf(void* handle, const rtb_shading* sds, const float* wi, const float* u, uint64_t n, float* eval,
                  float* pdf, float* s_wi, float* s_f, float* s_pdf)
{
	RefScene* h = (RefScene*)handle;
	bool swapped = cosineArgsSwapped();
	for (uint64_t i = 0; i < n; i++)
	{
		ShadingData sd = toShading(h->scene, sds[i]);
		if (!sd.bsdf) return -1;
		Vec3 w(wi[i * 3], wi[i * 3 + 1], wi[i * 3 + 2]);
		if (eval)
		{
			Colour c = sd.bsdf->evaluate(sd, w);
			eval[i * 3] = c.r, eval[i * 3 + 1] = c.g, eval[i * 3 + 2] = c.b;
		}
		if (pdf) pdf[i] = sd.bsdf->PDF(sd, w);
		if (s_wi || s_f || s_pdf)
		{
			ReplaySampler s;
			if (sd.bsdf->isPureSpecular())
			{
				s.v[0] = u[i * 3 + 2], s.n = 1; // glass: one draw; mirror: none
			}
			else
			{
				s.v[0] = swapped ? u[i * 3 + 1] : u[i * 3];
				s.v[1] = swapped ? u[i * 3] : u[i * 3 + 1];
				s.n = 2;
			}
			Colour c;
			float p = 0;
			Vec3 d = sd.bsdf->sample(sd, s, c, p);
			if (s_wi) s_wi[i * 3] = d.x, s_wi[i * 3 + 1] = d.y, s_wi[i * 3 + 2] = d.z;
			if (s_f) s_f[i * 3] = c.r, s_f[i * 3 + 1] = c.g, s_f[i * 3 + 2] = c.b;
			if (s_pdf) s_pdf[i] = p;
		}
	}
	return 0;
}

// Light::sample (u = n*2 uniforms r1,r2) and Light::evaluate(wi).
int ref_eval_light(void* handle, const int32_t* light, const float* wi, const float* u, uint64_t n, float* p_or_wi,
                   float* emitted, float* pdf, float* eval)
{
	RefScene* h = (RefScene*)handle;
	bool swapped = sphereArgsSwapped();
	ShadingData sd(Vec3(0, 0, 0), Vec3(0, 0, 1));
	for (uint64_t i = 0; i < n; i++)
	{
		if (light[i] < 0 || (size_t)light[i] >= h->scene->lights.size()) return -1;
		Light* L = h->scene->lights[light[i]];
		if (p_or_wi || emitted || pdf)
		{
			ReplaySampler s;
			bool sw = swapped && !L->isArea(); // Triangle::sample draws in two statements
			s.v[0] = sw ? u[i * 2 + 1] : u[i * 2];
			s.v[1] = sw ? u[i * 2] : u[i * 2 + 1];
			s.n = 2;
			Colour c;
			float p = 0;
			Vec3 d = L->sample(sd, s, c, p);
			if (p_or_wi) p_or_wi[i * 3] = d.x, p_or_wi[i * 3 + 1] = d.y, p_or_wi[i * 3 + 2] = d.z;
			if (emitted) emitted[i * 3] = c.r, emitted[i * 3 + 1] = c.g, emitted[i * 3 + 2] = c.b;
			if (pdf) pdf[i] = p;
		}
		if (eval)
		{
			Colour c = L->evaluate(Vec3(wi[i * 3], wi[i * 3 + 1], wi[i * 3 + 2]));
			eval[i * 3] = c.r, eval[i * 3 + 1] = c.g, eval[i * 3 + 2] = c.b;
		}
	}
	return 0;
}

// Scene::background->evaluate(dir) for n directions (Renderer.h:390).
int ref_eval_background(void* handle, const float* wi, uint64_t n, float* out)
{
	RefScene* h = (RefScene*)handle;
	for (uint64_t i = 0; i < n; i++)
	{
		Colour c = h->scene->background->evaluate(Vec3(wi[i * 3], wi[i * 3 + 1], wi[i * 3 + 2]));
		out[i * 3] = c.r, out[i * 3 + 1] = c.g, out[i * 3 + 2] = c.b;
	}
	return 0;
}

// `spp` calls of RayTracer::render() (Renderer.h:876-885) with `threads` workers
// (0 = all hardware threads).  film_sum receives Film::film (running sums, W*H*3 floats);
// *seconds the wall time of the render() calls only.  `fresh` != 0 clears the film first.
int ref_render(void* handle, int spp, int threads, int fresh, float* film_sum, double* seconds)
{
	RefScene* h = (RefScene*)handle;
	ensureRT(h, threads);
	if (fresh)
	{
		h->rt->clear();
		// every run starts from the reference's initial sampler state (seed 1, Sampling.h:18)
		for (int i = 0; i < h->rt->numProcs; i++) h->rt->samplers[i] = MTRandom();
	}
	auto t0 = std::chrono::steady_clock::now();
	for (int i = 0; i < spp; i++) h->rt->render();
	auto t1 = std::chrono::steady_clock::now();
	if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
	if (film_sum)
	{
		size_t n = (size_t)h->rt->film->width * h->rt->film->height;
		for (size_t i = 0; i < n; i++)
		{
			film_sum[i * 3] = h->rt->film->film[i].r;
			film_sum[i * 3 + 1] = h->rt->film->film[i].g;
			film_sum[i * 3 + 2] = h->rt->film->film[i].b;
		}
	}
	return h->rt->getSPP();
}

// One RayTracer::render() with the adaptive path switched on: film->incrementSPP() +
// adaptiveRender() (Renderer.h:679-749, :876-881).  tile_var = tileVariances; tile_samples = the counts
// sampleTileWithWeight derives (and prints) from tileWeights (:649-653).
int ref_render_adaptive(void* handle, int threads, int fresh, float* film_sum, float* tile_var, int* tile_samples, double* seconds)
{
	RefScene* h = (RefScene*)handle;
	ensureRT(h, threads);
	if (fresh)
	{
		h->rt->clear();
		for (int i = 0; i < h->rt->numProcs; i++) h->rt->samplers[i] = MTRandom();
	}
	std::streambuf* old = std::cout.rdbuf(nullptr); // sampleTileWithWeight prints one line per tile
	auto t0 = std::chrono::steady_clock::now();
	h->rt->film->incrementSPP();
	h->rt->adaptiveRender();
	auto t1 = std::chrono::steady_clock::now();
	std::cout.rdbuf(old);
	if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
	for (int i = 0; i < h->rt->totalTiles; i++)
	{
		if (tile_var) tile_var[i] = h->rt->tileVariances[i];
		if (tile_samples)
		{
			float weight = h->rt->tileWeights[i];
			weight = sqrt(weight);
			int sample = (int)(weight * MAX_SAMPLES);
			tile_samples[i] = sample > MIN_SAMPLES ? sample : MIN_SAMPLES;
		}
	}
	if (film_sum)
	{
		size_t n = (size_t)h->rt->film->width * h->rt->film->height;
		for (size_t i = 0; i < n; i++)
		{
			film_sum[i * 3] = h->rt->film->film[i].r;
			film_sum[i * 3 + 1] = h->rt->film->film[i].g;
			film_sum[i * 3 + 2] = h->rt->film->film[i].b;
		}
	}
	return h->rt->getSPP();
}

// `passes` x { film->incrementSPP(); lightTracer(); } (Renderer.h:220-231, 876-884: the commented-out
// alternative of render()).  Single-threaded in the reference (samplers[0]).
int ref_render_light(void* handle, int passes, int fresh, float* film_sum, double* seconds)
{
	RefScene* h = (RefScene*)handle;
	ensureRT(h, 1);
	if (fresh)
	{
		h->rt->clear();
		for (int i = 0; i < h->rt->numProcs; i++) h->rt->samplers[i] = MTRandom();
	}
	auto t0 = std::chrono::steady_clock::now();
	for (int i = 0; i < passes; i++)
	{
		h->rt->film->incrementSPP();
		h->rt->lightTracer();
	}
	auto t1 = std::chrono::steady_clock::now();
	if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
	size_t n = (size_t)h->rt->film->width * h->rt->film->height;
	for (size_t i = 0; i < n; i++)
	{
		film_sum[i * 3] = h->rt->film->film[i].r;
		film_sum[i * 3 + 1] = h->rt->film->film[i].g;
		film_sum[i * 3 + 2] = h->rt->film->film[i].b;
	}
	return h->rt->getSPP();
}

// `passes` x { film->incrementSPP(); instantRadiosity(); } (Renderer.h:102-123, render()'s other commented-out
// alternative, :884).  *vpls = total number of VPLs stored over the passes.
int ref_render_ir(void* handle, int passes, int threads, int fresh, float* film_sum, double* seconds, uint64_t* vpls)
{
	RefScene* h = (RefScene*)handle;
	ensureRT(h, threads);
	if (fresh)
	{
		h->rt->clear();
		for (int i = 0; i < h->rt->numProcs; i++) h->rt->samplers[i] = MTRandom();
	}
	uint64_t nv = 0;
	auto t0 = std::chrono::steady_clock::now();
	for (int i = 0; i < passes; i++)
	{
		h->rt->film->incrementSPP();
		h->rt->instantRadiosity();
		nv += h->rt->vpls.size();
	}
	auto t1 = std::chrono::steady_clock::now();
	if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
	if (vpls) *vpls = nv;
	size_t n = (size_t)h->rt->film->width * h->rt->film->height;
	for (size_t i = 0; i < n; i++)
	{
		film_sum[i * 3] = h->rt->film->film[i].r;
		film_sum[i * 3 + 1] = h->rt->film->film[i].g;
		film_sum[i * 3 + 2] = h->rt->film->film[i].b;
	}
	return h->rt->getSPP();
}

// stbi_loadf (Texture::load's .hdr branch, Imaging.h:37) on one file.  out may be NULL to query the size.
int ref_decode_hdr(const char* path, int* w, int* h, int* channels, float* out, uint64_t cap_floats)
{
	float* d = stbi_loadf(path, w, h, channels, 0);
	if (!d) return -1;
	uint64_t n = (uint64_t)(*w) * (*h) * (*channels);
	int rc = 0;
	if (out)
	{
		if (n > cap_floats) rc = -2;
		else memcpy(out, d, n * sizeof(float));
	}
	stbi_image_free(d);
	return rc;
}

// Camera::projectionMatrix (16), cameraToView (16), viewDirection (3), Afilm (1): Scene.h:11-41.
int ref_camera_ext(void* handle, float* out36)
{
	RefScene* h = (RefScene*)handle;
	Camera& c = h->scene->camera;
	for (int i = 0; i < 16; i++) out36[i] = c.projectionMatrix.m[i], out36[16 + i] = c.cameraToView.m[i];
	out36[32] = c.viewDirection.x, out36[33] = c.viewDirection.y, out36[34] = c.viewDirection.z;
	out36[35] = c.Afilm;
	return 0;
}

// stbi_load (the decoder behind Texture::load, Imaging.h:51) on one file: the golden for the product's
// own PNG / JPEG decoders.  out may be NULL to query the size.
int ref_decode_image(const char* path, int* w, int* h, int* channels, unsigned char* out, uint64_t cap)
{
	unsigned char* d = stbi_load(path, w, h, channels, 0);
	if (!d) return -1;
	uint64_t n = (uint64_t)(*w) * (*h) * (*channels);
	int rc = 0;
	if (out)
	{
		if (n > cap) rc = -2;
		else memcpy(out, d, n);
	}
	stbi_image_free(d);
	return rc;
}

// RayTracer::albedo (kind 0, Renderer.h:558-571) / viewNormals (kind 1, :572-581) /
// direct() with a fresh seed-1 MTRandom (kind 2, :393-407) at pixel centres.
int ref_aov(void* handle, int kind, float* out)
{
	RefScene* h = (RefScene*)handle;
	ensureRT(h, 1);
	Scene* s = h->scene;
	int W = (int)s->camera.width, H = (int)s->camera.height;
	MTRandom sampler;
	for (int y = 0; y < H; y++)
	{
		for (int x = 0; x < W; x++)
		{
			Ray ray = s->camera.generateRay(x + 0.5f, y + 0.5f);
			Colour c;
			if (kind == 0) c = h->rt->albedo(ray);
			else if (kind == 1) c = h->rt->viewNormals(ray);
			else c = h->rt->direct(ray, sampler);
			size_t i = ((size_t)y * W + x) * 3;
			out[i] = c.r, out[i + 1] = c.g, out[i + 2] = c.b;
		}
	}
	return 0;
}

// Film::tonemap on an arbitrary sum film (Imaging.h:233-242).
int ref_tonemap(void* handle, const float* film_sum, int spp, float exposure, uint8_t* rgb8)
{
	RefScene* h = (RefScene*)handle;
	ensureRT(h, 1);
	Film* f = h->rt->film;
	size_t n = (size_t)f->width * f->height;
	std::vector<Colour> saved(f->film, f->film + n);
	int savedSpp = f->SPP;
	for (size_t i = 0; i < n; i++) f->film[i] = Colour(film_sum[i * 3], film_sum[i * 3 + 1], film_sum[i * 3 + 2]);
	f->SPP = spp;
	for (unsigned int y = 0; y < f->height; y++)
		for (unsigned int x = 0; x < f->width; x++)
		{
			size_t i = ((size_t)y * f->width + x) * 3;
			f->tonemap(x, y, rgb8[i], rgb8[i + 1], rgb8[i + 2], exposure);
		}
	std::copy(saved.begin(), saved.end(), f->film);
	f->SPP = savedSpp;
	return 0;
}

// Film::splat with a GaussianFilter(radius, alpha) of a per-pixel colour image (one sample
// at every pixel centre): what switching Renderer.h:50 to :51 does to the film.
int ref_gaussian_splat(int width, int height, float radius, float alpha, const float* colours, float* film_sum)
{
	Film film;
	film.init(width, height, new GaussianFilter(radius, alpha));
	for (int y = 0; y < height; y++)
		for (int x = 0; x < width; x++)
		{
			size_t i = ((size_t)y * width + x) * 3;
			film.splat(x + 0.5f, y + 0.5f, Colour(colours[i], colours[i + 1], colours[i + 2]));
		}
	for (size_t i = 0; i < (size_t)width * height; i++)
	{
		film_sum[i * 3] = film.film[i].r, film_sum[i * 3 + 1] = film.film[i].g, film_sum[i * 3 + 2] = film.film[i].b;
	}
	return 0;
}

} // extern "C"
