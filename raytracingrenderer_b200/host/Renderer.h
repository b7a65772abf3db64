// Renderer.h — drop-in replacement for RTBase/Renderer.h: the same `RayTracer` class surface
// (RTBase/Renderer.h:30-67, 876-898: scene, canvas, film, numProcs; init, clear, render, getSPP,
// saveHDR, savePNG), with everything below render() running on a B200 through the C ABI of
// librtb200.so (include/rtb.h).
//
// Use: put this file where RTBase/Renderer.h was (RTBase/SceneLoader.h and Main.cpp include it
// by that name), keep the rest of RTBase/ untouched, add <repo>/include and
// <repo>/raytracingrenderer_b200/host to the include path, link -lrtb200.  See INTEGRATION.md.
//
// Like the original it expects GamesEngineeringBase.h (Window) to be available; headless builds
// may pass a null canvas.  Not part of the original surface (additions, all optional):
//   render(n)            n samples per pixel in one call (render() == render(1))
//   setPresentEveryFrame the original tonemaps and draws the film after every sample
//                        (presentFilmToCanvas, Renderer.h:69-80); switch it off for batch work
//   params()             the rtb_params the GPU uses (defaults = the reference's constants)
//   syncFilm()           copy the GPU film sums into film->film (done by saveHDR automatically)
//   renderAdaptive()     render() with the original's commented-out adaptiveRender() switched on (:880)
#pragma once

#include "Core.h"
#include "Sampling.h"
#include "Geometry.h"
#include "Imaging.h"
#include "Materials.h"
#include "Lights.h"
#include "Scene.h"
#include "GamesEngineeringBase.h"

#include "rtb.h"
#include "rtb_flatten.hpp"

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

const int TILE_SIZE = 32; // RTBase/Renderer.h:18 (the tile partition of rtb_params uses the same size)
const int MAX_DEPTH = 4;  // RTBase/Renderer.h:20 -> rtb_params.max_depth
const int MAX_SAMPLES = 10240; // RTBase/Renderer.h:21-23 -> the arguments of rtb_render_adaptive
const int MIN_SAMPLES = 1;
const int INIT_SAMPLES = 2;
const int MAX_VPL = 50; // RTBase/Renderer.h:24 -> n_paths of rtb_render_ir

class RayTracer
{
public:
	Scene* scene = NULL;
	GamesEngineeringBase::Window* canvas = NULL;
	Film* film = NULL;
	MTRandom* samplers = NULL; // kept for source compatibility; the GPU uses a counter-based RNG
	int numProcs = 0;          // number of GPUs driving this RayTracer (RTBase/Renderer.h:52-55: every processor)
	int tilesNumX = 0, tilesNumY = 0, totalTiles = 0;  // 32x32 tiles (RTBase/Renderer.h:40, 57-61)
	std::vector<float> tileVariances, tileWeights;     // filled by adaptiveRender()
	std::vector<uint32_t> tileSamples;                 // the counts sampleTileWithWeight derives (:649-653)

	// Like the original (numProcs = sysInfo.dwNumberOfProcessors, RTBase/Renderer.h:52-55) this uses every
	// processor of the machine: all visible GPUs as one device group (RTB_NUM_GPUS=n caps it).
	void init(Scene* _scene, GamesEngineeringBase::Window* _canvas)
	{
		int n = rtb_device_count();
		if (const char* e = getenv("RTB_NUM_GPUS"))
		{
			int v = atoi(e);
			if (v >= 1 && v < n) n = v;
		}
		init(_scene, _canvas, -1, n);
	}
	// one chosen device
	void init(Scene* _scene, GamesEngineeringBase::Window* _canvas, int device) { init(_scene, _canvas, device, 1); }
	void init(Scene* _scene, GamesEngineeringBase::Window* _canvas, int device, int nDevices)
	{
		scene = _scene;
		canvas = _canvas;
		film = new Film();
		film->init((unsigned int)scene->camera.width, (unsigned int)scene->camera.height, new BoxFilter());
		numProcs = nDevices > 1 ? nDevices : 1;
		samplers = new MTRandom[numProcs];
		tilesNumX = (film->width + TILE_SIZE - 1) / TILE_SIZE;
		tilesNumY = (film->height + TILE_SIZE - 1) / TILE_SIZE;
		totalTiles = tilesNumX * tilesNumY;
		tileVariances = std::vector<float>(totalTiles, 0.0f);
		tileWeights = std::vector<float>(totalTiles, 0.0f);
		tileSamples = std::vector<uint32_t>(totalTiles, 0u);
		if (device >= 0) check(rtb_create(device, &ctx), "rtb_create");
		else check(rtb_create_multi(NULL, numProcs, &ctx), "rtb_create_multi");
		numProcs = rtb_group_size(ctx);
		rtb_default_params(&prm);
		prm.max_depth = MAX_DEPTH;
		prm.epsilon = EPSILON;
		check(rtb_set_params(ctx, &prm), "rtb_set_params");
		upload();
		clear();
	}
	// Scene geometry / materials changed: flatten and upload again.
	void upload()
	{
		flush();
		rtb::FlatScene flat = rtb::flatten(*scene);
		rtb_scene_desc d = flat.desc();
		check(rtb_upload_scene(ctx, &d), "rtb_upload_scene");
		camera = flat.camera;
	}
	void clear()
	{
		pendingCount = 0; // samples nobody looked at are not traced
		film->clear();
		// the camera may have moved (RTCamera::updateCamera -> Camera::updateView)
		rtb_camera now = currentCamera();
		if (memcmp(&now, &camera, sizeof(now)) != 0)
		{
			camera = now;
			check(rtb_update_camera(ctx, &camera), "rtb_update_camera");
		}
		check(rtb_clear(ctx), "rtb_clear");
		filmOnHost = true;
	}
	// RTBase calls render() once per sample (Main.cpp:114).  A GPU render has a fixed cost — the pool fills and drains
	// once per rtb_render call, ~1 ms (profiles/r02_fixed_cost.txt) — so consecutive samples are collected and traced
	// in ONE rtb_render call: when somebody needs the film (present / save / getFilm / the other estimators), when the
	// parameters or the scene change, or after renderBatch samples.  Sample s of a pixel is the same sample whichever call
	// traces it (counter RNG keyed (pixel, sample); integer film sums), so the film is bit-identical to per-sample calls.
	// With a canvas and presentEveryFrame (the interactive shape of Main.cpp) every render() still shows its frame.
	void render() { render(1); }
	void render(int n)
	{
		if (n <= 0) return;
		if (pendingCount == 0) pendingBegin = (uint32_t)film->SPP;
		for (int i = 0; i < n; i++) film->incrementSPP();
		pendingCount += (uint32_t)n;
		filmOnHost = false;
		if (presentEveryFrame && canvas) presentFilmToCanvas();
		else if (pendingCount >= renderBatch) flush();
	}
	// Trace the samples render() has collected so far (rtb_render returns without waiting for the GPU).
	void flush()
	{
		if (pendingCount == 0) return;
		uint32_t begin = pendingBegin, count = pendingCount;
		pendingCount = 0;
		check(rtb_render(ctx, begin, count), "rtb_render");
	}
	void setRenderBatch(unsigned int samples) { renderBatch = samples ? samples : 1u; }
	// RTBase/Renderer.h:679-749.  Like the original it is meant to be called from render() in place of
	// pathTracerTileBased() (:880), after film->incrementSPP(); renderAdaptive() does both.
	void adaptiveRender()
	{
		flush();
		check(rtb_render_adaptive(ctx, INIT_SAMPLES, MIN_SAMPLES, MAX_SAMPLES, tileSamples.data(), tileVariances.data()),
		      "rtb_render_adaptive");
		check(rtb_set_spp(ctx, (uint32_t)film->SPP), "rtb_set_spp");
		float totalVariance = 0.0f;
		for (float v : tileVariances) totalVariance += v;
		for (int i = 0; i < totalTiles; i++) tileWeights[i] = (totalVariance > 0.0f) ? tileVariances[i] / totalVariance : 0.0f;
		filmOnHost = false;
		if (presentEveryFrame && canvas) presentFilmToCanvas();
	}
	// RTBase/Renderer.h:220-231 (render()'s commented-out alternative at :883): one light-tracing pass.
	void lightTracer()
	{
		flush();
		uint32_t pass = film->SPP > 0 ? (uint32_t)film->SPP - 1u : 0u; // render() has incremented SPP already
		check(rtb_render_light(ctx, pass, 1), "rtb_render_light");
		check(rtb_set_spp(ctx, (uint32_t)film->SPP), "rtb_set_spp");
		filmOnHost = false;
		if (presentEveryFrame && canvas) presentFilmToCanvas();
	}
	// RTBase/Renderer.h:102-123 (render()'s commented-out alternative at :884): one instant-radiosity pass.
	void instantRadiosity()
	{
		flush();
		uint32_t pass = film->SPP > 0 ? (uint32_t)film->SPP - 1u : 0u;
		check(rtb_render_ir(ctx, pass, 1, MAX_VPL), "rtb_render_ir");
		check(rtb_set_spp(ctx, (uint32_t)film->SPP), "rtb_set_spp");
		filmOnHost = false;
		if (presentEveryFrame && canvas) presentFilmToCanvas();
	}
	void renderInstantRadiosity(int n = 1)
	{
		for (int i = 0; i < n; i++)
		{
			film->incrementSPP();
			instantRadiosity();
		}
	}
	void renderLight(int n = 1)
	{
		for (int i = 0; i < n; i++)
		{
			film->incrementSPP();
			lightTracer();
		}
	}
	void renderAdaptive()
	{
		film->incrementSPP();
		adaptiveRender();
	}
	// RTBase/Renderer.h:750-792 runs OIDN on film->film in place.  Here the filter is whatever the caller
	// installs (e.g. a lambda around oidn::FilterRef): it sees the film sums on the host and edits them in
	// place; they are written back to the GPU.  Without a filter denoise() is a no-op.
	typedef void (*DenoiseFn)(Colour* film, unsigned int width, unsigned int height, void* user);
	void setDenoiser(DenoiseFn fn, void* user = NULL) { denoiseFn = fn, denoiseUser = user; }
	void denoise()
	{
		if (!denoiseFn) return;
		syncFilm();
		denoiseFn(film->film, film->width, film->height, denoiseUser);
		check(rtb_write_film(ctx, (const float*)film->film), "rtb_write_film");
	}
	void presentFilmToCanvas()
	{
		if (!canvas) return;
		flush();
		std::vector<uint8_t> rgb((size_t)film->width * film->height * 3);
		check(rtb_tonemap(ctx, rgb.data(), 1.0f), "rtb_tonemap");
		for (unsigned int y = 0; y < film->height; y++)
			for (unsigned int x = 0; x < film->width; x++)
			{
				const uint8_t* p = &rgb[((size_t)y * film->width + x) * 3];
				canvas->draw(x, y, p[0], p[1], p[2]);
			}
	}
	int getSPP() { return film->SPP; }
	void syncFilm()
	{
		if (filmOnHost) return;
		flush();
		uint32_t spp = 0;
		check(rtb_read_film(ctx, (float*)film->film, &spp), "rtb_read_film"); // Colour = 3 packed floats
		filmOnHost = true;
	}
	void saveHDR(std::string filename)
	{
		syncFilm();
		film->save(filename);
	}
	void savePNG(std::string filename)
	{
		if (!canvas) return;
		presentFilmToCanvas();
		stbi_write_png(filename.c_str(), canvas->getWidth(), canvas->getHeight(), 3, canvas->getBackBuffer(), canvas->getWidth() * 3);
	}
	void setPresentEveryFrame(bool on) { presentEveryFrame = on; }
	DenoiseFn denoiseFn = NULL;
	void* denoiseUser = NULL;
	rtb_params& params() { return prm; }
	void applyParams()
	{
		flush(); // the collected samples belong to the old parameters
		check(rtb_set_params(ctx, &prm), "rtb_set_params");
	}
	rtb_ctx* context()
	{
		flush(); // whoever talks to the C ABI directly sees every sample render() was asked for
		return ctx;
	}
	~RayTracer()
	{
		if (ctx) rtb_destroy(ctx);
	}

private:
	rtb_ctx* ctx = NULL;
	rtb_params prm;
	rtb_camera camera;
	bool presentEveryFrame = true;
	bool filmOnHost = true;
	uint32_t pendingBegin = 0, pendingCount = 0; // samples asked for by render() and not traced yet
	uint32_t renderBatch = 256;                  // at most this many are collected (setRenderBatch)

	rtb_camera currentCamera()
	{
		rtb_camera c;
		memset(&c, 0, sizeof(c));
		memcpy(c.inv_proj, scene->camera.inverseProjectionMatrix.m, 16 * sizeof(float));
		memcpy(c.cam_to_world, scene->camera.camera.m, 16 * sizeof(float));
		c.origin[0] = scene->camera.origin.x, c.origin[1] = scene->camera.origin.y, c.origin[2] = scene->camera.origin.z;
		c.width = scene->camera.width, c.height = scene->camera.height;
		return c;
	}
	void check(int rc, const char* what)
	{
		if (rc == RTB_OK) return;
		// the reference's own error handling is "print and exit" (GEMLoader.h:349-354)
		fprintf(stderr, "%s failed (%d): %s\n", what, rc, rtb_last_error(ctx));
		exit(1);
	}
};
