// rtb_render — headless counterpart of RTBase/Main.cpp (:60-140: loadScene, RayTracer::init, render() in
// a loop, saveHDR / savePNG) built ONLY from this repository: host/standalone/*.h resolve RTBase's header
// names to the product's own scene API and loader, host/Renderer.h is the drop-in RayTracer, and the
// path tracing runs on the GPU through librtb200.so.  No reference code is compiled in.
//
//   rtb_render <scene dir> <spp> <out.hdr> [--adaptive] [--mis] [--light] [--ir] [--png out.png] [--raw film.bin]
#include "GamesEngineeringBase.h"

#include "GEMLoader.h"
#include "Renderer.h"
#include "SceneLoader.h"

#include <chrono>

int main(int argc, char** argv)
{
	if (argc < 4)
	{
		fprintf(stderr, "usage: rtb_render <scene dir> <spp> <out.hdr> [--adaptive] [--mis] [--light] [--ir] [--png out.png] [--raw film.bin]\n");
		return 2;
	}
	bool adaptive = false, mis = false, light = false, ir = false;
	std::string png, raw;
	for (int i = 4; i < argc; i++)
	{
		std::string a = argv[i];
		if (a == "--adaptive") adaptive = true;
		else if (a == "--mis") mis = true;
		else if (a == "--light") light = true;
		else if (a == "--ir") ir = true;
		else if (a == "--png" && i + 1 < argc) png = argv[++i];
		else if (a == "--raw" && i + 1 < argc) raw = argv[++i];
	}
	try
	{
		auto t0 = std::chrono::steady_clock::now();
		Scene* scene = loadScene(argv[1]);
		GamesEngineeringBase::Window canvas;
		canvas.create((unsigned int)scene->camera.width, (unsigned int)scene->camera.height, "Tracer", 1.0f);
		RayTracer rt;
		rt.init(scene, &canvas);
		rt.setPresentEveryFrame(false);
		if (mis)
		{
			rt.params().integrator = RTB_INT_PATH_MIS;
			rt.applyParams();
		}
		auto t1 = std::chrono::steady_clock::now();
		int spp = atoi(argv[2]);
		if (light) rt.renderLight(spp); // lightTracer() per "sample", like Renderer.h:883
		else if (ir) rt.renderInstantRadiosity(spp); // instantRadiosity() per "sample", like Renderer.h:884
		else if (adaptive)
			for (int i = 0; i < spp; i++) rt.renderAdaptive(); // one adaptiveRender() per "sample", like Renderer.h:880
		else
			rt.render(spp);
		rt.syncFilm();
		auto t2 = std::chrono::steady_clock::now();
		printf("%zu triangles, %dx%d, SPP %d: load+build+upload %.3f s, render %.3f s\n", scene->triangles.size(), (int)scene->camera.width,
		       (int)scene->camera.height, rt.getSPP(), std::chrono::duration<double>(t1 - t0).count(),
		       std::chrono::duration<double>(t2 - t1).count());
		rt.saveHDR(argv[3]);
		if (!png.empty()) rt.savePNG(png);
		if (!raw.empty())
		{
			FILE* f = fopen(raw.c_str(), "wb");
			if (!f) return 3;
			fwrite(rt.film->film, sizeof(Colour), (size_t)rt.film->width * rt.film->height, f);
			fclose(f);
		}
	}
	catch (const std::exception& e)
	{
		fprintf(stderr, "rtb_render: %s\n", e.what());
		return 1;
	}
	return 0;
}
