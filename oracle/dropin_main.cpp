// TEST INFRASTRUCTURE (oracle/): proves the drop-in boundary.  This is the reference's own
// host program shape (RTBase/Main.cpp:67-71,114,135: loadScene -> RayTracer::init ->
// render() x SPP -> saveHDR) compiled against the UNMODIFIED RTBase headers — except that
// "Renderer.h" resolves to raytracingrenderer_b200/host/Renderer.h, so RayTracer renders on the
// GPU through librtb200.so.  Built by oracle/build_ref.py into oracle/_ref/dropin_main.
#include "GamesEngineeringBase.h"

#include "GEMLoader.h"
#include "Renderer.h"
#include "SceneLoader.h"

int main(int argc, char** argv)
{
	if (argc < 4)
	{
		fprintf(stderr, "usage: dropin_main <scene dir> <spp> <out.hdr> [raw film out]\n");
		return 2;
	}
	Scene* scene = loadScene(argv[1]);
	GamesEngineeringBase::Window canvas;
	canvas.create((unsigned int)scene->camera.width, (unsigned int)scene->camera.height, "Tracer", 1.0f);
	RayTracer rt;
	rt.init(scene, &canvas);
	int spp = atoi(argv[2]);
	rt.setPresentEveryFrame(false);
	// the reference's loop body: one render() per sample (Main.cpp:114).  The drop-in collects them (here at most 3 at a
	// time, so that 8 samples take two full batches and a partial one); the last two samples come in one call.
	rt.setRenderBatch(3);
	for (int i = 0; i + 2 < spp; i++) rt.render();
	rt.render(spp > 2 ? 2 : spp);
	printf("SPP: %d\n", rt.getSPP());
	rt.saveHDR(argv[3]);
	if (argc > 4)
	{
		rt.syncFilm();
		FILE* f = fopen(argv[4], "wb");
		fwrite(rt.film->film, sizeof(Colour), (size_t)rt.film->width * rt.film->height, f);
		fclose(f);
	}
	// moving the camera and clearing works like in Main.cpp:85-112
	viewcamera.forward();
	rt.clear();
	rt.render(2);
	rt.savePNG(std::string(argv[3]) + ".png");
	return 0;
}
