/*
 * rtb.h — C ABI of librtb200.so, the B200 (sm_100a) implementation of RTBase's
 * per-pixel path-tracing loop.
 *
 * The reference (NoSameRain/RayTracingRenderer, directory RTBase/) has no FFI: its seam is
 * the C++ class surface used by RTBase/Main.cpp:67-71,114,124,135
 *     Scene* loadScene(std::string)                       RTBase/SceneLoader.h:237
 *     RayTracer::init / clear / render / getSPP / saveHDR RTBase/Renderer.h:45-67, 876-898
 * Everything below `RayTracer::render()` (Renderer.h:876-885) runs on the device behind the
 * functions declared here.  Plain pointers and sizes only; no C++ or torch types; no
 * exceptions cross; every function returns an rtb_status (0 = OK, <0 = error) unless noted
 * and the message of the last error is available from rtb_last_error().
 *
 * A context drives ONE device (rtb_create) or a GROUP of up to 8 devices of one box
 * (rtb_create_multi): like RayTracer::init sizes the reference to every processor of the machine
 * (Renderer.h:52-55) and pathTracerTileBased fans out over them (:836-853), a group renders every
 * call on all its GPUs and sums their films at read-out — inside the library, exactly (64-bit
 * fixed-point sums), over NVLink peer memory or NCCL.  With one process per GPU instead
 * (rtb_params.partition), the caller reduces rtb_accum_device_ptr() itself (an int64 SUM).
 * RayTracer::render's alternatives that the reference ships commented out (adaptiveRender,
 * lightTracer, instantRadiosity, computeDirectMIS) are here too.  A context is not thread-safe.
 */
#ifndef RTB_H_
#define RTB_H_

#include <stddef.h>
#include <math.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTB_ABI_VERSION 1

typedef struct rtb_ctx rtb_ctx;

typedef enum rtb_status {
	RTB_OK = 0,
	RTB_ERR_ARG = -1,   /* null pointer, bad enum, size mismatch                     */
	RTB_ERR_CUDA = -2,  /* a CUDA runtime call or kernel failed                        */
	RTB_ERR_STATE = -3, /* call order: e.g. render before upload_scene                */
	RTB_ERR_OOM = -4,   /* host or device allocation failed                           */
	RTB_ERR_NODEV = -5  /* no CUDA device / device index out of range                 */
} rtb_status;

/* ------------------------------------------------------------------------------------
 * Scene description (POD mirror of RTBase/Scene.h:72-81 `Scene`).  All arrays are copied
 * by rtb_upload_scene; the caller keeps ownership.  float = IEEE binary32, little endian.
 * ---------------------------------------------------------------------------------- */

/* Camera (RTBase/Scene.h:10-54).  Matrices row-major like RTBase/Core.h:205-212.       */
typedef struct rtb_camera {
	float inv_proj[16];     /* Camera::inverseProjectionMatrix                       */
	float cam_to_world[16]; /* Camera::camera                                        */
	float origin[3];        /* Camera::origin                                        */
	float width, height;    /* Camera::width/height (floats, as in the reference)    */
	float pad_[3];
} rtb_camera; /* 160 B */

/* What Camera::init / Camera::updateView (RTBase/Scene.h:22-41) derive besides the fields above; only
 * light tracing (connectToCamera, Renderer.h:233-260) needs them.  rtb_camera carries the two matrices the
 * camera rays need bit-exactly; the rest is re-derived from them here — one definition, used by the library
 * and by its test oracle.  Matrices are row-major like RTBase's Matrix::m (Core.h:238-262). */
typedef struct rtb_camera_ext {
	float proj[16];         /* Camera::projectionMatrix = inverse of inv_proj         */
	float world_to_cam[16]; /* Camera::cameraToView     = inverse of cam_to_world     */
	float view_dir[3];      /* Camera::viewDirection                                  */
	float afilm;            /* Camera::Afilm                                          */
} rtb_camera_ext;

static inline int rtb_invert4(const float* m, float* out)
{
	double a[4][8];
	int i, j, k;
	for (i = 0; i < 4; i++)
		for (j = 0; j < 4; j++) a[i][j] = m[i * 4 + j], a[i][4 + j] = (i == j) ? 1.0 : 0.0;
	for (i = 0; i < 4; i++)
	{
		int piv = i;
		double best = a[i][i] < 0 ? -a[i][i] : a[i][i], f;
		for (k = i + 1; k < 4; k++)
		{
			double v = a[k][i] < 0 ? -a[k][i] : a[k][i];
			if (v > best) best = v, piv = k;
		}
		if (best == 0.0) return 0;
		if (piv != i)
			for (j = 0; j < 8; j++) f = a[i][j], a[i][j] = a[piv][j], a[piv][j] = f;
		f = 1.0 / a[i][i];
		for (j = 0; j < 8; j++) a[i][j] *= f;
		for (k = 0; k < 4; k++)
			if (k != i && a[k][i] != 0.0)
			{
				f = a[k][i];
				for (j = 0; j < 8; j++) a[k][j] -= f * a[i][j];
			}
	}
	for (i = 0; i < 4; i++)
		for (j = 0; j < 4; j++) out[i * 4 + j] = (float)a[i][4 + j];
	return 1;
}

static inline int rtb_camera_derive(const rtb_camera* c, rtb_camera_ext* e)
{
	const float* ip = c->inv_proj;
	const float* cw = c->cam_to_world;
	float vx, vy, vz, w, dx, dy, dz, l, wlens, aspect;
	if (!rtb_invert4(ip, e->proj) || !rtb_invert4(cw, e->world_to_cam)) return 0;
	/* viewDirection = normalize(camera.mulVec(inverseProjection.mulPointAndPerspectiveDivide((0,0,1)))) */
	w = 1.0f / (ip[14] + ip[15]);
	vx = (ip[2] + ip[3]) * w, vy = (ip[6] + ip[7]) * w, vz = (ip[10] + ip[11]) * w;
	dx = (vx * cw[0] + vy * cw[1]) + vz * cw[2];
	dy = (vx * cw[4] + vy * cw[5]) + vz * cw[6];
	dz = (vx * cw[8] + vy * cw[9]) + vz * cw[10];
	l = 1.0f / sqrtf((dx * dx + dy * dy) + dz * dz);
	e->view_dir[0] = dx * l, e->view_dir[1] = dy * l, e->view_dir[2] = dz * l;
	/* Afilm = Wlens * Hlens, Wlens = 2 / P[1][1], Hlens = Wlens * (P[0][0] / P[1][1])  (Scene.h:28-31) */
	wlens = 2.0f / e->proj[5];
	aspect = e->proj[0] / e->proj[5];
	e->afilm = wlens * (wlens * aspect);
	return 1;
}

/* One node of the reference BVH (RTBase/Geometry.h:294-313 `BVHNode`), flattened in
 * pre-order (left subtree directly follows its parent).  The device needs the reference's
 * own tree for EXACT traversal and its leaves (<= MAXNODE_TRIANGLES = 2 triangles,
 * Geometry.h:240,337) as the primitives of the accelerated tree.
 *   interior: a = index of l (>= 0), b = index of r
 *   leaf    : a = ~startIndex (< 0), b = endIndex - startIndex                         */
typedef struct rtb_ref_node {
	float bmin[3];
	int32_t a;
	float bmax[3];
	int32_t b;
} rtb_ref_node; /* 32 B */

/* Intersection record of one triangle (RTBase/Geometry.h:62-83 `Triangle`): exactly the
 * fields Triangle::rayIntersect (Geometry.h:89-105) reads.  n, d, area are the values
 * Triangle::init computed; inv_area = 1.0f / Dot(e1.cross(e2), n) (Geometry.h:98), hoisted.
 * e1 = v2 - v1 and e2 = v0 - v2 are re-derived on the device (single rounding, same bits). */
typedef struct rtb_tri_isect {
	float v0[3], d;
	float v1[3], inv_area;
	float v2[3];
	uint32_t material;
	float n[3], area;
} rtb_tri_isect; /* 64 B */

/* Shading attributes (Geometry.h:106-112, 127-130): vertex normals / uvs and the sign used
 * by Triangle::gNormal(): gsign = Dot(vertices[0].normal, n) > 0 ? +1 : -1.              */
typedef struct rtb_tri_shade {
	float n0[3], u0;
	float n1[3], u1;
	float n2[3], u2;
	float tv0, tv1, tv2, gsign;
} rtb_tri_shade; /* 64 B */

/* BSDF classes of RTBase/Materials.h:118-511. */
typedef enum rtb_bsdf_type {
	RTB_BSDF_DIFFUSE = 0,    /* DiffuseBSDF    Materials.h:118 */
	RTB_BSDF_MIRROR = 1,     /* MirrorBSDF     Materials.h:158 */
	RTB_BSDF_CONDUCTOR = 2,  /* ConductorBSDF  Materials.h:203 */
	RTB_BSDF_GLASS = 3,      /* GlassBSDF      Materials.h:252 */
	RTB_BSDF_DIELECTRIC = 4, /* DielectricBSDF Materials.h:320 */
	RTB_BSDF_ORENNAYAR = 5,  /* OrenNayarBSDF  Materials.h:369 */
	RTB_BSDF_PLASTIC = 6     /* PlasticBSDF    Materials.h:414 */
} rtb_bsdf_type;

#define RTB_MAT_SPECULAR 1u  /* isPureSpecular()                                      */
#define RTB_MAT_TWO_SIDED 2u /* isTwoSided() (LayeredBSDF forces true, :503-506)      */
#define RTB_MAT_LIGHT 4u     /* isLight(): emission.Lum() > 0 (:103-106)              */
#define RTB_MAT_LAYERED 8u   /* wrapped in LayeredBSDF (:467); `type` is the base's   */

typedef struct rtb_material {
	uint32_t type;  /* rtb_bsdf_type of the (base) BSDF                              */
	uint32_t flags; /* RTB_MAT_*                                                     */
	int32_t tex;    /* index into textures: the BSDF's `albedo`                      */
	float int_ior, ext_ior;
	float emission[3];
	float alpha;    /* 1.62142f*sqrtf(roughness) (Conductor/Dielectric/Plastic) or sigma (OrenNayar) */
	float eta[3], k[3]; /* ConductorBSDF                                             */
	float thickness;    /* LayeredBSDF                                               */
} rtb_material; /* 64 B */

/* Texture (RTBase/Imaging.h:19-31): `offset` counts TEXELS into the shared texel pool,
 * 3 floats per texel (Colour r,g,b), row-major.                                          */
typedef struct rtb_texture {
	uint32_t offset;
	int32_t width, height;
	int32_t pad_;
} rtb_texture; /* 16 B */

typedef enum rtb_light_type {
	RTB_LIGHT_AREA = 0,       /* AreaLight        RTBase/Lights.h:30  */
	RTB_LIGHT_BACKGROUND = 1, /* BackgroundColour RTBase/Lights.h:84  */
	RTB_LIGHT_ENVMAP = 2      /* EnvironmentMap   RTBase/Lights.h:135 */
} rtb_light_type;

/* Entry of Scene::lights (RTBase/Scene.h:77), in the reference's order: the background
 * first when totalIntegratedPower() > 0 (Scene.h:156-159), then one AreaLight per
 * emissive triangle in ascending triangle index (Scene.h:96-105).                       */
typedef struct rtb_light {
	uint32_t type;
	uint32_t triangle; /* area: index into triangles                                */
	float emission[3]; /* area: AreaLight::emission; background: colour             */
	float area;        /* area: Triangle::area                                      */
	int32_t tex;       /* envmap: texture index                                     */
	int32_t pad_;
} rtb_light; /* 32 B */

typedef struct rtb_scene_desc {
	rtb_camera camera;
	const rtb_ref_node* ref_nodes;
	uint32_t n_ref_nodes;
	uint32_t n_tris;
	const rtb_tri_isect* tri_isect; /* [n_tris], index = position in Scene::triangles after build() */
	const rtb_tri_shade* tri_shade; /* [n_tris]                                              */
	const rtb_material* materials;
	uint32_t n_materials;
	uint32_t n_textures;
	const rtb_texture* textures;
	const float* texels; /* [n_texels*3]                                              */
	uint64_t n_texels;
	const rtb_light* lights;
	uint32_t n_lights;
	/* Scene::background (always present; it is also lights[0] when it has power). */
	uint32_t background_type; /* RTB_LIGHT_BACKGROUND or RTB_LIGHT_ENVMAP          */
	float background_colour[3];
	int32_t background_tex;
} rtb_scene_desc;

/* ------------------------------------------------------------------------------------
 * Parameters (the reference's compile-time constants and commented-out switches made
 * run-time; defaults reproduce the reference as committed).
 * ---------------------------------------------------------------------------------- */
typedef enum rtb_integrator {
	RTB_INT_PATH = 0,    /* RayTracer::pathTrace   Renderer.h:328-392 (the default)   */
	RTB_INT_DIRECT = 1,  /* RayTracer::direct      Renderer.h:393-407                */
	RTB_INT_ALBEDO = 2,  /* RayTracer::albedo      Renderer.h:558-571                */
	RTB_INT_NORMALS = 3, /* RayTracer::viewNormals Renderer.h:572-581                */
	RTB_INT_PATH_MIS = 4 /* pathTrace with computeDirectMIS (Renderer.h:474-557) in place of
	                      * computeDirect: the estimator the reference ships but leaves switched
	                      * off.  Wavefront: the light strategy's segment and the BSDF strategy's
	                      * probe ray share one queue record (k_wf_mis).                        */
} rtb_integrator;

/* Philox block index of computeDirectMIS's BSDF-strategy uniforms at path depth k: RTB_RNG_MIS_BLOCK + k
 * (blocks 2k and 2k+1 belong to pathTrace's own draws). */
#define RTB_RNG_MIS_BLOCK 0x40000000u

typedef enum rtb_sampling {
	RTB_SAMPLING_STRICT = 0,    /* sampling decisions exactly as the reference: uniform
	                               light pick, uniform-sphere env, no MIS (Renderer.h:423-473) */
	RTB_SAMPLING_IMPORTANCE = 1 /* same estimator expectation, lower variance: env-map
	                               luminance CDF for the NEE direction      */
} rtb_sampling;

typedef enum rtb_traversal {
	RTB_TRAV_EXACT = 0, /* the reference's own tree, exhaustive DFS (Geometry.h:399-427) */
	RTB_TRAV_FAST = 1,  /* accelerated binary tree over the reference's leaves, ordered + culled;
	                       must return identical hits (tests assert it)                */
	RTB_TRAV_WIDE = 2,  /* the same tree collapsed to 4 children per node (selectable; measured
	                       ~5 % slower than FAST on B200, profiles/r01_wide_tree.txt)     */
	RTB_TRAV_CW = 3,    /* the same tree collapsed to 8 children per 80-byte node with 8-bit quantised,
	                       CONSERVATIVE child boxes (compressed wide BVH), children visited in ray-octant
	                       order, top levels staged in shared memory; the exact reference leaf box is still
	                       tested before a leaf's triangles, so the hits stay the reference's            */
	RTB_TRAV_Q16 = 4    /* FAST's binary tree in 32-byte nodes: both child boxes quantised conservatively to 16 bits
	                       per plane on one grid over the scene box (half the node bytes and fetches of FAST, a
	                       cheaper FMA-form box test); exact reference leaf boxes kept and tested before triangles */
} rtb_traversal;

typedef enum rtb_filter {
	RTB_FILTER_BOX = 0,     /* BoxFilter size 0 (Imaging.h:139-154, Renderer.h:50)     */
	RTB_FILTER_GAUSSIAN = 1 /* GaussianFilter(radius, alpha) (Imaging.h:155-187)       */
} rtb_filter;

typedef enum rtb_scheduler {
	RTB_SCHED_WAVEFRONT = 0, /* staged extend / shade / shadow kernels over a pool of path slots  */
	RTB_SCHED_MEGAKERNEL = 1 /* one thread per pixel runs whole paths (kept for A/B profiling)    */
} rtb_scheduler;

typedef enum rtb_partition {
	RTB_PART_NONE = 0,
	RTB_PART_SPP = 1, /* this rank renders sample indices s with s % world == rank      */
	RTB_PART_TILE = 2 /* this rank renders 32x32 tiles t with t % world == rank          */
} rtb_partition;

typedef struct rtb_params {
	int32_t max_depth;  /* MAX_DEPTH, Renderer.h:20 (4)                              */
	float epsilon;      /* EPSILON, Geometry.h:60 (1e-4f)                            */
	float rr_cap;       /* 0.9f, Renderer.h:353                                      */
	int32_t integrator; /* rtb_integrator                                            */
	int32_t sampling;   /* rtb_sampling                                              */
	int32_t traversal;  /* rtb_traversal                                             */
	int32_t filter;     /* rtb_filter                                                */
	float filter_radius, filter_alpha; /* Gaussian (2.0, 0.1 at Renderer.h:51)        */
	uint32_t seed;      /* key of the counter-based RNG                              */
	int32_t partition;  /* rtb_partition                                             */
	int32_t part_rank, part_world;
	float cull_rel;     /* FAST traversal: relative slack of the t-cull (1e-5)       */
	int32_t scheduler;  /* rtb_scheduler                                             */
	int32_t primary_reuse; /* wavefront: trace each pixel's camera ray once per rtb_render call and let
	                        * every sample of the pixel start from that hit.  Exact: the reference samples
	                        * pixel centres only (Renderer.h:806-807) and generateRay draws no random
	                        * numbers, so all samples of a pixel share one primary ray.  0 = trace it per
	                        * sample like the reference does; the film is bit-identical either way. */
} rtb_params; /* 64 B */

/* Ray / hit records of the batched parity entry points. */
typedef struct rtb_ray {
	float o[3];
	float tmax; /* any-hit: maxT of BVHNode::traverseVisible; closest-hit: ignored   */
	float d[3];
	float pad_;
} rtb_ray; /* 32 B */

/* == IntersectionData, RTBase/Geometry.h:231-238.  Miss: t = FLT_MAX, id = 0xFFFFFFFF. */
typedef struct rtb_hit {
	uint32_t id;
	float t, alpha, beta, gamma;
} rtb_hit; /* 20 B */

/* == the float members of ShadingData, RTBase/Materials.h:15-35 (Vec3 w dropped). */
typedef struct rtb_shading {
	float x[3], wo[3], s_normal[3], g_normal[3];
	float tu, tv;
	float frame_u[3], frame_v[3], frame_w[3];
	float t;
	int32_t material; /* index of `bsdf` in Scene::materials; -1 on miss             */
} rtb_shading; /* 100 B */

typedef struct rtb_stats {
	uint64_t samples;       /* pixel samples completed since the last clear        */
	uint64_t closest_rays;  /* Scene::traverse calls                               */
	uint64_t shadow_rays;   /* Scene::visible calls                                */
	uint64_t kernel_launches; /* kernels of this library launched since create     */
	double render_ms;       /* device time of the render calls since last clear    */
	uint64_t box_tests, tri_tests; /* closest-hit traversal work (both children of a visited node count) */
	uint64_t shadow_box_tests, shadow_tri_tests; /* any-hit traversal work          */
	/* wavefront schedule: device time of the three stage kernels, measured with CUDA events
	 * on every 8th iteration (timed_iterations of them) since the last clear             */
	double extend_ms, shade_ms, shadow_ms;
	uint64_t timed_iterations, iterations, host_syncs;
} rtb_stats;

/* ------------------------------------------------------------------------------------
 * Life cycle
 * ---------------------------------------------------------------------------------- */
int rtb_abi_version(void);
/* Creates a context on CUDA device `device`.  Fails with RTB_ERR_NODEV when there is no
 * such device: there is NO CPU fallback.                                                */
int rtb_create(int device, rtb_ctx** out);
/* Number of CUDA devices this process can see (0 = none: there is no renderer without one). */
int rtb_device_count(void);
/* A device GROUP: one context that renders on n GPUs of this box (RayTracer::init's numProcs =
 * every processor, Renderer.h:52-55).  devices = n distinct CUDA device indices, or NULL for the
 * first n visible devices (n <= 0: all of them).  The scene is replicated; every rtb_render /
 * rtb_render_light / rtb_render_ir call is split over the members (sample slices — tile slices
 * when a call has fewer samples than devices or rtb_params.partition asks for tiles — resp.
 * contiguous pass ranges), each member driven by its own host thread for the duration of the call;
 * a partition set in rtb_params (one process per node, say) is refined, not replaced.  The film
 * is read out on devices[0]: rtb_read_film / rtb_tonemap / rtb_*_device_ptr first add the other
 * members' fixed-point sums to devices[0]'s — one kernel over NVLink peer memory that also converts
 * to the float film, or ncclReduce (libnccl.so.2, loaded on demand) when a member is not
 * peer-accessible; RTB_GROUP_REDUCE=nccl|staged forces a path.  Integer sums: the group's film equals
 * the single-GPU film bit for bit.  rtb_render_adaptive: every member steers and samples the 32x32
 * tiles t with t % n == member (same plan, same film as one GPU).  The parity entry points run on
 * devices[0] alone.  n = 1 is a plain context.                                             */
int rtb_create_multi(const int* devices, int n, rtb_ctx** out);
/* Members of ctx (1 for rtb_create).  rtb_group_info: devices[n], p2p[n] (1 = devices[0] reads that
 * member's memory directly) and how many read-outs went over peer memory / NCCL; any may be NULL. */
int rtb_group_size(const rtb_ctx* ctx);
int rtb_group_info(const rtb_ctx* ctx, int* devices, int* p2p, uint64_t* gathers_p2p, uint64_t* gathers_nccl);
void rtb_destroy(rtb_ctx* ctx);
/* Message of the last failing call on ctx (ctx may be NULL: creation errors).  Never NULL. */
const char* rtb_last_error(const rtb_ctx* ctx);
/* All work of ctx is issued on `cuda_stream` (a cudaStream_t; NULL = the legacy default
 * stream) so that the caller's events/graphs order against it.                          */
int rtb_set_stream(rtb_ctx* ctx, void* cuda_stream);
int rtb_synchronize(rtb_ctx* ctx);

/* ------------------------------------------------------------------------------------
 * Scene + parameters
 * ---------------------------------------------------------------------------------- */
void rtb_default_params(rtb_params* p);
int rtb_set_params(rtb_ctx* ctx, const rtb_params* p);
int rtb_get_params(const rtb_ctx* ctx, rtb_params* p);
/* Copies the scene to the device, builds the accelerated tree over the reference leaves
 * (host threads: binned SAH; RTB_GPU_BUILD=1: a linear BVH built on the device in milliseconds,
 * same hits, ~10-25 % lower render rate) and (re)allocates a cleared film of camera.width x
 * camera.height.  Replaces RayTracer::init (Renderer.h:45-63).                             */
int rtb_upload_scene(rtb_ctx* ctx, const rtb_scene_desc* scene);
/* Camera::updateView / RTCamera::updateCamera (Scene.h:33-41): new camera, same scene.   */
int rtb_update_camera(rtb_ctx* ctx, const rtb_camera* cam);

/* ------------------------------------------------------------------------------------
 * The hot path
 * ---------------------------------------------------------------------------------- */
/* RayTracer::clear (Renderer.h:64-67): zero the film sums and SPP.                       */
int rtb_clear(rtb_ctx* ctx);
/* spp_count calls of RayTracer::render() (Renderer.h:876-885) in one go: accumulates the
 * samples with global indices [spp_begin, spp_begin + spp_count) of every pixel this rank
 * owns (see rtb_params.partition) into the device sum-film.  ASYNCHRONOUS once the context has
 * rendered this configuration (scene, sample count, parameters) before: the number of wavefront
 * iterations it needed is remembered, the call enqueues that many (+ 6 % + 3; launches past the
 * point where the pool drains exit at once) on the context's stream and returns; the next call
 * on the context checks that the pool did drain and enqueues the rest if not.  The FIRST render
 * of a configuration finds the iteration count with 2-5 small host probes and returns when its
 * last batch is enqueued.  RTB_ASYNC_RENDER=0 keeps every call on the probing path.  Resumable:
 * the RNG is keyed by (seed, pixel, sample index), not by call order.                        */
int rtb_render(rtb_ctx* ctx, uint32_t spp_begin, uint32_t spp_count);
/* RayTracer::adaptiveRender (Renderer.h:679-749; shipped commented out at :880), one call = one
 * render(): (1) adaptiveSampling (:583-641): init_samples (INIT_SAMPLES = 2, :23) paths per pixel,
 * per 32x32 tile the variance of the pixel means around the tile mean; (2) weight = variance share,
 * samples = max((int)(sqrt(weight) * max_samples), min_samples) (MAX_SAMPLES = 10240, MIN_SAMPLES = 1,
 * :21-22, :649-653); (3) sampleTileWithWeight (:645-677): that many FRESH samples per pixel of the tile,
 * their mean is splatted, i.e. the film sum grows by one mean image and SPP by one.  The initial samples
 * only steer (sample indices 0..init-1), the splatted ones use indices init.. .  tile_samples /
 * tile_variance (tilesX*tilesY entries, row-major 32x32 tiles; may be NULL) receive the plan.
 * Wavefront schedule; one context per image (a device group shares the tiles out).  Synchronous (the
 * plan goes through the host).                                                                       */
int rtb_render_adaptive(rtb_ctx* ctx, uint32_t init_samples, uint32_t min_samples, uint32_t max_samples,
                        uint32_t* tile_samples, float* tile_variance);
/* pass_count x RayTracer::lightTracer() (Renderer.h:220-326; a commented-out alternative in render(), :883):
 * each pass traces width*height paths FROM the area lights and connects every diffuse vertex to the camera
 * (connectToCamera, :233-260); splats land anywhere on the film (box filter).  Passes are numbered for the
 * counter-based RNG like samples are; SPP grows by pass_count.  Scenes without area lights add nothing.
 * Single device per image (shard passes, not pixels).  Asynchronous.                                   */
int rtb_render_light(rtb_ctx* ctx, uint32_t pass_begin, uint32_t pass_count);
/* pass_count x RayTracer::instantRadiosity() (Renderer.h:82-218; a commented-out alternative in render(), :884):
 * per pass n_paths light paths (MAX_VPL = 50, :24) leave virtual point lights at their diffuse vertices
 * (traceVPLs / VPLTracePath, :159-218), then every pixel's primary hit gathers all of them with one visibility
 * ray each (computeVPLsContribution, :124-158).  SPP grows by pass_count.  Single device per image (shard the
 * passes, not the pixels).  Asynchronous.                                                                 */
int rtb_render_ir(rtb_ctx* ctx, uint32_t pass_begin, uint32_t pass_count, uint32_t n_paths);
/* Film::film (Imaging.h:204): waits for the device and copies the running SUM (not the
 * mean; Film::save divides by SPP, Imaging.h:262-271) as width*height*3 floats (r,g,b per
 * pixel, row-major) to host memory; *spp receives Film::SPP.  Either may be NULL.         */
int rtb_read_film(rtb_ctx* ctx, float* rgb_sum, uint32_t* spp);
/* The inverse of rtb_read_film: the film sums become rgb_sum (width*height*3 floats).  This is the hook for
 * RayTracer::denoise (Renderer.h:750-792), which copies film->film out, runs an external filter (OIDN there)
 * and copies the result back: read, filter with anything, write.  SPP is unchanged.  Synchronous.          */
int rtb_write_film(rtb_ctx* ctx, const float* rgb_sum);
/* Device address of the sum film (width*height*3 floats), e.g. for an NCCL reduce.         */
int rtb_film_device_ptr(rtb_ctx* ctx, void** dptr, uint64_t* n_floats);
/* The film's master copy: width*height*3 signed 64-bit FIXED-POINT sums in units of 2^-32
 * (integer addition is associative: the film is bit-reproducible whatever the scheduling, and
 * an int64 SUM-reduce of this buffer over the ranks of a tile- or spp-partitioned render gives
 * exactly the single-GPU film).  After reducing into it, call rtb_set_spp with the global
 * sample count; the float film is re-derived on the next read.                             */
int rtb_accum_device_ptr(rtb_ctx* ctx, void** dptr, uint64_t* n_int64);
int rtb_set_spp(rtb_ctx* ctx, uint32_t spp);
/* Film::tonemap for every pixel (Imaging.h:233-242; what presentFilmToCanvas draws,
 * Renderer.h:69-80): width*height*3 bytes r,g,b.                                          */
int rtb_tonemap(rtb_ctx* ctx, uint8_t* rgb8, float exposure);
int rtb_get_stats(rtb_ctx* ctx, rtb_stats* out);
int rtb_film_size(const rtb_ctx* ctx, uint32_t* width, uint32_t* height);

/* ------------------------------------------------------------------------------------
 * Parity entry points (batched restatements of single reference calls; host buffers)
 * ---------------------------------------------------------------------------------- */
/* Camera::generateRay(x+0.5, y+0.5) + Scene::traverse for every pixel (Renderer.h:806-808,
 * Scene.h:43-54,107-130).  ids/t/rays row-major; any pointer may be NULL.  `traversal`
 * is an rtb_traversal.                                                                   */
int rtb_primary_hits(rtb_ctx* ctx, int traversal, uint32_t* ids, float* t, rtb_ray* rays);
/* Scene::traverse (any_hit = 0) or BVHNode::traverseVisible with maxT = ray.tmax
 * (any_hit = 1; hit.id = 1 if occluded, 0 if visible).                                  */
int rtb_trace(rtb_ctx* ctx, int traversal, int any_hit, const rtb_ray* rays, uint64_t n, rtb_hit* hits);
/* Scene::visible(p1, p2) (Scene.h:161-169): p1p2 = n*6 floats; out[i] = 1 if visible.    */
int rtb_visible(rtb_ctx* ctx, int traversal, const float* p1p2, uint64_t n, uint8_t* out);
/* Scene::calculateShadingData (Scene.h:174-203).                                         */
int rtb_shading_data(rtb_ctx* ctx, const rtb_ray* rays, const rtb_hit* hits, uint64_t n, rtb_shading* out);
/* BSDF::evaluate / PDF at wi and BSDF::sample with the uniforms u[3] (u[0],u[1] = the two
 * cosine-hemisphere draws in reference order r1,r2; u[2] = the glass reflect/refract draw).
 * Any output may be NULL.  eval: n*3, pdf: n, s_wi: n*3, s_f: n*3, s_pdf: n.             */
int rtb_eval_bsdf(rtb_ctx* ctx, const rtb_shading* sd, const float* wi, const float* u, uint64_t n,
                  float* eval, float* pdf, float* s_wi, float* s_f, float* s_pdf);
/* Light::sample with uniforms u[2] (r1, r2) for light `light[i]` -> p_or_wi (n*3),
 * emitted (n*3), pdf (n); Light::evaluate(wi) -> eval (n*3).  The direction sample of an environment map
 * follows rtb_params.sampling (uniform sphere, or the luminance CDF).                      */
int rtb_eval_light(rtb_ctx* ctx, const int32_t* light, const float* wi, const float* u, uint64_t n,
                   float* p_or_wi, float* emitted, float* pdf, float* eval);
/* The uniforms the render kernel draws: out[i] = u(pixel, sample, dim i), dims [0, n).    */
int rtb_rng_draws(rtb_ctx* ctx, uint32_t pixel, uint32_t sample, uint32_t n, float* out);

#ifdef __cplusplus
}
#endif
#endif /* RTB_H_ */
