/*
 * TEST INFRASTRUCTURE (oracle/): plain-C restatement of the reference's path-tracing hot
 * path on the flat scene arrays of include/rtb.h.  It is the CHECKER for the CUDA path in
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg; the product never links,
 * loads or calls it.
 *
 * Pinned (tests/test_oracle_cpu.py) against the reference itself — oracle/_ref/librtref.so,
 * the unmodified RTBase headers compiled headless — and against the committed fixtures in
 * tests/golden/: primary hit IDs and t bit-exact on all bundled scenes, shading data /
 * BSDF / light vectors within 1e-5 relative, rendered films statistically.
 *
 * Each function cites the reference code it follows (paths relative to
 * /root/reference/RTBase).  Build: gcc -O2 -ffp-contract=off (never -ffast-math): the hit
 * decisions must round like the reference's g++ -ffp-contract=off build.
 *
 * One deliberate difference from the reference: MTRandom (Sampling.h:13-26) is replaced by
 * the counter-based Philox-4x32-10 stream the CUDA kernels use, keyed by (seed, pixel,
 * sample) with the fixed dimension layout documented at rng_block() — so a pixel's value
 * does not depend on thread scheduling and GPU and oracle consume identical uniforms.
 */
#include "../include/rtb.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float x, y, z; } v3;

static v3 V(float x, float y, float z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
static v3 Vp(const float* p) { return V(p[0], p[1], p[2]); }
static v3 add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }            /* Core.h:128 */
static v3 sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }            /* Core.h:132 */
static v3 mul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }            /* Core.h:57,144 */
static v3 scl(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }               /* Core.h:136 */
static v3 dvd(v3 a, float s) { return V(a.x / s, a.y / s, a.z / s); }               /* Core.h:81 */
static v3 neg(v3 a) { return V(-a.x, -a.y, -a.z); }
static float dot3(v3 a, v3 b) { return ((a.x * b.x) + (a.y * b.y)) + (a.z * b.z); } /* Core.h:176 */
static v3 cross3(v3 a, v3 b)                                                        /* Core.h:170 */
{
	return V((a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x));
}
static v3 norm3(v3 a)                                                               /* Core.h:161 */
{
	float l = 1.0f / sqrtf(((a.x * a.x) + (a.y * a.y)) + (a.z * a.z));
	return V(a.x * l, a.y * l, a.z * l);
}
static float lum3(v3 c) { return ((0.2126f * c.x) + (0.7152f * c.y)) + (0.0722f * c.z); } /* Core.h:89 */
static float std_max(float a, float b) { return (a < b) ? b : a; }
static float std_min(float a, float b) { return (b < a) ? b : a; }
static float win_max(float a, float b) { return (a > b) ? a : b; } /* Windows max() / Core.h:187 Max */
static float win_min(float a, float b) { return (a < b) ? a : b; } /* Windows min() / Core.h:192 Min */

typedef struct { v3 o, d, inv; } ray_t;
static ray_t make_ray(v3 o, v3 d) /* Geometry.h:21-26 */
{
	ray_t r;
	r.o = o;
	r.d = d;
	r.inv = V(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
	return r;
}

typedef struct {
	uint64_t closest, shadow, samples;
	uint64_t cbox, ctri, sbox, stri; /* canonical-traversal work (oracle_render_counts only) */
	uint64_t cmax, smax, cbig, nanrays; /* largest box count of one closest / shadow ray, rays above 20 x 2^10 boxes, NaN rays */
} tally_t;

static int g_count_canonical = 0;

/* ---------------- Camera::generateRay, Scene.h:43-54; Core.h:295-309 ------------------ */
static ray_t generate_ray(const rtb_camera* c, float x, float y)
{
	float xprime = x / c->width;
	float yprime = 1.0f - (y / c->height);
	const float* m = c->inv_proj;
	const float* k = c->cam_to_world;
	v3 d, w;
	xprime = (xprime * 2.0f) - 1.0f;
	yprime = (yprime * 2.0f) - 1.0f;
	d.x = (xprime * m[0] + yprime * m[1] + 1.0f * m[2]) + m[3];
	d.y = (xprime * m[4] + yprime * m[5] + 1.0f * m[6]) + m[7];
	d.z = (xprime * m[8] + yprime * m[9] + 1.0f * m[10]) + m[11];
	w.x = (d.x * k[0] + d.y * k[1] + d.z * k[2]);
	w.y = (d.x * k[4] + d.y * k[5] + d.z * k[6]);
	w.z = (d.x * k[8] + d.y * k[9] + d.z * k[10]);
	return make_ray(Vp(c->origin), norm3(w));
}

/* ---------------- AABB::rayAABB, Geometry.h:173-184 ----------------------------------- */
static int ray_aabb(const rtb_ref_node* n, const ray_t* r)
{
	v3 tmin = mul(sub(Vp(n->bmin), r->o), r->inv);
	v3 tmax = mul(sub(Vp(n->bmax), r->o), r->inv);
	v3 en = V(win_min(tmin.x, tmax.x), win_min(tmin.y, tmax.y), win_min(tmin.z, tmax.z));
	v3 ex = V(win_max(tmin.x, tmax.x), win_max(tmin.y, tmax.y), win_max(tmin.z, tmax.z));
	float t_entry = std_max(std_max(en.x, en.y), en.z);
	float t_exit = std_min(std_min(ex.x, ex.y), ex.z);
	if (t_exit < t_entry || t_exit < 0) return 0;
	return 1;
}

/* ---------------- Triangle::rayIntersect, Geometry.h:89-105 --------------------------- */
static int tri_intersect(const rtb_tri_isect* q, const ray_t* r, float* t, float* u, float* v)
{
	v3 n = Vp(q->n), v0 = Vp(q->v0), v1 = Vp(q->v1), v2 = Vp(q->v2);
	v3 e1 = sub(v2, v1), e2 = sub(v0, v2), p;
	float denom = dot3(n, r->d);
	if (denom == 0) return 0;
	*t = (q->d - dot3(n, r->o)) / denom;
	if (*t < 0) return 0;
	p = add(r->o, scl(r->d, *t));
	*u = dot3(cross3(e1, sub(p, v1)), n) * q->inv_area;
	if (*u < 0 || *u > 1.0f) return 0;
	*v = dot3(cross3(e2, sub(p, v2)), n) * q->inv_area;
	if (*v < 0 || (*u + *v) > 1.0f) return 0;
	return 1;
}

/* ---------------- BVHNode::traverse, Geometry.h:399-427 (recursive, exhaustive) -------- */
static void bvh_traverse(const rtb_scene_desc* s, int32_t node, const ray_t* r, float eps, rtb_hit* best)
{
	const rtb_ref_node* n = &s->ref_nodes[node];
	if (!ray_aabb(n, r)) return;
	if (n->a < 0)
	{
		int32_t start = ~n->a, end = start + n->b, i;
		for (i = start; i < end; i++)
		{
			float t, u, v;
			if (tri_intersect(&s->tri_isect[i], r, &t, &u, &v))
			{
				if (t < best->t && t > eps)
				{
					best->t = t;
					best->id = (uint32_t)i;
					best->alpha = u;
					best->beta = v;
					best->gamma = 1.0f - (u + v);
				}
			}
		}
		return;
	}
	bvh_traverse(s, n->a, r, eps, best);
	bvh_traverse(s, n->b, r, eps, best);
}

static rtb_hit scene_traverse(const rtb_scene_desc* s, const ray_t* r, float eps) /* Scene.h:107-130 */
{
	rtb_hit h;
	h.id = 0xFFFFFFFFu;
	h.t = FLT_MAX;
	h.alpha = h.beta = h.gamma = 0;
	if (s->n_ref_nodes) bvh_traverse(s, 0, r, eps, &h);
	return h;
}

/* ---------------- BVHNode::traverseVisible, Geometry.h:435-462 ------------------------ */
static int bvh_visible(const rtb_scene_desc* s, int32_t node, const ray_t* r, float eps, float maxT)
{
	const rtb_ref_node* n = &s->ref_nodes[node];
	if (!ray_aabb(n, r)) return 1;
	if (n->a < 0)
	{
		int32_t start = ~n->a, end = start + n->b, i;
		for (i = start; i < end; i++)
		{
			float t, u, v;
			if (tri_intersect(&s->tri_isect[i], r, &t, &u, &v))
			{
				if (t >= maxT || t <= eps) continue;
				return 0;
			}
		}
		return 1;
	}
	if (!bvh_visible(s, n->a, r, eps, maxT)) return 0;
	return bvh_visible(s, n->b, r, eps, maxT);
}

static int scene_visible(const rtb_scene_desc* s, v3 p1, v3 p2, float eps) /* Scene.h:161-169 */
{
	v3 dir = sub(p2, p1);
	float maxT = sqrtf(dot3(dir, dir)) - (2.0f * eps);
	ray_t r;
	dir = norm3(dir);
	r = make_ray(add(p1, scl(dir, eps)), dir);
	if (!s->n_ref_nodes) return 1;
	return bvh_visible(s, 0, &r, eps, maxT);
}

/* ---------------- canonical traversal work counter -------------------------
 * The per-ray figures the roofline's algorithmic bytes/flops are defined on: iterative DFS on the
 * reference tree; at an interior node BOTH child boxes get the exact slab test; nearer child first;
 * a node is dropped when t_entry - |t_entry|*1e-5 > t_best (closest) or >= maxT (shadow); leaves
 * test all their triangles; shadow rays stop at the first accepted hit.  It only counts; the hit
 * decisions of the render stay those of bvh_traverse / bvh_visible above.                        */
static int slab_entry(const rtb_ref_node* n, const ray_t* r, float* t_in)
{
	v3 tmin = mul(sub(Vp(n->bmin), r->o), r->inv);
	v3 tmax = mul(sub(Vp(n->bmax), r->o), r->inv);
	v3 en = V(win_min(tmin.x, tmax.x), win_min(tmin.y, tmax.y), win_min(tmin.z, tmax.z));
	v3 ex = V(win_max(tmin.x, tmax.x), win_max(tmin.y, tmax.y), win_max(tmin.z, tmax.z));
	float t_entry = std_max(std_max(en.x, en.y), en.z);
	float t_exit = std_min(std_min(ex.x, ex.y), ex.z);
	if (t_exit < t_entry || t_exit < 0) return 0;
	*t_in = t_entry;
	return 1;
}

static void canonical_count(const rtb_scene_desc* s, const ray_t* r, float eps, int any_hit, float maxT, uint64_t* nbox, uint64_t* ntri)
{
	int32_t stack[256];
	float stackT[256];
	int sp = 0;
	float best = any_hit ? maxT : FLT_MAX, t0;
	if (!s->n_ref_nodes) return;
	(*nbox)++;
	if (!slab_entry(&s->ref_nodes[0], r, &t0)) return;
	stack[sp] = 0, stackT[sp] = t0, sp++;
	while (sp > 0)
	{
		int32_t node = stack[--sp];
		float te = stackT[sp];
		const rtb_ref_node* n = &s->ref_nodes[node];
		float lim = te - fabsf(te) * 1e-5f;
		if (any_hit ? (lim >= best) : (lim > best)) continue;
		if (n->a < 0)
		{
			int32_t start = ~n->a, end = start + n->b, i;
			for (i = start; i < end; i++)
			{
				float t, u, v;
				(*ntri)++;
				if (!tri_intersect(&s->tri_isect[i], r, &t, &u, &v)) continue;
				if (any_hit)
				{
					if (t < maxT && t > eps) return;
				}
				else if (t < best && t > eps)
					best = t;
			}
			continue;
		}
		{
			float ta = 0, tb = 0;
			int ha, hb;
			(*nbox) += 2;
			ha = slab_entry(&s->ref_nodes[n->a], r, &ta);
			hb = slab_entry(&s->ref_nodes[n->b], r, &tb);
			if (ha && hb)
			{
				int32_t nearN = (tb < ta) ? n->b : n->a, farN = (tb < ta) ? n->a : n->b;
				float nearT = (tb < ta) ? tb : ta, farT = (tb < ta) ? ta : tb;
				if (sp + 2 > 256) return;
				stack[sp] = farN, stackT[sp] = farT, sp++;
				stack[sp] = nearN, stackT[sp] = nearT, sp++;
			}
			else if (ha || hb)
			{
				if (sp + 1 > 256) return;
				stack[sp] = ha ? n->a : n->b, stackT[sp] = ha ? ta : tb, sp++;
			}
		}
	}
}

static rtb_hit scene_traverse_tl(const rtb_scene_desc* s, const ray_t* r, float eps, tally_t* tl)
{
	if (g_count_canonical)
	{
		uint64_t b0 = tl->cbox, nb;
		canonical_count(s, r, eps, 0, FLT_MAX, &tl->cbox, &tl->ctri);
		nb = tl->cbox - b0;
		if (nb > tl->cmax) tl->cmax = nb;
		if (nb > 20480) tl->cbig++;
		if (r->o.x != r->o.x || r->o.y != r->o.y || r->o.z != r->o.z || r->d.x != r->d.x || r->d.y != r->d.y || r->d.z != r->d.z) tl->nanrays++;
	}
	return scene_traverse(s, r, eps);
}

static int scene_visible_tl(const rtb_scene_desc* s, v3 p1, v3 p2, float eps, tally_t* tl)
{
	if (g_count_canonical)
	{
		v3 dir = sub(p2, p1);
		float maxT = sqrtf(dot3(dir, dir)) - (2.0f * eps);
		ray_t r;
		dir = norm3(dir);
		r = make_ray(add(p1, scl(dir, eps)), dir);
		{
			uint64_t b0 = tl->sbox;
			canonical_count(s, &r, eps, 1, maxT, &tl->sbox, &tl->stri);
			if (tl->sbox - b0 > tl->smax) tl->smax = tl->sbox - b0;
		}
	}
	return scene_visible(s, p1, p2, eps);
}

/* ---------------- ShadingData: Scene.h:174-203, Geometry.h:106-112,127-130, Core.h:513 -- */
typedef struct {
	v3 x, wo, sN, gN;
	float tu, tv;
	v3 fu, fv, fw;
	float t;
	int32_t mat;
} shade_t;

static void frame_from_vector(v3 n, v3* u, v3* v, v3* w)
{
	*w = norm3(n);
	if (fabsf(w->x) > fabsf(w->y))
	{
		float l = 1.0f / sqrtf(w->x * w->x + w->z * w->z);
		*u = V(w->z * l, 0.0f, -w->x * l);
	}
	else
	{
		float l = 1.0f / sqrtf(w->y * w->y + w->z * w->z);
		*u = V(0, w->z * l, -w->y * l);
	}
	*v = cross3(*w, *u);
}
static v3 to_local(const shade_t* s, v3 a) { return V(dot3(a, s->fu), dot3(a, s->fv), dot3(a, s->fw)); }
static v3 to_world(const shade_t* s, v3 a) { return add(add(scl(s->fu, a.x), scl(s->fv, a.y)), scl(s->fw, a.z)); }

static void shading_data(const rtb_scene_desc* s, const rtb_hit* h, const ray_t* r, shade_t* sd)
{
	memset(sd, 0, sizeof(*sd));
	sd->mat = -1;
	if (h->t < FLT_MAX)
	{
		const rtb_tri_isect* q = &s->tri_isect[h->id];
		const rtb_tri_shade* a = &s->tri_shade[h->id];
		v3 nrm;
		sd->x = add(r->o, scl(r->d, h->t));
		sd->gN = scl(Vp(q->n), a->gsign);
		nrm = add(add(scl(Vp(a->n0), h->alpha), scl(Vp(a->n1), h->beta)), scl(Vp(a->n2), h->gamma));
		sd->sN = norm3(nrm);
		sd->tu = a->u0 * h->alpha + a->u1 * h->beta + a->u2 * h->gamma;
		sd->tv = a->tv0 * h->alpha + a->tv1 * h->beta + a->tv2 * h->gamma;
		sd->mat = (int32_t)q->material;
		sd->wo = neg(r->d);
		if (s->materials[sd->mat].flags & RTB_MAT_TWO_SIDED)
		{
			if (dot3(sd->wo, sd->sN) < 0) sd->sN = neg(sd->sN);
			if (dot3(sd->wo, sd->gN) < 0) sd->gN = neg(sd->gN);
		}
		frame_from_vector(sd->sN, &sd->fu, &sd->fv, &sd->fw);
		sd->t = h->t;
	}
	else
	{
		sd->wo = neg(r->d);
		sd->t = h->t;
	}
}

/* ---------------- Texture::sample, Imaging.h:72-94 ------------------------------------ */
static v3 texture_sample(const rtb_scene_desc* s, int tex, float tu, float tv)
{
	const rtb_texture* T = &s->textures[tex];
	const float* px = s->texels + (size_t)T->offset * 3;
	float u = std_max(0.0f, fabsf(tu)) * T->width;
	float v = std_max(0.0f, fabsf(tv)) * T->height;
	int x = (int)floorf(u);
	int y = (int)floorf(v);
	float frac_u = u - x;
	float frac_v = v - y;
	float w0 = (1.0f - frac_u) * (1.0f - frac_v);
	float w1 = frac_u * (1.0f - frac_v);
	float w2 = (1.0f - frac_u) * frac_v;
	float w3 = frac_u * frac_v;
	v3 s0, s1, s2, s3;
	x = x % T->width;
	y = y % T->height;
	if (x < 0) x = 0; /* inf/NaN coordinates are UB in the reference; stay in bounds */
	if (y < 0) y = 0;
	s0 = Vp(px + 3 * (size_t)(y * T->width + x));
	s1 = Vp(px + 3 * (size_t)(y * T->width + ((x + 1) % T->width)));
	s2 = Vp(px + 3 * (size_t)(((y + 1) % T->height) * T->width + x));
	s3 = Vp(px + 3 * (size_t)(((y + 1) % T->height) * T->width + ((x + 1) % T->width)));
	return add(add(add(scl(s0, w0), scl(s1, w1)), scl(s2, w2)), scl(s3, w3));
}

/* ---------------- SamplingDistributions, Sampling.h:29-70; Core.h:547-550 --------------- */
static v3 spherical_to_world(float theta, float phi)
{
	return V(cosf(phi) * sinf(theta), sinf(phi) * sinf(theta), cosf(theta));
}
static v3 cosine_sample_hemisphere(float r1, float r2)
{
	float theta = acosf(sqrtf(r1));
	float phi = 2.0f * M_PI * r2;
	return spherical_to_world(theta, phi);
}
static float cosine_hemisphere_pdf(v3 wi) { return (wi.z >= 0.0f) ? (wi.z / M_PI) : 0.0f; }
static v3 uniform_sample_sphere(float r1, float r2)
{
	float theta = acosf(1 - 2 * r1);
	float phi = 2.0f * M_PI * r2;
	return spherical_to_world(theta, phi);
}

/* ---------------- BSDFs as Materials.h actually behaves -------------------- */
static v3 bsdf_evaluate(const rtb_scene_desc* s, const rtb_material* m, const shade_t* sd)
{
	v3 a;
	if (m->type == RTB_BSDF_GLASS) return V(0, 0, 0);   /* Materials.h:295-299 */
	a = texture_sample(s, m->tex, sd->tu, sd->tv);
	if (m->type == RTB_BSDF_MIRROR) return a;           /* Materials.h:178-183 */
	return dvd(a, (float)M_PI);                          /* :135-138, 227-231, 344-348, 389-393, 442-446 */
}
static float bsdf_pdf(const rtb_material* m, const shade_t* sd, v3 wi)
{
	if (m->type == RTB_BSDF_GLASS || m->type == RTB_BSDF_MIRROR) return 0.0f; /* :184-188, 300-305 */
	return cosine_hemisphere_pdf(to_local(sd, wi));
}
/* ShadingHelper::fresnelDielectric, Materials.h:55-77 (Fpe denominator as written there) */
static float fresnel_dielectric(float cosTheta, float iorInt, float iorExt, v3* wt, v3 wol)
{
	float ior = iorInt / iorExt;
	float sinTheta_i = sqrtf(1 - (cosTheta * cosTheta));
	float sinTheta_t = ior * sinTheta_i;
	float ior2sin2 = (ior * ior) * (1 - (cosTheta * cosTheta));
	float cosTheta_t, Fpa, Fpe, average;
	if (ior2sin2 > 1.0f) return 1.0f;
	cosTheta_t = sqrtf(1 - (sinTheta_t * sinTheta_t));
	*wt = V(-ior * wol.x, -ior * wol.y, -cosTheta_t);
	Fpa = (cosTheta - ior * cosTheta_t) / (cosTheta + ior * cosTheta_t);
	Fpe = (ior * cosTheta - cosTheta_t) / (ior * cosTheta + ior * cosTheta_t);
	average = ((Fpa * Fpa) + (Fpe * Fpe)) * 0.5f;
	return std_max(0.0f, std_min(1.0f, average));
}
static v3 bsdf_sample(const rtb_scene_desc* s, const rtb_material* m, const shade_t* sd, float r1, float r2, float r3,
                      v3* f, float* pdf)
{
	v3 albedo = texture_sample(s, m->tex, sd->tu, sd->tv);
	if (m->type == RTB_BSDF_MIRROR) /* Materials.h:167-177 */
	{
		v3 wol = to_local(sd, sd->wo);
		*pdf = 1.0f;
		*f = albedo;
		return to_world(sd, V(-wol.x, -wol.y, wol.z));
	}
	if (m->type == RTB_BSDF_GLASS) /* Materials.h:265-294 */
	{
		v3 wol = to_local(sd, sd->wo), wt = V(0, 0, 0), wi;
		float cosTheta_i = fabsf(wol.z);
		int enter = (wol.z > 0.0f);
		float etaI = enter ? m->ext_ior : m->int_ior;
		float etaT = enter ? m->int_ior : m->ext_ior;
		float R = fresnel_dielectric(cosTheta_i, etaI, etaT, &wt, wol);
		if (!enter) wt.z = -wt.z;
		if (R == 1.0f || r3 < R)
		{
			wi = V(-wol.x, -wol.y, wol.z);
			*pdf = R;
			*f = scl(albedo, R);
		}
		else
		{
			wi = wt;
			*pdf = 1.0f - R;
			*f = scl(albedo, (1.0f - R));
		}
		return to_world(sd, wi);
	}
	{
		v3 wl = cosine_sample_hemisphere(r1, r2);
		/* DiffuseBSDF: cosineHemispherePDF (:130); the stubs: wi.z / M_PI (:222,339,384,437) */
		*pdf = (m->type == RTB_BSDF_DIFFUSE) ? cosine_hemisphere_pdf(wl) : (float)(wl.z / M_PI);
		*f = dvd(albedo, (float)M_PI);
		return to_world(sd, wl);
	}
}

/* ---------------- Lights, Lights.h ----------------------------------------------------- */
static v3 env_evaluate(const rtb_scene_desc* s, int tex, v3 wi) /* Lights.h:158-165 */
{
	float u = atan2f(wi.z, wi.x);
	float v;
	u = (u < 0.0f) ? u + (2.0f * M_PI) : u;
	u = u / (2.0f * M_PI);
	v = acosf(wi.y) / M_PI;
	return texture_sample(s, tex, u, v);
}
static v3 background_evaluate(const rtb_scene_desc* s, v3 wi)
{
	if (s->background_type == RTB_LIGHT_ENVMAP) return env_evaluate(s, s->background_tex, wi);
	return Vp(s->background_colour);
}
static v3 triangle_gnormal(const rtb_scene_desc* s, uint32_t id) /* Geometry.h:127-130 */
{
	return scl(Vp(s->tri_isect[id].n), s->tri_shade[id].gsign);
}
static v3 triangle_sample(const rtb_scene_desc* s, uint32_t id, float r1, float r2, float* pdf) /* Geometry.h:114-126 */
{
	const rtb_tri_isect* q = &s->tri_isect[id];
	float alpha = 1 - sqrtf(r1);
	float beta = r2 * sqrtf(r1);
	float gamma = 1.0f - (alpha + beta);
	*pdf = 1.0f / q->area;
	return add(add(scl(Vp(q->v0), alpha), scl(Vp(q->v1), beta)), scl(Vp(q->v2), gamma));
}

/* ---------------- direction sample of a non-area light -----------------------------------
 * STRICT: the reference, uniform sphere with pdf 1/4pi (Lights.h:99-100, 143-149).
 * IMPORTANCE: NOT in the reference (EnvironmentMap::sample is uniform, there is no CDF anywhere);
 * this restates the product's env-map luminance sampler so that the GPU's (wi, pdf, emitted) can be checked
 * <= 1e-5 and its expectation against the reference's uniform sampler.  Tables: per texel cell (x, y) the
 * weight max-luminance-of-the-four-bilinear-taps x sin(theta_centre) + 5 % of the mean (the density is positive
 * wherever Texture::sample (Imaging.h:72-94) can return radiance); marginal CDF over rows, conditional CDF over
 * the columns of each row, both stored as floats and made non-decreasing.  Sample: row by r1, column by r2,
 * uniform inside the cell; (u, v) -> (phi, theta) inverts EnvironmentMap::evaluate's mapping (Lights.h:158-165).
 * The tables are cached per texel pointer (test infrastructure: not re-entrant across scenes being freed). */
typedef struct { const float* texels; int W, H; float* marginal; float* cond; } env_tables_t;
static env_tables_t g_env[8];
static int g_env_n = 0;
static pthread_mutex_t g_env_lock = PTHREAD_MUTEX_INITIALIZER;

static double env_lum_at(const float* texels, int W, int H, int x, int y)
{
	const float* p = texels + ((size_t)(y % H) * W + (size_t)(x % W)) * 3;
	return 0.2126 * p[0] + 0.7152 * p[1] + 0.0722 * p[2];
}
static double dmax(double a, double b) { return a > b ? a : b; }

static const env_tables_t* env_tables(const rtb_scene_desc* s, int tex)
{
	const rtb_texture* T = &s->textures[tex];
	const float* texels = s->texels + (size_t)T->offset * 3;
	int W = T->width, H = T->height, x, y, i;
	env_tables_t* e = NULL;
	double *w, *rowSum, sum = 0.0, mean, floorW, total = 0.0, acc = 0.0;
	pthread_mutex_lock(&g_env_lock);
	for (i = 0; i < g_env_n; i++)
		if (g_env[i].texels == texels && g_env[i].W == W && g_env[i].H == H) e = &g_env[i];
	if (e)
	{
		pthread_mutex_unlock(&g_env_lock);
		return e;
	}
	e = &g_env[g_env_n < 8 ? g_env_n++ : 7];
	if (e->marginal) free(e->marginal), free(e->cond);
	e->texels = texels, e->W = W, e->H = H;
	e->marginal = (float*)calloc((size_t)H + 1, sizeof(float));
	e->cond = (float*)calloc((size_t)H * (W + 1), sizeof(float));
	w = (double*)malloc((size_t)W * H * sizeof(double));
	rowSum = (double*)malloc((size_t)H * sizeof(double));
	for (y = 0; y < H; y++)
	{
		double st = sin(((double)y + 0.5) / (double)H * 3.14159265358979323846);
		for (x = 0; x < W; x++)
		{
			double l = dmax(dmax(env_lum_at(texels, W, H, x, y), env_lum_at(texels, W, H, x + 1, y)),
			                dmax(env_lum_at(texels, W, H, x, y + 1), env_lum_at(texels, W, H, x + 1, y + 1)));
			if (!(l > 0.0)) l = 0.0;
			w[(size_t)y * W + x] = l * st;
			sum += l * st;
		}
	}
	mean = sum / ((double)W * H);
	floorW = (mean > 0.0) ? 0.05 * mean : 1.0;
	for (y = 0; y < H; y++)
	{
		double st = sin(((double)y + 0.5) / (double)H * 3.14159265358979323846), rs = 0.0;
		for (x = 0; x < W; x++)
		{
			w[(size_t)y * W + x] += floorW * st;
			rs += w[(size_t)y * W + x];
		}
		rowSum[y] = rs;
		total += rs;
	}
	for (y = 0; y < H; y++)
	{
		double ca = 0.0;
		float* cd = e->cond + (size_t)y * (W + 1);
		e->marginal[y] = (float)(acc / total);
		acc += rowSum[y];
		for (x = 0; x < W; x++)
		{
			cd[x] = (float)(ca / rowSum[y]);
			ca += w[(size_t)y * W + x];
		}
		cd[W] = 1.0f;
	}
	e->marginal[H] = 1.0f;
	for (y = 1; y <= H; y++)
		if (e->marginal[y] < e->marginal[y - 1]) e->marginal[y] = e->marginal[y - 1];
	for (y = 0; y < H; y++)
	{
		float* cd = e->cond + (size_t)y * (W + 1);
		for (x = 1; x <= W; x++)
			if (cd[x] < cd[x - 1]) cd[x] = cd[x - 1];
	}
	free(w);
	free(rowSum);
	pthread_mutex_unlock(&g_env_lock);
	return e;
}

static int cdf_find(const float* cdf, int n, float r)
{
	int lo = 0, hi = n;
	while (hi - lo > 1)
	{
		int mid = (lo + hi) >> 1;
		if (cdf[mid] <= r) lo = mid;
		else hi = mid;
	}
	return lo;
}

/* returns 0 when the density is not positive */
static int sample_non_area_light(const rtb_scene_desc* s, const rtb_params* P, const rtb_light* L, float r1, float r2, v3* wi, float* pdf,
                                 v3* emitted)
{
	if (L->type == RTB_LIGHT_ENVMAP && P->sampling == RTB_SAMPLING_IMPORTANCE)
	{
		const env_tables_t* e = env_tables(s, L->tex);
		int W = e->W, H = e->H;
		int row = cdf_find(e->marginal, H, r1), col;
		float m0 = e->marginal[row], m1 = e->marginal[row + 1];
		float fr = (m1 > m0) ? (r1 - m0) / (m1 - m0) : 0.5f;
		const float* cd = e->cond + (size_t)row * (W + 1);
		float c0, c1, fc, v, u, theta, phi, st, pmfTexel;
		col = cdf_find(cd, W, r2);
		c0 = cd[col], c1 = cd[col + 1];
		fc = (c1 > c0) ? (r2 - c0) / (c1 - c0) : 0.5f;
		v = ((float)row + fr) / (float)H;
		u = ((float)col + fc) / (float)W;
		theta = v * (float)M_PI, phi = u * (2.0f * (float)M_PI);
		st = sinf(theta);
		*wi = V(cosf(phi) * st, cosf(theta), sinf(phi) * st);
		pmfTexel = (m1 - m0) * (c1 - c0);
		*pdf = pmfTexel * ((float)W * (float)H) / (2.0f * (float)M_PI * (float)M_PI * fmaxf(st, 1e-8f));
		if (!(*pdf > 0.0f)) return 0;
		*emitted = env_evaluate(s, L->tex, *wi);
		return 1;
	}
	*wi = uniform_sample_sphere(r1, r2);
	*pdf = 1.0f / (4.0f * M_PI);
	*emitted = (L->type == RTB_LIGHT_ENVMAP) ? env_evaluate(s, L->tex, *wi) : Vp(L->emission);
	return 1;
}

/* ---------------- counter-based RNG (replaces MTRandom) --------------------------------
 * Philox-4x32-10; counter = (pixel, sample, block, 0), key = (seed, 0x52544232).
 * Vertex at depth k: block 2k = [light pick, light r1, light r2, roulette],
 *                    block 2k+1 = [bsdf r1, bsdf r2, glass draw, -].
 * u = ((x >> 9) + 0.5) * 2^-23, strictly inside (0,1).                                   */
static void philox(uint32_t c[4], uint32_t k0, uint32_t k1)
{
	int i;
	for (i = 0; i < 10; i++)
	{
		uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
		uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
		uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
		c[0] = n0, c[1] = n1, c[2] = n2, c[3] = n3;
		k0 += 0x9E3779B9u;
		k1 += 0xBB67AE85u;
	}
}
static void rng_block(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t block, float u[4])
{
	uint32_t c[4];
	int i;
	c[0] = pixel, c[1] = sample, c[2] = block, c[3] = 0;
	philox(c, seed, 0x52544232u);
	for (i = 0; i < 4; i++) u[i] = ((float)(c[i] >> 9) + 0.5f) * 1.1920928955078125e-7f;
}

/* ---------------- RayTracer::computeDirect, Renderer.h:423-473; Scene.h:131-140 --------- */
static v3 compute_direct(const rtb_scene_desc* s, const rtb_params* P, const shade_t* sd, const float u[4], tally_t* tl)
{
	const rtb_material* m = &s->materials[sd->mat];
	const rtb_light* L;
	float pmf, pdf;
	int li;
	if (m->flags & RTB_MAT_SPECULAR) return V(0, 0, 0);
	if (s->n_lights == 0) return V(0, 0, 0);
	pmf = 1.f / s->n_lights;
	li = (int)(s->n_lights * u[0]);
	if (li > (int)s->n_lights - 1) li = (int)s->n_lights - 1;
	L = &s->lights[li];
	if (L->type == RTB_LIGHT_AREA)
	{
		v3 p = triangle_sample(s, L->triangle, u[1], u[2], &pdf);
		v3 wi = sub(p, sd->x);
		float l = dot3(wi, wi), G;
		wi = norm3(wi);
		G = (win_max(dot3(wi, sd->sN), 0.0f) * win_max(-dot3(wi, triangle_gnormal(s, L->triangle)), 0.0f)) / l;
		if (G > 0)
		{
			tl->shadow++;
			if (scene_visible_tl(s, sd->x, p, P->epsilon, tl))
				return dvd(scl(mul(bsdf_evaluate(s, m, sd), Vp(L->emission)), G), (pmf * pdf));
		}
	}
	else
	{
		v3 wi, emitted;
		int ok = sample_non_area_light(s, P, L, u[1], u[2], &wi, &pdf, &emitted);
		float G = ok ? win_max(dot3(wi, sd->sN), 0.0f) : 0.0f;
		if (G > 0)
		{
			tl->shadow++;
			if (scene_visible_tl(s, sd->x, add(sd->x, scl(wi, 10000.0f)), P->epsilon, tl))
				return dvd(scl(mul(bsdf_evaluate(s, m, sd), emitted), G), (pmf * pdf));
		}
	}
	return V(0, 0, 0);
}

/* ---------------- RayTracer::computeDirectMIS, Renderer.h:474-557 (+ balanceHeuristic :408-410,
 * convertPDFAreaToSolidAngle :411-422).  Unused by the reference's render(); selectable here as
 * RTB_INT_PATH_MIS.  Its quirks are kept: a visible NON-area light returns at once without a weight
 * and without the BSDF strategy; the BSDF strategy converts the area pdf of the light that was
 * SAMPLED (not of the emitter that was hit).  um = the BSDF strategy's own uniforms.           */
static float balance_heuristic(float a, float b) { return a / (a + b); }
static float pdf_area_to_solid_angle(float pdfArea, float dist2, float costheta)
{
	return (costheta > 0.0f) ? (pdfArea * dist2 / costheta) : 0.0f;
}

static v3 compute_direct_mis(const rtb_scene_desc* s, const rtb_params* P, const shade_t* sd, const float u[4], const float um[4],
                             tally_t* tl)
{
	const rtb_material* m = &s->materials[sd->mat];
	const rtb_light* L;
	float pmf, pdf, pdf_bsdf;
	int li;
	v3 result = V(0, 0, 0), val_bsdf, wi_bsdf;
	ray_t r;
	rtb_hit h;
	shade_t sh;
	if (m->flags & RTB_MAT_SPECULAR) return V(0, 0, 0);
	if (s->n_lights == 0) return V(0, 0, 0);
	pmf = 1.f / s->n_lights;
	li = (int)(s->n_lights * u[0]);
	if (li > (int)s->n_lights - 1) li = (int)s->n_lights - 1;
	L = &s->lights[li];
	if (L->type == RTB_LIGHT_AREA)
	{
		v3 p = triangle_sample(s, L->triangle, u[1], u[2], &pdf);
		v3 wi = sub(p, sd->x);
		float l = dot3(wi, wi), cs, cl, G;
		wi = norm3(wi);
		cs = win_max(dot3(wi, sd->sN), 0.0f);
		cl = win_max(-dot3(wi, triangle_gnormal(s, L->triangle)), 0.0f);
		G = cs * cl / l;
		if (G > 0)
		{
			tl->shadow++;
			if (scene_visible_tl(s, sd->x, p, P->epsilon, tl))
			{
				float pdfB = bsdf_pdf(m, sd, wi);
				float pdfL = pdf_area_to_solid_angle(pdf * pmf, l, cl);
				float w = balance_heuristic(pdfL, pdfB);
				result = add(result, dvd(scl(scl(mul(bsdf_evaluate(s, m, sd), Vp(L->emission)), G), w), (pmf * pdf)));
			}
		}
	}
	else
	{
		v3 wi, emitted;
		int ok = sample_non_area_light(s, P, L, u[1], u[2], &wi, &pdf, &emitted);
		float G = ok ? win_max(dot3(wi, sd->sN), 0.0f) : 0.0f;
		if (G > 0)
		{
			tl->shadow++;
			if (scene_visible_tl(s, sd->x, add(sd->x, scl(wi, 10000.0f)), P->epsilon, tl))
				return dvd(scl(mul(bsdf_evaluate(s, m, sd), emitted), G), (pmf * pdf));
		}
	}
	wi_bsdf = bsdf_sample(s, m, sd, um[0], um[1], um[2], &val_bsdf, &pdf_bsdf);
	r = make_ray(add(sd->x, scl(wi_bsdf, P->epsilon)), wi_bsdf);
	h = scene_traverse_tl(s, &r, P->epsilon, tl);
	tl->closest++;
	shading_data(s, &h, &r, &sh);
	if (sh.t < FLT_MAX && (s->materials[sh.mat].flags & RTB_MAT_LIGHT))
	{
		v3 wi = sub(sh.x, sd->x);
		float dist2 = dot3(wi, wi), cl, pdfL, w;
		wi = norm3(wi);
		cl = win_max(0.0f, dot3(neg(wi), sh.sN));
		pdfL = pdf_area_to_solid_angle(pdf * pmf, dist2, cl);
		w = balance_heuristic(pdf_bsdf, pdfL);
		result = add(result, dvd(scl(scl(mul(val_bsdf, Vp(s->materials[sh.mat].emission)), win_max(0.0f, dot3(wi_bsdf, sd->sN))), w), pdf_bsdf));
	}
	return result;
}

/* ---------------- RayTracer::pathTrace, Renderer.h:328-392 (recursive like the reference) */
static v3 path_trace(const rtb_scene_desc* s, const rtb_params* P, ray_t* r, v3* T, int depth, uint32_t pixel,
                     uint32_t sample, int canHitLight, tally_t* tl)
{
	rtb_hit h = scene_traverse_tl(s, r, P->epsilon, tl);
	shade_t sd;
	tl->closest++;
	shading_data(s, &h, r, &sd);
	if (sd.t < FLT_MAX)
	{
		const rtb_material* m = &s->materials[sd.mat];
		float ua[4], ub[4], rr, pdf;
		v3 direct, f, wi;
		if (m->flags & RTB_MAT_LIGHT)
		{
			if (canHitLight) return mul(*T, Vp(m->emission));
			return V(0, 0, 0);
		}
		rng_block(P->seed, pixel, sample, 2u * (uint32_t)depth, ua);
		if (P->integrator == RTB_INT_PATH_MIS)
		{
			float um[4];
			rng_block(P->seed, pixel, sample, RTB_RNG_MIS_BLOCK + (uint32_t)depth, um);
			direct = mul(*T, compute_direct_mis(s, P, &sd, ua, um, tl));
		}
		else
			direct = mul(*T, compute_direct(s, P, &sd, ua, tl));
		if (depth > P->max_depth) return direct;
		rr = win_min(lum3(*T), P->rr_cap);
		if (ua[3] < rr) *T = dvd(*T, rr);
		else return direct;
		rng_block(P->seed, pixel, sample, 2u * (uint32_t)depth + 1u, ub);
		wi = bsdf_sample(s, m, &sd, ub[0], ub[1], ub[2], &f, &pdf);
		if (m->flags & RTB_MAT_SPECULAR) *T = dvd(mul(*T, f), pdf);
		else *T = dvd(scl(mul(*T, f), fabsf(dot3(wi, sd.sN))), pdf);
		*r = make_ray(add(sd.x, scl(wi, P->epsilon)), wi);
		return add(direct, path_trace(s, P, r, T, depth + 1, pixel, sample, (m->flags & RTB_MAT_SPECULAR) != 0, tl));
	}
	return background_evaluate(s, r->d);
}

/* direct(), albedo(), viewNormals(): Renderer.h:393-407, 558-581 */
static v3 shade_simple(const rtb_scene_desc* s, const rtb_params* P, ray_t* r, uint32_t pixel, uint32_t sample, tally_t* tl)
{
	rtb_hit h = scene_traverse_tl(s, r, P->epsilon, tl);
	shade_t sd;
	tl->closest++;
	shading_data(s, &h, r, &sd);
	if (P->integrator == RTB_INT_NORMALS)
	{
		if (h.t < FLT_MAX) return V(fabsf(sd.sN.x), fabsf(sd.sN.y), fabsf(sd.sN.z));
		return V(0, 0, 0);
	}
	if (sd.t < FLT_MAX)
	{
		const rtb_material* m = &s->materials[sd.mat];
		float ua[4];
		if (m->flags & RTB_MAT_LIGHT) return Vp(m->emission);
		if (P->integrator == RTB_INT_ALBEDO) return bsdf_evaluate(s, m, &sd);
		rng_block(P->seed, pixel, sample, 0, ua);
		return compute_direct(s, P, &sd, ua, tl);
	}
	if (P->integrator == RTB_INT_ALBEDO) return background_evaluate(s, r->d);
	return V(0, 0, 0);
}

/* ---------------- render: renderTile + Film::splat(BoxFilter), Renderer.h:795-818 -------- */
typedef struct {
	const rtb_scene_desc* s;
	const rtb_params* P;
	uint32_t spp_begin, spp_count;
	float* film;
	int y0, y1;
	tally_t tl;
} job_t;

/* oracle_render_adaptive: per 32x32 tile sample counts replacing spp_count (not re-entrant) */
static const uint32_t* g_tile_count = NULL;
static uint32_t g_tile_tx = 0;

static int pixel_owned(const rtb_params* P, uint32_t W, uint32_t x, uint32_t y)
{
	if (P->partition == RTB_PART_TILE && P->part_world > 1)
	{
		uint32_t t32x = (W + 31u) >> 5;
		uint32_t tile = (y >> 5) * t32x + (x >> 5);
		return (int)(tile % (uint32_t)P->part_world) == P->part_rank;
	}
	return 1;
}

static void* render_rows(void* arg)
{
	job_t* j = (job_t*)arg;
	const rtb_scene_desc* s = j->s;
	const rtb_params* P = j->P;
	uint32_t W = (uint32_t)s->camera.width;
	int y;
	for (y = j->y0; y < j->y1; y++)
	{
		uint32_t x;
		for (x = 0; x < W; x++)
		{
			uint32_t pixel = (uint32_t)y * W + x, smp;
			v3 acc = V(0, 0, 0);
			float* f = j->film + (size_t)pixel * 3;
			if (!pixel_owned(P, W, x, (uint32_t)y)) continue;
			uint32_t count = g_tile_count ? g_tile_count[((uint32_t)y >> 5) * g_tile_tx + (x >> 5)] : j->spp_count;
			for (smp = j->spp_begin; smp < j->spp_begin + count; smp++)
			{
				ray_t r;
				v3 c;
				if (P->partition == RTB_PART_SPP && P->part_world > 1 &&
				    (int)(smp % (uint32_t)P->part_world) != P->part_rank)
					continue;
				r = generate_ray(&s->camera, x + 0.5f, y + 0.5f);
				if (P->integrator == RTB_INT_PATH || P->integrator == RTB_INT_PATH_MIS)
				{
					v3 T = V(1.0f, 1.0f, 1.0f);
					c = path_trace(s, P, &r, &T, 0, pixel, smp, 1, &j->tl);
				}
				else
					c = shade_simple(s, P, &r, pixel, smp, &j->tl);
				acc = add(acc, c);
				j->tl.samples++;
			}
			f[0] += acc.x, f[1] += acc.y, f[2] += acc.z;
		}
	}
	return NULL;
}

typedef struct {
	job_t* jobs;
	int first, step, n;
} span_t;

static void* span_runner(void* arg)
{
	span_t* sp = (span_t*)arg;
	int c;
	for (c = sp->first; c < sp->n; c += sp->step) render_rows(&sp->jobs[c]);
	return NULL;
}

/* spp_count samples per pixel accumulated into film_sum (running sums like Film::film);
 * rows are dealt to `threads` workers in interleaved blocks of 4 (no shared sampler state:
 * the result does not depend on the thread count). */
int oracle_render(const rtb_scene_desc* s, const rtb_params* P, uint32_t spp_begin, uint32_t spp_count, int threads,
                  float* film_sum, uint64_t* stats /* samples, closest, shadow */)
{
	int H = (int)s->camera.height, i, c;
	int nt = threads < 1 ? 1 : (threads > 256 ? 256 : threads);
	int chunks = (H + 3) / 4;
	job_t* jobs = (job_t*)calloc((size_t)chunks, sizeof(job_t));
	pthread_t* th = (pthread_t*)calloc((size_t)nt, sizeof(pthread_t));
	span_t* spans = (span_t*)calloc((size_t)nt, sizeof(span_t));
	for (c = 0; c < chunks; c++)
	{
		jobs[c].s = s, jobs[c].P = P, jobs[c].spp_begin = spp_begin, jobs[c].spp_count = spp_count;
		jobs[c].film = film_sum, jobs[c].y0 = c * 4, jobs[c].y1 = (c * 4 + 4 < H) ? c * 4 + 4 : H;
	}
	for (i = 0; i < nt; i++)
	{
		spans[i].jobs = jobs, spans[i].first = i, spans[i].step = nt, spans[i].n = chunks;
		if (nt == 1) span_runner(&spans[i]);
		else pthread_create(&th[i], NULL, span_runner, &spans[i]);
	}
	if (nt > 1)
		for (i = 0; i < nt; i++) pthread_join(th[i], NULL);
	if (stats)
	{
		stats[0] = stats[1] = stats[2] = 0;
		if (g_count_canonical) stats[3] = stats[4] = stats[5] = stats[6] = stats[7] = stats[8] = stats[9] = stats[10] = 0;
		for (c = 0; c < chunks; c++)
		{
			stats[0] += jobs[c].tl.samples, stats[1] += jobs[c].tl.closest, stats[2] += jobs[c].tl.shadow;
			if (g_count_canonical)
			{
				stats[3] += jobs[c].tl.cbox, stats[4] += jobs[c].tl.ctri, stats[5] += jobs[c].tl.sbox, stats[6] += jobs[c].tl.stri;
				if (jobs[c].tl.cmax > stats[7]) stats[7] = jobs[c].tl.cmax;
				if (jobs[c].tl.smax > stats[8]) stats[8] = jobs[c].tl.smax;
				stats[9] += jobs[c].tl.cbig, stats[10] += jobs[c].tl.nanrays;
			}
		}
	}
	free(spans);
	free(jobs);
	free(th);
	return 0;
}

/* ---------------- RayTracer::adaptiveRender, Renderer.h:583-749 ---------------------------
 * adaptiveSampling (:583-641): INIT samples per pixel -> per 32x32 tile the variance of the pixel
 * means around the tile mean, ((sum.r + sum.g + sum.b) / 3) / (n - 1), float arithmetic in pixel
 * order like the reference; adaptiveRender (:709-719): weight = variance / total; sampleTileWithWeight
 * (:645-677): samples = max((int)(sqrt(weight) * MAX), MIN) fresh samples per pixel, mean splatted.
 * Sample indices: 0..init-1 steer, init.. are splatted (the reference's per-thread MTRandom simply
 * runs on).  film_sum += one mean image.  tile_samples / tile_variance may be NULL.             */
/* sample_base: first sample index this call draws from.  The reference's MTRandom keeps advancing, so successive
 * adaptiveRender() calls are independent; with a counter-based RNG every call needs its own index range
 * (the product uses base = SPP-before-the-call x (init + max)). */
int oracle_render_adaptive_at(const rtb_scene_desc* s, const rtb_params* P, uint32_t sample_base, uint32_t init_samples, uint32_t min_samples,
                              uint32_t max_samples, int threads, float* film_sum, uint32_t* tile_samples, float* tile_variance);
int oracle_render_adaptive(const rtb_scene_desc* s, const rtb_params* P, uint32_t init_samples, uint32_t min_samples,
                           uint32_t max_samples, int threads, float* film_sum, uint32_t* tile_samples, float* tile_variance)
{
	return oracle_render_adaptive_at(s, P, 0, init_samples, min_samples, max_samples, threads, film_sum, tile_samples, tile_variance);
}
int oracle_render_adaptive_at(const rtb_scene_desc* s, const rtb_params* P, uint32_t sample_base, uint32_t init_samples, uint32_t min_samples,
                              uint32_t max_samples, int threads, float* film_sum, uint32_t* tile_samples, float* tile_variance)
{
	uint32_t W = (uint32_t)s->camera.width, H = (uint32_t)s->camera.height;
	uint32_t tx = (W + 31u) / 32u, ty = (H + 31u) / 32u, nT = tx * ty, t;
	size_t npx = (size_t)W * H, i;
	float* est = (float*)calloc(npx * 3, sizeof(float));
	float* var = (float*)calloc(nT, sizeof(float));
	uint32_t* cnt = (uint32_t*)calloc(nT, sizeof(uint32_t));
	uint32_t maxCount = 0;
	float total = 0.0f;
	uint64_t st[3];
	oracle_render(s, P, sample_base, init_samples, threads, est, st);
	for (t = 0; t < nT; t++)
	{
		uint32_t x0 = (t % tx) * 32u, y0 = (t / tx) * 32u, x1 = x0 + 32u < W ? x0 + 32u : W, y1 = y0 + 32u < H ? y0 + 32u : H, x, y;
		uint32_t n = (x1 - x0) * (y1 - y0);
		v3 sum = V(0, 0, 0), gt, sq = V(0, 0, 0);
		for (y = y0; y < y1; y++)
			for (x = x0; x < x1; x++)
			{
				const float* e = est + ((size_t)y * W + x) * 3;
				sum = add(sum, dvd(V(e[0], e[1], e[2]), (float)init_samples));
			}
		gt = dvd(sum, (float)n);
		for (y = y0; y < y1; y++)
			for (x = x0; x < x1; x++)
			{
				const float* e = est + ((size_t)y * W + x) * 3;
				v3 d = sub(dvd(V(e[0], e[1], e[2]), (float)init_samples), gt);
				sq = add(sq, mul(d, d));
			}
		var[t] = ((sq.x + sq.y + sq.z) / 3.0f) / (float)(n - 1u);
	}
	for (t = 0; t < nT; t++) total += var[t];
	for (t = 0; t < nT; t++)
	{
		float w = (total > 0.0f) ? var[t] / total : 0.0f;
		int sample;
		w = sqrtf(w);
		sample = (int)(w * (float)max_samples);
		cnt[t] = (uint32_t)(sample > (int)min_samples ? sample : (int)min_samples);
		if (cnt[t] > maxCount) maxCount = cnt[t];
	}
	/* fresh samples init .. init + count(tile) - 1 */
	memset(est, 0, npx * 3 * sizeof(float));
	g_tile_count = cnt, g_tile_tx = tx;
	oracle_render(s, P, sample_base + init_samples, maxCount, threads, est, st);
	g_tile_count = NULL;
	for (i = 0; i < npx; i++)
	{
		uint32_t x = (uint32_t)(i % W), y = (uint32_t)(i / W);
		float k = (float)cnt[(y >> 5) * tx + (x >> 5)];
		film_sum[i * 3] += est[i * 3] / k, film_sum[i * 3 + 1] += est[i * 3 + 1] / k, film_sum[i * 3 + 2] += est[i * 3 + 2] / k;
	}
	if (tile_samples) memcpy(tile_samples, cnt, nT * sizeof(uint32_t));
	if (tile_variance) memcpy(tile_variance, var, nT * sizeof(float));
	free(est);
	free(var);
	free(cnt);
	return 0;
}

/* ---------------- RayTracer::lightTracer, Renderer.h:220-326 ------------------------------
 * One pass = width*height light paths (:223-229).  lightTrace_init (:262-288): uniform light pick; only
 * area lights emit paths; position on the triangle, cosine direction about its geometric normal
 * (Lights.h:67-82); Le = evaluate(-wi) * cos / (pmf * pdfDir * pdfPos); the light vertex and every
 * non-specular, non-emitting hit are connected to the camera (connectToCamera :233-260: projection,
 * G term, visibility, W_e = 1 / (Afilm cos^4), box-filter splat at ((int)x, (int)y)); Russian roulette
 * with min(Lum(T), 0.9) and no depth limit (:307-315); T *= f |cos| / pdf (:321).
 * RNG: Philox counter (path, pass, block, 1): block 0 = [pick, pos r1, pos r2, dir r1], block 1 = [dir r2],
 * vertex k: block 2 + k = [roulette, bsdf r1, bsdf r2, bsdf r3].                                     */
static void rng_block_lt(uint32_t seed, uint32_t path, uint32_t pass, uint32_t block, float u[4])
{
	uint32_t c[4];
	int i;
	c[0] = path, c[1] = pass, c[2] = block, c[3] = 1;
	philox(c, seed, 0x52544232u);
	for (i = 0; i < 4; i++) u[i] = ((float)(c[i] >> 9) + 0.5f) * 1.1920928955078125e-7f;
}

static v3 mul_point(const float* m, v3 v) /* Core.h:302-309 */
{
	return V((v.x * m[0] + v.y * m[1] + v.z * m[2]) + m[3], (v.x * m[4] + v.y * m[5] + v.z * m[6]) + m[7],
	         (v.x * m[8] + v.y * m[9] + v.z * m[10]) + m[11]);
}

static void connect_to_camera(const rtb_scene_desc* s, const rtb_params* P, const rtb_camera_ext* ce, v3 p, v3 n, v3 col, float* film,
                              tally_t* tl)
{
	/* Camera::projectOntoCamera, Scene.h:55-69 */
	v3 pv = mul_point(ce->world_to_cam, p), v1 = mul_point(ce->proj, pv), dir;
	const float* m = ce->proj;
	float w = (m[12] * pv.x) + (m[13] * pv.y) + (m[14] * pv.z) + m[15], x, y, dist2, cs, cc, G, We;
	int px, py;
	w = 1.0f / w;
	v1 = scl(v1, w);
	x = (v1.x + 1.0f) * 0.5f;
	y = (v1.y + 1.0f) * 0.5f;
	if (x < 0 || x > 1.0f || y < 0 || y > 1.0f) return;
	x = x * s->camera.width;
	y = 1.0f - y;
	y = y * s->camera.height;
	dir = sub(Vp(s->camera.origin), p);
	dist2 = dot3(dir, dir);
	dir = norm3(dir);
	cs = dot3(n, dir);
	cc = dot3(Vp(ce->view_dir), neg(dir));
	if (cs < 0.0f || cc < 0.0f) return;
	G = (cs * cc) / dist2;
	tl->shadow++;
	if (!scene_visible_tl(s, p, Vp(s->camera.origin), P->epsilon, tl)) return;
	We = 1 / (ce->afilm * ((cc * cc) * (cc * cc)));
	col = scl(scl(col, We), G);
	px = (int)x, py = (int)y;
	if (px >= 0 && px < (int)s->camera.width && py >= 0 && py < (int)s->camera.height)
	{
		float* f = film + ((size_t)py * (size_t)s->camera.width + px) * 3;
		f[0] += col.x, f[1] += col.y, f[2] += col.z;
	}
}

int oracle_render_light(const rtb_scene_desc* s, const rtb_params* P, uint32_t pass_begin, uint32_t pass_count, float* film_sum,
                        uint64_t* stats /* paths, closest, shadow (camera connections) */)
{
	rtb_camera_ext ce;
	uint32_t W = (uint32_t)s->camera.width, H = (uint32_t)s->camera.height, pass, path;
	tally_t tl;
	memset(&tl, 0, sizeof(tl));
	if (!rtb_camera_derive(&s->camera, &ce)) return -1;
	for (pass = pass_begin; pass < pass_begin + pass_count; pass++)
		for (path = 0; path < W * H; path++)
		{
			float u0[4], u1[4], pdfPos, pdfDir, pmf, cosTheta;
			const rtb_light* L;
			v3 p, wl, wi, nL, Le, T = V(1.0f, 1.0f, 1.0f), fu, fv, fw;
			ray_t r;
			int li, k;
			tl.samples++;
			if (s->n_lights == 0) continue;
			rng_block_lt(P->seed, path, pass, 0, u0);
			rng_block_lt(P->seed, path, pass, 1, u1);
			pmf = 1.f / s->n_lights;
			li = (int)(s->n_lights * u0[0]);
			if (li > (int)s->n_lights - 1) li = (int)s->n_lights - 1;
			L = &s->lights[li];
			if (L->type != RTB_LIGHT_AREA) continue;
			p = triangle_sample(s, L->triangle, u0[1], u0[2], &pdfPos);
			wl = cosine_sample_hemisphere(u0[3], u1[0]);
			pdfDir = cosine_hemisphere_pdf(wl);
			nL = triangle_gnormal(s, L->triangle);
			frame_from_vector(nL, &fu, &fv, &fw);
			wi = add(add(scl(fu, wl.x), scl(fv, wl.y)), scl(fw, wl.z));
			cosTheta = dot3(nL, wi);
			/* AreaLight::evaluate(-wi): emission if dot(-wi, gNormal) < 0 (Lights.h:41-48) */
			Le = (dot3(neg(wi), nL) < 0) ? Vp(L->emission) : V(0, 0, 0);
			Le = dvd(scl(Le, cosTheta), (pmf * pdfDir * pdfPos));
			connect_to_camera(s, P, &ce, p, nL, Le, film_sum, &tl);
			r = make_ray(p, wi);
			for (k = 0; k < 100000; k++)
			{
				rtb_hit h = scene_traverse_tl(s, &r, P->epsilon, &tl);
				shade_t sd;
				const rtb_material* m;
				float uk[4], rr, pdf;
				v3 f, wi2;
				tl.closest++;
				shading_data(s, &h, &r, &sd);
				if (!(sd.t < FLT_MAX)) break;
				m = &s->materials[sd.mat];
				if ((m->flags & RTB_MAT_LIGHT) || (m->flags & RTB_MAT_SPECULAR)) break;
				connect_to_camera(s, P, &ce, sd.x, sd.sN, mul(mul(T, bsdf_evaluate(s, m, &sd)), Le), film_sum, &tl);
				rng_block_lt(P->seed, path, pass, 2u + (uint32_t)k, uk);
				rr = win_min(lum3(T), P->rr_cap);
				if (uk[0] < rr) T = dvd(T, rr);
				else break;
				wi2 = bsdf_sample(s, m, &sd, uk[1], uk[2], uk[3], &f, &pdf);
				T = dvd(scl(mul(T, f), fabsf(dot3(wi2, sd.sN))), pdf);
				r = make_ray(add(sd.x, scl(wi2, P->epsilon)), wi2);
			}
		}
	if (stats) stats[0] = tl.samples, stats[1] = tl.closest, stats[2] = tl.shadow;
	return 0;
}

/* ---------------- RayTracer::instantRadiosity, Renderer.h:82-218 ---------------------------
 * traceVPLs (:159-184): n_paths (MAX_VPL = 50, :24) light paths per pass; a VPL on the light itself
 * (Le = evaluate(-wi) / (pmf pdfPos N)) and, by VPLTracePath (:185-218), one at every hit that is neither an
 * emitter nor specular (Le = T * Le' * f * |cos|, Le' = evaluate(-wi) cos_light / (pmf pdfPos N)); the walk
 * goes on through emitters and specular surfaces alike: Russian roulette min(Lum(T), 0.9), BSDF sample,
 * T *= f |cos| / pdf, no depth limit.  Second pass (:82-101, computeVPLsContribution :124-158): every pixel's
 * primary hit gathers all VPLs: skip dist^2 < 1e-4 and back-facing pairs, G = cos cos / dist^2, visibility,
 * col += Le_vpl * f * G; splat at the pixel.
 * RNG: Philox counter (path, pass, block, 2), same block layout as the light tracer.                   */
typedef struct { v3 x, n, Le; } vpl_t;

static void rng_block_stream(uint32_t seed, uint32_t a, uint32_t b, uint32_t block, uint32_t stream, float u[4])
{
	uint32_t c[4];
	int i;
	c[0] = a, c[1] = b, c[2] = block, c[3] = stream;
	philox(c, seed, 0x52544232u);
	for (i = 0; i < 4; i++) u[i] = ((float)(c[i] >> 9) + 0.5f) * 1.1920928955078125e-7f;
}

static int trace_vpls(const rtb_scene_desc* s, const rtb_params* P, uint32_t pass, uint32_t n_paths, vpl_t* out, int cap, tally_t* tl)
{
	int n = 0;
	uint32_t i;
	for (i = 0; i < n_paths; i++)
	{
		float u0[4], u1[4], pdfPos, pmf;
		const rtb_light* L;
		v3 p, wl, wi, nL, Lev, Le, T = V(1.0f, 1.0f, 1.0f), fu, fv, fw;
		ray_t r;
		int li, k;
		if (s->n_lights == 0) continue;
		rng_block_stream(P->seed, i, pass, 0, 2, u0);
		rng_block_stream(P->seed, i, pass, 1, 2, u1);
		pmf = 1.f / s->n_lights;
		li = (int)(s->n_lights * u0[0]);
		if (li > (int)s->n_lights - 1) li = (int)s->n_lights - 1;
		L = &s->lights[li];
		if (L->type != RTB_LIGHT_AREA) continue;
		p = triangle_sample(s, L->triangle, u0[1], u0[2], &pdfPos);
		wl = cosine_sample_hemisphere(u0[3], u1[0]);
		nL = triangle_gnormal(s, L->triangle);
		frame_from_vector(nL, &fu, &fv, &fw);
		wi = add(add(scl(fu, wl.x), scl(fv, wl.y)), scl(fw, wl.z));
		Lev = (dot3(neg(wi), nL) < 0) ? Vp(L->emission) : V(0, 0, 0);
		if (n < cap) out[n].x = p, out[n].n = nL, out[n].Le = dvd(Lev, (pmf * pdfPos * (float)n_paths)), n++;
		Le = dvd(scl(Lev, dot3(wi, nL)), (pmf * pdfPos * (float)n_paths));
		r = make_ray(p, wi);
		for (k = 0; k < 100000; k++)
		{
			rtb_hit h = scene_traverse_tl(s, &r, P->epsilon, tl);
			shade_t sd;
			const rtb_material* m;
			float uk[4], rr, pdf;
			v3 f, wi2;
			tl->closest++;
			shading_data(s, &h, &r, &sd);
			if (!(sd.t < FLT_MAX)) break;
			m = &s->materials[sd.mat];
			if (!(m->flags & RTB_MAT_LIGHT) && !(m->flags & RTB_MAT_SPECULAR))
			{
				if (n < cap)
				{
					out[n].x = sd.x, out[n].n = sd.sN;
					out[n].Le = scl(mul(mul(T, Le), bsdf_evaluate(s, m, &sd)), fabsf(dot3(neg(r.d), sd.sN)));
					n++;
				}
			}
			rng_block_stream(P->seed, i, pass, 2u + (uint32_t)k, 2, uk);
			rr = win_min(lum3(T), P->rr_cap);
			if (uk[0] < rr) T = dvd(T, rr);
			else break;
			wi2 = bsdf_sample(s, m, &sd, uk[1], uk[2], uk[3], &f, &pdf);
			T = dvd(scl(mul(T, f), fabsf(dot3(wi2, sd.sN))), pdf);
			r = make_ray(add(sd.x, scl(wi2, P->epsilon)), wi2);
		}
	}
	return n;
}

typedef struct {
	const rtb_scene_desc* s;
	const rtb_params* P;
	const vpl_t* vpls;
	int n_vpl, y0, y1;
	float* film;
	tally_t tl;
} ir_job_t;

static void* ir_rows(void* arg)
{
	ir_job_t* j = (ir_job_t*)arg;
	const rtb_scene_desc* s = j->s;
	uint32_t W = (uint32_t)s->camera.width;
	int y, i;
	for (y = j->y0; y < j->y1; y++)
	{
		uint32_t x;
		for (x = 0; x < W; x++)
		{
			ray_t r = generate_ray(&s->camera, x + 0.5f, y + 0.5f);
			rtb_hit h = scene_traverse_tl(s, &r, j->P->epsilon, &j->tl);
			shade_t sd;
			const rtb_material* m;
			v3 col = V(0, 0, 0);
			float* f = j->film + ((size_t)y * W + x) * 3;
			j->tl.closest++;
			j->tl.samples++;
			shading_data(s, &h, &r, &sd);
			if (!(sd.t < FLT_MAX)) continue;
			m = &s->materials[sd.mat];
			if ((m->flags & RTB_MAT_LIGHT) || (m->flags & RTB_MAT_SPECULAR)) continue;
			for (i = 0; i < j->n_vpl; i++)
			{
				const vpl_t* v = &j->vpls[i];
				v3 d = sub(v->x, sd.x);
				float dist2 = dot3(d, d), cv, cx, G;
				if (dist2 < 1e-4f) continue;
				d = norm3(d);
				cv = dot3(v->n, neg(d));
				cx = dot3(sd.sN, d);
				if (cv <= 0.0f || cx <= 0.0f) continue;
				G = (cv * cx) / dist2;
				j->tl.shadow++;
				if (!scene_visible_tl(s, sd.x, v->x, j->P->epsilon, &j->tl)) continue;
				col = add(col, scl(mul(v->Le, bsdf_evaluate(s, m, &sd)), G));
			}
			f[0] += col.x, f[1] += col.y, f[2] += col.z;
		}
	}
	return NULL;
}

int oracle_render_ir(const rtb_scene_desc* s, const rtb_params* P, uint32_t pass_begin, uint32_t pass_count, uint32_t n_paths, int threads,
                     float* film_sum, uint64_t* stats /* pixels, closest, shadow, vpls */)
{
	int H = (int)s->camera.height, nt = threads < 1 ? 1 : (threads > 256 ? 256 : threads), cap = (int)n_paths * 256, i;
	vpl_t* vpls = (vpl_t*)calloc((size_t)cap, sizeof(vpl_t));
	ir_job_t* jobs = (ir_job_t*)calloc((size_t)nt, sizeof(ir_job_t));
	pthread_t* th = (pthread_t*)calloc((size_t)nt, sizeof(pthread_t));
	uint64_t tot[4] = {0, 0, 0, 0};
	uint32_t pass;
	for (pass = pass_begin; pass < pass_begin + pass_count; pass++)
	{
		tally_t tl;
		int n, rows = (H + nt - 1) / nt;
		memset(&tl, 0, sizeof(tl));
		n = trace_vpls(s, P, pass, n_paths, vpls, cap, &tl);
		tot[1] += tl.closest, tot[3] += (uint64_t)n;
		for (i = 0; i < nt; i++)
		{
			memset(&jobs[i], 0, sizeof(ir_job_t));
			jobs[i].s = s, jobs[i].P = P, jobs[i].vpls = vpls, jobs[i].n_vpl = n, jobs[i].film = film_sum;
			jobs[i].y0 = i * rows < H ? i * rows : H, jobs[i].y1 = (i + 1) * rows < H ? (i + 1) * rows : H;
			if (nt == 1) ir_rows(&jobs[i]);
			else pthread_create(&th[i], NULL, ir_rows, &jobs[i]);
		}
		for (i = 0; i < nt; i++)
		{
			if (nt > 1) pthread_join(th[i], NULL);
			tot[0] += jobs[i].tl.samples, tot[1] += jobs[i].tl.closest, tot[2] += jobs[i].tl.shadow;
		}
	}
	if (stats) memcpy(stats, tot, sizeof(tot));
	free(vpls);
	free(jobs);
	free(th);
	return 0;
}

/* include/rtb.h's rtb_camera_derive (what light tracing re-derives from the two camera matrices), exported so that
 * tests can compare it with the reference's own Camera fields. */
int oracle_camera_derive(const rtb_camera* c, rtb_camera_ext* e)
{
	return rtb_camera_derive(c, e);
}

/* oracle_render + the canonical-traversal work of every ray it traced.
 * stats = samples, closest, shadow, closest box tests, closest tri tests, shadow box, shadow tri.
 * Not re-entrant (one process-wide switch): tests/tools/canonical_counts.py is its only caller. */
int oracle_render_counts(const rtb_scene_desc* s, const rtb_params* P, uint32_t spp_begin, uint32_t spp_count, int threads,
                         float* film_sum, uint64_t* stats7 /* 11 entries: ... + max boxes of a closest / shadow ray, rays > 20480 boxes, NaN rays */)
{
	int rc;
	g_count_canonical = 1;
	rc = oracle_render(s, P, spp_begin, spp_count, threads, film_sum, stats7);
	g_count_canonical = 0;
	return rc;
}

/* ---------------- batched entry points mirroring include/rtb.h -------------------------- */
int oracle_primary_hits(const rtb_scene_desc* s, float eps, uint32_t* ids, float* t, rtb_ray* rays)
{
	uint32_t W = (uint32_t)s->camera.width, H = (uint32_t)s->camera.height, x, y;
	for (y = 0; y < H; y++)
		for (x = 0; x < W; x++)
		{
			ray_t r = generate_ray(&s->camera, x + 0.5f, y + 0.5f);
			rtb_hit h = scene_traverse(s, &r, eps);
			size_t i = (size_t)y * W + x;
			if (ids) ids[i] = h.id;
			if (t) t[i] = h.t;
			if (rays)
			{
				rays[i].o[0] = r.o.x, rays[i].o[1] = r.o.y, rays[i].o[2] = r.o.z, rays[i].tmax = FLT_MAX;
				rays[i].d[0] = r.d.x, rays[i].d[1] = r.d.y, rays[i].d[2] = r.d.z, rays[i].pad_ = 0;
			}
		}
	return 0;
}

int oracle_trace(const rtb_scene_desc* s, float eps, int any_hit, const rtb_ray* rays, uint64_t n, rtb_hit* hits)
{
	uint64_t i;
	for (i = 0; i < n; i++)
	{
		ray_t r = make_ray(Vp(rays[i].o), Vp(rays[i].d));
		if (any_hit)
		{
			int vis = s->n_ref_nodes ? bvh_visible(s, 0, &r, eps, rays[i].tmax) : 1;
			memset(&hits[i], 0, sizeof(rtb_hit));
			hits[i].id = vis ? 0u : 1u;
		}
		else
			hits[i] = scene_traverse(s, &r, eps);
	}
	return 0;
}

int oracle_visible(const rtb_scene_desc* s, float eps, const float* p1p2, uint64_t n, uint8_t* out)
{
	uint64_t i;
	for (i = 0; i < n; i++) out[i] = (uint8_t)scene_visible(s, Vp(p1p2 + i * 6), Vp(p1p2 + i * 6 + 3), eps);
	return 0;
}

static void store_shading(const shade_t* sd, rtb_shading* o)
{
	memset(o, 0, sizeof(*o));
	o->x[0] = sd->x.x, o->x[1] = sd->x.y, o->x[2] = sd->x.z;
	o->wo[0] = sd->wo.x, o->wo[1] = sd->wo.y, o->wo[2] = sd->wo.z;
	o->s_normal[0] = sd->sN.x, o->s_normal[1] = sd->sN.y, o->s_normal[2] = sd->sN.z;
	o->g_normal[0] = sd->gN.x, o->g_normal[1] = sd->gN.y, o->g_normal[2] = sd->gN.z;
	o->tu = sd->tu, o->tv = sd->tv;
	o->frame_u[0] = sd->fu.x, o->frame_u[1] = sd->fu.y, o->frame_u[2] = sd->fu.z;
	o->frame_v[0] = sd->fv.x, o->frame_v[1] = sd->fv.y, o->frame_v[2] = sd->fv.z;
	o->frame_w[0] = sd->fw.x, o->frame_w[1] = sd->fw.y, o->frame_w[2] = sd->fw.z;
	o->t = sd->t;
	o->material = sd->mat;
}
static void load_shading(const rtb_shading* o, shade_t* sd)
{
	sd->x = Vp(o->x), sd->wo = Vp(o->wo), sd->sN = Vp(o->s_normal), sd->gN = Vp(o->g_normal);
	sd->tu = o->tu, sd->tv = o->tv;
	sd->fu = Vp(o->frame_u), sd->fv = Vp(o->frame_v), sd->fw = Vp(o->frame_w);
	sd->t = o->t, sd->mat = o->material;
}

int oracle_shading_data(const rtb_scene_desc* s, const rtb_ray* rays, const rtb_hit* hits, uint64_t n, rtb_shading* out)
{
	uint64_t i;
	for (i = 0; i < n; i++)
	{
		ray_t r = make_ray(Vp(rays[i].o), Vp(rays[i].d));
		shade_t sd;
		shading_data(s, &hits[i], &r, &sd);
		store_shading(&sd, &out[i]);
	}
	return 0;
}

int oracle_eval_bsdf(const rtb_scene_desc* s, const rtb_shading* sds, const float* wi, const float* u, uint64_t n,
                     float* eval, float* pdf, float* s_wi, float* s_f, float* s_pdf)
{
	uint64_t i;
	for (i = 0; i < n; i++)
	{
		shade_t sd;
		const rtb_material* m;
		load_shading(&sds[i], &sd);
		if (sd.mat < 0 || (uint32_t)sd.mat >= s->n_materials) return -1;
		m = &s->materials[sd.mat];
		if (eval)
		{
			v3 e = bsdf_evaluate(s, m, &sd);
			eval[i * 3] = e.x, eval[i * 3 + 1] = e.y, eval[i * 3 + 2] = e.z;
		}
		if (pdf) pdf[i] = bsdf_pdf(m, &sd, Vp(wi + i * 3));
		if (s_wi || s_f || s_pdf)
		{
			v3 f, d;
			float p;
			d = bsdf_sample(s, m, &sd, u[i * 3], u[i * 3 + 1], u[i * 3 + 2], &f, &p);
			if (s_wi) s_wi[i * 3] = d.x, s_wi[i * 3 + 1] = d.y, s_wi[i * 3 + 2] = d.z;
			if (s_f) s_f[i * 3] = f.x, s_f[i * 3 + 1] = f.y, s_f[i * 3 + 2] = f.z;
			if (s_pdf) s_pdf[i] = p;
		}
	}
	return 0;
}

/* Light::sample / evaluate; env-map direction samples follow P->sampling (NULL = STRICT, the reference) */
int oracle_eval_light_p(const rtb_scene_desc* s, const rtb_params* P, const int32_t* light, const float* wi, const float* u, uint64_t n,
                        float* p_or_wi, float* emitted, float* pdf, float* eval);
int oracle_eval_light(const rtb_scene_desc* s, const int32_t* light, const float* wi, const float* u, uint64_t n,
                      float* p_or_wi, float* emitted, float* pdf, float* eval)
{
	rtb_params P;
	memset(&P, 0, sizeof(P));
	P.sampling = RTB_SAMPLING_STRICT;
	return oracle_eval_light_p(s, &P, light, wi, u, n, p_or_wi, emitted, pdf, eval);
}
int oracle_eval_light_p(const rtb_scene_desc* s, const rtb_params* P, const int32_t* light, const float* wi, const float* u, uint64_t n,
                        float* p_or_wi, float* emitted, float* pdf, float* eval)
{
	uint64_t i;
	for (i = 0; i < n; i++)
	{
		const rtb_light* L;
		v3 p, e, ev, w = Vp(wi + i * 3);
		float pd;
		if (light[i] < 0 || (uint32_t)light[i] >= s->n_lights) return -1;
		L = &s->lights[light[i]];
		if (L->type == RTB_LIGHT_AREA)
		{
			p = triangle_sample(s, L->triangle, u[i * 2], u[i * 2 + 1], &pd);
			e = Vp(L->emission);
			ev = (dot3(w, triangle_gnormal(s, L->triangle)) < 0) ? Vp(L->emission) : V(0, 0, 0); /* Lights.h:40-47 */
		}
		else
		{
			if (!sample_non_area_light(s, P, L, u[i * 2], u[i * 2 + 1], &p, &pd, &e)) pd = 0.0f, e = V(0, 0, 0);
			ev = (L->type == RTB_LIGHT_ENVMAP) ? env_evaluate(s, L->tex, w) : Vp(L->emission);
		}
		if (p_or_wi) p_or_wi[i * 3] = p.x, p_or_wi[i * 3 + 1] = p.y, p_or_wi[i * 3 + 2] = p.z;
		if (emitted) emitted[i * 3] = e.x, emitted[i * 3 + 1] = e.y, emitted[i * 3 + 2] = e.z;
		if (pdf) pdf[i] = pd;
		if (eval) eval[i * 3] = ev.x, eval[i * 3 + 1] = ev.y, eval[i * 3 + 2] = ev.z;
	}
	return 0;
}

int oracle_rng_draws(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t n, float* out)
{
	uint32_t b, k;
	for (b = 0; b * 4 < n; b++)
	{
		float u[4];
		rng_block(seed, pixel, sample, b, u);
		for (k = 0; k < 4 && b * 4 + k < n; k++) out[b * 4 + k] = u[k];
	}
	return 0;
}

/* Film::tonemap, Imaging.h:233-242 */
int oracle_tonemap(const float* film_sum, uint32_t n_pixels, int spp, float exposure, uint8_t* rgb8)
{
	uint32_t i;
	for (i = 0; i < n_pixels * 3; i++)
	{
		float c = film_sum[i] * exposure / (float)spp;
		rgb8[i] = (uint8_t)std_min(powf(std_max(c, 0.0f), 1.0f / 2.2f) * 255, 255.0f);
	}
	return 0;
}

/* Film::splat with GaussianFilter (Imaging.h:155-187, 209-232), one splat per pixel of an
 * image of per-pixel colours (sample positions are pixel centres, Renderer.h:806-807). */
int oracle_gaussian_splat(int width, int height, float radius, float alpha, const float* colours, float* film_sum)
{
	int size = (int)ceilf(radius), x, y, i, j;
	if (size > 2) size = 2;
	memset(film_sum, 0, (size_t)width * height * 3 * sizeof(float));
	for (y = 0; y < height; y++)
		for (x = 0; x < width; x++)
		{
			float w[25], total = 0;
			int idx[25], used = 0, k;
			const float* L = colours + ((size_t)y * width + x) * 3;
			for (i = -size; i <= size; i++)
				for (j = -size; j <= size; j++)
				{
					int px = x + j, py = y + i;
					if (px >= 0 && px < width && py >= 0 && py < height)
					{
						float gx = expf(-alpha * ((float)j * (float)j)) - expf(-alpha * (radius * radius));
						float gy = expf(-alpha * ((float)i * (float)i)) - expf(-alpha * (radius * radius));
						idx[used] = py * width + px;
						w[used] = gx * gy;
						total += w[used];
						used++;
					}
				}
			for (k = 0; k < used; k++)
			{
				float* f = film_sum + (size_t)idx[k] * 3;
				f[0] = f[0] + (L[0] * w[k] / total);
				f[1] = f[1] + (L[1] * w[k] / total);
				f[2] = f[2] + (L[2] * w[k] / total);
			}
		}
	return 0;
}

/* ---------------- raw primitives for known-answer tests -------------------------------- */
int oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
	uint32_t c[4];
	memcpy(c, ctr, sizeof(c));
	philox(c, key[0], key[1]);
	memcpy(out, c, sizeof(c));
	return 0;
}

/* AABB::rayAABB(const Ray&) on a free-standing box (the RTtest `RayAABB` case, RTtest.cpp:49-60) */
int oracle_ray_aabb(const float bmin[3], const float bmax[3], const float o[3], const float d[3])
{
	rtb_ref_node n;
	ray_t r = make_ray(Vp(o), Vp(d));
	memcpy(n.bmin, bmin, 12);
	memcpy(n.bmax, bmax, 12);
	n.a = n.b = 0;
	return ray_aabb(&n, &r);
}

/* Triangle::init + Triangle::rayIntersect on a free-standing triangle (Geometry.h:72-105);
 * out = t, u, v.  The plane stage is Plane::rayIntersect (RTtest.cpp:21-48). */
int oracle_ray_triangle(const float v0[3], const float v1[3], const float v2[3], const float o[3], const float d[3],
                        float out[3])
{
	rtb_tri_isect q;
	ray_t r = make_ray(Vp(o), Vp(d));
	v3 e1 = sub(Vp(v2), Vp(v1)), e2 = sub(Vp(v0), Vp(v2));
	v3 c = cross3(e1, e2), n = norm3(c);
	memset(&q, 0, sizeof(q));
	memcpy(q.v0, v0, 12), memcpy(q.v1, v1, 12), memcpy(q.v2, v2, 12);
	q.n[0] = n.x, q.n[1] = n.y, q.n[2] = n.z;
	q.area = sqrtf(dot3(c, c)) * 0.5f;
	q.d = dot3(n, Vp(v0));
	q.inv_area = 1.0f / dot3(c, n);
	out[0] = out[1] = out[2] = 0;
	return tri_intersect(&q, &r, &out[0], &out[1], &out[2]);
}
