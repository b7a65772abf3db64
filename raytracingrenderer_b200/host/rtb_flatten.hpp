// rtb_flatten.hpp — host side of the drop-in boundary: turns a live `Scene`
// (RTBase/Scene.h:72-81) into the POD arrays of include/rtb.h.
//
// Duck-typed on purpose: this header never includes a scene API.  Include it AFTER either
//   * the reference's own headers (Scene.h / Materials.h / Lights.h) — then it flattens the
//     reference's objects and the GPU renders exactly what the CPU renderer would, or
//   * raytracingrenderer_b200/host/rtb_scene.hpp — the stand-alone host API with the same
//     class names.
// Needs RTTI (dynamic_cast on BSDF* / Light*), like any code that must recover the class
// of a `BSDF*` stored in Scene::materials.
//
// Build host code that includes this with -ffp-contract=off: inv_area must be computed with
// the reference's expression and rounding (RTBase/Geometry.h:98).
#pragma once

#include "../../include/rtb.h"

#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace rtb
{

struct FlatScene
{
	rtb_camera camera;
	std::vector<rtb_ref_node> ref_nodes;
	std::vector<rtb_tri_isect> tri_isect;
	std::vector<rtb_tri_shade> tri_shade;
	std::vector<rtb_material> materials;
	std::vector<rtb_texture> textures;
	std::vector<float> texels;
	std::vector<rtb_light> lights;
	uint32_t background_type = RTB_LIGHT_BACKGROUND;
	float background_colour[3] = {0, 0, 0};
	int32_t background_tex = -1;

	rtb_scene_desc desc() const
	{
		rtb_scene_desc d;
		memset(&d, 0, sizeof(d));
		d.camera = camera;
		d.ref_nodes = ref_nodes.data();
		d.n_ref_nodes = (uint32_t)ref_nodes.size();
		d.n_tris = (uint32_t)tri_isect.size();
		d.tri_isect = tri_isect.data();
		d.tri_shade = tri_shade.data();
		d.materials = materials.data();
		d.n_materials = (uint32_t)materials.size();
		d.n_textures = (uint32_t)textures.size();
		d.textures = textures.data();
		d.texels = texels.data();
		d.n_texels = texels.size() / 3;
		d.lights = lights.data();
		d.n_lights = (uint32_t)lights.size();
		d.background_type = background_type;
		memcpy(d.background_colour, background_colour, sizeof(background_colour));
		d.background_tex = background_tex;
		return d;
	}

	// ".rtbs" container: fixed header, then the arrays back to back (every array size is a
	// multiple of 16 bytes except texels, which comes last).
	struct FileHeader
	{
		char magic[8]; // "RTBS0001"
		uint32_t n_ref_nodes, n_tris, n_materials, n_textures, n_lights;
		uint32_t background_type;
		int32_t background_tex;
		float background_colour[3];
		uint32_t pad_[2];
		uint64_t n_texels;
		rtb_camera camera;
	};

	bool save(const std::string& path) const
	{
		FILE* f = fopen(path.c_str(), "wb");
		if (!f) return false;
		FileHeader h;
		memset(&h, 0, sizeof(h));
		memcpy(h.magic, "RTBS0001", 8);
		h.n_ref_nodes = (uint32_t)ref_nodes.size();
		h.n_tris = (uint32_t)tri_isect.size();
		h.n_materials = (uint32_t)materials.size();
		h.n_textures = (uint32_t)textures.size();
		h.n_lights = (uint32_t)lights.size();
		h.background_type = background_type;
		h.background_tex = background_tex;
		memcpy(h.background_colour, background_colour, sizeof(background_colour));
		h.n_texels = texels.size() / 3;
		h.camera = camera;
		bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
		auto put = [&](const void* p, size_t bytes) {
			if (bytes && fwrite(p, 1, bytes, f) != bytes) ok = false;
		};
		put(ref_nodes.data(), ref_nodes.size() * sizeof(rtb_ref_node));
		put(tri_isect.data(), tri_isect.size() * sizeof(rtb_tri_isect));
		put(tri_shade.data(), tri_shade.size() * sizeof(rtb_tri_shade));
		put(materials.data(), materials.size() * sizeof(rtb_material));
		put(textures.data(), textures.size() * sizeof(rtb_texture));
		put(lights.data(), lights.size() * sizeof(rtb_light));
		put(texels.data(), texels.size() * sizeof(float));
		if (fclose(f) != 0) ok = false;
		return ok;
	}

	bool load(const std::string& path)
	{
		FILE* f = fopen(path.c_str(), "rb");
		if (!f) return false;
		FileHeader h;
		bool ok = fread(&h, sizeof(h), 1, f) == 1 && memcmp(h.magic, "RTBS0001", 8) == 0;
		auto get = [&](void* p, size_t bytes) {
			if (ok && bytes && fread(p, 1, bytes, f) != bytes) ok = false;
		};
		if (ok)
		{
			ref_nodes.resize(h.n_ref_nodes);
			tri_isect.resize(h.n_tris);
			tri_shade.resize(h.n_tris);
			materials.resize(h.n_materials);
			textures.resize(h.n_textures);
			lights.resize(h.n_lights);
			texels.resize((size_t)h.n_texels * 3);
			background_type = h.background_type;
			background_tex = h.background_tex;
			memcpy(background_colour, h.background_colour, sizeof(background_colour));
			camera = h.camera;
			get(ref_nodes.data(), ref_nodes.size() * sizeof(rtb_ref_node));
			get(tri_isect.data(), tri_isect.size() * sizeof(rtb_tri_isect));
			get(tri_shade.data(), tri_shade.size() * sizeof(rtb_tri_shade));
			get(materials.data(), materials.size() * sizeof(rtb_material));
			get(textures.data(), textures.size() * sizeof(rtb_texture));
			get(lights.data(), lights.size() * sizeof(rtb_light));
			get(texels.data(), texels.size() * sizeof(float));
		}
		fclose(f);
		return ok;
	}
};

namespace detail
{
	template <class TextureT>
	int32_t internTexture(FlatScene& out, std::map<const void*, int32_t>& seen, const TextureT* tex)
	{
		if (!tex) return -1;
		auto it = seen.find((const void*)tex);
		if (it != seen.end()) return it->second;
		rtb_texture t;
		t.offset = (uint32_t)(out.texels.size() / 3);
		t.width = tex->width;
		t.height = tex->height;
		t.pad_ = 0;
		size_t n = (size_t)tex->width * (size_t)tex->height;
		out.texels.reserve(out.texels.size() + n * 3);
		for (size_t i = 0; i < n; i++)
		{
			out.texels.push_back(tex->texels[i].r);
			out.texels.push_back(tex->texels[i].g);
			out.texels.push_back(tex->texels[i].b);
		}
		int32_t id = (int32_t)out.textures.size();
		out.textures.push_back(t);
		seen[(const void*)tex] = id;
		return id;
	}

	// Pre-order flattening of the BVHNode pointer tree (Geometry.h:294-313).
	template <class NodeT>
	int32_t flattenNode(std::vector<rtb_ref_node>& out, const NodeT* node)
	{
		int32_t self = (int32_t)out.size();
		out.push_back(rtb_ref_node());
		rtb_ref_node n;
		n.bmin[0] = node->bounds.min.x;
		n.bmin[1] = node->bounds.min.y;
		n.bmin[2] = node->bounds.min.z;
		n.bmax[0] = node->bounds.max.x;
		n.bmax[1] = node->bounds.max.y;
		n.bmax[2] = node->bounds.max.z;
		if (!node->l && !node->r)
		{
			n.a = ~(int32_t)node->startIndex;
			n.b = (int32_t)(node->endIndex - node->startIndex);
		}
		else
		{
			// buildRecursive (Geometry.h:387-390) always creates both children.
			n.a = flattenNode(out, node->l);
			n.b = flattenNode(out, node->r);
		}
		out[self] = n;
		return self;
	}
} // namespace detail

template <class SceneT>
FlatScene flatten(SceneT& scene)
{
	FlatScene out;
	std::map<const void*, int32_t> texIds;

	// ---- camera (Scene.h:10-41) ----
	memset(&out.camera, 0, sizeof(out.camera));
	memcpy(out.camera.inv_proj, scene.camera.inverseProjectionMatrix.m, 16 * sizeof(float));
	memcpy(out.camera.cam_to_world, scene.camera.camera.m, 16 * sizeof(float));
	out.camera.origin[0] = scene.camera.origin.x;
	out.camera.origin[1] = scene.camera.origin.y;
	out.camera.origin[2] = scene.camera.origin.z;
	out.camera.width = scene.camera.width;
	out.camera.height = scene.camera.height;

	// ---- materials (Materials.h:94-511) ----
	for (size_t i = 0; i < scene.materials.size(); i++)
	{
		BSDF* outer = scene.materials[i];
		BSDF* b = outer;
		rtb_material m;
		memset(&m, 0, sizeof(m));
		m.tex = -1;
		m.int_ior = 1.33f;
		m.ext_ior = 1.0f;
		bool layered = false;
		while (LayeredBSDF* lay = dynamic_cast<LayeredBSDF*>(b))
		{
			layered = true;
			m.thickness = lay->thickness;
			b = lay->base;
		}
		if (DiffuseBSDF* d = dynamic_cast<DiffuseBSDF*>(b))
		{
			m.type = RTB_BSDF_DIFFUSE;
			m.tex = detail::internTexture(out, texIds, d->albedo);
		}
		else if (MirrorBSDF* d = dynamic_cast<MirrorBSDF*>(b))
		{
			m.type = RTB_BSDF_MIRROR;
			m.tex = detail::internTexture(out, texIds, d->albedo);
		}
		else if (ConductorBSDF* d = dynamic_cast<ConductorBSDF*>(b))
		{
			m.type = RTB_BSDF_CONDUCTOR;
			m.tex = detail::internTexture(out, texIds, d->albedo);
			m.alpha = d->alpha;
			m.eta[0] = d->eta.r, m.eta[1] = d->eta.g, m.eta[2] = d->eta.b;
			m.k[0] = d->k.r, m.k[1] = d->k.g, m.k[2] = d->k.b;
		}
		else if (GlassBSDF* d = dynamic_cast<GlassBSDF*>(b))
		{
			m.type = RTB_BSDF_GLASS;
			m.tex = detail::internTexture(out, texIds, d->albedo);
			m.int_ior = d->intIOR;
			m.ext_ior = d->extIOR;
		}
		else if (DielectricBSDF* d = dynamic_cast<DielectricBSDF*>(b))
		{
			m.type = RTB_BSDF_DIELECTRIC;
			m.tex = detail::internTexture(out, texIds, d->albedo);
			m.int_ior = d->intIOR;
			m.ext_ior = d->extIOR;
			m.alpha = d->alpha;
		}
		else if (OrenNayarBSDF* d = dynamic_cast<OrenNayarBSDF*>(b))
		{
			m.type = RTB_BSDF_ORENNAYAR;
			m.tex = detail::internTexture(out, texIds, d->albedo);
			m.alpha = d->sigma;
		}
		else if (PlasticBSDF* d = dynamic_cast<PlasticBSDF*>(b))
		{
			m.type = RTB_BSDF_PLASTIC;
			m.tex = detail::internTexture(out, texIds, d->albedo);
			m.int_ior = d->intIOR;
			m.ext_ior = d->extIOR;
			m.alpha = d->alpha;
		}
		// The virtuals of the OUTER object decide the flags: LayeredBSDF answers
		// isTwoSided() = true even over glass and carries its own (zero) emission.
		m.flags = 0;
		if (outer->isPureSpecular()) m.flags |= RTB_MAT_SPECULAR;
		if (outer->isTwoSided()) m.flags |= RTB_MAT_TWO_SIDED;
		if (outer->isLight()) m.flags |= RTB_MAT_LIGHT;
		if (layered) m.flags |= RTB_MAT_LAYERED;
		m.emission[0] = outer->emission.r;
		m.emission[1] = outer->emission.g;
		m.emission[2] = outer->emission.b;
		out.materials.push_back(m);
	}

	// ---- triangles (Geometry.h:62-131), in Scene::triangles order after build() ----
	out.tri_isect.resize(scene.triangles.size());
	out.tri_shade.resize(scene.triangles.size());
	for (size_t i = 0; i < scene.triangles.size(); i++)
	{
		auto& t = scene.triangles[i];
		rtb_tri_isect& q = out.tri_isect[i];
		for (int k = 0; k < 3; k++)
		{
			q.v0[k] = t.vertices[0].p.coords[k];
			q.v1[k] = t.vertices[1].p.coords[k];
			q.v2[k] = t.vertices[2].p.coords[k];
			q.n[k] = t.n.coords[k];
		}
		q.d = t.d;
		q.inv_area = 1.0f / Dot(t.e1.cross(t.e2), t.n); // Geometry.h:98
		q.material = t.materialIndex;
		q.area = t.area;
		rtb_tri_shade& s = out.tri_shade[i];
		for (int k = 0; k < 3; k++)
		{
			s.n0[k] = t.vertices[0].normal.coords[k];
			s.n1[k] = t.vertices[1].normal.coords[k];
			s.n2[k] = t.vertices[2].normal.coords[k];
		}
		s.u0 = t.vertices[0].u, s.u1 = t.vertices[1].u, s.u2 = t.vertices[2].u;
		s.tv0 = t.vertices[0].v, s.tv1 = t.vertices[1].v, s.tv2 = t.vertices[2].v;
		s.gsign = (Dot(t.vertices[0].normal, t.n) > 0 ? 1.0f : -1.0f); // Geometry.h:129
	}

	// ---- reference BVH ----
	if (scene.bvh) detail::flattenNode(out.ref_nodes, scene.bvh);

	// ---- background + light list (Scene.h:96-105, 155-159) ----
	auto flattenBackground = [&](Light* bg, rtb_light& l) {
		memset(&l, 0, sizeof(l));
		l.tex = -1;
		if (EnvironmentMap* e = dynamic_cast<EnvironmentMap*>(bg))
		{
			l.type = RTB_LIGHT_ENVMAP;
			l.tex = detail::internTexture(out, texIds, e->env);
		}
		else if (BackgroundColour* c = dynamic_cast<BackgroundColour*>(bg))
		{
			l.type = RTB_LIGHT_BACKGROUND;
			l.emission[0] = c->emission.r, l.emission[1] = c->emission.g, l.emission[2] = c->emission.b;
		}
	};
	if (scene.background)
	{
		rtb_light l;
		flattenBackground(scene.background, l);
		out.background_type = l.type;
		out.background_tex = l.tex;
		memcpy(out.background_colour, l.emission, sizeof(l.emission));
	}
	for (size_t i = 0; i < scene.lights.size(); i++)
	{
		Light* L = scene.lights[i];
		rtb_light l;
		memset(&l, 0, sizeof(l));
		l.tex = -1;
		if (AreaLight* a = dynamic_cast<AreaLight*>(L))
		{
			l.type = RTB_LIGHT_AREA;
			l.triangle = (uint32_t)(a->triangle - &scene.triangles[0]);
			l.emission[0] = a->emission.r, l.emission[1] = a->emission.g, l.emission[2] = a->emission.b;
			l.area = a->triangle->area;
		}
		else
		{
			flattenBackground(L, l);
		}
		out.lights.push_back(l);
	}
	return out;
}

} // namespace rtb
