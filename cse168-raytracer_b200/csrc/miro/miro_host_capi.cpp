// miro_host_capi.cpp -- a flat C surface over the host API layer so that Python (tests, bench.py) can drive
// Scene / BVH / Camera / Photon_map exactly as a C++ user of the reference would.  Function names mirror the
// checker drivers (oracle/ref_driver.cpp: ref_*, oracle/miro_oracle.cpp: orc_*) with the prefix mh_.
#include <cstdio>
#include <cstring>
#include <map>
#include <vector>

#include "Miro.h"
#include "Scene.h"
#include "Texture.h"
#include "Camera.h"
#include "Image.h"
#include "Triangle.h"
#include "TriangleMesh.h"
#include "Sphere.h"
#include "Plane.h"
#include "PointLight.h"
#include "PhotonMap.h"

namespace {
std::vector<Material*> g_materials;
std::map<const Object*, int> g_prim_id;
int g_layout = MIROGPU_LAYOUT_QBVH4;
int g_builder = MIROGPU_BUILDER_SAH_HOST;
}

extern "C" {

void mh_new_scene()
{
    delete g_scene; delete g_camera; delete g_image;
    g_scene = new Scene; g_camera = new Camera; g_image = new Image;
    g_scene->PhotonsPerLightSource = g_scene->CausticPhotonsPerLightSource = 0;   // see mh_set_photon_counts
    g_materials.clear(); g_prim_id.clear();
}

void mh_set_layout(int layout) { g_layout = layout; }
void mh_set_builder(int builder) { g_builder = builder; }

int mh_new_material(const float* kd, const float* ks, const float* kt, float shininess, float refr_index)
{
    if (shininess < 0) shininess = infinity;
    g_materials.push_back(new Phong(Vector3(kd[0], kd[1], kd[2]), Vector3(ks[0], ks[1], ks[2]), Vector3(kt[0], kt[1], kt[2]), shininess, refr_index));
    return (int)g_materials.size() - 1;
}

// TexturedPhong(texture, ks, kt, shininess, refractIndex) over one of the procedural textures; tp: the texture's constructor
// arguments in mirogpu_material::tex order (include/mirogpu.h)
int mh_new_textured_material(int kind, const float* tp, const float* ks, const float* kt, float shininess, float refr_index)
{
    if (shininess < 0) shininess = infinity;
    Texture* t = 0;
    switch (kind) {
    case MIROGPU_TEX_CHECKER: t = new CheckerBoardTexture(Vector3(tp[0], tp[1], tp[2]), Vector3(tp[3], tp[4], tp[5]), tp[6]); break;
    case MIROGPU_TEX_STONE: t = new StoneTexture(tp[0]); break;
    case MIROGPU_TEX_STEM: t = new StemTexture(tp[0]); break;
    case MIROGPU_TEX_PETAL: t = new PetalTexture(Vector3(tp[0], tp[1], tp[2]), tp[3]); break;
    case MIROGPU_TEX_LEAF: t = new LeafTexture(Vector3(0, 0, 0), Vector3(1, 0, 0), tp[0]); break;
    case MIROGPU_TEX_FLOWER_CENTER: t = new FlowerCenterTexture(Vector3(tp[0], tp[1], tp[2]), tp[3]); break;
    default: return -1;
    }
    g_materials.push_back(new TexturedPhong(t, Vector3(ks[0], ks[1], ks[2]), Vector3(kt[0], kt[1], kt[2]), shininess, refr_index));
    return (int)g_materials.size() - 1;
}

int mh_add_obj(const char* path, const float* ctm, int material)
{
    Matrix4x4 m;
    if (ctm) m.set(ctm[0], ctm[1], ctm[2], ctm[3], ctm[4], ctm[5], ctm[6], ctm[7], ctm[8], ctm[9], ctm[10], ctm[11], ctm[12], ctm[13], ctm[14], ctm[15]);
    TriangleMesh* mesh = new TriangleMesh;
    if (!mesh->load(path, m)) return -1;
    for (int i = 0; i < mesh->numTris(); ++i) {
        Triangle* t = new Triangle;
        t->setIndex(i); t->setMesh(mesh); t->setMaterial(g_materials[material]);
        g_scene->addObject(t);
    }
    return mesh->numTris();
}

void mh_add_triangle(const float* v9, const float* n9, int material)
{
    TriangleMesh* mesh = new TriangleMesh;
    mesh->createSingleTriangle();
    mesh->setV1(Vector3(v9[0], v9[1], v9[2])); mesh->setV2(Vector3(v9[3], v9[4], v9[5])); mesh->setV3(Vector3(v9[6], v9[7], v9[8]));
    mesh->setN1(Vector3(n9[0], n9[1], n9[2])); mesh->setN2(Vector3(n9[3], n9[4], n9[5])); mesh->setN3(Vector3(n9[6], n9[7], n9[8]));
    Triangle* t = new Triangle;
    t->setIndex(0); t->setMesh(mesh); t->setMaterial(g_materials[material]);
    g_scene->addObject(t);
}

void mh_add_sphere(const float* center, float radius, int material)
{
    Sphere* sp = new Sphere;
    sp->setCenter(Vector3(center[0], center[1], center[2])); sp->setRadius(radius); sp->setMaterial(g_materials[material]);
    g_scene->addObject(sp);
}
void mh_add_plane(const float* normal, const float* origin, int material)
{
    Plane* pl = new Plane;
    pl->setNormal(Vector3(normal[0], normal[1], normal[2])); pl->setOrigin(Vector3(origin[0], origin[1], origin[2])); pl->setMaterial(g_materials[material]);
    g_scene->addObject(pl);
}
// 0 = use MIROGPU_DEVICES / the current device; else the scene is replicated on devices 0 .. n-1 at the next precalc
void mh_set_device_count(int n)
{
    std::vector<int> d;
    for (int k = 0; k < n; ++k) d.push_back(k);
    g_scene->bvh().setDevices(d);
}

void mh_add_point_light(const float* pos, const float* color, float wattage)
{
    PointLight* l = new PointLight;
    l->setPosition(Vector3(pos[0], pos[1], pos[2])); l->setColor(Vector3(color[0], color[1], color[2])); l->setWattage(wattage);
    g_scene->addLight(l);
}

void mh_add_directional_light(const float* pos, const float* normal, float radius, const float* color, float wattage)
{
    DirectionalAreaLight* l = new DirectionalAreaLight(radius);
    l->setPosition(Vector3(pos[0], pos[1], pos[2])); l->setNormal(Vector3(normal[0], normal[1], normal[2]));
    l->setColor(Vector3(color[0], color[1], color[2])); l->setWattage(wattage);
    g_scene->addLight(l);
}

// Photon targets of the next preCalc (Scene.h:67-68).  The driver starts every scene at 0 / 0 so that scenes with a
// DirectionalAreaLight do not trace photons unless a test asks for them.
void mh_set_photon_counts(int global_photons, int caustic_photons)
{
    g_scene->PhotonsPerLightSource = global_photons; g_scene->CausticPhotonsPerLightSource = caustic_photons;
}
long mh_trace_photons(int which) { return which == 0 ? g_scene->tracePhotons() : g_scene->traceCausticPhotons(); }

void mh_set_bg_color(const float* c) { g_scene->setBgColor(Vector3(c[0], c[1], c[2])); }

// Object::preCalc only (bounds), no device work: lets CPU-only tests inspect the ingest.
void mh_precalc_host_only()
{
    const Objects* objs = g_scene->objects();
    for (size_t i = 0; i < objs->size(); ++i) (*objs)[i]->preCalc();
}

double mh_precalc()
{
    g_scene->bvh().setLayout(g_layout); g_scene->bvh().setBuilder(g_builder);
    g_scene->preCalc();
    g_prim_id.clear();
    const Objects* objs = g_scene->objects();
    for (size_t i = 0; i < objs->size(); ++i) g_prim_id[(*objs)[i]] = (int)i;
    const Objects* ub = g_scene->unboundedObjects();   // planes are numbered after the bounded objects (as the checker drivers do)
    for (size_t i = 0; i < ub->size(); ++i) g_prim_id[(*ub)[i]] = (int)(objs->size() + i);
    mirogpu_scene_info info;
    mirogpu_scene_info_get(g_scene->bvh().handle(), &info);
    return info.build_seconds + info.flatten_seconds + info.upload_seconds;
}

int mh_num_objects() { return (int)g_scene->objects()->size(); }

void* mh_scene_handle() { return (void*)g_scene->bvh().handle(); }

void mh_dump_triangles(float* out18)
{
    const Objects* objs = g_scene->objects();
    for (size_t i = 0; i < objs->size(); ++i) {
        Triangle* t = dynamic_cast<Triangle*>((*objs)[i]);
        float* o = out18 + 18 * i;
        if (!t) { memset(o, 0, 18 * sizeof(float)); continue; }
        TriangleMesh* m = t->getMesh();
        const TriangleMesh::TupleI3 vi = m->vIndices()[t->getIndex()], ni = m->nIndices()[t->getIndex()];
        for (int k = 0; k < 3; ++k) {
            const Vector3& v = m->vertices()[vi.v[k]]; const Vector3& n = m->normals()[ni.v[k]];
            o[3 * k] = v.x; o[3 * k + 1] = v.y; o[3 * k + 2] = v.z;
            o[9 + 3 * k] = n.x; o[9 + 3 * k + 1] = n.y; o[9 + 3 * k + 2] = n.z;
        }
    }
}

// Scene::traceBatch -> same outputs as ref_trace / orc_trace.
void mh_trace(const float* rays, long n, float* out_t, int* out_id, float* out_P, float* out_N, int)
{
    std::vector<Ray> rr(n);
    std::vector<HitInfo> hh(n);
    std::vector<char> flags(n);
    // per-ray tMin/tMax are not part of Scene::trace's batch signature: group by the common case, else one by one
    bool uniform = true;
    for (long i = 1; i < n && uniform; ++i) uniform = rays[8 * i + 3] == rays[3] && rays[8 * i + 7] == rays[7];
    for (long i = 0; i < n; ++i) rr[i] = Ray(Vector3(rays[8 * i], rays[8 * i + 1], rays[8 * i + 2]), Vector3(rays[8 * i + 4], rays[8 * i + 5], rays[8 * i + 6]));
    if (uniform && n > 0) g_scene->traceBatch(rr.data(), n, hh.data(), reinterpret_cast<bool*>(flags.data()), rays[3], rays[7]);
    else for (long i = 0; i < n; ++i) flags[i] = g_scene->trace(hh[i], rr[i], rays[8 * i + 3], rays[8 * i + 7]);
    for (long i = 0; i < n; ++i) {
        const HitInfo& h = hh[i];
        if (flags[i]) {
            std::map<const Object*, int>::const_iterator it = g_prim_id.find(h.object);
            out_id[i] = it == g_prim_id.end() ? -2 : it->second;
            if (out_t) out_t[i] = h.t;
            if (out_P) { out_P[3 * i] = h.P.x; out_P[3 * i + 1] = h.P.y; out_P[3 * i + 2] = h.P.z; }
            if (out_N) { out_N[3 * i] = h.N.x; out_N[3 * i + 1] = h.N.y; out_N[3 * i + 2] = h.N.z; }
        } else {
            out_id[i] = -1;
            if (out_t) out_t[i] = h.t;
            if (out_P) out_P[3 * i] = out_P[3 * i + 1] = out_P[3 * i + 2] = 0;
            if (out_N) out_N[3 * i] = out_N[3 * i + 1] = out_N[3 * i + 2] = 0;
        }
    }
}

void mh_set_camera(const float* eye, const float* lookat, const float* up, float fov)
{
    g_camera->setEye(Vector3(eye[0], eye[1], eye[2]));
    g_camera->setLookAt(Vector3(lookat[0], lookat[1], lookat[2]));
    g_camera->setUp(Vector3(up[0], up[1], up[2]));
    g_camera->setFOV(fov);
}

void mh_get_camera(mirogpu_camera* out) { *out = g_camera->abi(); }

void mh_eye_rays(int w, int h, float* rays)
{
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const Ray r = g_camera->eyeRay(x, y, w, h, false);
            float* o = rays + 8 * ((long)y * w + x);
            o[0] = r.o.x; o[1] = r.o.y; o[2] = r.o.z; o[3] = 0.0f; o[4] = r.d.x; o[5] = r.d.y; o[6] = r.d.z; o[7] = MIRO_TMAX;
        }
}

void mh_set_render(int spp, int jitter, int mode, int shadows, unsigned seed, int use_photon_maps)
{
    g_scene->renderSpp = spp; g_scene->renderJitter = jitter; g_scene->renderMode = mode; g_scene->renderShadows = shadows;
    g_scene->renderSeed = seed; g_scene->setPhotonMapsEnabled(use_photon_maps != 0);
}

// Camera::click -> Scene::raytraceImage -> 8-bit image, row 0 = bottom.
double mh_render(int w, int h, unsigned char* rgb8)
{
    g_image->resize(w, h);
    g_camera->click(g_scene, g_image);
    memcpy(rgb8, g_image->getCharPixels(), (size_t)w * h * 3);
    return g_scene->lastRenderSeconds;
}

// ---- photon maps: which 0 / 1 = the scene's global / caustic map -------------------------------------------
static Photon_map* pm_of(int which) { return which == 0 ? &g_scene->photonMap() : &g_scene->causticMap(); }
void mh_pm_store(int which, const float* power, const float* pos, const float* dir, long n)
{
    for (long i = 0; i < n; ++i) pm_of(which)->store(power + 3 * i, pos + 3 * i, dir + 3 * i);
}
void mh_pm_scale(int which, float s) { pm_of(which)->scale_photon_power(s); }
void mh_pm_balance(int which) { pm_of(which)->balance(); }
int mh_pm_stored(int which) { return pm_of(which)->stored(); }
void mh_pm_dump(int which, void* out) { memcpy(out, pm_of(which)->data(), sizeof(Photon) * ((size_t)pm_of(which)->stored() + 1)); }
void mh_pm_attach(int which) { pm_of(which)->attach(g_scene->bvh().handle(), which); }
void mh_pm_irradiance(int which, const float* pos, const float* nrm, long n, float max_dist, int k, float* irr, int)
{
    pm_of(which)->irradiance_estimate_batch(irr, pos, nrm, (size_t)n, max_dist, k);
}

}  // extern "C"
