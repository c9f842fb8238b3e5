// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, not product code.
//
// A thin extern "C" driver around the UNMODIFIED reference sources, which are
// compiled in place from /root/reference by oracle/Makefile into
// oracle/_ref/libmiro_ref*.so.  Nothing here re-implements reference
// behaviour: it only constructs reference objects (Scene, TriangleMesh,
// Triangle, Phong, PointLight, Camera, Photon_map) and calls their own
// methods (TriangleMesh::load, Scene::preCalc, Scene::trace, Camera::eyeRay,
// Scene::raytraceImage, Photon_map::irradiance_estimate), so that
//   * the CPU restatement in oracle/miro_oracle.cpp can be pinned against it,
//   * golden fixtures under tests/golden/ can be generated from it, and
//   * bench.py --impl reference can time the reference's own CPU path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
// legs may load the resulting library.
//
// Geometry reaches the reference through its own OBJ loader
// (TriangleMeshLoad.cpp:64) from a path the caller supplies, so the library
// also works on the GPU box, where /root/reference does not exist and the
// caller regenerates an .obj from a committed fixture.

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <map>
#include <tr1/unordered_map>
#include <string>
#include <iostream>
#include <limits>
#include <algorithm>
#include <omp.h>

// Reach Scene's photon maps / BVH and Photon_map's arrays without touching the sources.
#define protected public
#define private public
#include "Miro.h"
#include "Scene.h"
#include "Texture.h"
#include "Camera.h"
#include "Image.h"
#include "Triangle.h"
#include "TriangleMesh.h"
#include "Sphere.h"
#include "Plane.h"
#include "Phong.h"
#include "PointLight.h"
#include "DirectionalAreaLight.h"
#include "PhotonMap.h"
#include "Ray.h"
#include "BVH.h"
#ifdef STATS
#include "Stats.h"
#endif
#undef protected
#undef private

// g_scene, g_camera and g_image are defined in Scene.cpp / Camera.cpp / Image.cpp.

namespace {
std::vector<Material*> g_materials;
std::tr1::unordered_map<const Object*, int> g_prim_id;   // hit.object -> index in Scene::objects()

void rebuild_prim_ids()
{
    g_prim_id.clear();
    const Objects* objs = g_scene->objects();
    g_prim_id.rehash(2 * objs->size() + 16);
    for (size_t i = 0; i < objs->size(); ++i) g_prim_id[(*objs)[i]] = (int)i;
    // unbounded objects (planes) are numbered after the bounded ones
    for (size_t i = 0; i < g_scene->m_unboundedObjects.size(); ++i) g_prim_id[g_scene->m_unboundedObjects[i]] = (int)(objs->size() + i);
}
}  // namespace

// scene scripts of the reference (C++ linkage), see ref_make_scene
void makeTeapotScene(); void makeBunny1Scene(); void makeBunny20Scene(); void makeCornellScene();
void A1makeSphereScene(); void A1makeTeapotScene();

extern "C" {

int ref_build_flags()
{
    int f = 0;
#ifdef STATS
    f |= 1;
#endif
#ifdef __SSE4_1__
    f |= 2;
#endif
#ifdef OPENMP
    f |= 4;
#endif
    return f;
}

void ref_new_scene()
{
    // The reference never frees scenes (no destructors on the path); neither do we.
    g_scene = new Scene;
    g_camera = new Camera;
    g_image = new Image;
    g_materials.clear();
    g_prim_id.clear();
#ifdef STATS
    Stats::BVH_Nodes = Stats::BVH_LeafNodes = 0;   // the reference only ever increments these (BVH.cpp:64,88)
#endif
}

// Phong(kd, ks, kt, shininess, refractIndex), Phong.cpp:13.  shininess < 0 means "infinity" (the ctor default).
int ref_new_material(const float* kd, const float* ks, const float* kt, float shininess, float refr_index)
{
    if (shininess < 0) shininess = infinity;
    Material* m = new Phong(Vector3(kd[0], kd[1], kd[2]), Vector3(ks[0], ks[1], ks[2]),
                            Vector3(kt[0], kt[1], kt[2]), shininess, refr_index);
    g_materials.push_back(m);
    return (int)g_materials.size() - 1;
}

// The reference's own texture classes (Texture.h) by kind number (include/mirogpu.h: MIROGPU_TEX_*), constructor arguments in tp.
static Texture* make_texture(int kind, const float* tp)
{
    switch (kind) {
    case 1: return new CheckerBoardTexture(Vector3(tp[0], tp[1], tp[2]), Vector3(tp[3], tp[4], tp[5]), tp[6]);
    case 2: return new StoneTexture(tp[0]);
    case 3: return new StemTexture(tp[0]);
    case 4: return new PetalTexture(Vector3(tp[0], tp[1], tp[2]), tp[3]);
    case 5: return new LeafTexture(Vector3(0, 0, 0), Vector3(1, 0, 0), tp[0]);
    case 6: return new FlowerCenterTexture(Vector3(tp[0], tp[1], tp[2]), tp[3]);
    }
    return 0;
}

int ref_new_textured_material(int kind, const float* tp, const float* ks, const float* kt, float shininess, float refr_index)
{
    if (shininess < 0) shininess = infinity;
    Texture* t = make_texture(kind, tp);
    if (!t) return -1;
    g_materials.push_back(new TexturedPhong(t, Vector3(ks[0], ks[1], ks[2]), Vector3(kt[0], kt[1], kt[2]), shininess, refr_index));
    return (int)g_materials.size() - 1;
}

// Texture::lookup2D / lookup3D and bumpHeight2D of the reference's classes at n coordinates (3 floats each: u, v, w)
void ref_texture_lookup(int kind, const float* tp, const float* coords3, long n, float* rgb3, float* bump)
{
    Texture* t = make_texture(kind, tp);
    for (long i = 0; i < n; ++i) {
        const float* c = coords3 + 3 * i;
        Vector3 col = t->GetLookupCoordinates() == UV ? t->lookup2D(tex_coord2d_t(c[0], c[1])) : t->lookup3D(tex_coord3d_t(c[0], c[1], c[2]));
        rgb3[3 * i] = col.x; rgb3[3 * i + 1] = col.y; rgb3[3 * i + 2] = col.z;
        if (bump) bump[i] = t->bumpHeight2D(tex_coord2d_t(c[0], c[1]));
    }
}

// ctm: 16 floats m11..m44 in ROW order (the member order of Matrix4x4.h:21-24), or NULL for identity.
int ref_add_obj(const char* path, const float* ctm, int material)
{
    Matrix4x4 m;
    if (ctm)
        m = Matrix4x4(ctm[0], ctm[1], ctm[2], ctm[3], ctm[4], ctm[5], ctm[6], ctm[7],
                      ctm[8], ctm[9], ctm[10], ctm[11], ctm[12], ctm[13], ctm[14], ctm[15]);
    TriangleMesh* mesh = new TriangleMesh;
    if (!mesh->load(path, m)) return -1;
    // same loop as addMeshTrianglesToScene, assignment2.cpp:449-461
    for (int i = 0; i < mesh->numTris(); ++i) {
        Triangle* t = new Triangle;
        t->setIndex(i);
        t->setMesh(mesh);
        t->setMaterial(g_materials[material]);
        g_scene->addObject(t);
    }
    return mesh->numTris();
}

// One free-standing triangle, as the floor triangles in assignment2.cpp:52-66.
void ref_add_triangle(const float* v9, const float* n9, int material)
{
    TriangleMesh* mesh = new TriangleMesh;
    mesh->createSingleTriangle();
    mesh->setV1(Vector3(v9[0], v9[1], v9[2]));
    mesh->setV2(Vector3(v9[3], v9[4], v9[5]));
    mesh->setV3(Vector3(v9[6], v9[7], v9[8]));
    mesh->setN1(Vector3(n9[0], n9[1], n9[2]));
    mesh->setN2(Vector3(n9[3], n9[4], n9[5]));
    mesh->setN3(Vector3(n9[6], n9[7], n9[8]));
    Triangle* t = new Triangle;
    t->setIndex(0);
    t->setMesh(mesh);
    t->setMaterial(g_materials[material]);
    g_scene->addObject(t);
}

// Reference Sphere / Plane objects (Sphere.h, Plane.h); Scene::addObject files the plane under the unbounded objects.
void ref_add_sphere(const float* center, float radius, int material)
{
    Sphere* sp = new Sphere;
    sp->setCenter(Vector3(center[0], center[1], center[2]));
    sp->setRadius(radius);
    sp->setMaterial(g_materials[material]);
    g_scene->addObject(sp);
}
void ref_add_plane(const float* normal, const float* origin, int material)
{
    Plane* pl = new Plane;
    pl->setNormal(Vector3(normal[0], normal[1], normal[2]));
    pl->setOrigin(Vector3(origin[0], origin[1], origin[2]));
    pl->setMaterial(g_materials[material]);
    g_scene->addObject(pl);
}

void ref_add_point_light(const float* pos, const float* color, float wattage)
{
    PointLight* l = new PointLight;
    l->setPosition(Vector3(pos[0], pos[1], pos[2]));
    l->setColor(Vector3(color[0], color[1], color[2]));
    l->setWattage(wattage);
    g_scene->addLight(l);
}

void ref_add_directional_light(const float* pos, const float* normal, float radius, const float* color, float wattage)
{
    DirectionalAreaLight* l = new DirectionalAreaLight(radius);
    l->setPosition(Vector3(pos[0], pos[1], pos[2]));
    l->setNormal(Vector3(normal[0], normal[1], normal[2]));
    l->setColor(Vector3(color[0], color[1], color[2]));
    l->setWattage(wattage);
    g_scene->addLight(l);
}

// The reference's OWN scene scripts (assignment1.cpp / assignment2.cpp, compiled in place, unmodified): each builds g_scene /
// g_camera / g_image, loads its models from paths relative to the working directory (the caller chdir()s to the reference
// tree) and calls Scene::preCalc itself.  Returns 0, or -1 for an unknown name.  Pins tests/../scenes.py against the scripts.
int ref_make_scene(const char* name)
{
    const std::string n(name);
    g_materials.clear(); g_prim_id.clear();
#ifdef STATS
    Stats::BVH_Nodes = Stats::BVH_LeafNodes = 0;
#endif
    if (n == "teapot") makeTeapotScene();
    else if (n == "bunny1") makeBunny1Scene();
    else if (n == "bunny20") makeBunny20Scene();
    else if (n == "cornell") makeCornellScene();
    else if (n == "a1_sphere") A1makeSphereScene();
    else if (n == "a1_teapot") A1makeTeapotScene();
    else return -1;
    rebuild_prim_ids();
    return 0;
}

void ref_set_bg_color(const float* c) { g_scene->setBgColor(Vector3(c[0], c[1], c[2])); }

void ref_srand(unsigned seed) { srand(seed); }

// Scene::preCalc (Scene.cpp:50-84): Object::preCalc, BVH::build, photon passes.
double ref_precalc()
{
    double t = -omp_get_wtime();
    g_scene->preCalc();
    t += omp_get_wtime();
    rebuild_prim_ids();
    return t;
}

int ref_num_objects() { return (int)g_scene->objects()->size(); }

// Per primitive, in Scene::objects() insertion order: A,B,C,nA,nB,nC (18 floats) as the reference's
// loader produced them -- pins the host loader / normal synthesis bit for bit.
void ref_dump_triangles(float* out18)
{
    const Objects* objs = g_scene->objects();
    for (size_t i = 0; i < objs->size(); ++i) {
        Triangle* t = dynamic_cast<Triangle*>((*objs)[i]);
        float* o = out18 + 18 * i;
        if (!t) { for (int k = 0; k < 18; ++k) o[k] = 0; continue; }
        TriangleMesh* m = t->getMesh();
        TriangleMesh::TupleI3 vi = m->vIndices()[t->getIndex()];
        TriangleMesh::TupleI3 ni = m->nIndices()[t->getIndex()];
        for (int k = 0; k < 3; ++k) {
            const Vector3& v = m->vertices()[vi.v[k]];
            const Vector3& n = m->normals()[ni.v[k]];
            o[3 * k + 0] = v.x; o[3 * k + 1] = v.y; o[3 * k + 2] = v.z;
            o[9 + 3 * k + 0] = n.x; o[9 + 3 * k + 1] = n.y; o[9 + 3 * k + 2] = n.z;
        }
    }
}

// rays: n x 8 floats {o.xyz, tMin, d.xyz, tMax}.  Calls Scene::trace (Scene.cpp:214) per ray.
// out_id = index in Scene::objects() of hit.object, -1 on miss.  out_t/out_P/out_N may be NULL.
// N is the normal after Scene::trace's UV-material normalisation (Scene.cpp:262).
void ref_trace(const float* rays, long n, float* out_t, int* out_id, float* out_P, float* out_N, int nthreads)
{
    if (nthreads <= 0) nthreads = omp_get_num_procs();
#pragma omp parallel for schedule(dynamic, 4096) num_threads(nthreads)
    for (long i = 0; i < n; ++i) {
        const float* r = rays + 8 * i;
        Ray ray(Vector3(r[0], r[1], r[2]), Vector3(r[4], r[5], r[6]));
        HitInfo hit;
        bool h = g_scene->trace(hit, ray, r[3], r[7]);
        if (h) {
            std::tr1::unordered_map<const Object*, int>::const_iterator it = g_prim_id.find(hit.object);
            out_id[i] = (it == g_prim_id.end()) ? -2 : it->second;
            if (out_t) out_t[i] = hit.t;
            if (out_P) { out_P[3 * i] = hit.P.x; out_P[3 * i + 1] = hit.P.y; out_P[3 * i + 2] = hit.P.z; }
            if (out_N) { out_N[3 * i] = hit.N.x; out_N[3 * i + 1] = hit.N.y; out_N[3 * i + 2] = hit.N.z; }
        } else {
            out_id[i] = -1;
            if (out_t) out_t[i] = hit.t;   // the reference leaves minHit.t = tMax on a miss (BVH.cpp:444)
            if (out_P) { out_P[3 * i] = out_P[3 * i + 1] = out_P[3 * i + 2] = 0; }
            if (out_N) { out_N[3 * i] = out_N[3 * i + 1] = out_N[3 * i + 2] = 0; }
        }
    }
}

// Same loop, results discarded except a checksum: the timing leg (no output traffic).
double ref_trace_time(const float* rays, long n, int nthreads, long* out_hits)
{
    if (nthreads <= 0) nthreads = omp_get_num_procs();
    long hits = 0;
    double t = -omp_get_wtime();
#pragma omp parallel for schedule(dynamic, 4096) num_threads(nthreads) reduction(+ : hits)
    for (long i = 0; i < n; ++i) {
        const float* r = rays + 8 * i;
        Ray ray(Vector3(r[0], r[1], r[2]), Vector3(r[4], r[5], r[6]));
        HitInfo hit;
        if (g_scene->trace(hit, ray, r[3], r[7])) hits++;
    }
    t += omp_get_wtime();
    if (out_hits) *out_hits = hits;
    return t;
}

// The timing leg that KEEPS its answers (bench.py's parity block compares them with the GPU's hits on the same rays).  Inside the
// timed loop only hit.t and the hit.object pointer are stored; pointers are mapped to prim ids after the clock stops.
double ref_trace_time_hits(const float* rays, long n, int nthreads, float* out_t, int* out_id)
{
    if (nthreads <= 0) nthreads = omp_get_num_procs();
    std::vector<const Object*> obj((size_t)n);
    double t = -omp_get_wtime();
#pragma omp parallel for schedule(dynamic, 4096) num_threads(nthreads)
    for (long i = 0; i < n; ++i) {
        const float* r = rays + 8 * i;
        Ray ray(Vector3(r[0], r[1], r[2]), Vector3(r[4], r[5], r[6]));
        HitInfo hit;
        obj[i] = g_scene->trace(hit, ray, r[3], r[7]) ? hit.object : 0;
        out_t[i] = hit.t;
    }
    t += omp_get_wtime();
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (long i = 0; i < n; ++i) {
        if (!obj[i]) { out_id[i] = -1; continue; }
        std::tr1::unordered_map<const Object*, int>::const_iterator it = g_prim_id.find(obj[i]);
        out_id[i] = (it == g_prim_id.end()) ? -2 : it->second;
    }
    return t;
}

// Host threads the timing legs use when asked for "all" (nthreads <= 0): the processor count, not OMP_NUM_THREADS -- torchrun
// exports OMP_NUM_THREADS=1 to its workers.
int ref_host_threads() { return omp_get_num_procs(); }

// out9: BVH_Nodes, BVH_LeafNodes, Rays, Primary, Secondary, Shadow, Photon_Bounces, Ray_Box, Ray_Tri.
int ref_stats_get(long long* out9)
{
#ifdef STATS
    out9[0] = Stats::BVH_Nodes; out9[1] = Stats::BVH_LeafNodes; out9[2] = Stats::Rays;
    out9[3] = Stats::Primary_Rays; out9[4] = Stats::Secondary_Rays; out9[5] = Stats::Shadow_Rays;
    out9[6] = Stats::Photon_Bounces; out9[7] = Stats::Ray_Box_Intersect; out9[8] = Stats::Ray_Tri_Intersect;
    return 1;
#else
    for (int i = 0; i < 9; ++i) out9[i] = -1;
    return 0;
#endif
}

void ref_stats_reset_rays()
{
#ifdef STATS
    Stats::Rays = Stats::Primary_Rays = Stats::Secondary_Rays = Stats::Shadow_Rays = 0;
    Stats::Ray_Box_Intersect = Stats::Ray_Tri_Intersect = 0;
#endif
}

// Camera (one per process: Camera::eyeRay caches its basis in function statics, Camera.cpp:106-125).
void ref_set_camera(const float* eye, const float* lookat, const float* up, float fov)
{
    g_camera->setEye(Vector3(eye[0], eye[1], eye[2]));
    g_camera->setLookAt(Vector3(lookat[0], lookat[1], lookat[2]));
    g_camera->setUp(Vector3(up[0], up[1], up[2]));
    g_camera->setFOV(fov);
}

// Pixel-centre eye rays for the whole w x h image, row 0 = bottom; rays: n x 8 {o,0,d,MIRO_TMAX}.
void ref_eye_rays(int w, int h, float* rays)
{
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            Ray r = g_camera->eyeRay(x, y, w, h, false);
            float* o = rays + 8 * ((long)y * w + x);
            o[0] = r.o.x; o[1] = r.o.y; o[2] = r.o.z; o[3] = 0.0f;
            o[4] = r.d.x; o[5] = r.d.y; o[6] = r.d.z; o[7] = MIRO_TMAX;
        }
}

// Scene::raytraceImage (Scene.cpp:93) into an 8-bit RGB buffer, row 0 = bottom (Image's own order).
double ref_render(int w, int h, unsigned char* rgb8)
{
    g_image->resize(w, h);
    double t = -omp_get_wtime();
    g_scene->raytraceImage(g_camera, g_image);
    t += omp_get_wtime();
    memcpy(rgb8, g_image->getCharPixels(), (size_t)w * h * 3);
    return t;
}

// Scene::traceScene per ray: the pre-tone-map radiance the render accumulates (Scene.cpp:270).
void ref_trace_scene(const float* rays, long n, int depth, float* rgb, int nthreads)
{
    if (nthreads <= 0) nthreads = omp_get_num_procs();
#pragma omp parallel for schedule(dynamic, 1024) num_threads(nthreads)
    for (long i = 0; i < n; ++i) {
        const float* r = rays + 8 * i;
        Ray ray(Vector3(r[0], r[1], r[2]), Vector3(r[4], r[5], r[6]));
        Vector3 c(0.f);
        g_scene->traceScene(ray, c, depth);
        rgb[3 * i] = c.x; rgb[3 * i + 1] = c.y; rgb[3 * i + 2] = c.z;
    }
}

// ---- photon maps -------------------------------------------------------------------------------
// which: 0 = Scene::m_photonMap, 1 = Scene::m_causticMap, else a standalone map made by ref_pm_new.
static std::vector<Photon_map*> g_pms;

static Photon_map* pm_of(int which)
{
    if (which == 0) return &g_scene->m_photonMap;
    if (which == 1) return &g_scene->m_causticMap;
    return g_pms[which - 2];
}

int ref_pm_new(int max_photons)
{
    g_pms.push_back(new Photon_map(max_photons));
    return (int)g_pms.size() + 1;
}

void ref_pm_store(int which, const float* power, const float* pos, const float* dir, long n)
{
    Photon_map* pm = pm_of(which);
    for (long i = 0; i < n; ++i) pm->store(power + 3 * i, pos + 3 * i, dir + 3 * i);
}

void ref_pm_scale(int which, float s) { pm_of(which)->scale_photon_power(s); }
void ref_pm_balance(int which) { pm_of(which)->balance(); }
int ref_pm_stored(int which) { return pm_of(which)->stored_photons; }
int ref_pm_half_stored(int which) { return pm_of(which)->half_stored_photons; }
int ref_pm_sizeof_photon() { return (int)sizeof(Photon); }

// Copies photons[0..stored] (stored+1 records of 28 B, entry 0 unused) -- the heap-ordered kd-tree.
void ref_pm_dump(int which, void* out)
{
    Photon_map* pm = pm_of(which);
    memcpy(out, pm->photons, sizeof(Photon) * (size_t)(pm->stored_photons + 1));
}

void ref_pm_irradiance(int which, const float* pos, const float* nrm, long n, float max_dist, int k,
                       float* irr, int nthreads)
{
    Photon_map* pm = pm_of(which);
    if (nthreads <= 0) nthreads = omp_get_num_procs();
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads)
    for (long i = 0; i < n; ++i)
        pm->irradiance_estimate(irr + 3 * i, pos + 3 * i, nrm + 3 * i, max_dist, k);
}

}  // extern "C"
