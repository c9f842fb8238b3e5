// miro_host.cpp -- host API layer: the reference's Object / BVH / Scene / Camera / Image / Photon_map
// interfaces implemented over the C ABI in include/mirogpu.h.  Host code keeps the reference's binary32
// operand order wherever a value feeds the device (geometry ingest, camera basis, hit reconstruction), so
// this file is compiled with -ffp-contract=off and no -march flag.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstdarg>
#include <cstring>
#include <map>
#include <omp.h>
#include <vector>

#include "Miro.h"
#include "Vector3.h"
#include "Matrix4x4.h"
#include "Ray.h"
#include "Material.h"
#include "Object.h"
#include "TriangleMesh.h"
#include "Triangle.h"
#include "PointLight.h"
#include "Sphere.h"
#include "Plane.h"
#include "Console.h"
#include "Utility.h"
#include "BVH.h"
#include "Camera.h"
#include "Image.h"
#include "PhotonMap.h"
#include "Scene.h"

Camera* g_camera = 0;
Scene* g_scene = 0;
Image* g_image = 0;

namespace {
[[noreturn]] void die(const char* what)
{
    // the reference's fatal() prints and exits (Console.cpp:119-129); the device layer has no CPU fallback
    fprintf(stderr, "fatal: %s: %s\n", what, mirogpu_last_error());
    exit(-1);
}
double wall()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
}  // namespace

// ================================= TriangleMesh ===========================================================
TriangleMesh::TriangleMesh() : m_normals(0), m_vertices(0), m_normalIndices(0), m_vertexIndices(0), m_texCoords(0), m_texCoordIndices(0),
                               m_numVertices(0), m_numTris(0), m_numTextCoords(0) {}

TriangleMesh::~TriangleMesh()
{
    delete[] m_normals; delete[] m_vertices; delete[] m_normalIndices; delete[] m_vertexIndices; delete[] m_texCoords; delete[] m_texCoordIndices;
}

void TriangleMesh::createSingleTriangle()
{
    m_normals = new Vector3[3];
    m_vertices = new Vector3[3];
    m_normalIndices = new TupleI3[1];
    m_vertexIndices = new TupleI3[1];
    // TriangleMeshLoad.cpp:20-42: the single triangle carries the texture coordinates (0,0) (1,0) (0,1)
    m_texCoords = new VectorR2[3];
    m_texCoords[0].x = 0.0f; m_texCoords[0].y = 0.0f; m_texCoords[1].x = 1.0f; m_texCoords[1].y = 0.0f; m_texCoords[2].x = 0.0f; m_texCoords[2].y = 1.0f;
    m_texCoordIndices = new TupleI3[1];
    for (unsigned k = 0; k < 3; ++k) m_vertexIndices[0].v[k] = m_normalIndices[0].v[k] = m_texCoordIndices[0].v[k] = k;
    m_numVertices = 3;
    m_numTris = 1;
    m_numTextCoords = 3;
}

bool TriangleMesh::load(const char* file, const Matrix4x4& ctm)
{
    FILE* fp = fopen(file, "rb");
    if (!fp) { fprintf(stderr, "error: Cannot open \"%s\" for reading\n", file); return false; }
    loadObj(fp, ctm);
    fclose(fp);
    return true;
}

namespace {
// "v", "v/t", "v//n", "v/t/n" -> (v, t, n), missing parts 0
void splitFaceToken(char* tok, int& v, int& t, int& n)
{
    char* second = 0; char* third = 0;
    for (char* p = tok; *p; ++p)
        if (*p == '/') { *p = 0; if (!second) second = p + 1; else third = p + 1; }
    v = atoi(tok); t = second ? atoi(second) : 0; n = third ? atoi(third) : 0;
}
}  // namespace

// OBJ ingest (TriangleMeshLoad.cpp:25-168), same results as the reference's loader bit for bit, organised for speed: the file
// is read once, cut into the reference's fgets(line, 80) pieces (a longer line continues as a new "line", as there), the
// numbers of all pieces are parsed by all host threads (strtof / atoi: the conversions sscanf itself performs), and only the
// order-dependent assembly -- transforms, index tuples, synthesised normals -- runs sequentially over the parsed records.
namespace {
struct ObjPiece {
    const char* p; int len;
    char kind;            // 'v' vertex, 'n' normal, 't' texture coordinate, 'f' face, 0 anything else
    float f[3];
    int v[3], n[3], t[3];
};
inline void parsePiece(ObjPiece& q)
{
    char line[81];
    memcpy(line, q.p, (size_t)q.len); line[q.len] = 0;
    if (q.kind == 'v' || q.kind == 'n' || q.kind == 't') {
        char* c = line + (q.kind == 'v' ? 1 : 2);
        q.f[0] = q.f[1] = q.f[2] = 0.f;
        for (int k = 0; k < 3; ++k) { char* e = 0; const float x = strtof(c, &e); if (e == c) break; q.f[k] = x; c = e; }
    } else if (q.kind == 'f') {
        char* c = line + 1;
        for (int k = 0; k < 3; ++k) {
            while (*c == ' ' || *c == '\t' || *c == '\n' || *c == '\r' || *c == '\v' || *c == '\f') ++c;
            char* tok = c;
            while (*c && !(*c == ' ' || *c == '\t' || *c == '\n' || *c == '\r' || *c == '\v' || *c == '\f')) ++c;
            if (*c) *c++ = 0;
            splitFaceToken(tok, q.v[k], q.t[k], q.n[k]);
        }
    }
}
}  // namespace

void TriangleMesh::loadObj(FILE* fp, const Matrix4x4& ctm)
{
    static const bool timing = getenv("MIROHOST_TIMING") != 0;
    const double tt0 = wall();
    fseek(fp, 0, SEEK_END);
    const long fsize = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    std::vector<char> text((size_t)std::max(fsize, 0L) + 1);
    const size_t got = fsize > 0 ? fread(text.data(), 1, (size_t)fsize, fp) : 0;
    text[got] = 0;
    // The pieces fgets(line, 80, fp) would return: up to 79 bytes, ending after a newline if one comes first.  A physical line is
    // cut into pieces independently of every other line, so the text is split at newlines into one chunk per thread.
    const int nthreads = std::max(1, omp_get_max_threads());
    std::vector<size_t> bounds((size_t)nthreads + 1, got);
    bounds[0] = 0;
    for (int t = 1; t < nthreads; ++t) {
        size_t s0 = std::max(bounds[(size_t)t - 1], got * (size_t)t / (size_t)nthreads);
        const void* nl = s0 < got ? memchr(text.data() + s0, '\n', got - s0) : 0;
        bounds[(size_t)t] = nl ? (size_t)((const char*)nl - text.data()) + 1 : got;
    }
    std::vector<std::vector<ObjPiece> > local((size_t)nthreads);
#pragma omp parallel num_threads(nthreads)
    {
        const int t = omp_get_thread_num();
        std::vector<ObjPiece>& out = local[(size_t)t];
        const size_t end = bounds[(size_t)t + 1];
        out.reserve((end - bounds[(size_t)t]) / 24 + 16);
        for (size_t s0 = bounds[(size_t)t]; s0 < end;) {
            const size_t lim = std::min(end - s0, (size_t)79);
            const void* nl = memchr(text.data() + s0, '\n', lim);
            const size_t len = nl ? (size_t)((const char*)nl - (text.data() + s0)) + 1 : lim;
            ObjPiece q;
            q.p = text.data() + s0; q.len = (int)len; q.kind = 0;
            const char c0 = q.p[0], c1 = len > 1 ? q.p[1] : 0;
            if (c0 == 'v') q.kind = c1 == 'n' ? 'n' : (c1 == 't' ? 't' : 'v');
            else if (c0 == 'f') q.kind = 'f';
            q.v[0] = q.v[1] = q.v[2] = q.n[0] = q.n[1] = q.n[2] = q.t[0] = q.t[1] = q.t[2] = 0;
            if (q.kind) parsePiece(q);           // numbers: strtof / atoi, the conversions sscanf performs
            out.push_back(q);
            s0 += len;
        }
    }
    std::vector<ObjPiece> pieces;
    {
        size_t total = 0;
        for (int t = 0; t < nthreads; ++t) total += local[(size_t)t].size();
        pieces.reserve(total);
        for (int t = 0; t < nthreads; ++t) pieces.insert(pieces.end(), local[(size_t)t].begin(), local[(size_t)t].end());
    }
    const double tt2 = wall();
    // Where every record lands is a prefix count over the pieces: vertex / normal / texture-coordinate slots, triangle number,
    // and -- for a face whose LAST token carries no normal (the reference's test) -- the three synthesised normal slots, which
    // the reference allocates in file order between the explicit ones.
    const size_t np = pieces.size();
    std::vector<int> slot(np, 0), nslot(np, 0), tbase(np, 0);
    int nv = 0, nf = 0, nt = 0, nnormals = 0, ntouch = 0;
    for (size_t i = 0; i < np; ++i) {
        const ObjPiece& q = pieces[i];
        if (q.kind == 'v') slot[i] = nv++;
        else if (q.kind == 't') slot[i] = nt++;
        else if (q.kind == 'n') slot[i] = nnormals++;
        else if (q.kind == 'f') {
            slot[i] = nf++; nslot[i] = nnormals; tbase[i] = ntouch;
            for (int k = 0; k < 3; ++k) if (q.n[k]) ntouch++;
            if (!q.n[2]) { nnormals += 3; ntouch += 3; }
        }
    }
    const int ncap = std::max(nv, nf * 3);
    m_normals = new Vector3[std::max(ncap, nnormals)];
    m_vertices = new Vector3[nv];
    m_numVertices = nv;
    m_normalIndices = new TupleI3[nf];
    m_vertexIndices = new TupleI3[nf];
    if (nt) {   // TriangleMeshLoad.cpp:152-157
        m_texCoords = new VectorR2[nt];
        m_texCoordIndices = new TupleI3[nf];
        memset(m_texCoordIndices, 0, sizeof(TupleI3) * (size_t)nf);
    }
    m_numTextCoords = nt;
    m_numTris = nf;
    std::vector<char> synthesised((size_t)std::max(ncap, nnormals), 0);
    Matrix4x4 nctm = ctm;
    nctm.invert();
    nctm.transpose();
    // normal slots touching each vertex, in order of appearance: (vertex, slot) pairs now, grouped per vertex below
    std::vector<std::pair<int, int> > touch((size_t)ntouch);
    // pass 1: positions, normals, texture coordinates (independent records)
#pragma omp parallel for schedule(static)
    for (long pi = 0; pi < (long)np; ++pi) {
        const ObjPiece& q = pieces[(size_t)pi];
        if (q.kind == 'n') {
            Vector3& n = m_normals[slot[(size_t)pi]];
            n = nctm * Vector3(q.f[0], q.f[1], q.f[2]);
            n.normalize();
        } else if (q.kind == 'v') m_vertices[slot[(size_t)pi]] = ctm * Vector3(q.f[0], q.f[1], q.f[2]);
        else if (q.kind == 't') { m_texCoords[slot[(size_t)pi]].x = q.f[0]; m_texCoords[slot[(size_t)pi]].y = q.f[1]; }
    }
    // pass 2: faces.  A synthesised normal is cross(e1, e2) of the face's transformed vertices (a well-formed file defines its
    // vertices before the faces that use them, so every position is final here, as it is when the reference reads the face).
#pragma omp parallel for schedule(static)
    for (long pi = 0; pi < (long)np; ++pi) {
        const ObjPiece& q = pieces[(size_t)pi];
        if (q.kind != 'f') continue;
        const int tri = slot[(size_t)pi];
        TupleI3& vi = m_vertexIndices[tri];
        TupleI3& ni = m_normalIndices[tri];
        int tb = tbase[(size_t)pi];
        int n = 0;
        for (int k = 0; k < 3; ++k) {
            n = q.n[k];
            vi.v[k] = q.v[k] - 1;
            if (n) { ni.v[k] = n - 1; touch[(size_t)tb++] = std::make_pair(q.v[k] - 1, n - 1); }
            if (q.t[k] && nt) m_texCoordIndices[tri].v[k] = q.t[k] - 1;
        }
        if (!n) {   // the LAST token decides, as in the reference
            const Vector3 e1 = m_vertices[vi.v[1]] - m_vertices[vi.v[0]];
            const Vector3 e2 = m_vertices[vi.v[2]] - m_vertices[vi.v[0]];
            for (int k = 0; k < 3; ++k) {
                const int sl = nslot[(size_t)pi] + k;
                m_normals[sl] = cross(e1, e2);
                m_normals[sl].normalize();
                synthesised[(size_t)sl] = 1;
                ni.v[k] = sl;
                touch[(size_t)tb++] = std::make_pair((int)vi.v[k], sl);
            }
        }
    }
    const double tt3 = wall();
    // group the touches per vertex, keeping their order (stable counting sort; sequential: two passes over the list)
    std::vector<int> start((size_t)nv + 1, 0), slots(touch.size());
    for (size_t j = 0; j < touch.size(); ++j) if (touch[j].first >= 0 && touch[j].first < nv) start[(size_t)touch[j].first + 1]++;
    for (int i = 0; i < nv; ++i) start[(size_t)i + 1] += start[i];
    {
        std::vector<int> fill(start.begin(), start.end() - 1);
        for (size_t j = 0; j < touch.size(); ++j) if (touch[j].first >= 0 && touch[j].first < nv) slots[(size_t)fill[touch[j].first]++] = touch[j].second;
    }
    // per-vertex averages in order of appearance; a synthesised slot belongs to exactly one vertex, so vertices are independent
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nv; ++i) {
        const int b0 = start[i], e0 = start[(size_t)i + 1];
        if (b0 == e0) continue;
        Vector3 avg;                             // default-constructed: (0,1,2), as the reference accumulates from
        for (int j = b0; j < e0; ++j) avg += m_normals[slots[j]];
        avg /= (float)(e0 - b0);
        avg.normalize();
        for (int j = b0; j < e0; ++j) if (synthesised[slots[j]]) m_normals[slots[j]] = avg;
    }
    if (timing) fprintf(stderr, "loadObj: read + cut + parse %.1f ms, assemble %.1f ms, group + average %.1f ms\n", (tt2 - tt0) * 1e3, (tt3 - tt2) * 1e3,
                        (wall() - tt3) * 1e3);
}

// ================================= Triangle ================================================================
Vector3 Triangle::center() const
{
    const TriangleMesh::TupleI3 ti = m_mesh->vIndices()[m_index];
    const Vector3 A = m_mesh->vertices()[ti.v[0]], B = m_mesh->vertices()[ti.v[1]], C = m_mesh->vertices()[ti.v[2]];
    return A + (B - A) / 3 + (C - A) / 3;
}

void Triangle::preCalc()
{
    const TriangleMesh::TupleI3 ti = m_mesh->vIndices()[m_index];
    m_cachedMin = m_cachedMax = m_mesh->vertices()[ti.v[0]];
    for (int k = 1; k < 3; ++k) {
        const Vector3& p = m_mesh->vertices()[ti.v[k]];
        for (int a = 0; a < 3; ++a) {
            if (p[a] < m_cachedMin[a]) m_cachedMin[a] = p[a];
            if (p[a] > m_cachedMax[a]) m_cachedMax[a] = p[a];
        }
    }
}

void Triangle::fillHit(HitInfo& result, float t, float beta, float gamma) const
{
    const TriangleMesh::TupleI3 ti = m_mesh->vIndices()[m_index];
    const TriangleMesh::TupleI3 ni = m_mesh->nIndices()[m_index];
    const Vector3& A = m_mesh->vertices()[ti.v[0]];
    const Vector3 BmA = m_mesh->vertices()[ti.v[1]] - A, CmA = m_mesh->vertices()[ti.v[2]] - A;
    result.P = A + beta * BmA + gamma * CmA;
    result.t = t;
    result.N = (1 - beta - gamma) * m_mesh->normals()[ni.v[0]] + beta * m_mesh->normals()[ni.v[1]] + gamma * m_mesh->normals()[ni.v[2]];
    result.material = m_material;
}

// Single triangle, single ray on the host: plane / barycentric form with the epsilon slop of the reference.
bool Triangle::intersect(HitInfo& result, const Ray& r, float tMin, float tMax)
{
    const TriangleMesh::TupleI3 ti = m_mesh->vIndices()[m_index];
    const Vector3& A = m_mesh->vertices()[ti.v[0]];
    const Vector3 BmA = m_mesh->vertices()[ti.v[1]] - A, CmA = m_mesh->vertices()[ti.v[2]] - A;
    const Vector3 n = cross(BmA, CmA);
    const Vector3 nd = -r.d, oa = r.o - A;
    const float den = dot(nd, n);
    const float t = dot(oa, n) / den;
    const float beta = dot(nd, cross(oa, CmA)) / den;
    const float gamma = dot(nd, cross(BmA, oa)) / den;
    if (beta < -epsilon || gamma < -epsilon || beta + gamma > 1 + epsilon || t < tMin || t > tMax) return false;
    fillHit(result, t, beta, gamma);
    return true;
}

// ================================= Sphere / Plane ============================================================
// Sphere.cpp:28-69, statement for statement (this file is compiled without FMA contraction).
bool Sphere::intersect(HitInfo& result, const Ray& ray, float tMin, float tMax)
{
    const Vector3 toO = ray.o - m_center;
    const float a = ray.d.length2();
    const float b = dot(2 * ray.d, toO);
    const float c = toO.length2() - m_radius * m_radius;
    const float discrim = b * b - 4.0f * a * c;
    if (discrim < 0) return false;
    const float sqrt_discrim = sqrtf(discrim);
    const float t[2] = {(-b - sqrt_discrim) / (2.0f * a), (-b + sqrt_discrim) / (2.0f * a)};
    if ((t[0] > tMin) && (t[0] < tMax)) result.t = t[0];
    else if ((t[1] > tMin) && (t[1] < tMax)) result.t = t[1];
    else return false;
    fillHit(result, ray, result.t);
    return true;
}
void Sphere::fillHit(HitInfo& result, const Ray& ray, float t) const
{
    result.t = t;
    result.P = ray.o + t * ray.d;
    result.N = (result.P - m_center);
    result.N.normalize();
    result.material = m_material;
}
// Plane.cpp:33-48
bool Plane::intersect(HitInfo& result, const Ray& r, float tMin, float tMax)
{
    const float ndotd = dot(m_normal, r.d);
    if (fabs(ndotd) < 1e-6) return false;
    const float t = dot(m_normal, (m_origin - r.o)) / ndotd;
    if (t < tMin || t > tMax) return false;
    fillHit(result, r, t);
    return true;
}
void Plane::fillHit(HitInfo& result, const Ray& r, float t) const
{
    result.P = r.o + t * r.d;
    result.t = t;
    result.N = m_normal;
    result.material = m_material;
}

// ================================= Console / Utility ==========================================================
#define MIRO_CONSOLE_FN(name, prefix, stream, after)            \
    void name(const char* fmt, ...)                             \
    {                                                           \
        va_list ap; va_start(ap, fmt);                          \
        fputs(prefix, stream); vfprintf(stream, fmt, ap);       \
        va_end(ap); after;                                      \
    }
MIRO_CONSOLE_FN(warning, "warning: ", stderr, (void)0)
MIRO_CONSOLE_FN(error, "error: ", stderr, (void)0)
MIRO_CONSOLE_FN(debug, "", stdout, (void)0)
MIRO_CONSOLE_FN(fatal, "fatal: ", stderr, exit(-1))
#undef MIRO_CONSOLE_FN
double getTime() { return wall(); }

// ================================= BVH =======================================================================
BVH::~BVH() { if (m_handle) mirogpu_scene_destroy(m_handle); }

namespace {
uint32_t materialIndex(const Material* mat, std::map<const Material*, uint32_t>& matIndex, std::vector<mirogpu_material>& mats)
{
    std::map<const Material*, uint32_t>::iterator it = matIndex.find(mat);
    if (it == matIndex.end()) {
        mirogpu_material mm; memset(&mm, 0, sizeof mm);
        const Vector3 kd = mat ? mat->getDiffuse() : Vector3(1.f), ks = mat ? mat->getReflection() : Vector3(0.f),
                      kt = mat ? mat->getRefraction() : Vector3(0.f);
        for (int k = 0; k < 3; ++k) { mm.kd[k] = kd[k]; mm.ks[k] = ks[k]; mm.kt[k] = kt[k]; }
        mm.shininess = mat ? mat->getShininess() : 1.f;
        mm.refract_index = mat ? mat->getRefractionIndex() : 1.f;
        if (mat) mat->describeTexture(mm);                     // TexturedPhong: texture kind + constructor arguments
        it = matIndex.insert(std::make_pair(mat, (uint32_t)mats.size())).first;
        mats.push_back(mm);
    }
    return it->second;
}
}  // namespace

// BVH::build (BVH.h:33): the bounded objects of the list -- triangles and spheres -- become the leaf primitives of the device
// tree, the scene's planes (setUnbounded) ride along as the device's post-walk list; objects of any other class stay on the
// host and are tested after the device query.  Device primitive ids: triangles, then spheres, then planes, each in list order.
void BVH::build(Objects* objs, int)
{
    if (m_handle) { mirogpu_scene_destroy(m_handle); m_handle = 0; }
    m_objects = objs;
    m_prims.clear(); m_other.clear();
    std::map<const Material*, uint32_t> matIndex;
    std::vector<mirogpu_material> mats;
    std::vector<float> verts, norms, uvs;
    bool anyUv = false;
    std::vector<uint32_t> matIds;
    std::vector<Object*> sphereObjs, planeObjs;
    std::vector<mirogpu_sphere> spheres;
    std::vector<mirogpu_plane> planes;
    for (size_t i = 0; i < objs->size(); ++i) {
        Object* o = (*objs)[i];
        if (!o->isBounded()) continue;
        if (Sphere* sp = dynamic_cast<Sphere*>(o)) {
            mirogpu_sphere ms; memset(&ms, 0, sizeof ms);
            const Vector3 c = sp->center();
            ms.center[0] = c.x; ms.center[1] = c.y; ms.center[2] = c.z; ms.radius = sp->radius();
            ms.material_id = materialIndex(sp->getMaterial(), matIndex, mats);
            spheres.push_back(ms); sphereObjs.push_back(o);
            continue;
        }
        Triangle* t = dynamic_cast<Triangle*>(o);
        if (!t) { m_other.push_back(o); continue; }
        TriangleMesh* m = t->getMesh();
        const TriangleMesh::TupleI3 vi = m->vIndices()[t->getIndex()], ni = m->nIndices()[t->getIndex()];
        for (int k = 0; k < 3; ++k) {
            const Vector3& p = m->vertices()[vi.v[k]];
            verts.push_back(p.x); verts.push_back(p.y); verts.push_back(p.z);
        }
        for (int k = 0; k < 3; ++k) {
            const Vector3& q = m->normals()[ni.v[k]];
            norms.push_back(q.x); norms.push_back(q.y); norms.push_back(q.z);
        }
        // Triangle::toUVCoordinates interpolates the corners' texture coordinates; a mesh without them answers (0, 0)
        if (m->numTextCoords() > 0 && m->tIndices()) {
            const TriangleMesh::TupleI3 ti = m->tIndices()[t->getIndex()];
            for (int k = 0; k < 3; ++k) {
                const unsigned idx = ti.v[k] < (unsigned)m->numTextCoords() ? ti.v[k] : 0u;
                uvs.push_back(m->texCoords()[idx].x); uvs.push_back(m->texCoords()[idx].y);
            }
            anyUv = true;
        } else uvs.insert(uvs.end(), 6, 0.f);
        matIds.push_back(materialIndex(t->getMaterial(), matIndex, mats));
        m_prims.push_back(t);
    }
    m_ntris = (uint32_t)m_prims.size();
    if (m_unbounded)
        for (size_t i = 0; i < m_unbounded->size(); ++i)
            if (Plane* pl = dynamic_cast<Plane*>((*m_unbounded)[i])) {
                mirogpu_plane mp; memset(&mp, 0, sizeof mp);
                for (int k = 0; k < 3; ++k) { mp.normal[k] = pl->normal()[k]; mp.origin[k] = pl->origin()[k]; }
                mp.material_id = materialIndex(pl->getMaterial(), matIndex, mats);
                planes.push_back(mp); planeObjs.push_back(pl);
            }
    m_nspheres = (uint32_t)spheres.size();
    m_prims.insert(m_prims.end(), sphereObjs.begin(), sphereObjs.end());
    m_prims.insert(m_prims.end(), planeObjs.begin(), planeObjs.end());
    // devices: setDevices(), else MIROGPU_DEVICES = "all" | "0,1,..." , else the current device
    std::vector<int32_t> devs(m_devices.begin(), m_devices.end());
    if (devs.empty())
        if (const char* e = getenv("MIROGPU_DEVICES")) {
            if (!strcmp(e, "all")) { int n = 0; mirogpu_device_count(&n); for (int k = 0; k < n; ++k) devs.push_back(k); }
            else for (const char* q = e; *q;) { devs.push_back(atoi(q)); while (*q && *q != ',') ++q; if (*q == ',') ++q; }
        }
    mirogpu_build_options opt; opt.layout = m_layout; opt.max_leaf = 0; opt.sah_bins = 32; opt.device = -1; opt.builder = m_builder;
    mirogpu_scene_desc d; memset(&d, 0, sizeof d);
    d.tri_vertices = verts.data(); d.tri_normals = norms.data(); d.tri_material_ids = matIds.data(); d.ntris = m_ntris;
    d.spheres = spheres.empty() ? 0 : spheres.data(); d.nspheres = m_nspheres;
    d.planes = planes.empty() ? 0 : planes.data(); d.nplanes = (uint32_t)planes.size();
    d.materials = mats.empty() ? 0 : mats.data(); d.nmaterials = (uint32_t)mats.size();
    d.devices = devs.empty() ? 0 : devs.data(); d.ndevices = (uint32_t)devs.size();
    d.tri_texcoords = (anyUv && m_ntris) ? uvs.data() : 0;
    const int rc = mirogpu_scene_create_ex(&d, &opt, &m_handle);
    if (rc != MIROGPU_OK) die("BVH::build");
}

bool BVH::onDevice(const Object* o) const
{
    for (size_t i = m_ntris; i < m_prims.size(); ++i) if (m_prims[i] == o) return true;   // spheres and planes (few)
    return dynamic_cast<const Triangle*>(o) != 0;
}

bool BVH::finish(HitInfo& result, const mirogpu_hit& h, const Ray& ray, float tMin, float tMax) const
{
    bool hit = false;
    result.t = tMax;                                  // miss contract of the reference (BVH.cpp:444)
    if (h.prim_id != MIROGPU_MISS) {
        Object* o = m_prims[h.prim_id];
        if (h.prim_id < m_ntris) static_cast<const Triangle*>(o)->fillHit(result, h.t, h.beta, h.gamma);
        else if (h.prim_id < m_ntris + m_nspheres) static_cast<const Sphere*>(o)->fillHit(result, ray, h.t);
        else static_cast<const Plane*>(o)->fillHit(result, ray, h.t);
        result.object = o;
        hit = true;
    }
    for (size_t i = 0; i < m_other.size(); ++i) {     // object classes the device does not know: tested here
        HitInfo tmp;
        if (m_other[i]->intersect(tmp, ray, tMin, result.t) && tmp.t < result.t) { result = tmp; result.object = m_other[i]; hit = true; }
    }
    return hit;
}

bool BVH::intersect(HitInfo& result, const Ray& ray, float tMin, float tMax) const
{
    if (!m_handle) die("BVH::intersect before build");
    mirogpu_ray r = {ray.o.x, ray.o.y, ray.o.z, tMin, ray.d.x, ray.d.y, ray.d.z, tMax};
    mirogpu_hit h;
    if (mirogpu_intersect_batch(m_handle, &r, 1, &h, MIROGPU_CLOSEST_HIT) != MIROGPU_OK) die("BVH::intersect");
    return finish(result, h, ray, tMin, tMax);
}

size_t BVH::intersectBatch(const Ray* rays, size_t n, HitInfo* results, bool* hitFlags, float tMin, float tMax) const
{
    if (!m_handle) die("BVH::intersectBatch before build");
    std::vector<mirogpu_ray> rr(n);
    std::vector<mirogpu_hit> hh(n);
    for (size_t i = 0; i < n; ++i) {
        const Ray& q = rays[i];
        mirogpu_ray r = {q.o.x, q.o.y, q.o.z, tMin, q.d.x, q.d.y, q.d.z, tMax};
        rr[i] = r;
    }
    if (mirogpu_intersect_batch(m_handle, rr.data(), n, hh.data(), MIROGPU_CLOSEST_HIT) != MIROGPU_OK) die("BVH::intersectBatch");
    size_t nh = 0;
    for (size_t i = 0; i < n; ++i) {
        const bool h = finish(results[i], hh[i], rays[i], tMin, tMax);
        if (hitFlags) hitFlags[i] = h;
        nh += h;
    }
    return nh;
}

// ================================= Camera / Image ============================================================
mirogpu_camera Camera::abi() const
{
    mirogpu_camera c;
    for (int k = 0; k < 3; ++k) { c.eye[k] = m_eye[k]; c.up[k] = m_up[k]; c.view_dir[k] = m_viewDir[k]; }
    c.fov_degrees = m_fov;
    return c;
}

Ray Camera::eyeRay(int x, int y, int imageWidth, int imageHeight, bool randomize)
{
    const float HalfDegToRad = DegToRad / 2.0f;
    Vector3 w = -m_viewDir; w.normalize();
    Vector3 u = cross(m_up, w); u.normalize();
    const Vector3 v = cross(w, u);
    const float aspect = (float)imageWidth / (float)imageHeight;
    const float top = tanf(m_fov * HalfDegToRad), right = aspect * top, bottom = -top, left = -right;
    float dx = 0.5f, dy = 0.5f;
    if (randomize) { dx = (float)rand() / (float)RAND_MAX; dy = (float)rand() / (float)RAND_MAX; }
    const float U = left + (right - left) * (((float)x + dx) / (float)imageWidth);
    const float V = bottom + (top - bottom) * (((float)y + dy) / (float)imageHeight);
    Vector3 d = U * u + V * v - w;
    d.normalize();
    return Ray(m_eye, d);
}

void Camera::click(Scene* pScene, Image* pImage)
{
    pImage->clear(bgColor());
    pScene->raytraceImage(this, pImage);
}

void Image::release()
{
    if (m_pinned) mirogpu_host_free(m_pixels); else delete[] m_pixels;
    m_pixels = 0; m_pinned = false;
}
void Image::resize(int width, int height)
{
    if (m_pixels && width == m_width && height == m_height) return;   // Camera::click on the same Image, frame after frame
    release();
    const size_t n = (size_t)std::max(width, 0) * std::max(height, 0);
    m_pixels = static_cast<Pixel*>(mirogpu_host_alloc(n * sizeof(Pixel)));
    m_pinned = m_pixels != 0;
    if (m_pixels) for (size_t i = 0; i < n; ++i) m_pixels[i] = Pixel();
    else m_pixels = new Pixel[n];
    m_width = width; m_height = height;
}
namespace {
unsigned char mapByte(float r) { const float m = 255 * r; return m > 255 ? 255 : (unsigned char)m; }
}
void Image::setPixel(int x, int y, const Vector3& p)
{
    if (x >= 0 && x < m_width && y >= 0 && y < m_height) m_pixels[(size_t)y * m_width + x] = Pixel(mapByte(p.x), mapByte(p.y), mapByte(p.z));
}
void Image::setPixel(int x, int y, const Pixel& p)
{
    if (x >= 0 && x < m_width && y >= 0 && y < m_height) m_pixels[(size_t)y * m_width + x] = p;
}
void Image::clear(const Vector3& c)
{
    std::fill(m_pixels, m_pixels + (size_t)m_width * m_height, Pixel(mapByte(c.x), mapByte(c.y), mapByte(c.z)));
}
void Image::writePPM(const char* pcFile)
{
    FILE* fp = fopen(pcFile, "wb");
    if (!fp) return;
    fprintf(fp, "P6\n%d %d\n255\n", m_width, m_height);
    for (int y = m_height - 1; y >= 0; --y) fwrite(m_pixels + (size_t)y * m_width, 3, m_width, fp);   // row 0 is the bottom scanline
    fclose(fp);
}

// ================================= Scene =====================================================================
Scene::Scene()
    : PhotonsPerLightSource(200000), CausticPhotonsPerLightSource(200000),
      renderSpp(1), renderJitter(0), renderMode(MIROGPU_RENDER_WHITTED), renderShadows(1), renderSeed(168), lastRenderSeconds(0),
      m_bgColor(0.f), m_usePhotonMaps(false)
{
    // the reference reserves 20.1 M photons per map up front (Scene.h:18); maps here grow on store()
    m_photonMap = new Photon_map(0);
    m_causticMap = new Photon_map(0);
}

Scene::~Scene() { delete m_photonMap; delete m_causticMap; }

void Scene::preCalc()
{
    for (Objects::iterator it = m_objects.begin(); it != m_objects.end(); ++it) (*it)->preCalc();
    for (Lights::iterator it = m_lights.begin(); it != m_lights.end(); ++it) (*it)->preCalc();
    for (Objects::iterator it = m_unboundedObjects.begin(); it != m_unboundedObjects.end(); ++it) (*it)->preCalc();
    m_bvh.setUnbounded(&m_unboundedObjects);
    m_bvh.build(&m_objects);
    std::vector<mirogpu_light> ls;
    for (size_t i = 0; i < m_lights.size(); ++i) {
        mirogpu_light l; memset(&l, 0, sizeof l);
        const PointLight* p = m_lights[i];
        for (int k = 0; k < 3; ++k) { l.position[k] = p->position()[k]; l.color[k] = p->color()[k]; }
        l.wattage = p->wattage();
        if (DirectionalAreaLight* d = dynamic_cast<DirectionalAreaLight*>(m_lights[i])) {
            l.kind = 1; l.radius = d->getRadius();
            const Vector3 n = d->getNormal();
            for (int k = 0; k < 3; ++k) l.normal[k] = n[k];
        }
        ls.push_back(l);
    }
    if (mirogpu_scene_set_lights(m_bvh.handle(), ls.empty() ? 0 : ls.data(), (uint32_t)ls.size()) != MIROGPU_OK) die("Scene::preCalc lights");
    // Scene.cpp:76-82: generate the photon maps (only DirectionalAreaLights emit, Scene.cpp:368)
    bool emitter = false;
    for (size_t i = 0; i < ls.size(); ++i) emitter |= ls[i].kind == 1;
    if (emitter) {
        tracePhotons();
        traceCausticPhotons();
    }
    if (m_photonMap->stored() > 0) m_photonMap->attach(m_bvh.handle(), 0);
    if (m_causticMap->stored() > 0) m_causticMap->attach(m_bvh.handle(), 1);
}

long Scene::tracePhotons() { return tracePhotonPass(*m_photonMap, 0, PhotonsPerLightSource, false); }
long Scene::traceCausticPhotons() { return tracePhotonPass(*m_causticMap, 1, CausticPhotonsPerLightSource, true); }

long Scene::tracePhotonPass(Photon_map& map, int which, int target, bool caustic)
{
    if (target == 0) {                                         // Scene.cpp:353-357
        if (!map.balanced()) map.balance();                    // (a map the caller filled and balanced by hand stays as it is)
        return 0;
    }
    // The whole pass -- emission loop with the reference's stop rule, store, scale_photon_power, balance -- runs on the
    // device (mirogpu_photon_pass); the map is attached there and mirrored to this object only when somebody reads it.
    long long emissions = 0;
    int stored = 0;
    const long long kMaxEmissions = 1LL << 28;                 // the reference loops forever when nothing can be stored
    if (mirogpu_photon_pass(m_bvh.handle(), which, caustic ? 1 : 0, renderSeed + (caustic ? 1u : 0u), target, kMaxEmissions,
                            &emissions, &stored) != MIROGPU_OK) die("Scene::tracePhotons");
    if (emissions >= kMaxEmissions) fprintf(stderr, "Scene::tracePhotons: gave up after %lld emissions (%d of %d photons stored)\n", emissions, stored, target);
    map.adoptDevice(m_bvh.handle(), which, stored);
    return (long)emissions;
}

// Scene.cpp:232-266 for UV-lookup materials with zero bump height: the perturbation vanishes, N is normalised.
// Scene.cpp:232-262: materials with UV lookup coordinates (plain Phong, 2-D textures) get the bump-mapped, normalised normal --
// the perturbation vanishes for every texture but StoneTexture --, materials with UVW coordinates keep the normal as the object
// computed it.  (The device applies the same rule to every hit it shades: resolve_hit_textured, csrc/kernels.cuh.)
void Scene::postProcess(HitInfo& minHit) const
{
    if (minHit.material && minHit.material->GetLookupCoordinates() != UV) return;
    const float delta = 0.0001f;
    float dx = 0.f, dy = 0.f;
    if (minHit.material && minHit.object) {
        const tex_coord2d_t c = minHit.object->toUVCoordinates(minHit.P);
        const float u1 = minHit.material->bumpHeight2D(tex_coord2d_t(c.u - delta, c.v)), u2 = minHit.material->bumpHeight2D(tex_coord2d_t(c.u + delta, c.v));
        const float v1 = minHit.material->bumpHeight2D(tex_coord2d_t(c.u, c.v - delta)), v2 = minHit.material->bumpHeight2D(tex_coord2d_t(c.u, c.v + delta));
        dx = (u2 - u1) / (2 * delta); dy = (v2 - v1) / (2 * delta);
    }
    if (dx != 0.f || dy != 0.f) {
        const float n[3] = {minHit.N.x, minHit.N.y, minHit.N.z};
        int m = 0;
        if (n[1] > n[0]) m = 1;
        if (n[2] > n[m]) m = 2;
        const Vector3 randomVec(m == 2 ? -n[2] : 0, m == 0 ? -n[0] : 0, m == 1 ? -n[1] : 0);
        const Vector3 t1 = cross(minHit.N, randomVec);
        minHit.N += dx * (cross(minHit.N, t1)) - dy * (cross(minHit.N, cross(minHit.N, t1)));
    }
    minHit.N.normalize();
}

tex_coord2d_t Sphere::toUVCoordinates(const Vector3& xyz) const
{
    Vector3 dir = xyz - m_center;
    dir.normalize();
    tex_coord2d_t coords;
    coords.u = (atan2f(dir.x, dir.z)) / (2.0f * PI) + 0.5;
    coords.v = (std::max(-1.0f, std::min(1.0f, asinf(dir.y)))) / PI + 0.5;
    return coords;
}

tex_coord2d_t Triangle::toUVCoordinates(const Vector3& xyz) const
{
    if (m_mesh->numTextCoords() == 0 || !m_mesh->tIndices()) return tex_coord2d_t();
    const TriangleMesh::TupleI3 vi3 = m_mesh->vIndices()[m_index], ti3 = m_mesh->tIndices()[m_index];
    const Vector3 &vA = m_mesh->vertices()[vi3.v[0]], &vB = m_mesh->vertices()[vi3.v[1]], &vC = m_mesh->vertices()[vi3.v[2]];
    const TriangleMesh::VectorR2 &tA = m_mesh->texCoords()[ti3.v[0]], &tB = m_mesh->texCoords()[ti3.v[1]], &tC = m_mesh->texCoords()[ti3.v[2]];
    // barycentrics in the plane that drops one axis, by Cramer's rule
    const Vector3 BmA = vB - vA, CmA = vC - vA;
    const Vector3 normal = cross(BmA, CmA);
    int i = 0, j = 1;
    if (normal.x > normal.z) i = 2;
    else if (normal.y > normal.z) j = 2;
    const float p[3] = {xyz.x - vA.x, xyz.y - vA.y, xyz.z - vA.z}, B[3] = {BmA.x, BmA.y, BmA.z}, C[3] = {CmA.x, CmA.y, CmA.z};
    const float detPC = p[i] * C[j] - C[i] * p[j], detBP = B[i] * p[j] - p[i] * B[j], detBC = B[i] * C[j] - C[i] * B[j];
    const float beta = std::max(detPC / detBC, 0.f), gamma = std::max(detBP / detBC, 0.f);
    const float alpha = std::max(1 - (beta + gamma), 0.f);
    tex_coord2d_t UV;
    UV.u = alpha * tA.x + beta * tB.x + gamma * tC.x;
    UV.v = alpha * tA.y + beta * tB.y + gamma * tC.y;
    return UV;
}

bool Scene::trace(HitInfo& minHit, const Ray& ray, float tMin, float tMax) const
{
    bool result = m_bvh.intersect(minHit, ray, tMin, tMax);    // planes included (the device's post-walk list)
    for (size_t i = 0; i < m_unboundedObjects.size(); ++i) {
        if (m_bvh.onDevice(m_unboundedObjects[i])) continue;
        HitInfo tmp;
        if (m_unboundedObjects[i]->intersect(tmp, ray, tMin, tMax) && (!result || tmp.t < minHit.t)) {
            result = true; minHit = tmp; minHit.object = m_unboundedObjects[i];
        }
    }
    if (result) postProcess(minHit);
    return result;
}

size_t Scene::traceBatch(const Ray* rays, size_t n, HitInfo* results, bool* hitFlags, float tMin, float tMax) const
{
    std::vector<char> flags(n);
    m_bvh.intersectBatch(rays, n, results, reinterpret_cast<bool*>(flags.data()), tMin, tMax);
    size_t nh = 0;
    for (size_t r = 0; r < n; ++r) {
        bool result = flags[r] != 0;
        for (size_t i = 0; i < m_unboundedObjects.size(); ++i) {
            if (m_bvh.onDevice(m_unboundedObjects[i])) continue;
            HitInfo tmp;
            if (m_unboundedObjects[i]->intersect(tmp, rays[r], tMin, tMax) && (!result || tmp.t < results[r].t)) {
                result = true; results[r] = tmp; results[r].object = m_unboundedObjects[i];
            }
        }
        if (result) postProcess(results[r]);
        if (hitFlags) hitFlags[r] = result;
        nh += result;
    }
    return nh;
}

void Scene::raytraceImage(Camera* cam, Image* img)
{
    const int w = img->width(), h = img->height();
    mirogpu_render_params p; memset(&p, 0, sizeof p);
    p.width = w; p.height = h; p.spp = renderSpp; p.jitter = renderJitter; p.max_depth = (int)TRACE_DEPTH; p.mode = renderMode;
    p.seed = renderSeed; p.tonemap = 1; p.row_begin = 0; p.row_end = h; p.row_stride = 1; p.row_phase = 0;
    for (int k = 0; k < 3; ++k) p.bg_color[k] = m_bgColor[k];
    p.use_photon_maps = m_usePhotonMaps ? 1 : 0;
    p.shadows = renderShadows;
    const mirogpu_camera c = cam->abi();
    const double t0 = wall();
    // the device applies the tone map and Image::setPixel's 8-bit mapping; Image::Pixel is 3 bytes, row 0 = bottom
    if (mirogpu_render_rgb8(m_bvh.handle(), &c, &p, img->getCharPixels()) != MIROGPU_OK) die("Scene::raytraceImage");
    lastRenderSeconds = wall() - t0;
    printf("Time spent raytracing image: %lf seconds.\n", lastRenderSeconds);
}

// ================================= Photon_map ================================================================
Photon_map::Photon_map(int max_phot)
    : photons(0), stored_photons(0), half_stored_photons(0), max_photons(max_phot), prev_scale(1), m_handle(0), m_which(0), m_balanced(false),
      m_on_device(false), m_host_stale(false)
{
    photons = (Photon*)malloc(sizeof(Photon) * ((size_t)std::max(max_photons, 0) + 1));
    memset(photons, 0, sizeof(Photon));
    bbox_min[0] = bbox_min[1] = bbox_min[2] = 1e8f;
    bbox_max[0] = bbox_max[1] = bbox_max[2] = -1e8f;
}

Photon_map::~Photon_map() { free(photons); }

int Photon_map::balanceDevice = 0;

// A map the device built (mirogpu_photon_pass): the host array is filled in on first use.
void Photon_map::adoptDevice(mirogpu_handle h, int which, int stored)
{
    m_handle = h; m_which = which;
    stored_photons = stored; half_stored_photons = stored / 2 - 1; prev_scale = stored;
    m_balanced = true; m_on_device = stored > 0; m_host_stale = stored > 0;
}

void Photon_map::syncHost() const
{
    if (!m_host_stale) return;
    Photon_map* self = const_cast<Photon_map*>(this);
    if (stored_photons > max_photons) {
        self->max_photons = stored_photons;
        self->photons = (Photon*)realloc(photons, sizeof(Photon) * ((size_t)max_photons + 1));
    }
    int n = 0;
    if (mirogpu_photon_download(m_handle, m_which, self->photons, max_photons, &n) != MIROGPU_OK || n != stored_photons) die("Photon_map: device map download");
    for (int i = 1; i <= n; ++i)
        for (int k = 0; k < 3; ++k) {
            self->bbox_min[k] = std::min(bbox_min[k], photons[i].pos[k]);
            self->bbox_max[k] = std::max(bbox_max[k], photons[i].pos[k]);
        }
    self->m_host_stale = false;
}

void Photon_map::store(const float power[3], const float pos[3], const float dir[3])
{
    syncHost();
    if (stored_photons >= max_photons) {        // grow instead of the reference's silent drop at capacity
        max_photons = max_photons ? 2 * max_photons : 1024;
        photons = (Photon*)realloc(photons, sizeof(Photon) * ((size_t)max_photons + 1));
    }
    m_balanced = false; m_on_device = false;
    Photon* const node = &photons[++stored_photons];
    for (int i = 0; i < 3; ++i) {
        node->pos[i] = pos[i];
        bbox_min[i] = std::min(bbox_min[i], node->pos[i]);
        bbox_max[i] = std::max(bbox_max[i], node->pos[i]);
        node->power[i] = power[i];
    }
    node->plane = 0;
    // direction quantised to two bytes (PhotonMap.cpp:275-287)
    const int theta = int(acos(dir[2]) * (256.0 / M_PI));
    node->theta = theta > 255 ? 255 : (unsigned char)theta;
    const int phi = int(atan2(dir[1], dir[0]) * (256.0 / (2.0 * M_PI)));
    node->phi = phi > 255 ? 255 : (phi < 0 ? (unsigned char)(phi + 256) : (unsigned char)phi);
}

void Photon_map::scale_photon_power(const float scale)
{
    syncHost();
    m_on_device = false;
    for (int i = prev_scale; i <= stored_photons; ++i)
        for (int k = 0; k < 3; ++k) photons[i].power[k] *= scale;
    prev_scale = stored_photons;
}

// PhotonMap.cpp:314-466 on the device (mirogpu_photon_balance): the array comes back in the reference's heap order --
// the quickselect of median_split is evaluated round for round there, so equal keys end up where Jensen's code puts them.
void Photon_map::balance(void)
{
    if (m_on_device && m_balanced) return;                     // the device pass balanced it
    syncHost();
    if (stored_photons > 1) {
        if (mirogpu_photon_balance(balanceDevice, photons, stored_photons, bbox_min, bbox_max) != MIROGPU_OK) die("Photon_map::balance");
    }
    half_stored_photons = stored_photons / 2 - 1;
    m_balanced = true;
}

void Photon_map::attach(mirogpu_handle h, int which)
{
    if (m_on_device && h == m_handle && which == m_which) return;   // built there already
    syncHost();
    m_handle = h; m_which = which;
    if (mirogpu_photon_upload(h, which, photons, stored_photons) != MIROGPU_OK) die("Photon_map::attach");
}

void Photon_map::irradiance_estimate_batch(float* irrad3, const float* pos3, const float* normal3, size_t n, float max_dist, int nphotons) const
{
    if (!m_handle) die("Photon_map::irradiance_estimate before attach");
    if (mirogpu_photon_gather(m_handle, m_which, pos3, normal3, n, max_dist, nphotons, irrad3) != MIROGPU_OK) die("Photon_map::irradiance_estimate");
}

void Photon_map::irradiance_estimate(float irrad[3], const float pos[3], const float normal[3], const float max_dist, const int nphotons) const
{
    irradiance_estimate_batch(irrad, pos, normal, 1, max_dist, nphotons);
}
