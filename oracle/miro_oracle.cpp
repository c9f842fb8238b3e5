// oracle/miro_oracle.cpp -- TEST INFRASTRUCTURE (the parity oracle), not product code.
//
// A plain-CPU restatement of the reference's algorithm for the ray-intersection hot path
// (hallgeirl/cse168-raytracer, scalar build -- never the SSE build, whose _mm_rcp_ps makes it
// unfit as a parity oracle).  Each function cites the reference file:line it follows.  Arithmetic is
// IEEE binary32 in the reference's operand order; this file is compiled with -ffp-contract=off and
// no -march flag, so no FMA is ever formed and results are bit-identical to the reference's own
// scalar build (checked by tests/test_oracle_vs_reference.py against oracle/_ref, and by the
// known-answer BVH node counts of writeup/A2/Readme.tex:96-97).
//
// Parity status: PINNED -- (1) bunny BVH 42 881 nodes / 21 441 leaves (published KAT);
// (2) bit-exact hits, counters, loader output, eye rays and photon gathers against the reference
// compiled in place (oracle/_ref) on identical inputs; (3) committed fixtures in tests/golden/.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library.  The product (libmirogpu.so) never links or calls it.

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>
#include <algorithm>
#include <omp.h>

namespace {

const float kInf = std::numeric_limits<float>::infinity();
const float MIRO_TMAX = 1e12f;                                         // Miro.h:8
const float kEps = 1e-4f;                                              // Miro.h:9
const float PI = 3.1415926535897932384626433832795028841972f;          // Miro.h:10
const float DegToRad = PI / 180.0f;                                    // Miro.h:11
const float HalfDegToRad = DegToRad / 2.0f;                            // Camera.cpp:15

// ---- Vector3.h arithmetic, same operand order ---------------------------------------------------
struct V3 {
    float x, y, z;
    V3() : x(0), y(1), z(2) {}                                         // Vector3.h:26-27 (yes: 0,1,2)
    V3(float s) : x(s), y(s), z(s) {}
    V3(float a, float b, float c) : x(a), y(b), z(c) {}
    float operator[](int i) const { return (&x)[i]; }
    float& operator[](int i) { return (&x)[i]; }
};
inline V3 operator+(const V3& a, const V3& b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(const V3& a, const V3& b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator-(const V3& a) { return V3(-a.x, -a.y, -a.z); }
inline V3 operator*(const V3& a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(float s, const V3& a) { return V3(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(const V3& a, const V3& b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline V3 operator/(const V3& a, float s) { float inv = float(1) / s; return V3(a.x * inv, a.y * inv, a.z * inv); }  // Vector3.h:125-129
inline float dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }      // Vector3.h:243-246
inline V3 cross(const V3& a, const V3& b)                                                     // Vector3.h:251-256
{
    return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline float length2(const V3& a) { return dot(a, a); }
inline float length(const V3& a) { return sqrtf(length2(a)); }
inline V3 normalized(const V3& a) { return a / length(a); }                                   // Vector3.h:205-208
inline float average(const V3& a) { return (a.x + a.y + a.z) / 3.0f; }                        // Vector3.h:226-229

// ---- Matrix4x4.h ---------------------------------------------------------------------------------
struct M4 {
    float m[4][4];  // m[i][j] = m_(i+1)(j+1), row i column j
    M4() { for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) m[i][j] = (i == j) ? 1.0f : 0.0f; }
};
inline V3 xform(const M4& A, const V3& u)                                                     // Matrix4x4.h:581-588
{
    return V3(A.m[0][0] * u.x + A.m[0][1] * u.y + A.m[0][2] * u.z + A.m[0][3],
              A.m[1][0] * u.x + A.m[1][1] * u.y + A.m[1][2] * u.z + A.m[1][3],
              A.m[2][0] * u.x + A.m[2][1] * u.y + A.m[2][2] * u.z + A.m[2][3]);
}
M4 inverted(const M4& A)                                                                      // Matrix4x4.h:309-369
{
    const float m11 = A.m[0][0], m12 = A.m[0][1], m13 = A.m[0][2], m14 = A.m[0][3];
    const float m21 = A.m[1][0], m22 = A.m[1][1], m23 = A.m[1][2], m24 = A.m[1][3];
    const float m31 = A.m[2][0], m32 = A.m[2][1], m33 = A.m[2][2], m34 = A.m[2][3];
    const float m41 = A.m[3][0], m42 = A.m[3][1], m43 = A.m[3][2], m44 = A.m[3][3];
    float T34C12 = m31 * m42 - m32 * m41, T34C13 = m31 * m43 - m33 * m41, T34C14 = m31 * m44 - m34 * m41;
    float T34C23 = m32 * m43 - m33 * m42, T34C24 = m32 * m44 - m34 * m42, T34C34 = m33 * m44 - m34 * m43;
    float T24C12 = m21 * m42 - m22 * m41, T24C13 = m21 * m43 - m23 * m41, T24C14 = m21 * m44 - m24 * m41;
    float T24C23 = m22 * m43 - m23 * m42, T24C24 = m22 * m44 - m24 * m42, T24C34 = m23 * m44 - m24 * m43;
    float T23C12 = m21 * m32 - m22 * m31, T23C13 = m21 * m33 - m23 * m31, T23C14 = m21 * m34 - m24 * m31;
    float T23C23 = m22 * m33 - m23 * m32, T23C24 = m22 * m34 - m24 * m32, T23C34 = m23 * m34 - m24 * m33;
    float sd11 = m22 * T34C34 - m23 * T34C24 + m24 * T34C23;
    float sd12 = m21 * T34C34 - m23 * T34C14 + m24 * T34C13;
    float sd13 = m21 * T34C24 - m22 * T34C14 + m24 * T34C12;
    float sd14 = m21 * T34C23 - m22 * T34C13 + m23 * T34C12;
    float sd21 = m12 * T34C34 - m13 * T34C24 + m14 * T34C23;
    float sd22 = m11 * T34C34 - m13 * T34C14 + m14 * T34C13;
    float sd23 = m11 * T34C24 - m12 * T34C14 + m14 * T34C12;
    float sd24 = m11 * T34C23 - m12 * T34C13 + m13 * T34C12;
    float sd31 = m12 * T24C34 - m13 * T24C24 + m14 * T24C23;
    float sd32 = m11 * T24C34 - m13 * T24C14 + m14 * T24C13;
    float sd33 = m11 * T24C24 - m12 * T24C14 + m14 * T24C12;
    float sd34 = m11 * T24C23 - m12 * T24C13 + m13 * T24C12;
    float sd41 = m12 * T23C34 - m13 * T23C24 + m14 * T23C23;
    float sd42 = m11 * T23C34 - m13 * T23C14 + m14 * T23C13;
    float sd43 = m11 * T23C24 - m12 * T23C14 + m14 * T23C12;
    float sd44 = m11 * T23C23 - m12 * T23C13 + m13 * T23C12;
    float detInv = 1.0 / (m11 * sd11 - m12 * sd12 + m13 * sd13 - m14 * sd14);   // double divide, as the reference
    M4 R;
    R.m[0][0] = sd11 * detInv;  R.m[0][1] = -sd21 * detInv; R.m[0][2] = sd31 * detInv;  R.m[0][3] = -sd41 * detInv;
    R.m[1][0] = -sd12 * detInv; R.m[1][1] = sd22 * detInv;  R.m[1][2] = -sd32 * detInv; R.m[1][3] = sd42 * detInv;
    R.m[2][0] = sd13 * detInv;  R.m[2][1] = -sd23 * detInv; R.m[2][2] = sd33 * detInv;  R.m[2][3] = -sd43 * detInv;
    R.m[3][0] = -sd14 * detInv; R.m[3][1] = sd24 * detInv;  R.m[3][2] = -sd34 * detInv; R.m[3][3] = sd44 * detInv;
    return R;
}
M4 transposed(const M4& A) { M4 R; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) R.m[i][j] = A.m[j][i]; return R; }

// ---- scene storage -------------------------------------------------------------------------------
struct Mesh {                                                          // TriangleMesh.h
    std::vector<V3> vertices, normals;
    std::vector<int> vidx, nidx;                                       // 3 per triangle
    int ntris;
};
struct Material { V3 kd, ks, kt; float shininess, refr; };             // Phong.cpp:13-32 after clamping
struct Prim { int mesh, index, material; V3 cmin, cmax, center; float radius; };   // Triangle.h:40-42 + cached bounds; mesh < 0: a Sphere (Sphere.h:34-35: centre in `center`, radius)
struct PlaneObj { V3 normal, origin; int material; };                  // Plane.h:35
struct Light { int kind; V3 pos, color, normal; float wattage, radius; };  // 0 = PointLight, 1 = DirectionalAreaLight

struct Node {                                                          // BVH.h:30-62
    float corners[2][4];
    bool leaf;
    Node* child[2];
    std::vector<int> objs;
    Node() : leaf(false) { corners[0][0] = kInf; child[0] = child[1] = 0; }   // BVH.h:32
};

struct Photon { float pos[3]; short plane; unsigned char theta, phi; float power[3]; };   // PhotonMap.h:16-22
static_assert(sizeof(Photon) == 28, "Photon must be 28 bytes like the reference's");

struct PhotonMap {                                                     // PhotonMap.h:41-101
    std::vector<Photon> photons;                                       // 1-indexed, entry 0 unused
    int stored, half_stored, max_photons, prev_scale;
    float costheta[256], sintheta[256], cosphi[256], sinphi[256];
    float bbox_min[3], bbox_max[3];
    explicit PhotonMap(int max_phot)                                   // PhotonMap.cpp:26-54
    {
        stored = 0; prev_scale = 1; max_photons = max_phot; half_stored = 0;
        photons.resize((size_t)max_phot + 1);
        bbox_min[0] = bbox_min[1] = bbox_min[2] = 1e8f;
        bbox_max[0] = bbox_max[1] = bbox_max[2] = -1e8f;
        for (int i = 0; i < 256; i++) {
            double angle = double(i) * (1.0 / 256.0) * M_PI;
            costheta[i] = cos(angle); sintheta[i] = sin(angle);
            cosphi[i] = cos(2.0 * angle); sinphi[i] = sin(2.0 * angle);
        }
    }
};

struct Scene {
    std::vector<Mesh*> meshes;
    std::vector<Material> materials;
    std::vector<Prim> prims;
    std::vector<PlaneObj> planes;                                      // Scene::m_unboundedObjects (Scene.h:24-25)
    std::vector<Light> lights;
    Node* root;
    long long n_nodes, n_leaves;
    V3 bg;
    // camera (Camera.h) -- basis derived lazily like Camera::eyeRay's statics
    V3 eye, up, viewDir; float fov;
    Scene() : root(0), n_nodes(0), n_leaves(0), bg(0.f), eye(0, 0, 0), up(0, 1, 0), viewDir(0, 0, -1), fov(45.f) {}
};

Scene* g = 0;
std::vector<PhotonMap*> g_pms;
long long g_ray_box = 0, g_ray_tri = 0;

// ---- TriangleMeshLoad.cpp:81-111 ------------------------------------------------------------------
void getIndices(char* word, int* vindex, int* tindex, int* nindex)
{
    char* null = (char*)" ";
    char *ptr, *tp = null, *np = null;
    for (ptr = word; *ptr != '\0'; ptr++) {
        if (*ptr == '/') {
            if (tp == null) tp = ptr + 1; else np = ptr + 1;
            *ptr = '\0';
        }
    }
    *vindex = atoi(word); *tindex = atoi(tp); *nindex = atoi(np);
}

// TriangleMeshLoad.cpp:114-311 (two passes, 80-char fgets, first three face tokens, normal synthesis + averaging)
Mesh* loadObj(FILE* fp, const M4& ctm)
{
    int nv = 0, nt = 0, nn = 0, nf = 0;
    char line[81];
    while (fgets(line, 80, fp) != 0) {
        if (line[0] == 'v') { if (line[1] == 'n') nn++; else if (line[1] == 't') nt++; else nv++; }
        else if (line[0] == 'f') nf++;
    }
    fseek(fp, 0, 0);
    Mesh* M = new Mesh;
    const int ncap = std::max(nv, nf * 3);
    M->normals.resize(ncap); M->vertices.resize(nv);
    std::vector<std::vector<int> > neighboringNormals(nv);
    std::vector<char> fixNormal(ncap, 0);
    M->vidx.resize((size_t)nf * 3); M->nidx.resize((size_t)nf * 3);
    int numTris = 0, nvertices = 0, nnormals = 0;
    const M4 nctm = transposed(inverted(ctm));                         // :176-178
    while (fgets(line, 80, fp) != 0) {
        if (line[0] == 'v') {
            if (line[1] == 'n') {
                float x, y, z; sscanf(&line[2], "%f %f %f\n", &x, &y, &z);
                M->normals[nnormals] = normalized(xform(nctm, V3(x, y, z)));
                fixNormal[nnormals] = 0; nnormals++;
            } else if (line[1] == 't') {
                // texture coordinates are not on the hot path
            } else {
                float x, y, z; sscanf(&line[1], "%f %f %f\n", &x, &y, &z);
                M->vertices[nvertices] = xform(ctm, V3(x, y, z)); nvertices++;
            }
        } else if (line[0] == 'f') {
            char s1[32], s2[32], s3[32]; int v, t, n;
            sscanf(&line[1], "%s %s %s\n", s1, s2, s3);
            char* toks[3] = {s1, s2, s3};
            for (int k = 0; k < 3; ++k) {
                getIndices(toks[k], &v, &t, &n);
                M->vidx[3 * numTris + k] = v - 1;
                if (n) { M->nidx[3 * numTris + k] = n - 1; neighboringNormals[v - 1].push_back(n - 1); }
            }
            if (!n) {   // decided by the THIRD token only (:252)
                V3 e1 = M->vertices[M->vidx[3 * numTris + 1]] - M->vertices[M->vidx[3 * numTris + 0]];
                V3 e2 = M->vertices[M->vidx[3 * numTris + 2]] - M->vertices[M->vidx[3 * numTris + 0]];
                for (int i = 0; i < 3; i++) {
                    M->normals[nnormals] = normalized(cross(e1, e2));
                    fixNormal[nnormals] = 1; nnormals++;
                }
                M->nidx[3 * numTris + 0] = nnormals - 3; M->nidx[3 * numTris + 1] = nnormals - 2; M->nidx[3 * numTris + 2] = nnormals - 1;
                for (int k = 0; k < 3; ++k) neighboringNormals[M->vidx[3 * numTris + k]].push_back(M->nidx[3 * numTris + k]);
            }
            numTris++;
        }
    }
    for (int i = 0; i < nvertices; i++) {                              // :288-308
        if (neighboringNormals[i].size() == 0) continue;
        V3 avg;                                                        // starts at (0,1,2): Vector3's default ctor
        for (size_t j = 0; j < neighboringNormals[i].size(); j++) avg = avg + M->normals[neighboringNormals[i][j]];
        avg = avg / (float)neighboringNormals[i].size();
        avg = normalized(avg);
        for (size_t j = 0; j < neighboringNormals[i].size(); j++) {
            if (!fixNormal[neighboringNormals[i][j]]) continue;
            M->normals[neighboringNormals[i][j]] = avg;
        }
    }
    M->ntris = numTris;
    return M;
}

void addPrim(int mesh, int index, int material)
{
    Prim p; p.mesh = mesh; p.index = index; p.material = material;
    g->prims.push_back(p);
}

// Triangle::updateMinMax (Triangle.cpp:97-118) and Triangle::center (Triangle.cpp:41-48)
void primPreCalc(Prim& p)
{
    if (p.mesh < 0) { p.cmin = p.center - V3(p.radius); p.cmax = p.center + V3(p.radius); return; }   // Sphere.h:21-23
    const Mesh* m = g->meshes[p.mesh];
    V3 v[3] = {m->vertices[m->vidx[3 * p.index]], m->vertices[m->vidx[3 * p.index + 1]], m->vertices[m->vidx[3 * p.index + 2]]};
    p.cmin = v[0]; p.cmax = v[0];
    for (int i = 1; i < 3; i++) {
        if (v[i].x < p.cmin.x) p.cmin.x = v[i].x;
        if (v[i].y < p.cmin.y) p.cmin.y = v[i].y;
        if (v[i].z < p.cmin.z) p.cmin.z = v[i].z;
        if (v[i].x > p.cmax.x) p.cmax.x = v[i].x;
        if (v[i].y > p.cmax.y) p.cmax.y = v[i].y;
        if (v[i].z > p.cmax.z) p.cmax.z = v[i].z;
    }
    V3 BmA = v[1] - v[0], CmA = v[2] - v[0];
    p.center = v[0] + BmA / 3 + CmA / 3;
}

// ---- BVH.cpp:14-58 ----------------------------------------------------------------------------------
void getCornerPoints(float (&out)[2][4], const std::vector<int>& objs)
{
    for (int i = 0; i < 3; i++) { out[0][i] = kInf; out[1][i] = -kInf; }
    for (size_t i = 0; i < objs.size(); ++i) {
        const Prim& p = g->prims[objs[i]];
        for (int j = 0; j < 3; j++) {
            if (out[1][j] < p.cmax[j]) out[1][j] = p.cmax[j];
            if (out[0][j] > p.cmin[j]) out[0][j] = p.cmin[j];
        }
    }
}
float getArea(const float (&c)[2][4])
{
    float area = 0;
    for (int d = 0; d < 3; d++) {
        int dim1 = (d + 1) % 3, dim2 = (d + 2) % 3;
        area += (c[1][dim1] - c[0][dim1]) * (c[1][dim2] - c[0][dim2]);
    }
    return 2 * area;
}
inline float getCost(const float (&c)[2][4], const std::vector<int>& objs)
{
    if (objs.size() == 0) return 0;
    return ((float)objs.size()) * getArea(c);
}

const int OBJECTS_PER_LEAF = 4;   // BVH.h:61 (scalar build)
const int MAX_TREE_DEPTH = 32;    // BVH.h:55

// BVH::build, BVH.cpp:60-339 (scalar branch).  objs is consumed.
void build(Node* node, std::vector<int>& objs, int depth)
{
    g->n_nodes++;
    if (node->corners[0][0] == kInf) getCornerPoints(node->corners, objs);
    for (int i = 0; i < 3; i++) { node->corners[0][i] -= kEps; node->corners[1][i] += kEps; }

    if ((int)objs.size() <= OBJECTS_PER_LEAF || depth >= MAX_TREE_DEPTH) {
        node->objs.swap(objs);
        node->leaf = true;
        g->n_leaves++;
        return;
    }
    float bestCost = kInf, bestPosition = 0;
    int bestDim = 0;
    float bestCorners[2][2][3];
    const int maxSearchDepth = 32;
    node->leaf = false;

    for (int dim = 0; dim < 3; dim++) {
        float current = (node->corners[1][dim] + node->corners[0][dim]) / 2.0f, beg = node->corners[0][dim], end = node->corners[1][dim];
        int nocheckBoundary[2] = {0, 0};
        std::vector<int> children[2];
        float corners[2][2][4];
        for (size_t i = 0; i < objs.size(); i++) {
            if (g->prims[objs[i]].center[dim] < current) children[0].push_back(objs[i]);
            else children[1].push_back(objs[i]);
        }
        getCornerPoints(corners[0], children[0]);
        getCornerPoints(corners[1], children[1]);

        for (int searchDepth = 0; searchDepth < maxSearchDepth; searchDepth++) {
            float costLeft = getCost(corners[0], children[0]), costRight = getCost(corners[1], children[1]);
            if (costLeft + costRight < bestCost) {
                bestCost = costLeft + costRight; bestDim = dim; bestPosition = current;
                for (int i = 0; i < 2; i++) for (int j = 0; j < 3; j++) { bestCorners[0][i][j] = corners[0][i][j]; bestCorners[1][i][j] = corners[1][i][j]; }
            }
            bool canShrink = false;
            int largest = (costLeft > costRight ? 0 : 1);
            int smallest = (largest + 1) % 2;
            if (largest == 0) end = current; else beg = current;
            current = (beg + end) / 2;
            nocheckBoundary[smallest] = (int)children[smallest].size();

            for (int i = (int)children[largest].size() - 1; i >= nocheckBoundary[largest]; i--) {
                const Prim& p = g->prims[children[largest][i]];
                if ((largest == 0 && p.center[dim] > current) || (largest == 1 && p.center[dim] < current)) {
                    const V3& cmax = p.cmax; const V3& cmin = p.cmin;
                    for (int j = 0; j < 3; j++) {
                        if (cmax[j] > corners[smallest][1][j]) corners[smallest][1][j] = cmax[j];
                        if (cmin[j] < corners[smallest][0][j]) corners[smallest][0][j] = cmin[j];
                        if (!canShrink) {
                            if (cmax[j] >= corners[largest][1][j] - kEps || cmin[j] <= corners[largest][0][j] + kEps) canShrink = true;
                        }
                    }
                    children[smallest].push_back(children[largest][i]);
                    std::swap(children[largest][i], children[largest][children[largest].size() - 1]);
                    children[largest].pop_back();
                }
            }
            if (canShrink) getCornerPoints(corners[largest], children[largest]);
        }
        float costLeft = getCost(corners[0], children[0]), costRight = getCost(corners[1], children[1]);
        if (costLeft + costRight < bestCost) {
            bestCost = costLeft + costRight; bestDim = dim; bestPosition = current;
            for (int i = 0; i < 2; i++) for (int j = 0; j < 3; j++) { bestCorners[0][i][j] = corners[0][i][j]; bestCorners[1][i][j] = corners[1][i][j]; }
        }
    }

    std::vector<int> part[2];                                          // :308-319
    for (size_t i = 0; i < objs.size(); i++) {
        if (g->prims[objs[i]].center[bestDim] < bestPosition) part[0].push_back(objs[i]);
        else part[1].push_back(objs[i]);
    }
    std::vector<int>().swap(objs);
    for (int i = 0; i < 2; i++) {                                      // :321-337
        Node* c = new Node;
        node->child[i] = c;
        for (int j = 0; j < 3; j++) { c->corners[0][j] = bestCorners[i][0][j]; c->corners[1][j] = bestCorners[i][1][j]; }
        build(c, part[i], depth + 1);
    }
}

struct Hit {                                                           // Ray.h:21-38
    float t; V3 P, N; int material; int object;
    Hit() : t(0.0f), P(), N(0.0f, 1.0f, 0.0f), material(-1), object(-1) {}
};
struct Ray { V3 o, d; };

struct Counters { long long box, tri; };

// Triangle::intersect, Triangle.cpp:136-169 (scalar branch)
inline bool triIntersect(int prim, Hit& result, const Ray& r, float tMin, float tMax)
{
    const Prim& p = g->prims[prim];
    const Mesh* m = g->meshes[p.mesh];
    const V3& A = m->vertices[m->vidx[3 * p.index]];
    const V3& B = m->vertices[m->vidx[3 * p.index + 1]];
    const V3& C = m->vertices[m->vidx[3 * p.index + 2]];
    const V3& nA = m->normals[m->nidx[3 * p.index]];
    const V3& nB = m->normals[m->nidx[3 * p.index + 1]];
    const V3& nC = m->normals[m->nidx[3 * p.index + 2]];
    V3 BmA = B - A, CmA = C - A;
    V3 normal = cross(BmA, CmA);
    float ddotn = dot(-r.d, normal);
    float t = dot(r.o - A, normal) / ddotn;
    float beta = dot(-r.d, cross(r.o - A, CmA)) / ddotn;
    float gamma = dot(-r.d, cross(BmA, r.o - A)) / ddotn;
    if (beta < -kEps || gamma < -kEps || beta + gamma > 1 + kEps || t < tMin || t > tMax) return false;
    result.P = A + beta * BmA + gamma * CmA;
    result.t = t;
    result.N = (1 - beta - gamma) * nA + beta * nB + gamma * nC;
    result.material = p.material;
    return true;
}

// Sphere::intersect, Sphere.cpp:28-69
inline bool sphereIntersect(const Prim& p, Hit& result, const Ray& ray, float tMin, float tMax)
{
    const V3 toO = ray.o - p.center;
    const float a = length2(ray.d);
    const float b = dot(2 * ray.d, toO);
    const float c = length2(toO) - p.radius * p.radius;
    const float discrim = b * b - 4.0f * a * c;
    if (discrim < 0) return false;
    const float sqrt_discrim = sqrtf(discrim);
    const float t[2] = {(-b - sqrt_discrim) / (2.0f * a), (-b + sqrt_discrim) / (2.0f * a)};
    if ((t[0] > tMin) && (t[0] < tMax)) result.t = t[0];
    else if ((t[1] > tMin) && (t[1] < tMax)) result.t = t[1];
    else return false;
    result.P = ray.o + result.t * ray.d;
    result.N = normalized(result.P - p.center);
    result.material = p.material;
    return true;
}
// Plane::intersect, Plane.cpp:33-48
inline bool planeIntersect(const PlaneObj& pl, Hit& result, const Ray& r, float tMin, float tMax)
{
    float ndotd = dot(pl.normal, r.d);
    if (fabs(ndotd) < 1e-6) return false;
    float t = dot(pl.normal, (pl.origin - r.o)) / ndotd;
    if (t < tMin || t > tMax) return false;
    result.P = r.o + t * r.d;
    result.t = t;
    result.N = pl.normal;
    result.material = pl.material;
    return true;
}
inline bool primIntersect(int prim, Hit& result, const Ray& r, float tMin, float tMax)
{
    const Prim& p = g->prims[prim];
    return p.mesh < 0 ? sphereIntersect(p, result, r, tMin, tMax) : triIntersect(prim, result, r, tMin, tMax);
}

// BVH::intersectChildren, BVH.cpp:471-658 (scalar branch)
bool intersectChildren(const Node* node, Hit& minHit, const Ray& ray, float tMin, float tMax, Counters& cn)
{
    bool hit = false;
    Hit tempMinHit;
    minHit.t = tMax;
    if (node->leaf) {
        for (size_t i = 0; i < node->objs.size(); ++i) {
            if (g->prims[node->objs[i]].mesh >= 0) cn.tri++;           // Stats::Ray_Tri_Intersect counts Triangles only (BVH.cpp:495-497)
            if (primIntersect(node->objs[i], tempMinHit, ray, tMin, minHit.t)) {
                if (tempMinHit.t < minHit.t) {
                    hit = true;
                    minHit = tempMinHit;
                    minHit.object = node->objs[i];
                }
            }
        }
        return hit;
    }
    float minT = kInf, minTother = kInf;
    int minIndex = -1, otherIndex = -1;
    for (int i = 0; i < 2; i++) {
        const Node* child = node->child[i];
        float minOverlap = -kInf, maxOverlap = kInf;
        float t[2];
        for (int j = 0; j < 3; ++j) {
            t[0] = (child->corners[0][j] - ray.o[j]) / ray.d[j];
            t[1] = (child->corners[1][j] - ray.o[j]) / ray.d[j];
            int m = t[0] > t[1];
            if (t[m] > minOverlap) minOverlap = t[m];
            if (t[m ^ 1] < maxOverlap) maxOverlap = t[m ^ 1];
        }
        if (minOverlap > maxOverlap || minOverlap > tMax || maxOverlap < tMin) continue;
        if (minT > minOverlap) { minTother = minT; otherIndex = minIndex; minT = minOverlap; minIndex = i; }
        else if (minTother > minOverlap) { otherIndex = i; minTother = minOverlap; }
    }
    if (minIndex == -1) return false;
    cn.box++;
    if (intersectChildren(node->child[minIndex], tempMinHit, ray, tMin, minHit.t, cn)) { minHit = tempMinHit; hit = true; }
    if (otherIndex != -1) {
        cn.box++;
        if (intersectChildren(node->child[otherIndex], tempMinHit, ray, tMin, minHit.t, cn)) { minHit = tempMinHit; hit = true; }
    }
    return hit;
}

// BVH::intersect, BVH.cpp:438-469
bool bvhIntersect(Hit& minHit, const Ray& ray, float tMin, float tMax, Counters& cn)
{
    const Node* root = g->root;
    minHit.t = tMax;
    float minOverlap = -kInf, maxOverlap = kInf;
    float t[2];
    for (int i = 0; i < 3; ++i) {
        t[0] = (root->corners[0][i] - ray.o[i]) / ray.d[i];
        t[1] = (root->corners[1][i] - ray.o[i]) / ray.d[i];
        int m = t[0] > t[1];
        if (t[m] > minOverlap) minOverlap = t[m];
        if (t[m ^ 1] < maxOverlap) maxOverlap = t[m ^ 1];
    }
    cn.box++;
    if (minOverlap > maxOverlap || minOverlap > tMax || maxOverlap < tMin) return false;
    return intersectChildren(root, minHit, ray, tMin, tMax, cn);
}

// Scene::trace, Scene.cpp:214-268.  All materials are
// UV-lookup Phong with bumpHeight2D == 0, so dx = dy = 0 and the branch reduces to N.normalize().
bool sceneTrace(Hit& minHit, const Ray& ray, float tMin, float tMax, Counters& cn)
{
    bool result = bvhIntersect(minHit, ray, tMin, tMax, cn);
    for (size_t i = 0; i < g->planes.size(); i++) {                    // the unbounded objects, Scene.cpp:219-230
        Hit tempMinHit;
        bool ubresult = planeIntersect(g->planes[i], tempMinHit, ray, tMin, tMax);
        if (ubresult && (!result || tempMinHit.t < minHit.t)) {
            result = true;
            minHit = tempMinHit;
            minHit.object = (int)(g->prims.size() + i);
        }
    }
    if (result) {
        const float dx = 0.0f, dy = 0.0f;                              // (u2-u1)/(2*delta) with all heights 0
        float n[3] = {minHit.N.x, minHit.N.y, minHit.N.z};
        int m = 0;
        if (n[1] > n[0]) m = 1;
        if (n[2] > n[m]) m = 2;
        V3 randomVec(m == 2 ? -n[2] : 0, m == 0 ? -n[0] : 0, m == 1 ? -n[1] : 0);
        V3 t1 = cross(minHit.N, randomVec);
        minHit.N = minHit.N + (dx * (cross(minHit.N, t1)) - dy * (cross(minHit.N, cross(minHit.N, t1))));
        minHit.N = normalized(minHit.N);
    }
    return result;
}

// Exhaustive closest hit over every primitive with the reference's triangle test.  Ties in t go to
// the smallest primitive index.  This is the order-independent statement the GPU result must equal.
bool bruteTrace(Hit& best, const Ray& ray, float tMin, float tMax)
{
    bool hit = false;
    best.t = tMax;
    Hit tmp;
    for (size_t i = 0; i < g->prims.size(); ++i) {
        if (primIntersect((int)i, tmp, ray, tMin, tMax)) {
            if (!(tmp.t == tmp.t)) continue;                           // NaN t: the leaf's strict < discards it (BVH.cpp:500)
            if (!hit ? (tmp.t <= best.t) : (tmp.t < best.t)) { best = tmp; best.object = (int)i; hit = true; }
        }
    }
    for (size_t i = 0; i < g->planes.size(); i++) {                    // then the unbounded objects, as Scene::trace does
        if (planeIntersect(g->planes[i], tmp, ray, tMin, tMax) && (!hit || tmp.t < best.t)) { best = tmp; best.object = (int)(g->prims.size() + i); hit = true; }
    }
    return hit;
}

// ---- Camera::eyeRay, Camera.cpp:104-161 (no DOF) ------------------------------------------------------
struct CamBasis { V3 w, u, v; float top, left, bottom, right; };
CamBasis camBasis(int W, int H)
{
    CamBasis b;
    b.w = normalized(-g->viewDir);
    b.u = normalized(cross(g->up, b.w));
    b.v = cross(b.w, b.u);
    float aspect = (float)W / (float)H;
    b.top = std::tan(g->fov * HalfDegToRad);
    b.right = aspect * b.top; b.bottom = -b.top; b.left = -b.right;
    return b;
}
inline Ray eyeRay(const CamBasis& b, int x, int y, int W, int H, float dx, float dy)
{
    const float U = b.left + (b.right - b.left) * (((float)x + dx) / (float)W);
    const float V = b.bottom + (b.top - b.bottom) * (((float)y + dy) / (float)H);
    Ray r; r.o = g->eye; r.d = normalized(U * b.u + V * b.v - b.w);
    return r;
}

// alignHemisphereToVector, Utility.h:34-50
inline V3 alignHemisphereToVector(const V3& v, float theta, float phi)
{
    float u1 = std::sin(phi) * std::cos(theta);
    float u2 = std::sin(phi) * std::sin(theta);
    float u3 = std::cos(phi);
    V3 t1 = cross(V3(0, 0, 1), v);
    if (length2(t1) < 1e-6) t1 = cross(V3(0, 1, 0), v);
    V3 aligned_d(u1 * t1 + u2 * cross(t1, v) + u3 * v);
    return normalized(aligned_d);
}

// Ray::reflect (Ray.h:143-165, non-PATH_TRACING), getReflectionCoefficient (:168-200), refract (:202-243)
inline Ray reflectRay(const Ray& in, const Hit& h)
{
    V3 d_r = in.d - 2 * dot(h.N, in.d) * h.N;
    d_r = normalized(d_r);
    Ray r; r.o = h.P + d_r * kEps; r.d = d_r; return r;
}
inline float reflectionCoefficient(const Ray& in, const Hit& h)
{
    float n1, n2; V3 n;
    const float ri = g->materials[h.material].refr;
    if (dot(in.d, h.N) < 0) { n1 = 1.0f; n2 = ri; n = h.N; } else { n1 = ri; n2 = 1.0f; n = -h.N; }
    float cosTheta = dot(-in.d, n);
    float sinTheta = std::sin(std::acos(cosTheta));
    float powsomething = std::pow((n1 / n2) * sinTheta, 2.f);
    if (powsomething > 1.f) return 1;
    float sqrtSinTheta = std::sqrt(1.f - (powsomething));
    return std::pow((n1 * cosTheta - sqrtSinTheta) / (n1 * cosTheta + sqrtSinTheta), 2.f);
}
inline Ray refractRay(const Ray& in, const Hit& h)
{
    float n1, n2; V3 n;
    const float ri = g->materials[h.material].refr;
    if (dot(in.d, h.N) < 0) { n1 = 1.0f; n2 = ri; n = h.N; } else { n1 = ri; n2 = 1.0f; n = -h.N; }
    // The reference is built as gnu++98, where pow(float, int) is libstdc++'s __builtin_powif overload
    // (Ray.h:221); calling the builtin here makes the same compiler emit the same multiply chain.
    float energy = 1 - (__builtin_powif(n1, 2) * (1 - __builtin_powif(dot(in.d, n), 2)) / __builtin_powif(n2, 2));
    if (energy < 0) return reflectRay(in, h);
    V3 d_r = n1 * (in.d - n * dot(in.d, n)) / n2 - n * std::sqrt(energy);
    Ray r; r.o = h.P + d_r * kEps; r.d = d_r; return r;
}

inline bool isReflective(const Material& m) { return m.ks.x > 0.f || m.ks.y > 0.f || m.ks.z > 0.f; }   // Material.h:32
inline bool isRefractive(const Material& m) { return m.kt.x > 0.f || m.kt.y > 0.f || m.kt.z > 0.f; }   // Material.h:33
inline bool isDiffuse(const Material& m) { return m.kd.x > 0.f || m.kd.y > 0.f || m.kd.z > 0.f; }      // Phong.cpp:39-42

// Phong::shade, Phong.cpp:44-161.  Phong::diffuse2D returns m_diffuse (Phong.h:20), so the diffuse term
// carries kd twice (diffuseColor * m_diffuse, Phong.cpp:146) -- restated as is.
V3 phongShade(const Ray& ray, const Hit& hit, Counters& cn, long long* shadow_rays)
{
    const Material& mat = g->materials[hit.material];
    V3 L(0.0f, 0.0f, 0.0f);
    const V3 diffuseColor = mat.kd;                                    // Phong::diffuse2D, Phong.h:20
    for (size_t li = 0; li < g->lights.size(); ++li) {
        const Light& lt = g->lights[li];
        const float samples = 1;
        V3 result = lt.color;
        V3 l = (lt.kind == 1) ? -lt.normal : (lt.pos - hit.P);         // getLightDirection: PointLight.h:40-43 / DirectionalAreaLight.h:27-31
        float intensity = 1.f;
        float falloff = length2(l);
        l = l / sqrtf(falloff);
        Ray shadow; shadow.o = hit.P + (l * kEps); shadow.d = l;
        Hit sh;
        if (shadow_rays) (*shadow_rays)++;
        if (sceneTrace(sh, shadow, 0.f, sqrtf(falloff), cn)) {
            if (!isRefractive(g->materials[sh.material])) continue;
            if (dot(sh.N, l) < 0) continue;
            intensity = dot(sh.N, l);
            if (intensity < kEps) continue;
        }
        float nDotL;
        if (lt.kind == 1) {
            V3 lightNormal = lt.normal;
            nDotL = dot(hit.N, -lightNormal);
            float t = dot(lightNormal, lt.pos - hit.P) / -1.0f;
            if (length2((hit.P - t * lightNormal) - lt.pos) > lt.radius * lt.radius) continue;
            falloff = 1.0f / PI;
        } else {
            nDotL = dot(hit.N, l);
            falloff = 1.0f / (falloff * 4.0f * PI * PI);
        }
        L = L + result * (std::max(0.0f, nDotL * falloff * lt.wattage / ((float)samples)) * diffuseColor * mat.kd) * intensity;
        if (mat.shininess < kInf) {
            V3 r = -l + 2 * dot(l, hit.N) * hit.N;
            float eDotr = __builtin_powif(std::max(0.0f, std::min(1.f, dot(-ray.d, r))), 500);   // pow(float,int), gnu++98
            float highlights = std::max(0.0f, eDotr * falloff * lt.wattage / (float)samples);
            L = L + V3(highlights);
        }
    }
    return L;
}

void irradianceEstimate(const PhotonMap* pm, float irrad[3], const float pos[3], const float normal[3], float max_dist, int nphotons, long long* visited);

// Scene::traceScene, Scene.cpp:270-346.  Environment map absent -> background colour (Scene.cpp:657-).
bool traceScene(const Ray& ray, V3& shadeResult, int depth, Counters& cn, long long* nrays)
{
    Hit hitInfo;
    shadeResult = V3(0.f);
    bool hit = false;
    if (depth >= 0) {
        if (nrays) nrays[0]++;
        if (sceneTrace(hitInfo, ray, 0.0f, MIRO_TMAX, cn)) {
            hit = true;
            --depth;
            const Material& mat = g->materials[hitInfo.material];
            shadeResult = phongShade(ray, hitInfo, cn, nrays ? nrays + 1 : 0);
            if (isDiffuse(mat) && g_pms.size() >= 2 && (g_pms[0]->stored > 0 || g_pms[1]->stored > 0)) {
                float pos[3] = {hitInfo.P.x, hitInfo.P.y, hitInfo.P.z};
                float normal[3] = {hitInfo.N.x, hitInfo.N.y, hitInfo.N.z};
                float irr[3] = {0, 0, 0}, cau[3] = {0, 0, 0};
                irradianceEstimate(g_pms[0], irr, pos, normal, 1e10f, 500, 0);
                irradianceEstimate(g_pms[1], cau, pos, normal, 1e10f, 500, 0);
                shadeResult = shadeResult + V3(irr[0] + cau[0], irr[1] + cau[1], irr[2] + cau[2]);
            }
            if (isReflective(mat)) {
                V3 rr; Ray r2 = reflectRay(ray, hitInfo);
                if (traceScene(r2, rr, depth, cn, nrays)) shadeResult = shadeResult + mat.ks * rr;
            }
            if (isRefractive(mat)) {
                float Rs = reflectionCoefficient(ray, hitInfo);
                V3 rr; Ray r2 = reflectRay(ray, hitInfo);
                if (Rs > 0.01) { if (traceScene(r2, rr, depth, cn, nrays)) shadeResult = shadeResult + mat.kt * rr * Rs; }
                V3 fr; Ray r3 = refractRay(ray, hitInfo);
                if (traceScene(r3, fr, depth, cn, nrays)) shadeResult = shadeResult + mat.kt * fr * (1.f - Rs);
            }
        } else {
            shadeResult = g->bg;
            hit = true;
        }
    }
    return hit;
}

// ---- PhotonMap.cpp ---------------------------------------------------------------------------------------
struct NearestPhotons { int max, found, got_heap; float pos[3]; float* dist2; const Photon** index; };

inline void photonDir(const PhotonMap* pm, float* dir, const Photon* p)           // PhotonMap.cpp:68-74
{
    dir[0] = pm->sintheta[p->theta] * pm->cosphi[p->phi];
    dir[1] = pm->sintheta[p->theta] * pm->sinphi[p->phi];
    dir[2] = pm->costheta[p->theta];
}

// Photon_map::locate_photons, PhotonMap.cpp:152-243 (with the reference's direction filter at :186)
void locatePhotons(const PhotonMap* pm, NearestPhotons* const np, const int index, const float normal[3], long long* visited)
{
    const Photon* p = &pm->photons[index];
    float dist1;
    if (visited) (*visited)++;
    if (index < pm->half_stored) {
        dist1 = np->pos[p->plane] - p->pos[p->plane];
        if (dist1 > 0.0) {
            locatePhotons(pm, np, 2 * index + 1, normal, visited);
            if (dist1 * dist1 < np->dist2[0]) locatePhotons(pm, np, 2 * index, normal, visited);
        } else {
            locatePhotons(pm, np, 2 * index, normal, visited);
            if (dist1 * dist1 < np->dist2[0]) locatePhotons(pm, np, 2 * index + 1, normal, visited);
        }
    }
    dist1 = p->pos[0] - np->pos[0];
    float dist2 = dist1 * dist1;
    dist1 = p->pos[1] - np->pos[1];
    dist2 += dist1 * dist1;
    dist1 = p->pos[2] - np->pos[2];
    dist2 += dist1 * dist1;
    float pdir[3];
    photonDir(pm, pdir, p);
    if (dist2 < np->dist2[0] && (pdir[0] * normal[0] + pdir[1] * normal[1] + pdir[2] * normal[2]) < 0.0f) {
        if (np->found < np->max) {
            np->found++;
            np->dist2[np->found] = dist2;
            np->index[np->found] = p;
        } else {
            int j, parent;
            if (np->got_heap == 0) {
                float dst2; const Photon* phot;
                int half_found = np->found >> 1;
                for (int k = half_found; k >= 1; k--) {
                    parent = k; phot = np->index[k]; dst2 = np->dist2[k];
                    while (parent <= half_found) {
                        j = parent + parent;
                        if (j < np->found && np->dist2[j] < np->dist2[j + 1]) j++;
                        if (dst2 >= np->dist2[j]) break;
                        np->dist2[parent] = np->dist2[j]; np->index[parent] = np->index[j];
                        parent = j;
                    }
                    np->dist2[parent] = dst2; np->index[parent] = phot;
                }
                np->got_heap = 1;
            }
            parent = 1; j = 2;
            while (j <= np->found) {
                if (j < np->found && np->dist2[j] < np->dist2[j + 1]) j++;
                if (dist2 > np->dist2[j]) break;
                np->dist2[parent] = np->dist2[j]; np->index[parent] = np->index[j];
                parent = j; j += j;
            }
            np->index[parent] = p; np->dist2[parent] = dist2;
            np->dist2[0] = np->dist2[1];
        }
    }
}

// Photon_map::irradiance_estimate, PhotonMap.cpp:81-145
void irradianceEstimate(const PhotonMap* pm, float irrad[3], const float pos[3], const float normal[3], float max_dist, int nphotons, long long* visited)
{
    irrad[0] = irrad[1] = irrad[2] = 0.0;
    NearestPhotons np;
    std::vector<float> d2(nphotons + 1);
    std::vector<const Photon*> idx(nphotons + 1);
    np.dist2 = d2.data(); np.index = idx.data();
    np.pos[0] = pos[0]; np.pos[1] = pos[1]; np.pos[2] = pos[2];
    np.max = nphotons; np.found = 0; np.got_heap = 0;
    np.dist2[0] = max_dist * max_dist;
    locatePhotons(pm, &np, 1, normal, visited);
    for (int i = 1; i <= np.found; i++) {
        const Photon* p = np.index[i];
        irrad[0] += p->power[0]; irrad[1] += p->power[1]; irrad[2] += p->power[2];
    }
    const float tmp = (1.0f / M_PI) / (np.dist2[0]);                   // double arithmetic, rounded to float (:136)
    irrad[0] *= tmp; irrad[1] *= tmp; irrad[2] *= tmp;
}

// Photon_map::store, PhotonMap.cpp:252-288
void pmStore(PhotonMap* pm, const float power[3], const float pos[3], const float dir[3])
{
    if (pm->stored >= pm->max_photons) return;
    pm->stored++;
    Photon* const node = &pm->photons[pm->stored];
    for (int i = 0; i < 3; i++) {
        node->pos[i] = pos[i];
        if (node->pos[i] < pm->bbox_min[i]) pm->bbox_min[i] = node->pos[i];
        if (node->pos[i] > pm->bbox_max[i]) pm->bbox_max[i] = node->pos[i];
        node->power[i] = power[i];
    }
    int theta = int(acos(dir[2]) * (256.0 / M_PI));
    if (theta > 255) node->theta = 255; else node->theta = (unsigned char)theta;
    int phi = int(atan2(dir[1], dir[0]) * (256.0 / (2.0 * M_PI)));
    if (phi > 255) node->phi = 255; else if (phi < 0) node->phi = (unsigned char)(phi + 256); else node->phi = (unsigned char)phi;
}

// median_split, PhotonMap.cpp:371-402
void medianSplit(Photon** p, const int start, const int end, const int median, const int axis)
{
    int left = start, right = end;
    while (right > left) {
        const float v = p[right]->pos[axis];
        int i = left - 1, j = right;
        for (;;) {
            while (p[++i]->pos[axis] < v) ;
            while (p[--j]->pos[axis] > v && j > left) ;
            if (i >= j) break;
            std::swap(p[i], p[j]);
        }
        std::swap(p[i], p[right]);
        if (i >= median) right = i - 1;
        if (i <= median) left = i + 1;
    }
}

// balance_segment, PhotonMap.cpp:408-477
void balanceSegment(PhotonMap* pm, Photon** pbal, Photon** porg, const int index, const int start, const int end)
{
    int median = 1;
    while ((4 * median) <= (end - start + 1)) median += median;
    if ((3 * median) <= (end - start + 1)) { median += median; median += start - 1; }
    else median = end - median + 1;
    int axis = 2;
    if ((pm->bbox_max[0] - pm->bbox_min[0]) > (pm->bbox_max[1] - pm->bbox_min[1]) &&
        (pm->bbox_max[0] - pm->bbox_min[0]) > (pm->bbox_max[2] - pm->bbox_min[2])) axis = 0;
    else if ((pm->bbox_max[1] - pm->bbox_min[1]) > (pm->bbox_max[2] - pm->bbox_min[2])) axis = 1;
    medianSplit(porg, start, end, median, axis);
    pbal[index] = porg[median];
    pbal[index]->plane = axis;
    if (median > start) {
        if (start < median - 1) {
            const float tmp = pm->bbox_max[axis];
            pm->bbox_max[axis] = pbal[index]->pos[axis];
            balanceSegment(pm, pbal, porg, 2 * index, start, median - 1);
            pm->bbox_max[axis] = tmp;
        } else pbal[2 * index] = porg[start];
    }
    if (median < end) {
        if (median + 1 < end) {
            const float tmp = pm->bbox_min[axis];
            pm->bbox_min[axis] = pbal[index]->pos[axis];
            balanceSegment(pm, pbal, porg, 2 * index + 1, median + 1, end);
            pm->bbox_min[axis] = tmp;
        } else pbal[2 * index + 1] = porg[end];
    }
}

// Photon_map::balance, PhotonMap.cpp:314-359
void pmBalance(PhotonMap* pm)
{
    const int stored = pm->stored;
    if (stored > 1) {
        Photon* photons = pm->photons.data();
        // the reference indexes pa1 up to 2*index+1 of leaf parents: allocate generously
        std::vector<Photon*> pa1((size_t)2 * stored + 4, (Photon*)0), pa2((size_t)stored + 1);
        for (int i = 0; i <= stored; i++) pa2[i] = &photons[i];
        balanceSegment(pm, pa1.data(), pa2.data(), 1, 1, stored);
        int d, j = 1, foo = 1;
        Photon foo_photon = photons[j];
        for (int i = 1; i <= stored; i++) {
            d = (int)(pa1[j] - photons);
            pa1[j] = 0;
            if (d != foo) photons[j] = photons[d];
            else {
                photons[j] = foo_photon;
                if (i < stored) {
                    for (; foo <= stored; foo++) if (pa1[foo] != 0) break;
                    foo_photon = photons[foo];
                    j = foo;
                }
                continue;
            }
            j = d;
        }
    }
    pm->half_stored = stored / 2 - 1;
}

}  // namespace

// =================================================================================================
// C interface -- mirrors oracle/ref_driver.cpp name for name (ref_* -> orc_*) so the same Python test
// helper drives the reference, the oracle and (through mirogpu) the product.
// =================================================================================================

// ---- photon tracing: Scene::tracePhoton (Scene.cpp:526-641) and the emission of Scene::tracePhotons /
// traceCausticPhotons (Scene.cpp:351-472).  The reference draws from an unseeded, racy rand(); every frand() here is a
// counter-based uniform instead, a pure function of (seed, emission index, segment, purpose), the same stream the
// device draws (Philox4x32-10, Salmon et al. SC'11: restated from the published algorithm), so a photon walk can be
// compared emission by emission.  Per segment (1-based depth) one draw yields, in order: the roulette number
// (Scene.cpp:543), the Fresnel coin (:621), and the two numbers of Ray::random (Ray.h:132-133).  The disc rejection
// loop of sampleDisc (Utility.h:82-95) draws pair k = 0, 1, ... of purpose 4.
static void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4])
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int round = 0; round < 10; ++round) {
        const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static void uniform4(uint32_t seed, uint64_t index, uint32_t sample, uint32_t purpose, float u[4])
{
    uint32_t r[4];
    philox4x32_10((uint32_t)index, sample, purpose, (uint32_t)(index >> 32), seed, 0x4D49524Fu, r);
    for (int k = 0; k < 4; ++k) u[k] = (float)(r[k] >> 8) * (1.0f / 16777216.0f);
}

// One emission; returns the number of photons recorded (<= 5), records = {power[3], pos[3], dir[3]} each.
static int tracePhotonWalk(const Light& L, const V3& t1, const V3& t2, bool caustic, uint32_t seed, uint64_t e, float* records)
{
    // Scene.cpp:379-388 (power) / :431-434 (caustic: / 10), DirectionalAreaLight.h:20-35
    V3 power = L.color * L.wattage;
    if (caustic) power = power * (PI * L.radius * L.radius / 10.f); else power = power * (PI * L.radius * L.radius);
    V3 direction = L.normal;
    V3 position;
    {
        float x_rand = 0, y_rand = 0;
        for (uint32_t k = 0; k < 64; ++k) {
            float u[4]; uniform4(seed, e, k, 4, u);
            x_rand = (2 * u[0] - 1) * L.radius;
            y_rand = (2 * u[1] - 1) * L.radius;
            if (!(x_rand * x_rand + y_rand * y_rand > L.radius * L.radius)) break;
        }
        position = L.pos + (x_rand * t1 + y_rand * t2);
    }
    int depth = 0, n = 0;
    Counters cn = {0, 0};
    for (;;) {
        if (depth > 5) return n;                                       // TRACE_DEPTH_PHOTONS, Miro.h:14
        Ray ray; ray.o = position + kEps * direction; ray.d = direction;
        Hit hit;
        ++depth;
        if (!sceneTrace(hit, ray, 0.0f, MIRO_TMAX, cn)) return n;
        const Material& m = g->materials[hit.material];
        float u[4]; uniform4(seed, e, (uint32_t)depth, 3, u);
        const float rnd = u[0];
        const V3 diffuseColor = m.kd;                                  // Phong::diffuse2D returns m_diffuse (Phong.h:20)
        float prob[3];
        prob[0] = average(diffuseColor);
        prob[1] = prob[0] + average(m.ks);
        prob[2] = prob[1] + average(m.kt);
        if (rnd > prob[2]) return n;                                   // absorbed
        if (rnd < prob[0]) {
            if (depth > 1) {                                           // only indirect light is stored
                float* r = records + 9 * n;
                r[0] = power.x; r[1] = power.y; r[2] = power.z; r[3] = hit.P.x; r[4] = hit.P.y; r[5] = hit.P.z;
                r[6] = direction.x; r[7] = direction.y; r[8] = direction.z;
                ++n;
            } else if (caustic) return 0;                              // caustic photons must leave a specular surface first
            const float phi = std::asin(std::sqrt(u[2]));              // Ray::random, Ray.h:124-140
            const float theta = 2.0f * PI * u[3];
            const V3 dir = alignHemisphereToVector(hit.N, theta, phi);
            position = hit.P + kEps * dir; direction = dir;
            power = diffuseColor * power / prob[0];
        } else if (rnd < prob[1]) {
            if (!caustic && depth == 1) return n;
            const Ray refl = reflectRay(ray, hit);
            position = hit.P; direction = refl.d;
        } else if (rnd < prob[2]) {
            if (!caustic && depth == 1) return n;
            const float Rs = reflectionCoefficient(ray, hit);
            if (u[1] < Rs) { const Ray refl = reflectRay(ray, hit); position = hit.P; direction = refl.d; }
            else { const Ray refr = refractRay(ray, hit); position = hit.P; direction = refr.d; }
        } else return n;
    }
}

extern "C" {

void orc_new_scene()
{
    g = new Scene;
    g_pms.clear();
    g_pms.push_back(new PhotonMap(1));
    g_pms.push_back(new PhotonMap(1));
}

int orc_new_material(const float* kd, const float* ks, const float* kt, float shininess, float refr_index)
{
    Material m;                                                        // Phong::Phong, Phong.cpp:13-32
    if (shininess < 0) shininess = kInf;
    m.kd = V3(kd[0], kd[1], kd[2]); m.ks = V3(ks[0], ks[1], ks[2]); m.kt = V3(kt[0], kt[1], kt[2]);
    m.shininess = shininess; m.refr = refr_index;
    for (int i = 0; i < 3; ++i) m.kt[i] = std::max(std::min(m.kt[i], 1.0f - m.ks[i]), 0.f);
    for (int i = 0; i < 3; ++i) m.kd[i] = std::max(std::min(m.kd[i], 1.0f - m.ks[i] - m.kt[i]), 0.f);
    g->materials.push_back(m);
    return (int)g->materials.size() - 1;
}

int orc_add_obj(const char* path, const float* ctm, int material)
{
    M4 m;
    if (ctm) for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) m.m[i][j] = ctm[4 * i + j];
    FILE* fp = fopen(path, "rb");
    if (!fp) return -1;
    Mesh* M = loadObj(fp, m);
    fclose(fp);
    g->meshes.push_back(M);
    for (int i = 0; i < M->ntris; ++i) addPrim((int)g->meshes.size() - 1, i, material);
    return M->ntris;
}

void orc_add_triangle(const float* v9, const float* n9, int material)
{
    Mesh* M = new Mesh;                                                // TriangleMesh::createSingleTriangle, TriangleMeshLoad.cpp:15-56
    for (int k = 0; k < 3; ++k) {
        M->vertices.push_back(V3(v9[3 * k], v9[3 * k + 1], v9[3 * k + 2]));
        M->normals.push_back(V3(n9[3 * k], n9[3 * k + 1], n9[3 * k + 2]));
        M->vidx.push_back(k); M->nidx.push_back(k);
    }
    M->ntris = 1;
    g->meshes.push_back(M);
    addPrim((int)g->meshes.size() - 1, 0, material);
}

// Sphere (Sphere.h) as a bounded scene object / Plane (Plane.h) as an unbounded one.  Plane ids follow the bounded objects.
void orc_add_sphere(const float* center, float radius, int material)
{
    Prim p; p.mesh = -1; p.index = 0; p.material = material; p.center = V3(center[0], center[1], center[2]); p.radius = radius;
    g->prims.push_back(p);
}
void orc_add_plane(const float* normal, const float* origin, int material)
{
    PlaneObj pl; pl.normal = V3(normal[0], normal[1], normal[2]); pl.origin = V3(origin[0], origin[1], origin[2]); pl.material = material;
    g->planes.push_back(pl);
}

void orc_add_point_light(const float* pos, const float* color, float wattage)
{
    Light l; l.kind = 0; l.pos = V3(pos[0], pos[1], pos[2]); l.color = V3(color[0], color[1], color[2]);
    l.normal = V3(0, 1, 0); l.wattage = wattage; l.radius = 0;
    g->lights.push_back(l);
}

void orc_add_directional_light(const float* pos, const float* normal, float radius, const float* color, float wattage)
{
    Light l; l.kind = 1; l.pos = V3(pos[0], pos[1], pos[2]); l.color = V3(color[0], color[1], color[2]);
    l.normal = V3(normal[0], normal[1], normal[2]); l.wattage = wattage; l.radius = radius;
    g->lights.push_back(l);
}

void orc_set_bg_color(const float* c) { g->bg = V3(c[0], c[1], c[2]); }

// Scene::preCalc (Scene.cpp:50-84) minus the photon passes (driven separately through orc_pm_*).
double orc_precalc()
{
    double t = -omp_get_wtime();
    for (size_t i = 0; i < g->prims.size(); ++i) primPreCalc(g->prims[i]);
    std::vector<int> all(g->prims.size());
    for (size_t i = 0; i < all.size(); ++i) all[i] = (int)i;
    g->n_nodes = g->n_leaves = 0;
    g->root = new Node;
    build(g->root, all, 0);
    t += omp_get_wtime();
    return t;
}

int orc_num_objects() { return (int)g->prims.size(); }

void orc_dump_triangles(float* out18)
{
    for (size_t i = 0; i < g->prims.size(); ++i) {
        const Prim& p = g->prims[i];
        float* o = out18 + 18 * i;
        if (p.mesh < 0) { for (int k = 0; k < 18; ++k) o[k] = 0; continue; }
        const Mesh* m = g->meshes[p.mesh];
        for (int k = 0; k < 3; ++k) {
            const V3& v = m->vertices[m->vidx[3 * p.index + k]];
            const V3& n = m->normals[m->nidx[3 * p.index + k]];
            o[3 * k] = v.x; o[3 * k + 1] = v.y; o[3 * k + 2] = v.z;
            o[9 + 3 * k] = n.x; o[9 + 3 * k + 1] = n.y; o[9 + 3 * k + 2] = n.z;
        }
    }
}

void orc_material_ids(int* out) { for (size_t i = 0; i < g->prims.size(); ++i) out[i] = g->prims[i].material; }

// mode 0: the reference's BVH traversal (Scene::trace).  mode 1: exhaustive, min-index tie-break,
// same normal post-processing.  Extra outputs beta/gamma are not produced here: P and N carry them.
static void trace_impl(const float* rays, long n, float* out_t, int* out_id, float* out_P, float* out_N, int nthreads, int mode)
{
    if (nthreads <= 0) nthreads = omp_get_num_procs();
    long long box = 0, tri = 0;
#pragma omp parallel for schedule(dynamic, 1024) num_threads(nthreads) reduction(+ : box, tri)
    for (long i = 0; i < n; ++i) {
        const float* r = rays + 8 * i;
        Ray ray; ray.o = V3(r[0], r[1], r[2]); ray.d = V3(r[4], r[5], r[6]);
        Hit hit; Counters cn = {0, 0};
        bool h;
        if (mode == 0) h = sceneTrace(hit, ray, r[3], r[7], cn);
        else {
            h = bruteTrace(hit, ray, r[3], r[7]);
            if (h) hit.N = normalized(hit.N);
        }
        box += cn.box; tri += cn.tri;
        if (h) {
            out_id[i] = hit.object;
            if (out_t) out_t[i] = hit.t;
            if (out_P) { out_P[3 * i] = hit.P.x; out_P[3 * i + 1] = hit.P.y; out_P[3 * i + 2] = hit.P.z; }
            if (out_N) { out_N[3 * i] = hit.N.x; out_N[3 * i + 1] = hit.N.y; out_N[3 * i + 2] = hit.N.z; }
        } else {
            out_id[i] = -1;
            if (out_t) out_t[i] = hit.t;
            if (out_P) { out_P[3 * i] = out_P[3 * i + 1] = out_P[3 * i + 2] = 0; }
            if (out_N) { out_N[3 * i] = out_N[3 * i + 1] = out_N[3 * i + 2] = 0; }
        }
    }
    g_ray_box += box; g_ray_tri += tri;
}

void orc_trace(const float* rays, long n, float* out_t, int* out_id, float* out_P, float* out_N, int nthreads)
{
    trace_impl(rays, n, out_t, out_id, out_P, out_N, nthreads, 0);
}
void orc_trace_brute(const float* rays, long n, float* out_t, int* out_id, float* out_P, float* out_N, int nthreads)
{
    trace_impl(rays, n, out_t, out_id, out_P, out_N, nthreads, 1);
}

double orc_trace_time(const float* rays, long n, int nthreads, long* out_hits)
{
    if (nthreads <= 0) nthreads = omp_get_num_procs();
    long hits = 0;
    double t = -omp_get_wtime();
#pragma omp parallel for schedule(dynamic, 4096) num_threads(nthreads) reduction(+ : hits)
    for (long i = 0; i < n; ++i) {
        const float* r = rays + 8 * i;
        Ray ray; ray.o = V3(r[0], r[1], r[2]); ray.d = V3(r[4], r[5], r[6]);
        Hit hit; Counters cn = {0, 0};
        if (sceneTrace(hit, ray, r[3], r[7], cn)) hits++;
    }
    t += omp_get_wtime();
    if (out_hits) *out_hits = hits;
    return t;
}

// Timing leg that keeps its answers (bench.py's parity block).
double orc_trace_time_hits(const float* rays, long n, int nthreads, float* out_t, int* out_id)
{
    if (nthreads <= 0) nthreads = omp_get_num_procs();
    double t = -omp_get_wtime();
#pragma omp parallel for schedule(dynamic, 4096) num_threads(nthreads)
    for (long i = 0; i < n; ++i) {
        const float* r = rays + 8 * i;
        Ray ray; ray.o = V3(r[0], r[1], r[2]); ray.d = V3(r[4], r[5], r[6]);
        Hit hit; Counters cn = {0, 0};
        out_id[i] = sceneTrace(hit, ray, r[3], r[7], cn) ? hit.object : -1;
        out_t[i] = hit.t;
    }
    t += omp_get_wtime();
    return t;
}
int orc_host_threads() { return omp_get_num_procs(); }

int orc_stats_get(long long* out9)
{
    for (int i = 0; i < 9; ++i) out9[i] = 0;
    out9[0] = g->n_nodes; out9[1] = g->n_leaves; out9[7] = g_ray_box; out9[8] = g_ray_tri;
    return 1;
}
void orc_stats_reset_rays() { g_ray_box = g_ray_tri = 0; }

void orc_set_camera(const float* eye, const float* lookat, const float* up, float fov)
{
    g->eye = V3(eye[0], eye[1], eye[2]);                                // Camera.h setEye / setUp / setLookAt -> setViewDir
    g->up = normalized(V3(up[0], up[1], up[2]));
    g->viewDir = normalized(V3(lookat[0], lookat[1], lookat[2]) - g->eye);
    g->fov = fov;
}

void orc_eye_rays(int w, int h, float* rays)
{
    CamBasis b = camBasis(w, h);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            Ray r = eyeRay(b, x, y, w, h, 0.5f, 0.5f);
            float* o = rays + 8 * ((long)y * w + x);
            o[0] = r.o.x; o[1] = r.o.y; o[2] = r.o.z; o[3] = 0.0f;
            o[4] = r.d.x; o[5] = r.d.y; o[6] = r.d.z; o[7] = MIRO_TMAX;
        }
}

// Jittered eye rays: jitter holds (dx,dy) per pixel in [0,1) (Camera.cpp:129-133 with the caller's numbers).
void orc_eye_rays_jitter(int w, int h, const float* jitter, float* rays)
{
    CamBasis b = camBasis(w, h);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            long i = (long)y * w + x;
            Ray r = eyeRay(b, x, y, w, h, jitter[2 * i], jitter[2 * i + 1]);
            float* o = rays + 8 * i;
            o[0] = r.o.x; o[1] = r.o.y; o[2] = r.o.z; o[3] = 0.0f;
            o[4] = r.d.x; o[5] = r.d.y; o[6] = r.d.z; o[7] = MIRO_TMAX;
        }
}

// Camera basis as the reference derives it (for the host mirror's test): out = w,u,v (9) + top,right (2)
void orc_camera_basis(int w, int h, float* out11)
{
    CamBasis b = camBasis(w, h);
    out11[0] = b.w.x; out11[1] = b.w.y; out11[2] = b.w.z; out11[3] = b.u.x; out11[4] = b.u.y; out11[5] = b.u.z;
    out11[6] = b.v.x; out11[7] = b.v.y; out11[8] = b.v.z; out11[9] = b.top; out11[10] = b.right;
}

// Ray::diffuse / Ray::random (Ray.h:109-140) with the caller's uniform numbers u1,u2 in place of rand():
// phi = asin(sqrt(u1)), theta = 2*PI*u2, dir = alignHemisphereToVector(N, theta, phi), origin P + eps*dir.
// ids < 0 (miss) produce a zero ray with tMax = -1 (never hits).
void orc_diffuse_rays(const float* P, const float* N, const int* ids, const float* u, long n, float* rays)
{
    for (long i = 0; i < n; ++i) {
        float* o = rays + 8 * i;
        if (ids[i] < 0) { for (int k = 0; k < 8; ++k) o[k] = 0; o[6] = 1.0f; o[7] = -1.0f; continue; }
        float phi = std::asin(std::sqrt(u[2 * i]));
        float theta = 2.0f * PI * u[2 * i + 1];
        V3 nn(N[3 * i], N[3 * i + 1], N[3 * i + 2]);
        V3 dir = alignHemisphereToVector(nn, theta, phi);
        V3 org = V3(P[3 * i], P[3 * i + 1], P[3 * i + 2]) + kEps * dir;
        o[0] = org.x; o[1] = org.y; o[2] = org.z; o[3] = 0.0f;
        o[4] = dir.x; o[5] = dir.y; o[6] = dir.z; o[7] = MIRO_TMAX;
    }
}

// Scene::traceScene per ray (pre-tone-map radiance).  counts (may be NULL): [0] rays through trace() from
// traceScene, [1] shadow rays.
void orc_trace_scene(const float* rays, long n, int depth, float* rgb, int nthreads, long long* counts)
{
    if (nthreads <= 0) nthreads = omp_get_num_procs();
    long long c0 = 0, c1 = 0, box = 0, tri = 0;
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads) reduction(+ : c0, c1, box, tri)
    for (long i = 0; i < n; ++i) {
        const float* r = rays + 8 * i;
        Ray ray; ray.o = V3(r[0], r[1], r[2]); ray.d = V3(r[4], r[5], r[6]);
        V3 c(0.f); Counters cn = {0, 0}; long long nr[2] = {0, 0};
        traceScene(ray, c, depth, cn, nr);
        rgb[3 * i] = c.x; rgb[3 * i + 1] = c.y; rgb[3 * i + 2] = c.z;
        c0 += nr[0]; c1 += nr[1]; box += cn.box; tri += cn.tri;
    }
    g_ray_box += box; g_ray_tri += tri;
    if (counts) { counts[0] = c0; counts[1] = c1; }
}

// Tone map of Scene::raytraceImage's second pass (Scene.cpp:87-91,177-202) + Image::setPixel's Map()
// (Image.cpp:47-52): NaN -> global max, sigmoid(6v-3), truncate 255*v.
void orc_tonemap(const float* rgb, long npix, unsigned char* out8)
{
    float maxI = -kInf;
    for (long i = 0; i < 3 * npix; ++i) if (rgb[i] > maxI) maxI = rgb[i];
    for (long i = 0; i < 3 * npix; ++i) {
        float v = rgb[i];
        if (v != v) v = maxI;
        v = 1 / (1 + std::exp(-(6 * v - 3)));                          // sigmoid, Utility.h:19-22
        float m = 255 * v;
        out8[i] = m > 255 ? 255 : (unsigned char)m;
    }
}


// Emissions [first, first + count) of light `light`: counts[i] photons recorded by emission first + i, records
// [i*45 .. ) = up to five {power, pos, dir} triples.  Powers are unscaled (the caller divides by the number of
// emissions, Scene.cpp:400).
void orc_trace_photons(int light, int caustic, unsigned seed, unsigned long long first, unsigned count, unsigned char* counts,
                       float* records, int nthreads)
{
    const Light& L = g->lights[light];
    V3 t1 = cross(V3(0, 0, 1), L.normal);                               // getTangents, Utility.h:25-31 (SquareLight::preCalc)
    if (length2(t1) < 1e-6) t1 = cross(V3(0, 1, 0), L.normal);
    const V3 t2 = cross(t1, L.normal);
    if (nthreads <= 0) nthreads = omp_get_num_procs();
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads)
    for (long i = 0; i < (long)count; ++i)
        counts[i] = (unsigned char)tracePhotonWalk(L, t1, t2, caustic != 0, seed, first + (unsigned long long)i, records + 45 * i);
}

// ---- photon maps: which 0 / 1 = the scene's global / caustic map, >= 2 standalone -----------------
int orc_pm_new(int max_photons) { g_pms.push_back(new PhotonMap(max_photons)); return (int)g_pms.size() - 1; }
void orc_pm_reset(int which, int max_photons) { delete g_pms[which]; g_pms[which] = new PhotonMap(max_photons); }
void orc_pm_store(int which, const float* power, const float* pos, const float* dir, long n)
{
    for (long i = 0; i < n; ++i) pmStore(g_pms[which], power + 3 * i, pos + 3 * i, dir + 3 * i);
}
void orc_pm_scale(int which, float s)                                  // PhotonMap.cpp:297-306
{
    PhotonMap* pm = g_pms[which];
    for (int i = pm->prev_scale; i <= pm->stored; i++) { pm->photons[i].power[0] *= s; pm->photons[i].power[1] *= s; pm->photons[i].power[2] *= s; }
    pm->prev_scale = pm->stored;
}
void orc_pm_balance(int which) { pmBalance(g_pms[which]); }
int orc_pm_stored(int which) { return g_pms[which]->stored; }
int orc_pm_half_stored(int which) { return g_pms[which]->half_stored; }
int orc_pm_sizeof_photon() { return (int)sizeof(Photon); }
void orc_pm_dump(int which, void* out) { memcpy(out, g_pms[which]->photons.data(), sizeof(Photon) * (size_t)(g_pms[which]->stored + 1)); }
// Load an already balanced heap-ordered array (stored+1 records).
void orc_pm_load(int which, const void* photons, int stored)
{
    PhotonMap* pm = g_pms[which];
    if (pm->max_photons < stored) { delete pm; pm = g_pms[which] = new PhotonMap(stored); }
    memcpy(pm->photons.data(), photons, sizeof(Photon) * (size_t)(stored + 1));
    pm->stored = stored; pm->half_stored = stored / 2 - 1;
}
void orc_pm_irradiance(int which, const float* pos, const float* nrm, long n, float max_dist, int k, float* irr, int nthreads)
{
    if (nthreads <= 0) nthreads = omp_get_num_procs();
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads)
    for (long i = 0; i < n; ++i) irradianceEstimate(g_pms[which], irr + 3 * i, pos + 3 * i, nrm + 3 * i, max_dist, k, 0);
}
// Instrumented: photons visited per query (bytes/query = 28*visited + 48).
long long orc_pm_visited(int which, const float* pos, const float* nrm, long n, float max_dist, int k)
{
    long long v = 0;
    float irr[3];
    for (long i = 0; i < n; ++i) irradianceEstimate(g_pms[which], irr, pos + 3 * i, nrm + 3 * i, max_dist, k, &v);
    return v;
}

}  // extern "C"
