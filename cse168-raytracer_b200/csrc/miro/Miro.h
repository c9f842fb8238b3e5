// Miro.h -- constants of the host API layer (same names and values as the reference's Miro.h:8-20).
#ifndef MIROHOST_MIRO_H
#define MIROHOST_MIRO_H
#include <limits>

const float MIRO_TMAX = 1e12f;
const float epsilon = 1e-4f;
const float PI = 3.1415926535897932384626433832795028841972f;
const float DegToRad = PI / 180.0f;
const float RadToDeg = 180.0f / PI;
const float TRACE_DEPTH = 10;
const float TRACE_DEPTH_PHOTONS = 5;
const float TRACE_SAMPLES = 1000;
const float PHOTON_MAX_DIST = 1e10;
const float PHOTON_SAMPLES = 500.f;
const float infinity = std::numeric_limits<float>::infinity();

// Miro.h:34-46
typedef struct tex_coord2d_s {
    tex_coord2d_s() : u(0), v(0) {}
    tex_coord2d_s(float _u, float _v) : u(_u), v(_v) {}
    float u, v;
} tex_coord2d_t;
typedef struct tex_coord3d_s {
    tex_coord3d_s() : u(0), v(0), w(0) {}
    tex_coord3d_s(float _u, float _v, float _w) : u(_u), v(_v), w(_w) {}
    float u, v, w;
} tex_coord3d_t;

class Ray;
class HitInfo;
class Object;
class Triangle;
class TriangleMesh;
class PointLight;
class Camera;
class Image;
class Scene;
class Material;

extern Camera* g_camera;
extern Scene* g_scene;
extern Image* g_image;
#endif
