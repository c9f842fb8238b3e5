// Utility.h -- sampling helpers of the reference (Utility.h:14-108) for host code that calls them (scene scripts, light
// classes).  Same arithmetic and the same rand() stream; the device has its own counter-based versions (csrc/rng.cuh).
#ifndef MIROHOST_UTILITY_H
#define MIROHOST_UTILITY_H
#include <cmath>
#include <cstdlib>
#include "Material.h"
#include "Vector3.h"

double getTime();

inline float frand() { return (float)rand() / (float)RAND_MAX; }
inline float sigmoid(float x) { return 1 / (1 + exp(-x)); }

inline void getTangents(const Vector3& normal, Vector3& t1, Vector3& t2)
{
    t1 = cross(Vector3(0, 0, 1), normal);
    if (t1.length2() < 1e-6) t1 = cross(Vector3(0, 1, 0), normal);
    t2 = cross(t1, normal);
}

inline Vector3 alignHemisphereToVector(const Vector3& v, float theta, float phi)
{
    const float u1 = sin(phi) * cos(theta), u2 = sin(phi) * sin(theta), u3 = cos(phi);
    Vector3 t1 = cross(Vector3(0, 0, 1), v);
    if (t1.length2() < 1e-6) t1 = cross(Vector3(0, 1, 0), v);
    Vector3 aligned_d(u1 * t1 + u2 * cross(t1, v) + u3 * v);
    aligned_d.normalize();
    return aligned_d;
}

inline Vector3 sampleSphericalDirection()
{
    float x, y, z;
    do {
        x = 2 * frand() - 1; y = 2 * frand() - 1; z = 2 * frand() - 1;
    } while (x * x + y * y + z * z > 1.0f);
    return Vector3(x, y, z).normalize();
}

inline VectorR2 sampleDisc(float radius)
{
    float x_rand, y_rand;
    do {
        x_rand = (2 * frand() - 1) * radius;
        y_rand = (2 * frand() - 1) * radius;
    } while (x_rand * x_rand + y_rand * y_rand > radius * radius);
    VectorR2 v; v.x = x_rand; v.y = y_rand;
    return v;
}
#endif
