// Ray.h -- Ray and HitInfo of the host API layer (reference Ray.h:21-84).  The secondary-ray generators
// of the reference (diffuse / reflect / refract) run on the device; the host types only carry data.
#ifndef MIROHOST_RAY_H
#define MIROHOST_RAY_H
#include "Vector3.h"
#include "Material.h"

class HitInfo {
public:
    float t;                   // hit distance
    Vector3 P;                 // hit point
    Vector3 N;                 // shading normal
    const Material* material;  // material of the intersected object
    const Object* object;      // the intersected object
    explicit HitInfo(float t = 0.0f, const Vector3& P = Vector3(), const Vector3& N = Vector3(0.0f, 1.0f, 0.0f))
        : t(t), P(P), N(N), material(0), object(0) {}
};

class Ray {
public:
    bool isDiffuse;
    Vector3 o, d;
    Ray() : isDiffuse(false), o(), d(Vector3(0.0f, 0.0f, 1.0f)) {}
    Ray(const Vector3& o, const Vector3& d) : isDiffuse(false), o(o), d(d) {}
};
#endif
