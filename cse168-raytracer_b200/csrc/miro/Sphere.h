// Sphere.h -- the reference's Sphere (Sphere.h:7-38) in the host API layer.  A bounded object: BVH::build hands it to the
// device as a leaf primitive of the tree (mirogpu_sphere); intersect() is the single-ray host form with the reference's
// arithmetic (Sphere.cpp:28-69).
#ifndef MIROHOST_SPHERE_H
#define MIROHOST_SPHERE_H
#include "Vector3.h"
#include "Object.h"

class Sphere : public Object {
public:
    Sphere() : m_center(0.f), m_radius(1.f) {}
    virtual ~Sphere() {}
    void setCenter(const Vector3& v) { m_center = v; }
    void setRadius(const float f) { m_radius = f; }
    float radius() const { return m_radius; }
    virtual Vector3 coordsMin() const { return m_center - Vector3(m_radius); }
    virtual Vector3 coordsMax() const { return m_center + Vector3(m_radius); }
    virtual Vector3 center() const { return m_center; }
    virtual bool intersect(HitInfo& result, const Ray& ray, float tMin = 0.0f, float tMax = MIRO_TMAX);
    virtual tex_coord2d_t toUVCoordinates(const Vector3& xyz) const;   // Sphere.cpp:83-95
    // P = o + t d, N = (P - c).normalize(), material (Sphere.cpp:62-66), for a hit the device found at distance t
    void fillHit(HitInfo& result, const Ray& ray, float t) const;
protected:
    Vector3 m_center;
    float m_radius;
};
#endif
