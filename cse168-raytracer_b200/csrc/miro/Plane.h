// Plane.h -- the reference's Plane (Plane.h:12-36) in the host API layer.  Unbounded: Scene keeps it outside the tree
// (Scene.h:22-25) and the device tests it after the walk (mirogpu_plane); intersect() is Plane.cpp:33-48.
#ifndef MIROHOST_PLANE_H
#define MIROHOST_PLANE_H
#include <limits>
#include "Object.h"

class Plane : public Object {
public:
    Plane() : m_normal(0, 1, 0), m_origin(0, 0, 0) {}
    virtual ~Plane() {}
    virtual Vector3 coordsMin() const { return -Vector3(infinity); }
    virtual Vector3 coordsMax() const { return Vector3(infinity); }
    virtual Vector3 center() const { return m_origin; }
    void setNormal(Vector3 normal) { m_normal = normal; }
    void setOrigin(Vector3 origin) { m_origin = origin; }
    const Vector3& normal() const { return m_normal; }
    const Vector3& origin() const { return m_origin; }
    virtual bool isBounded() const { return false; }
    virtual bool intersect(HitInfo& result, const Ray& ray, float tMin = 0.0f, float tMax = MIRO_TMAX);
    void fillHit(HitInfo& result, const Ray& ray, float t) const;
protected:
    Vector3 m_normal, m_origin;
};
#endif
