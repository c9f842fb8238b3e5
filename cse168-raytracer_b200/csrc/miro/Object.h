// Object.h -- abstract scene object of the host API layer (reference Object.h:9-47).
#ifndef MIROHOST_OBJECT_H
#define MIROHOST_OBJECT_H
#include <vector>
#include "Miro.h"
#include "Material.h"
#include "Ray.h"

class Object {
public:
    Object() : m_material(0) {}
    virtual ~Object() {}
    void setMaterial(const Material* m) { m_material = m; }
    const Material* getMaterial() const { return m_material; }
    virtual void preCalc() {}
    virtual Vector3 coordsMin() const = 0;
    virtual Vector3 coordsMax() const = 0;
    virtual Vector3 center() const = 0;
    virtual bool isBounded() const { return true; }
    // the texture mapping function of 2-D lookups (Object.h:37); Plane keeps this default (Plane.cpp:50-60)
    virtual tex_coord2d_t toUVCoordinates(const Vector3& xyz) const { return tex_coord2d_t(xyz.x, xyz.z); }
    virtual bool intersect(HitInfo& result, const Ray& ray, float tMin = 0.0f, float tMax = MIRO_TMAX) = 0;
protected:
    const Material* m_material;
};
typedef std::vector<Object*> Objects;
#endif
