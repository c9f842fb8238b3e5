// Triangle.h -- one face of a TriangleMesh as a scene Object (reference Triangle.h).  Triangle::intersect
// on the host is the single-ray convenience; the batched path runs on the device with identical arithmetic.
#ifndef MIROHOST_TRIANGLE_H
#define MIROHOST_TRIANGLE_H
#include "Object.h"
#include "TriangleMesh.h"

class Triangle : public Object {
public:
    Triangle(TriangleMesh* m = 0, unsigned int i = 0) : m_mesh(m), m_index(i) {}
    virtual Vector3 coordsMin() const { return m_cachedMin; }
    virtual Vector3 coordsMax() const { return m_cachedMax; }
    virtual Vector3 center() const;
    virtual void preCalc();
    void setIndex(unsigned int i) { m_index = i; }
    unsigned int getIndex() { return m_index; }
    void setMesh(TriangleMesh* m) { m_mesh = m; }
    TriangleMesh* getMesh() { return m_mesh; }
    virtual bool intersect(HitInfo& result, const Ray& ray, float tMin = 0.0f, float tMax = MIRO_TMAX);
    virtual tex_coord2d_t toUVCoordinates(const Vector3& xyz) const;   // Triangle.cpp:172-222
    // Fills result.P / N / material from barycentrics exactly as Triangle.cpp:160-166 does.
    void fillHit(HitInfo& result, float t, float beta, float gamma) const;
protected:
    TriangleMesh* m_mesh;
    unsigned int m_index;
    Vector3 m_cachedMin, m_cachedMax;
};
#endif
