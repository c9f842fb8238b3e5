// PointLight.h -- light parameter holders of the host API layer (reference PointLight.h, SquareLight.h,
// DirectionalAreaLight.h).  Only what crosses the C ABI as mirogpu_light is kept.
#ifndef MIROHOST_POINTLIGHT_H
#define MIROHOST_POINTLIGHT_H
#include <vector>
#include "Vector3.h"

class PointLight {
public:
    PointLight() : m_position(0.f), m_color(1.f), m_wattage(100.f) {}
    virtual ~PointLight() {}
    void setPosition(const Vector3& v) { m_position = v; }
    void setColor(const Vector3& v) { m_color = v; }
    void setWattage(float f) { m_wattage = f; }
    float wattage() const { return m_wattage; }
    const Vector3& color() const { return m_color; }
    const Vector3& position() const { return m_position; }
    virtual void preCalc() {}
protected:
    Vector3 m_position, m_color;
    float m_wattage;
};

class SquareLight : public PointLight {
public:
    SquareLight() : m_normal(0, 1, 0) { m_dimensions[0] = m_dimensions[1] = 1.f; }
    void setNormal(Vector3 n) { m_normal = n; }
    Vector3 getNormal() { return m_normal; }
    void setDimensions(float width, float height) { m_dimensions[0] = width; m_dimensions[1] = height; }
protected:
    Vector3 m_normal;
    float m_dimensions[2];
};

class DirectionalAreaLight : public SquareLight {
public:
    DirectionalAreaLight(float radius = 1) : m_radius(radius) {}
    virtual float getRadius() { return m_radius; }
protected:
    float m_radius;
};
typedef std::vector<PointLight*> Lights;
#endif
