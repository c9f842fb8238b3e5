// Material.h / Phong / Lambert of the host API layer: plain parameter holders.  Shading itself runs on the
// device (mirogpu_render); the host keeps the reference's constructor semantics (energy clamp,
// Phong.cpp:13-32) and the predicates the tracer branches on (Material.h:32-34, Phong.cpp:39-42).
#ifndef MIROHOST_MATERIAL_H
#define MIROHOST_MATERIAL_H
#include <algorithm>
#include "Miro.h"
#include "Vector3.h"

struct mirogpu_material;

// Material.h:7-15: how the shader queries the material for its colour
enum LookupCoordinates { UV = 0, UVW = 1 };

class Material {
public:
    Material() : m_specular(0.f), m_transmission(0.f), m_refractIndex(1.f), m_shininess(infinity) {}
    virtual ~Material() {}
    virtual LookupCoordinates GetLookupCoordinates() const { return UV; }
    virtual Vector3 diffuse2D(const tex_coord2d_t&) const { return Vector3(0); }
    virtual Vector3 diffuse3D(const tex_coord3d_t&) const { return Vector3(0); }
    virtual float bumpHeight2D(const tex_coord2d_t&) const { return 0; }
    virtual float bumpHeight3D(const tex_coord3d_t&) const { return 0; }
    // the texture part of this material's device description (TexturedPhong fills it in; Texture.h)
    virtual void describeTexture(mirogpu_material&) const {}
    bool isReflective() const { return m_specular.x > 0.f || m_specular.y > 0.f || m_specular.z > 0.f; }
    bool isRefractive() const { return m_transmission.x > 0.f || m_transmission.y > 0.f || m_transmission.z > 0.f; }
    virtual bool isDiffuse() const { return true; }
    Vector3 getReflection() const { return m_specular; }
    Vector3 getRefraction() const { return m_transmission; }
    virtual Vector3 getDiffuse() const { return Vector3(1.f); }
    float getRefractionIndex() const { return m_refractIndex; }
    float getShininess() const { return m_shininess; }
    void setReflection(const Vector3& r) { m_specular = r; }
    void setRefraction(const Vector3& t, float index) { m_refractIndex = index; m_transmission = t; }
    void setShininess(float s) { m_shininess = s; }
    virtual void preCalc() {}
protected:
    Vector3 m_specular, m_transmission;
    float m_refractIndex, m_shininess;
};

class Phong : public Material {
public:
    Phong(const Vector3& kd = Vector3(1), const Vector3& ks = Vector3(0), const Vector3& kt = Vector3(0),
          const float shininess = 1.f, const float refractIndex = 1)
    {
        m_diffuse = kd; m_specular = ks; m_transmission = kt; m_shininess = shininess; m_refractIndex = refractIndex;
        for (int i = 0; i < 3; ++i) m_transmission[i] = std::max(std::min(m_transmission[i], 1.0f - m_specular[i]), 0.f);
        for (int i = 0; i < 3; ++i) m_diffuse[i] = std::max(std::min(m_diffuse[i], 1.0f - m_specular[i] - m_transmission[i]), 0.f);
    }
    virtual bool isDiffuse() const { return m_diffuse.x > 0.f || m_diffuse.y > 0.f || m_diffuse.z > 0.f; }
    virtual Vector3 diffuse2D(const tex_coord2d_t&) const { return m_diffuse; }   // Phong.h:20-21
    virtual Vector3 diffuse3D(const tex_coord3d_t&) const { return m_diffuse; }
    virtual Vector3 getDiffuse() const { return m_diffuse; }
    void setDiffuse(const Vector3& kd) { m_diffuse = kd; }
protected:
    Vector3 m_diffuse;
};
typedef Phong Lambert;
#endif
