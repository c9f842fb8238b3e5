// Texture.h -- the procedural textures and TexturedPhong of the host API layer (reference Texture.h:62-318, Texture.cpp:358-537).
// A texture is its constructor arguments: evaluation happens on the device inside the shading kernels (csrc/texture.cuh:
// Perlin's improved noise, Worley's cellular basis, the reference's colour formulas); the lookup2D / lookup3D / bumpHeight2D
// members answer single queries through the same code compiled for the host (mirogpu_texture_lookup / _bump).
// Not provided: LoadedTexture (image files through FreeImage), CellularTexture2D (unused by the reference's scenes),
// CloudTexture (its lookup never reaches a material: it overrides lookup2D with a 3-D argument, Texture.h:151).
#ifndef MIROHOST_TEXTURE_H
#define MIROHOST_TEXTURE_H
#include <cstring>
#include "Phong.h"
#include "../../../include/mirogpu.h"

class Texture {
public:
    virtual ~Texture() {}
    virtual LookupCoordinates GetLookupCoordinates() const = 0;
    virtual int deviceKind() const = 0;                       // MIROGPU_TEX_*
    virtual void deviceParams(float tp[12]) const = 0;        // constructor arguments in mirogpu_material::tex order
    virtual float bumpHeight2D(const tex_coord2d_t& c) const
    {
        float tp[12]; memset(tp, 0, sizeof tp); deviceParams(tp);
        float h = 0.f; mirogpu_texture_bump(deviceKind(), tp, c.u, c.v, &h); return h;
    }
    virtual float bumpHeight3D(const tex_coord3d_t&) const { return 0; }
    virtual Vector3 lookup2D(const tex_coord2d_t& c) const { return GetLookupCoordinates() == UV ? eval(c.u, c.v, 0.f) : Vector3(0, 0, 0); }
    virtual Vector3 lookup3D(const tex_coord3d_t& c) const { return GetLookupCoordinates() == UVW ? eval(c.u, c.v, c.w) : Vector3(0, 0, 0); }
    virtual Vector3 lowresLookup2D(const tex_coord2d_t& c) const { return lookup2D(c); }
    virtual Vector3 lowresLookup3D(const tex_coord2d_t& c) const { return lookup2D(c); }
protected:
    Vector3 eval(float u, float v, float w) const
    {
        float tp[12]; memset(tp, 0, sizeof tp); deviceParams(tp);
        float rgb[3] = {0.f, 0.f, 0.f};
        mirogpu_texture_lookup(deviceKind(), tp, u, v, w, rgb);
        return Vector3(rgb[0], rgb[1], rgb[2]);
    }
};
class Texture2D : public Texture { public: virtual LookupCoordinates GetLookupCoordinates() const { return UV; } };
class Texture3D : public Texture { public: virtual LookupCoordinates GetLookupCoordinates() const { return UVW; } };

class CheckerBoardTexture : public Texture2D {
public:
    CheckerBoardTexture(Vector3 color1 = Vector3(1), Vector3 color2 = Vector3(0), float scale = 1) : m_scale(scale), m_color1(color1), m_color2(color2) {}
    virtual int deviceKind() const { return MIROGPU_TEX_CHECKER; }
    virtual void deviceParams(float tp[12]) const { for (int k = 0; k < 3; ++k) { tp[k] = m_color1[k]; tp[3 + k] = m_color2[k]; } tp[6] = m_scale; }
protected:
    float m_scale; Vector3 m_color1, m_color2;
};
class StoneTexture : public Texture2D {
public:
    StoneTexture(float scale = 1) : m_scale(scale) {}
    virtual int deviceKind() const { return MIROGPU_TEX_STONE; }
    virtual void deviceParams(float tp[12]) const { tp[0] = m_scale; }
protected:
    float m_scale;
};
class StemTexture : public Texture2D {
public:
    StemTexture(float scale = 1) : m_scale(scale) {}
    virtual int deviceKind() const { return MIROGPU_TEX_STEM; }
    virtual void deviceParams(float tp[12]) const { tp[0] = m_scale; }
    virtual float bumpHeight2D(const tex_coord2d_t&) const { return 0.0f; }
protected:
    float m_scale;
};
class PetalTexture : public Texture3D {
public:
    PetalTexture(const Vector3& Pivot, float Radius = 1, float scale = 1) : m_scale(scale), m_radius(Radius), m_pivot(Pivot) {}
    virtual int deviceKind() const { return MIROGPU_TEX_PETAL; }
    virtual void deviceParams(float tp[12]) const { for (int k = 0; k < 3; ++k) tp[k] = m_pivot[k]; tp[3] = m_radius; tp[4] = m_scale; }
protected:
    float m_scale, m_radius; Vector3 m_pivot;
};
class LeafTexture : public Texture3D {
public:
    LeafTexture(const Vector3& pivot, const Vector3& direction, float scale = 1) : m_scale(scale), m_direction(direction), m_pivot(pivot) { m_direction.normalize(); }
    virtual int deviceKind() const { return MIROGPU_TEX_LEAF; }
    virtual void deviceParams(float tp[12]) const { tp[0] = m_scale; }
protected:
    float m_scale; Vector3 m_direction, m_pivot;
};
class FlowerCenterTexture : public Texture3D {
public:
    FlowerCenterTexture(const Vector3& Pivot, float Radius = 1, float scale = 1) : m_scale(scale), m_radius(Radius), m_pivot(Pivot) {}
    virtual int deviceKind() const { return MIROGPU_TEX_FLOWER_CENTER; }
    virtual void deviceParams(float tp[12]) const { for (int k = 0; k < 3; ++k) tp[k] = m_pivot[k]; tp[3] = m_radius; }
protected:
    float m_scale, m_radius; Vector3 m_pivot;
};

// Shading model that also does textures (Texture.h:299-318); its Phong diffuse term is Vector3(1) before the energy clamp
// (Texture.cpp:513-517).
class TexturedPhong : public Phong {
public:
    TexturedPhong(Texture* texture, const Vector3& specularColor = Vector3(0), const Vector3& transparentColor = Vector3(0),
                  const float shinyness = 1.0f, const float refractIndex = 1)
        : Phong(Vector3(1.f), specularColor, transparentColor, shinyness, refractIndex), m_texture(texture) {}
    virtual LookupCoordinates GetLookupCoordinates() const { return m_texture->GetLookupCoordinates(); }
    virtual Vector3 diffuse2D(const tex_coord2d_t& c) const { return m_texture->lookup2D(c); }
    virtual Vector3 diffuse3D(const tex_coord3d_t& c) const { return m_texture->lookup3D(c); }
    virtual float bumpHeight2D(const tex_coord2d_t& c) const { return m_texture->bumpHeight2D(c); }
    virtual float bumpHeight3D(const tex_coord3d_t& c) const { return m_texture->bumpHeight3D(c); }
    virtual void describeTexture(mirogpu_material& m) const { m.texture = m_texture->deviceKind(); m_texture->deviceParams(m.tex); }
protected:
    Texture* m_texture;
};
#endif
