// Camera.h -- look-at pinhole camera of the host API layer (reference Camera.h / Camera.cpp).
#ifndef MIROHOST_CAMERA_H
#define MIROHOST_CAMERA_H
#include <cfloat>
#include "Miro.h"
#include "Ray.h"
#include "../../../include/mirogpu.h"

class Camera {
public:
    Camera() : m_bgColor(0, 0, 0), m_eye(0, 0, 0), m_up(0, 1, 0), m_viewDir(0, 0, -1), m_fov(45.f) {}
    virtual ~Camera() {}
    void click(Scene* pScene, Image* pImage);   // renders pScene into pImage (Camera.cpp:38-70, ray-trace branch)
    void setEye(float x, float y, float z) { m_eye.set(x, y, z); }
    void setEye(const Vector3& e) { m_eye = e; }
    void setUp(float x, float y, float z) { m_up.set(x, y, z); m_up.normalize(); }
    void setUp(const Vector3& u) { setUp(u.x, u.y, u.z); }
    void setViewDir(float x, float y, float z) { m_viewDir.set(x, y, z); m_viewDir.normalize(); }
    void setViewDir(const Vector3& v) { setViewDir(v.x, v.y, v.z); }
    void setLookAt(float x, float y, float z) { Vector3 d = Vector3(x, y, z) - m_eye; setViewDir(d); }
    void setLookAt(const Vector3& l) { setLookAt(l.x, l.y, l.z); }
    void setBGColor(float x, float y, float z) { m_bgColor.set(x, y, z); }
    void setBGColor(const Vector3& c) { m_bgColor = c; }
    void setFOV(float fov) { m_fov = fov; }
    float fov() const { return m_fov; }
    const Vector3& viewDir() const { return m_viewDir; }
    const Vector3& up() const { return m_up; }
    const Vector3& eye() const { return m_eye; }
    const Vector3& bgColor() const { return m_bgColor; }
    // Unlike the reference (function statics, Camera.cpp:106-125) the basis is derived per call, so several
    // cameras can live in one process.
    Ray eyeRay(int x, int y, int imageWidth, int imageHeight, bool randomize);
    mirogpu_camera abi() const;
private:
    Vector3 m_bgColor, m_eye, m_up, m_viewDir;
    float m_fov;
};
#endif
