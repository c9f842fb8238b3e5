// Scene.h -- the reference's Scene interface (Scene.h:14-69) over the device engine.
//   preCalc()        Object::preCalc for every object, then BVH::build (= flatten + upload)
//   trace()          Scene::trace: BVH query + unbounded objects + the N normalisation of Scene.cpp:262
//   traceBatch()     the same for n rays per call
//   raytraceImage()  whole frame on the device (mirogpu_render), tone map, Image::setPixel
#ifndef MIROHOST_SCENE_H
#define MIROHOST_SCENE_H
#include "Miro.h"
#include "Object.h"
#include "PointLight.h"
#include "BVH.h"
#include "PhotonMap.h"

class Scene {
public:
    Scene();
    ~Scene();
    void addObject(Object* pObj)
    {
        if (pObj->isBounded()) m_objects.push_back(pObj); else m_unboundedObjects.push_back(pObj);
    }
    const Objects* objects() const { return &m_objects; }
    const Objects* unboundedObjects() const { return &m_unboundedObjects; }
    void addLight(PointLight* pObj) { m_lights.push_back(pObj); }
    const Lights* lights() const { return &m_lights; }
    void preCalc();
    void raytraceImage(Camera* cam, Image* img);
    bool trace(HitInfo& minHit, const Ray& ray, float tMin = 0.0f, float tMax = MIRO_TMAX) const;
    size_t traceBatch(const Ray* rays, size_t n, HitInfo* results, bool* hitFlags, float tMin = 0.0f, float tMax = MIRO_TMAX) const;
    void setBgColor(Vector3 color) { m_bgColor = color; }
    BVH& bvh() { return m_bvh; }
    Photon_map& photonMap() { return *m_photonMap; }
    Photon_map& causticMap() { return *m_causticMap; }
    void setPhotonMapsEnabled(bool on) { m_usePhotonMaps = on; }
    // Scene::tracePhotons / traceCausticPhotons (Scene.cpp:351-472): emit from every DirectionalAreaLight until the
    // target number of photons is stored, scale by 1 / emissions, balance, upload.  The walks run on the device
    // (mirogpu_photon_trace); store / scale / balance are the host's Photon_map.  Both are called by preCalc(), like
    // the reference; a target of 0 only balances.  Return value: emissions consumed (the reference's totalPhotons).
    long tracePhotons();
    long traceCausticPhotons();
    // Scene.h:67-68 (static const in the reference)
    int PhotonsPerLightSource, CausticPhotonsPerLightSource;
    // render controls the reference fixes at compile time (Miro.h:13-15, -DDISABLE_SHADOWS)
    int renderSpp, renderJitter, renderMode, renderShadows;
    unsigned renderSeed;
    double lastRenderSeconds;
protected:
    long tracePhotonPass(Photon_map& map, int which, int target, bool caustic);
    void postProcess(HitInfo& minHit) const;
    Objects m_objects, m_unboundedObjects;
    Photon_map* m_photonMap;
    Photon_map* m_causticMap;
    BVH m_bvh;
    Lights m_lights;
    Vector3 m_bgColor;
    bool m_usePhotonMaps;
};
#endif
