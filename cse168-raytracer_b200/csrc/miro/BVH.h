// BVH.h -- the reference's BVH interface (BVH.h:29-62) backed by the device engine.
//   build()      gathers the Triangle objects of the list, hands per-triangle arrays to
//                mirogpu_scene_create (host SAH build -> flat GPU layout -> HBM) and keeps the handle.
//   intersect()  single-ray convenience over the batched device query; same result contract as the
//                reference (minHit.t = tMax on a miss, t/P/N/material/object on a hit).
//   intersectBatch()  the throughput path: n rays per call, host buffers in and out.
#ifndef MIROHOST_BVH_H
#define MIROHOST_BVH_H
#include <vector>
#include "Miro.h"
#include "Object.h"
#include "../../../include/mirogpu.h"

class BVH {
public:
    BVH() : m_handle(0), m_objects(0), m_unbounded(0), m_layout(MIROGPU_LAYOUT_QBVH4), m_builder(MIROGPU_BUILDER_SAH_HOST) {}
    ~BVH();
    void build(Objects* objs, int depth = 0);
    bool intersect(HitInfo& result, const Ray& ray, float tMin = 0.0f, float tMax = MIRO_TMAX) const;
    bool intersectChildren(HitInfo& result, const Ray& ray, float tMin, float tMax) const { return intersect(result, ray, tMin, tMax); }
    // results[i] is filled like intersect() would; returns the number of hits.
    size_t intersectBatch(const Ray* rays, size_t n, HitInfo* results, bool* hitFlags, float tMin = 0.0f, float tMax = MIRO_TMAX) const;
    void setLayout(int layout) { m_layout = layout; }
    void setBuilder(int builder) { m_builder = builder; }   // MIROGPU_BUILDER_*: host SAH (default) or device LBVH
    // Scene's unbounded objects (planes): not part of the tree, but answered by the same device query (Scene.cpp:219-230).
    void setUnbounded(const Objects* unbounded) { m_unbounded = unbounded; }
    // The CUDA devices the scene is replicated on (SURVEY 8b device_mask): image rows of raytraceImage shard over them.
    // Default: environment MIROGPU_DEVICES ("all" or a comma list), else the current device.
    void setDevices(const std::vector<int>& devices) { m_devices = devices; }
    bool onDevice(const Object* o) const;   // true: the device query already accounts for this object
    mirogpu_handle handle() const { return m_handle; }
    const std::vector<Object*>& fallbackObjects() const { return m_other; }
protected:
    bool finish(HitInfo& result, const mirogpu_hit& h, const Ray& ray, float tMin, float tMax) const;
    mirogpu_handle m_handle;
    Objects* m_objects;               // borrowed, as in the reference (BVH.cpp:84)
    const Objects* m_unbounded;       // borrowed (Scene::m_unboundedObjects)
    std::vector<Object*> m_prims;     // device prim id -> object: triangles, then spheres, then planes
    uint32_t m_ntris, m_nspheres;
    std::vector<Object*> m_other;     // objects of other classes: tested on the host after the device query
    std::vector<int> m_devices;
    int m_layout, m_builder;
};
#endif
