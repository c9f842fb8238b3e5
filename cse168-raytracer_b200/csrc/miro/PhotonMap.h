// PhotonMap.h -- Photon_map of the host API layer (reference PhotonMap.h; H. W. Jensen's layout).
// Two ways to fill a map: (1) Scene::tracePhotons / traceCausticPhotons build it on the device in one call
// (mirogpu_photon_pass: emission, store, scale, balance -- SURVEY 8f-2) and this object adopts it, mirroring the
// 28-byte Photon array to the host only when data() / store() ask for it; (2) the caller store()s photons here, and
// balance() sends the array through the device balance (mirogpu_photon_balance: the reference's heap order, ties
// included) before attach() uploads it for irradiance_estimate.
#ifndef MIROHOST_PHOTONMAP_H
#define MIROHOST_PHOTONMAP_H
#include <vector>
#include "../../../include/mirogpu.h"

typedef struct Photon {
    float pos[3];
    short plane;
    unsigned char theta, phi;
    float power[3];
} Photon;

class Photon_map {
public:
    Photon_map(int max_phot);
    ~Photon_map();
    void store(const float power[3], const float pos[3], const float dir[3]);
    void scale_photon_power(const float scale);
    void balance(void);
    // Device gather through the scene handle given to attach(); one query per call (convenience) ...
    void irradiance_estimate(float irrad[3], const float pos[3], const float normal[3], const float max_dist, const int nphotons) const;
    // ... or n queries per call (throughput path).
    void irradiance_estimate_batch(float* irrad3, const float* pos3, const float* normal3, size_t n, float max_dist, int nphotons) const;
    void attach(mirogpu_handle h, int which);   // uploads the balanced array to the device as map `which`
    int stored() const { return stored_photons; }
    bool balanced() const { return m_balanced; }
    const Photon* data() const { syncHost(); return photons; }
    void adoptDevice(mirogpu_handle h, int which, int stored);   // the map mirogpu_photon_pass built as map `which` of h
    static int balanceDevice;                                     // CUDA device balance() runs on (default 0)
private:
    void syncHost() const;
    Photon* photons;
    int stored_photons, half_stored_photons, max_photons, prev_scale;
    float bbox_min[3], bbox_max[3];
    mirogpu_handle m_handle;
    int m_which;
    bool m_balanced, m_on_device;
    mutable bool m_host_stale;
};
#endif
