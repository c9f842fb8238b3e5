// PhotonMap.h -- Photon_map of the host API layer (reference PhotonMap.h; H. W. Jensen's layout).
// store / scale_photon_power / balance stay on the host (balance is SURVEY 8f-2, "next"); the balanced,
// heap-ordered 28-byte Photon array is uploaded once and irradiance_estimate runs on the device.
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
    const Photon* data() const { return photons; }
private:
    void balance_segment(Photon** pbal, Photon** porg, const int index, const int start, const int end);
    void balance_segment_box(Photon** pbal, Photon** porg, int index, int start, int end, const float* lo3, const float* hi3);
    void median_split(Photon** p, const int start, const int end, const int median, const int axis);
    Photon* photons;
    int stored_photons, half_stored_photons, max_photons, prev_scale;
    float bbox_min[3], bbox_max[3];
    mirogpu_handle m_handle;
    int m_which;
    bool m_balanced;
};
#endif
