// photon.cuh -- photon-map kNN gather on the device (Photon_map::irradiance_estimate / locate_photons,
// reference PhotonMap.cpp:81-243) against the reference's own heap-ordered left-balanced kd-tree.
#ifndef MIROGPU_PHOTON_CUH
#define MIROGPU_PHOTON_CUH
#include <cuda_runtime.h>
#include <cstdint>
#include <cstddef>
#include <string>
#include "../../include/mirogpu.h"

#define MIRO_PHOTON_KMAX 512
#define MIRO_GW_TICKET_SPAN 8192   /* segment cursors of one warp-per-query launch: one per warp of the grid (the grid is capped accordingly) */
#define MIRO_GW_TICKET_SLOTS 8      /* launches in flight */

namespace mirogpu {

// One map under construction on the device (photon_build.cu): the photons in store order, 1-based like the reference's array.
struct PhotonBuild {
    float4* pos = nullptr;        // (x, y, z, word 3 of the reference's Photon: plane | theta << 16 | phi << 24)
    float4* pow = nullptr;        // (r, g, b, 0)
    uint32_t* bbox = nullptr;     // bbox_min xyz, bbox_max xyz as order-preserving words (PhotonMap.cpp:36-39, 264-267)
    int* ctl = nullptr;           // [0] emissions of the last batch consumed, [1] photons stored after it
    uint32_t* excl = nullptr;     // exclusive prefix of the batch's per-emission record counts
    void* scan_tmp = nullptr;
    size_t scan_bytes = 0;
    uint32_t excl_cap = 0;
    int cap = 0;
    cudaError_t alloc(int capacity);
    void release();
    // Photon_map::store over one batch of k_photon_trace records, in emission order, while fewer than `target` photons are
    // stored (Scene.cpp:370-377); returns (after a stream sync) how many emissions were consumed and the new photon count.
    cudaError_t store_batch(const unsigned char* d_counts, const float* d_records, uint32_t batch, int base, int target, cudaStream_t st,
                            int* used, int* stored_after);
    cudaError_t scale(int first, int last, float s, cudaStream_t st);   // scale_photon_power (PhotonMap.cpp:297-306)
    ~PhotonBuild() { release(); }
    PhotonBuild() = default;
    PhotonBuild(const PhotonBuild&) = delete;
    PhotonBuild& operator=(const PhotonBuild&) = delete;
};

struct PhotonMapDevice {
    float4* d_photons = nullptr;  // 2 float4 per photon: (pos.xyz, plane|theta<<8|phi<<16 bits) (power.xyz, 0)
    float* d_tables = nullptr;    // costheta[256] sintheta[256] cosphi[256] sinphi[256]  (PhotonMap.cpp:47-53)
    float4* d_search = nullptr;   // warp-per-query search records, 32 B per photon: (pos.xyz, same bits) (direction.xyz from the tables, 0)
    unsigned int* d_tickets = nullptr;   // segment cursors of the warp-per-query launches (MIRO_GW_TICKET_SLOTS arrays of MIRO_GW_TICKET_SPAN)
    int stored = 0, half_stored = 0;
    bool exact = false;           // true: the reference's search verbatim, one query per thread (bit-identical estimates); see photon_impl.cuh
    int upload(const void* photons28, int stored, std::string& err);
    // device to device: balance (PhotonMap.cpp:314-466, same heap array) the photons of `b` and lay out the gather records
    int build_from_store(const PhotonBuild& b, int n, bool balance, cudaStream_t st, std::string& err);
    // copy of another device's map (multi-device handles: the pass runs once, the result is replicated)
    int clone_from(const PhotonMapDevice& src, int src_device, int dst_device, std::string& err);
    cudaError_t export28(void* photons28, cudaStream_t st) const;   // the map as the reference's Photon array (host mirror)
    void release()
    {
        cudaFree(d_photons); cudaFree(d_tables); cudaFree(d_search); cudaFree(d_tickets);
        d_photons = nullptr; d_tables = nullptr; d_search = nullptr; d_tickets = nullptr; stored = half_stored = 0;
    }
};

// active (may be NULL): per-query float4 whose .w == 0 marks a query to skip (its irradiance is written as 0).
cudaError_t photon_gather_launch(const PhotonMapDevice& pm, const float* d_pos3, const float* d_normal3, size_t n,
                                 float max_dist, int k, float* d_irrad3, cudaStream_t st, const float4* active = nullptr,
                                 const uint32_t* d_n = nullptr);

// Photon_map::balance for a host-filled array of n + 1 reference Photon records (in place); lo3 / hi3 = bbox_min / bbox_max.
cudaError_t photon_balance_host_array(void* photons28, int n, const float* lo3, const float* hi3, cudaStream_t st);
cudaError_t photon_balance_device(const float4* d_pos, int n, const uint32_t* d_bbox, const float* lo3, const float* hi3, uint32_t* d_heap,
                                  cudaStream_t st);

}  // namespace mirogpu
#endif
