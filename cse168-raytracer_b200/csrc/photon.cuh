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

struct PhotonMapDevice {
    float4* d_photons = nullptr;  // 2 float4 per photon: (pos.xyz, plane|theta<<8|phi<<16 bits) (power.xyz, 0)
    float* d_tables = nullptr;    // costheta[256] sintheta[256] cosphi[256] sinphi[256]  (PhotonMap.cpp:47-53)
    float4* d_search = nullptr;   // warp-per-query search records, 32 B per photon: (pos.xyz, same bits) (direction.xyz from the tables, 0)
    unsigned int* d_tickets = nullptr;   // segment cursors of the warp-per-query launches (MIRO_GW_TICKET_SLOTS arrays of MIRO_GW_TICKET_SPAN)
    int stored = 0, half_stored = 0;
    bool exact = false;           // true: the reference's search verbatim, one query per thread (bit-identical estimates); see photon_impl.cuh
    int upload(const void* photons28, int stored, std::string& err);
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

}  // namespace mirogpu
#endif
