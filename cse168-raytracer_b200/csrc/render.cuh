// render.cuh -- declarations for the wavefront renderer (Scene::raytraceImage on the device).
#ifndef MIROGPU_RENDER_CUH
#define MIROGPU_RENDER_CUH
#include <cuda_runtime.h>
#include <cstdint>
#include <cstddef>

namespace mirogpu {

// Device buffers reused across render calls on one handle (grown on demand, never shrunk).
struct RenderScratch {
    void* buf[16] = {};
    size_t cap[16] = {};
    cudaError_t ensure(int i, size_t bytes)
    {
        if (cap[i] >= bytes) return cudaSuccess;
        cudaFree(buf[i]); buf[i] = nullptr; cap[i] = 0;
        cudaError_t e = cudaMalloc(&buf[i], bytes);
        if (e == cudaSuccess) cap[i] = bytes;
        return e;
    }
    void release()
    {
        for (int i = 0; i < 16; ++i) { cudaFree(buf[i]); buf[i] = nullptr; cap[i] = 0; }
    }
};

}  // namespace mirogpu
#endif
