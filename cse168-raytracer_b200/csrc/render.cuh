// render.cuh -- declarations for the wavefront renderer (Scene::raytraceImage on the device).
#ifndef MIROGPU_RENDER_CUH
#define MIROGPU_RENDER_CUH
#include <cuda_runtime.h>
#include <cstdint>
#include <cstddef>

namespace mirogpu {

// Device buffers reused across render calls on one handle (grown on demand, never shrunk).
struct RenderScratch {
    void* buf[20] = {};
    size_t cap[20] = {};
    // the fused diffuse-bounce path runs the two halves of a 16-sample batch on two streams (render_impl.cuh)
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaError_t ensure_side()
    {
        if (side) return cudaSuccess;
        cudaError_t e = cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming);
        return e;
    }
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
        for (int i = 0; i < 20; ++i) { cudaFree(buf[i]); buf[i] = nullptr; cap[i] = 0; }
        if (side) cudaStreamDestroy(side);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
        side = nullptr; ev_fork = ev_join = nullptr;
    }
};

}  // namespace mirogpu
#endif
