// render_impl.cuh -- included by mirogpu.cu after mirogpu_scene is complete.  (Filled in below.)
#ifndef MIROGPU_RENDER_IMPL_CUH
#define MIROGPU_RENDER_IMPL_CUH
namespace {
int render_device(mirogpu_scene*, const mirogpu_camera&, const mirogpu_render_params&, float*, cudaStream_t, std::string& err)
{
    err = "mirogpu_render: not implemented yet";
    return MIROGPU_ERR_UNSUPPORTED;
}
int render_host(mirogpu_scene*, const mirogpu_camera&, const mirogpu_render_params&, float*, std::string& err)
{
    err = "mirogpu_render: not implemented yet";
    return MIROGPU_ERR_UNSUPPORTED;
}
}  // namespace
namespace mirogpu {
int PhotonMapDevice::upload(const void*, int, std::string& err) { err = "photon upload: not implemented yet"; return MIROGPU_ERR_UNSUPPORTED; }
cudaError_t photon_gather_launch(const PhotonMapDevice&, const float*, const float*, size_t, float, int, float*, cudaStream_t) { return cudaErrorNotSupported; }
}
#endif
