"""B200-native ray-intersection engine for the Miro ray tracer -- Python front end.

The product is ``libmirogpu.so`` (C ABI in ``include/mirogpu.h``, hand-written sm_100a kernels).  This
module only binds it with ctypes and lends it PyTorch's device memory, streams and ``torch.distributed``
for plumbing; no intersection work is done in Python or PyTorch.  If the shared library is missing the
import fails loudly -- there is no CPU or eager fallback.

Reference interfaces mirrored here (hallgeirl/cse168-raytracer):
    Scene.preCalc / BVH::build      -> MiroScene(...)                     Scene.cpp:50-84, BVH.cpp:60-339
    Scene::trace / BVH::intersect   -> MiroScene.intersect(...)           Scene.cpp:214-268, BVH.cpp:438-658
    Camera::eyeRay                  -> MiroScene.generate_primary(...)    Camera.cpp:104-161
    Ray::diffuse                    -> MiroScene.generate_bounce(...)     Ray.h:109-122
    Scene::raytraceImage            -> MiroScene.render(...)              Scene.cpp:93-212
    Photon_map::irradiance_estimate -> MiroScene.photon_gather(...)       PhotonMap.cpp:81-145
"""
import ctypes
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libmirogpu.so")

MISS = 0xFFFFFFFF
TMAX = np.float32(1e12)
LAYOUT_BVH2, LAYOUT_CWBVH8, LAYOUT_BVH4, LAYOUT_QBVH4 = 0, 1, 2, 3
BUILDER_SAH_HOST, BUILDER_LBVH_DEVICE, BUILDER_PLOC_DEVICE = 0, 1, 2
CLOSEST_HIT, ANY_HIT = 0, 1
HINT_COHERENT = 0x100   # or-ed into a query mode: camera-like batch -> packet kernel
RENDER_WHITTED, RENDER_DIFFUSE_BOUNCE, RENDER_PRIMARY_ONLY = 0, 1, 2

RAY_DTYPE = np.dtype([("o", np.float32, 3), ("tmin", np.float32), ("d", np.float32, 3), ("tmax", np.float32)])
HIT_DTYPE = np.dtype([("t", np.float32), ("prim_id", np.uint32), ("beta", np.float32), ("gamma", np.float32)])
TEX_NONE, TEX_CHECKER, TEX_STONE, TEX_STEM, TEX_PETAL, TEX_LEAF, TEX_FLOWER_CENTER = range(7)   # MIROGPU_TEX_*
PHOTON_DTYPE = np.dtype([("pos", np.float32, 3), ("plane", np.int16), ("theta", np.uint8), ("phi", np.uint8), ("power", np.float32, 3)])
assert RAY_DTYPE.itemsize == 32 and HIT_DTYPE.itemsize == 16 and PHOTON_DTYPE.itemsize == 28


class MiroGpuError(RuntimeError):
    pass


class BuildOptions(ctypes.Structure):
    _fields_ = [("layout", ctypes.c_int32), ("max_leaf", ctypes.c_int32), ("sah_bins", ctypes.c_int32), ("device", ctypes.c_int32),
                ("builder", ctypes.c_int32)]


class SceneInfo(ctypes.Structure):
    _fields_ = [("num_triangles", ctypes.c_uint32), ("num_nodes", ctypes.c_uint32), ("num_binary_nodes", ctypes.c_uint32),
                ("num_binary_leaves", ctypes.c_uint32), ("max_depth", ctypes.c_uint32), ("layout", ctypes.c_int32),
                ("builder", ctypes.c_int32), ("_pad", ctypes.c_int32), ("node_bytes", ctypes.c_uint64), ("triangle_bytes", ctypes.c_uint64), ("shading_bytes", ctypes.c_uint64),
                ("build_seconds", ctypes.c_double), ("flatten_seconds", ctypes.c_double), ("upload_seconds", ctypes.c_double),
                ("bounds_min", ctypes.c_float * 3), ("bounds_max", ctypes.c_float * 3)]


class Material(ctypes.Structure):
    _fields_ = [("kd", ctypes.c_float * 3), ("ks", ctypes.c_float * 3), ("kt", ctypes.c_float * 3),
                ("shininess", ctypes.c_float), ("refract_index", ctypes.c_float), ("texture", ctypes.c_int32), ("tex", ctypes.c_float * 12)]


class Light(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("position", ctypes.c_float * 3), ("color", ctypes.c_float * 3),
                ("wattage", ctypes.c_float), ("normal", ctypes.c_float * 3), ("radius", ctypes.c_float)]


class Camera(ctypes.Structure):
    _fields_ = [("eye", ctypes.c_float * 3), ("up", ctypes.c_float * 3), ("view_dir", ctypes.c_float * 3), ("fov_degrees", ctypes.c_float)]


class RenderParams(ctypes.Structure):
    _fields_ = [("width", ctypes.c_int32), ("height", ctypes.c_int32), ("spp", ctypes.c_int32), ("jitter", ctypes.c_int32),
                ("max_depth", ctypes.c_int32), ("mode", ctypes.c_int32), ("seed", ctypes.c_uint32), ("tonemap", ctypes.c_int32),
                ("row_begin", ctypes.c_int32), ("row_end", ctypes.c_int32), ("row_stride", ctypes.c_int32), ("row_phase", ctypes.c_int32),
                ("bg_color", ctypes.c_float * 3), ("use_photon_maps", ctypes.c_int32), ("shadows", ctypes.c_int32)]


class Counters(ctypes.Structure):
    _fields_ = [("rays", ctypes.c_uint64), ("node_visits", ctypes.c_uint64), ("box_tests", ctypes.c_uint64),
                ("triangle_tests", ctypes.c_uint64), ("hits", ctypes.c_uint64), ("bytes_fetched", ctypes.c_uint64)]


EXPORTS = [
    "mirogpu_version", "mirogpu_last_error", "mirogpu_device_count", "mirogpu_scene_create", "mirogpu_scene_destroy",
    "mirogpu_scene_info_get", "mirogpu_scene_set_lights", "mirogpu_debug_copy_nodes", "mirogpu_debug_copy_triangles",
    "mirogpu_intersect_batch", "mirogpu_intersect_batch_device", "mirogpu_intersect_batch_counted", "mirogpu_set_kernel_variant",
    "mirogpu_resolve_hits_device", "mirogpu_generate_primary_device", "mirogpu_generate_bounce_device", "mirogpu_rng_uniforms",
    "mirogpu_render", "mirogpu_render_rgb8", "mirogpu_tonemap_rgb8_device", "mirogpu_render_device", "mirogpu_last_call_stats", "mirogpu_photon_upload", "mirogpu_photon_gather",
    "mirogpu_photon_gather_device", "mirogpu_photon_trace", "mirogpu_photon_set_exact", "mirogpu_host_alloc", "mirogpu_host_free",
    "mirogpu_frame_max_device", "mirogpu_tonemap_rows_rgb8_device", "mirogpu_release_build_scratch",
    "mirogpu_scene_create_ex", "mirogpu_scene_devices", "mirogpu_resolve_hits_rays_device",
    "mirogpu_photon_pass", "mirogpu_photon_download", "mirogpu_photon_balance", "mirogpu_texture_lookup", "mirogpu_texture_bump",
]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "The engine has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    lib.mirogpu_last_error.restype = ctypes.c_char_p
    lib.mirogpu_host_alloc.restype = ctypes.c_void_p
    lib.mirogpu_host_alloc.argtypes = [ctypes.c_size_t]
    lib.mirogpu_host_free.argtypes = [ctypes.c_void_p]
    lib.mirogpu_release_build_scratch.restype = None
    return lib


lib = _load()


def _check(rc):
    if rc != 0:
        raise MiroGpuError(f"mirogpu error {rc}: {lib.mirogpu_last_error().decode()}")


def _ptr(a):
    """Raw address of a numpy array / torch tensor / None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return ctypes.c_void_p(a.ctypes.data)
    return ctypes.c_void_p(a.data_ptr())  # torch tensor


def _stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def make_camera(eye, lookat, up=(0, 1, 0), fov=45.0):
    """Camera::setEye / setUp / setLookAt (Camera.h:80-124): up and view_dir are normalised in binary32."""
    eye = np.asarray(eye, np.float32)
    up = np.asarray(up, np.float32)
    vd = np.asarray(lookat, np.float32) - eye

    def _norm(v):
        l2 = np.float32(v[0] * v[0]) + np.float32(v[1] * v[1])
        l2 = np.float32(l2 + np.float32(v[2] * v[2]))
        inv = np.float32(1.0) / np.sqrt(l2, dtype=np.float32)
        return (v * inv).astype(np.float32)

    c = Camera()
    c.eye[:] = [float(x) for x in eye]
    c.up[:] = [float(x) for x in _norm(up)]
    c.view_dir[:] = [float(x) for x in _norm(vd)]
    c.fov_degrees = float(fov)
    return c


def phong(kd=(1, 1, 1), ks=(0, 0, 0), kt=(0, 0, 0), shininess=1.0, refract_index=1.0):
    """Phong's constructor with its energy clamp (Phong.cpp:13-32), in binary32."""
    f = np.float32
    kd = np.asarray(kd, f).copy(); ks = np.asarray(ks, f).copy(); kt = np.asarray(kt, f).copy()
    kt = np.maximum(np.minimum(kt, f(1.0) - ks), f(0))
    kd = np.maximum(np.minimum(kd, f(1.0) - ks - kt), f(0))
    m = Material()
    m.kd[:] = [float(x) for x in kd]; m.ks[:] = [float(x) for x in ks]; m.kt[:] = [float(x) for x in kt]
    m.shininess = float(shininess) if shininess >= 0 else float("inf")
    m.refract_index = float(refract_index)
    return m


def texture_lookup(kind, params, coords):
    """Texture::lookup2D / lookup3D of a procedural texture at coords (n, 2) or (n, 3) -> (n, 3) colours (host evaluation of the
    device code, for checkers)."""
    coords = np.ascontiguousarray(coords, np.float32)
    tp = (ctypes.c_float * 12)(*([float(x) for x in params] + [0.0] * (12 - len(params))))
    out = np.zeros((coords.shape[0], 3), np.float32)
    rgb = (ctypes.c_float * 3)()
    for i in range(coords.shape[0]):
        w = float(coords[i, 2]) if coords.shape[1] > 2 else 0.0
        _check(lib.mirogpu_texture_lookup(int(kind), tp, ctypes.c_float(coords[i, 0]), ctypes.c_float(coords[i, 1]), ctypes.c_float(w), rgb))
        out[i] = rgb[:]
    return out


def texture_bump(kind, params, coords):
    coords = np.ascontiguousarray(coords, np.float32)
    tp = (ctypes.c_float * 12)(*([float(x) for x in params] + [0.0] * (12 - len(params))))
    out = np.zeros(coords.shape[0], np.float32)
    h = ctypes.c_float(0)
    for i in range(coords.shape[0]):
        _check(lib.mirogpu_texture_bump(int(kind), tp, ctypes.c_float(coords[i, 0]), ctypes.c_float(coords[i, 1]), ctypes.byref(h)))
        out[i] = h.value
    return out


def photon_balance(photons, bbox_min, bbox_max, device=0):
    """Photon_map::balance (PhotonMap.cpp:314-466) on the device for a store-order Photon array (entry 0 unused); returns the
    heap-ordered copy."""
    out = np.ascontiguousarray(photons).copy()
    assert out.dtype == PHOTON_DTYPE
    lo = np.ascontiguousarray(bbox_min, dtype=np.float32); hi = np.ascontiguousarray(bbox_max, dtype=np.float32)
    _check(lib.mirogpu_photon_balance(int(device), _ptr(out), int(out.shape[0] - 1), _ptr(lo), _ptr(hi)))
    return out


class MiroScene:
    """One scene resident in HBM on the current CUDA device (replicated per rank in multi-GPU runs)."""

    def __init__(self, tri_vertices, tri_normals=None, material_ids=None, materials=None, layout=LAYOUT_QBVH4,
                 max_leaf=0, sah_bins=32, device=-1, builder=BUILDER_SAH_HOST):
        v = np.ascontiguousarray(tri_vertices, np.float32).reshape(-1, 9)
        n = None if tri_normals is None else np.ascontiguousarray(tri_normals, np.float32).reshape(-1, 9)
        m = None if material_ids is None else np.ascontiguousarray(material_ids, np.uint32).reshape(-1)
        if n is not None and n.shape[0] != v.shape[0]:
            raise ValueError("tri_normals must have one 9-float row per triangle")
        if m is not None and m.shape[0] != v.shape[0]:
            raise ValueError("material_ids must have one entry per triangle")
        mats, nmats = None, 0
        if materials:
            mats = (Material * len(materials))(*materials)
            nmats = len(materials)
        opt = BuildOptions(int(layout), int(max_leaf), int(sah_bins), int(device), int(builder))
        self._h = ctypes.c_void_p()
        _check(lib.mirogpu_scene_create(_ptr(v), _ptr(n), _ptr(m), ctypes.c_uint32(v.shape[0]), mats, ctypes.c_uint32(nmats),
                                        ctypes.byref(opt), ctypes.byref(self._h)))
        self.num_triangles = v.shape[0]

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib.mirogpu_scene_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- introspection ------------------------------------------------------------------------------
    @property
    def info(self):
        i = SceneInfo()
        _check(lib.mirogpu_scene_info_get(self._h, ctypes.byref(i)))
        return i

    def nodes_bytes(self):
        n = ctypes.c_uint64(0)
        _check(lib.mirogpu_debug_copy_nodes(self._h, None, ctypes.byref(n)))
        out = np.zeros(n.value, np.uint8)
        _check(lib.mirogpu_debug_copy_nodes(self._h, _ptr(out), ctypes.byref(n)))
        return out

    def triangles_bytes(self):
        n = ctypes.c_uint64(0)
        _check(lib.mirogpu_debug_copy_triangles(self._h, None, ctypes.byref(n)))
        out = np.zeros(n.value, np.uint8)
        _check(lib.mirogpu_debug_copy_triangles(self._h, _ptr(out), ctypes.byref(n)))
        return out

    def devices(self):
        """CUDA device ordinals the scene is replicated on (mirogpu_scene_devices)."""
        n = ctypes.c_uint32(0)
        _check(lib.mirogpu_scene_devices(self._h, None, ctypes.c_uint32(0), ctypes.byref(n)))
        arr = (ctypes.c_int32 * n.value)()
        _check(lib.mirogpu_scene_devices(self._h, arr, ctypes.c_uint32(n.value), ctypes.byref(n)))
        return list(arr)

    def set_kernel_variant(self, v):
        _check(lib.mirogpu_set_kernel_variant(self._h, int(v)))

    def set_lights(self, lights):
        arr = (Light * len(lights))(*lights) if lights else None
        _check(lib.mirogpu_scene_set_lights(self._h, arr, ctypes.c_uint32(len(lights))))

    def last_call_stats(self):
        r, k = ctypes.c_uint64(0), ctypes.c_uint64(0)
        _check(lib.mirogpu_last_call_stats(self._h, ctypes.byref(r), ctypes.byref(k)))
        return r.value, k.value

    # ---- BVH::intersect over a batch ----------------------------------------------------------------
    def intersect(self, rays, mode=CLOSEST_HIT, out=None):
        """HOST buffers: rays (n,8) float32 numpy (or pinned torch CPU tensor) -> hits structured array."""
        if isinstance(rays, np.ndarray):
            rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
            n = rays.shape[0]
            hits = np.zeros(n, HIT_DTYPE) if out is None else out
            _check(lib.mirogpu_intersect_batch(self._h, _ptr(rays), ctypes.c_size_t(n), _ptr(hits), int(mode)))
            return hits
        n = rays.shape[0]
        assert out is not None and not rays.is_cuda and not out.is_cuda
        _check(lib.mirogpu_intersect_batch(self._h, _ptr(rays), ctypes.c_size_t(n), _ptr(out), int(mode)))
        return out

    def intersect_counted(self, rays, mode=CLOSEST_HIT):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        hits = np.zeros(rays.shape[0], HIT_DTYPE)
        c = Counters()
        _check(lib.mirogpu_intersect_batch_counted(self._h, _ptr(rays), ctypes.c_size_t(rays.shape[0]), _ptr(hits), int(mode), ctypes.byref(c)))
        return hits, c

    def intersect_device(self, d_rays, d_hits, mode=CLOSEST_HIT, n=None):
        """DEVICE tensors (torch, float32 (n,8) and (n,4)); asynchronous on torch's current stream."""
        n = d_rays.shape[0] if n is None else n
        _check(lib.mirogpu_intersect_batch_device(self._h, _ptr(d_rays), ctypes.c_size_t(n), _ptr(d_hits), int(mode), _stream()))

    def resolve_hits_device(self, d_hits, d_P=None, d_N=None, d_mat=None, d_rays=None):
        """P, N (normalised as Scene::trace leaves it), material id per hit.  d_rays is required when the scene holds spheres / planes."""
        if d_rays is not None:
            _check(lib.mirogpu_resolve_hits_rays_device(self._h, _ptr(d_rays), _ptr(d_hits), ctypes.c_size_t(d_hits.shape[0]), _ptr(d_P), _ptr(d_N), _ptr(d_mat), _stream()))
        else:
            _check(lib.mirogpu_resolve_hits_device(self._h, _ptr(d_hits), ctypes.c_size_t(d_hits.shape[0]), _ptr(d_P), _ptr(d_N), _ptr(d_mat), _stream()))

    # ---- device ray generation ----------------------------------------------------------------------
    def generate_primary(self, cam, width, height, d_rays, rows=None, jitter=0, seed=168, sample=0, samples=1):
        """rows = (row_begin, row_end, row_stride, row_phase); output is sample-major then row-major."""
        rb, re_, rs, rp = rows if rows is not None else (0, height, 1, 0)
        _check(lib.mirogpu_generate_primary_device(self._h, ctypes.byref(cam), int(width), int(height), int(rb), int(re_), int(rs), int(rp),
                                                   int(jitter), ctypes.c_uint32(seed), ctypes.c_uint32(sample), ctypes.c_uint32(samples),
                                                   _ptr(d_rays), _stream()))

    def generate_bounce(self, d_rays, d_hits, d_out, seed=168, sample=0, n=None, index_base=0, d_live_count=None):
        n = d_rays.shape[0] if n is None else n
        _check(lib.mirogpu_generate_bounce_device(self._h, _ptr(d_rays), _ptr(d_hits), ctypes.c_size_t(n), ctypes.c_uint32(seed),
                                                  ctypes.c_uint32(sample), ctypes.c_uint32(index_base), _ptr(d_out), _ptr(d_live_count), _stream()))

    # ---- Scene::raytraceImage ------------------------------------------------------------------------
    def render_params(self, width, height, spp=1, jitter=0, max_depth=10, mode=RENDER_WHITTED, seed=168, tonemap=0,
                      rows=None, bg=(0, 0, 0), use_photon_maps=0, shadows=1):
        p = RenderParams()
        p.width, p.height, p.spp, p.jitter, p.max_depth, p.mode = int(width), int(height), int(spp), int(jitter), int(max_depth), int(mode)
        p.seed, p.tonemap = int(seed), int(tonemap)
        rb, re_, rs, rp = rows if rows is not None else (0, height, 1, 0)
        p.row_begin, p.row_end, p.row_stride, p.row_phase = int(rb), int(re_), int(rs), int(rp)
        p.bg_color[:] = [float(x) for x in bg]
        p.use_photon_maps = int(use_photon_maps)
        p.shadows = int(shadows)
        return p

    def render(self, cam, params, out=None):
        """HOST framebuffer (height, width, 3) float32, row 0 = bottom."""
        if out is None:
            out = np.zeros((params.height, params.width, 3), np.float32)
        _check(lib.mirogpu_render(self._h, ctypes.byref(cam), ctypes.byref(params), _ptr(out)))
        return out

    def render_rgb8(self, cam, params, out=None):
        """HOST 8-bit framebuffer (height, width, 3), tone-mapped like the reference's Image; row 0 = bottom."""
        if out is None:
            out = np.zeros((params.height, params.width, 3), np.uint8)
        _check(lib.mirogpu_render_rgb8(self._h, ctypes.byref(cam), ctypes.byref(params), _ptr(out)))
        return out

    def tonemap_rgb8_device(self, d_rgb, d_rgb8):
        h, w = d_rgb.shape[0], d_rgb.shape[1]
        _check(lib.mirogpu_tonemap_rgb8_device(self._h, _ptr(d_rgb), int(w), int(h), _ptr(d_rgb8), _stream()))

    def frame_max_device(self, d_rgb, rows, d_max):
        """Largest non-NaN value over this shard's rows of the full-frame float buffer -> d_max (1 float, device)."""
        h, w = d_rgb.shape[0], d_rgb.shape[1]
        rb, re_, rs, rp = rows
        _check(lib.mirogpu_frame_max_device(self._h, _ptr(d_rgb), int(w), int(h), int(rb), int(re_), int(rs), int(rp), _ptr(d_max), _stream()))

    def tonemap_rows_rgb8_device(self, d_rgb, rows, d_max, d_rgb8):
        """Tone map + 8-bit conversion of this shard's rows with the frame-wide maximum d_max (see mirogpu.h)."""
        h, w = d_rgb.shape[0], d_rgb.shape[1]
        rb, re_, rs, rp = rows
        _check(lib.mirogpu_tonemap_rows_rgb8_device(self._h, _ptr(d_rgb), int(w), int(h), int(rb), int(re_), int(rs), int(rp), _ptr(d_max), _ptr(d_rgb8), _stream()))

    def render_device(self, cam, params, d_rgb):
        _check(lib.mirogpu_render_device(self._h, ctypes.byref(cam), ctypes.byref(params), _ptr(d_rgb), _stream()))

    # ---- photon map ----------------------------------------------------------------------------------
    def photon_upload(self, which, photons):
        photons = np.ascontiguousarray(photons)
        assert photons.dtype == PHOTON_DTYPE
        _check(lib.mirogpu_photon_upload(self._h, int(which), _ptr(photons), int(photons.shape[0] - 1)))

    def photon_pass(self, which, caustic, seed, target, max_emissions=0):
        """Scene::tracePhotons / traceCausticPhotons on the device (emit, store, scale, balance); returns (emissions, stored)."""
        em = ctypes.c_longlong(0); st = ctypes.c_int(0)
        _check(lib.mirogpu_photon_pass(self._h, int(which), int(caustic), ctypes.c_uint32(seed), int(target), ctypes.c_longlong(max_emissions),
                                       ctypes.byref(em), ctypes.byref(st)))
        return int(em.value), int(st.value)

    def photon_download(self, which):
        """Map `which` as the reference's Photon array (stored + 1 records, heap order, entry 0 unused)."""
        st = ctypes.c_int(0)
        _check(lib.mirogpu_photon_download(self._h, int(which), None, 0, ctypes.byref(st)))
        out = np.zeros(st.value + 1, dtype=PHOTON_DTYPE)
        if st.value:
            _check(lib.mirogpu_photon_download(self._h, int(which), _ptr(out), int(st.value), ctypes.byref(st)))
        return out

    def photon_set_exact(self, which, exact):
        """exact=True: the reference's search verbatim, one query per thread (bit-identical); default False: one query per warp."""
        _check(lib.mirogpu_photon_set_exact(self._h, int(which), int(bool(exact))))

    def photon_gather(self, which, pos, normal, max_dist=1e10, k=500):
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
        normal = np.ascontiguousarray(normal, np.float32).reshape(-1, 3)
        irr = np.zeros_like(pos)
        _check(lib.mirogpu_photon_gather(self._h, int(which), _ptr(pos), _ptr(normal), ctypes.c_size_t(pos.shape[0]),
                                         ctypes.c_float(max_dist), int(k), _ptr(irr)))
        return irr

    def photon_trace(self, light_index, caustic, seed, first, count):
        """Scene::tracePhoton for emissions [first, first+count): returns (counts uint8 (count,), records float32 (count, 5, 9))."""
        counts = np.zeros(count, np.uint8)
        records = np.zeros((count, 5, 9), np.float32)
        _check(lib.mirogpu_photon_trace(self._h, int(light_index), int(caustic), ctypes.c_uint32(seed), ctypes.c_uint64(first),
                                        ctypes.c_uint32(count), _ptr(counts), _ptr(records)))
        return counts, records

    def photon_gather_device(self, which, d_pos, d_normal, d_irr, max_dist=1e10, k=500):
        _check(lib.mirogpu_photon_gather_device(self._h, int(which), _ptr(d_pos), _ptr(d_normal), ctypes.c_size_t(d_pos.shape[0]),
                                                ctypes.c_float(max_dist), int(k), _ptr(d_irr), _stream()))


def rng_uniforms(seed, sample, dimension, first, n):
    """The (u1,u2) pairs the device generators draw for elements first..first+n (host evaluation of the same Philox)."""
    out = np.zeros((n, 2), np.float32)
    _check(lib.mirogpu_rng_uniforms(ctypes.c_uint32(seed), ctypes.c_uint32(sample), ctypes.c_uint32(dimension),
                                    ctypes.c_size_t(first), ctypes.c_size_t(n), _ptr(out)))
    return out


def device_count():
    n = ctypes.c_int(0)
    rc = lib.mirogpu_device_count(ctypes.byref(n))
    return n.value if rc == 0 else 0


# ---- host API layer (libmiro_host.so): the reference's Scene / BVH / Camera interfaces over the C ABI --------
HOST_LIB_PATH = os.path.join(_PKG, "libmiro_host.so")
_host = None


def host_lib():
    global _host
    if _host is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise ImportError(f"{HOST_LIB_PATH} is missing: run __graft_entry__.build()")
        _host = ctypes.CDLL(HOST_LIB_PATH)
        _host.mh_precalc.restype = ctypes.c_double
        _host.mh_render.restype = ctypes.c_double
        _host.mh_scene_handle.restype = ctypes.c_void_p
        _host.mh_trace_photons.restype = ctypes.c_long
    return _host


def _f3(v):
    return (ctypes.c_float * 3)(*[float(x) for x in v])


class HostScene:
    """Drives the C++ host layer the way a user of the reference drives Scene / Camera (one global scene,
    like the reference's g_scene).  Same method names as the checker drivers in tests/miro_driver.py."""

    def __init__(self, layout=LAYOUT_QBVH4, builder=BUILDER_SAH_HOST):
        self.h = host_lib()
        self.layout = layout
        self.builder = builder

    def new_scene(self):
        self.h.mh_new_scene()
        self.h.mh_set_layout(int(self.layout))
        self.h.mh_set_builder(int(self.builder))

    def new_material(self, kd=(1, 1, 1), ks=(0, 0, 0), kt=(0, 0, 0), shininess=1.0, refr=1.0):
        return self.h.mh_new_material(_f3(kd), _f3(ks), _f3(kt), ctypes.c_float(shininess), ctypes.c_float(refr))

    def new_textured_material(self, kind, params, ks=(0, 0, 0), kt=(0, 0, 0), shininess=1.0, refr=1.0):
        """TexturedPhong over a procedural texture: kind = TEX_*, params = the texture's constructor arguments (mirogpu.h)."""
        tp = (ctypes.c_float * 12)(*([float(x) for x in params] + [0.0] * (12 - len(params))))
        return self.h.mh_new_textured_material(int(kind), tp, _f3(ks), _f3(kt), ctypes.c_float(shininess), ctypes.c_float(refr))

    def add_obj(self, path, ctm=None, material=0):
        c = None
        if ctm is not None:
            c = (ctypes.c_float * 16)(*[float(x) for x in np.asarray(ctm, np.float32).reshape(16)])
        n = self.h.mh_add_obj(os.fsencode(path), c, int(material))
        if n < 0:
            raise FileNotFoundError(path)
        return n

    def add_triangle(self, v9, n9, material=0):
        self.h.mh_add_triangle((ctypes.c_float * 9)(*map(float, v9)), (ctypes.c_float * 9)(*map(float, n9)), int(material))

    def add_sphere(self, center, radius, material=0):
        self.h.mh_add_sphere(_f3(center), ctypes.c_float(radius), int(material))

    def add_plane(self, normal, origin, material=0):
        self.h.mh_add_plane(_f3(normal), _f3(origin), int(material))

    def set_device_count(self, n):
        """Replicate the scene on CUDA devices 0..n-1 at the next precalc(): Scene::raytraceImage then shards its rows over them."""
        self.h.mh_set_device_count(int(n))

    def add_point_light(self, pos, color, wattage):
        self.h.mh_add_point_light(_f3(pos), _f3(color), ctypes.c_float(wattage))

    def add_directional_light(self, pos, normal, radius, color, wattage):
        self.h.mh_add_directional_light(_f3(pos), _f3(normal), ctypes.c_float(radius), _f3(color), ctypes.c_float(wattage))

    def set_bg_color(self, c):
        self.h.mh_set_bg_color(_f3(c))

    def set_camera(self, eye, lookat, up, fov):
        self.h.mh_set_camera(_f3(eye), _f3(lookat), _f3(up), ctypes.c_float(fov))

    def camera(self):
        c = Camera()
        self.h.mh_get_camera(ctypes.byref(c))
        return c

    def precalc_host_only(self):
        self.h.mh_precalc_host_only()

    def precalc(self):
        """Scene::preCalc: BVH::build = host SAH build + flatten + upload to the current CUDA device."""
        return self.h.mh_precalc()

    def num_objects(self):
        return self.h.mh_num_objects()

    def dump_triangles(self):
        out = np.zeros((self.num_objects(), 18), np.float32)
        self.h.mh_dump_triangles(_ptr(out))
        return out

    def scene(self):
        """The device scene BVH::build created, as a non-owning MiroScene."""
        s = MiroScene.__new__(MiroScene)
        s._h = ctypes.c_void_p(self.h.mh_scene_handle())
        s.num_triangles = self.num_objects()
        s.close = lambda: None
        return s

    def trace(self, rays, nthreads=0):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        n = rays.shape[0]
        t = np.zeros(n, np.float32); ids = np.zeros(n, np.int32)
        P = np.zeros((n, 3), np.float32); N = np.zeros((n, 3), np.float32)
        self.h.mh_trace(_ptr(rays), ctypes.c_long(n), _ptr(t), _ptr(ids), _ptr(P), _ptr(N), 0)
        return t, ids, P, N

    def eye_rays(self, w, h):
        rays = np.zeros((w * h, 8), np.float32)
        self.h.mh_eye_rays(int(w), int(h), _ptr(rays))
        return rays

    def set_render(self, spp=1, jitter=0, mode=RENDER_WHITTED, shadows=1, seed=168, use_photon_maps=0):
        self.h.mh_set_render(int(spp), int(jitter), int(mode), int(shadows), ctypes.c_uint(seed), int(use_photon_maps))

    def render(self, w, h, out=None):
        """Camera::click -> Scene::raytraceImage; returns the 8-bit image (h, w, 3), row 0 = bottom."""
        if out is None:
            out = np.zeros((h, w, 3), np.uint8)
        assert out.dtype == np.uint8 and out.shape == (h, w, 3) and out.flags.c_contiguous
        self.last_render_seconds = float(self.h.mh_render(int(w), int(h), _ptr(out)))   # Scene::raytraceImage's own timer (Scene.cpp:206)
        return out

    # photon maps
    def set_photon_counts(self, global_photons, caustic_photons):
        """Scene::PhotonsPerLightSource / CausticPhotonsPerLightSource for the next precalc() (this driver starts at 0 / 0)."""
        self.h.mh_set_photon_counts(int(global_photons), int(caustic_photons))

    def trace_photons(self, which):
        """Scene::tracePhotons (which = 0) / traceCausticPhotons (1) on the device; returns the emissions consumed."""
        return int(self.h.mh_trace_photons(int(which)))

    def pm_store(self, which, power, pos, direction):
        power = np.ascontiguousarray(power, np.float32).reshape(-1, 3)
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
        direction = np.ascontiguousarray(direction, np.float32).reshape(-1, 3)
        self.h.mh_pm_store(int(which), _ptr(power), _ptr(pos), _ptr(direction), ctypes.c_long(pos.shape[0]))

    def pm_scale(self, which, s):
        self.h.mh_pm_scale(int(which), ctypes.c_float(s))

    def pm_balance(self, which):
        self.h.mh_pm_balance(int(which))

    def pm_stored(self, which):
        return self.h.mh_pm_stored(int(which))

    def pm_dump(self, which):
        out = np.zeros((self.pm_stored(which) + 1) * 28, np.uint8)
        self.h.mh_pm_dump(int(which), _ptr(out))
        return out.view(PHOTON_DTYPE)

    def pm_attach(self, which):
        self.h.mh_pm_attach(int(which))

    def pm_irradiance(self, which, pos, nrm, max_dist, k, nthreads=0):
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
        nrm = np.ascontiguousarray(nrm, np.float32).reshape(-1, 3)
        irr = np.zeros_like(pos)
        self.h.mh_pm_irradiance(int(which), _ptr(pos), _ptr(nrm), ctypes.c_long(pos.shape[0]), ctypes.c_float(max_dist), int(k), _ptr(irr), 0)
        return irr
