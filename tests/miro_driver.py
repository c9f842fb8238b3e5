"""ctypes front end shared by the checkers.

The reference compiled in place (oracle/_ref/libmiro_ref*.so, prefix ``ref_``) and the CPU
restatement (oracle/libmiro_oracle.so, prefix ``orc_``) export the same driver functions, so one
wrapper serves both.  TEST INFRASTRUCTURE: imported only from tests/, bench.py's cpu_baseline /
--impl reference legs and __graft_entry__.smoke().
"""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "libmiro_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libmiro_ref.so")
REF_STATS_SO = os.path.join(ROOT, "oracle", "_ref", "libmiro_ref_stats.so")
REF_SSE_SO = os.path.join(ROOT, "oracle", "_ref", "libmiro_ref_sse.so")

MIRO_TMAX = np.float32(1e12)

_c_float_p = ctypes.POINTER(ctypes.c_float)


def _fp(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _f3(v):
    return (ctypes.c_float * 3)(*[float(x) for x in v])


class Driver:
    """One scene in one checker library (the libraries keep a single global scene each)."""

    def __init__(self, so_path, prefix):
        self.lib = ctypes.CDLL(so_path)
        self.p = prefix
        f = self._f
        f("precalc").restype = ctypes.c_double
        f("trace_time").restype = ctypes.c_double
        if prefix == "ref_":
            f("render").restype = ctypes.c_double
        if prefix == "orc_":
            f("pm_visited").restype = ctypes.c_longlong

    def _f(self, name):
        return getattr(self.lib, self.p + name)

    # ---- scene construction -------------------------------------------------------------------
    def new_scene(self):
        self._f("new_scene")()

    def new_material(self, kd=(1, 1, 1), ks=(0, 0, 0), kt=(0, 0, 0), shininess=1.0, refr=1.0):
        """Phong(kd, ks, kt, shininess, refractIndex); shininess < 0 means infinity."""
        return self._f("new_material")(_f3(kd), _f3(ks), _f3(kt), ctypes.c_float(shininess), ctypes.c_float(refr))

    def new_textured_material(self, kind, params, ks=(0, 0, 0), kt=(0, 0, 0), shininess=1.0, refr=1.0):
        """reference only: TexturedPhong over one of its procedural textures (kind numbers of include/mirogpu.h)."""
        tp = (ctypes.c_float * 12)(*([float(x) for x in params] + [0.0] * (12 - len(params))))
        return self._f("new_textured_material")(int(kind), tp, _f3(ks), _f3(kt), ctypes.c_float(shininess), ctypes.c_float(refr))

    def texture_lookup(self, kind, params, coords, bump=False):
        """reference only: the real Texture classes' lookup2D / lookup3D (and bumpHeight2D) at coords (n, 3)."""
        coords = np.ascontiguousarray(coords, np.float32).reshape(-1, 3)
        tp = (ctypes.c_float * 12)(*([float(x) for x in params] + [0.0] * (12 - len(params))))
        rgb = np.zeros((coords.shape[0], 3), np.float32); b = np.zeros(coords.shape[0], np.float32)
        self._f("texture_lookup")(int(kind), tp, _fp(coords), ctypes.c_long(coords.shape[0]), _fp(rgb), _fp(b) if bump else None)
        return (rgb, b) if bump else rgb

    def add_obj(self, path, ctm=None, material=0):
        c = None
        if ctm is not None:
            c = (ctypes.c_float * 16)(*[float(x) for x in np.asarray(ctm, np.float32).reshape(16)])
        n = self._f("add_obj")(os.fsencode(path), c, int(material))
        if n < 0:
            raise FileNotFoundError(path)
        return n

    def add_triangle(self, v9, n9, material=0):
        self._f("add_triangle")((ctypes.c_float * 9)(*map(float, v9)), (ctypes.c_float * 9)(*map(float, n9)), int(material))

    def add_sphere(self, center, radius, material=0):
        self._f("add_sphere")(_f3(center), ctypes.c_float(radius), int(material))

    def add_plane(self, normal, origin, material=0):
        self._f("add_plane")(_f3(normal), _f3(origin), int(material))

    def add_point_light(self, pos, color, wattage):
        self._f("add_point_light")(_f3(pos), _f3(color), ctypes.c_float(wattage))

    def add_directional_light(self, pos, normal, radius, color, wattage):
        self._f("add_directional_light")(_f3(pos), _f3(normal), ctypes.c_float(radius), _f3(color), ctypes.c_float(wattage))

    def set_bg_color(self, c):
        self._f("set_bg_color")(_f3(c))

    def precalc(self):
        return self._f("precalc")()

    def num_objects(self):
        return self._f("num_objects")()

    def dump_triangles(self):
        out = np.zeros((self.num_objects(), 18), np.float32)
        self._f("dump_triangles")(_fp(out))
        return out

    # ---- queries ------------------------------------------------------------------------------
    def _trace(self, fn, rays, nthreads):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        n = rays.shape[0]
        t = np.zeros(n, np.float32)
        ids = np.zeros(n, np.int32)
        P = np.zeros((n, 3), np.float32)
        N = np.zeros((n, 3), np.float32)
        self._f(fn)(_fp(rays), ctypes.c_long(n), _fp(t), _fp(ids), _fp(P), _fp(N), int(nthreads))
        return t, ids, P, N

    def trace(self, rays, nthreads=0):
        """Scene::trace per ray -> (t, prim_id (-1 = miss), P, N normalised)."""
        return self._trace("trace", rays, nthreads)

    def trace_brute(self, rays, nthreads=0):
        return self._trace("trace_brute", rays, nthreads)

    def trace_time(self, rays, nthreads=0):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        hits = ctypes.c_long(0)
        s = self._f("trace_time")(_fp(rays), ctypes.c_long(rays.shape[0]), int(nthreads), ctypes.byref(hits))
        return s, hits.value

    def trace_time_hits(self, rays, nthreads=0):
        """The timing leg that keeps its answers: (seconds, t, prim_id (-1 = miss))."""
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        n = rays.shape[0]
        t = np.zeros(n, np.float32)
        ids = np.zeros(n, np.int32)
        f = self._f("trace_time_hits")
        f.restype = ctypes.c_double
        s = f(_fp(rays), ctypes.c_long(n), int(nthreads), _fp(t), _fp(ids))
        return s, t, ids

    def host_threads(self):
        return int(self._f("host_threads")())

    def stats(self):
        out = (ctypes.c_longlong * 9)()
        self._f("stats_get")(out)
        k = ["nodes", "leaves", "rays", "primary", "secondary", "shadow", "photon_bounces", "ray_box", "ray_tri"]
        return dict(zip(k, list(out)))

    def stats_reset_rays(self):
        self._f("stats_reset_rays")()

    def set_camera(self, eye, lookat, up, fov):
        self._f("set_camera")(_f3(eye), _f3(lookat), _f3(up), ctypes.c_float(fov))

    def eye_rays(self, w, h):
        rays = np.zeros((w * h, 8), np.float32)
        self._f("eye_rays")(int(w), int(h), _fp(rays))
        return rays

    def trace_scene(self, rays, depth=10, nthreads=0):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        rgb = np.zeros((rays.shape[0], 3), np.float32)
        if self.p == "orc_":
            self._f("trace_scene")(_fp(rays), ctypes.c_long(rays.shape[0]), int(depth), _fp(rgb), int(nthreads), None)
        else:
            self._f("trace_scene")(_fp(rays), ctypes.c_long(rays.shape[0]), int(depth), _fp(rgb), int(nthreads))
        return rgb

    # ---- photon maps --------------------------------------------------------------------------
    def trace_photons(self, light, caustic, seed, first, count, nthreads=0):
        """oracle only: Scene::tracePhoton over emissions [first, first+count) with the counter-based uniforms."""
        counts = np.zeros(count, np.uint8)
        records = np.zeros((count, 5, 9), np.float32)
        self._f("trace_photons")(int(light), int(caustic), ctypes.c_uint(seed), ctypes.c_ulonglong(first), ctypes.c_uint(count),
                                 counts.ctypes.data_as(ctypes.c_void_p), records.ctypes.data_as(ctypes.c_void_p), int(nthreads))
        return counts, records

    def pm_new(self, max_photons):
        return self._f("pm_new")(int(max_photons))

    def pm_store(self, which, power, pos, direction):
        power = np.ascontiguousarray(power, np.float32).reshape(-1, 3)
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
        direction = np.ascontiguousarray(direction, np.float32).reshape(-1, 3)
        self._f("pm_store")(int(which), _fp(power), _fp(pos), _fp(direction), ctypes.c_long(pos.shape[0]))

    def pm_scale(self, which, s):
        self._f("pm_scale")(int(which), ctypes.c_float(s))

    def pm_balance(self, which):
        self._f("pm_balance")(int(which))

    def pm_stored(self, which):
        return self._f("pm_stored")(int(which))

    def pm_dump(self, which):
        n = self.pm_stored(which)
        out = np.zeros((n + 1) * 28, np.uint8)
        self._f("pm_dump")(int(which), _fp(out))
        return out.view(PHOTON_DTYPE)

    def pm_irradiance(self, which, pos, nrm, max_dist, k, nthreads=0):
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
        nrm = np.ascontiguousarray(nrm, np.float32).reshape(-1, 3)
        irr = np.zeros_like(pos)
        self._f("pm_irradiance")(int(which), _fp(pos), _fp(nrm), ctypes.c_long(pos.shape[0]), ctypes.c_float(max_dist), int(k), _fp(irr), int(nthreads))
        return irr


# Photon, PhotonMap.h:16-22 (28 bytes)
PHOTON_DTYPE = np.dtype([("pos", np.float32, 3), ("plane", np.int16), ("theta", np.uint8), ("phi", np.uint8), ("power", np.float32, 3)])
assert PHOTON_DTYPE.itemsize == 28


def oracle():
    return Driver(ORACLE_SO, "orc_")


def reference(kind="scalar"):
    return Driver({"scalar": REF_SO, "stats": REF_STATS_SO, "sse": REF_SSE_SO}[kind], "ref_")
