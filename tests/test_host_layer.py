"""CPU tier: the host API layer (cse168-raytracer_b200/csrc/miro) and the C-ABI library.

 * libmirogpu.so loads and exports every function include/mirogpu.h declares (no compute call is made:
   there is no GPU in this tier), and compute entry points fail loudly without a device -- there is no CPU path;
 * the host geometry ingest (TriangleMesh::load: transforms, normal synthesis/averaging) and Camera::eyeRay
   reproduce the reference bit for bit (checked against the oracle, itself pinned to the reference in
   test_oracle_vs_reference.py).  Photon_map::balance runs on the device (tests/test_gpu_photon_build.py; its
   algorithm is checked on the CPU in tests/test_photon_build_model.py).
"""
import ctypes
import os
import re

import numpy as np
import pytest

import objio
from conftest import ROOT, bits


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "mirogpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mirogpu_[a-z0-9_]+)\s*\(", src)))


def test_abi_exports_every_declared_symbol(pkg):
    names = _declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(pkg.lib, n), f"libmirogpu.so does not export {n}"
    assert sorted(pkg.EXPORTS) == names
    assert pkg.lib.mirogpu_version() == 1


def test_struct_sizes_match_header(pkg):
    assert ctypes.sizeof(pkg.Material) == 96 and ctypes.sizeof(pkg.Light) == 48 and ctypes.sizeof(pkg.Camera) == 40
    assert ctypes.sizeof(pkg.RenderParams) == 17 * 4 and ctypes.sizeof(pkg.Counters) == 48
    assert pkg.RAY_DTYPE.itemsize == 32 and pkg.HIT_DTYPE.itemsize == 16


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    v = np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0]], np.float32)
    with pytest.raises(pkg.MiroGpuError) as e:
        pkg.MiroScene(v)
    assert "error 2" in str(e.value)          # MIROGPU_ERR_NO_DEVICE
    assert pkg.device_count() == 0


def test_invalid_arguments_are_rejected(pkg):
    h = ctypes.c_void_p()
    assert pkg.lib.mirogpu_scene_create(None, None, None, ctypes.c_uint32(3), None, ctypes.c_uint32(0), None, ctypes.byref(h)) == 1
    assert pkg.lib.mirogpu_scene_create(None, None, None, ctypes.c_uint32(0), None, ctypes.c_uint32(0), None, None) == 1
    assert pkg.lib.mirogpu_intersect_batch(None, None, ctypes.c_size_t(0), None, 0) == 1
    assert b"NULL" in pkg.lib.mirogpu_last_error()


@pytest.mark.parametrize("name", ["testobj", "cornell", "teapot", "bunny_teapot"])
def test_host_ingest_matches_reference_semantics(pkg, oracle, scenes, name):
    H = pkg.HostScene()
    for d in (oracle, H):
        scenes.realise(d, name, objio.obj_path)
    H.precalc_host_only()
    assert np.array_equal(bits(oracle.dump_triangles()), bits(H.dump_triangles()))
    assert np.array_equal(bits(oracle.eye_rays(48, 32)), bits(H.eye_rays(48, 32)))


def test_host_ingest_with_transforms(pkg, oracle, scenes):
    """A rotated / scaled / translated mesh: ctm on vertices, normalised inverse transpose on normals."""
    H = pkg.HostScene()
    ctm = scenes._bunny20_transforms()[10]
    for d in (oracle, H):
        d.new_scene(); d.new_material()
        d.add_obj(objio.obj_path("teapot"), ctm, 0)
        d.add_obj(objio.obj_path("sphere"), ctm, 0)
    H.precalc_host_only()
    assert np.array_equal(bits(oracle.dump_triangles()), bits(H.dump_triangles()))


def test_scene_descriptions_are_complete(scenes):
    assert set(["cornell", "bunny_teapot", "bunny20"]).issubset(scenes.SCENES)
    assert len(scenes.SCENES["bunny20"]["meshes"]) == 20
    assert scenes.SCENES["bunny20"]["size"] == (1920, 1080)
