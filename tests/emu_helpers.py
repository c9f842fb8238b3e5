"""ctypes helpers for the test-only host emulation library (tests/cpu_emu/libmiro_emu.so)."""
import ctypes

import numpy as np

HIT_DTYPE = np.dtype([("t", np.float32), ("prim_id", np.uint32), ("beta", np.float32), ("gamma", np.float32)])
BVH2_NODE = np.dtype([("f", np.float32, 12), ("link", np.int32, 4)])
CW_NODE = np.dtype([("p", np.float32, 3), ("e", np.uint8, 3), ("imask", np.uint8), ("child_base", np.uint32), ("tri_base", np.uint32),
                    ("meta", np.uint8, 8), ("qlox", np.uint8, 8), ("qloy", np.uint8, 8), ("qloz", np.uint8, 8),
                    ("qhix", np.uint8, 8), ("qhiy", np.uint8, 8), ("qhiz", np.uint8, 8)])
BVH4_NODE = np.dtype([("lox", np.float32, 4), ("hix", np.float32, 4), ("loy", np.float32, 4), ("hiy", np.float32, 4), ("loz", np.float32, 4), ("hiz", np.float32, 4),
                      ("link", np.int32, 4), ("pad", np.int32, 4)])
QBVH4_NODE = np.dtype([("origin", np.float32, 3), ("e", np.uint8, 3), ("pad0", np.uint8), ("qlox", np.uint8, 4), ("qhix", np.uint8, 4), ("qloy", np.uint8, 4),
                       ("qhiy", np.uint8, 4), ("qloz", np.uint8, 4), ("qhiz", np.uint8, 4), ("link", np.int32, 4), ("cell", np.uint32, 2)])
TRI_REC = np.dtype([("a", np.float32, 3), ("prim_id", np.uint32), ("e1", np.float32, 3), ("nx", np.float32), ("e2", np.float32, 3), ("ny", np.float32),
                    ("nz", np.float32), ("pad", np.float32, 3)])
assert BVH2_NODE.itemsize == 64 and CW_NODE.itemsize == 80 and TRI_REC.itemsize == 64 and BVH4_NODE.itemsize == 128 and QBVH4_NODE.itemsize == 64


def emu_trace(emu, verts, rays, layout, any_hit=False, max_leaf=0):
    verts = np.ascontiguousarray(verts, np.float32).reshape(-1, 9)
    rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
    hits = np.zeros(rays.shape[0], HIT_DTYPE)
    cnt = (ctypes.c_ulonglong * 3)()
    info = (ctypes.c_uint32 * 4)()
    emu.emu_trace(verts.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(verts.shape[0]), int(layout), int(max_leaf),
                  rays.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(rays.shape[0]), hits.ctypes.data_as(ctypes.c_void_p),
                  int(any_hit), cnt, info)
    return hits, dict(nodes=cnt[0], boxes=cnt[1], tris=cnt[2]), list(info)


def emu_build(emu, verts, layout, max_leaf=0):
    verts = np.ascontiguousarray(verts, np.float32).reshape(-1, 9)
    sizes = (ctypes.c_uint64 * 8)()
    vp = verts.ctypes.data_as(ctypes.c_void_p)
    emu.emu_build(vp, ctypes.c_uint32(verts.shape[0]), int(layout), int(max_leaf), None, None, None, sizes)
    nodes = np.zeros(sizes[0], np.uint8)
    order = np.zeros(sizes[1], np.uint32)
    tris = np.zeros(sizes[1], TRI_REC)
    emu.emu_build(vp, ctypes.c_uint32(verts.shape[0]), int(layout), int(max_leaf), nodes.ctypes.data_as(ctypes.c_void_p),
                  order.ctypes.data_as(ctypes.c_void_p), tris.ctypes.data_as(ctypes.c_void_p), sizes)
    return nodes.view({0: BVH2_NODE, 1: CW_NODE, 3: BVH4_NODE, 4: QBVH4_NODE}[layout]), order, tris, dict(binary_nodes=sizes[2], binary_leaves=sizes[3], depth=sizes[4])


def hit_ids(hits):
    ids = hits["prim_id"].astype(np.int64)
    ids[ids == 0xFFFFFFFF] = -1
    return ids
