"""Scene descriptions for the BASELINE configs -- data only (models, transforms, materials, lights, camera).

The constants come from the reference's scene scripts (assignment2.cpp:24-442); a description is then
realised by whichever implementation is under test -- the reference compiled in place, the oracle, or
the product's host layer -- through the same builder calls, so all three see identical inputs.
Transforms are composed in float64 and rounded once to binary32; every implementation receives those
same sixteen floats (row order, Matrix4x4.h:21-24).
"""
import math

import numpy as np


def translate(x, y, z):
    m = np.eye(4)
    m[:3, 3] = [x, y, z]
    return m


def scale(x, y, z):
    return np.diag([x, y, z, 1.0])


def rotate(angle_deg, x, y, z):
    """assignment2.cpp:486-518 -- note the axis is used as given (not normalised), like the reference."""
    rad = angle_deg * (math.pi / 180.0)
    c, s = math.cos(rad), math.sin(rad)
    cinv = 1 - c
    return np.array([[x * x + c * (1 - x * x), x * y * cinv + z * s, x * z * cinv - y * s, 0],
                     [x * y * cinv - z * s, y * y + c * (1 - y * y), y * z * cinv + x * s, 0],
                     [x * z * cinv + y * s, y * z * cinv - x * s, z * z + c * (1 - z * z), 0],
                     [0, 0, 0, 1.0]])


def _chain(*ms):
    out = np.eye(4)
    for m in ms:
        out = out @ m
    return out.astype(np.float32)


FLOOR_BIG = dict(v=[-100, 0, -100, 0, 0, 100, 100, 0, -100], n=[0, 1, 0] * 3)   # assignment2.cpp:101-109
FLOOR_SMALL = dict(v=[-10, 0, -10, 0, 0, 10, 10, 0, -10], n=[0, 1, 0] * 3)      # assignment2.cpp:52-60
LAMBERT_WHITE = dict(kd=(1, 1, 1), ks=(0, 0, 0), kt=(0, 0, 0), shininess=1.0, refr=1.0)  # Lambert(Vector3(1.0f)) == Phong defaults


def _bunny20_transforms():
    x2 = rotate(110, 0, 1, 0) @ scale(.6, 1, 1.1)
    base = [
        [scale(0.3, 2.0, 0.7), translate(-1, .4, .3), rotate(25, .3, .1, .6)],
        [scale(.6, 1.2, .9), translate(7.6, .8, .6)],
        [translate(.7, 0, -2), rotate(120, 0, .6, 1)],
        [translate(3.6, 3, -1)],
        [translate(-2.4, 2, 3), scale(1, .8, 2)],
        [translate(5.5, -.5, 1), scale(1, 2, 1)],
        [rotate(15, 0, 0, 1), translate(-4, -.5, -6), scale(1, 2, 1)],
        [rotate(60, 0, 1, 0), translate(5, .1, 3)],
        [translate(-3, .4, 6), rotate(-30, 0, 1, 0)],
        [translate(3, 0.5, -2), rotate(180, 0, 1, 0), scale(1.5, 1.5, 1.5)],
    ]
    return [_chain(*b) for b in base] + [_chain(x2, *b) for b in base]


SCENES = {
    # config 1 -- makeCornellScene's camera/light on cornell_box.obj (assignment2.cpp:380-405)
    "cornell": dict(
        meshes=[("cornell_box", None, 0)], triangles=[], materials=[LAMBERT_WHITE],
        lights=[dict(kind=0, pos=(2.5, 4.9, -1), color=(1, 1, 1), wattage=160)],
        camera=dict(eye=(2.5, 3, 3), lookat=(2.5, 2.5, 0), up=(0, 1, 0), fov=90), size=(512, 512)),
    # makeTeapotScene (assignment2.cpp:24-70)
    "teapot": dict(
        meshes=[("teapot", None, 0)], triangles=[(FLOOR_SMALL, 0)], materials=[LAMBERT_WHITE],
        lights=[dict(kind=0, pos=(10, 10, 10), color=(1, 1, 1), wattage=700)],
        camera=dict(eye=(0, 3, 6), lookat=(0, 0, 0), up=(0, 1, 0), fov=45), size=(512, 512)),
    # makeBunny1Scene (assignment2.cpp:73-119)
    "bunny1": dict(
        meshes=[("bunny", None, 0)], triangles=[(FLOOR_BIG, 0)], materials=[LAMBERT_WHITE],
        lights=[dict(kind=0, pos=(10, 20, 10), color=(1, 1, 1), wattage=1000)],
        camera=dict(eye=(0, 5, 15), lookat=(0, 0, 0), up=(0, 1, 0), fov=45), size=(512, 512)),
    # config 2 -- bunny + teapot at (-4,0,2) + floor, bunny camera, 1024^2
    "bunny_teapot": dict(
        meshes=[("bunny", None, 0), ("teapot", translate(-4, 0, 2).astype(np.float32), 0)], triangles=[(FLOOR_BIG, 0)],
        materials=[LAMBERT_WHITE], lights=[dict(kind=0, pos=(10, 20, 10), color=(1, 1, 1), wattage=1000)],
        camera=dict(eye=(0, 5, 15), lookat=(0, 0, 0), up=(0, 1, 0), fov=45), size=(1024, 1024)),
    # config 3 stand-in -- makeBunny20Scene (assignment2.cpp:123-338): sponza.obj is absent from the reference tree
    "bunny20": dict(
        meshes=[("bunny", t, 0) for t in _bunny20_transforms()], triangles=[(FLOOR_BIG, 0)], materials=[LAMBERT_WHITE],
        lights=[dict(kind=0, pos=(10, 20, 10), color=(1, 1, 1), wattage=1000)],
        camera=dict(eye=(0, 5, 15), lookat=(0, 0, 0), up=(0, 1, 0), fov=45), size=(1920, 1080)),
    # config 4 -- makeFlowerScene (assignment3.cpp:54-121) with the assets the reference tree holds: Petals2 + Stem + Leaf +
    # WaterDrops.obj (stand-in for the missing WaterDropsMany.obj; FlowerCenter.obj is missing too).  The procedural
    # textures are out of scope: flat colours; background colour instead of the missing HDR map.
    "flower": dict(
        meshes=[("Petals2", None, 0), ("Stem", None, 1), ("Leaf", None, 2), ("WaterDrops", None, 3)], triangles=[],
        materials=[dict(kd=(0.9, 0.35, 0.55), ks=(0, 0, 0), kt=(0, 0, 0), shininess=500.0, refr=1.5),
                   dict(kd=(0.25, 0.6, 0.2), ks=(0, 0, 0), kt=(0, 0, 0), shininess=1.0, refr=1.0),
                   dict(kd=(0.1, 0.5, 0.15), ks=(0, 0, 0), kt=(0, 0, 0), shininess=1.0, refr=1.0),
                   dict(kd=(1, 1, 1), ks=(0, 0, 0), kt=(1, 1, 1), shininess=250.0, refr=1.33)],       # assignment3.cpp:105
        lights=[dict(kind=1, pos=(50, 50, 40), normal=tuple(-np.array([50, 50, 40.0]) / np.linalg.norm([50, 50, 40.0])), radius=7.0,
                     color=(1, 1, 1), wattage=4)],                                                        # assignment3.cpp:78-86
        bg=(1, 1, 1),
        camera=dict(eye=(2, 4.4, 16.8), lookat=(3, 0, 4), up=(0, 1, 0), fov=30), size=(2048, 1365)),
    # config 5 -- the cornell box in four material groups + WaterDrops.obj as caustic caster (assignment2.cpp:414-438),
    # lit by a DirectionalAreaLight so that photons are emitted (Scene.cpp:368).  The reference script's placement
    # (-2,-0.5,0) leaves the drops outside the light's downward beam (z > -0.23 against the disc's z in [-3.5,-1.5]); its
    # caustic pass then never stores a photon and Scene::traceCausticPhotons loops forever (observed with the reference
    # compiled here).  The drops are therefore moved under the light: translate(-1.1, 0.9, -3.8).
    "cornell_drops": dict(
        meshes=[("cornell_box_1", None, 0), ("cornell_box_2", None, 1), ("cornell_box_3", None, 2), ("cornell_box_4", None, 0),
                ("WaterDrops", translate(-1.1, 0.9, -3.8).astype(np.float32), 3)], triangles=[],
        materials=[LAMBERT_WHITE, dict(kd=(1, 0, 0), ks=(0, 0, 0), kt=(0, 0, 0), shininess=1.0, refr=1.0),
                   dict(kd=(0, 1, 0), ks=(0, 0, 0), kt=(0, 0, 0), shininess=1.0, refr=1.0),
                   dict(kd=(1, 1, 1), ks=(0, 0, 0), kt=(1, 1, 1), shininess=5.0, refr=1.5)],            # assignment2.cpp:433
        lights=[dict(kind=1, pos=(2.5, 5.4, -2.5), normal=(0, -1, 0), radius=1.0, color=(1, 1, 1), wattage=160)],
        camera=dict(eye=(2.5, 3, 3), lookat=(2.5, 2.5, 0), up=(0, 1, 0), fov=90), size=(512, 512)),
    # two-triangle smoke geometry (models/testobj.obj)
    "testobj": dict(
        meshes=[("testobj", None, 0)], triangles=[], materials=[LAMBERT_WHITE],
        lights=[dict(kind=0, pos=(0, 5, 5), color=(1, 1, 1), wattage=100)],
        camera=dict(eye=(0, 1, 5), lookat=(0, 0, 0), up=(0, 1, 0), fov=60), size=(64, 64)),
}


def realise(builder, name, obj_path_of):
    """Drives `builder` (an object with new_scene/new_material/add_obj/add_triangle/add_point_light/
    add_directional_light/set_camera/precalc) through scene `name`.  obj_path_of(model) -> .obj path."""
    sc = SCENES[name]
    builder.new_scene()
    for m in sc["materials"]:
        builder.new_material(m["kd"], m["ks"], m["kt"], m["shininess"], m["refr"])
    for model, ctm, mat in sc["meshes"]:
        builder.add_obj(obj_path_of(model), ctm, mat)
    for tri, mat in sc["triangles"]:
        builder.add_triangle(tri["v"], tri["n"], mat)
    for l in sc["lights"]:
        if l["kind"] == 0:
            builder.add_point_light(l["pos"], l["color"], l["wattage"])
        else:
            builder.add_directional_light(l["pos"], l["normal"], l["radius"], l["color"], l["wattage"])
    if "bg" in sc and hasattr(builder, "set_bg_color"):
        builder.set_bg_color(sc["bg"])
    c = sc["camera"]
    builder.set_camera(c["eye"], c["lookat"], c["up"], c["fov"])
    return sc
