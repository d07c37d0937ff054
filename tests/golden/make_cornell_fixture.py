"""Builds tests/golden/cornell_box.npz from the reference's Cornell box test scene.

Run in the build container only (it reads /root/reference):
    python tests/golden/make_cornell_fixture.py

Source: sandbox/tests/test scenes/sppm/01 - cornell box.appleseed + its 8 OBJ meshes
(32 triangles; each object instance scaled by diag(0.001...); pinhole camera, film 0.025^2,
focal length 0.035, 512x512) -- SURVEY.md section 8(d), config C1.
"""
import os
import re
import sys
import xml.etree.ElementTree as ET

import numpy as np

REF = "/root/reference/sandbox/tests/test scenes/sppm"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cornell_box.npz")


def read_obj(path):
    verts, tris = [], []
    for line in open(path):
        p = line.split()
        if not p:
            continue
        if p[0] == "v":
            verts.append([float(x) for x in p[1:4]])
        elif p[0] == "f":
            idx = [int(tok.split("/")[0]) - 1 for tok in p[1:]]
            for k in range(1, len(idx) - 1):        # fan triangulation, as the OBJ reader does
                tris.append([idx[0], idx[k], idx[k + 1]])
    return np.array(verts, dtype=np.float32), np.array(tris, dtype=np.uint32)


def matrix(elem):
    return np.array([float(x) for x in elem.find("transform").find("matrix").text.split()]).reshape(4, 4)


def main():
    root = ET.parse(os.path.join(REF, "01 - cornell box.appleseed")).getroot()
    scene = root.find("scene")
    cam = scene.find("camera")
    asm = scene.find("assembly")
    objects = {o.get("name"): o.find("parameter[@name='filename']").get("value") for o in asm.findall("object")}
    out = {"camera_matrix": matrix(cam)}
    for p in cam.findall("parameter"):
        if p.get("name") == "film_dimensions":
            out["film_dimensions"] = np.array([float(x) for x in p.get("value").split()])
        if p.get("name") == "focal_length":
            out["focal_length"] = np.array(float(p.get("value")))
    res = root.find("output").find("frame").find("parameter[@name='resolution']").get("value")
    out["resolution"] = np.array([int(x) for x in res.split()])
    names = []
    for i, inst in enumerate(asm.findall("object_instance")):
        obj = inst.get("object").split(".")[0]
        v, t = read_obj(os.path.join(REF, objects[obj]))
        out["vertices_%d" % i] = v
        out["triangles_%d" % i] = t
        out["transform_%d" % i] = matrix(inst)
        names.append(obj)
    out["names"] = np.array(names)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, "objects:", names, "triangles:", sum(out["triangles_%d" % i].shape[0] for i in range(len(names))))


if __name__ == "__main__":
    sys.exit(main())
