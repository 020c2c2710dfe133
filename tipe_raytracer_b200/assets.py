"""OBJ/MTL + P3-PPM loading in Python, with the reference's conventions (list_of_mesh, mesh.h:96-218;
tex_path_from_mtl, rtutility.h:233-290; create_mat_list_mtl, texture.h:175-354).  Same outputs as the C
loaders of apps/trt_assets.h; used by tests and tools to build `Scene`s from files."""
from __future__ import annotations

import os
import re

import numpy as np

from .scene import MATERIAL, SKY, TRIANGLE


def _texture_of(mtl_text: str, mtl_path: str, name: str):
    lines = mtl_text.splitlines()
    for k, line in enumerate(lines):
        if line.startswith("newmtl ") and line[7:].strip() == name:
            for later in lines[k + 1:]:
                if later.startswith("map_Kd "):
                    rel = later[7:].strip()
                    if rel.startswith("./"):
                        rel = rel[2:]
                    return os.path.join(os.path.dirname(mtl_path), rel) if os.path.dirname(mtl_path) else rel
            return None
    return None


def load_mesh(obj_path: str, mtl_path: str):
    """-> (TRIANGLE[n], quel_mat int32[n], [texture path per usemtl])"""
    mtl_text = open(mtl_path, newline="").read()
    v, vt, faces, quel, paths = [], [], [], [], []
    mat = -1
    for line in open(obj_path, newline=""):
        if line.startswith("v "):
            v.append([float(x) for x in line.split()[1:4]])
        elif line.startswith("vt"):
            vt.append([float(x) for x in line.split()[1:3]])
        elif line.startswith("f"):
            m = re.match(r"f (\d+)/(\d+)/\d+ (\d+)/(\d+)/\d+ (\d+)/(\d+)/\d+", line)
            if not m:
                raise ValueError(f"{obj_path}: face is not a v/vt/vn triangle: {line!r}")
            faces.append([int(x) for x in m.groups()])
            quel.append(mat)
        elif line.startswith("usemtl "):
            mat += 1
            p = _texture_of(mtl_text, mtl_path, line[7:].strip())
            if p is None:
                raise ValueError(f"{mtl_path}: no map_Kd for material {line[7:].strip()!r}")
            paths.append(p)
    v, vt, f = np.array(v), np.array(vt), np.array(faces)
    tris = np.zeros(len(faces), TRIANGLE)
    for c, name in enumerate(("A", "B", "C")):
        tris[name] = v[f[:, 2 * c] - 1]
        tris["uv" + name] = vt[f[:, 2 * c + 1] - 1]
    tris["mat"]["diffuseColor"] = SKY
    return tris, np.array(quel, np.int32), paths


def _read_p3(path: str):
    tok = re.sub(r"#[^\n]*", " ", open(path).read()).split()
    if tok[0] != "P3":
        raise ValueError(f"{path} is not a P3 PPM")
    w, h, mx = int(tok[1]), int(tok[2]), int(tok[3])
    data = np.array(tok[4:4 + w * h * 3], np.float64).reshape(h, w, 3)
    return data, mx


def load_textures(paths):
    """-> (MATERIAL[nb*th*tw], tw, th): diffuse = rgb/max, alpha = first channel of <stem>_alpha.ppm / max,
    rows flipped so that row 0 is the bottom of the image."""
    slabs, size = [], None
    for p in paths:
        stem = p[:p.index(".png")]
        col, mx = _read_p3(stem + ".ppm")
        alp, amx = _read_p3(stem + "_alpha.ppm")
        if size is None:
            size = col.shape[:2]
        if col.shape[:2] != size or alp.shape[:2] != size:
            raise ValueError("all textures and alpha maps must share one size")
        m = np.zeros(size, MATERIAL)
        m["diffuseColor"] = col[::-1] / mx
        m["alpha"] = alp[::-1, :, 0] / amx
        slabs.append(m.reshape(-1))
    th, tw = size
    return np.concatenate(slabs), tw, th
