"""CPU checks of the restated UnpackNRCInput (oracle/nrc_oracle.c) against an independent float64 numpy evaluation of
shader/src/NRCRecord.glsl:98-125 + shader/src/Scene.glsl:50-64 on a random scene."""
import numpy as np

import oracle
from util import make_scene, random_packed_inputs


def _srgb(c):
    x = c / 255.0
    return np.where(x <= 0.04045, x / 12.92, ((x + 0.055) / 1.055) ** 2.4)


def _sample(tex, u, v):
    h, w = tex.shape[:2]
    x, y = u * w - 0.5, v * h - 0.5
    fx, fy = np.floor(x), np.floor(y)
    tx, ty = x - fx, y - fy
    x0, y0 = int(fx) % w, int(fy) % h
    x1, y1 = (x0 + 1) % w, (y0 + 1) % h
    t = _srgb(tex[..., :3].astype(np.float64))
    return (t[y0, x0] * (1 - tx) + t[y0, x1] * tx) * (1 - ty) + (t[y1, x0] * (1 - tx) + t[y1, x1] * tx) * ty


def test_unpack_against_float64_numpy():
    sc = make_scene(3, n_prims=200)
    pk = random_packed_inputs(4, 400, sc)
    got = oracle.unpack(sc, pk)
    for i in range(pk.shape[0]):
        prim, inst, flip = int(pk[i, 0]), int(pk[i, 1] & 0x7FFFFFFF), bool(pk[i, 1] >> 31)
        M = sc.transforms[inst].reshape(3, 4).astype(np.float64)
        v = [M[:, :3] @ sc.vertices[sc.vertex_indices[prim, k]].astype(np.float64) + M[:, 3] for k in range(3)]
        nrm = np.cross(v[1] - v[0], v[2] - v[0])
        nrm = nrm / np.linalg.norm(nrm) * (-1.0 if flip else 1.0)
        by, bz = (int(pk[i, 2]) & 0xFFFF) / 65535.0, (int(pk[i, 2]) >> 16) / 65535.0
        bx = 1.0 - by - bz
        pos = v[0] * bx + v[1] * by + v[2] * bz
        tc = [sc.texcoords[sc.texcoord_indices[prim, k]].astype(np.float64) for k in range(3)]
        uv = tc[0] * bx + tc[1] * by + tc[2] * bz
        mat = sc.materials[sc.material_ids[prim]]
        diff = mat["diffuse"] if mat["diffuse_texture_id"] == 0xFFFFFFFF else _sample(sc.textures[mat["diffuse_texture_id"]], uv[0], uv[1])
        spec = mat["specular"] if mat["specular_texture_id"] == 0xFFFFFFFF else _sample(sc.textures[mat["specular_texture_id"]], uv[0], uv[1])
        exp = np.concatenate([pos, [(int(pk[i, 3]) & 0xFFFF) / 65535.0, (int(pk[i, 3]) >> 16) / 65535.0],
                              [0.5 + np.arctan2(nrm[1], nrm[0]) / (2 * np.pi), np.arccos(np.clip(nrm[2], -1, 1)) / np.pi],
                              [mat["roughness"]], diff, spec])
        tol = np.full(14, 5e-6)
        tol[8:] = 2e-4  # a texel-boundary sample can pick the neighbouring footprint in fp32; colours are in [0,1]
        assert np.all(np.abs(got[i] - exp) <= tol), (i, got[i], exp)
