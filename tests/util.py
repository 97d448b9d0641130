"""Shared tolerances. north_star: 'at most 1e-2 relative error on fp16 radiance outputs' (with an absolute floor for
outputs ~ 0), 'loss curves within 2% after N steps', 'bit-exact for record and query indexing'."""
import numpy as np

OUT_REL_TOL = 1e-2      # relative error on fp16 outputs ...
OUT_ABS_FRAC = 2e-3     # ... plus an absolute floor of this fraction of the output scale (max |ref|), for outputs ~ 0
# The reference's CPU `Evaluate` accumulates every dot product in fp16 (Eigen with a _Float16 scalar), which alone puts
# it ~0.6 % of the output scale away from the exactly-rounded network; comparisons AGAINST IT use this wider floor.
REF_EVALUATE_ABS_FRAC = 1e-2
GRAD_REL_TOL = 2e-3     # dW against the oracle's fp32-accumulate mode, relative to the per-layer max |dW|
LOSS_CURVE_TOL = 0.02


def out_err(got, ref, abs_frac: float = OUT_ABS_FRAC, rel: float = OUT_REL_TOL) -> float:
    """max over elements of |got - ref| / (rel * |ref| + abs_frac * max|ref|); the outputs agree iff this is <= 1."""
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape
    if got.size == 0:
        return 0.0
    scale = np.abs(ref).max()
    if scale == 0:
        return float(np.abs(got).max() > 0) * 1e9
    return float((np.abs(got - ref) / (rel * np.abs(ref) + abs_frac * scale)).max())


def layer_rel_err(got, ref):
    errs = []
    for l in range(6):
        a, b = got[l * 4096:(l + 1) * 4096], ref[l * 4096:(l + 1) * 4096]
        if l == 5:
            a, b = got[20480:20672], ref[20480:20672]
        errs.append(float(np.abs(np.asarray(a, np.float64) - b).max() / (np.abs(b).max() + 1e-30)))
    return errs


def he_weights(seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return (rng.standard_normal(20672) * np.sqrt(2.0 / 64.0)).astype(np.float32)


def random_records(seed: int, n: int, pos_scale: float = 4.0) -> np.ndarray:
    """UnpackedNRCInput-shaped records: position U(-s, s)^3, the 11 remaining fields U(0,1)."""
    rng = np.random.default_rng(seed)
    return np.concatenate([rng.uniform(-pos_scale, pos_scale, (n, 3)), rng.uniform(0, 1, (n, 11))], axis=1).astype(np.float32)


def make_scene(seed: int, n_prims: int = 500, n_instances: int = 3, n_materials: int = 12, n_textures: int = 3):
    """A random scene in the reference's buffer layouts (oracle.Scene): non-degenerate triangles inside [-2, 2]^3, rigid
    instance transforms, materials of which about half sample a diffuse and / or specular sRGB texture."""
    import oracle
    rng = np.random.default_rng(seed)
    n_verts = 3 * n_prims
    centres = rng.uniform(-1.5, 1.5, (n_prims, 1, 3))
    # well-conditioned triangles (no slivers: a sliver's fp32 normal is ill-defined in the reference as well)
    frame = np.linalg.qr(rng.standard_normal((n_prims, 3, 3)))[0]
    ang = 2 * np.pi * np.arange(3)[None, :] / 3 + rng.uniform(-0.5, 0.5, (n_prims, 3))
    rad = rng.uniform(0.2, 0.5, (n_prims, 3))
    offs = (rad * np.cos(ang))[..., None] * frame[:, None, :, 0] + (rad * np.sin(ang))[..., None] * frame[:, None, :, 1]
    vertices = (centres + offs).reshape(n_verts, 3).astype(np.float32)
    vertex_indices = np.arange(n_verts, dtype=np.uint32).reshape(n_prims, 3)
    texcoords = rng.uniform(-1.0, 2.0, (n_verts, 2)).astype(np.float32)  # outside [0,1] too: REPEAT addressing
    texcoord_indices = rng.permutation(n_verts).astype(np.uint32).reshape(n_prims, 3)
    mats = np.zeros(n_materials, oracle.MATERIAL_DTYPE)
    mats["diffuse"], mats["specular"], mats["emission"] = rng.uniform(0, 1, (3, n_materials, 3))
    mats["roughness"], mats["metallic"], mats["ior"] = rng.uniform(0, 1, n_materials), rng.uniform(0, 1, n_materials), 1.5
    for key in ("diffuse_texture_id", "specular_texture_id", "emission_texture_id"):
        ids = rng.integers(0, n_textures, n_materials).astype(np.uint32)
        ids[rng.uniform(size=n_materials) < 0.5] = 0xFFFFFFFF
        mats[key] = ids if n_textures else 0xFFFFFFFF
    material_ids = rng.integers(0, n_materials, n_prims).astype(np.uint32)
    transforms = np.zeros((n_instances, 12), np.float32)
    for i in range(n_instances):
        q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
        t = rng.uniform(-1, 1, 3)
        for j in range(3):  # column j of the mat3x4 = {row j of the rotation, translation j}
            transforms[i, 4 * j:4 * j + 3], transforms[i, 4 * j + 3] = q[j], t[j]
    textures = [rng.integers(0, 256, (int(rng.integers(3, 40)), int(rng.integers(3, 40)), 4), dtype=np.uint8) for _ in range(n_textures)]
    return oracle.Scene(vertices, vertex_indices, texcoords, texcoord_indices, mats, material_ids, transforms, textures)


def random_packed_inputs(seed: int, n: int, scene) -> np.ndarray:
    """[n,4] uint32 PackedNRCInput (NRCRecord.glsl:6-10): primitive, flip bit | instance, barycentric y,z, scattered dir."""
    rng = np.random.default_rng(seed)
    prim = rng.integers(0, scene.vertex_indices.shape[0], n).astype(np.uint32)
    inst = rng.integers(0, scene.transforms.shape[0], n).astype(np.uint32) | (rng.integers(0, 2, n).astype(np.uint32) << 31)
    b = rng.dirichlet((1, 1, 1), n)
    bary = (np.round(b[:, 1] * 65535).astype(np.uint32) & 0xFFFF) | (np.round(b[:, 2] * 65535).astype(np.uint32) << 16)
    sd = rng.integers(0, 65536, (n, 2)).astype(np.uint32)
    return np.stack([prim, inst, bary, sd[:, 0] | (sd[:, 1] << 16)], axis=1).astype(np.uint32)
