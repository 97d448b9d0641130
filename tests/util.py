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
    """The synthetic scene of vknrc_b200.synth as an oracle.Scene (host buffers in the reference's layouts)."""
    import oracle
    from vknrc_b200.synth import make_scene_arrays
    a = make_scene_arrays(seed, n_prims, n_instances, n_materials, n_textures)
    return oracle.Scene(a["vertices"], a["vertex_indices"], a["texcoords"], a["texcoord_indices"], a["materials"], a["material_ids"],
                        a["transforms"], a["textures"])


def random_packed_inputs(seed: int, n: int, scene) -> np.ndarray:
    from vknrc_b200.synth import random_packed_inputs as gen
    return gen(seed, n, scene.vertex_indices.shape[0], scene.transforms.shape[0])
