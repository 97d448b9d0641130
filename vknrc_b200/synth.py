"""Synthetic inputs in the reference's buffer layouts, for tests and bench.py (numpy only, no CUDA): a random scene
(shader/src/Scene.glsl:8-71 layouts) and PackedNRCInput / NRCEvalRecord / NRCTrainRecord arrays (shader/src/NRCRecord.glsl:6-38)
with the statistics SURVEY 8d lists for configs 3 and 4. There is no network access for real scenes; the path tracer that
produces real records is out of scope."""
from __future__ import annotations

import numpy as np

from .api import EVAL_RECORD_DTYPE, MATERIAL_DTYPE, TRAIN_RECORD_DTYPE


def make_scene_arrays(seed: int, n_prims: int = 500, n_instances: int = 3, n_materials: int = 12, n_textures: int = 3) -> dict:
    """Non-degenerate triangles inside about [-2, 2]^3, rigid instance transforms, materials of which about half sample a
    diffuse and / or specular sRGB texture (texture coordinates also outside [0,1]: REPEAT addressing)."""
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
    texcoords = rng.uniform(-1.0, 2.0, (n_verts, 2)).astype(np.float32)
    texcoord_indices = rng.permutation(n_verts).astype(np.uint32).reshape(n_prims, 3)
    mats = np.zeros(n_materials, MATERIAL_DTYPE)
    mats["diffuse"], mats["specular"], mats["emission"] = rng.uniform(0, 1, (3, n_materials, 3))
    mats["roughness"], mats["metallic"], mats["ior"] = rng.uniform(0, 1, n_materials), rng.uniform(0, 1, n_materials), 1.5
    for key in ("diffuse_texture_id", "specular_texture_id", "emission_texture_id"):
        ids = rng.integers(0, max(1, n_textures), n_materials).astype(np.uint32)
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
    return {"vertices": vertices, "vertex_indices": vertex_indices, "texcoords": texcoords, "texcoord_indices": texcoord_indices,
            "materials": mats, "material_ids": material_ids, "transforms": transforms, "textures": textures}


def prim_instance_ids(n_prims: int, n_instances: int) -> np.ndarray:
    """The instance every primitive belongs to: contiguous primitive ranges, as the reference's scene loader produces (one OBJ
    split into instances, each primitive in exactly one; src/Scene.cpp) - the path tracer's records always pair a primitive with
    its own instance."""
    return (np.arange(n_prims, dtype=np.uint64) * n_instances // max(1, n_prims)).astype(np.uint32)


def random_packed_inputs(seed: int, n: int, n_prims: int, n_instances: int, any_instance: bool = False) -> np.ndarray:
    """[n,4] uint32 PackedNRCInput (NRCRecord.glsl:6-10): primitive, flip bit | instance, barycentric y,z, scattered dir.
    The instance is the primitive's own (prim_instance_ids) unless `any_instance` (arbitrary pairs: legal for the shader, never
    produced by the path tracer; exercises the per-record fallback of the normal table)."""
    rng = np.random.default_rng(seed)
    prim = rng.integers(0, n_prims, n).astype(np.uint32)
    inst = rng.integers(0, n_instances, n).astype(np.uint32)
    if not any_instance:
        inst = prim_instance_ids(n_prims, n_instances)[prim]
    inst = inst | (rng.integers(0, 2, n).astype(np.uint32) << 31)
    b = rng.dirichlet((1, 1, 1), n)
    bary = (np.round(b[:, 1] * 65535).astype(np.uint32) & 0xFFFF) | (np.round(b[:, 2] * 65535).astype(np.uint32) << 16)
    sd = rng.integers(0, 65536, (n, 2)).astype(np.uint32)
    return np.stack([prim, inst, bary, sd[:, 0] | (sd[:, 1] << 16)], axis=1).astype(np.uint32)


def eval_records_screen(seed: int, width: int, height: int, n_prims: int, n_instances: int) -> np.ndarray:
    """One NRCEvalRecord per pixel, dst = screen (x, y) (NRCRecord.glsl:19): the inference workload of a frame whose every
    primary path ends in a cache query."""
    n = width * height
    ev = np.zeros(n, EVAL_RECORD_DTYPE)
    idx = np.arange(n, dtype=np.uint32)
    ev["dst"] = ((idx % width) | ((idx // width) << 15)) << 1
    ev["packed_input"] = np.ascontiguousarray(random_packed_inputs(seed, n, n_prims, n_instances)).view(EVAL_RECORD_DTYPE["packed_input"]).reshape(n)
    return ev


def train_records(seed: int, n: int, n_prims: int, n_instances: int) -> np.ndarray:
    """[n] NRCTrainRecord: bias (= target radiance) and factor U(0,1)^3 + a PackedNRCInput."""
    rng = np.random.default_rng(seed)
    r = np.zeros(n, TRAIN_RECORD_DTYPE)
    r["bias"], r["factor"] = rng.uniform(0, 1, (n, 3)), rng.uniform(0, 1, (n, 3))
    r["packed_input"] = np.ascontiguousarray(random_packed_inputs(seed + 1, n, n_prims, n_instances)).view(TRAIN_RECORD_DTYPE["packed_input"]).reshape(n)
    return r


MAX_BOUNCE = 8  # shader/src/path_tracer.comp:11
CONST_LIGHT = 10.0  # kConstLight (path_tracer.comp:142): radiance of a ray that leaves the scene


def frame_records(seed: int, width: int, height: int, n_prims: int, n_instances: int, train_probability: float = 0.03,
                  batch_size: int = 16384, batch_count: int = 4, light_terminate_probability: float = 0.3) -> dict:
    """One frame's record buffers with the STRUCTURE the reference's path tracer produces (path_tracer.comp:254-375, SURVEY
    appendix B), from random hits instead of traced ones (the tracer itself is out of scope):
      * every pixel appends one screen-destined NRCEvalRecord;
      * with probability `train_probability` a whole subgroup (32 consecutive pixels of a row, path_tracer.comp:389-394) runs
        extended "train" paths: `bounce` in [1, MAX_BOUNCE] vertices, per-vertex throughput colour and emitted light; the
        suffix scan of :347-350 turns them into bias_i = radiance collected from vertex i on and factor_i = throughput from
        vertex i to the path's tail; the `bounce` records go CONTIGUOUSLY into a random batch (atomic append, truncated at
        the batch capacity while the count keeps growing, :343-351); unless the path left the scene, one more eval record
        with dst = train(batch, first, last) asks the cache for the tail radiance, which nrc_inference.comp:60-72 adds to
        every bias_i, i in [first, last], scaled by factor_i.
    Returns eval_records, eval_count, train_records[batch_count] (capacity `batch_size` each) and the raw (unclamped) counts."""
    rng = np.random.default_rng(seed)
    n_pix = width * height
    idx = np.arange(n_pix, dtype=np.uint32)
    screen = np.zeros(n_pix, EVAL_RECORD_DTYPE)
    screen["dst"] = ((idx % width) | ((idx // width) << 15)) << 1
    screen["packed_input"] = np.ascontiguousarray(random_packed_inputs(seed + 1, n_pix, n_prims, n_instances)).view(EVAL_RECORD_DTYPE["packed_input"]).reshape(n_pix)
    # which pixels run train paths: per subgroup of 32 consecutive pixels in a row
    groups_per_row = (width + 31) // 32
    train_group = rng.uniform(size=(height, groups_per_row)) < train_probability
    train_pix = np.flatnonzero(np.repeat(train_group, 32, axis=1)[:, :width].reshape(-1))
    trains = [np.zeros(batch_size, TRAIN_RECORD_DTYPE) for _ in range(batch_count)]
    counts = np.zeros(batch_count, np.int64)
    tails = []
    n_paths = len(train_pix)
    bounces = rng.integers(1, MAX_BOUNCE + 1, n_paths)
    batches = np.minimum((rng.uniform(size=n_paths) * batch_count).astype(np.int64), batch_count - 1)
    leaves = rng.uniform(size=n_paths) < light_terminate_probability
    vertex_inputs = random_packed_inputs(seed + 2, int(bounces.sum()), n_prims, n_instances)
    v0 = 0
    for path in range(n_paths):
        b, batch = int(bounces[path]), int(batches[path])
        colors = rng.uniform(0.05, 0.95, (b, 3)).astype(np.float32)
        lights = np.where(rng.uniform(size=(b, 1)) < 0.05, rng.uniform(0, 5, (b, 3)), 0.0).astype(np.float32)  # few emitters
        if leaves[path]:
            lights[b - 1] = CONST_LIGHT
        for i in range(b - 2, -1, -1):  # path_tracer.comp:347-350 (fp32, same order)
            lights[i] = lights[i] + colors[i] * lights[i + 1]
            colors[i] = colors[i] * colors[i + 1]
        first = int(counts[batch])
        counts[batch] += b  # the atomic keeps counting past the capacity (:343-345)
        if first < batch_size:
            cnt = min(b, batch_size - first)
            rec = trains[batch][first:first + cnt]
            rec["bias"], rec["factor"] = lights[:cnt], colors[:cnt]
            rec["packed_input"] = np.ascontiguousarray(vertex_inputs[v0:v0 + cnt]).view(TRAIN_RECORD_DTYPE["packed_input"]).reshape(cnt)
            if not leaves[path]:
                tail = np.zeros(1, EVAL_RECORD_DTYPE)
                tail["dst"] = (((batch | (first << 2) | ((first + cnt - 1) << 16)) << 1) | 1) & 0xFFFFFFFF
                tail["packed_input"] = np.ascontiguousarray(vertex_inputs[v0 + b - 1:v0 + b]).view(EVAL_RECORD_DTYPE["packed_input"]).reshape(1)
                tails.append(tail)
        v0 += b
    ev = np.concatenate([screen] + tails) if tails else screen
    return {"eval_records": ev, "eval_count": len(ev), "train_records": trains, "train_counts": counts, "train_paths": n_paths}
