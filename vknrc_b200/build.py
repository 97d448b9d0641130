"""Builds libnrc_b200.so (the CUDA kernels + the C ABI of include/nrc_b200.h) in-tree with nvcc for sm_100a.

In-tree so that the built library travels to the GPU box with the repo snapshot; no JIT cache is involved.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libnrc_b200.so")
SOURCES = ["nrc_infer.cu", "nrc_train.cu", "nrc_state.cu"]
HEADERS = ["sm100_ptx.cuh", "nrc_config.h", "nrc_kernels.h", "nrc_encode.cuh", "nrc_unpack.cuh", "nrc_state.hpp", "../../include/nrc_b200.h",
           "../../include/nrc_b200_types.h"]
STAMP = os.path.join(HERE, "build", "build_stamp.json")  # what the last build ran: written by build(), read by tests / the driver
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
              "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def source_hash() -> str:
    import hashlib
    h = hashlib.sha256()
    for f in sorted(SOURCES + HEADERS):
        h.update(open(os.path.join(CSRC, f), "rb").read())
    return h.hexdigest()


def needs_build() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    try:
        import json
        if json.load(open(STAMP)).get("source_sha256") != source_hash():
            return True  # (content, not mtime: a checkout or a snapshot copy rewrites every mtime)
    except Exception:
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False, extra_flags=(), out: str = LIB) -> str:
    """`out` / `extra_flags`: development variants (tools/lab_train.py) built beside the product library."""
    if out == LIB and not force and not needs_build():
        return LIB
    objdir = os.path.join(HERE, "build") if out == LIB else out + ".obj"
    os.makedirs(objdir, exist_ok=True)
    env = dict(os.environ)
    env.pop("CC", None), env.pop("CXX", None)  # this image exports a gcc wrapper that nvcc must not pick up
    objs, procs = [], []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [_nvcc(), *NVCC_FLAGS, *extra_flags, "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)))
        objs.append(obj)
    log = []
    for src, p in procs:
        text, _ = p.communicate()
        log.append(text)
        if p.returncode != 0:
            sys.stderr.write(text)
            raise RuntimeError(f"nvcc failed on {src}")
    link = [_nvcc(), "-shared", "-o", out, *objs, "-lcudart"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    with open(os.path.join(objdir, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if out == LIB:
        import hashlib
        import json
        import time
        ver = subprocess.run([_nvcc(), "--version"], stdout=subprocess.PIPE, text=True, env=env).stdout.strip().splitlines()[-1]
        h = hashlib.sha256()
        for f in sorted(SOURCES + HEADERS):
            h.update(open(os.path.join(CSRC, f), "rb").read())
        with open(STAMP, "w") as f:
            json.dump({"built_at": time.time(), "nvcc": ver, "flags": NVCC_FLAGS, "sources": SOURCES, "source_sha256": h.hexdigest(),
                       "kernels_ptxas_lines": sum(l.count("Used ") for l in log)}, f)
    if verbose:
        print("\n".join(log))
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
