"""CPU tests of the drop-in boundary: the shared library loads, exports every symbol include/nrc_b200.h declares,
its pure size helpers agree with the reference, and it fails loudly (no fallback) without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    import vknrc_b200
    return vknrc_b200.lib()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "nrc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nrc_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(L):
    from vknrc_b200.api import SIGNATURES
    declared = _declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/nrc_b200.h but not exported"
        assert name in SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(SIGNATURES) == declared


def test_size_helpers_match_reference(L):
    # src/VkNRCState.cpp:34-37 with sizeof(NRCEvalRecord) = 20, sizeof(NRCTrainRecord) = 40; src/VkNRCState.hpp:23-25
    assert L.nrc_get_eval_record_buffer_size(1920, 1080) == (1920 * 1080 + 16384 * 4) * 20
    assert L.nrc_get_eval_record_buffer_size(0, 0) == 65536 * 20
    assert L.nrc_get_batch_train_record_buffer_size() == 16384 * 40
    assert L.nrc_get_train_batch_count() == 4
    assert L.nrc_get_train_batch_size() == 16384
    assert L.nrc_get_weight_count() == 5 * 64 * 64 + 3 * 64 == 20672
    assert abs(L.nrc_get_default_train_probability() - 0.03) < 1e-9


def test_record_layouts():
    from vknrc_b200 import api
    assert api.EVAL_RECORD_DTYPE.itemsize == 20 and api.EVAL_RECORD_DTYPE.fields["packed_input"][1] == 4
    assert api.TRAIN_RECORD_DTYPE.itemsize == 40 and api.TRAIN_RECORD_DTYPE.fields["factor"][1] == 12
    assert api.OPT_ENTRY_DTYPE.itemsize == 16 and api.OPT_STATE_DTYPE.itemsize == 20


def test_argument_errors_are_reported_not_thrown(L):
    assert L.nrc_create(None, 0, None) != 0
    assert b"null" in L.nrc_last_error()
    assert L.nrc_adam_step(None, 1, None) != 0
    assert L.nrc_infer_encoded(None, None, None, 10, 0, None) != 0
    # n == 0 is a no-op that needs no device
    assert L.nrc_mlp_evaluate_encoded(None, None, None, 0, None) == 0
    assert L.nrc_mlp_gradient_encoded(None, None, None, None, 0, None) == 0
    # the stand-alone encode stage: empty input is a no-op, null buffers and a stride below one record are argument errors
    assert L.nrc_encode_inputs(None, 56, 0, None, None) == 0
    assert L.nrc_encode_inputs(None, 56, 10, None, None) != 0 and b"null" in L.nrc_last_error()
    buf = (C.c_uint8 * 4096)()
    assert L.nrc_encode_inputs(buf, 40, 4, buf, None) != 0 and b"stride" in L.nrc_last_error()
    assert L.nrc_encode_packed_inputs(buf, 12, 4, None, buf, None) != 0


def test_fails_loudly_without_a_gpu(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from vknrc_b200.api import nrc_config_t, NrcState, NrcError
    h = C.c_void_p()
    cfg = nrc_config_t(64, 64, 1)
    rc = L.nrc_create(C.byref(cfg), 0, C.byref(h))
    assert rc != 0 and not h.value and len(L.nrc_last_error()) > 0
    with pytest.raises(NrcError):
        NrcState(0, (64, 64), 1)


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under vknrc_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "vknrc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text and "libnrc_oracle" not in text, f


def test_headers_are_plain_c(tmp_path):
    """The boundary is a C ABI: both public headers must compile as C99 with nothing but <stdint.h>."""
    import subprocess
    tu = tmp_path / "tu.c"
    tu.write_text('#include <nrc_b200.h>\nint main(void) { nrc_config_t c = {1u, 1u, 0u}; NrcEvalRecord e; (void)c; (void)e;\n'
                  '  return sizeof(NrcEvalRecord) == 20 && sizeof(NrcTrainRecord) == 40 && sizeof(NrcMaterial) == 64 &&\n'
                  '         sizeof(NrcOptimizerEntry) == 16 && sizeof(NrcOptimizerState) == 20 && sizeof(NrcPrimRow) == 64 ? 0 : 1; }\n')
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    exe = tmp_path / "tu"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(tu), "-o", str(exe)],
                       capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    assert subprocess.run([str(exe)]).returncode == 0  # the struct sizes are the reference's (NRCRecord.glsl, Scene.glsl std430)
