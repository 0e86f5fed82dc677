"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol declared in
include/exl_b200.h, the pybind shim exports the reference's 16 names, and compute calls fail loudly
(no CPU fallback) when no GPU is present."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REFERENCE_EXPORTS = ["set_tuning_params", "prepare_buffers", "cleanup", "make_q4", "q4_matmul", "q4_matmul_lora",
                     "q4_attn", "q4_attn_2", "q4_mlp", "column_remap", "rms_norm", "rope_", "half_matmul",
                     "half_matmul_cublas", "rep_penalty", "apply_rep_penalty"]   # exllama_ext.cpp:743-762


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "exl_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(exl_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from exllama_b200 import capi
    lib = capi.lib()
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/exl_b200.h but not exported"
    assert set(names) == set(capi.SIGNATURES), set(names) ^ set(capi.SIGNATURES)


def test_pybind_surface_matches_reference():
    from exllama_b200 import cuda_ext
    for n in REFERENCE_EXPORTS:
        assert hasattr(cuda_ext.exllama_ext, n), n
    for n in ["ext_make_q4", "ext_q4_matmul", "ext_half_matmul", "ext_rope_", "ext_rms_norm", "ext_rms_norm_",
              "ext_rep_penalty_mask_cpu", "ext_apply_rep_penalty_mask_cpu", "none_tensor", "exllama_ext"]:
        assert hasattr(cuda_ext, n), n
    assert cuda_ext.none_tensor.is_meta


def test_reference_error_messages():
    """dtype / shape errors surface as RuntimeError with the reference's wording (exllama_ext.cpp:53-58,166-172)."""
    import torch
    from exllama_b200 import cuda_ext
    qw = torch.zeros((8, 16), dtype=torch.int32)
    qz = torch.zeros((1, 2), dtype=torch.int32)
    sc = torch.zeros((1, 16), dtype=torch.float16)
    with pytest.raises(RuntimeError, match="scales is incorrect datatype, must be kHalf"):
        cuda_ext.ext_make_q4(qw, qz, sc.float(), None, 0)
    with pytest.raises(RuntimeError, match="qweight and qzeros have incompatible shapes"):
        cuda_ext.ext_make_q4(qw, torch.zeros((1, 3), dtype=torch.int32), sc, None, 0)
    with pytest.raises(RuntimeError, match="x is incorrect datatype, must be kHalf"):
        cuda_ext.rms_norm(torch.zeros(1, 8), torch.zeros(8, dtype=torch.half), torch.zeros(1, 8, dtype=torch.half), 1e-6)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from exllama_b200 import capi
    lib = capi.lib()
    h = C.c_void_p()
    rc = lib.exl_make_q4(None, None, None, None, 64, 32, 1, 0, None, C.byref(h))
    assert rc != 0 and b"no CPU fallback" in lib.exl_last_error()


def test_rep_penalty_through_c_abi(oracle):
    """CPU op of the surface: bit-exact against the oracle (and through it the reference, test_oracle.py)."""
    import numpy as np
    from exllama_b200 import capi
    lib = capi.lib()
    rng = np.random.default_rng(0)
    for (vocab, n, pmax, sustain, decay) in [(16, 4, 1.15, 2, 2), (1000, 300, 1.3, 64, 128), (50, 10, 1.2, -1, 0),
                                             (50, 7, 1.1, 0, 5), (32000, 2048, 1.18, 256, 256)]:
        seq = rng.integers(0, vocab, size=n).astype(np.int64)
        mask = np.empty(vocab, dtype=np.float32)
        assert lib.exl_rep_penalty(vocab, seq.ctypes.data, mask.ctypes.data, pmax, sustain, decay, n) == 0
        np.testing.assert_array_equal(mask, oracle.rep_penalty(vocab, seq, pmax, sustain, decay))
        logits = rng.standard_normal((1, vocab)).astype(np.float32)
        mine = logits.copy()
        assert lib.exl_apply_rep_penalty(vocab, seq.ctypes.data, pmax, sustain, decay, n, mine.ctypes.data) == 0
        np.testing.assert_array_equal(mine, oracle.apply_rep_penalty(seq[None], pmax, sustain, decay, logits))
