"""CPU: the C-ABI library builds, loads, and exports every symbol include/fvta_b200.h declares
(no compute calls: there is no GPU here); the host layer refuses to run without CUDA."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "fvta_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fvta_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from fvta_memexqa_b200 import _lib
    from fvta_memexqa_b200.build import LIB_PATH
    lib = ctypes.CDLL(LIB_PATH)
    names = header_functions()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(_lib.SYMBOLS), "ctypes table and header disagree"
    assert _lib.load().fvta_version() == 100


def test_descriptor_queries_without_gpu():
    from fvta_memexqa_b200 import _lib
    lib = _lib.load()
    d = _lib.FvtaAttnDesc(N=32, K=6, T=256, D=100, JQ=25, simi=2, add_tanh=1, use_time_warp=1, warp_type=5,
                          window=3, time_warp_att=0, has_mask=1)
    assert lib.fvta_attn_workspace_bytes(ctypes.byref(d)) > 0
    assert lib.fvta_attn_stats_floats(ctypes.byref(d)) == 32 * 6 * (8 + 100)
    bad = _lib.FvtaAttnDesc(N=32, K=6, T=256, D=100, JQ=25, simi=7, add_tanh=1, use_time_warp=0, warp_type=1,
                            window=3, time_warp_att=0, has_mask=1)
    assert lib.fvta_attn_workspace_bytes(ctypes.byref(bad)) == 0      # unknown simiMatrix
    L = _lib.FvtaLstmDesc(B=128, J=8, In=200, d=50, path=0)
    assert lib.fvta_bilstm_reserve_floats(ctypes.byref(L)) == 2 * 128 * 8 * 6 * 50


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    """The product path must fail loudly without CUDA: no silent CPU/PyTorch fallback."""
    from fvta_memexqa_b200 import ops, _lib
    from fvta_memexqa_b200.config import fvta_config
    from fvta_memexqa_b200.model import get_model
    with pytest.raises(_lib.FvtaError):
        get_model(fvta_config(batch_size=2))
    with pytest.raises(_lib.FvtaError):
        ops.gemm(torch.zeros(2, 2), torch.zeros(2, 2))
    with pytest.raises(_lib.FvtaError):
        ops.attn_forward(ops.AttnConfig(2, True), torch.zeros(1, 1, 4, 4), None, torch.zeros(1, 2, 4), None,
                         None, torch.zeros(8, 1), torch.zeros(1))
