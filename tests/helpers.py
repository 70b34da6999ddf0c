"""Shared helpers for the parity tests (CUDA path vs the CPU oracle)."""
import torch

from fvta_memexqa_b200.config import fvta_config
from fvta_memexqa_b200.params import ParamStore
from fvta_memexqa_b200.synthetic import make_attention_inputs


def attn_case(N, K, M, JMAX, D, JQ, simi=2, add_tanh=True, use_time_warp=True, warp_type=5,
              time_warp_att=False, density=0.5, q_density=0.6, seed=0, dtype=torch.float32):
    """Config + params + inputs of one attention parity case (CPU tensors)."""
    assert D % 2 == 0
    cfg = fvta_config(batch_size=N, hidden_size=D // 2, simiMatrix=simi, add_tanh=add_tanh,
                      use_time_warp=use_time_warp, warp_type=warp_type, use_time_warp_att=time_warp_att)
    ps = ParamStore(cfg, seed=seed + 100, dtype=dtype)
    params = {k: v.clone() for k, v in ps.as_dict().items()}
    # zero-initialised biases would hide bias bugs
    g = torch.Generator().manual_seed(seed + 7)
    for k in params:
        if k.endswith("/b"):
            params[k] = torch.randn(params[k].shape, generator=g, dtype=dtype) * 0.1
    inp = make_attention_inputs(N, K, M, JMAX, D, JQ, seed=seed, density=density, dtype=dtype,
                                q_density=q_density)
    return cfg, params, inp


def rel_err(a, b, mask=None):
    """max |a-b| / (max|b| + tiny) over (optionally masked) entries."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    if mask is not None:
        a = a * mask
        b = b * mask
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()
