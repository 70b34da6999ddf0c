"""Generates the golden fixtures under tests/golden/ (run in the BUILD container, where
/root/reference is mounted; the GPU box never reads the reference).

  main_flags.json   the reference's argparse block (main.py:28-147), extracted textually:
                    flag -> {type, default, action}.  Pins the CLI boundary.
  fvta_tiny.npz     inputs, weights and every oracle output (forward tensors, loss, all
                    gradients, post-Adadelta weights) of one training step of the LITERAL
                    oracle in fp64 at the tiny shape of SURVEY 8c pin 3 (README FVTA flags).
  attn_c2.npz       attention-block inputs/outputs of the collapsed oracle at the C2 shape
                    (N=2), fp64 math on fp32 inputs.

PARITY UNPINNED: these vectors come from this repo's own CPU restatement, not from
TensorFlow (which cannot run here).  They freeze the restatement so that refactors of the
oracle or kernels cannot drift silently.
"""
import ast
import json
import os
import re
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def extract_flags(path="/root/reference/main.py"):
    src = open(path).read()
    flags = {}
    # one add_argument call per line in the reference
    for m in re.finditer(r"parser\.add_argument\((.*)\)\s*(?:#.*)?$", src, re.M):
        call = m.group(1)
        try:
            node = ast.parse("f(%s)" % call).body[0].value
        except SyntaxError:
            continue
        name = node.args[0].value
        kw = {}
        for k in node.keywords:
            if k.arg == "type":
                kw["type"] = k.value.id
            elif k.arg in ("default", "action"):
                try:
                    kw[k.arg] = ast.literal_eval(k.value)
                except ValueError:
                    kw[k.arg] = ast.unparse(k.value)
        flags[name] = kw
    return flags


def tiny_case():
    from fvta_memexqa_b200.config import fvta_config, set_shapes
    from fvta_memexqa_b200.params import ParamStore, NEVER_UPDATED
    from fvta_memexqa_b200.synthetic import make_feed
    from oracle import full
    cfg = fvta_config(batch_size=2, hidden_size=4, char_emb_size=6, char_out_size=5, image_feat_dim=11,
                      image_trans_dim=7)
    set_shapes(cfg, M=2, JI=2, JXP=3, JD=4, JXA=3, JT=2, JG=2, JQ=4, JA=3, W=6, word_vocab_size=20,
               char_vocab_size=15, word_emb_size=8)
    ps = ParamStore(cfg, seed=42)
    params = {k: v.double() for k, v in ps.as_dict().items()}
    g = torch.Generator().manual_seed(1)
    for k in params:
        if k.endswith("/b") or k.endswith("/bias"):
            params[k] = torch.randn(params[k].shape, generator=g, dtype=torch.float64) * 0.1
    feed = make_feed(cfg, seed=9, glove_vocab=10)
    feed = {k: (v.double() if v.is_floating_point() else v) for k, v in feed.items()}
    loss, grads, new_params, state, out = full.train_step(feed, params, {}, cfg, block="literal",
                                                          never_updated=NEVER_UPDATED)
    arrays = {}
    for k, v in feed.items():
        arrays["feed/" + k] = v.numpy()
    for k, v in params.items():
        arrays["param/" + k] = v.numpy()
    for k, v in grads.items():
        if v is not None:
            arrays["grad/" + k] = v.numpy()
    for k, v in new_params.items():
        arrays["new_param/" + k] = v.numpy()
    for k in ("hall", "hall_mask", "C", "warp_h", "att_logits", "g1_all", "gq", "q_att_logits", "logits", "yp"):
        arrays["out/" + k] = out[k].detach().numpy()
    arrays["out/loss"] = loss.numpy()
    np.savez_compressed(os.path.join(GOLD, "fvta_tiny.npz"), **arrays)


def attn_case_c2():
    from tests.helpers import attn_case
    from oracle import collapsed
    cfg, params, inp = attn_case(2, 6, 4, 64, 100, 25, seed=77)
    p64 = {k: v.double().requires_grad_(True) for k, v in params.items()}
    hall = inp["hall"].double().requires_grad_(True)
    hq = inp["hq"].double().requires_grad_(True)
    lq = inp["lq"].double().requires_grad_(True)
    o = collapsed.fvta_attention(hall, inp["hall_mask"], hq, inp["q_mask"], lq, p64, cfg)
    g = torch.Generator().manual_seed(3)
    dg1 = torch.randn(2, 100, generator=g, dtype=torch.float64)
    (o["g1_all"] * dg1).sum().backward()
    arrays = {"hall": inp["hall"].numpy(), "hall_mask": inp["hall_mask"].numpy(), "hq": inp["hq"].numpy(),
              "q_mask": inp["q_mask"].numpy(), "lq": inp["lq"].numpy(), "d_g1_all": dg1.numpy(),
              "g1_all": o["g1_all"].detach().numpy(), "c": o["c"].detach().numpy(),
              "att_logits": o["att_logits"].detach().numpy().astype(np.float32),
              "d_hall": hall.grad.numpy().astype(np.float32), "d_hq": hq.grad.numpy(), "d_lq": lq.grad.numpy()}
    for k, v in params.items():
        if k.startswith("time_warp/W") or k.startswith("attention/"):
            arrays["param/" + k] = v.numpy()
            arrays["grad/" + k] = p64[k].grad.numpy()
    np.savez_compressed(os.path.join(GOLD, "attn_c2.npz"), **arrays)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    flags = extract_flags()
    with open(os.path.join(GOLD, "main_flags.json"), "w") as f:
        json.dump(flags, f, indent=1, sort_keys=True)
    print("flags:", len(flags))
    tiny_case()
    attn_case_c2()
    for fn in sorted(os.listdir(GOLD)):
        print(fn, os.path.getsize(os.path.join(GOLD, fn)))
