"""CPU, world_size 2 over gloo: the host-side logic of data parallelism (north_star d) --
sharding, the flat-buffer sum-all-reduce and its 1/world scale -- and, with the CPU oracle
standing in for the kernels, that a 2-rank step equals the 1-rank step on the concatenated
batch when padded shapes are pinned."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from fvta_memexqa_b200.dp import DataParallel
    from fvta_memexqa_b200.config import fvta_config, set_shapes
    from fvta_memexqa_b200.params import ParamStore, NEVER_UPDATED
    from fvta_memexqa_b200.synthetic import make_feed
    from oracle import full
    from oracle.tf_ops import adadelta_update
    torch.set_num_threads(1)
    dp = DataParallel(backend="gloo")
    lo, hi = dp.shard(5)
    assert (lo, hi) == ((0, 3) if rank == 0 else (3, 5))

    def cfg_for(n):
        cfg = fvta_config(batch_size=n, hidden_size=4, char_emb_size=6, char_out_size=5, image_feat_dim=11,
                          image_trans_dim=7)
        set_shapes(cfg, M=2, JI=2, JXP=3, JD=4, JXA=3, JT=2, JG=2, JQ=4, JA=3, W=6, word_vocab_size=20,
                   char_vocab_size=15, word_emb_size=8)
        return cfg
    G = 4                                                  # global batch, 2 per rank
    gfeed = make_feed(cfg_for(G), seed=3, glove_vocab=10, num_photos_pool=16)
    gfeed = {k: (v.double() if v.is_floating_point() else v) for k, v in gfeed.items()}
    batched = {k for k, v in gfeed.items() if v.shape[0] == G and k not in ("image_emb_mat", "existing_emb_mat")}
    lo, hi = dp.shard(G)
    feed = {k: (v[lo:hi] if k in batched else v) for k, v in gfeed.items()}
    cfg = cfg_for(hi - lo)
    ps = ParamStore(cfg, seed=11, dtype=torch.float64)     # same seed on every rank = replicated weights
    dp.broadcast_params(ps.flat)
    params = ps.as_dict()
    leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    loss = full.forward(feed, leaves, cfg, is_train=True, block="literal")["loss"]
    names = [k for k in leaves if k not in NEVER_UPDATED]
    grads = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    for k, g in zip(names, grads):
        if g is not None:
            ps.grad(k).copy_(g)
    scale = dp.all_reduce_grads(ps.grad_flat)              # ONE collective over ONE flat buffer
    assert scale == 0.5
    new_flat, _, _ = adadelta_update(ps.flat, ps.grad_flat * scale, ps.accum, ps.accum_update, cfg.init_lr)
    gl = dp.global_mean(loss.detach())
    if rank == 0:
        # single-process reference on the whole batch
        cfg1 = cfg_for(G)
        ps1 = ParamStore(cfg1, seed=11, dtype=torch.float64)
        l1, g1, p1, _, _ = full.train_step(gfeed, ps1.as_dict(), {}, cfg1, block="literal",
                                           never_updated=NEVER_UPDATED)
        err_loss = abs(l1.item() - gl.item())
        err_g = max((ps.grad(k) * scale - g).abs().max().item() for k, g in g1.items() if g is not None)
        err_p = max((new_flat[ps.offsets[k][0]:ps.offsets[k][0] + ps.offsets[k][1]].view(ps.offsets[k][2])
                     - p1[k]).abs().max().item() for k in p1 if k not in NEVER_UPDATED)
        q.put((err_loss, err_g, err_p))
    dp.barrier()
    dist.destroy_process_group()


def test_two_rank_step_equals_single_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    err_loss, err_g, err_p = q.get(timeout=5)
    assert err_loss < 1e-12 and err_g < 1e-12 and err_p < 1e-10, (err_loss, err_g, err_p)
