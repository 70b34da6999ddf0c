"""GPU parity of the whole FVTA step (embeddings -> reader -> context tensor -> attention ->
question attention -> output -> loss -> gradients -> Adadelta) against the CPU oracle, through
the reference-shaped API (get_model / Trainer.step / Tester.step).

Tolerances (BASELINE.json north_star): 1e-4 relative in fp32 for logits, loss, gradients and
post-update weights; argmax of yp identical."""
import pytest
import torch

from oracle import full as oracle_full
from fvta_memexqa_b200.config import fvta_config, set_shapes
from fvta_memexqa_b200.data import TensorBatch
from fvta_memexqa_b200.model import get_model
from fvta_memexqa_b200.params import NEVER_UPDATED
from fvta_memexqa_b200.synthetic import make_feed
from fvta_memexqa_b200.tester import Tester
from fvta_memexqa_b200.trainer import Trainer
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


def tiny_config(**kw):
    cfg = fvta_config(batch_size=3, hidden_size=4, char_emb_size=6, char_out_size=5, image_feat_dim=11,
                      image_trans_dim=7, **kw)
    set_shapes(cfg, M=2, JI=2, JXP=3, JD=4, JXA=3, JT=2, JG=2, JQ=4, JA=3, W=6, word_vocab_size=20,
               char_vocab_size=15, word_emb_size=8)
    return cfg


def run_case(cfg, block, n_ex=None, steps=2, glove=10):
    feed = make_feed(cfg, seed=5, num_examples=n_ex, glove_vocab=glove)
    model = get_model(cfg)
    trainer = Trainer(model, cfg)
    p0 = {k: v.detach().cpu().double() for k, v in model.params.as_dict().items()}
    # randomise biases so bias-gradient bugs cannot hide behind zeros
    g = torch.Generator().manual_seed(1)
    for k in p0:
        if k.endswith("/b") or k.endswith("/bias"):
            p0[k] = torch.randn(p0[k].shape, generator=g, dtype=torch.float64) * 0.1
            model.params.view(k).copy_(p0[k].float())
    p0 = {k: model.params.view(k).detach().cpu().double() for k in p0}
    feed64 = {k: (v.double() if v.is_floating_point() else v) for k, v in feed.items()}
    batch = TensorBatch(feed, cfg.batch_size if n_ex is None else n_ex)

    # inference parity first (Tester.step trims padded rows)
    tester = Tester(model, cfg)
    yp = tester.step(None, ((), batch))
    ref = oracle_full.forward(feed64, p0, cfg, block=block)
    assert yp.shape[0] == batch.num_examples
    assert rel_err(torch.as_tensor(yp), ref["yp"][:batch.num_examples]) < TOL
    assert (torch.as_tensor(yp).argmax(1) == ref["yp"][:batch.num_examples].argmax(1)).all()

    params, state = p0, {}
    for s in range(steps):
        loss, _, _ = trainer.step(None, ((), batch))
        rloss, rgrads, params, state, _ = oracle_full.train_step(feed64, params, state, cfg, block=block,
                                                                 never_updated=NEVER_UPDATED)
        assert abs(loss - rloss.item()) <= TOL * max(1.0, abs(rloss.item())), (s, loss, rloss.item())
        # gradients that are analytically ~0 (e.g. the bias of a softmax logit) are pure rounding
        # noise: measure every error against at least 1e-3 of the largest gradient in the model
        gmax = max(g.abs().max().item() for g in rgrads.values() if g is not None)
        for k, gref in rgrads.items():
            if gref is None:
                continue
            got = model.params.grad(k).detach().cpu().double()
            err = (got - gref).abs().max().item()
            scale = max(gref.abs().max().item(), 1e-3 * gmax)
            assert err < 5 * TOL * scale + 2e-7, (s, k, err, scale)
        for k in params:
            assert rel_err(model.params.view(k), params[k]) < 5 * TOL, (s, k)
    assert model.global_step == steps
    return model


def test_tiny_step_vs_literal_oracle(cuda_device):
    """Tiny shape (SURVEY 8c pin 3) against the op-for-op literal oracle, incl. a short batch."""
    run_case(tiny_config(), "literal", n_ex=2)


@pytest.mark.parametrize("kw", [
    dict(),                                           # README FVTA flags
    dict(use_question_att=False),                     # FVTA_no_question_att ablation
    dict(use_time_warp=False),                        # FVTA_no_timewarp ablation
    dict(no_photo=True),                              # FVTA_no_photo ablation (K=5)
    dict(simiMatrix=4),                               # FVTA cosine
    dict(simiMatrix=1, add_tanh=False, use_eu_output=True),
    dict(simiMatrix=3, warp_type=3, use_time_warp_att=True),
    dict(wd=0.002),
    dict(use_char=False, use_image_trans=False),
], ids=lambda kw: ",".join("%s=%s" % kv for kv in kw.items()) or "fvta")
def test_tiny_step_flag_matrix(cuda_device, kw):
    run_case(tiny_config(**kw), "literal", steps=1)


def test_c2_step_vs_collapsed_oracle(cuda_device):
    """C2 shape (README training flags, N=4 to keep the CPU oracle in seconds)."""
    cfg = fvta_config(batch_size=4)
    run_case(cfg, "collapsed", steps=1, glove=50)


def test_sparse_adadelta_on_embedding_tables(cuda_device):
    """The embedding tables are read through tf.nn.embedding_lookup, so TF updates them with
    sparse_apply_adadelta: a row that the CURRENT batch does not look up keeps its accumulators
    (a dense update would decay them by rho and change the row's NEXT update).  Batches A, B, A:
    rows looked up by A only see the difference at step 3."""
    cfg = tiny_config()
    feeds = [make_feed(cfg, seed=s, glove_vocab=10) for s in (5, 6, 5)]
    # make batch B avoid a few word / char ids that batch A uses, so the two sets really differ
    for k in ("q", "at", "ad", "when", "where", "pts", "choices"):
        feeds[1][k] = torch.where((feeds[1][k] >= 2) & (feeds[1][k] < 8), torch.full_like(feeds[1][k], 9), feeds[1][k])
        feeds[1][k + "_c"] = torch.where((feeds[1][k + "_c"] >= 2) & (feeds[1][k + "_c"] < 6),
                                         torch.full_like(feeds[1][k + "_c"], 7), feeds[1][k + "_c"])
    model = get_model(cfg)
    trainer = Trainer(model, cfg)
    params = {k: model.params.view(k).detach().cpu().double() for k in model.params.offsets}
    dense = dict(params)
    state, dstate = {}, {}
    for feed in feeds:
        feed64 = {k: (v.double() if v.is_floating_point() else v) for k, v in feed.items()}
        trainer.step(None, ((), TensorBatch(feed, cfg.batch_size)))
        _, _, params, state, _ = oracle_full.train_step(feed64, params, state, cfg, block="literal",
                                                        never_updated=NEVER_UPDATED)
        saved = dict(oracle_full.SPARSE_TABLES)
        oracle_full.SPARSE_TABLES.clear()                   # the dense variant, to prove the test can tell
        try:
            _, _, dense, dstate, _ = oracle_full.train_step(feed64, dense, dstate, cfg, block="literal",
                                                            never_updated=NEVER_UPDATED)
        finally:
            oracle_full.SPARSE_TABLES.update(saved)
    told_apart = False
    for k in ("emb/var/char_emb", "emb/word/var/word_emb_mat"):
        got = model.params.view(k).detach().cpu().double()
        assert rel_err(got, params[k]) < 5 * TOL, k
        # the accumulators are where sparse and dense differ first (a factor rho on every row that a
        # step did not look up); the weights follow one step later
        for which, buf in ((0, model.params.accum), (1, model.params.accum_update)):
            acc = model.params.view(k, buf).detach().cpu().double()
            assert rel_err(acc, state[k][which]) < 5 * TOL, (k, which)
            told_apart |= rel_err(dstate[k][which], state[k][which]) > 1e-3
    assert told_apart, "batches A and B look up the same rows: the test cannot see sparse vs dense"


def test_step_vis_surface(cuda_device):
    cfg = tiny_config()
    model = get_model(cfg)
    feed = make_feed(cfg, seed=2, num_examples=2, glove_vocab=10)
    out = Tester(model, cfg).step_vis(None, ((), TensorBatch(feed, 2)))
    assert len(out) == 28
    yp, C, C_win, att_logits, q_att = out[:5]
    T = 2 * 6
    assert yp.shape == (2, 4) and C.shape == (2, T, T) and att_logits.shape == (2, 6, T, 4)
    # tester.py:27-28 trim() turns every non-ndarray (the scalar window) into -1
    assert q_att.shape == (2, 4, 1) and C_win == -1


def test_main_train_then_test_roundtrip(cuda_device, tmp_path):
    """The reference's CLI end to end on a synthetic dataset in the reference's on-disk format:
    main.py --is_train (list-of-lists batches -> get_feed_dict -> Trainer.step, periodic save +
    validation, best checkpoint) then --is_test --load_best (restore, Tester.step, answers.json)."""
    import json
    import os
    import main as cli
    from fvta_memexqa_b200.config import get_args
    base = [str(tmp_path / "prepro"), str(tmp_path / "models"), "--modelname", "fvta_t", "--synthetic", "24",
            "--batch_size", "4", "--hidden_size", "8", "--image_feat_dim", "16", "--use_image_trans",
            "--image_trans_dim", "12", "--use_char", "--char_emb_size", "6", "--char_out_size", "8", "--add_tanh",
            "--simiMatrix", "2", "--use_question_att", "--use_3d", "--use_time_warp", "--warp_type", "5",
            "--num_albums_thres", "4", "--num_photos_thres", "8"]
    cfg = get_args(base + ["--is_train", "--num_epochs", "1", "--save_period", "3", "--val_num_batches", "2"])
    model = cli.train(cfg)
    assert model.global_step == 6                                   # ceil(24/4) steps
    assert os.path.exists(cfg.self_summary_path)
    assert any(f.endswith(".pt") for f in os.listdir(cfg.save_dir))
    assert any(f.endswith(".pt") for f in os.listdir(cfg.save_dir_best))
    cfg_t = get_args(base + ["--is_test", "--load_best", "--get_yp"])
    acc = cli.test(cfg_t)
    assert 0.0 <= acc <= 1.0
    answers = json.load(open(os.path.join(cfg_t.val_path, "answers.json")))
    assert len(answers) == 6 and all(0 <= v < 4 for v in answers.values())   # test split = synthetic // 4


def _dp_worker(rank, world, port, q, fused=False):
    import os
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    from fvta_memexqa_b200.dp import DataParallel
    dp = DataParallel(backend="nccl")
    cfg = tiny_config()
    cfg.batch_size = 2
    G = 4
    gcfg = tiny_config()
    gcfg.batch_size = G
    gfeed = make_feed(gcfg, seed=5, glove_vocab=10, num_photos_pool=16)
    lo, hi = dp.shard(G)
    batched = {k for k, v in gfeed.items() if v.shape[0] == G and k not in ("image_emb_mat", "existing_emb_mat")}
    feed = {k: (v[lo:hi] if k in batched else v) for k, v in gfeed.items()}
    model = get_model(cfg)
    dp.broadcast_params(model.params.flat)
    trainer = Trainer(model, cfg, dp=dp)
    if fused:
        assert dp.enable_p2p(model.params), "peer-memory all-reduce could not be set up"
    for _ in range(3):          # three steps: both staging halves and the flag epochs are exercised
        loss, _, _ = trainer.step(None, ((), TensorBatch(feed, hi - lo)))
    torch.cuda.synchronize()
    flat = model.params.flat.detach().clone()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    torch.distributed.all_gather(gathered, flat)
    if rank == 0:
        q.put(({k: model.params.view(k).detach().cpu() for k in model.params.offsets},
               [bool(torch.equal(g, flat)) for g in gathered]))
    dp.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("fused", [False, True], ids=["nccl", "p2p-fused"])
def test_two_gpu_step_equals_one_gpu_step(cuda_device, fused):
    """Data parallel (north_star d): two ranks, each half of a 4-sample batch, the flat gradient summed
    over the ranks -- by ncclAllReduce + fvta_adadelta, or by the fused peer-memory kernel
    fvta_p2p_allreduce_adadelta -- gives the same post-Adadelta weights as one GPU on the whole batch
    (padded shapes pinned) over three steps.  The fused kernel sums in rank order on every rank: its
    replicas must be bitwise identical.  Skipped on a single-GPU box."""
    import socket
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q, fused)) for r in range(2)]
    for p in procs:
        p.start()
    got, same = q.get(timeout=240)
    if fused:
        assert all(same), "replicas diverged bitwise under the fused all-reduce"
    for p in procs:
        p.join(timeout=60)
    gcfg = tiny_config()
    gcfg.batch_size = 4
    gfeed = make_feed(gcfg, seed=5, glove_vocab=10, num_photos_pool=16)
    model = get_model(gcfg)
    one = Trainer(model, gcfg)
    for _ in range(3):
        one.step(None, ((), TensorBatch(gfeed, 4)))
    for k, v in got.items():
        ref = model.params.view(k).detach().cpu()
        assert rel_err(v, ref) < 1e-4 or (v - ref).abs().max() < 1e-6, k


def test_pipeline_feeds_and_resident_tables(cuda_device):
    """A raw dataset in the reference's on-disk format through the input routes: the literal get_feed_dict
    restatement (data.build_feed_dict) and the vectorised store must train to the same losses and
    weights (same placeholders); so must the store at pinned shapes with per-batch photo tables and the
    loader (packed by worker processes, photo / GloVe tables resident on the device): same arithmetic,
    only where the bytes live differs.  (Shrunk and pinned shapes are not compared with each other: the
    padded sizes enter the model)"""
    import numpy as np
    from fvta_memexqa_b200.config import default_config, FVTA_FLAGS
    from fvta_memexqa_b200.data import Dataset, make_synthetic_raw, add_vocab, update_config, build_feed_dict
    from fvta_memexqa_b200.pipeline import TokenizedStore, BatchLoader, resident_photo_table

    def setup():
        cfg = default_config(**dict(FVTA_FLAGS, batch_size=4, is_train=True, image_feat_dim=16, hidden_size=8,
                                    image_trans_dim=12, seed=3))
        data, shared = make_synthetic_raw(12, num_albums=9, seed=3, feat_dim=16, emb_dim=100, vocab=60)
        add_vocab(shared, cfg)
        ds = Dataset(data, "train", shared=shared)
        update_config(cfg, [ds])
        torch.manual_seed(0)
        model = get_model(cfg)
        return cfg, ds, model, Trainer(model, cfg)

    order = None
    results = []
    for route in ("literal", "store", "store-pinned", "loader"):
        cfg, ds, model, trainer = setup()
        store = TokenizedStore(ds, cfg)
        loader = BatchLoader(store, model, 4, num_steps=3, shuffle=True, seed=5, workers=0 if route != "loader" else 2)
        tasks = list(loader._tasks())
        losses = []
        if route == "loader":
            model.set_resident_tables(image_emb_mat=resident_photo_table(store, model.device),
                                      existing_emb_mat=store.existing_emb_mat)
            loader.pin_caps()
            for batch in loader:
                losses.append(trainer.step(None, batch)[0])
            loader.close()
        else:
            for step, idxs, seed, *_ in tasks:
                rng = np.random.RandomState((seed * 1000003 + step) & 0x7fffffff)
                if route == "literal":
                    sub = next(b for i, b in _batches_by_idxs(ds, [idxs]))
                    feed = build_feed_dict(model, sub, is_train=True, rng=rng)
                else:
                    feed = store.feed(idxs, is_train=True, rng=rng, pin_shapes=route == "store-pinned")
                feed = {k: torch.as_tensor(v) for k, v in feed.items()}
                losses.append(trainer.step(None, ((), TensorBatch(feed, len(idxs))))[0])
        results.append((losses, model.params.flat.detach().cpu().clone()))
    # (not bitwise: split-K products and the attention backward accumulate with atomics, so two runs of the
    # SAME inputs differ in the last bits)
    for a, b in ((0, 1), (2, 3)):
        for x, y in zip(results[a][0], results[b][0]):
            assert abs(x - y) <= 1e-5 * abs(x), (results[a][0], results[b][0])
        assert rel_err(results[a][1], results[b][1]) < 1e-5


def _batches_by_idxs(ds, groups):
    """Dataset.get_batches' batch construction (utils.py:117-198) for explicit index groups."""
    from collections import defaultdict
    import numpy as np
    from fvta_memexqa_b200.data import Dataset
    feat_dim = len(next(iter(ds.shared["pid2feat"].values())))
    for idxs in groups:
        bd = ds.get_by_idxs(idxs)
        pid2idx = {}
        for albumIds in bd["aid"]:
            for a in albumIds:
                for pid in ds.shared["albums"][a]["photo_ids"]:
                    pid2idx.setdefault(pid, len(pid2idx))
        feats = np.zeros((len(pid2idx), feat_dim), dtype="float32")
        for pid, i in pid2idx.items():
            feats[i] = ds.shared["pid2feat"][pid]
        bd["pidx2feat"] = feats
        cols = defaultdict(list)
        for albumIds in bd["aid"]:
            c = defaultdict(list)
            for a in albumIds:
                al = ds.shared["albums"][a]
                for dst, src in (("album_title", "title"), ("album_description", "description"), ("where", "where"),
                                 ("when", "when"), ("photo_titles", "photo_titles")):
                    c[dst].append(al[src])
                    c[dst + "_c"].append(al[src + "_c"])
                c["photo_idxs"].append([pid2idx[p] for p in al["photo_ids"]])
                c["photo_ids"].append(al["photo_ids"])
            for k, v in c.items():
                cols[k].append(v)
        bd.update(cols)
        yield idxs, Dataset(bd, ds.datatype, shared=ds.shared)


def test_step_handle_reads_the_same_losses(cuda_device):
    """Trainer.step_handle (loss copied back behind the step, read later) vs Trainer.step, with and without
    the CUDA graph: same losses (1e-5: the step accumulates with atomics)."""
    cfg = tiny_config()
    feed = make_feed(cfg, seed=5, glove_vocab=10)
    batch = ((), TensorBatch(feed, cfg.batch_size))
    losses = {}
    for mode in ("step", "handle", "handle+graph"):
        torch.manual_seed(0)
        trainer = Trainer(get_model(cfg), cfg)
        if mode.endswith("graph"):
            trainer.enable_graph()
        if mode == "step":
            losses[mode] = [trainer.step(None, batch)[0] for _ in range(4)]
        else:
            hs = [trainer.step_handle(None, batch) for _ in range(4)]     # four outstanding
            losses[mode] = [h.result() for h in hs]
    for mode in ("handle", "handle+graph"):
        for a, b in zip(losses["step"], losses[mode]):
            assert abs(a - b) <= 1e-5 * max(1.0, abs(a)), (mode, losses)
    assert losses["step"][0] != losses["step"][3]          # the weights moved
