"""CPU: the vectorised input pipeline (fvta_memexqa_b200/pipeline.py) reproduces the reference's
get_feed_dict restatement (data.build_feed_dict, model_v2.py:1099-1565) bit for bit -- every
placeholder, every batch, training and testing -- and the loader hands batches over in order."""
import numpy as np
import pytest
import torch

from fvta_memexqa_b200.config import default_config, FVTA_FLAGS
from fvta_memexqa_b200.data import Dataset, make_synthetic_raw, add_vocab, update_config, build_feed_dict
from fvta_memexqa_b200.pipeline import TokenizedStore


class _M:
    def __init__(self, config):
        self.config = config
        self.num_choice = 4


def _dataset(n=23, is_train=True, **over):
    cfg = default_config(**dict(FVTA_FLAGS, batch_size=4, is_train=is_train, image_feat_dim=16, **over))
    data, shared = make_synthetic_raw(n, num_albums=9, seed=3, feat_dim=16, emb_dim=100, vocab=60)
    add_vocab(shared, cfg)
    ds = Dataset(data, "train", shared=shared)
    update_config(cfg, [ds])
    return cfg, ds


@pytest.mark.parametrize("is_train", [True, False])
@pytest.mark.parametrize("thres", [{}, {"num_albums_thres": 2, "num_photos_thres": 3, "sent_des_size_thres": 4,
                                        "sent_photo_title_size_thres": 3, "answer_size_thres": 2}])
def test_store_feed_equals_build_feed_dict(is_train, thres):
    cfg, ds = _dataset(is_train=is_train, **thres)
    store = TokenizedStore(ds, cfg)
    for idxs, batch in ds.get_batches(4, num_steps=6, shuffle=False):       # includes the short last batch
        ref = build_feed_dict(_M(cfg), batch, is_train=is_train, rng=np.random.RandomState(7))
        got = store.feed(idxs, is_train=is_train, rng=np.random.RandomState(7))
        assert set(got) == set(ref)
        for k in ref:
            assert got[k].shape == ref[k].shape, (k, got[k].shape, ref[k].shape)
            assert got[k].dtype == ref[k].dtype, (k, got[k].dtype, ref[k].dtype)
            assert np.array_equal(got[k], ref[k]), k


def test_resident_photos_and_pinned_shapes():
    """pis over the resident table addresses the same features as the per-batch table; pinned shapes are
    the config maxima for every batch."""
    cfg, ds = _dataset()
    store = TokenizedStore(ds, cfg)
    shapes = set()
    for idxs, batch in ds.get_batches(4, num_steps=6, shuffle=False):
        loc = store.feed(idxs, is_train=False)
        res = store.feed(idxs, is_train=False, resident_photos=True, pin_shapes=True)
        assert "image_emb_mat" not in res
        m = loc["pis_mask"]
        JI, M = m.shape[2], m.shape[1]
        a = loc["image_emb_mat"][loc["pis"][m]]
        b = store.photo_feats[res["pis"][:, :M, :JI][m]]
        assert np.array_equal(a, b)
        shapes.add(tuple((k, v.shape) for k, v in sorted(res.items())))
        assert res["at"].shape[1:] == (cfg.max_num_albums, cfg.max_sent_album_title_size)
        assert res["q"].shape[1] == cfg.max_question_size
    assert len(shapes) == 1


class _Params:
    touch_rows, touch_len = {}, 0


class _Model:
    def __init__(self, config):
        self.config = config
        self.params = _Params()


@pytest.mark.parametrize("workers", [0, 2])
def test_loader_order_and_content(workers):
    """Worker processes (spawned) and the inline path give the same batches, in get_batches order."""
    from fvta_memexqa_b200.pipeline import BatchLoader
    from fvta_memexqa_b200.packing import host_step_input
    cfg, ds = _dataset()
    store = TokenizedStore(ds, cfg)
    loader = BatchLoader(store, _Model(cfg), 4, num_steps=7, shuffle=True, seed=5, workers=workers, depth=3)
    import random
    order = random.Random(5).sample(list(ds.valid_idxs), len(ds.valid_idxs))
    seen = 0
    for step, (idxs, hb) in enumerate(loader):     # (a batch lives in a ring slot: look at it before moving on)
        seen += 1
        b = step % 6
        assert idxs == tuple(order[b * 4:(b + 1) * 4])
        feed = store.feed(idxs, is_train=True, rng=np.random.RandomState((5 * 1000003 + step) & 0x7fffffff),
                          pin_shapes=True, resident_photos=True)
        feed["is_train"] = True
        ref = host_step_input(feed, cfg, pin=False)
        hsi = hb.step_input(None, True)
        assert hsi["index"] == ref["index"] and hsi["meta"] == ref["meta"]
        from fvta_memexqa_b200.packing import blob_views
        a, b = blob_views(hsi["blob"].cpu(), hsi["index"]), blob_views(ref["blob"], ref["index"])
        for k in b:                       # (the padding bytes between the tensors are not initialised)
            assert torch.equal(a[k], b[k]), k
    loader.close()
    assert seen == 7


def test_token_counts_match_the_pack():
    from fvta_memexqa_b200.packing import build_pack
    cfg, ds = _dataset()
    store = TokenizedStore(ds, cfg)
    text, img = store.token_counts()
    for idxs, _ in ds.get_batches(4, num_steps=6, shuffle=False):
        feed = store.feed(idxs, is_train=True, rng=np.random.RandomState(1), pin_shapes=True, resident_photos=True)
        _, meta = build_pack({k: torch.as_tensor(v) for k, v in feed.items()}, cfg)
        assert meta["text_ntok"] == int(text[list(idxs)].sum())
        assert meta["img_ntok"] == int(img[list(idxs)].sum())
