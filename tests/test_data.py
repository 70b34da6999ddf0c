"""CPU: the reference's data layer restated (utils.py Dataset/get_batches/update_config,
model_v2.get_feed_dict) on a synthetic dataset in the reference's on-disk format, and the host
packing that feeds the packed reader."""
import numpy as np
import torch

from fvta_memexqa_b200.config import default_config, FVTA_FLAGS
from fvta_memexqa_b200.data import (Dataset, make_synthetic_raw, add_vocab, update_config, build_feed_dict,
                                    getAnswers, getEvalScore, grouper)
from fvta_memexqa_b200.packing import build_pack, to_blob, blob_views


class _M:   # the two attributes build_feed_dict reads from the model
    def __init__(self, config):
        self.config = config
        self.num_choice = 4


def _dataset(n=11, is_train=True):
    cfg = default_config(**dict(FVTA_FLAGS, batch_size=4, is_train=is_train, image_feat_dim=16))
    data, shared = make_synthetic_raw(n, num_albums=9, seed=3, feat_dim=16, emb_dim=100, vocab=60)
    add_vocab(shared, cfg)
    ds = Dataset(data, "train", shared=shared)
    update_config(cfg, [ds])
    return cfg, ds


def test_grouper_and_batches():
    assert grouper([1, 2, 3, 4, 5], 2) == [(1, 2), (3, 4), (5, None)]
    cfg, ds = _dataset()
    batches = list(ds.get_batches(4, num_steps=3, shuffle=False))
    assert [len(b[0]) for b in batches] == [4, 4, 3]          # short last batch (utils.py:117)
    b = batches[0][1]
    assert b.data["pidx2feat"].shape[1] == 16
    # every photo index points at the feature of the right photo
    for albums_idx, albums_ids in zip(b.data["photo_idxs"], b.data["photo_ids"]):
        for idxs, pids in zip(albums_idx, albums_ids):
            for i, pid in zip(idxs, pids):
                assert np.array_equal(b.data["pidx2feat"][i], ds.shared["pid2feat"][pid])


def test_update_config_clamps_like_the_reference():
    cfg, ds = _dataset()
    assert cfg.max_num_albums <= cfg.num_albums_thres and cfg.max_num_photos <= cfg.num_photos_thres
    assert cfg.max_sent_des_size <= cfg.sent_des_size_thres and cfg.max_word_size <= cfg.word_size_thres
    assert cfg.word_vocab_size == len(ds.shared["word2idx"]) and cfg.word_emb_size == 100
    cfg2, _ = _dataset(is_train=False)        # test mode clamps only description / photo-title length
    assert cfg2.max_num_photos >= cfg.max_num_photos


def test_feed_dict_semantics():
    cfg, ds = _dataset()
    (idxs, batch) = list(ds.get_batches(4, num_steps=3, shuffle=False))[2]    # the short batch
    rng = np.random.RandomState(0)
    feed = build_feed_dict(_M(cfg), batch, is_train=True, rng=rng)
    N = cfg.batch_size
    assert feed["q"].shape[0] == N and feed["y"].shape == (N, 4)
    n_ex = len(idxs)
    assert feed["y"][:n_ex].sum(1).tolist() == [1] * n_ex and feed["y"][n_ex:].sum() == 0   # padded rows
    assert feed["q_mask"][n_ex:].sum() == 0
    # masks are prefixes and ids are zero exactly where the mask is off
    for k in ("at", "ad", "when", "where", "pts", "q", "choices"):
        m = feed[k + "_mask"]
        assert np.all((feed[k] != 0) <= m)
        assert np.all(m[..., 1:] <= m[..., :-1])
    # word ids: trainable words < |word2idx| <= GloVe words; the correct answer sits where y says
    w2i, e2i = ds.shared["word2idx"], ds.shared["existing_word2idx"]
    i = 0
    slot = int(np.argmax(feed["y"][i]))
    ans = batch.data["y"][i]
    got = feed["choices"][i, slot, :len(ans)]
    want = [w2i[w] if w in w2i else e2i[w] + len(w2i) for w in ans][:cfg.max_answer_size]
    assert got.tolist()[:len(want)] == want
    # test mode places the answer at yidx
    feed_t = build_feed_dict(_M(cfg), batch, is_train=False)
    assert int(np.argmax(feed_t["y"][0])) == batch.data["yidx"][0]
    # evaluation helpers
    yp = np.eye(4)[[batch.data["yidx"][j] for j in range(n_ex)]]
    pred, gt = getAnswers(yp, (idxs, batch))
    assert getEvalScore(pred, gt) == 1.0


def test_pack_roundtrip():
    cfg, ds = _dataset()
    (idxs, batch) = next(iter(ds.get_batches(4, num_steps=1, shuffle=False)))
    feed = {k: torch.as_tensor(v) for k, v in build_feed_dict(_M(cfg), batch, is_train=True,
                                                              rng=np.random.RandomState(1)).items()}
    tensors, meta = build_pack(feed, cfg)
    # every valid token appears once, sequence after sequence; offsets are the exclusive prefix sum
    lens = tensors["text_seq_len"].long()
    assert int(lens.sum()) == meta["text_ntok"] <= meta["text_cap"] and meta["text_cap"] % 512 == 0
    assert torch.equal(tensors["text_seq_off"].long(), torch.cumsum(lens, 0) - lens)
    q0, q1 = meta["q_range"]
    n0 = int(tensors["text_seq_off"][q0])
    L0 = int(lens[q0])
    assert torch.equal(tensors["text_tok_ids"][n0:n0 + L0], feed["q"][0, :L0].long())
    # output rows: the question goes after the context tensor, choices store nothing
    assert int(tensors["text_seq_out"][q0]) == meta["hq_row0"]
    c0, c1 = meta["choices_range"]
    assert torch.all(tensors["text_seq_out"][c0:c1] == -1)
    # modality k, album m of sample n starts at row ((n*K+k)*M+m)*JMAX
    a0 = meta["q_range"][1]            # first album-title sequence
    assert int(tensors["text_seq_out"][a0 + 1]) == meta["JMAX"]
    assert tensors["hall_mask"].shape == (meta["N"], meta["K"], meta["T"])
    # one blob, typed views back
    blob, index = to_blob(tensors, pin=False)
    views = blob_views(blob, index)
    for k, v in tensors.items():
        assert torch.equal(views[k], v), k
