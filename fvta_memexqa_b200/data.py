"""Datasets and batches.

``TensorBatch`` / ``SyntheticDataset`` feed the model with ready tensors at pinned padded
shapes (benchmarks, parity tests, data-parallel runs).  They expose what ``Trainer.step`` /
``Tester.step`` read from the reference's ``Dataset`` (utils.py:31-198): ``num_examples`` and
the ``(idxs, batch)`` tuples of ``get_batches``.
"""
import math
import random

import torch

from .synthetic import make_feed


class TensorBatch:
    """One batch already in placeholder form (model_v2.py:420-477 names)."""

    def __init__(self, feed, num_examples, pin=False):
        if pin and torch.cuda.is_available():
            feed = {k: (v.pin_memory() if isinstance(v, torch.Tensor) else v) for k, v in feed.items()}
        self.feed = feed
        self.num_examples = num_examples
        self.data = {}
        self.shared = {}
        self._step_input = {}

    def step_input(self, model, is_train):
        """The packed, pinned host buffer of this batch (Model.host_step_input), built once: what a
        data-loader worker hands to the training loop."""
        key = (is_train, getattr(model, "pin_text_cap", 0), getattr(model, "pin_img_cap", 0))
        if key not in self._step_input:
            feed = dict(self.feed)
            feed["is_train"] = is_train
            self._step_input[key] = model.host_step_input(feed)
        return self._step_input[key]

    def h2d_bytes(self):
        return sum(v.numel() * v.element_size() for v in self.feed.values() if isinstance(v, torch.Tensor))


class SyntheticDataset:
    """``num_examples`` synthetic QAs served in fixed-order or shuffled batches like
    ``Dataset.get_batches`` (utils.py:89-198): the last batch of an epoch is short and the
    model pads it with all-zero rows up to the static ``batch_size``."""

    def __init__(self, config, num_examples, seed=1234, pin=False, distinct_batches=4):
        self.config = config
        self.num_examples = num_examples
        self.valid_idxs = list(range(num_examples))
        self.seed = seed
        self.pin = pin
        self.distinct = distinct_batches
        self._cache = {}

    def _batch(self, b, n_ex):
        key = (b % self.distinct, n_ex)
        if key not in self._cache:
            cfg = self.config
            # fixed-size photo table so every batch has identical shapes (CUDA-graph replay)
            # pidx2feat holds the batch's own photos (utils.py:127-143); a fixed size (about the
            # expected number of valid photos, rounded up) keeps every batch the same shape
            pool = (cfg.batch_size * cfg.max_num_albums * cfg.max_num_photos * 7 // 16 + 63) // 64 * 64
            feed = make_feed(cfg, seed=self.seed + 17 * key[0], num_examples=n_ex, num_photos_pool=pool)
            self._cache[key] = TensorBatch(feed, n_ex, pin=self.pin)
        return self._cache[key]

    def get_batches(self, batch_size, num_steps=None, shuffle=True, cap=False):
        per_epoch = int(math.ceil(self.num_examples / float(batch_size)))
        if num_steps is None or (cap and num_steps > per_epoch):
            num_steps = per_epoch
        idxs = list(self.valid_idxs)
        if shuffle:
            random.Random(self.seed).shuffle(idxs)
        for s in range(num_steps):
            b = s % per_epoch
            chunk = tuple(idxs[b * batch_size:(b + 1) * batch_size])
            yield chunk, self._batch(b, len(chunk))


# =============================================================================================
# The reference's own data layer (utils.py) -- list-of-lists batches, same keys, same order.
# =============================================================================================
import itertools  # noqa: E402
from collections import defaultdict  # noqa: E402

import numpy as np  # noqa: E402


def grouper(l, n):
    """utils.py:12-18: chunks of n, the last one filled with None."""
    args = [iter(l)] * n
    return list(itertools.zip_longest(*args, fillvalue=None))


class Dataset:
    """Mirrors the reference's ``utils.Dataset`` (utils.py:31-198): ``data`` is a dict of
    per-example lists (q, cq, y, cy, yidx, aid, qid, idxs, cs, ccs), ``shared`` holds albums,
    pid2feat and the vocabularies.  ``get_batches`` joins the album information of each batch
    and builds its photo-feature matrix ``pidx2feat`` exactly as the reference does."""

    def __init__(self, data, datatype, shared=None, valid_idxs=None):
        self.data = data
        self.datatype = datatype
        self.shared = shared
        self.valid_idxs = list(range(self.get_data_size())) if valid_idxs is None else list(valid_idxs)
        self.num_examples = len(self.valid_idxs)

    def get_data_size(self):
        return len(next(iter(self.data.values())))

    def get_by_idxs(self, idxs):
        out = defaultdict(list)
        for key, val in self.data.items():
            out[key].extend(val[idx] for idx in idxs)
        return out

    def get_batches(self, batch_size, num_steps, shuffle=True, cap=False, rng=None):
        """utils.py:89-198."""
        num_batches_per_epoch = int(math.ceil(self.num_examples / float(batch_size)))
        if cap and (num_steps > num_batches_per_epoch):
            num_steps = num_batches_per_epoch
        num_epochs = int(math.ceil(num_steps / float(num_batches_per_epoch)))
        if shuffle:
            rnd = rng or random
            random_idxs = rnd.sample(self.valid_idxs, len(self.valid_idxs))   # same order every epoch (:104)
            grouped = lambda: list(grouper(random_idxs, batch_size))          # noqa: E731
        else:
            grouped = lambda: list(grouper(self.valid_idxs, batch_size))      # noqa: E731
        batch_idxs_iter = itertools.chain.from_iterable(grouped() for _ in range(num_epochs))
        feat_dim = len(next(iter(self.shared["pid2feat"].values())))
        for _ in range(num_steps):
            batch_idxs = tuple(i for i in next(batch_idxs_iter) if i is not None)
            batch_data = self.get_by_idxs(batch_idxs)
            pid2idx = {}
            for albumIds in batch_data["aid"]:
                for albumId in albumIds:
                    for pid in self.shared["albums"][albumId]["photo_ids"]:
                        if pid not in pid2idx:
                            pid2idx[pid] = len(pid2idx)
            image_feats = np.zeros((len(pid2idx), feat_dim), dtype="float32")
            for pid, i in pid2idx.items():
                image_feats[i] = self.shared["pid2feat"][pid]
            batch_data["pidx2feat"] = image_feats
            shared_batch_data = defaultdict(list)
            for albumIds in batch_data["aid"]:
                cols = defaultdict(list)
                for albumId in albumIds:
                    album = self.shared["albums"][albumId]
                    cols["album_title"].append(album["title"])
                    cols["album_title_c"].append(album["title_c"])
                    cols["album_description"].append(album["description"])
                    cols["album_description_c"].append(album["description_c"])
                    cols["where"].append(album["where"])
                    cols["where_c"].append(album["where_c"])
                    cols["when"].append(album["when"])
                    cols["when_c"].append(album["when_c"])
                    cols["photo_titles"].append(album["photo_titles"])
                    cols["photo_titles_c"].append(album["photo_titles_c"])
                    cols["photo_idxs"].append([pid2idx[pid] for pid in album["photo_ids"]])
                    cols["photo_ids"].append(album["photo_ids"])
                for k, v in cols.items():
                    shared_batch_data[k].append(v)
            batch_data.update(shared_batch_data)
            yield batch_idxs, Dataset(batch_data, self.datatype, shared=self.shared)


def getAnswers(yp, batch):
    """utils.py:247-259."""
    id2predanswers, id2realanswers = {}, {}
    for qid, yidxi, ypi in zip(batch[1].data["qid"], batch[1].data["yidx"], yp):
        id2predanswers[qid] = int(np.argmax(ypi))
        id2realanswers[qid] = yidxi
        assert yidxi < 4
        assert np.argmax(ypi) < 4
    return id2predanswers, id2realanswers


def getAnswers_yp(yp, batch):
    """utils.py:262-276."""
    pred, gt = getAnswers(yp, batch)
    id2yp = {qid: ypi for qid, ypi in zip(batch[1].data["qid"], yp)}
    return pred, gt, id2yp


def getEvalScore(pred, gt):
    """utils.py:279-287."""
    assert len(pred) == len(gt)
    assert len(pred) > 0
    return sum(1 for qid in pred if pred[qid] == gt[qid]) / float(len(pred))


def update_config(config, datasets, showMeta=False):
    """utils.py:294-396: scan the data for the maximum lengths, clamp them to the *_thres flags
    (training clamps everything; testing only description and photo-title length), set the vocab
    sizes."""
    names = ["max_num_albums", "max_num_photos", "max_sent_album_title_size", "max_sent_photo_title_size",
             "max_sent_des_size", "max_when_size", "max_where_size", "max_answer_size", "max_question_size",
             "max_word_size"]
    m = dict.fromkeys(names, 0)

    def wl(words):
        return max(len(w) for w in words)
    for dataset in datasets:
        for idx in dataset.valid_idxs:
            question = dataset.data["q"][idx]
            m["max_question_size"] = max(m["max_question_size"], len(question))
            m["max_word_size"] = max(m["max_word_size"], wl(question))
            for sent in dataset.data["cs"][idx] + [dataset.data["y"][idx]]:
                m["max_answer_size"] = max(m["max_answer_size"], len(sent))
                m["max_word_size"] = max(m["max_word_size"], wl(sent))
            albums = [dataset.shared["albums"][aid] for aid in dataset.data["aid"][idx]]
            m["max_num_albums"] = max(m["max_num_albums"], len(albums))
            for album in albums:
                m["max_num_photos"] = max(m["max_num_photos"], len(album["photo_ids"]))
                m["max_sent_album_title_size"] = max(m["max_sent_album_title_size"], len(album["title"]))
                for title in album["photo_titles"]:
                    if len(title) > 0:
                        m["max_sent_photo_title_size"] = max(m["max_sent_photo_title_size"], len(title))
                        m["max_word_size"] = max(m["max_word_size"], wl(title))
                if len(album["description"]) > 0:
                    m["max_sent_des_size"] = max(m["max_sent_des_size"], len(album["description"]))
                    m["max_word_size"] = max(m["max_word_size"], wl(album["description"]))
                m["max_when_size"] = max(m["max_when_size"], len(album["when"]))
                m["max_word_size"] = max(m["max_word_size"], wl(album["title"]), wl(album["when"]))
                if len(album["where"]) != 0:
                    m["max_word_size"] = max(m["max_word_size"], wl(album["where"]))
                    m["max_where_size"] = max(m["max_where_size"], len(album["where"]))
    for k, v in m.items():
        setattr(config, k, v)
    if showMeta:
        print("max meta:")
        print("\t" + " ,".join("%s:%s" % (k, getattr(config, k)) for k in names))
    if config.is_train:
        config.max_num_albums = min(config.max_num_albums, config.num_albums_thres)
        config.max_num_photos = min(config.max_num_photos, config.num_photos_thres)
        config.max_sent_album_title_size = min(config.max_sent_album_title_size, config.sent_album_title_size_thres)
        config.max_sent_photo_title_size = min(config.max_sent_photo_title_size, config.sent_photo_title_size_thres)
        config.max_sent_des_size = min(config.max_sent_des_size, config.sent_des_size_thres)
        config.max_when_size = min(config.max_when_size, config.sent_when_size_thres)
        config.max_where_size = min(config.max_where_size, config.sent_where_size_thres)
        config.max_answer_size = min(config.max_answer_size, config.answer_size_thres)
    else:
        config.max_sent_des_size = min(config.max_sent_des_size, config.sent_des_size_thres)
        config.max_sent_photo_title_size = min(config.max_sent_photo_title_size, config.sent_photo_title_size_thres)
    config.max_word_size = min(config.max_word_size, config.word_size_thres)
    config.char_vocab_size = len(datasets[0].shared["char2idx"])
    config.word_emb_size = len(next(iter(datasets[0].shared["word2vec"].values())))
    config.word_vocab_size = len(datasets[0].shared["word2idx"])


def build_feed_dict(model, batch, is_train=False, rng=None):
    """model_v2.py:1099-1565: a list-of-lists batch -> {placeholder name: numpy array}.

    Same semantics as the reference: padded sizes shrink to the batch maxima (:1126-1165, empty
    -> 1), truncation by ``break`` at the config maxima, word lookup tries the word as is /
    lower / capitalize / upper in the trainable vocabulary then in the GloVe vocabulary (offset by
    the trainable size) else UNK=1 (:1322-1336), char lookup else 1 (:1338-1343); in training the
    answer is inserted at a random slot (:1274-1283), in testing at ``yidx`` (:1298-1302).  The
    per-word Python work is done once per distinct word (a cache) rather than per occurrence."""
    config = model.config
    N = 2 if getattr(config, "showspecs", False) else config.batch_size
    data, shared = batch.data, batch.shared
    C = model.num_choice
    W = config.max_word_size

    def bmax(seqs, default=0):
        vals = [len(s) for s in seqs]
        v = max(vals + [default]) if vals else default
        return 1 if v == 0 else v

    JXA = min(config.max_sent_album_title_size, bmax(t for s in data["album_title"] for t in s))
    JXP = min(config.max_sent_photo_title_size, bmax(t for s in data["photo_titles"] for a in s for t in a))
    JD = min(config.max_sent_des_size, bmax(t for s in data["album_description"] for t in s))
    JG = min(config.max_where_size, bmax(t for s in data["where"] for t in s))
    JT = min(config.max_when_size, bmax(t for s in data["when"] for t in s))
    JI = min(config.max_num_photos, bmax(a for s in data["photo_ids"] for a in s))
    JQ = min(config.max_question_size, bmax(data["q"]))
    M = min(config.max_num_albums, bmax(data["album_title"]))
    JA = config.max_answer_size

    w2i, e2i, c2i = shared["word2idx"], shared["existing_word2idx"], shared["char2idx"]
    nw = len(w2i)
    wcache, ccache = {}, {}

    def get_word(word):
        r = wcache.get(word)
        if r is None:
            r = 1
            forms = (word, word.lower(), word.capitalize(), word.upper())
            for each in forms:
                if each in w2i:
                    r = w2i[each]
                    break
            else:
                for each in forms:
                    if each in e2i:
                        r = e2i[each] + nw
                        break
            wcache[word] = r
        return r

    def get_chars(cword):
        key = cword if isinstance(cword, str) else tuple(cword)
        r = ccache.get(key)
        if r is None:
            r = [c2i.get(ch, 1) for ch in list(cword)[:W]]
            ccache[key] = r
        return r

    def alloc(*shape):
        return (np.zeros(shape, dtype="int32"), np.zeros(shape + (W,), dtype="int32"), np.zeros(shape, dtype="bool"))

    at, at_c, at_mask = alloc(N, M, JXA)
    ad, ad_c, ad_mask = alloc(N, M, JD)
    when, when_c, when_mask = alloc(N, M, JT)
    where, where_c, where_mask = alloc(N, M, JG)
    pts, pts_c, pts_mask = alloc(N, M, JI, JXP)
    q, q_c, q_mask = alloc(N, JQ)
    choices, choices_c, choices_mask = alloc(N, C, JA)
    pis = np.zeros([N, M, JI], dtype="int32")
    pis_mask = np.zeros([N, M, JI], dtype="bool")

    def fill_album_site(words, chars, ids, cids, mask, J):
        for i, (ws, cs) in enumerate(zip(words, chars)):
            for j, (wj, cj) in enumerate(zip(ws[:M], cs[:M])):
                for k, (w, c) in enumerate(zip(wj[:J], cj[:J])):
                    ids[i, j, k] = get_word(w)
                    mask[i, j, k] = True
                    ch = get_chars(c)
                    cids[i, j, k, :len(ch)] = ch

    fill_album_site(data["album_title"], data["album_title_c"], at, at_c, at_mask, config.max_sent_album_title_size)
    fill_album_site(data["album_description"], data["album_description_c"], ad, ad_c, ad_mask, config.max_sent_des_size)
    fill_album_site(data["when"], data["when_c"], when, when_c, when_mask, config.max_when_size)
    fill_album_site(data["where"], data["where_c"], where, where_c, where_mask, config.max_where_size)
    for i, (pt, ptc) in enumerate(zip(data["photo_titles"], data["photo_titles_c"])):
        for j, (ptj, ptcj) in enumerate(zip(pt[:config.max_num_albums], ptc[:config.max_num_albums])):
            for k, (ptjk, ptcjk) in enumerate(zip(ptj[:config.max_num_photos], ptcj[:config.max_num_photos])):
                for l, (w, c) in enumerate(zip(ptjk[:config.max_sent_photo_title_size],
                                               ptcjk[:config.max_sent_photo_title_size])):
                    pts[i, j, k, l] = get_word(w)
                    pts_mask[i, j, k, l] = True
                    ch = get_chars(c)
                    pts_c[i, j, k, l, :len(ch)] = ch
    for i, pi in enumerate(data["photo_idxs"]):
        for j, pij in enumerate(pi[:config.max_num_albums]):
            for k, pijk in enumerate(pij[:config.max_num_photos]):
                assert isinstance(pijk, (int, np.integer))
                pis[i, j, k] = pijk
                pis_mask[i, j, k] = True
    # choices: insert the correct answer among the wrong ones
    Cs = [list(c) for c in data["cs"]]
    Ccs = [list(c) for c in data["ccs"]]
    y = np.zeros([N, C], dtype="bool")
    n_ex = len(data["q"])
    if is_train:
        gen = rng or np.random
        correctIndex = gen.choice(C, N)
        for i in range(len(data["y"])):
            y[i, correctIndex[i]] = True
            assert len(Cs[i]) == C - 1
            Cs[i].insert(correctIndex[i], data["y"][i])
            Ccs[i].insert(correctIndex[i], data["cy"][i])
    elif "y" in data and "cy" in data and "yidx" in data:
        for i in range(len(data["y"])):
            assert len(Cs[i]) == C - 1
            Cs[i].insert(data["yidx"][i], data["y"][i])
            Ccs[i].insert(data["yidx"][i], data["cy"][i])
            y[i, data["yidx"][i]] = True      # not fed by the reference at test time; harmless here
    for i, (ci, cci) in enumerate(zip(Cs, Ccs)):
        assert len(ci) == C
        for j, (cij, ccij) in enumerate(zip(ci, cci)):
            for k, (w, c) in enumerate(zip(cij[:JA], ccij[:JA])):
                choices[i, j, k] = get_word(w)
                choices_mask[i, j, k] = True
                ch = get_chars(c)
                choices_c[i, j, k, :len(ch)] = ch
    for i, (qi, cqi) in enumerate(zip(data["q"], data["cq"])):
        for j, (w, c) in enumerate(zip(qi, cqi)):     # the question is NOT truncated (:1538-1542)
            q[i, j] = get_word(w)
            q_mask[i, j] = True
            ch = get_chars(c)
            q_c[i, j, :len(ch)] = ch
    feed = dict(at=at, at_c=at_c, at_mask=at_mask, ad=ad, ad_c=ad_c, ad_mask=ad_mask, when=when, when_c=when_c,
                when_mask=when_mask, where=where, where_c=where_c, where_mask=where_mask, pts=pts, pts_c=pts_c,
                pts_mask=pts_mask, pis=pis, pis_mask=pis_mask, q=q, q_c=q_c, q_mask=q_mask, choices=choices,
                choices_c=choices_c, choices_mask=choices_mask, y=y,
                image_emb_mat=np.asarray(data["pidx2feat"], dtype="float32"),
                existing_emb_mat=np.asarray(shared["existing_emb_mat"], dtype="float32"))
    assert n_ex <= N
    return feed


def make_synthetic_raw(num_examples, num_albums=40, seed=0, feat_dim=2537, emb_dim=100, vocab=300):
    """A MemexQA-shaped dataset in the reference's ON-DISK format (what preprocess.py:252-275
    writes): ``data`` = per-question lists, ``shared`` = albums / pid2feat / counters / word2vec.
    Stands in for prepro_v1.1/{train,val,test}_{data,shared}.p, which cannot be downloaded here."""
    rnd = random.Random(seed)
    nrng = np.random.RandomState(seed)
    # word lengths 5..10: the char-CNN needs max_word_size >= its filter height 5
    words = ["tok%02d%s" % (i, "abcde"[: i % 6]) for i in range(vocab)]
    glove = set(words[: vocab * 2 // 3])            # the rest are "new" (trainable) words

    def sent(lo, hi):
        return [rnd.choice(words) for _ in range(rnd.randint(lo, hi))]

    def chars(s):
        return [list(w) for w in s]
    albums, pid2feat = {}, {}
    for a in range(num_albums):
        nphoto = rnd.randint(1, 10)
        pids = ["p%d_%d" % (a, i) for i in range(nphoto)]
        for pid in pids:
            f = nrng.randn(feat_dim).astype("float32")
            pid2feat[pid] = f / np.linalg.norm(f)
        title, desc, when, where = sent(1, 8), sent(0, 12), sent(1, 4), sent(0, 4)
        ptitles = [sent(0, 8) for _ in pids]
        albums["a%d" % a] = dict(title=title, title_c=chars(title), description=desc, description_c=chars(desc),
                                 when=when, when_c=chars(when), where=where, where_c=chars(where),
                                 photo_titles=ptitles, photo_titles_c=[chars(t) for t in ptitles], photo_ids=pids)
    data = defaultdict(list)
    for i in range(num_examples):
        qs, ys = sent(3, 20), sent(1, 5)
        cs = [sent(1, 5) for _ in range(3)]
        data["q"].append(qs); data["cq"].append(chars(qs))
        data["y"].append(ys); data["cy"].append(chars(ys))
        data["cs"].append(cs); data["ccs"].append([chars(c) for c in cs])
        data["yidx"].append(rnd.randint(0, 3))
        data["aid"].append(["a%d" % rnd.randrange(num_albums) for _ in range(rnd.randint(1, 5))])
        data["qid"].append("q%d" % i)
        data["idxs"].append(i)
    wordCounter, charCounter = defaultdict(int), defaultdict(int)
    for w in words:
        wordCounter[w] = 5
        for ch in w:
            charCounter[ch] += 20
    word2vec = {w: nrng.randn(emb_dim).astype("float32") * 0.4 for w in glove}
    shared = dict(albums=albums, pid2feat=pid2feat, wordCounter=dict(wordCounter), charCounter=dict(charCounter),
                  word2vec=word2vec)
    return dict(data), shared


def add_vocab(shared, config, model_shared=None):
    """main.py:232-270: build (or reuse) word2idx / char2idx, then existing_word2idx and the GloVe
    matrix.  Returns the side-car dict the reference pickles next to the model (shared.p)."""
    if model_shared is not None:
        shared.update(model_shared)
    else:
        new_words = [w for w, c in shared["wordCounter"].items() if c > config.word_count_thres and w not in shared["word2vec"]]
        shared["word2idx"] = {w: i + 2 for i, w in enumerate(new_words)}
        shared["char2idx"] = {c: i + 2 for i, c in enumerate(
            [c for c, n in shared["charCounter"].items() if n > config.char_count_thres])}
        shared["word2idx"]["<NULL>"] = 0
        shared["char2idx"]["<NULL>"] = 0
        shared["word2idx"]["<UNK>"] = 1
        shared["char2idx"]["<UNK>"] = 1
    shared["existing_word2idx"] = {w: i for i, w in enumerate(
        [w for w in sorted(shared["word2vec"].keys()) if w not in shared["word2idx"]])}
    idx2vec = {i: shared["word2vec"][w] for w, i in shared["existing_word2idx"].items()}
    shared["existing_emb_mat"] = np.array([idx2vec[i] for i in range(len(idx2vec))], dtype="float32")
    return {"word2idx": shared["word2idx"], "char2idx": shared["char2idx"]}
