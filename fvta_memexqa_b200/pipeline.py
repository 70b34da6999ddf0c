"""Vectorised input pipeline: the reference's ``get_feed_dict`` without its per-word loops.

The reference turns every batch into placeholders with five nested Python loops over albums, photos,
sentences, words and characters, looking each word up in two dictionaries (model_v2.py:1099-1565) --
tens of milliseconds per batch of 32, forty times the device step on a B200.  Nothing in that work
depends on the batch: an album's title is tokenised the same way every time it is drawn.  So

* :class:`TokenizedStore` tokenises every album and every question ONCE (same lookup rule,
  model_v2.py:1322-1343) into integer arrays padded to the config maxima, and keeps the photo
  features as one table;
* :meth:`TokenizedStore.feed` assembles a batch's placeholders with a handful of ``numpy`` gathers
  (album rows by album index), reproducing ``build_feed_dict`` bit for bit (tests/test_pipeline.py);
* with ``resident_photos`` the photo-feature table lives on the device for the whole run (14k photos
  x 2537 floats = 140 MB of 180 GB) and ``pis`` carries global photo rows: the image transform reads
  them in place (fvta_gemm_gather), and the per-step host->device copy drops from ~5 MB to the token
  ids alone;
* :class:`BatchLoader` runs that on worker processes a few batches ahead of the training loop and hands
  over packed, pinned step inputs (packing.host_step_input), so ``Trainer.step`` never waits for Python.
"""
import copy
import math
import random
import threading

import numpy as np
import torch


def _lookup_tables(shared):
    w2i, e2i, c2i = shared["word2idx"], shared["existing_word2idx"], shared["char2idx"]
    nw = len(w2i)
    cache = {}

    def get_word(word):                       # model_v2.py:1322-1336
        r = cache.get(word)
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
            cache[word] = r
        return r

    def get_chars(cword, W):                  # model_v2.py:1338-1343
        return [c2i.get(ch, 1) for ch in list(cword)[:W]]

    return get_word, get_chars


class TokenizedStore:
    """Integer arrays of a reference-format dataset (``data`` + ``shared``, utils.py:31-88)."""

    def __init__(self, dataset, config, num_choice=4):
        self.config = config
        self.C = num_choice
        data, shared = dataset.data, dataset.shared
        self.data = data
        self.valid_idxs = list(dataset.valid_idxs)
        self.num_examples = len(self.valid_idxs)
        W = self.W = config.max_word_size
        get_word, get_chars = _lookup_tables(shared)
        cm = config
        # truncation by the config maxima happens at tokenisation time (the reference `break`s there)
        self.JXA, self.JD = cm.max_sent_album_title_size, cm.max_sent_des_size
        self.JT, self.JG = cm.max_when_size, cm.max_where_size
        self.JI, self.JXP = cm.max_num_photos, cm.max_sent_photo_title_size
        self.M, self.JA = cm.max_num_albums, cm.max_answer_size

        def sent(words, chars, ids, cids, J):
            n = min(len(words), J)
            for k in range(n):
                ids[k] = get_word(words[k])
                ch = get_chars(chars[k], W)
                cids[k, :len(ch)] = ch
            return n

        # ---- photos: one feature table, rows in sorted-pid order
        pids = sorted(shared["pid2feat"].keys())
        self.pid_row = {p: i for i, p in enumerate(pids)}
        self.photo_feats = np.stack([np.asarray(shared["pid2feat"][p], dtype="float32") for p in pids]) if pids \
            else np.zeros((0, cm.image_feat_dim), "float32")
        # ---- albums
        aids = sorted(shared["albums"].keys())
        self.album_index = {a: i for i, a in enumerate(aids)}
        A = len(aids)
        z = lambda *s: np.zeros(s, dtype="int32")      # noqa: E731
        self.at, self.at_c, self.at_len = z(A, self.JXA), z(A, self.JXA, W), z(A)
        self.ad, self.ad_c, self.ad_len = z(A, self.JD), z(A, self.JD, W), z(A)
        self.when, self.when_c, self.when_len = z(A, self.JT), z(A, self.JT, W), z(A)
        self.where, self.where_c, self.where_len = z(A, self.JG), z(A, self.JG, W), z(A)
        self.pts, self.pts_c, self.pts_len = z(A, self.JI, self.JXP), z(A, self.JI, self.JXP, W), z(A, self.JI)
        self.n_photos = z(A)                           # min(#photos, JI)
        self.album_photo = z(A, self.JI)               # global photo rows of the first JI photos
        self.album_photo_full = []                     # ALL photos of the album (pidx2feat order, utils.py:127-143)
        # raw lengths, before the config maxima: the reference's per-batch padded sizes look at them
        self.raw = {k: z(A) for k in ("at", "ad", "when", "where", "photos", "pts")}
        for a, aid in enumerate(aids):
            al = shared["albums"][aid]
            self.at_len[a] = sent(al["title"], al["title_c"], self.at[a], self.at_c[a], self.JXA)
            self.ad_len[a] = sent(al["description"], al["description_c"], self.ad[a], self.ad_c[a], self.JD)
            self.when_len[a] = sent(al["when"], al["when_c"], self.when[a], self.when_c[a], self.JT)
            self.where_len[a] = sent(al["where"], al["where_c"], self.where[a], self.where_c[a], self.JG)
            rows = [self.pid_row[p] for p in al["photo_ids"]]
            self.album_photo_full.append(np.asarray(rows, dtype="int64"))
            n = min(len(rows), self.JI)
            self.n_photos[a] = n
            self.album_photo[a, :n] = rows[:n]
            for k in range(min(len(al["photo_titles"]), self.JI)):
                self.pts_len[a, k] = sent(al["photo_titles"][k], al["photo_titles_c"][k], self.pts[a, k],
                                          self.pts_c[a, k], self.JXP)
            self.raw["at"][a] = len(al["title"])
            self.raw["ad"][a] = len(al["description"])
            self.raw["when"][a] = len(al["when"])
            self.raw["where"][a] = len(al["where"])
            self.raw["photos"][a] = len(al["photo_ids"])
            self.raw["pts"][a] = max([len(t) for t in al["photo_titles"]] + [0])
        # ---- questions (not truncated, model_v2.py:1538-1542), choices, album lists
        Q = len(data["q"])
        JQd = max([len(q) for q in data["q"]] + [1, cm.max_question_size])
        Md = max([len(a) for a in data["aid"]] + [1, cm.max_num_albums])
        self.q, self.q_c, self.q_len = z(Q, JQd), z(Q, JQd, W), z(Q)
        self.wrong, self.wrong_c, self.wrong_len = z(Q, self.C - 1, self.JA), z(Q, self.C - 1, self.JA, W), z(Q, self.C - 1)
        self.ans, self.ans_c, self.ans_len = z(Q, self.JA), z(Q, self.JA, W), z(Q)
        self.qa_albums = np.full((Q, Md), -1, dtype="int64")
        self.n_albums = z(Q)
        self.yidx = np.asarray(data.get("yidx", [0] * Q), dtype="int64")
        for i in range(Q):
            self.q_len[i] = sent(data["q"][i], data["cq"][i], self.q[i], self.q_c[i], JQd)
            self.ans_len[i] = sent(data["y"][i], data["cy"][i], self.ans[i], self.ans_c[i], self.JA)
            assert len(data["cs"][i]) == self.C - 1
            for c in range(self.C - 1):
                self.wrong_len[i, c] = sent(data["cs"][i][c], data["ccs"][i][c], self.wrong[i, c], self.wrong_c[i, c], self.JA)
            al = [self.album_index[a] for a in data["aid"][i]]
            self.qa_albums[i, :len(al)] = al
            self.n_albums[i] = len(al)
        self.existing_emb_mat = np.asarray(shared["existing_emb_mat"], dtype="float32")

    def token_counts(self):
        """(text tokens, image tokens) each question contributes to a packed batch (packing.build_pack):
        its own words, its choices, and the words / photos of its first M albums."""
        alb = self.qa_albums[:, :self.M]
        have = alb >= 0
        safe = np.where(have, alb, 0)
        per_album = (self.at_len + self.ad_len + self.when_len + self.where_len + self.pts_len.sum(1))
        text = self.q_len + self.ans_len + self.wrong_len.sum(1) + (per_album[safe] * have).sum(1)
        img = (self.n_photos[safe] * have).sum(1)
        return text.astype("int64"), img.astype("int64")

    # ------------------------------------------------------------------------------------------
    def feed(self, idxs, is_train=False, rng=None, pin_shapes=False, resident_photos=False):
        """Placeholders of the batch ``idxs`` (indices into ``data``), as ``build_feed_dict`` returns them.

        ``pin_shapes``: pad to the config maxima instead of shrinking to the batch maxima (the reference
        shrinks, model_v2.py:1126-1165; fixed shapes are what a captured CUDA graph replays).
        ``resident_photos``: ``pis`` holds rows of the store's full photo table and no per-batch
        ``image_emb_mat`` is produced (the model must hold the table: Model.set_photo_table)."""
        cm = self.config
        N = 2 if getattr(cm, "showspecs", False) else cm.batch_size
        idxs = np.asarray(list(idxs), dtype="int64")
        n = len(idxs)
        assert n <= N
        C, W = self.C, self.W
        alb_all = self.qa_albums[idxs]                                  # [n, Md], -1 = none
        have = alb_all >= 0
        safe = np.where(have, alb_all, 0)

        def bmax(raw, cap):
            """min(config maximum, batch maximum), an empty batch maximum counting as 1."""
            if pin_shapes:
                return cap
            v = int(raw.max()) if raw.size else 0
            return min(cap, 1 if v == 0 else v)

        def raw_of(key):
            return np.where(have, self.raw[key][safe], 0)
        JXA, JD = bmax(raw_of("at"), self.JXA), bmax(raw_of("ad"), self.JD)
        JT, JG = bmax(raw_of("when"), self.JT), bmax(raw_of("where"), self.JG)
        JI, JXP = bmax(raw_of("photos"), self.JI), bmax(raw_of("pts"), self.JXP)
        M = bmax(self.n_albums[idxs], self.M)
        JQ = cm.max_question_size if pin_shapes else min(cm.max_question_size, max(1, int(self.q_len[idxs].max()) if n else 1))
        JA = self.JA
        alb, have = safe[:, :M], have[:, :M]
        if alb.shape[1] < M:                                            # config maximum above the data's
            pad = M - alb.shape[1]
            alb = np.pad(alb, ((0, 0), (0, pad)))
            have = np.pad(have, ((0, 0), (0, pad)))

        def rows(x, n_extra_dims):
            """[n, ...] -> [N, ...] (all-zero rows behind a short batch, utils.py:117)."""
            if n == N:
                return x
            out = np.zeros((N,) + x.shape[1:], dtype=x.dtype)
            out[:n] = x
            return out

        def site(ids, cids, lens, J):
            hv = have[..., None]
            i = ids[alb][:, :, :J] * hv
            c = cids[alb][:, :, :J] * hv[..., None]
            ln = np.minimum(lens[alb], J) * have
            m = np.arange(J)[None, None, :] < ln[..., None]
            return rows(i.astype("int32"), 0), rows(c.astype("int32"), 0), rows(m, 0)

        f = {}
        f["at"], f["at_c"], f["at_mask"] = site(self.at, self.at_c, self.at_len, JXA)
        f["ad"], f["ad_c"], f["ad_mask"] = site(self.ad, self.ad_c, self.ad_len, JD)
        f["when"], f["when_c"], f["when_mask"] = site(self.when, self.when_c, self.when_len, JT)
        f["where"], f["where_c"], f["where_mask"] = site(self.where, self.where_c, self.where_len, JG)
        nph = np.minimum(self.n_photos[alb], JI) * have                 # [n, M]
        pis_mask = np.arange(JI)[None, None, :] < nph[..., None]
        pl = np.minimum(self.pts_len[alb][:, :, :JI], JXP) * pis_mask
        f["pts"] = rows((self.pts[alb][:, :, :JI, :JXP] * pis_mask[..., None]).astype("int32"), 0)
        f["pts_c"] = rows((self.pts_c[alb][:, :, :JI, :JXP] * pis_mask[..., None, None]).astype("int32"), 0)
        f["pts_mask"] = rows(np.arange(JXP)[None, None, None, :] < pl[..., None], 0)
        glob = self.album_photo[alb][:, :, :JI]                         # global photo rows
        if resident_photos:
            f["pis"] = rows((glob * pis_mask).astype("int32"), 0)
        else:
            # the batch's own photo table, rows in order of first appearance over ALL photos of ALL albums
            # of the batch (utils.py:127-143 -- before any truncation)
            seq = [self.album_photo_full[a] for i in range(n) for a in alb_all[i][alb_all[i] >= 0]]
            flat = np.concatenate(seq) if seq else np.zeros(0, "int64")
            uniq, first = np.unique(flat, return_index=True)
            order = uniq[np.argsort(first, kind="stable")]
            local = np.zeros(max(len(self.photo_feats), 1), dtype="int64")
            local[order] = np.arange(len(order))
            f["pis"] = rows((local[glob] * pis_mask).astype("int32"), 0)
            f["image_emb_mat"] = self.photo_feats[order]
        f["pis_mask"] = rows(pis_mask, 0)
        # ---- question
        ql = np.minimum(self.q_len[idxs], JQ)
        f["q"] = rows(self.q[idxs][:, :JQ].copy(), 0)
        f["q_c"] = rows(self.q_c[idxs][:, :JQ].copy(), 0)
        if f["q"].shape[1] < JQ:                                        # pinned JQ above the data's longest
            pad = JQ - f["q"].shape[1]
            f["q"] = np.pad(f["q"], ((0, 0), (0, pad)))
            f["q_c"] = np.pad(f["q_c"], ((0, 0), (0, pad), (0, 0)))
        f["q_mask"] = rows(np.arange(JQ)[None, :] < ql[:, None], 0)
        # ---- choices: the answer goes to a random slot in training (:1274-1283), to yidx in testing
        y = np.zeros((N, C), dtype="bool")
        if is_train:
            gen = rng or np.random
            slot = np.asarray(gen.choice(C, N))[:n]
        else:
            slot = self.yidx[idxs]
        y[np.arange(n), slot] = True
        cpos = np.arange(C)[None, :]
        src = np.clip(cpos - (cpos > slot[:, None]), 0, C - 2)          # which wrong answer sits in slot c
        is_ans = cpos == slot[:, None]
        qi = idxs[:, None]
        ch = np.where(is_ans[..., None], self.ans[idxs][:, None, :], self.wrong[qi, src])
        ch_c = np.where(is_ans[..., None, None], self.ans_c[idxs][:, None], self.wrong_c[qi, src])
        ch_l = np.where(is_ans, self.ans_len[idxs][:, None], self.wrong_len[qi, src])
        f["choices"] = rows(ch.astype("int32"), 0)
        f["choices_c"] = rows(ch_c.astype("int32"), 0)
        f["choices_mask"] = rows(np.arange(JA)[None, None, :] < ch_l[..., None], 0)
        f["y"] = y
        f["existing_emb_mat"] = self.existing_emb_mat
        return f


class HostBatch:
    """A batch whose packed, pinned step input already exists (what ``Trainer.step`` asks a batch for)."""

    def __init__(self, hsi, idxs, data=None, shared=None):
        self._hsi = hsi
        self.num_examples = len(idxs)
        self.data = data or {}
        self.shared = shared or {}

    def step_input(self, model, is_train):
        return self._hsi


# ---- worker-process side (module level: the workers are spawned and import this module) -------------
_W = {}


def _worker_init(store, pack, ring, slot_bytes):
    torch.set_num_threads(1)
    _W["store"], _W["pack"], _W["ring"], _W["slot_bytes"] = store, pack, ring, slot_bytes


def _worker_build(task):
    step, idxs, seed, is_train, pin_shapes, resident, slot = task
    from .packing import host_step_input
    store, pack = _W["store"], _W["pack"]
    rng = np.random.RandomState((seed * 1000003 + step) & 0x7fffffff)
    feed = store.feed(idxs, is_train=is_train, rng=rng, pin_shapes=pin_shapes, resident_photos=resident)
    feed["is_train"] = is_train
    hsi = host_step_input(feed, store.config, pack["min_text_cap"], pack["min_img_cap"], pack["touch_rows"],
                          pack["touch_len"], pin=False)
    blob = hsi["blob"]
    ring = _W.get("ring")
    if ring is None or slot < 0:
        return idxs, blob, hsi["index"], hsi["meta"]
    sb = _W["slot_bytes"]
    if blob.numel() > sb:
        raise ValueError("step input of %d bytes does not fit the %d-byte ring slot (BatchLoader.pin_caps)" % (blob.numel(), sb))
    ring[slot * sb: slot * sb + blob.numel()].copy_(blob)      # shared (and, in the parent, pinned) memory
    return idxs, blob.numel(), hsi["index"], hsi["meta"]


class BatchLoader:
    """``Dataset.get_batches`` (utils.py:89-198: same epoch structure and shuffling) producing packed,
    pinned step inputs ``depth`` batches ahead of the consumer, in order.

    The per-batch host work (numpy gathers + packing: ~4 ms for a batch of 32) is seven device steps on a
    B200 and mostly interpreter time, so threads do not scale it: ``workers`` PROCESSES (spawned -- the
    parent holds a CUDA context) each build whole batches and write them into a ring of shared-memory
    slots that the parent has page-locked (cudaHostRegister): the device copies a batch straight out of
    the slot its worker filled, and the parent touches no payload bytes.  A slot is reused ``depth + 3``
    batches later, after the step that consumed it has finished (Trainer.step reads the loss back).
    ``workers=0`` builds inline (tests, debugging)."""

    def __init__(self, store, model, batch_size, num_steps, shuffle=True, seed=0, is_train=True, workers=8,
                 depth=16, resident_photos=True, pin_shapes=True):
        self.store, self.model = store, model
        self.batch_size, self.num_steps = batch_size, num_steps
        self.shuffle, self.seed, self.is_train = shuffle, seed, is_train
        self.workers, self.depth = workers, max(depth, 2)
        self.resident_photos, self.pin_shapes = resident_photos, pin_shapes
        self.pack = dict(min_text_cap=getattr(model, "pin_text_cap", 0), min_img_cap=getattr(model, "pin_img_cap", 0),
                         touch_rows=model.params.touch_rows, touch_len=model.params.touch_len)
        self._pool = None
        self._ring = None
        self._registered = False

    def _tasks(self):
        s = self.store
        per_epoch = int(math.ceil(s.num_examples / float(self.batch_size)))
        idxs = list(s.valid_idxs)
        if self.shuffle:
            idxs = random.Random(self.seed).sample(idxs, len(idxs))     # same order every epoch (utils.py:104)
        nslots = self.depth + 3
        for step in range(self.num_steps):
            b = step % per_epoch
            yield (step, tuple(idxs[b * self.batch_size:(b + 1) * self.batch_size]), self.seed, self.is_train,
                   self.pin_shapes, self.resident_photos, (step % nslots) if self.pin_shapes else -1)

    def pin_caps(self, bucket=512):
        """Pin the model's packed-token capacities to the largest batch this loader will produce, so that
        every batch has the same step-input layout (one captured step graph).  Returns the capacities."""
        text, img = self.store.token_counts()
        t = max(int(text[list(task[1])].sum()) for task in self._tasks())
        i = max(int(img[list(task[1])].sum()) for task in self._tasks())
        caps = (max(bucket, (t + bucket - 1) // bucket * bucket), max(64, (i + 63) // 64 * 64))
        self.model.pin_text_cap = max(getattr(self.model, "pin_text_cap", 0), caps[0])
        self.model.pin_img_cap = max(getattr(self.model, "pin_img_cap", 0), caps[1])
        self.pack["min_text_cap"], self.pack["min_img_cap"] = self.model.pin_text_cap, self.model.pin_img_cap
        return self.model.pin_text_cap, self.model.pin_img_cap

    def start(self):
        """Spawn the workers now (they import torch: seconds) instead of on the first batch."""
        if self.workers > 0 and self._pool is None:
            import torch.multiprocessing as mp
            store = self.store
            if self.resident_photos:                 # the workers never touch the features
                store = copy.copy(store)
                store.photo_feats = store.photo_feats[:0]
                store.data = None
            slot_bytes = 0
            if self.pin_shapes:
                # with pinned shapes and capacities every batch has the size of the first one
                _worker_init(self.store, self.pack, None, 0)
                first = _worker_build(next(iter(self._tasks())))
                slot_bytes = (first[1].numel() + 4095) // 4096 * 4096
                self._ring = torch.empty((self.depth + 3) * slot_bytes, dtype=torch.uint8).share_memory_()
                if torch.cuda.is_available():
                    rc = torch.cuda.cudart().cudaHostRegister(self._ring.data_ptr(), self._ring.numel(), 0)
                    self._registered = int(rc) == 0
            self._slot_bytes = slot_bytes
            self._pool = mp.get_context("spawn").Pool(self.workers, _worker_init,
                                                      (store, self.pack, self._ring, slot_bytes))
        return self

    def close(self):
        if self._pool is not None:
            self._pool.terminate()
            self._pool.join()
            self._pool = None
        if self._registered:
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaHostUnregister(self._ring.data_ptr())
            self._registered = False
        self._ring = None

    def _finish(self, task, res):
        idxs, payload, index, meta = res
        if isinstance(payload, int):                 # the worker wrote the ring slot
            o = task[6] * self._slot_bytes
            blob = self._ring[o:o + payload]
        else:
            blob = payload
            if torch.cuda.is_available():
                blob = blob.pin_memory()             # caching host allocator: reuse is ordered after the H2D copies
        hsi = {"blob": blob, "index": index, "meta": meta, "host": {"is_train": self.is_train}}
        return idxs, HostBatch(hsi, idxs)

    def __iter__(self):
        if self.workers <= 0:
            _worker_init(self.store, self.pack, None, 0)
            for task in self._tasks():
                yield self._finish(task, _worker_build(task))
            return
        self.start()
        ahead = threading.Semaphore(self.depth)      # bounds how far imap's feeder thread runs ahead
        issued = []

        def gated():
            for task in self._tasks():
                ahead.acquire()
                issued.append(task)
                yield task
        for k, res in enumerate(self._pool.imap(_worker_build, gated())):
            ahead.release()
            yield self._finish(issued[k], res)
            issued[k] = None


def resident_photo_table(store, device):
    """The store's photo features on the device, for Model.set_photo_table."""
    return torch.from_numpy(store.photo_feats).to(device)
