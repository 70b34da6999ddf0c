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
