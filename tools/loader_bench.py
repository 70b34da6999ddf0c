"""Throughput of the input pipeline alone (no device work): batches/s of pipeline.BatchLoader for several
worker counts, and the cost of the parent-side hand-over."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from fvta_memexqa_b200.config import fvta_config  # noqa: E402
from fvta_memexqa_b200.data import Dataset, make_synthetic_raw, add_vocab  # noqa: E402
from fvta_memexqa_b200.pipeline import TokenizedStore, BatchLoader  # noqa: E402


class _P:
    touch_rows, touch_len = {}, 0


class _M:
    def __init__(self, cfg):
        self.config, self.params = cfg, _P()


if __name__ == "__main__":
    cfg = fvta_config(batch_size=32, seed=1)
    data, shared = make_synthetic_raw(4096, num_albums=1200, seed=4242, feat_dim=cfg.image_feat_dim,
                                      emb_dim=cfg.word_emb_size, vocab=3 * (cfg.word_vocab_size - 2))
    add_vocab(shared, cfg)
    store = TokenizedStore(Dataset(data, "train", shared=shared), cfg)
    print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
    for workers in (0, 2, 4, 8, 14):
        n = 60 if workers == 0 else 400
        loader = BatchLoader(store, _M(cfg), 32, num_steps=n + 20, workers=workers, depth=max(4, 2 * workers))
        loader.pin_caps()
        loader.start()
        it = iter(loader)
        for _ in range(20):
            next(it)
        t = time.perf_counter()
        for _ in range(n):
            b = next(it)
        dt = time.perf_counter() - t
        loader.close()
        print("workers %2d: %.3f ms per batch (%d B)" % (workers, 1e3 * dt / n, b[1].step_input(None, True)["blob"].numel()))
    blob = torch.empty(774144, dtype=torch.uint8)
    if torch.cuda.is_available():
        t = time.perf_counter()
        for _ in range(100):
            p = blob.pin_memory()
        print("pin_memory() of the blob: %.3f ms" % (1e3 * (time.perf_counter() - t) / 100))
