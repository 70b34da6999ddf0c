"""Short driver for profilers: W warm-up steps + K timed training steps at the C2 shape
(same model/config as bench.py), or the attention microbenchmark (--attn).

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python tools/profile_step.py --steps 1 --warmup 1
    python tools/profile_step.py --torch-profiler      # kernel table via CUPTI, no ncu
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--attn", action="store_true")
    ap.add_argument("--torch-profiler", action="store_true")
    ap.add_argument("--batch", type=int, default=32)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    if args.attn:
        import bench
        print(bench.attention_roofline(dev, iters=args.steps))
        return
    from fvta_memexqa_b200.config import fvta_config
    from fvta_memexqa_b200.data import SyntheticDataset
    from fvta_memexqa_b200.model import get_model
    from fvta_memexqa_b200.trainer import Trainer
    cfg = fvta_config(batch_size=args.batch, seed=1234)
    model = get_model(cfg)
    trainer = Trainer(model, cfg)
    data = SyntheticDataset(cfg, num_examples=args.batch, seed=1234, pin=True)
    batch = next(iter(data.get_batches(args.batch, num_steps=1, shuffle=False)))
    dfeed = model.to_device(model.get_feed_dict(batch[1], is_train=True))
    for _ in range(args.warmup):
        trainer._device_step(dfeed)
    torch.cuda.synchronize()
    if args.torch_profiler:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for _ in range(args.steps):
                trainer._device_step(dfeed)
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=60))
    else:
        for _ in range(args.steps):
            trainer._device_step(dfeed)
        torch.cuda.synchronize()


if __name__ == "__main__":
    main()
