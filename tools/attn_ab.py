"""A/B of the two forward kernels (tensor-core vs FFMA) on one shape: max difference of every
output and the time of each.  FVTA_FORCE_FFMA is read once per process, so the FFMA leg runs in
a child process."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def leg():
    import torch
    import bench
    dev = torch.device("cuda:0")
    r = bench.attention_roofline(dev, iters=10)
    r2 = bench.attention_roofline(dev, iters=10, with_logits=True)
    print(json.dumps({"launch_ms": r["launch_ms"], "frac": r["frac"], "with_logits_ms": r2["launch_ms"],
                      "with_logits_frac": r2["frac"]}))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "leg":
        leg()
    else:
        for name, env in (("tensor-core", {}), ("ffma", {"FVTA_FORCE_FFMA": "1"})):
            e = dict(os.environ, **env)
            out = subprocess.run([sys.executable, __file__, "leg"], env=e, capture_output=True, text=True)
            print(name, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-500:])
