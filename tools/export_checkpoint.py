#!/usr/bin/env python
"""Write a synthetic Llama (weights.SyntheticLlama) as a Hugging Face checkpoint directory, so
the reference's own scripts (`generate.py --model <dir>`) and this engine (`--model <dir>`) can be
pointed at the very same weights.

    python tools/export_checkpoint.py synthetic:llama2-7b /data/syn7b --alpha 0.1 --exit_layer 8

Tensors are generated one at a time on `--device` (the per-tensor seeds make CPU and CUDA streams
different: export on the device class you benchmarked on)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawTextHelpFormatter)
    ap.add_argument("model", help="synthetic:<arch> (see layerskip_b200.weights.ARCHS)")
    ap.add_argument("out_dir")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--alpha", type=float, default=1.0)
    ap.add_argument("--exit_layer", type=int, default=-1, help="alpha damps layers >= exit_layer")
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--max_shard_gb", type=float, default=4.0)
    a = ap.parse_args(argv)
    from layerskip_b200.checkpoint import save_checkpoint
    from layerskip_b200.weights import ARCHS, SyntheticLlama
    if not a.model.startswith("synthetic:"):
        ap.error("model must be synthetic:<arch>")
    arch = ARCHS[a.model.split(":", 1)[1]]
    syn = SyntheticLlama(arch, seed=a.seed, alpha=a.alpha,
                         damp_from=a.exit_layer if a.exit_layer > 0 else None, device=a.device)
    stream = ((name, syn.tensor(name, shape)) for name, shape in syn.names())
    files = save_checkpoint(a.out_dir, arch, stream, max_shard_bytes=int(a.max_shard_gb * (1 << 30)))
    print(f"{a.out_dir}: {len(files)} shard(s), {arch.param_bytes() / 1e9:.2f} GB of bf16 weights")


if __name__ == "__main__":
    main()
