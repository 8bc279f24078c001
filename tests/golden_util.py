"""Shared helpers for tests that replay tests/golden/*.json (written by oracle/gen_golden.py)."""
import json
import os

import torch

from oracle import llama_oracle as orc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(GOLDEN_DIR, name)) as f:
        return json.load(f)


def dims_from_list(d):
    v, h, i, l, nh, nkv, hd = d
    return orc.LlamaDims(vocab=v, hidden=h, inter=i, layers=l, heads=nh, kv_heads=nkv,
                         head_dim=hd, rms_eps=1e-5, rope_theta=10000.0)


def checksum(sd) -> str:
    acc = 0.0
    for k in sorted(sd):
        t = sd[k].to(torch.float64)
        acc += float((t.abs().sum() + (t * torch.arange(1, t.numel() + 1, dtype=torch.float64)
                                       .view(t.shape) % 7).sum()))
    return f"{acc:.6f}"


def checksum_matches(sd, text: str) -> bool:
    """The checksum is a float64 sum over ~1e9 terms: thread count changes the reduction order,
    so compare to 1e-12 relative instead of digit for digit."""
    a, b = float(checksum(sd)), float(text)
    return abs(a - b) <= 1e-12 * max(abs(a), abs(b), 1.0) + 1e-5


def state_dict_for(case, alpha_key=True):
    dims = dims_from_list(case["dims"])
    sd = orc.random_state_dict(dims, case["weight_seed"],
                               case.get("damp_from") if alpha_key else None,
                               case.get("alpha", 1.0) if alpha_key else 1.0)
    assert checksum_matches(sd, case["weights_checksum"]), (
        "seeded weights differ from the ones the golden file was generated with "
        "(torch CPU RNG changed?) — regenerate with oracle/gen_golden.py")
    return dims, sd


def spec_cases(greedy=None):
    cases = load("spec_traces.json")["cases"]
    if greedy is None:
        return cases
    return [c for c in cases if (not c["cfg"]["sample"]) == greedy]
