"""Tiny driver for ncu: a 7B-shaped engine, a short prompt, a few speculation rounds.
   python tools/profile_round.py [arch] [rounds] [ctx]

Cost of a kernel class INSIDE the replayed graph (what eager per-class timing cannot give):
   for c in "" qkv attn o gate_up down lm_head; do LSK_ABLATE=$c python tools/profile_round.py; done
and subtract the round times (LSK_ABLATE skips the named classes; outputs are garbage)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from layerskip_b200.engine import Engine
from layerskip_b200.weights import ARCHS, SyntheticLlama

arch = ARCHS[sys.argv[1] if len(sys.argv) > 1 else "llama2-7b"]
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 400
E, D = 8, 6
eng = Engine(arch, max_ctx=1024, use_pdl=not os.environ.get('LSK_NO_PDL'), use_graph=not os.environ.get('LSK_NO_GRAPH'))
eng.load_model(SyntheticLlama(arch, seed=0))
g = torch.Generator().manual_seed(1234)
prompt = torch.randint(3, arch.vocab - 1, (ctx,), generator=g).tolist()
eng.begin(exit_layer=E, max_steps=512, eos_token_ids=[arch.vocab - 1])
eng.prefill(prompt)
for _ in range(2):
    eng.round(D)
tot = 0.0
nbytes = 0.0
for _ in range(rounds):
    nbytes += eng.round_bytes(D, eng.kv_len)
    r = eng.round(D)
    tot += eng.last_device_ms
print(f"rounds={rounds} avg_ms={tot / rounds:.3f} kv_len={eng.kv_len} -> {nbytes / (tot * 1e-3) / 1e9:.0f} GB/s "
      f"env={ {k: v for k, v in os.environ.items() if k.startswith('LSK_')} }", flush=True)
if os.environ.get("LSK_PROFILE_CLASSES"):
    acc = {}
    for _ in range(3):
        _r, ms, cnt, _t = eng.profile_round(D)
        for k in ms:
            acc[k] = acc.get(k, 0.0) + ms[k] / 3
    print("  per-class ms/round (eager, isolated):", {k: round(v, 3) for k, v in acc.items()}, flush=True)
eng.close()
