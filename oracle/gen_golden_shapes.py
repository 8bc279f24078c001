"""TEST INFRASTRUCTURE ONLY — writes tests/golden/shape_parity.json from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python oracle/gen_golden_shapes.py

BASELINE-shaped decoder layers (Llama-2-7B / Llama-3-8B / Llama-2-13B widths, and the
llama3.2-1B shape of the reference's own test checkpoint, tests/tests_constants.py:9), two layers
deep so the fp32 CPU forward stays cheap.  Weights = oracle.random_state_dict(seed) (rebuildable
anywhere from the seed; a checksum proves it).  For one 1109-token synthetic sequence the
reference's own `forward` (self_speculation/llama_model_utils.py:155-209, imported through
oracle/ref_shim.py) produces teacher-forced logits; the fixture keeps, for the rows that the
parity tests look at (contexts 70 / 520 / 1100, i.e. 1 / 2 / 3 key groups per attention split,
up to 9 rows each), 128 sampled logits per row plus arg-max, max and log-sum-exp.
"""
from __future__ import annotations

import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import llama_oracle as orc  # noqa: E402
from oracle import ref_shim  # noqa: E402
from oracle.gen_golden import checksum  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

LLAMA3_SCALING = {"rope_type": "llama3", "factor": 32.0, "low_freq_factor": 1.0,
                  "high_freq_factor": 4.0, "original_max_position_embeddings": 8192}

# name: dims (vocab, hidden, inter, layers, heads, kv_heads, head_dim), theta, scaling, tied, seed
WIDTHS = {
    "w7b": dict(dims=(32000, 4096, 11008, 2, 32, 32, 128), theta=10000.0, scaling=None, tied=False, seed=11),
    "w8b": dict(dims=(128256, 4096, 14336, 2, 32, 8, 128), theta=500000.0, scaling=None, tied=False, seed=12),
    "w13b": dict(dims=(32000, 5120, 13824, 2, 40, 40, 128), theta=10000.0, scaling=None, tied=False, seed=13),
    "l32_1b": dict(dims=(128256, 2048, 8192, 2, 32, 8, 64), theta=500000.0, scaling=LLAMA3_SCALING,
                   tied=True, seed=14),
}
CONTEXTS = (70, 520, 1100)
MAX_ROWS = 9
SEQ_LEN = CONTEXTS[-1] + MAX_ROWS
N_SAMPLED = 128


def dims_of(spec) -> orc.LlamaDims:
    v, h, i, l, nh, nkv, hd = spec["dims"]
    return orc.LlamaDims(vocab=v, hidden=h, inter=i, layers=l, heads=nh, kv_heads=nkv, head_dim=hd,
                         rms_eps=1e-5, rope_theta=spec["theta"], rope_scaling=spec["scaling"])


def state_dict_of(spec):
    dims = dims_of(spec)
    sd = orc.random_state_dict(dims, spec["seed"])
    if spec["tied"]:
        sd["lm_head.weight"] = sd["model.embed_tokens.weight"]
    return dims, sd


def sequence_of(spec):
    g = torch.Generator().manual_seed(1000 + spec["seed"])
    return torch.randint(3, spec["dims"][0] - 1, (SEQ_LEN,), generator=g).tolist()


def sampled_columns(spec):
    g = torch.Generator().manual_seed(2000 + spec["seed"])
    return torch.randperm(spec["dims"][0], generator=g)[:N_SAMPLED].tolist()


def build_hf(dims: orc.LlamaDims, sd, tied: bool):
    from transformers import LlamaConfig, LlamaForCausalLM
    kw = dict(vocab_size=dims.vocab, hidden_size=dims.hidden, intermediate_size=dims.inter,
              num_hidden_layers=dims.layers, num_attention_heads=dims.heads,
              num_key_value_heads=dims.kv_heads, head_dim=dims.head_dim,
              max_position_embeddings=16384 if dims.rope_scaling else 4096,   # llama3 rule: > original 8192
              rms_norm_eps=dims.rms_eps, tie_word_embeddings=tied)
    rp = {"rope_type": "default", "rope_theta": dims.rope_theta}
    if dims.rope_scaling:
        rp = {**dims.rope_scaling, "rope_theta": dims.rope_theta}
    try:
        cfg = LlamaConfig(rope_parameters=rp, **kw)                    # transformers 5.x
    except TypeError:                                                   # transformers 4.x
        cfg = LlamaConfig(rope_theta=dims.rope_theta, rope_scaling=dims.rope_scaling, **kw)
    model = LlamaForCausalLM(cfg)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not [m for m in missing if "rotary" not in m], missing
    assert not unexpected, unexpected
    return model.eval()


def main() -> None:
    ref = ref_shim.load_reference()
    lmu = ref.llama_model_utils
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    cases = []
    for name, spec in WIDTHS.items():
        dims, sd = state_dict_of(spec)
        model = build_hf(dims, sd, spec["tied"])
        ids = sequence_of(spec)
        cols = sampled_columns(spec)
        with torch.inference_mode():
            out = lmu.forward(model, torch.tensor([ids]), None)
        logits = out.logits[0].float()                                    # [SEQ_LEN, V]
        rows = {}
        for ctx in CONTEXTS:
            # row j predicts the token after ids[: ctx + 1 + j]
            blk = logits[ctx: ctx + MAX_ROWS]
            rows[str(ctx)] = dict(
                argmax=[int(t) for t in blk.argmax(-1)],
                max=[float(x) for x in blk.max(-1).values],
                logsumexp=[float(x) for x in torch.logsumexp(blk.double(), -1)],
                sampled=[[float(x) for x in r[cols]] for r in blk])
        cases.append(dict(name=name, dims=list(spec["dims"]), rope_theta=spec["theta"],
                          rope_scaling=spec["scaling"], tied=spec["tied"], weight_seed=spec["seed"],
                          weights_checksum=checksum(sd), seq_len=SEQ_LEN, contexts=list(CONTEXTS),
                          sampled_columns=cols, rows=rows))
        print(f"{name}: logits {tuple(logits.shape)} |max| {float(logits.abs().max()):.3f}")
        del model, sd, out, logits
    with open(os.path.join(GOLDEN_DIR, "shape_parity.json"), "w") as f:
        json.dump(dict(generator="oracle/gen_golden_shapes.py",
                       reference="facebookresearch/LayerSkip self_speculation/llama_model_utils.py "
                                 "forward(), run unmodified under oracle/ref_shim.py",
                       torch=torch.__version__, cases=cases), f)
    print("wrote", os.path.join(GOLDEN_DIR, "shape_parity.json"))


if __name__ == "__main__":
    main()
