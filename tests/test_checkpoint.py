"""Checkpoint ingest: HF directory format in, (role, layer, bf16) stream
out — no GPU needed.  Interop is checked against the installed transformers in both directions."""
import json
import os

import pytest
import torch

from layerskip_b200 import _lib
from layerskip_b200.checkpoint import (CheckpointError, CheckpointLlama, arch_from_config_json,
                                       config_json_of, expected_shapes, save_checkpoint)
from layerskip_b200.weights import ARCHS, LlamaArch, classify

ARCH = LlamaArch(vocab=96, hidden=64, inter=160, layers=3, heads=2, kv_heads=1, head_dim=32,
                 rms_eps=1e-6, rope_theta=50000.0)


def _random_tensors(arch, seed=0, dtype=torch.bfloat16, tied=False):
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in expected_shapes(arch, tied).items():
        out[name] = (torch.randn(shape, generator=g) * 0.05).to(dtype)
    return out


def _same(a, b):
    assert a.keys() == b.keys()
    for k in a:
        assert a[k].dtype == b[k].dtype and torch.equal(a[k], b[k]), k


def test_sharded_roundtrip_and_index(tmp_path):
    sd = _random_tensors(ARCH)
    files = save_checkpoint(str(tmp_path), ARCH, sd.items(), max_shard_bytes=40_000)
    assert len(files) > 2 and all(f.endswith(".safetensors") for f in files)
    index = json.load(open(tmp_path / "model.safetensors.index.json"))
    assert set(index["weight_map"]) == set(sd)
    assert index["metadata"]["total_size"] == sum(t.numel() * 2 for t in sd.values())
    ck = CheckpointLlama(str(tmp_path), device="cpu")
    assert ck.arch == ARCH and not ck.tied
    _same(ck.state_dict(), sd)
    # every shard is opened exactly once
    order = [f for f, _ in ck.plan()]
    assert [f for i, f in enumerate(order) if i == 0 or order[i - 1] != f] == sorted(set(order))


def test_single_file_without_index(tmp_path):
    sd = _random_tensors(ARCH, seed=1)
    assert save_checkpoint(str(tmp_path), ARCH, sd.items()) == ["model.safetensors"]
    assert not (tmp_path / "model.safetensors.index.json").exists()
    _same(CheckpointLlama(str(tmp_path), device="cpu").state_dict(), sd)


def test_iter_weights_roles_dtype_and_fp16_conversion(tmp_path):
    sd = _random_tensors(ARCH, seed=2, dtype=torch.float16)
    save_checkpoint(str(tmp_path), ARCH, sd.items(), max_shard_bytes=100_000)
    ck = CheckpointLlama(str(tmp_path), device="cpu")
    seen = {}
    for role, layer, t in ck.iter_weights(torch.device("cpu")):
        assert t.dtype == torch.bfloat16 and t.is_contiguous()
        seen[(role, layer)] = t
    assert len(seen) == 3 + 9 * ARCH.layers
    for name, src in sd.items():
        assert torch.equal(seen[classify(name)], src.to(torch.bfloat16)), name


def test_tied_embeddings_yield_lm_head_from_embed(tmp_path):
    sd = _random_tensors(ARCH, seed=3)
    save_checkpoint(str(tmp_path), ARCH, sd.items(), tie_word_embeddings=True)
    ck = CheckpointLlama(str(tmp_path), device="cpu")
    assert ck.tied and "lm_head.weight" not in ck.weight_map
    got = {(r, l): t for r, l, t in ck.iter_weights(torch.device("cpu"))}
    assert torch.equal(got[(_lib.LSK_W_LM_HEAD, 0)], sd["model.embed_tokens.weight"])
    assert torch.equal(ck.state_dict()["lm_head.weight"], sd["model.embed_tokens.weight"])


def test_tied_detected_when_config_flag_is_missing(tmp_path):
    sd = _random_tensors(ARCH, seed=4)
    del sd["lm_head.weight"]
    save_checkpoint(str(tmp_path), ARCH, sd.items())
    assert CheckpointLlama(str(tmp_path), device="cpu").tied


def test_torch_bin_checkpoint(tmp_path):
    sd = _random_tensors(ARCH, seed=5)
    names = list(sd)
    half = len(names) // 2
    torch.save({k: sd[k] for k in names[:half]}, tmp_path / "pytorch_model-00001-of-00002.bin")
    torch.save({k: sd[k] for k in names[half:]}, tmp_path / "pytorch_model-00002-of-00002.bin")
    wm = {k: "pytorch_model-00001-of-00002.bin" for k in names[:half]}
    wm.update({k: "pytorch_model-00002-of-00002.bin" for k in names[half:]})
    json.dump({"metadata": {}, "weight_map": wm}, open(tmp_path / "pytorch_model.bin.index.json", "w"))
    json.dump(config_json_of(ARCH), open(tmp_path / "config.json", "w"))
    _same(CheckpointLlama(str(tmp_path), device="cpu").state_dict(), sd)


def test_config_spellings():
    base = config_json_of(ARCH)
    assert arch_from_config_json(base) == (ARCH, False)
    v5 = dict(base)
    del v5["rope_theta"]
    v5["rope_parameters"] = {"rope_type": "default", "rope_theta": 50000.0}
    assert arch_from_config_json(v5)[0] == ARCH
    mha = dict(base)
    del mha["num_key_value_heads"], mha["head_dim"]
    a = arch_from_config_json(mha)[0]
    assert a.kv_heads == ARCH.heads and a.head_dim == ARCH.hidden // ARCH.heads
    # RoPE scaling rules, transformers-4 spelling (`rope_scaling`, keys `rope_type` or `type`) and
    # transformers-5 spelling (`rope_parameters`)
    l3 = {"rope_type": "llama3", "factor": 32.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
          "original_max_position_embeddings": 8192}
    a = arch_from_config_json({**base, "rope_scaling": l3})[0]
    assert (a.rope_scaling, a.rope_factor, a.rope_high_freq_factor, a.rope_original_max_pos) == \
        ("llama3", 32.0, 4.0, 8192)
    assert arch_from_config_json({**v5, "rope_parameters": {**l3, "rope_theta": 50000.0}})[0] == a
    lin = arch_from_config_json({**base, "rope_scaling": {"type": "linear", "factor": 2.0}})[0]
    assert (lin.rope_scaling, lin.rope_factor) == ("linear", 2.0)
    for bad in ({"rope_scaling": {"rope_type": "yarn", "factor": 8.0}},
                {"rope_scaling": {"type": "dynamic", "factor": 2.0}},
                {"model_type": "mistral"}, {"attention_bias": True}):
        with pytest.raises(CheckpointError):
            arch_from_config_json({**base, **bad})


def test_hf_model_config_with_transformers4_rope_scaling_is_not_silently_unscaled():
    """a 4.x-style config object carries llama3 scaling in `cfg.rope_scaling`, not in
    `rope_parameters`; it must reach the engine's RoPE table (or be refused), never be ignored."""
    from layerskip_b200.weights import LlamaArch
    cfg = type("Cfg", (), dict(
        vocab_size=128256, hidden_size=2048, intermediate_size=8192, num_hidden_layers=16,
        num_attention_heads=32, num_key_value_heads=8, head_dim=64, rms_norm_eps=1e-5,
        rope_theta=500000.0,
        rope_scaling={"rope_type": "llama3", "factor": 32.0, "low_freq_factor": 1.0,
                      "high_freq_factor": 4.0, "original_max_position_embeddings": 8192}))()
    a = LlamaArch.from_hf_config(cfg)
    assert a == ARCHS["llama3.2-1b"]
    cfg.rope_scaling = {"type": "yarn", "factor": 4.0}
    with pytest.raises(NotImplementedError):
        LlamaArch.from_hf_config(cfg)


def test_named_archs_survive_config_json():
    for name, arch in ARCHS.items():
        assert arch_from_config_json(json.loads(json.dumps(config_json_of(arch))))[0] == arch, name


def test_missing_and_misshapen_tensors_fail_loudly(tmp_path):
    sd = _random_tensors(ARCH, seed=6)
    broken = dict(sd)
    del broken["model.layers.1.mlp.up_proj.weight"]
    save_checkpoint(str(tmp_path / "a"), ARCH, broken.items())
    with pytest.raises(CheckpointError, match="1 tensors missing"):
        CheckpointLlama(str(tmp_path / "a"), device="cpu")
    bent = dict(sd)
    bent["model.layers.0.self_attn.k_proj.weight"] = torch.zeros(ARCH.q_dim, ARCH.hidden,
                                                                 dtype=torch.bfloat16)
    save_checkpoint(str(tmp_path / "b"), ARCH, bent.items())
    with pytest.raises(CheckpointError, match="k_proj.*shape"):
        CheckpointLlama(str(tmp_path / "b"), device="cpu").state_dict()
    os.makedirs(tmp_path / "c")
    with pytest.raises(CheckpointError, match="config.json"):
        CheckpointLlama(str(tmp_path / "c"), device="cpu")
    json.dump(config_json_of(ARCH), open(tmp_path / "c" / "config.json", "w"))
    with pytest.raises(CheckpointError, match="no model.safetensors"):
        CheckpointLlama(str(tmp_path / "c"), device="cpu")


def _hf_model(arch, seed):
    import transformers
    cfg = transformers.LlamaConfig(
        vocab_size=arch.vocab, hidden_size=arch.hidden, intermediate_size=arch.inter,
        num_hidden_layers=arch.layers, num_attention_heads=arch.heads,
        num_key_value_heads=arch.kv_heads, head_dim=arch.head_dim, rms_norm_eps=arch.rms_eps,
        rope_theta=arch.rope_theta, tie_word_embeddings=False, attention_bias=False)
    torch.manual_seed(seed)
    return transformers.LlamaForCausalLM(cfg).to(torch.bfloat16).eval()


def test_reads_what_transformers_save_pretrained_writes(tmp_path):
    model = _hf_model(ARCH, 7)
    model.save_pretrained(str(tmp_path), max_shard_size="60KB")
    assert (tmp_path / "model.safetensors.index.json").exists()
    ck = CheckpointLlama(str(tmp_path), device="cpu")
    assert ck.arch == ARCH
    want = {k: v for k, v in model.state_dict().items() if classify(k) is not None}
    _same(ck.state_dict(), want)


def test_transformers_from_pretrained_reads_what_we_write(tmp_path):
    import transformers
    sd = _random_tensors(ARCH, seed=8)
    save_checkpoint(str(tmp_path), ARCH, sd.items(), max_shard_bytes=50_000)
    model = transformers.LlamaForCausalLM.from_pretrained(str(tmp_path), dtype=torch.bfloat16)
    got = {k: v for k, v in model.state_dict().items() if classify(k) is not None}
    _same(got, sd)
    assert model.config.num_key_value_heads == ARCH.kv_heads
