"""Streaming ingest of Hugging Face Llama checkpoint directories.

The reference loads `facebook/layerskip-*` with `AutoModelForCausalLM.from_pretrained`
(`/root/reference/generate.py:54-67`), which materialises the whole model in host memory first.
The engine only ever needs ONE tensor at a time (`lsk_load_weights` slices the tensor-parallel
shard and repacks it into its own HBM buffers), so `CheckpointLlama` walks the checkpoint's
shards lazily instead: `config.json` -> `LlamaArch`, then every tensor the engine consumes is
read from its `.safetensors` (or `.bin`) shard, converted to bf16 on the target GPU, handed over
and dropped.  Peak host memory = one tensor; with tensor parallelism every rank streams the same
files and keeps only its slice on the device.

`save_checkpoint` writes the same directory format (sharded safetensors + index + config.json):
it exports a synthetic model so that the *reference* scripts can be pointed at the very weights
the engine was benchmarked on, and it is what the CPU tests use as a fixture.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Iterable, Iterator, List, Optional, Tuple

import torch

from . import _lib
from .weights import LlamaArch, classify, parse_rope

SAFETENSORS_INDEX = "model.safetensors.index.json"
SAFETENSORS_SINGLE = "model.safetensors"
TORCH_INDEX = "pytorch_model.bin.index.json"
TORCH_SINGLE = "pytorch_model.bin"


class CheckpointError(RuntimeError):
    pass


def arch_from_config_json(cfg: Dict) -> Tuple[LlamaArch, bool]:
    """(`LlamaArch`, tie_word_embeddings) from a parsed HF `config.json` — both the
    transformers-4 spelling (`rope_theta`, `rope_scaling`) and the transformers-5 one
    (`rope_parameters`) are accepted."""
    mt = cfg.get("model_type", "llama")
    if mt != "llama":
        raise CheckpointError(f"model_type {mt!r}: only Llama checkpoints are supported")
    if cfg.get("attention_bias") or cfg.get("mlp_bias"):
        raise CheckpointError("attention_bias / mlp_bias checkpoints are not supported")
    heads = int(cfg["num_attention_heads"])
    hidden = int(cfg["hidden_size"])
    head_dim = int(cfg.get("head_dim") or hidden // heads)
    try:
        rope = parse_rope(cfg.get)
    except NotImplementedError as exc:
        raise CheckpointError(str(exc)) from exc
    arch = LlamaArch(vocab=int(cfg["vocab_size"]), hidden=hidden,
                     inter=int(cfg["intermediate_size"]), layers=int(cfg["num_hidden_layers"]),
                     heads=heads, kv_heads=int(cfg.get("num_key_value_heads") or heads),
                     head_dim=head_dim, rms_eps=float(cfg.get("rms_norm_eps", 1e-5)), **rope)
    return arch, bool(cfg.get("tie_word_embeddings", False))


def config_json_of(arch: LlamaArch, tie_word_embeddings: bool = False) -> Dict:
    return {
        "architectures": ["LlamaForCausalLM"], "model_type": "llama",
        "vocab_size": arch.vocab, "hidden_size": arch.hidden, "intermediate_size": arch.inter,
        "num_hidden_layers": arch.layers, "num_attention_heads": arch.heads,
        "num_key_value_heads": arch.kv_heads, "head_dim": arch.head_dim,
        "rms_norm_eps": arch.rms_eps, "rope_theta": arch.rope_theta,
        "rope_scaling": arch.rope_config(), "hidden_act": "silu",
        "attention_bias": False, "mlp_bias": False, "tie_word_embeddings": tie_word_embeddings,
        "torch_dtype": "bfloat16", "max_position_embeddings": 4096,
    }


def expected_shapes(arch: LlamaArch, tie_word_embeddings: bool = False
                    ) -> Dict[str, Tuple[int, ...]]:
    """HF parameter name -> shape, for every tensor the engine ingests."""
    out: Dict[str, Tuple[int, ...]] = {"model.embed_tokens.weight": (arch.vocab, arch.hidden),
                                       "model.norm.weight": (arch.hidden,)}
    if not tie_word_embeddings:
        out["lm_head.weight"] = (arch.vocab, arch.hidden)
    for i in range(arch.layers):
        p = f"model.layers.{i}."
        out[p + "input_layernorm.weight"] = (arch.hidden,)
        out[p + "post_attention_layernorm.weight"] = (arch.hidden,)
        out[p + "self_attn.q_proj.weight"] = (arch.q_dim, arch.hidden)
        out[p + "self_attn.k_proj.weight"] = (arch.kv_dim, arch.hidden)
        out[p + "self_attn.v_proj.weight"] = (arch.kv_dim, arch.hidden)
        out[p + "self_attn.o_proj.weight"] = (arch.hidden, arch.q_dim)
        out[p + "mlp.gate_proj.weight"] = (arch.inter, arch.hidden)
        out[p + "mlp.up_proj.weight"] = (arch.inter, arch.hidden)
        out[p + "mlp.down_proj.weight"] = (arch.hidden, arch.inter)
    return out


class _ShardReader:
    """Keeps at most one shard file open; tensors come out one at a time."""

    def __init__(self, root: str):
        self.root = root
        self._path: Optional[str] = None
        self._handle = None          # safetensors handle or a dict from torch.load(mmap=True)
        self._ctx = None

    def close(self) -> None:
        if self._ctx is not None:
            self._ctx.__exit__(None, None, None)
        self._path = self._handle = self._ctx = None

    def _open(self, fname: str):
        if fname == self._path:
            return self._handle
        self.close()
        path = os.path.join(self.root, fname)
        if fname.endswith(".safetensors"):
            from safetensors import safe_open
            self._ctx = safe_open(path, framework="pt", device="cpu")
            self._handle = self._ctx.__enter__()
        else:
            self._handle = torch.load(path, map_location="cpu", mmap=True, weights_only=True)
        self._path = fname
        return self._handle

    def keys(self, fname: str) -> List[str]:
        return list(self._open(fname).keys())

    def get(self, fname: str, name: str) -> torch.Tensor:
        h = self._open(fname)
        return h.get_tensor(name) if fname.endswith(".safetensors") else h[name]


def _weight_map(root: str) -> Dict[str, str]:
    """tensor name -> shard file name."""
    for index in (SAFETENSORS_INDEX, TORCH_INDEX):
        p = os.path.join(root, index)
        if os.path.exists(p):
            with open(p) as f:
                return dict(json.load(f)["weight_map"])
    reader = _ShardReader(root)
    try:
        for single in (SAFETENSORS_SINGLE, TORCH_SINGLE):
            if os.path.exists(os.path.join(root, single)):
                return {k: single for k in reader.keys(single)}
        shards = sorted(f for f in os.listdir(root) if f.endswith(".safetensors"))
        if shards:
            out: Dict[str, str] = {}
            for s in shards:
                for k in reader.keys(s):
                    out[k] = s
            return out
    finally:
        reader.close()
    raise CheckpointError(f"no model.safetensors / pytorch_model.bin (or index) under {root}")


class CheckpointLlama:
    """A Llama checkpoint directory, duck-typed like the `model` argument of the reference's
    `generate_token_ids` where the B200 strategies need it (`.config`, `.device`) and like
    `SyntheticLlama` for the engine (`.arch`, `.iter_weights(device)`)."""

    def __init__(self, path: str, device: str = "cuda"):
        self.path = os.path.abspath(path)
        cfg_path = os.path.join(self.path, "config.json")
        if not os.path.exists(cfg_path):
            raise CheckpointError(f"{cfg_path} not found")
        with open(cfg_path) as f:
            self.config_json = json.load(f)
        self.arch, self.tied = arch_from_config_json(self.config_json)
        self.device = torch.device(device)
        self.weight_map = _weight_map(self.path)
        self._validate_names()

    def _validate_names(self) -> None:
        missing = [n for n in expected_shapes(self.arch, self.tied) if n not in self.weight_map]
        if "lm_head.weight" in missing and "model.embed_tokens.weight" in self.weight_map:
            missing.remove("lm_head.weight")      # tied head stored once (HF omits the alias)
            self.tied = True
        if missing:
            head = ", ".join(missing[:4]) + (" ..." if len(missing) > 4 else "")
            raise CheckpointError(f"{self.path}: {len(missing)} tensors missing ({head})")

    @property
    def config(self):
        a = self.arch
        return type("Cfg", (), dict(
            vocab_size=a.vocab, hidden_size=a.hidden, intermediate_size=a.inter,
            num_hidden_layers=a.layers, num_attention_heads=a.heads,
            num_key_value_heads=a.kv_heads, head_dim=a.head_dim, rms_norm_eps=a.rms_eps,
            rope_theta=a.rope_theta, rope_scaling=a.rope_config()))()

    def plan(self) -> List[Tuple[str, str]]:
        """(shard file, tensor name) in an order that opens every shard exactly once."""
        want = expected_shapes(self.arch, self.tied)
        by_file: Dict[str, List[str]] = {}
        for name in want:
            by_file.setdefault(self.weight_map[name], []).append(name)
        return [(fname, n) for fname in sorted(by_file) for n in by_file[fname]]

    def iter_named(self) -> Iterator[Tuple[str, torch.Tensor]]:
        """(HF name, host tensor in the checkpoint's dtype), shapes checked against config.json."""
        want = expected_shapes(self.arch, self.tied)
        reader = _ShardReader(self.path)
        try:
            for fname, name in self.plan():
                t = reader.get(fname, name)
                if tuple(t.shape) != want[name]:
                    raise CheckpointError(f"{name}: shape {tuple(t.shape)} in {fname}, config.json "
                                          f"implies {want[name]}")
                if not t.is_floating_point():
                    raise CheckpointError(f"{name}: dtype {t.dtype} (quantised checkpoints are "
                                          "not supported)")
                yield name, t
        finally:
            reader.close()

    def iter_weights(self, device: torch.device) -> Iterator[Tuple[int, int, torch.Tensor]]:
        """(role, layer, contiguous bf16 tensor on `device`) — what `Engine.load_weights` eats."""
        embed_dev = None
        for name, t in self.iter_named():
            role, layer = classify(name)
            d = t.to(device=device, dtype=torch.bfloat16, non_blocking=False).contiguous()
            if self.tied and role == _lib.LSK_W_EMBED:
                embed_dev = d
            yield role, layer, d
        if self.tied:
            yield _lib.LSK_W_LM_HEAD, 0, embed_dev

    def state_dict(self, dtype: torch.dtype = torch.bfloat16, device: str = "cpu"
                   ) -> Dict[str, torch.Tensor]:
        """Everything at once (small models / tests / handing the CPU baseline the same weights)."""
        sd = {n: t.to(device=device, dtype=dtype) for n, t in self.iter_named()}
        if self.tied:
            sd["lm_head.weight"] = sd["model.embed_tokens.weight"]
        return sd


def save_checkpoint(path: str, arch: LlamaArch, tensors: Iterable[Tuple[str, torch.Tensor]],
                    max_shard_bytes: int = 4 << 30, tie_word_embeddings: bool = False) -> List[str]:
    """Write `config.json` + sharded safetensors (+ index) from a stream of (HF name, tensor).
    Only one shard is held in host memory at a time.  Returns the shard file names."""
    from safetensors.torch import save_file
    os.makedirs(path, exist_ok=True)
    pending: Dict[str, torch.Tensor] = {}
    pending_bytes = 0
    shards: List[Dict[str, int]] = []          # per shard: name -> nbytes
    tmp_names: List[str] = []

    def flush():
        nonlocal pending, pending_bytes
        if not pending:
            return
        tmp = f"model-{len(shards) + 1:05d}.safetensors.part"
        save_file(pending, os.path.join(path, tmp), metadata={"format": "pt"})
        shards.append({k: v.numel() * v.element_size() for k, v in pending.items()})
        tmp_names.append(tmp)
        pending, pending_bytes = {}, 0

    for name, t in tensors:
        if tie_word_embeddings and name == "lm_head.weight":
            continue
        t = t.detach().to("cpu").contiguous()
        nbytes = t.numel() * t.element_size()
        if pending and pending_bytes + nbytes > max_shard_bytes:
            flush()
        pending[name] = t
        pending_bytes += nbytes
    flush()

    n = len(shards)
    final: List[str] = []
    weight_map: Dict[str, str] = {}
    for i, (tmp, content) in enumerate(zip(tmp_names, shards)):
        fname = SAFETENSORS_SINGLE if n == 1 else f"model-{i + 1:05d}-of-{n:05d}.safetensors"
        os.replace(os.path.join(path, tmp), os.path.join(path, fname))
        final.append(fname)
        for k in content:
            weight_map[k] = fname
    if n > 1:
        total = sum(sum(c.values()) for c in shards)
        with open(os.path.join(path, SAFETENSORS_INDEX), "w") as f:
            json.dump({"metadata": {"total_size": total}, "weight_map": weight_map}, f, indent=1)
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(config_json_of(arch, tie_word_embeddings), f, indent=1)
    return final
