"""TEST INFRASTRUCTURE ONLY — CPU oracle for the LayerSkip self-speculative decoding path.

A plain-torch, fp32, single-thread-friendly *restatement* (not a copy) of the algorithm the
reference implements in
  /root/reference/self_speculation/self_speculation_generator.py:32-229   (round loop)
  /root/reference/self_speculation/llama_model_utils.py:109-391           (stepping)
  /root/reference/self_speculation/autoregressive_generator.py:26-80      (AR counterpart)
plus the third-party arithmetic those call (HuggingFace `LlamaDecoderLayer`, transformers
pinned ==4.50.0 in /root/reference/requirements.txt:4; restated from the published Llama
equations as implemented in the installed copy
`transformers/models/llama/modeling_llama.py:52-70` RMSNorm, `:73-135` rotary table,
`:138-168` rotate_half RoPE, `:171-184` SwiGLU MLP, `:187-221` eager attention).

PINNING: this oracle is pinned against outputs of the reference itself, run unmodified in
the build container under `oracle/ref_shim.py`; the vectors live in `tests/golden/*.json`
and were written by `oracle/gen_golden.py`.  `tests/test_oracle_golden.py` re-checks them on
every CPU run, and (when /root/reference is present) re-runs the live reference too.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module.  The product package (`layerskip_b200/`) never does.

State model (differs from the reference on purpose, same results): instead of legacy KV
tuples that are re-concatenated and cropped, each layer owns a growing [kv_heads, len, hd]
key and value tensor and the two facts the reference derives from tuple shapes are explicit
integers: `early_len` (entries in layers < E) and `full_len` (entries in layers >= E).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch


# ----------------------------------------------------------------------------------------
# weights
# ----------------------------------------------------------------------------------------
@dataclass
class LlamaDims:
    vocab: int
    hidden: int
    inter: int
    layers: int
    heads: int
    kv_heads: int
    head_dim: int
    rms_eps: float = 1e-5
    rope_theta: float = 10000.0
    # HF `rope_scaling` dict (rope_type "linear" | "llama3" + its parameters) or None
    rope_scaling: Optional[Dict] = None

    @property
    def q_dim(self) -> int:
        return self.heads * self.head_dim

    @property
    def kv_dim(self) -> int:
        return self.kv_heads * self.head_dim


@dataclass
class OracleWeights:
    dims: LlamaDims
    embed: torch.Tensor                 # [V, h]
    final_norm: torch.Tensor            # [h]
    lm_head: torch.Tensor               # [V, h]
    layers: List[Dict[str, torch.Tensor]] = field(default_factory=list)
    # per layer: ln1 [h], wq [q_dim,h], wk [kv_dim,h], wv [kv_dim,h], wo [h,q_dim],
    #            ln2 [h], wg [I,h], wu [I,h], wd [h,I]

    def to(self, dtype: torch.dtype) -> "OracleWeights":
        conv = lambda t: t.to(dtype)
        return OracleWeights(
            dims=self.dims, embed=conv(self.embed), final_norm=conv(self.final_norm),
            lm_head=conv(self.lm_head),
            layers=[{k: conv(v) for k, v in layer.items()} for layer in self.layers])


def dims_from_hf_config(cfg) -> LlamaDims:
    head_dim = getattr(cfg, "head_dim", None) or cfg.hidden_size // cfg.num_attention_heads
    theta = None
    rp = getattr(cfg, "rope_parameters", None)
    if isinstance(rp, dict):
        theta = rp.get("rope_theta")
    if theta is None:
        theta = getattr(cfg, "rope_theta", 10000.0)
    scaling = None
    for src in (rp, getattr(cfg, "rope_scaling", None)):
        if isinstance(src, dict) and src.get("rope_type", src.get("type", "default")) not in (None, "default"):
            scaling = {k: v for k, v in src.items() if k != "rope_theta"}
    return LlamaDims(
        vocab=cfg.vocab_size, hidden=cfg.hidden_size, inter=cfg.intermediate_size,
        layers=cfg.num_hidden_layers, heads=cfg.num_attention_heads,
        kv_heads=cfg.num_key_value_heads, head_dim=head_dim,
        rms_eps=float(cfg.rms_norm_eps), rope_theta=float(theta), rope_scaling=scaling)


def weights_from_state_dict(dims: LlamaDims, sd: Dict[str, torch.Tensor],
                            dtype: torch.dtype = torch.float32) -> OracleWeights:
    """HF parameter names -> oracle weights (values converted to `dtype`, default fp32)."""
    g = lambda name: sd[name].detach().to("cpu").to(dtype)
    lm = sd.get("lm_head.weight", sd["model.embed_tokens.weight"])
    w = OracleWeights(dims=dims, embed=g("model.embed_tokens.weight"),
                      final_norm=g("model.norm.weight"),
                      lm_head=lm.detach().to("cpu").to(dtype))
    for i in range(dims.layers):
        p = f"model.layers.{i}."
        w.layers.append(dict(
            ln1=g(p + "input_layernorm.weight"),
            wq=g(p + "self_attn.q_proj.weight"), wk=g(p + "self_attn.k_proj.weight"),
            wv=g(p + "self_attn.v_proj.weight"), wo=g(p + "self_attn.o_proj.weight"),
            ln2=g(p + "post_attention_layernorm.weight"),
            wg=g(p + "mlp.gate_proj.weight"), wu=g(p + "mlp.up_proj.weight"),
            wd=g(p + "mlp.down_proj.weight")))
    return w


def weights_from_hf(model, dtype: torch.dtype = torch.float32) -> OracleWeights:
    return weights_from_state_dict(dims_from_hf_config(model.config), model.state_dict(), dtype)


# ----------------------------------------------------------------------------------------
# arithmetic (third-party layer restated)
# ----------------------------------------------------------------------------------------
def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    """modeling_llama.py:52-70 — fp32 mean of squares, rsqrt, cast back, then * weight."""
    xf = x.to(torch.float32)
    inv = torch.rsqrt(xf.pow(2).mean(dim=-1, keepdim=True) + eps)
    return weight * (xf * inv).to(x.dtype)


def rope_tables(dims: LlamaDims, positions: torch.Tensor, dtype: torch.dtype
                ) -> Tuple[torch.Tensor, torch.Tensor]:
    """modeling_llama.py:73-135 — inv_freq = theta^(-2i/d); angle table duplicated over halves."""
    half = torch.arange(0, dims.head_dim, 2, dtype=torch.int64).to(torch.float32)
    inv_freq = scale_inv_freq(1.0 / (dims.rope_theta ** (half / dims.head_dim)), dims.rope_scaling)
    ang = positions.to(torch.float32)[:, None] * inv_freq[None, :]          # [s, d/2]
    ang = torch.cat([ang, ang], dim=-1)                                      # [s, d]
    return ang.cos().to(dtype), ang.sin().to(dtype)


def scale_inv_freq(inv_freq: torch.Tensor, rs: Optional[Dict]) -> torch.Tensor:
    """transformers modeling_rope_utils.py, restated: `_compute_linear_scaling_rope_parameters`
    (inv_freq / factor) and `_compute_llama3_parameters` (wavelengths longer than
    old_ctx / low_freq_factor are divided by factor, shorter than old_ctx / high_freq_factor kept,
    the band in between interpolated).  Both return attention_factor 1."""
    if not rs:
        return inv_freq
    kind = rs.get("rope_type", rs.get("type", "default"))
    if kind == "default":
        return inv_freq
    if kind == "linear":
        return inv_freq / rs["factor"]
    if kind != "llama3":
        raise NotImplementedError(kind)
    factor, lo, hi = rs["factor"], rs["low_freq_factor"], rs["high_freq_factor"]
    old = rs["original_max_position_embeddings"]
    wavelen = 2 * math.pi / inv_freq
    scaled = torch.where(wavelen > old / lo, inv_freq / factor, inv_freq)
    smooth = (old / wavelen - lo) / (hi - lo)
    mid = (1 - smooth) * scaled / factor + smooth * scaled
    medium = ~(wavelen < old / hi) & ~(wavelen > old / lo)
    return torch.where(medium, mid, scaled)


def apply_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """modeling_llama.py:138-168 — split-halves ("rotate_half") layout. x: [heads, s, d]."""
    d = x.shape[-1] // 2
    rotated = torch.cat([-x[..., d:], x[..., :d]], dim=-1)
    return x * cos[None] + rotated * sin[None]


class KVStore:
    """Per-layer growing K/V, keys stored post-RoPE (modeling_llama.py:262-289)."""

    def __init__(self, n_layers: int):
        self.k: List[Optional[torch.Tensor]] = [None] * n_layers
        self.v: List[Optional[torch.Tensor]] = [None] * n_layers

    def length(self, layer: int) -> int:
        return 0 if self.k[layer] is None else self.k[layer].shape[1]

    def append(self, layer: int, k: torch.Tensor, v: torch.Tensor):
        if self.k[layer] is None:
            self.k[layer], self.v[layer] = k, v
        else:
            self.k[layer] = torch.cat([self.k[layer], k], dim=1)
            self.v[layer] = torch.cat([self.v[layer], v], dim=1)
        return self.k[layer], self.v[layer]

    def crop(self, max_len: int) -> None:
        """llama_model_utils.py:134-149 — every initialised layer sliced to [:max_len]."""
        for i in range(len(self.k)):
            if self.k[i] is not None:
                self.k[i] = self.k[i][:, :max_len]
                self.v[i] = self.v[i][:, :max_len]


def additive_causal_mask(n_query: int, n_past: int, dtype: torch.dtype) -> Optional[torch.Tensor]:
    """llama_model_utils.py:21-73 — with an all-ones padding mask the result is: zeros when
    n_query == 1 (no causal part is built, :25), else `finfo.min` strictly above the diagonal
    shifted by n_past.  Returned as [n_query, n_past + n_query]."""
    total = n_past + n_query
    mask = torch.zeros(n_query, total, dtype=dtype)
    if n_query > 1:
        q_idx = torch.arange(n_query)[:, None] + n_past
        k_idx = torch.arange(total)[None, :]
        mask = mask.masked_fill(k_idx > q_idx, torch.finfo(dtype).min)
    return mask


def decoder_layer(w: OracleWeights, li: int, x: torch.Tensor, positions: torch.Tensor,
                  kv: KVStore, mask: torch.Tensor) -> torch.Tensor:
    """One LlamaDecoderLayer on x:[s,h] at `positions`, appending to the layer's K/V.
    modeling_llama.py:292-332 (layer), :225-289 (attention), :187-221 (eager math)."""
    d, L = w.dims, w.layers[li]
    s = x.shape[0]
    h1 = rms_norm(x, L["ln1"], d.rms_eps)
    q = (h1 @ L["wq"].T).view(s, d.heads, d.head_dim).transpose(0, 1)       # [H, s, hd]
    k = (h1 @ L["wk"].T).view(s, d.kv_heads, d.head_dim).transpose(0, 1)
    v = (h1 @ L["wv"].T).view(s, d.kv_heads, d.head_dim).transpose(0, 1)
    cos, sin = rope_tables(d, positions, x.dtype)
    q, k = apply_rope(q, cos, sin), apply_rope(k, cos, sin)
    k_all, v_all = kv.append(li, k, v)                                      # [KV, n, hd]
    rep = d.heads // d.kv_heads
    if rep > 1:                                                              # repeat_kv :171-184
        k_all = k_all.repeat_interleave(rep, dim=0)
        v_all = v_all.repeat_interleave(rep, dim=0)
    scores = (q @ k_all.transpose(1, 2)) * (d.head_dim ** -0.5) + mask[None]
    probs = torch.softmax(scores, dim=-1, dtype=torch.float32).to(x.dtype)
    ctx = (probs @ v_all).transpose(0, 1).reshape(s, d.q_dim)
    x = x + ctx @ L["wo"].T
    h2 = rms_norm(x, L["ln2"], d.rms_eps)
    act = torch.nn.functional.silu(h2 @ L["wg"].T) * (h2 @ L["wu"].T)
    return x + act @ L["wd"].T


def lm_logits(w: OracleWeights, x: torch.Tensor) -> torch.Tensor:
    return rms_norm(x, w.final_norm, w.dims.rms_eps) @ w.lm_head.T


# ----------------------------------------------------------------------------------------
# stepping (llama_model_utils.py:155-391 restated on explicit state)
# ----------------------------------------------------------------------------------------
def step_all_layers(w: OracleWeights, ids: Sequence[int], kv: KVStore) -> torch.Tensor:
    """`forward` (llama_model_utils.py:155-209): every layer on ids, logits for all rows."""
    past = kv.length(0)
    s = len(ids)
    pos = torch.arange(past, past + s)
    x = w.embed[torch.tensor(list(ids), dtype=torch.long)]
    mask = additive_causal_mask(s, past, x.dtype)
    for li in range(w.dims.layers):
        x = decoder_layer(w, li, x, pos, kv, mask)
    return lm_logits(w, x)


def step_early(w: OracleWeights, ids: Sequence[int], kv: KVStore, exit_layer: int,
               exit_rows: Optional[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    """`forward_early` (llama_model_utils.py:213-276): layers [0,E) on ids; the pre-norm
    hidden rows are appended to the exit-query cache (:266-269); logits for ALL rows (:271-273).
    Returns (logits [s,V], exit_rows [.., h])."""
    past = kv.length(0)
    s = len(ids)
    pos = torch.arange(past, past + s)
    x = w.embed[torch.tensor(list(ids), dtype=torch.long)]
    mask = additive_causal_mask(s, past, x.dtype)
    for li in range(exit_layer):
        x = decoder_layer(w, li, x, pos, kv, mask)
    exit_rows = x if exit_rows is None else torch.cat([exit_rows, x], dim=0)
    return lm_logits(w, x), exit_rows


def step_remainder(w: OracleWeights, ids: Sequence[int], kv: KVStore, exit_layer: int,
                   exit_rows: Optional[torch.Tensor], kv_touched_layers: int) -> torch.Tensor:
    """`forward_remainder` (llama_model_utils.py:280-391).

    ids = [round input tokens ; draft tokens]  (T + D of them).
    Layers < E run ONLY the last id (:350-362) at position draft_len (:353) against all
    draft_len+1 keys (mask :323-329 is all zeros for one query).  Layers >= E run
    [exit_rows ; that last row] (:364-371) — or just the current rows when there is no exit
    cache (:372-374) — at positions arange(full_len, draft_len+1) (:312-318) under a causal
    mask offset by full_len (:331-337).  `kv_touched_layers` restates `len(past_key_values)`
    (:301-305): unless every layer already holds entries, full_len is 0.
    """
    n_layers = w.dims.layers
    has_past = kv_touched_layers > 0
    if not has_past and len(ids) > 1 and exit_layer > 0:
        # reference quirk: a D_req == 0 round on a multi-token
        # prompt feeds one row against a [1,1,1,T] mask and raises; keep that visible.
        raise RuntimeError("forward_remainder without past on a multi-token input is "
                           "ill-formed in the reference (llama_model_utils.py:350-362)")
    draft_len = kv.length(0) if has_past else 0
    full_len = kv.length(n_layers - 1) if (has_past and kv_touched_layers == n_layers) else 0
    s = len(ids)
    total = (1 + draft_len) if has_past else s
    pos = torch.arange(full_len, total)
    # NB: the reference reshapes this to [1, s]; a length mismatch raises there (:318) and here.
    assert pos.numel() == s, "position/row mismatch (reference would raise in .view, :318)"
    x_all = w.embed[torch.tensor(list(ids), dtype=torch.long)]
    early_mask = additive_causal_mask(1, draft_len, x_all.dtype)
    full_mask = additive_causal_mask(s, full_len, x_all.dtype)
    x = x_all
    full_rows: Optional[torch.Tensor] = None
    for li in range(n_layers):
        if li < exit_layer:
            x = decoder_layer(w, li, x[-1:], pos[-1:], kv, early_mask)
        else:
            if full_rows is None and exit_rows is not None:
                full_rows = torch.cat([exit_rows, x[-1:]], dim=0)
            else:
                full_rows = x
            x = decoder_layer(w, li, full_rows, pos, kv, full_mask)
            full_rows = x
    return lm_logits(w, x)


# ----------------------------------------------------------------------------------------
# token selection (llama_model_utils.py:75-131)
# ----------------------------------------------------------------------------------------
def warp_top_k_top_p(logits: torch.Tensor, top_k: int, top_p: float) -> torch.Tensor:
    """llama_model_utils.py:75-107.  top-k only when > 0 (:97); nucleus ALWAYS when
    0 <= top_p <= 1 (:102).  Semantics of the two HF warpers: top-k removes everything below
    the k-th largest logit; top-p sorts ascending and removes the prefix whose cumulative
    softmax mass is <= 1 - top_p, always keeping the largest entry."""
    neg_inf = -float("inf")
    if top_k > 0:
        k = min(top_k, logits.shape[-1])
        kth = torch.topk(logits, k)[0][..., -1, None]
        logits = logits.masked_fill(logits < kth, neg_inf)
    if 0 <= top_p <= 1.0:
        srt, idx = torch.sort(logits, descending=False)
        mass = srt.softmax(dim=-1).cumsum(dim=-1)
        drop = mass <= (1 - top_p)
        drop[..., -1:] = False
        drop = drop.scatter(1, idx, drop)
        logits = logits.masked_fill(drop, neg_inf)
    return logits


def ban_repeated_ngrams(logits: torch.Tensor, seqs: Sequence[Sequence[int]], n: int) -> torch.Tensor:
    """HF `NoRepeatNGramLogitsProcessor` (transformers generation/logits_process.py,
    `_calc_banned_ngram_tokens`; built by the reference at generator_base.py:77-85) with its
    documented inputs: row r of `logits` continues the FULL token sequence `seqs[r]`; every token that
    would complete an n-gram already present in that sequence gets -inf.  (The reference hands the
    processor only the current step's ids — one token after the first step — so that its ban list
    stays empty; the engine and this oracle implement the processor's documented semantics.)"""
    if n <= 0:
        return logits
    out = logits.clone()
    for r, seq in enumerate(seqs):
        seq = list(seq)
        L = len(seq)
        if L + 1 < n:
            continue
        prefix = tuple(seq[L - (n - 1):]) if n > 1 else ()
        for i in range(0, L - n + 1):
            if tuple(seq[i:i + n - 1]) == prefix:
                out[r, seq[i + n - 1]] = -float("inf")
    return out


def pick_tokens(logits: torch.Tensor, last_only: bool, sample: bool, temperature: float,
                top_k: int, top_p: float) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """`decode_next_token` (llama_model_utils.py:109-131) on logits [s, V].
    Returns (tokens [rows], probs [rows, V] or None); rows = 1 if last_only else s.
    RNG call shape matches the reference (one multinomial over [rows, V]) so seeded
    sampling traces agree with the reference draw for draw."""
    rows = logits[-1:] if last_only else logits
    if not sample:
        return rows.argmax(dim=-1), None
    probs = torch.softmax(warp_top_k_top_p(rows / temperature, top_k, top_p), dim=-1)
    return torch.multinomial(probs, num_samples=1)[:, 0], probs


def residual_distribution(p_target: torch.Tensor, p_draft: torch.Tensor, eps: float = 1e-6
                          ) -> torch.Tensor:
    """`max_fn` (self_speculation_generator.py:27-29): norm(max(p_target - p_draft, 0))."""
    pos = torch.where(p_target - p_draft > 0, p_target - p_draft, torch.zeros(()))
    return pos / (pos.sum() + eps)


# ----------------------------------------------------------------------------------------
# strategies
# ----------------------------------------------------------------------------------------
@dataclass
class RoundTrace:
    n_input: int               # T: tokens fed this round (prompt length on round 0, else 1)
    d_req: int                 # requested speculations after the max_steps clamp
    draft: List[int]           # draft tokens actually produced (EOS can cut it short)
    verified: List[int]        # the D_actual+1 tokens the full model chose
    n_matches: int
    emitted: List[int]         # draft[:n] + [verified[n]]
    kv_len_after: int          # == n_prompt + n_out - 1


@dataclass
class OracleResult:
    predicted_tokens: List[int]
    acceptance_rate: Optional[float]
    rounds: List[RoundTrace] = field(default_factory=list)


def self_speculative_generate(
        w: OracleWeights, prompt: Sequence[int], eos_token_ids: Sequence[int], *,
        max_steps: int, exit_layer: int, num_speculations: int, sample: bool = False,
        temperature: float = 0.6, top_k: int = 0, top_p: float = 0.9,
        no_repeat_ngram_size: int = 0) -> OracleResult:
    """self_speculation_generator.py:32-99 (outer loop) and :102-229 (one round)."""
    prompt = list(prompt)
    kv = KVStore(w.dims.layers)
    kv_touched = 0                       # len(past_key_values) in the reference
    cur: List[int] = list(prompt)        # `input_ids` (:45, :203)
    out: List[int] = []
    matches_total = 0
    drafted_total = 0
    rounds: List[RoundTrace] = []
    while len(out) < max_steps:                                                   # :51
        d_req = min(num_speculations, max_steps - len(out) - 1)                   # :63-66
        # ---- draft (:127-148)
        draft: List[int] = []
        draft_probs: List[torch.Tensor] = []
        exit_rows: Optional[torch.Tensor] = None
        feed = list(cur)
        for _ in range(d_req):
            logits, exit_rows = step_early(w, feed, kv, exit_layer, exit_rows)
            kv_touched = max(kv_touched, exit_layer)
            if no_repeat_ngram_size:
                logits = ban_repeated_ngrams(logits[-1:], [prompt + out + draft], no_repeat_ngram_size)
            tok, prob = pick_tokens(logits, True, sample, temperature, top_k, top_p)
            t = int(tok[0])
            draft.append(t)
            if sample:
                draft_probs.append(prob)
            feed = [t]
            if t in eos_token_ids:                                                # :146-148
                break
        # ---- verify (:152-182)
        n_in = len(cur)
        logits = step_remainder(w, cur + draft, kv, exit_layer, exit_rows, kv_touched)
        kv_touched = w.dims.layers
        ver_logits = logits[n_in - 1:]                                            # :177
        if no_repeat_ngram_size:
            ver_logits = ban_repeated_ngrams(
                ver_logits, [prompt + out + draft[:j] for j in range(len(draft) + 1)], no_repeat_ngram_size)
        ver_tok, ver_prob = pick_tokens(ver_logits, False, sample, temperature, top_k, top_p)
        verified = [int(t) for t in ver_tok]
        # ---- accept (:185-199)
        n = 0
        if not sample:
            while n < len(draft) and draft[n] == verified[n]:
                n += 1
        else:
            u = torch.rand(1, len(draft), dtype=torch.float)                      # :193
            for i in range(len(draft)):
                ratio = ver_prob[i, draft[i]].item() / draft_probs[i][0, draft[i]].item()
                if u[0, i] < min(1, ratio):
                    n += 1
                else:
                    resid = residual_distribution(ver_prob[i, :], draft_probs[i])
                    verified[n] = int(torch.multinomial(resid, num_samples=1).item())
                    break
        emitted = draft[:n] + [verified[n]]                                       # :203-205
        out.extend(emitted)
        cur = [verified[n]]
        kv.crop(len(prompt) + len(out) - 1)                                       # :219-221
        rounds.append(RoundTrace(n_in, d_req, list(draft), verified, n, emitted,
                                 kv.length(0)))
        matches_total += n                                                        # :80
        drafted_total += len(draft)                                               # :81
        hit = False
        for e in eos_token_ids:                                                   # :82-91
            if e in out:
                out = out[: out.index(e)]
                hit = True
                break
        if hit:
            break
    rate = matches_total / drafted_total          # ZeroDivisionError as in the reference (:98)
    return OracleResult(out, rate, rounds)


def autoregressive_generate(
        w: OracleWeights, prompt: Sequence[int], eos_token_ids: Sequence[int], *,
        max_steps: int, exit_layer: int = -1, sample: bool = False, temperature: float = 0.6,
        top_k: int = 0, top_p: float = 0.9, no_repeat_ngram_size: int = 0) -> OracleResult:
    """autoregressive_generator.py:26-80: `forward` each step, or `forward_early` when
    exit_layer > 0 (:44-51); EOS is checked BEFORE the token is appended (:66-67)."""
    kv = KVStore(w.dims.layers)
    feed = list(prompt)
    out: List[int] = []
    exit_rows = None
    for _ in range(max_steps):
        if exit_layer > 0:
            logits, exit_rows = step_early(w, feed, kv, exit_layer, exit_rows)
        else:
            logits = step_all_layers(w, feed, kv)
        if no_repeat_ngram_size:
            logits = ban_repeated_ngrams(logits[-1:], [list(prompt) + out], no_repeat_ngram_size)
        tok, _ = pick_tokens(logits, True, sample, temperature, top_k, top_p)
        t = int(tok[0])
        if t in eos_token_ids:
            break
        out.append(t)
        feed = [t]
    return OracleResult(out, None, [])


# ----------------------------------------------------------------------------------------
# helpers for the margin-gated parity protocol (SURVEY.md §7.3 / BASELINE.md §5)
# ----------------------------------------------------------------------------------------
def teacher_forced_logits(w: OracleWeights, prompt: Sequence[int], continuation: Sequence[int]
                          ) -> torch.Tensor:
    """Full-model logits predicting each continuation token: row j is the distribution after
    prompt + continuation[:j].  One causal pass (same maths as `forward`)."""
    ids = list(prompt) + list(continuation)
    kv = KVStore(w.dims.layers)
    logits = step_all_layers(w, ids[:-1] if len(continuation) else ids, kv)
    return logits[len(prompt) - 1:]


def early_exit_logits(w: OracleWeights, prompt: Sequence[int], continuation: Sequence[int],
                      exit_layer: int) -> torch.Tensor:
    """Same as above for the draft sub-model (layers < E + shared head)."""
    ids = list(prompt) + list(continuation)
    kv = KVStore(w.dims.layers)
    logits, _ = step_early(w, ids[:-1] if len(continuation) else ids, kv, exit_layer, None)
    return logits[len(prompt) - 1:]


def top2_margin(logits_row: torch.Tensor) -> float:
    top = torch.topk(logits_row.to(torch.float32), 2).values
    return float(top[0] - top[1])


# ----------------------------------------------------------------------------------------
# synthetic model factory shared by tests / bench (deterministic, CPU generator)
# ----------------------------------------------------------------------------------------
def random_state_dict(dims: LlamaDims, seed: int = 0, damp_from_layer: Optional[int] = None,
                      alpha: float = 1.0, round_bf16: bool = True,
                      std: float = 0.02) -> Dict[str, torch.Tensor]:
    """HF-named fp32 tensors, N(0, std^2) linears/embeddings and unit RMSNorm weights (the HF
    default init); optionally rounded through bf16 so a bf16 engine and the
    fp32 oracle hold identical values; `alpha` damps o_proj/down_proj of layers >=
    damp_from_layer to control greedy acceptance."""
    g = torch.Generator().manual_seed(seed)

    def rnd(*shape):
        t = torch.randn(*shape, generator=g, dtype=torch.float32) * std
        return t.to(torch.bfloat16).to(torch.float32) if round_bf16 else t

    sd: Dict[str, torch.Tensor] = {}
    sd["model.embed_tokens.weight"] = rnd(dims.vocab, dims.hidden)
    for i in range(dims.layers):
        p = f"model.layers.{i}."
        damp = alpha if (damp_from_layer is not None and i >= damp_from_layer) else 1.0
        sd[p + "input_layernorm.weight"] = torch.ones(dims.hidden)
        sd[p + "self_attn.q_proj.weight"] = rnd(dims.q_dim, dims.hidden)
        sd[p + "self_attn.k_proj.weight"] = rnd(dims.kv_dim, dims.hidden)
        sd[p + "self_attn.v_proj.weight"] = rnd(dims.kv_dim, dims.hidden)
        wo = rnd(dims.hidden, dims.q_dim) * damp
        sd[p + "self_attn.o_proj.weight"] = wo.to(torch.bfloat16).to(torch.float32) if round_bf16 else wo
        sd[p + "post_attention_layernorm.weight"] = torch.ones(dims.hidden)
        sd[p + "mlp.gate_proj.weight"] = rnd(dims.inter, dims.hidden)
        sd[p + "mlp.up_proj.weight"] = rnd(dims.inter, dims.hidden)
        wd = rnd(dims.hidden, dims.inter) * damp
        sd[p + "mlp.down_proj.weight"] = wd.to(torch.bfloat16).to(torch.float32) if round_bf16 else wd
    sd["model.norm.weight"] = torch.ones(dims.hidden)
    sd["lm_head.weight"] = rnd(dims.vocab, dims.hidden)
    return sd


def _selfcheck() -> None:  # pragma: no cover - manual smoke
    dims = LlamaDims(vocab=97, hidden=64, inter=176, layers=4, heads=4, kv_heads=2, head_dim=16)
    w = weights_from_state_dict(dims, random_state_dict(dims, 0, 2, 0.1))
    a = self_speculative_generate(w, [3, 4, 5, 6], [96], max_steps=12, exit_layer=2,
                                  num_speculations=3)
    b = autoregressive_generate(w, [3, 4, 5, 6], [96], max_steps=12)
    assert a.predicted_tokens == b.predicted_tokens, (a, b)
    print("ok", a.predicted_tokens, a.acceptance_rate, math.isfinite(a.acceptance_rate))


if __name__ == "__main__":
    _selfcheck()
