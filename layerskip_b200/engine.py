"""Python handle on one liblsk engine (one per process / GPU).

Host code here is plumbing only: it marshals arguments into the C ABI (include/lsk.h).  All
arithmetic, the draft/verify/accept logic and the KV bookkeeping run in the CUDA library.
"""
from __future__ import annotations

import os

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch

from . import _lib
from .weights import ROPE_KINDS, LlamaArch, SyntheticLlama, iter_state_dict


@dataclass
class RoundOutput:
    """One speculation round as seen by the host (lsk_round_out)."""
    n_drafted: int
    n_matches: int
    emitted: List[int]
    draft: List[int]
    verified: List[int]
    kv_len: int


class Engine:
    def __init__(self, arch: LlamaArch, max_ctx: int = 4096, tp_rank: int = 0, tp_size: int = 1,
                 keep_logits: bool = False, use_pdl: bool = True, use_graph: bool = True,
                 attn_splits: int = 0, device: Optional[torch.device] = None,
                 tp_nccl: bool = False, prefill_tc: Optional[bool] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("layerskip_b200 needs a CUDA device (B200); there is no CPU path")
        self._lib = _lib.load()
        self.arch = arch
        self.device = torch.device(device) if device is not None else \
            torch.device("cuda", torch.cuda.current_device())
        self.tp_rank, self.tp_size = tp_rank, tp_size
        # the tcgen05 prompt pass needs a second (canonical-layout) copy of the layer weights: on by
        # default, dropped automatically when the two copies would not fit this GPU
        try:
            free_now = torch.cuda.mem_get_info(self.device)[0]
        except Exception:  # pragma: no cover
            free_now = None
        if prefill_tc is None:
            prefill_tc = os.environ.get("LSK_PREFILL_TC", "1") not in ("0",)
            if prefill_tc and free_now is not None:
                from .memory import plan_memory
                plan = plan_memory(arch, max_ctx=max_ctx, tp_size=tp_size, keep_logits=keep_logits,
                                   prefill_tc=True)
                if plan["total"] + plan["weights_source_peak"] > free_now:
                    prefill_tc = False
        self.prefill_tc = bool(prefill_tc)
        flags = (_lib.LSK_FLAG_KEEP_LOGITS if keep_logits else 0) | \
                (0 if prefill_tc else _lib.LSK_FLAG_NO_PREFILL_TC) | \
                (0 if use_pdl else _lib.LSK_FLAG_NO_PDL) | (0 if use_graph else _lib.LSK_FLAG_NO_GRAPH) | \
                (_lib.LSK_FLAG_TP_NCCL if tp_nccl else 0)
        cfg = _lib.lsk_config(
            vocab=arch.vocab, hidden=arch.hidden, inter=arch.inter, n_layers=arch.layers,
            n_heads=arch.heads, n_kv_heads=arch.kv_heads, head_dim=arch.head_dim,
            rms_eps=arch.rms_eps, rope_theta=arch.rope_theta, max_ctx=max_ctx, tp_rank=tp_rank,
            tp_size=tp_size, attn_splits=attn_splits, flags=flags,
            rope_scaling=ROPE_KINDS[arch.rope_scaling], rope_factor=arch.rope_factor,
            rope_low_freq_factor=arch.rope_low_freq_factor,
            rope_high_freq_factor=arch.rope_high_freq_factor,
            rope_original_max_pos=arch.rope_original_max_pos)
        self.max_ctx = max_ctx
        self.keep_logits = keep_logits
        # refuse a configuration that cannot fit BEFORE cudaMalloc fails half-way (memory.py)
        try:
            free_bytes = torch.cuda.mem_get_info(self.device)[0]
        except Exception:  # pragma: no cover - very old drivers
            free_bytes = None
        if free_bytes is not None:
            from .memory import check_fits
            check_fits(arch, free_bytes, max_ctx=max_ctx, tp_size=tp_size, keep_logits=keep_logits,
                       sampling=False, lm_head_tc=os.environ.get("LSK_LMHEAD_TC", "0") not in ("", "0"),
                       prefill_tc=self.prefill_tc)
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.lsk_create(C.byref(cfg), C.byref(handle)))
        self._h = handle
        self._exit_layer = -1
        # token rows one step can carry (engine.cu: max_rows): 16 when the host planner finds a
        # schedule for the 16-row RMSNorm GEMM at K = hidden (whole rows resident up to hidden 4096,
        # K-chunked normalisation above), else 8
        plan = _lib.lsk_gemm_plan()
        qkv_rows = (arch.heads + 2 * arch.kv_heads) // tp_size * arch.head_dim
        ok = self._lib.lsk_plan_gemm((qkv_rows + 15) // 16 * 16, arch.hidden, 16, 0, 0, 148, C.byref(plan))
        self.max_rows = 16 if (ok == 0 and plan.ok) else 8

    # ------------------------------------------------------------------ lifetime
    def close(self) -> None:
        if getattr(self, "_h", None):
            with torch.cuda.device(self.device):
                self._lib.lsk_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ tensor parallel
    def init_comm(self, process_group=None) -> None:
        """Create the engine's NCCL communicator; the unique id travels over torch.distributed
        (plumbing only — the data path uses the engine's own communicator and stream)."""
        if self.tp_size == 1:
            return
        import torch.distributed as dist
        from .parallel_util import broadcast_bytes
        uid = (C.c_uint8 * 128)()
        if dist.get_rank(process_group) == 0:
            _lib.check(self._lib.lsk_comm_unique_id(uid))
        raw = broadcast_bytes(bytes(uid), 128, src=0, group=process_group, device=self.device)
        arr = (C.c_uint8 * 128)(*raw)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.lsk_comm_init(self._h, arr))

    # ------------------------------------------------------------------ weights
    def load_weights(self, source: Iterable[Tuple[int, int, torch.Tensor]]) -> None:
        with torch.cuda.device(self.device):
            for role, layer, t in source:
                assert t.is_cuda and t.dtype == torch.bfloat16 and t.is_contiguous()
                # the tensor was produced on torch's stream; the engine packs on its own
                torch.cuda.current_stream(self.device).synchronize()
                rows = t.shape[0]
                cols = t.shape[1] if t.dim() == 2 else 1
                desc = _lib.lsk_weight_desc(role=role, layer=layer, data=t.data_ptr(), rows=rows,
                                            cols=cols)
                _lib.check(self._lib.lsk_load_weights(self._h, C.byref(desc), 1))

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        self.load_weights(iter_state_dict(sd, self.device))

    def load_model(self, model) -> None:
        """HF `LlamaForCausalLM` (any device / float dtype), or a streaming source with
        `iter_weights(device)` (`SyntheticLlama`, `checkpoint.CheckpointLlama`)."""
        if hasattr(model, "iter_weights"):
            self.load_weights(model.iter_weights(self.device))
        else:
            self.load_state_dict(model.state_dict())
        if not self._lib.lsk_weights_complete(self._h):
            raise RuntimeError("model did not provide every tensor the engine needs")

    # ------------------------------------------------------------------ generation
    def begin(self, exit_layer: int, max_steps: int, eos_token_ids: Sequence[int],
              sample: bool = False, temperature: float = 0.6, top_k: int = 0, top_p: float = 0.9,
              seed: int = 0, no_repeat_ngram_size: int = 0) -> None:
        eos = list(eos_token_ids)
        if len(eos) > _lib.LSK_MAX_EOS:
            raise ValueError(f"at most {_lib.LSK_MAX_EOS} eos ids are supported")
        gen = _lib.lsk_generation(exit_layer=exit_layer, max_steps=max_steps, n_eos=len(eos),
                                  sample=int(bool(sample)), temperature=temperature, top_k=top_k,
                                  top_p=top_p, seed=seed,
                                  no_repeat_ngram_size=int(no_repeat_ngram_size or 0))
        for i, t in enumerate(eos):
            gen.eos_ids[i] = int(t)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.lsk_begin(self._h, C.byref(gen)))
        self._exit_layer = exit_layer

    def prefill(self, prompt_ids: Sequence[int]) -> None:
        n = len(prompt_ids)
        arr = (C.c_int32 * n)(*[int(t) for t in prompt_ids])
        with torch.cuda.device(self.device):
            _lib.check(self._lib.lsk_prefill(self._h, arr, n))

    def round(self, d_req: int) -> RoundOutput:
        out = _lib.lsk_round_out()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.lsk_round(self._h, d_req, C.byref(out)))
        return RoundOutput(
            n_drafted=out.n_drafted, n_matches=out.n_matches,
            emitted=list(out.emitted_ids[:out.n_emitted]),
            draft=list(out.draft_ids[:out.n_drafted]),
            verified=list(out.verified_ids[:out.n_drafted + 1]), kv_len=out.kv_len)

    KERNEL_CLASSES = ("qkv", "attention", "o_proj", "gate_up", "down", "lm_head", "small", "comm")

    def profile_round(self, d_req: int):
        """Eager round with per-kernel-class device times (ms) and launch counts."""
        out = _lib.lsk_round_out()
        ms = (C.c_float * 8)()
        cnt = (C.c_int64 * 8)()
        total = C.c_float()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.lsk_profile_round(self._h, d_req, C.byref(out), ms, cnt,
                                                   C.byref(total)))
        r = RoundOutput(n_drafted=out.n_drafted, n_matches=out.n_matches,
                        emitted=list(out.emitted_ids[:out.n_emitted]),
                        draft=list(out.draft_ids[:out.n_drafted]),
                        verified=list(out.verified_ids[:out.n_drafted + 1]), kv_len=out.kv_len)
        return r, dict(zip(self.KERNEL_CLASSES, [float(x) for x in ms])), \
            dict(zip(self.KERNEL_CLASSES, [int(x) for x in cnt])), float(total.value)

    def ar_step(self) -> int:
        tok = C.c_int32()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.lsk_ar_step(self._h, C.byref(tok)))
        return tok.value

    # ------------------------------------------------------------------ introspection
    @property
    def kv_len(self) -> int:
        v = C.c_int32()
        _lib.check(self._lib.lsk_kv_len(self._h, C.byref(v)))
        return v.value

    @property
    def launch_count(self) -> int:
        v = C.c_int64()
        _lib.check(self._lib.lsk_launch_count(self._h, C.byref(v)))
        return v.value

    @property
    def last_device_ms(self) -> float:
        v = C.c_float()
        _lib.check(self._lib.lsk_last_device_ms(self._h, C.byref(v)))
        return v.value

    def round_bytes(self, d: int, ctx: int) -> float:
        v = C.c_double()
        _lib.check(self._lib.lsk_round_bytes(self._h, d, ctx, C.byref(v)))
        return v.value

    def ar_bytes(self, ctx: int) -> float:
        v = C.c_double()
        _lib.check(self._lib.lsk_ar_bytes(self._h, ctx, C.byref(v)))
        return v.value

    def debug_forward_rows(self, ids: Sequence[int]) -> torch.Tensor:
        """Teacher-forced block (parity tests): logits [len(ids), vocab_local] of the given ids
        run as ONE block on top of the committed context; nothing is committed."""
        m = len(ids)
        arr = (C.c_int32 * m)(*[int(t) for t in ids])
        with torch.cuda.device(self.device):
            _lib.check(self._lib.lsk_debug_forward_rows(self._h, arr, m))
        return self.debug_logits(m)

    def debug_hidden(self, rows: int = 16) -> torch.Tensor:
        n = rows * self.arch.hidden
        buf = (C.c_float * n)()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.lsk_debug_read(self._h, _lib.LSK_DBG_HIDDEN, 0, 0, buf, n))
        return torch.tensor(list(buf), dtype=torch.float32).view(rows, self.arch.hidden)

    def debug_logits(self, rows: int = 16) -> torch.Tensor:
        vloc = self.arch.vocab // self.tp_size
        vpad = (vloc + 15) // 16 * 16
        n = rows * vpad
        buf = (C.c_float * n)()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.lsk_debug_read(self._h, _lib.LSK_DBG_LOGITS, 0, 0, buf, n))
        return torch.frombuffer(buf, dtype=torch.float32).clone().view(rows, vpad)[:, :vloc]

    def debug_probs(self, which: str, rows: int = 16) -> torch.Tensor:
        """Warped sampling distributions of the last round: 'draft' or 'verify' -> [rows, vocab]."""
        n = rows * self.arch.vocab
        buf = (C.c_float * n)()
        what = _lib.LSK_DBG_PROBS_DRAFT if which == "draft" else _lib.LSK_DBG_PROBS_VERIFY
        with torch.cuda.device(self.device):
            _lib.check(self._lib.lsk_debug_read(self._h, what, 0, 0, buf, n))
        return torch.frombuffer(buf, dtype=torch.float32).clone().view(rows, self.arch.vocab)

    def debug_residual(self) -> torch.Tensor:
        """max(p_verify - p_draft, 0) of the last rejected draft position (unnormalised) [vocab]."""
        n = self.arch.vocab
        buf = (C.c_float * n)()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.lsk_debug_read(self._h, _lib.LSK_DBG_RESIDUAL, 0, 0, buf, n))
        return torch.frombuffer(buf, dtype=torch.float32).clone()

    def debug_kv_row(self, which: str, layer: int, kv_head: int, pos: int) -> torch.Tensor:
        hd = self.arch.head_dim
        buf = (C.c_float * hd)()
        what = _lib.LSK_DBG_KROW if which == "k" else _lib.LSK_DBG_VROW
        with torch.cuda.device(self.device):
            _lib.check(self._lib.lsk_debug_read(self._h, what, layer, kv_head * self.max_ctx + pos,
                                                buf, hd))
        return torch.tensor(list(buf), dtype=torch.float32)

    def debug_set_page_table(self, pages: Sequence[int]) -> None:
        arr = (C.c_int32 * len(pages))(*[int(p) for p in pages])
        with torch.cuda.device(self.device):
            _lib.check(self._lib.lsk_debug_set_page_table(self._h, arr, len(pages)))
