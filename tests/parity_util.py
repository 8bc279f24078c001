"""Margin-gated parity protocol shared by the GPU tests.

A bf16 engine cannot be token-identical to an fp32 oracle on random-init weights over long
runs: near-tie arg-maxes flip.  A mismatch at position j is BENIGN iff, under the oracle's own
teacher-forced logits at j, the engine's token is within `tau` of the oracle's best logit;
anything else is a failure.  After a benign flip the comparison restarts from the oracle's
prefix (teacher forcing) so the rest of the stream is still checked.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import torch

from oracle import llama_oracle as orc

TAU = 0.06   # logit units; 2 x the max |delta logit| we measure for bf16 activations (tests assert it)


def check_stream(w: orc.OracleWeights, prompt: Sequence[int], oracle_tokens: Sequence[int],
                 generate: Callable[[List[int], int], List[int]], exit_layer: int = -1,
                 tau: float = TAU) -> Tuple[int, List[float]]:
    """`generate(prompt_ids, n_tokens)` must return the engine's continuation.
    Returns (number of benign flips, their oracle margins); raises AssertionError on a real one."""
    oracle_tokens = list(oracle_tokens)
    flips: List[float] = []
    start = 0
    while start < len(oracle_tokens):
        want = oracle_tokens[start:]
        got = generate(list(prompt) + oracle_tokens[:start], len(want))
        j = 0
        while j < len(want) and j < len(got) and got[j] == want[j]:
            j += 1
        if j == len(want):
            break
        assert j < len(got), f"engine stopped early at {start + j} (got {len(got)} of {len(want)})"
        # oracle distribution at the diverging position
        if exit_layer > 0:
            logits = orc.early_exit_logits(w, prompt, oracle_tokens[:start + j + 1], exit_layer)
        else:
            logits = orc.teacher_forced_logits(w, prompt, oracle_tokens[:start + j + 1])
        row = logits[start + j]
        gap = float(row[want[j]] - row[got[j]])
        assert int(row.argmax()) == want[j]
        assert gap < tau, (f"token {start + j}: engine chose {got[j]}, oracle {want[j]} with logit "
                           f"gap {gap:.4f} >= tau {tau}")
        flips.append(gap)
        start = start + j + 1
    return len(flips), flips


def usable_cpus() -> int:
    """CPUs this process may really use (affinity mask capped by the cgroup quota): a GPU box can
    show 128 cores to a container throttled to 16, and 128 OpenMP threads are then far slower."""
    import os
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        txt = open("/sys/fs/cgroup/cpu.max").read().split()
        if txt[0] != "max":
            n = min(n, max(1, int(int(txt[0]) / int(txt[1]))))
    except Exception:
        pass
    return max(1, n)


def set_oracle_threads() -> int:
    n = min(32, usable_cpus())
    torch.set_num_threads(n)
    return n
