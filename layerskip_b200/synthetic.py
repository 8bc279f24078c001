"""Offline stand-ins for what the reference downloads: an integer tokenizer and seeded prompts
(SURVEY.md §8(b) "tokenizer for offline runs", §8(d) "synthetic prompts")."""
from __future__ import annotations

from typing import List

import torch


class IntegerTokenizer:
    """`"17 942 3"` <-> [17, 942, 3].  Implements the three things
    `HuggingfaceLlamaGenerator` uses (generator_base.py:103,106,119-121)."""

    def __init__(self, vocab_size: int, eos_token_id: int = None):
        self.vocab_size = vocab_size
        self.eos_token_id = vocab_size - 1 if eos_token_id is None else eos_token_id

    def __call__(self, text: str, return_tensors: str = "pt", add_special_tokens: bool = True):
        ids = [int(t) for t in text.split()]
        return {"input_ids": torch.tensor([ids], dtype=torch.long)}

    def decode(self, ids, **_kw) -> str:
        if hasattr(ids, "tolist"):
            ids = ids.tolist()
        return " ".join(str(int(t)) for t in ids)


def synthetic_prompts(vocab_size: int, n_prompts: int = 8, prompt_len: int = 128,
                      seed: int = 1234) -> List[List[int]]:
    """`n_prompts` x `prompt_len` ids uniform in [3, vocab-2]."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(3, vocab_size - 1, (n_prompts, prompt_len), generator=g)
    return [row.tolist() for row in ids]
