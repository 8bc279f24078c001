"""GPU tests of code paths that were written after the round-1 GPU budget ran out (never
executed yet).  They are opt-in so that a defect in an experimental path cannot mask the
validated suite:  LSK_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gpu_zz_experimental.py"""
import os

import pytest
import torch

from tests import golden_util as gu

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.environ.get("LSK_TEST_EXPERIMENTAL"),
                                 reason="experimental paths: enable with LSK_TEST_EXPERIMENTAL=1")]


def _generate(case, **engine_kwargs):
    from layerskip_b200 import GenerationConfig
    from layerskip_b200.strategy import B200SelfSpeculativeGenerationStrategy
    from tests.test_gpu_engine import _Model
    dims, sd = gu.state_dict_for(case)
    strat = B200SelfSpeculativeGenerationStrategy(max_ctx=512, **engine_kwargs)
    try:
        r = strat.generate_token_ids(_Model(dims, sd), case["prompt"], case["eos"],
                                     GenerationConfig(**case["cfg"]))
        rounds = [(x.n_drafted, x.n_matches, tuple(x.emitted)) for x in strat.last_rounds]
    finally:
        strat.engines.close()
    return r.predicted_tokens, r.acceptance_rate, rounds


@pytest.mark.parametrize("name", ["gqa128_a0.1", "mha128_a0.1", "gqa128_a0.05_long"])
def test_push_merge_attention_is_bit_identical_to_the_pull_merge_kernel(name, monkeypatch):
    """attn_cluster_push_kernel: same partials, same merge arithmetic and order — only the
    direction of the distributed-shared-memory exchange differs, so every round must be equal."""
    case = next(c for c in gu.spec_cases() if c["name"] == name)
    monkeypatch.delenv("LSK_ATTN_PUSH", raising=False)
    want = _generate(case)
    monkeypatch.setenv("LSK_ATTN_PUSH", "1")
    got = _generate(case)
    assert got == want
