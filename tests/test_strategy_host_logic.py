"""CPU: the host side of the strategies (the reference's OUTER loop: max_steps clamp, acceptance
accounting, EOS truncation, streamer / stopping-criteria hand-off) replayed against the golden
traces of the unmodified reference with a scripted engine in place of the GPU."""
import pytest
import torch

from layerskip_b200 import GenerationConfig
from layerskip_b200.engine import RoundOutput
from layerskip_b200.plugin import HuggingfaceLlamaGenerator
from layerskip_b200.strategy import (B200AutoRegressiveGenerationStrategy,
                                     B200SelfSpeculativeGenerationStrategy)
from layerskip_b200.synthetic import IntegerTokenizer, synthetic_prompts
from layerskip_b200.weights import ARCHS, LlamaArch, classify
from tests import golden_util as gu


class ScriptedEngine:
    """Replays the reference's recorded rounds; checks what the strategy asks for."""

    def __init__(self, case):
        self.case = case
        self.rounds = list(case["reference"]["rounds"])
        self.i = 0
        self.began = None
        self.prompt = None

    def begin(self, **kw):
        self.began = kw
        self.i = 0

    def prefill(self, ids):
        self.prompt = list(ids)

    def round(self, d_req):
        rec = self.rounds[self.i]
        self.i += 1
        assert d_req == rec["d_req"], "max_steps clamp differs from the reference (:63-66)"
        n = rec["n_matches"]
        return RoundOutput(n_drafted=rec["d_actual"], n_matches=n, emitted=list(rec["emitted"]),
                           draft=list(rec["emitted"][:n]) + [0] * (rec["d_actual"] - n),
                           verified=[0] * (rec["d_actual"] + 1), kv_len=rec["kv_len_after"])


def _strategy_with(engine):
    s = B200SelfSpeculativeGenerationStrategy.__new__(B200SelfSpeculativeGenerationStrategy)
    s.engines = type("Cache", (), {"get": lambda self, model: engine})()
    s.last_rounds = []
    return s


@pytest.mark.parametrize("case", gu.spec_cases(greedy=True), ids=lambda c: c["name"])
def test_outer_loop_matches_reference(case):
    eng = ScriptedEngine(case)
    strat = _strategy_with(eng)
    res = strat.generate_token_ids(object(), case["prompt"], case["eos"],
                                   GenerationConfig(**case["cfg"]))
    ref = case["reference"]
    assert res.predicted_tokens == ref["spec_tokens"]
    assert res.acceptance_rate == pytest.approx(ref["acceptance_rate"], abs=1e-12)
    assert eng.i == len(ref["rounds"])                     # stopped exactly where the reference did
    assert eng.prompt == case["prompt"]
    assert eng.began["exit_layer"] == case["cfg"]["exit_layer"]


def test_streamers_receive_what_the_reference_sends():
    case = next(c for c in gu.spec_cases(greedy=True) if c["name"] == "mha128_a0.1")

    class Plain:
        def __init__(self):
            self.got = []

        def put(self, t):
            self.got += t.tolist()

    class Speculative(Plain):
        def __init__(self):
            super().__init__()
            self.deleted = 0
            self.drafts = 0

        def put(self, t, is_draft=False):
            if is_draft:
                self.drafts += t.numel()
            else:
                self.got += t.tolist()

        def delete(self, n):
            self.deleted += n

    flat = [t for r in case["reference"]["rounds"] for t in r["emitted"]]
    p = Plain()
    _strategy_with(ScriptedEngine(case)).generate_token_ids(
        object(), case["prompt"], case["eos"], GenerationConfig(**case["cfg"]), streamer=p)
    assert p.got == flat
    s = Speculative()
    _strategy_with(ScriptedEngine(case)).generate_token_ids(
        object(), case["prompt"], case["eos"], GenerationConfig(**case["cfg"]), streamer=s)
    assert s.got == flat
    assert s.drafts == s.deleted == sum(r["d_actual"] for r in case["reference"]["rounds"])


def test_stopping_criteria_and_unsupported_processors():
    case = next(c for c in gu.spec_cases(greedy=True) if c["name"] == "mha128_a0.1")
    seen = []

    def stop(ids, scores=None):          # bare callable, like tests/test_autoregressive_generator.py:43
        seen.append(ids.tolist())
        return torch.tensor([True])

    res = _strategy_with(ScriptedEngine(case)).generate_token_ids(
        object(), case["prompt"], case["eos"], GenerationConfig(**case["cfg"]),
        stopping_criteria=stop)
    first = case["reference"]["rounds"][0]["emitted"]
    assert res.predicted_tokens == first and seen == [[[first[-1]]]]   # 1x1 NEXT token (:94)
    with pytest.raises(NotImplementedError):
        _strategy_with(ScriptedEngine(case)).generate_token_ids(
            object(), case["prompt"], case["eos"], GenerationConfig(**case["cfg"]),
            logits_processors=[object()])


def test_autoregressive_host_loop_eos_before_append():
    toks = iter([5, 6, 7, 99, 8])

    class Eng:
        def begin(self, **kw): pass
        def prefill(self, ids): pass
        def ar_step(self): return next(toks)

    ar = B200AutoRegressiveGenerationStrategy.__new__(B200AutoRegressiveGenerationStrategy)
    ar.engines = type("Cache", (), {"get": lambda self, m: Eng()})()
    res = ar.generate_token_ids(object(), [1, 2], [99], GenerationConfig(max_steps=10, sample=False))
    assert res.predicted_tokens == [5, 6, 7] and res.acceptance_rate is None   # :66-67


def test_generator_facade_and_integer_tokenizer():
    case = next(c for c in gu.spec_cases(greedy=True) if c["name"] == "mha128_a0.1")
    tok = IntegerTokenizer(512)
    gen = HuggingfaceLlamaGenerator(tok, object(), _strategy_with(ScriptedEngine(case)))
    cfg = GenerationConfig(**case["cfg"])
    out = gen.generate(" ".join(map(str, case["prompt"])), cfg)
    assert out.generation_strategy_result.predicted_tokens == case["reference"]["spec_tokens"]
    assert out.decoded_prediction == " ".join(map(str, case["reference"]["spec_tokens"]))
    assert out.num_tokens_generated == len(case["reference"]["spec_tokens"])
    assert out.tokens_per_second > 0 and out.total_time > 0


def test_generation_config_defaults_match_the_reference():
    c = GenerationConfig()          # generator_base.py:33-49
    assert (c.max_steps, c.exit_layer, c.num_speculations, c.generation_strategy, c.sample,
            c.temperature, c.top_k, c.top_p, c.no_repeat_ngram_size, c.stop_words,
            c.stop_token_ids) == (512, -1, -1, "autoregressive", True, 0.6, 0, 0.9, None, None, [])


def test_weight_name_classification_and_arch_table():
    from layerskip_b200 import _lib
    assert classify("model.layers.7.self_attn.q_proj.weight") == (_lib.LSK_W_Q, 7)
    assert classify("model.layers.0.mlp.down_proj.weight") == (_lib.LSK_W_DOWN, 0)
    assert classify("lm_head.weight") == (_lib.LSK_W_LM_HEAD, 0)
    assert classify("model.layers.0.self_attn.rotary_emb.inv_freq") is None
    a = ARCHS["llama2-7b"]
    assert isinstance(a, LlamaArch) and abs(a.param_bytes() / 1e9 - 13.48) < 0.01
    assert ARCHS["llama2-70b"].kv_heads == 8 and ARCHS["llama3-8b"].vocab == 128256
    prompts = synthetic_prompts(32000, 8, 128)
    assert len(prompts) == 8 and all(len(p) == 128 and 3 <= min(p) and max(p) <= 31998 for p in prompts)


def test_context_overflow_is_reported_before_any_kernel_runs():
    from layerskip_b200.plugin import GenerationConfig
    from layerskip_b200.strategy import _check_context
    eng = type("E", (), dict(max_ctx=256))()
    _check_context(eng, 100, GenerationConfig(max_steps=155))                 # 100 + 155 + 1 == 256
    with pytest.raises(ValueError, match="max_ctx=256"):
        _check_context(eng, 100, GenerationConfig(max_steps=156))
    _check_context(object(), 10 ** 6, GenerationConfig(max_steps=512))        # engines without the attribute


def test_logits_processor_list_is_mapped_to_the_device_ngram_ban():
    """generator_base.py:77-85 builds at most one processor (NoRepeatNGram); the strategy turns it
    into the engine's `no_repeat_ngram_size` and refuses anything it cannot run on the device."""
    from transformers.generation.logits_process import (LogitsProcessorList, NoRepeatNGramLogitsProcessor,
                                                         TemperatureLogitsWarper)
    from layerskip_b200.strategy import _ngram_size_of
    assert _ngram_size_of(None) == 0 and _ngram_size_of([]) == 0 and _ngram_size_of(LogitsProcessorList()) == 0
    assert _ngram_size_of(LogitsProcessorList([NoRepeatNGramLogitsProcessor(3)])) == 3
    with pytest.raises(NotImplementedError):
        _ngram_size_of(LogitsProcessorList([TemperatureLogitsWarper(0.7)]))
    with pytest.raises(NotImplementedError):
        _ngram_size_of([lambda ids, scores: scores])
    with pytest.raises(NotImplementedError):
        _ngram_size_of([NoRepeatNGramLogitsProcessor(17)])
