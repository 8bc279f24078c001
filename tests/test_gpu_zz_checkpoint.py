"""GPU: an engine fed from a checkpoint directory (streamed shards) must hold exactly the weights
of an engine fed from the equivalent state dict — same tokens, same acceptance, greedy spec."""
import pytest
import torch

from layerskip_b200.checkpoint import CheckpointLlama, expected_shapes, save_checkpoint
from layerskip_b200.weights import ARCHS

pytestmark = pytest.mark.gpu


class _Model:
    def __init__(self, arch, sd):
        self._sd = sd
        self.arch = arch
        self.config = None

    def state_dict(self):
        return self._sd


def _tensors(arch, seed, tied):
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in expected_shapes(arch, tied).items():
        if len(shape) == 1:
            out[name] = torch.ones(shape, dtype=torch.bfloat16)
        else:
            out[name] = (torch.randn(shape, generator=g) * 0.02).to(torch.bfloat16)
    return out


@pytest.mark.parametrize("tied", [False, True])
def test_checkpoint_engine_equals_state_dict_engine(tmp_path, tied):
    from layerskip_b200 import GenerationConfig
    from layerskip_b200.strategy import B200SelfSpeculativeGenerationStrategy
    arch = ARCHS["tiny-gqa"]
    sd = _tensors(arch, 21, tied)
    save_checkpoint(str(tmp_path), arch, sd.items(), max_shard_bytes=3 << 20,
                    tie_word_embeddings=tied)
    full = dict(sd)
    if tied:
        full["lm_head.weight"] = sd["model.embed_tokens.weight"]
    cfg = GenerationConfig(max_steps=48, exit_layer=2, num_speculations=4,
                           generation_strategy="self_speculative", sample=False)
    g = torch.Generator().manual_seed(5)
    prompt = torch.randint(3, arch.vocab - 1, (24,), generator=g).tolist()
    outs = []
    for model in (_Model(arch, full), CheckpointLlama(str(tmp_path))):
        strat = B200SelfSpeculativeGenerationStrategy(max_ctx=256)
        try:
            r = strat.generate_token_ids(model, prompt, [arch.vocab - 1], cfg)
        finally:
            strat.engines.close()
        outs.append((r.predicted_tokens, r.acceptance_rate))
    assert outs[0] == outs[1]
    assert 0 < len(outs[0][0]) <= 48
