"""Engine acceptance statistics over many prompts (diagnostic for tests/test_gpu_sampling.py)."""
import json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from layerskip_b200 import GenerationConfig
from layerskip_b200.strategy import B200SelfSpeculativeGenerationStrategy
from tests import golden_util as gu
from tests.test_gpu_engine import _Model
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
case = next(c for c in gu.spec_cases() if c["name"] == "gqa128_sample_s3")
dims, sd = gu.state_dict_for(case)
model = _Model(dims, sd)
strat = B200SelfSpeculativeGenerationStrategy(max_ctx=512)
g = torch.Generator().manual_seed(2024)
prompts = torch.randint(3, dims.vocab - 1, (256, 12), generator=g).tolist()
cfg = GenerationConfig(max_steps=128, exit_layer=3, num_speculations=6, sample=True, temperature=0.6, top_k=0, top_p=0.9)
rounds = []
for i, p in enumerate(prompts[:n]):
    torch.manual_seed(100 + i)
    strat.generate_token_ids(model, p, [dims.vocab - 1], cfg)
    rounds += [(r.n_matches, r.n_drafted) for r in strat.last_rounds]
m = sum(a for a, b in rounds); d = sum(b for a, b in rounds); p = m / d
var = sum((a - p * b) ** 2 for a, b in rounds) / d ** 2
print(json.dumps(dict(p=p, d=d, se_binom=math.sqrt(p * (1 - p) / d), se_robust=math.sqrt(var))))
