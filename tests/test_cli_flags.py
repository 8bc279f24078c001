"""CPU: the command lines parse the reference's flags (README.md:54-156 of the reference)."""
import sys

from layerskip_b200 import cli
from layerskip_b200.plugin import GenerationConfig


def test_reference_benchmark_flags_parse(monkeypatch):
    monkeypatch.setattr(sys, "argv", ["benchmark.py", "--model", "synthetic:llama2-7b", "--dataset", "synthetic",
                                      "--num_samples", "4", "--generation_strategy", "self_speculative",
                                      "--exit_layer", "8", "--num_speculations", "6", "--max_steps", "512",
                                      "--sample", "False", "--output_dir", "./logs",
                                      "--model_args", "alpha=0.1,seed=3,max_ctx=1024"])
    args, bargs, gcfg = cli.parse(cli.Arguments, cli.BenchmarkArguments, GenerationConfig)
    assert (gcfg.exit_layer, gcfg.num_speculations, gcfg.max_steps, gcfg.sample) == (8, 6, 512, False)
    assert gcfg.generation_strategy == "self_speculative" and bargs.num_samples == 4
    assert cli.parse_model_args(args.model_args) == {"alpha": 0.1, "seed": 3, "max_ctx": 1024}


def test_sweep_flags_and_defaults(monkeypatch):
    monkeypatch.setattr(sys, "argv", ["sweep.py", "--model", "synthetic:tiny-gqa", "--exit_layer_first", "1",
                                      "--exit_layer_last", "3", "--num_speculations_last", "4"])
    _a, _b, sargs, gcfg = cli.parse(cli.Arguments, cli.BenchmarkArguments, cli.SweepArguments, GenerationConfig)
    assert (sargs.exit_layer_first, sargs.exit_layer_last, sargs.exit_layer_step) == (1, 3, 1)
    assert (sargs.num_speculations_first, sargs.num_speculations_last) == (1, 4)
    assert gcfg.temperature == 0.6 and gcfg.top_p == 0.9          # generator_base.py:39-42


def test_model_args_string_parser():
    assert cli.parse_model_args("a=1,b=true,c=x,d=0.5") == {"a": 1, "b": True, "c": "x", "d": 0.5}
    assert cli.parse_model_args(None) == {} and cli.parse_model_args("") == {}


def test_model_flag_accepts_a_checkpoint_directory(tmp_path):
    """`--model <dir>` (reference: generate.py:54-67 loads with from_pretrained) resolves to the
    streaming checkpoint reader + the integer tokenizer when the directory has no tokenizer files;
    nothing touches a GPU until an engine is built."""
    import torch
    from layerskip_b200.checkpoint import CheckpointLlama, expected_shapes, save_checkpoint
    from layerskip_b200.synthetic import IntegerTokenizer
    from layerskip_b200.weights import ARCHS
    arch = ARCHS["tiny-mha"]
    g = torch.Generator().manual_seed(0)
    tensors = ((n, (torch.randn(s, generator=g) * 0.02).to(torch.bfloat16))
               for n, s in expected_shapes(arch).items())
    save_checkpoint(str(tmp_path), arch, tensors, max_shard_bytes=1 << 20)
    model, tok, margs = cli.load_model_and_tokenizer(
        cli.Arguments(model=str(tmp_path), model_args="max_ctx=256"), exit_layer=2)
    assert isinstance(model, CheckpointLlama) and model.arch == arch
    assert isinstance(tok, IntegerTokenizer) and margs == {"max_ctx": 256}
    assert model.config.num_hidden_layers == arch.layers       # what the strategies read
    syn, tok2, _ = cli.load_model_and_tokenizer(cli.Arguments(model="synthetic:tiny-mha"), exit_layer=2)
    assert syn.arch == arch and isinstance(tok2, IntegerTokenizer)
