"""CPU: the shared library loads without a GPU and exports every symbol include/lsk.h declares;
the ctypes table (layerskip_b200/_lib.py) covers exactly that list."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "lsk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lsk_[a-z_0-9]+)\s*\(", text)))


def test_header_and_ctypes_table_agree():
    from layerskip_b200 import _lib
    assert _header_symbols() == sorted(_lib.SIGNATURES)


def test_library_exports_every_symbol():
    from layerskip_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    lib = _lib.load()
    for name in _header_symbols():
        assert hasattr(lib, name), name
    assert lib.lsk_abi_version() == 2
    assert isinstance(lib.lsk_last_error(), bytes)


def test_struct_sizes_match_the_header():
    import ctypes as C
    from layerskip_b200 import _lib
    assert C.sizeof(_lib.lsk_config) == 19 * 4
    assert C.sizeof(_lib.lsk_round_out) == 4 * 4 + 3 * 16 * 4
    assert C.sizeof(_lib.lsk_weight_desc) == 32
    assert C.sizeof(_lib.lsk_generation) == 72      # 60 bytes of 32-bit fields, pad to 8, uint64 seed


def test_engine_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from layerskip_b200.engine import Engine
    from layerskip_b200.weights import ARCHS
    with pytest.raises(RuntimeError, match="no CPU path"):
        Engine(ARCHS["tiny-mha"])


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "layerskip_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn


def test_flag_constants_match_the_header():
    import re
    from layerskip_b200 import _lib
    text = open(os.path.join(os.path.dirname(__file__), "..", "include", "lsk.h")).read()
    flags = dict(re.findall(r"#define (LSK_FLAG_\w+) (\d+)u", text))
    assert set(flags) == {n for n in dir(_lib) if n.startswith("LSK_FLAG_")}
    for name, value in flags.items():
        assert getattr(_lib, name) == int(value), name
    assert len(set(flags.values())) == len(flags)       # distinct bits


def test_create_fails_cleanly_without_a_gpu_and_validates_arguments():
    """No compute: `lsk_create` must turn a missing device / bad config into an error code and a
    message, never a crash or a half-built handle (a failed create releases what it allocated)."""
    import ctypes as C
    import torch
    from layerskip_b200 import _lib
    lib = _lib.load()

    def cfg(**over):
        base = dict(vocab=512, hidden=256, inter=704, n_layers=2, n_heads=2, n_kv_heads=2,
                    head_dim=128, rms_eps=1e-5, rope_theta=1e4, max_ctx=128, tp_rank=0, tp_size=1,
                    attn_splits=0, flags=0, rope_scaling=0, rope_factor=1.0)
        base.update(over)
        return _lib.lsk_config(**base)

    for bad, needle in ((dict(head_dim=96), "head_dim"), (dict(tp_rank=2, tp_size=2), "tp_rank"),
                        (dict(n_heads=3), "heads"), (dict(hidden=8200), "hidden"),
                        (dict(inter=700), "intermediate"), (dict(rope_scaling=2, rope_factor=0.0), "rope"), (dict(max_ctx=1), "max_ctx"),
                        (dict(vocab=511, tp_size=2), "vocab")):
        h = C.c_void_p()
        c = cfg(**bad)
        assert lib.lsk_create(C.byref(c), C.byref(h)) == -1, bad       # LSK_ERR_INVALID
        assert needle in lib.lsk_last_error().decode(), bad
        assert not h.value
    assert lib.lsk_create(None, None) == -1
    if not torch.cuda.is_available():
        h = C.c_void_p()
        c = cfg()
        assert lib.lsk_create(C.byref(c), C.byref(h)) == -2            # LSK_ERR_CUDA
        assert "cuda" in lib.lsk_last_error().decode().lower()
        assert not h.value
    lib.lsk_destroy(None)                                              # tolerated
