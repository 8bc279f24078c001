"""GPU: the paged split-KV attention kernel alone (through the C ABI: `lsk_test_attn`) against a
plain fp32 torch restatement of HF's eager attention (transformers modeling_llama.py:187-221:
scores * head_dim^-0.5 + causal mask, fp32 softmax, probabilities rounded to bf16, P.V) on
BASELINE head layouts: 32 heads x 32 kv (Llama-2-7B), 32 x 8 (Llama-3-8B), 40 x 40 (13B), 64 x 8
(70B), head_dim 64 (llama3.2-1B) and 32 (correctness.py's tiny model); contexts that give a split
1, 2, 3 and more key groups; 1 / 7 / 9 / 16 query rows; permuted page tables."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _reference(q, k, v, n_heads, n_kv, hd, ctx, m):
    group = n_heads // n_kv
    qf = q.float().view(m, n_heads, hd).transpose(0, 1)                 # [H, m, hd]
    kf = k.float().repeat_interleave(group, 0)                          # [H, ctx, hd]
    vf = v.float().repeat_interleave(group, 0)
    scores = (qf @ kf.transpose(1, 2)) * hd ** -0.5
    pos = torch.arange(ctx - m, ctx, device=q.device)[:, None]
    scores = scores.masked_fill(torch.arange(ctx, device=q.device)[None, :] > pos, float("-inf"))
    probs = torch.softmax(scores, -1).to(torch.bfloat16).float()
    return (probs @ vf).transpose(0, 1).reshape(m, n_heads * hd)


def _run(n_heads, n_kv, hd, ctx, m, splits=8, perm=False, iters=0, seed=0):
    from layerskip_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(seed + ctx * 31 + m)
    q = torch.randn(m, n_heads * hd, generator=g, device="cuda").to(torch.bfloat16)
    k = torch.randn(n_kv, ctx, hd, generator=g, device="cuda").to(torch.bfloat16)
    v = torch.randn(n_kv, ctx, hd, generator=g, device="cuda").to(torch.bfloat16)
    out = torch.full((m, n_heads * hd), float("nan"), device="cuda", dtype=torch.bfloat16)
    n_pages = (ctx + 63) // 64
    pp = None
    if perm:
        order = torch.randperm(n_pages, generator=torch.Generator().manual_seed(seed)).tolist()
        pp = (C.c_int32 * n_pages)(*order)
    ms = C.c_float(0)
    torch.cuda.synchronize()
    _lib.check(lib.lsk_test_attn(q.data_ptr(), k.data_ptr(), v.data_ptr(), n_heads, n_kv, hd, ctx, m,
                                 splits, pp, out.data_ptr(), iters, C.byref(ms)))
    torch.cuda.synchronize()
    return out.float(), _reference(q, k, v, n_heads, n_kv, hd, ctx, m), ms.value


CASES = [
    # heads, kv, hd, ctx, m
    (32, 32, 128, 70, 1), (32, 32, 128, 70, 7), (32, 32, 128, 520, 1), (32, 32, 128, 520, 9),
    (32, 32, 128, 640, 7), (32, 32, 128, 1100, 7), (32, 32, 128, 1100, 16),
    (32, 8, 128, 520, 7), (32, 8, 128, 1100, 9), (32, 8, 128, 70, 1),
    (40, 40, 128, 1100, 7), (64, 8, 128, 520, 7), (64, 8, 128, 1100, 1),
    (32, 8, 64, 700, 9), (32, 8, 64, 70, 1), (8, 8, 32, 70, 5), (8, 8, 32, 200, 16),
    (2, 2, 128, 63, 1), (2, 2, 128, 64, 1), (2, 2, 128, 65, 2), (4, 2, 128, 9, 9),
]


@pytest.mark.parametrize("n_heads,n_kv,hd,ctx,m", CASES)
def test_attention_matches_fp32_reference(n_heads, n_kv, hd, ctx, m):
    got, want, _ = _run(n_heads, n_kv, hd, ctx, m)
    assert torch.isfinite(got).all()
    torch.testing.assert_close(got, want, rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("n_heads,n_kv,hd,ctx,m,splits", [
    (32, 32, 128, 2100, 7, 8),     # 33 key groups: 4-5 per split, the K/V ring wraps
    (32, 8, 128, 2100, 9, 8),      # ... with two row blocks re-streaming the ring
    (32, 32, 128, 1100, 7, 4), (32, 32, 128, 520, 7, 1), (32, 8, 64, 1500, 16, 2),
])
def test_long_contexts_fewer_splits_and_permuted_pages(n_heads, n_kv, hd, ctx, m, splits):
    got, want, _ = _run(n_heads, n_kv, hd, ctx, m, splits=splits, perm=True, seed=5)
    torch.testing.assert_close(got, want, rtol=2e-2, atol=2e-2)


def test_rows_are_batch_invariant():
    """A query row computed alone is bit-identical to the same row inside a 7-row block (the
    property `speculative == autoregressive` rests on): same keys, same partition, same order."""
    from layerskip_b200 import _lib
    lib = _lib.load()
    n_heads, n_kv, hd, ctx, m = 32, 8, 128, 700, 7
    g = torch.Generator(device="cuda").manual_seed(3)
    q = torch.randn(m, n_heads * hd, generator=g, device="cuda").to(torch.bfloat16)
    k = torch.randn(n_kv, ctx, hd, generator=g, device="cuda").to(torch.bfloat16)
    v = torch.randn(n_kv, ctx, hd, generator=g, device="cuda").to(torch.bfloat16)
    out = torch.zeros(m, n_heads * hd, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.lsk_test_attn(q.data_ptr(), k.data_ptr(), v.data_ptr(), n_heads, n_kv, hd, ctx, m, 8,
                                 None, out.data_ptr(), 0, None))
    for row in (0, 3, 6):
        c1 = ctx - m + row + 1                       # keys visible to that row
        one = torch.zeros(1, n_heads * hd, device="cuda", dtype=torch.bfloat16)
        q1 = q[row:row + 1].contiguous()
        k1, v1 = k[:, :c1].contiguous(), v[:, :c1].contiguous()
        _lib.check(lib.lsk_test_attn(q1.data_ptr(), k1.data_ptr(), v1.data_ptr(), n_heads, n_kv, hd, c1, 1, 8,
                                     None, one.data_ptr(), 0, None))
        torch.cuda.synchronize()
        assert torch.equal(one[0], out[row]), row


def test_attention_latency_at_the_bench_shape():
    """Llama-2-7B, ctx 640, 7 rows (the verify block of the headline config): report the latency
    (informative; the roofline for 4.3 MB of K/V is ~0.7 us, the kernel is latency-bound)."""
    for m in (1, 7):
        _, _, ms = _run(32, 32, 128, 640, m, iters=200)
        print(f"attention 7B ctx 640 m={m}: {ms * 1e3:.2f} us per launch (back-to-back launches)")
        assert ms < 0.05
