"""GPU: the streaming GEMM kernel against a plain fp32 torch matmul of the same bf16 values."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _gemm(lib, w, x, iters=0):
    from layerskip_b200 import _lib
    n, k = w.shape
    m = x.shape[0]
    packed = torch.empty_like(w)
    _lib.check(lib.lsk_test_pack(w.data_ptr(), n, k, packed.data_ptr()))
    y = torch.zeros(m, n, dtype=torch.float32, device="cuda")
    ms = C.c_float(0)
    _lib.check(lib.lsk_test_gemm(packed.data_ptr(), n, k, x.data_ptr(), m, y.data_ptr(), iters,
                                 C.byref(ms)))
    torch.cuda.synchronize()
    return y, ms.value


@pytest.mark.parametrize("n,k", [(64, 256), (512, 704), (4096, 4096), (1024, 11008), (32000, 4096)])
@pytest.mark.parametrize("m", [1, 7, 8, 9, 16])
def test_skinny_gemm_matches_torch(n, k, m):
    from layerskip_b200 import _lib
    lib = _lib.load()
    if m > 8 and k > 8192:
        pytest.skip("16-row blocks do not fit next to K > 8192 (engine splits them)")
    g = torch.Generator(device="cuda").manual_seed(n * 31 + k + m)
    w = (torch.randn(n, k, generator=g, device="cuda") * 0.02).to(torch.bfloat16)
    x = torch.randn(m, k, generator=g, device="cuda").to(torch.bfloat16)
    y, _ = _gemm(lib, w, x)
    ref = x.float() @ w.float().T
    torch.testing.assert_close(y, ref, rtol=1e-3, atol=2e-4 * (k ** 0.5) * 0.02 * 4)


def test_skinny_gemm_is_batch_invariant():
    """Row j of an m-row block is bit-identical to the same row run alone."""
    from layerskip_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(5)
    w = (torch.randn(2048, 4096, generator=g, device="cuda") * 0.02).to(torch.bfloat16)
    x = torch.randn(16, 4096, generator=g, device="cuda").to(torch.bfloat16)
    y16, _ = _gemm(lib, w, x)
    y7, _ = _gemm(lib, w, x[:7].contiguous())
    assert torch.equal(y16[:7], y7)
    for j in (0, 3, 6):
        y1, _ = _gemm(lib, w, x[j:j + 1].contiguous())
        assert torch.equal(y1[0], y7[j])
    y1, _ = _gemm(lib, w, x[12:13].contiguous())
    assert torch.equal(y1[0], y16[12])
