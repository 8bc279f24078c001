"""CPU model of the shared-memory operand layouts of csrc/lmhead_tc.cuh (tcgen05, SWIZZLE_NONE,
K-major).  The kernel itself cannot run here; what can be checked is that the three pieces of
address arithmetic agree with each other under the documented descriptor semantics
(cute::UMMA canonical K-major layout  ((8, n), 2) : ((16 B, SBO), LBO)  per K = 16 MMA):
  * pack_canonical_kernel's mapping  (tile, k stage, 16-byte chunk) -> (row, k),
  * the prologue's mapping           (token, k) -> byte offset of the B operand,
  * the descriptors the MMA issuer builds (start address advance, LBO, SBO).
For every MMA the model gathers the 128 x 16 A block and the 16 x 16 B block the hardware would
read and compares them with the source matrices."""
import numpy as np

TILE_ROWS, STAGE_K, STAGE_BYTES = 128, 64, 128 * 64 * 2


def pack_canonical(w):
    """numpy restatement of pack_canonical_kernel (uint16 = bf16 bit patterns)."""
    n, k = w.shape
    n_tiles, kst = (n + TILE_ROWS - 1) // TILE_ROWS, k // STAGE_K
    out = np.zeros(n_tiles * kst * STAGE_BYTES // 2, dtype=np.uint16)
    for c in range(n_tiles * kst * (STAGE_BYTES // 16)):
        blk, inn = divmod(c, STAGE_BYTES // 16)
        tile, s = divmod(blk, kst)
        core, r = inn >> 3, inn & 7
        i, j = core >> 3, core & 7
        row = tile * TILE_ROWS + i * 8 + r
        if row < n:
            out[c * 8:(c + 1) * 8] = w[row, s * STAGE_K + j * 8: s * STAGE_K + j * 8 + 8]
    return out


def write_b_operand(x):
    """numpy restatement of the prologue's store pattern: token m, 4 consecutive k per store."""
    m_rows, k = x.shape
    xb = np.zeros(16 * k, dtype=np.uint16)
    for m in range(16):
        for idx in range(k // 4):
            kk = idx * 4
            j = kk >> 3
            off = ((j * 2 + (m >> 3)) * 128 + (m & 7) * 16 + (kk & 7) * 2) // 2      # in uint16
            xb[off:off + 4] = x[m, kk:kk + 4] if m < m_rows else 0
    return xb


def umma_read(mem_u16, start_bytes, lbo, sbo, mn):
    """What one K = 16 MMA reads through a SWIZZLE_NONE K-major descriptor: [mn, 16] elements."""
    out = np.zeros((mn, 16), dtype=np.uint16)
    for r in range(mn):
        for kc in range(2):
            base = (start_bytes + (r // 8) * sbo + kc * lbo + (r % 8) * 16) // 2
            out[r, kc * 8:(kc + 1) * 8] = mem_u16[base:base + 8]
    return out


def test_descriptors_walk_exactly_the_source_matrices():
    rng = np.random.default_rng(0)
    n, k, m = 300, 256, 7                       # 3 tiles (last one partial), 4 k stages
    w = rng.integers(1, 2 ** 16, size=(n, k), dtype=np.uint16)
    x = rng.integers(1, 2 ** 16, size=(m, k), dtype=np.uint16)
    packed, xb = pack_canonical(w), write_b_operand(x)
    kst = k // STAGE_K
    for tile in range((n + TILE_ROWS - 1) // TILE_ROWS):
        for s in range(kst):
            stage = packed[(tile * kst + s) * STAGE_BYTES // 2:(tile * kst + s + 1) * STAGE_BYTES // 2]
            for kk in range(STAGE_K // 16):
                a_blk = umma_read(stage, kk * 256, lbo=128, sbo=1024, mn=TILE_ROWS)      # kernel: a_addr + k * 256
                b_blk = umma_read(xb, (s * 8 + kk * 2) * 256, lbo=256, sbo=128, mn=16)   # kernel: xb + (s*8 + 2k) * 256
                k0 = s * STAGE_K + kk * 16
                want_a = np.zeros((TILE_ROWS, 16), dtype=np.uint16)
                rows = min(TILE_ROWS, n - tile * TILE_ROWS)
                want_a[:rows] = w[tile * TILE_ROWS:tile * TILE_ROWS + rows, k0:k0 + 16]
                want_b = np.zeros((16, 16), dtype=np.uint16)
                want_b[:m] = x[:, k0:k0 + 16]
                assert np.array_equal(a_blk, want_a), (tile, s, kk)
                assert np.array_equal(b_blk, want_b), (tile, s, kk)


def test_descriptor_and_instruction_words():
    """Bit layout of the hand-built descriptors (cute::UMMA::SmemDescriptor / InstrDescriptor)."""
    def umma_desc(addr, lbo, sbo):
        return ((addr & 0x3FFFF) >> 4) | (((lbo >> 4) & 0x3FFF) << 16) | (((sbo >> 4) & 0x3FFF) << 32) | (1 << 46)

    d = umma_desc(0x12340, 128, 1024)
    assert d & 0x3FFF == 0x1234 and (d >> 16) & 0x3FFF == 8 and (d >> 32) & 0x3FFF == 64
    assert (d >> 46) & 3 == 1 and d >> 61 == 0 and (d >> 49) & 0xF == 0
    idesc = (1 << 4) | (1 << 7) | (1 << 10) | ((16 >> 3) << 17) | ((128 >> 4) << 24)
    assert (idesc >> 4) & 3 == 1                  # D = f32
    assert (idesc >> 7) & 7 == 1 and (idesc >> 10) & 7 == 1        # A, B = bf16
    assert (idesc >> 15) & 3 == 0                 # both K-major
    assert (idesc >> 17) & 0x3F == 2 and (idesc >> 24) & 0x1F == 8  # N = 16, M = 128
