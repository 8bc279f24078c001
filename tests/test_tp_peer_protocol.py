"""Model check of the one-shot peer collective protocol (layerskip_b200/csrc/tp_peer.cuh) on CPU.

The CUDA kernels cannot run here; what CAN be checked without a GPU is the protocol they
implement: epoch counter in owner memory, monotonic per-(source, CTA) flags, data slots
double-buffered by epoch parity, "last CTA advances the epoch".  This file restates that protocol
with one Python thread per (rank, CTA), random scheduling jitter and grids that change from
instance to instance, and asserts that every rank obtains the rank-ordered sum of every instance
— i.e. no slot is overwritten while it is still being read and no wait is satisfied early.
A deliberately broken variant (single-buffered slots) must be caught by the same harness."""
import random
import threading
import time

import pytest

MAX_CTAS = 4


class Region:
    """One rank's peer-visible memory."""

    def __init__(self, tp, n):
        self.data = [[[None] * n for _ in range(tp)] for _ in range(2)]   # [parity][src][elem]
        self.flags = [[0] * MAX_CTAS for _ in range(tp)]                  # [src][cta]
        self.epoch = 0
        self.ticket = 0
        self.lock = threading.Lock()                                      # models atomicAdd


def _cta(rank, cta, grid, regions, partial, x, n, double_buffer, jitter, errors):
    tp = len(regions)
    mine = regions[rank]
    epoch = mine.epoch + 1
    par = (epoch & 1) if double_buffer else 0
    lo, hi = cta * n // grid, (cta + 1) * n // grid
    jitter()
    for r in range(tp):                                   # 1. push
        if r != rank:
            for i in range(lo, hi):
                regions[r].data[par][rank][i] = partial[i]
    jitter()
    for r in range(tp):                                   # 2. signal ...
        if r != rank:
            regions[r].flags[rank][cta] = epoch
    deadline = time.time() + 20.0
    for r in range(tp):                                   # ... and wait
        if r != rank:
            while mine.flags[r][cta] - epoch < 0:
                if time.time() > deadline:
                    errors.append(f"rank {rank} cta {cta} timed out at epoch {epoch}")
                    return
                time.sleep(0)
    jitter()
    for i in range(lo, hi):                               # 3. rank-ordered sum + residual
        acc = None
        for r in range(tp):
            v = partial[i] if r == rank else mine.data[par][r][i]
            acc = v if acc is None else acc + v
        x[i] += acc
    with mine.lock:                                       # 4. last CTA advances the epoch
        mine.ticket += 1
        if mine.ticket == grid:
            mine.ticket = 0
            mine.epoch = epoch


def _rank(rank, regions, partials, x, grids, n, double_buffer, seed, errors):
    rng = random.Random(seed)

    def jitter():
        if rng.random() < 0.3:
            time.sleep(rng.random() * 0.002)

    for inst, grid in enumerate(grids):
        # a kernel's CTAs run concurrently; the next instance starts when all of them are done
        threads = [threading.Thread(target=_cta, args=(rank, c, grid, regions, partials[inst][rank],
                                                       x, n, double_buffer, jitter, errors))
                   for c in range(grid)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            return
        if rng.random() < 0.2:
            time.sleep(rng.random() * 0.004)              # one rank falls behind for a while


def _simulate(tp, n_inst, double_buffer, seed):
    n = 12
    rng = random.Random(seed)
    grids = [rng.choice([1, 2, 3, MAX_CTAS]) for _ in range(n_inst)]
    partials = [[[rng.randrange(1, 1000) * 1000 ** r for _ in range(n)] for r in range(tp)]
                for _ in range(n_inst)]
    regions = [Region(tp, n) for _ in range(tp)]
    xs = [[0] * n for _ in range(tp)]
    errors = []
    threads = [threading.Thread(target=_rank, args=(r, regions, partials, xs[r], grids, n,
                                                    double_buffer, seed * 31 + r, errors))
               for r in range(tp)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    want = [sum(partials[k][r][i] for k in range(n_inst) for r in range(tp)) for i in range(n)]
    return xs, want, errors, regions


@pytest.mark.parametrize("tp,seed", [(2, 1), (2, 2), (3, 3), (4, 4)])
def test_protocol_delivers_every_sum_to_every_rank(tp, seed):
    xs, want, errors, regions = _simulate(tp, n_inst=40, double_buffer=True, seed=seed)
    assert not errors, errors
    for r in range(tp):
        assert xs[r] == want, f"rank {r}"
        assert regions[r].epoch == 40 and regions[r].ticket == 0


def test_harness_catches_single_buffered_slots():
    """Without parity double-buffering a fast peer overwrites a slot that is still being read:
    the harness must be able to see that (otherwise the test above proves nothing)."""
    caught = False
    for seed in range(1, 30):
        xs, want, errors, _ = _simulate(2, n_inst=60, double_buffer=False, seed=seed)
        if errors or any(x != want for x in xs):
            caught = True
            break
    assert caught


# ---- fused mode (LSK_TP_ONESHOT=2): the GEMM's CTAs push + flag, a finish kernel waits for all ----
class FusedRegion(Region):
    def __init__(self, tp, n):
        super().__init__(tp, n)
        self.gemm_flags = [[0] * MAX_CTAS for _ in range(tp)]


def _gemm_cta(rank, cta, grid, regions, partial, n, jitter):
    mine = regions[rank]
    epoch = mine.epoch + 1
    par = epoch & 1
    jitter()
    for i in range(cta, n, grid):                       # this CTA's tiles, strided like the kernel
        for r in range(len(regions)):                   # every rank, this one included
            regions[r].data[par][rank][i] = partial[i]
    jitter()
    for r in range(len(regions)):
        regions[r].gemm_flags[rank][cta] = epoch


def _finish_cta(rank, cta, grid, n_src, regions, x, n, jitter, errors):
    tp = len(regions)
    mine = regions[rank]
    epoch = mine.epoch + 1
    par = epoch & 1
    deadline = time.time() + 20.0
    for r in range(tp):
        for c in range(n_src):
            while mine.gemm_flags[r][c] - epoch < 0:
                if time.time() > deadline:
                    errors.append(f"rank {rank} finish cta {cta} timed out at epoch {epoch}")
                    return
                time.sleep(0)
    jitter()
    for i in range(cta * n // grid, (cta + 1) * n // grid):
        acc = None
        for r in range(tp):
            v = mine.data[par][r][i]
            acc = v if acc is None else acc + v
        x[i] += acc
    with mine.lock:
        mine.ticket += 1
        if mine.ticket == grid:
            mine.ticket = 0
            mine.epoch = epoch


def _fused_rank(rank, regions, partials, x, gemm_grids, fin_grids, n, seed, errors):
    rng = random.Random(seed)

    def jitter():
        if rng.random() < 0.3:
            time.sleep(rng.random() * 0.002)

    def run(target, grid, args):
        threads = [threading.Thread(target=target, args=(rank, c, grid) + args) for c in range(grid)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()

    for inst in range(len(gemm_grids)):
        run(_gemm_cta, gemm_grids[inst], (regions, partials[inst][rank], n, jitter))
        run(_finish_cta, fin_grids[inst], (gemm_grids[inst], regions, x, n, jitter, errors))
        if errors:
            return
        if rng.random() < 0.2:
            time.sleep(rng.random() * 0.004)


@pytest.mark.parametrize("tp,seed", [(2, 5), (3, 6), (4, 7)])
def test_fused_protocol_delivers_every_sum_to_every_rank(tp, seed):
    n, n_inst = 12, 40
    rng = random.Random(seed)
    gemm_grids = [rng.choice([1, 2, 3, MAX_CTAS]) for _ in range(n_inst)]
    fin_grids = [rng.choice([1, 2, 3]) for _ in range(n_inst)]
    partials = [[[rng.randrange(1, 1000) * 1000 ** r for _ in range(n)] for r in range(tp)]
                for _ in range(n_inst)]
    regions = [FusedRegion(tp, n) for _ in range(tp)]
    xs = [[0] * n for _ in range(tp)]
    errors = []
    threads = [threading.Thread(target=_fused_rank, args=(r, regions, partials, xs[r], gemm_grids,
                                                          fin_grids, n, seed * 31 + r, errors))
               for r in range(tp)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    want = [sum(partials[k][r][i] for k in range(n_inst) for r in range(tp)) for i in range(n)]
    for r in range(tp):
        assert xs[r] == want, f"rank {r}"
        assert regions[r].epoch == n_inst
