"""CPU model of the three pipelines of csrc/lmhead_tc.cuh (TMA ring full/empty, TMEM accumulator
full/empty, persistent tile loop): one Python thread per role, an mbarrier model with the PTX
phase/parity semantics, the kernel's own parity expressions.  Checks that the protocol terminates
for every (tiles per CTA, ring depth, k stages) and that no buffer is overwritten before its
consumer released it — the class of bug that would hang or corrupt the (not yet executed) kernel."""
import itertools
import threading
import time

import pytest


class MBarrier:
    """mbarrier with an arrival count; wait(parity) returns once the phase of that parity is over."""

    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0
        self.cv = threading.Condition()

    def arrive(self):
        with self.cv:
            self.pending -= 1
            assert self.pending >= 0, "more arrivals than the barrier expects"
            if self.pending == 0:
                self.pending = self.count
                self.phase ^= 1
                self.cv.notify_all()

    def wait(self, parity, timeout=10.0):
        deadline = time.time() + timeout
        with self.cv:
            while self.phase == parity:           # phase `parity` still in progress
                left = deadline - time.time()
                if left <= 0:
                    raise TimeoutError("mbarrier wait: protocol deadlock")
                self.cv.wait(left)


def run_pipeline(n_tiles, NS, n_kst):
    full = [MBarrier(1) for _ in range(NS)]
    empty = [MBarrier(1) for _ in range(NS)]
    tfull = [MBarrier(1) for _ in range(2)]
    tempty = [MBarrier(4) for _ in range(2)]
    ring = [None] * NS                 # what the producer last wrote: (tile, s)
    ring_free = [True] * NS
    tmem = [None] * 2                  # tile whose accumulation is complete in this buffer
    tmem_busy = [False] * 2            # True from "MMA starts accumulating" to "4 warps drained"
    drained = [0, 0]
    errors, results = [], []
    lock = threading.Lock()

    def guard(fn):
        def wrapped(*a):
            try:
                fn(*a)
            except Exception as exc:   # noqa: BLE001
                errors.append(repr(exc))
        return wrapped

    @guard
    def producer():
        q = 0
        for tile in range(n_tiles):
            for s in range(n_kst):
                st = q % NS
                empty[st].wait(((q // NS) & 1) ^ 1)
                assert ring_free[st], f"stage {st} overwritten before the MMAs released it"
                ring_free[st] = False
                ring[st] = (tile, s)
                full[st].arrive()                     # arrive.expect_tx + TMA completion
                q += 1

    @guard
    def mma():
        q = 0
        for it, tile in enumerate(range(n_tiles)):
            buf = it & 1
            tempty[buf].wait(((it >> 1) & 1) ^ 1)
            assert not tmem_busy[buf], f"TMEM buffer {buf} overwritten before the epilogue drained it"
            tmem_busy[buf] = True
            for s in range(n_kst):
                st = q % NS
                full[st].wait((q // NS) & 1)
                assert ring[st] == (tile, s), f"MMA read stage {st} holding {ring[st]}, wanted {(tile, s)}"
                ring_free[st] = True
                empty[st].arrive()                    # tcgen05.commit -> empty
                q += 1
            tmem[buf] = tile
            tfull[buf].arrive()                       # tcgen05.commit -> accumulator ready

    @guard
    def epilogue(w):
        for it, tile in enumerate(range(n_tiles)):
            buf = it & 1
            tfull[buf].wait((it >> 1) & 1)
            assert tmem[buf] == tile, f"epilogue warp {w} read tile {tmem[buf]}, wanted {tile}"
            with lock:
                results.append((w, tile))
                drained[buf] += 1
                if drained[buf] == 4:
                    drained[buf] = 0
                    tmem_busy[buf] = False
            tempty[buf].arrive()

    threads = [threading.Thread(target=producer), threading.Thread(target=mma)] + \
              [threading.Thread(target=epilogue, args=(w,)) for w in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(30)
    assert not any(t.is_alive() for t in threads), "pipeline did not terminate"
    assert not errors, errors
    assert sorted(results) == sorted((w, t) for w in range(4) for t in range(n_tiles))


@pytest.mark.parametrize("n_tiles,NS,n_kst", list(itertools.product([1, 2, 3, 5, 8], [3, 4, 6], [1, 4, 7])))
def test_tma_ring_tmem_double_buffer_and_tile_loop_terminate_cleanly(n_tiles, NS, n_kst):
    run_pipeline(n_tiles, NS, n_kst)
