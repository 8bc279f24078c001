"""CPU, world_size 2, gloo: the N > 1 host plumbing (prompt sharding for replicas, the NCCL
unique-id broadcast the tensor-parallel engines use, max/sum reductions of bench.py)."""
import os
import socket

import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from layerskip_b200 import parallel_util as pu
        from layerskip_b200.synthetic import synthetic_prompts
        prompts = synthetic_prompts(512, 8, 16)
        mine = pu.shard_prompts(prompts, rank, world)
        payload = bytes(range(128)) if rank == 0 else b""
        got = pu.broadcast_bytes(payload, 128, src=0)
        mx = pu.reduce_scalar(10.0 + rank, "max")
        sm = pu.reduce_scalar(len(mine), "sum")
        q.put((rank, [p[0] for p in mine], got == bytes(range(128)), mx, sm))
    finally:
        dist.destroy_process_group()


def test_world_size_2_plumbing():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from layerskip_b200.synthetic import synthetic_prompts
    firsts = [p[0] for p in synthetic_prompts(512, 8, 16)]
    assert res[0][1] == firsts[0::2] and res[1][1] == firsts[1::2]      # disjoint, covering
    assert all(r[2] for r in res)                                          # same id bytes everywhere
    assert all(r[3] == 11.0 for r in res) and all(r[4] == 8.0 for r in res)


def _seed_worker(rank, world, port, q):
    import torch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from layerskip_b200.strategy import _EngineCache, _generation_seed
        torch.manual_seed(1000 + rank)                      # ranks disagree, like unseeded processes
        eng = type("E", (), dict(tp_size=world, device=None))()
        cache = _EngineCache(tp_size=world)
        a, b = _generation_seed(cache, eng, True), _generation_seed(cache, eng, True)
        q.put((rank, a, _generation_seed(cache, eng, False), b))
    finally:
        dist.destroy_process_group()


def test_sampling_seed_is_rank_zeros_on_every_tp_rank():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_seed_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    import torch
    torch.manual_seed(1000)
    want = [int(torch.randint(0, 2 ** 31 - 1, ()).item()) for _ in range(2)]
    assert res[0][1] == res[1][1] == want[0]               # sampling: rank 0's draw on every rank
    assert res[0][3] == res[1][3] == want[1] != want[0]    # the next call gets a fresh seed
    assert (res[0][2], res[1][2]) == (0, 0)                # greedy: no collective, no RNG consumed
