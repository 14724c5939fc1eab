"""world_size-2 gloo test of the multi-rank host logic (modem_b200/shard.py): block partition + the single payload
gather.  The per-rank decode is stood in for by the CPU oracle here (this is a CPU test; the GPU path is covered by
tests/test_gpu_parity.py and the N>1 bench)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_total, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib as O
    from modem_b200.shard import shard_range, gather_payload
    lo, hi = shard_range(n_total, rank, world)
    pcm, ns, sent = O.encode_batch(hi - lo, seed0=7000 + lo, nthreads=2)
    st, out = O.decode_batch(pcm, nthreads=2)
    assert (st == 0).all()
    full = gather_payload(torch.from_numpy(out), n_total)
    if rank == 0:
        q.put(full.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges():
    from modem_b200.shard import shard_range
    for n, w in ((10, 2), (7, 2), (5, 8), (1000000, 8), (0, 4)):
        r = [shard_range(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) <= 1


def test_two_rank_decode_and_gather(oracle):
    import oracle_lib as O
    n_total, world = 5, 2   # ragged on purpose: 3 + 2 windows
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    full = q.get(timeout=300)
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    expect = np.stack([O.make_payload(7000 + i) for i in range(n_total)])
    assert full.shape == (n_total, 5380) and (full == expect).all()
