"""Host-side multi-GPU logic on CPU: world_size-2 gloo process group, chain sharding + final gather."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    from mcmc_symreg_b200.parallel import shard_range
    for n in (0, 1, 7, 50, 4096, 65536):
        for world in (1, 2, 3, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mcmc_symreg_b200 import parallel
    ctx = parallel.context()
    assert (ctx.rank, ctx.world) == (rank, world)
    seed = ctx.broadcast_int(1234 + rank)              # rank 0's value wins
    MM, K = 7, 2
    lo, hi = parallel.shard_range(MM, rank, world)
    # stand-in for what parallel.collect(engine) returns: arrays tagged with the global chain id
    ids = np.arange(lo, hi)
    # (packed trees: every tree is "lt(x_id)": two tokens, one (a, b) pair carrying the chain id; the RMSE traces of the ranks have
    # different capacities, as after bsr_reserve_err)
    ptok = np.tile(np.array([2, 0], dtype=np.uint32), (hi - lo) * K) | np.repeat((ids.astype(np.uint32) << 16), 2 * K) * np.tile(np.array([0, 1], dtype=np.uint32), (hi - lo) * K)
    res = dict(nn=np.full((hi - lo, K), 2, dtype=np.int32), ptok=ptok, pab=np.repeat(ids.astype(float), K)[:, None] * np.ones((1, 2)),
               beta=np.tile(ids[:, None], (1, K + 1)).astype(float), sigma=ids.astype(float), sse=np.zeros(hi - lo),
               sa=np.zeros((hi - lo, K)), sb=np.zeros((hi - lo, K)), counters=np.zeros((hi - lo, 8), dtype=np.int64),
               done=np.ones(hi - lo, dtype=np.int32), nerr=np.zeros(hi - lo, dtype=np.int32), err=np.zeros((hi - lo, 4 + 3 * rank)), sweeps=10 + rank)
    out = ctx.gather_results(res, MM, K)
    from mcmc_symreg_b200 import capi
    pk = capi.PackedTrees(out["nn"], out["ptok"], out["pab"])
    tok, pa, pb, nn = pk.dense()
    ok = (seed == 1234 and out["nn"].shape == (MM, K) and np.array_equal(out["sigma"], np.arange(MM, dtype=float))
          and np.array_equal(tok[:, 0, 1] >> 16, np.arange(MM)) and np.array_equal(pa[:, 1, 0], np.arange(MM, dtype=float))
          and np.array_equal(pk.tree(5, 1)[1][:2], [5.0, 0.0]) and out["err"].shape == (MM, 4 + 3 * (world - 1))
          and out["sweeps"] == 10 + world - 1)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_chain_sharding_gather_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_capi_library_exports_every_declared_symbol():
    """The C-ABI library loads on a CPU-only box and exports every entry point include/bsr_b200.h declares."""
    import ctypes
    import re
    import __graft_entry__ as g
    g.build()
    from mcmc_symreg_b200 import capi
    lib = ctypes.CDLL(capi.LIB_PATH)
    hdr = open(os.path.join(ROOT, "include", "bsr_b200.h")).read()
    names = sorted(set(re.findall(r"\b(bsr_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
    assert set(capi.EXPORTED) <= set(names)
    assert lib.bsr_version() >= 100 and lib.bsr_max_nodes() == 64


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mcmc_symreg_b200 import BSR, capi
    with pytest.raises(capi.BsrError, match="no CUDA device"):
        BSR(2, 2, seed=1).fit(np.random.rand(20, 2), np.random.rand(20))
