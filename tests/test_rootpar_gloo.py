"""The N>1 path on CPU: two processes over the gloo backend exercise the sharding plan and the once-per-move merge
(the GPU run does the same sum with NCCL through gai_allreduce_root / gai_allreduce_u64)."""
import os
import sys
import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import gameai_b200 as g  # noqa: F401
    import importlib
    rootpar = importlib.import_module("game-ai_b200.rootpar")
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.RandomState(100 + rank)                          # every rank searched its own tree
    n = 36
    visits = rng.randint(0, 5000, size=n).astype(np.uint32)
    values = (rng.rand(n, 2) * visits[:, None]).astype(np.float32)
    mv, mf = rootpar.merge_root_stats(visits, values, dist)
    move = rootpar.select_best_move(mv, mf, player=1, eps=0.05, seed=1234)
    plan = rootpar.shard_plan(rank, world, 1 << 20, 16, 0x10A5EED)
    q.put((rank, visits, values, mv, mf, move, plan))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_merge_and_sharding():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in range(2)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, v0, f0, mv0, mf0, m0, plan0), (r1, v1, f1, mv1, mf1, m1, plan1) = out
    assert (mv0 == v0 + v1).all() and (mv1 == mv0).all()                        # the sum, identical on both ranks
    assert np.allclose(mf0, f0 + f1) and (mf0 == mf1).all()
    assert m0 == m1                                                            # same merged numbers + same seed -> same move
    # sharding: disjoint, contiguous playout ids and distinct corpus shards
    assert plan0["first_playout_id"] == 0 and plan1["first_playout_id"] == plan0["playouts"]
    assert plan0["corpus_key"] != plan1["corpus_key"] and plan0["global_playouts"] == 2 * (1 << 20) * 16


def test_single_rank_is_identity():
    import importlib
    rootpar = importlib.import_module("game-ai_b200.rootpar")
    v = np.arange(10, dtype=np.uint32); f = np.ones((10, 2), np.float32)
    mv, mf = rootpar.merge_root_stats(v, f)
    assert (mv == v).all() and (mf == f).all()
    assert rootpar.select_best_move(v, f, 0, 0.005, 1) == 0                    # value 1/(1+0) is the largest
