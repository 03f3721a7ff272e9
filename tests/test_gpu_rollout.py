"""GPU parity, tier 2: Philox-driven playouts bit-exact against the oracle (= reference rules driven by the
same choice stream), through the C ABI; plus size-independent properties at BASELINE.json's full sizes."""
import numpy as np
import pytest
from util import ref_corpus

pytestmark = pytest.mark.gpu
SEED = 0x10A5EED


def test_playout_trace_vs_committed_reference(gpu):
    G = ref_corpus()
    o, pl, f = gpu.playout_trace(G["play_states"], 50000, int(G["seed"]), 1000)
    assert (o == G["play_outcome"]).all() and (pl == G["play_plies"]).all() and (f == G["play_final"]).all()
    o, pl, f = gpu.playout_trace(G["play_states"], 40, int(G["seed"]), 1000)         # ply cap
    assert (o == G["cap_outcome"]).all() and (pl == G["cap_plies"]).all() and (f == G["cap_final"]).all()


def test_playout_trace_vs_oracle(gpu, oracle):
    S = np.concatenate([oracle.gen_reachable(3000, 42), oracle.gen_uniform(1000, 42)])
    for key, first in ((1, 0), (0xDEADBEEFCAFEF00D, (1 << 40) + 5)):           # 64-bit key and ids
        o, pl, f = gpu.playout_trace(S, 50000, key, first)
        oo, opl, of = oracle.playout_trace_batch(S, 50000, key, first)
        assert (o == oo).all() and (pl == opl).all() and (f == of).all()


def test_opening_playouts(gpu, oracle):
    S = np.tile(np.array([[0x0081818181818100, 0x7e0000000000007e, 1]], dtype=np.uint64), (2000, 1))
    o, pl, f = gpu.playout_trace(S, 50000, SEED, 0)
    oo, opl, of = oracle.playout_trace_batch(S, 50000, SEED, 0)
    assert (o == oo).all() and (pl == opl).all() and (f == of).all()
    assert 150 < pl.mean() < 280                                                # mean 213.8 plies


@pytest.mark.parametrize("n,ppl", [(1, 1), (33, 1), (257, 7), (4096, 2), (1000, 16)])
def test_rollout_counts_vs_oracle(gpu, oracle, n, ppl):
    S = np.concatenate([oracle.gen_reachable(n - n // 4, 7 + n), oracle.gen_uniform(n // 4, 7 + n)]) if n > 3 else oracle.gen_reachable(n, 7)
    counts, st = gpu.rollout(S, ppl, 50000, 0xFEED, 12345)
    oc, oplies = oracle.rollout_states(S, ppl, 50000, 0xFEED, 12345)
    assert (counts == oc).all()
    assert st["playouts"] == len(S) * ppl and st["plies"] == oplies


def test_rollout_edge_cases(gpu, oracle):
    # empty batch
    counts, st = gpu.rollout(np.zeros((0, 3), np.uint64), 4)
    assert counts.shape == (0, 4) and st["playouts"] == 0
    # terminal leaves score at 0 plies; ply cap 0 and 1
    S = oracle.gen_uniform(512, 5)
    t, _ = oracle.terminal_batch(S)
    assert t.sum() > 20
    for cap in (0, 1, 3, 17):
        counts, st = gpu.rollout(S, 3, cap, 9, 0)
        oc, oplies = oracle.rollout_states(S, 3, cap, 9, 0)
        assert (counts == oc).all() and st["plies"] == oplies
    # invalid arguments fail loudly
    import gameai_b200 as g
    with pytest.raises(g.GaiError):
        gpu.rollout(S, 0)


def test_rollout_independent_of_launch_shape(gpu, oracle):
    """Playout identity is (leaf, playout) -> Philox counter, never the lane: any launch shape, same counts."""
    S = oracle.gen_reachable(3000, 77)
    ref_counts, ref_st = gpu.rollout(S, 5, 50000, 3, 0)
    default = gpu.launch_config()["threads_per_block"]
    for tpb, bps in ((32, 1), (64, 3), (128, 8), (256, 2), (512, 1), (640, 1), (768, 1), (896, 1), (1024, 1)):
        gpu.set_launch_config(tpb, bps)
        assert gpu.launch_config()["threads_per_block"] == (tpb if gpu.impl else min(tpb, 256))
        counts, st = gpu.rollout(S, 5, 50000, 3, 0)
        assert (counts == ref_counts).all() and st["plies"] == ref_st["plies"]
    gpu.set_launch_config(default, 4)
    oc, oplies = oracle.rollout_states(S[:500], 5, 50000, 3, 0)
    assert (ref_counts[:500] == oc).all()


def test_sweep_kernel_unroll_variants(oracle):
    """The static-sweep kernel with 1, 2 and 4 plies per loop trip (GAI_SWEEP_UNROLL, read at gai_create): same playouts."""
    import os
    import gameai_b200 as g
    S = oracle.gen_reachable(4000, 91)
    oc, oplies = oracle.rollout_states(S[:600], 3, 50000, 11, 5)
    ref = None
    try:
        for u in ("1", "2", "4"):
            os.environ["GAI_SWEEP_UNROLL"] = u
            eng = g.GpuRollout(0)
            assert eng.launch_config()["plies_per_trip"] == int(u)
            counts, st = eng.rollout(S, 3, 50000, 11, 5)
            eng.close()
            assert (counts[:600] == oc).all()
            if ref is None:
                ref = (counts, st["plies"])
            assert (counts == ref[0]).all() and st["plies"] == ref[1]
    finally:
        os.environ.pop("GAI_SWEEP_UNROLL", None)


def test_device_resident_path_matches_host_path(gpu, oracle):
    n = 5000
    S = oracle.gen_reachable(n, 31)
    import gameai_b200 as g
    boards, stm = g.pack_states(S)
    d_b = gpu.dev_alloc(boards.nbytes); d_s = gpu.dev_alloc(stm.nbytes); d_o = gpu.dev_alloc(n * 16)
    try:
        gpu.h2d(d_b, boards); gpu.h2d(d_s, stm)
        st = gpu.rollout_device(d_b, d_s, n, d_o, 4, 50000, 0xAB, 10)
        out = np.zeros((n, 4), np.uint32); gpu.d2h(out, d_o)
        hc, hst = gpu.rollout(S, 4, 50000, 0xAB, 10)
        assert (out == hc).all() and st["plies"] == hst["plies"]
    finally:
        gpu.dev_free(d_b); gpu.dev_free(d_s); gpu.dev_free(d_o)


def test_full_size_properties_1m_leaves(gpu, oracle):
    """Config 3 at full size (1 Mi leaves): properties that do not need the oracle at that size, plus an
    oracle check on a strided sample of leaves (same playout ids)."""
    n, ppl = 1 << 20, 2
    d_b = gpu.dev_alloc(n * 16); d_s = gpu.dev_alloc(((n + 31) // 32) * 4); d_o = gpu.dev_alloc(n * 16)
    try:
        S = gpu.gen_reachable(n, SEED, 300, d_boards=d_b, d_stm=d_s, host=True)
        st = gpu.rollout_device(d_b, d_s, n, d_o, ppl, 50000, SEED, 0)
        out = np.zeros((n, 4), np.uint32); gpu.d2h(out, d_o)
        assert (out.sum(axis=1) == ppl).all()                        # every playout accounted for exactly once
        assert st["playouts"] == n * ppl
        assert out[:, 3].sum() == 0                                  # nothing reaches the 50 000-ply cap
        mean_plies = st["plies"] / st["playouts"]
        assert 80 < mean_plies < 180
        w, b, d = out[:, 0].sum(), out[:, 1].sum(), out[:, 2].sum()
        assert abs(int(w) - int(b)) < 0.03 * n * ppl and d < 0.02 * n * ppl
        # idempotence: same key/ids -> identical counts
        st2 = gpu.rollout_device(d_b, d_s, n, d_o, ppl, 50000, SEED, 0)
        out2 = np.zeros((n, 4), np.uint32); gpu.d2h(out2, d_o)
        assert (out == out2).all() and st2["plies"] == st["plies"]
        # oracle on a strided sample: leaf i's playouts have ids i*ppl + j -> reproduce with first_id = i*ppl
        idx = np.arange(0, n, 4099)
        for i in idx[:200]:
            oc, _ = oracle.rollout_states(S[i:i + 1], ppl, 50000, SEED, int(i) * ppl)
            assert (oc[0] == out[i]).all(), i
    finally:
        gpu.dev_free(d_b); gpu.dev_free(d_s); gpu.dev_free(d_o)


def test_async_rollout_matches_sync(gpu, oracle):
    """gai_loa_rollout_async + gai_sync + gai_last_rollout_stats (the pipelined path of the host player)"""
    import gameai_b200 as g
    S = oracle.gen_reachable(2000, 61)
    sync_counts, sync_st = gpu.rollout(S, 3, 50000, 17, 5)
    out = np.zeros((len(S), 4), np.uint32)
    gpu.rollout_async(S, out, 3, 50000, 17, 5)
    with pytest.raises(g.GaiError):                      # one batch in flight per context
        gpu.rollout_async(S, out, 3, 50000, 17, 5)
    gpu.sync()
    st = gpu.last_rollout_stats()
    assert (out == sync_counts).all() and st["plies"] == sync_st["plies"] and st["playouts"] == sync_st["playouts"]


# ---- ticketed batches (several queued per context) and batch-by-expansion ------------------------------------------
def _children_by_oracle(oracle, P):
    """all successors of every parent in reference move order -> (children (m,3), offsets (n+1,))"""
    cnt, mv = oracle.movegen_batch(P)
    kids, off = [], [0]
    for i in range(len(P)):
        n = int(cnt[i])
        if n:
            kids.append(oracle.apply_batch(np.repeat(P[i:i + 1], n, axis=0), mv[i, :n]))
        off.append(off[-1] + n)
    return (np.concatenate(kids) if kids else np.zeros((0, 3), np.uint64)), np.array(off, np.uint32)


def test_ticketed_batches_queue_and_match_oracle(gpu, oracle):
    import gameai_b200 as g
    batches = [oracle.gen_reachable(700 + 100 * k, 100 + k) for k in range(4)]
    outs = [np.zeros((len(b), 4), np.uint32) for b in batches]
    tickets = [gpu.submit(b, o, 3, 50000, 0xABCD + k, 1000 * k) for k, (b, o) in enumerate(zip(batches, outs))]
    assert gpu.inflight() == 4
    with pytest.raises(g.GaiError):                                # a fifth batch has no slot
        gpu.submit(batches[0], np.zeros((len(batches[0]), 4), np.uint32), 3, 50000, 1, 0)
    with pytest.raises(g.GaiError):                                # the synchronous calls refuse while tickets are out
        gpu.movegen(batches[0][:4])
    for k in (2, 0, 3, 1):                                         # any order
        st = gpu.wait(tickets[k])
        oc, opl = oracle.rollout_states(batches[k], 3, 50000, 0xABCD + k, 1000 * k)
        assert (outs[k] == oc).all() and st["plies"] == opl and st["playouts"] == 3 * len(batches[k])
    assert gpu.inflight() == 0
    with pytest.raises(g.GaiError):
        gpu.wait(tickets[0])                                       # stale ticket
    c, _ = gpu.movegen(batches[0][:4])                             # the context is usable again
    assert c.shape == (4,)


def test_children_expansion_vs_oracle(gpu, oracle):
    import gameai_b200 as g
    P = np.concatenate([oracle.gen_reachable(300, 5), oracle.gen_uniform(100, 6)])
    kids, off = _children_by_oracle(oracle, P)
    out = np.zeros((len(kids), 4), np.uint32)
    t = gpu.submit_children(P, off, out, 2, 50000, 0x77, 555)
    st = gpu.wait(t)
    oc, opl = oracle.rollout_states(kids, 2, 50000, 0x77, 555)
    assert (out == oc).all() and st["plies"] == opl and st["playouts"] == 2 * len(kids)
    # a wrong child count is caught on the device
    bad = off.copy(); bad[1:] += 1
    out_bad = np.zeros((int(bad[-1]), 4), np.uint32)               # (kept alive until wait: the library writes into it)
    t = gpu.submit_children(P, bad, out_bad, 1, 10, 1, 0)
    with pytest.raises(g.GaiError):
        gpu.wait(t)
    # empty batch
    P0, off0, out0 = np.zeros((0, 3), np.uint64), np.zeros(1, np.uint32), np.zeros((0, 4), np.uint32)
    t = gpu.submit_children(P0, off0, out0, 1, 10, 1, 0)
    assert gpu.wait(t)["playouts"] == 0
