"""CPU tests: pin the Pan oracle restatement (oracle/pan_oracle.c) against the reference's own unit-test vectors,
committed outputs of the reference itself, the live reference when present -- and check the PRODUCT's
__host__ __device__ Pan rules (game-ai_b200/csrc/pan_rules.cuh, run on the host) against the oracle."""
import ctypes
import os
import subprocess
import numpy as np
import pytest
from pan_util import ut_vectors, ref_corpus, mkmb, ones, case_state, move_fields, OPS, PROB

V = ut_vectors()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def po():
    from oracle import pan
    return pan.port()


@pytest.fixture(scope="module")
def pref():
    from oracle import pan
    r = pan.ref()
    if r is None:
        pytest.skip("oracle/_ref not built (needs /root/reference; container only)")
    return r


def test_make11_ut(po):
    for a, b in V["make11"]["cases"]:
        assert po.make11(int(a, 2)) == int(b, 2)


def test_stack_masks_ut(po):
    for c in V["stack_masks"]["cases"]:
        allowed, quad, take = po.make_stack_masks(mkmb(c["stack"]))
        assert allowed == ones(c["allowed_ones"]) + (mkmb(c["allowed_plus"]) if c["allowed_plus"] else 0)
        assert quad == mkmb(c["quad"])
        assert take == (mkmb(c["take"]) if c["take"] else 0)


@pytest.mark.parametrize("case", V["movegen_cases"], ids=lambda c: c["name"])
def test_movegen_ut(po, case):
    s = case_state(case)
    cnt, mv = po.movegen_batch(case["players"], s[None])
    assert cnt[0] == len(case["moves"])
    for k, e in enumerate(case["moves"]):
        f = move_fields(mv[0, k])
        assert f["op"] == OPS[e["op"]] and f["cards"] == mkmb(e["cards"]) and f["count"] == e["count"]
        if "p" in e:
            assert abs(PROB[f["prob"]] - e["p"]) < 1e-6


def test_noop_and_terminal_ut(po):
    for c in V["terminal_cases"]:
        t, sc, _, _ = po.terminal_batch(c["players"], case_state(c)[None])
        assert bool(t[0]) == c["terminal"]


def test_port_vs_committed_reference_outputs(po):
    G = ref_corpus()
    for npl in (2, 3, 4):
        k = "p%d_" % npl
        S = G[k + "states"]
        assert (po.gen_reachable(npl, len(S), int(G["seed"]), 100) == S).all()
        c, m = po.movegen_batch(npl, S)
        assert (c == G[k + "counts"]).all() and (m == G[k + "moves"]).all()
        assert (po.apply_batch(npl, S, G[k + "pick"]) == G[k + "applied"]).all()
        t, sc, e0, e1 = po.terminal_batch(npl, S)
        assert (t == G[k + "terminal"]).all() and (sc == G[k + "score"]).all() and (e0 == G[k + "eval0"]).all() and (e1 == G[k + "eval1"]).all()
        for tag, cap, ek in (("cap50w", 50, 1), ("cap50n", 50, 0), ("full", 100000, 1)):
            code, pl, v, f = po.playout_trace_batch(npl, S[:400], cap, ek, int(G["seed"]), 500)
            assert (code == G[k + tag + "_code"]).all() and (pl == G[k + tag + "_plies"]).all()
            assert (v == G[k + tag + "_value"]).all() and (f == G[k + tag + "_final"]).all()
        vc, vm = po.movegen_batch(npl, G[k + "views"])
        assert (vc == G[k + "view_counts"]).all() and (vm == G[k + "view_moves"]).all()


def test_port_vs_live_reference(po, pref):
    for npl in (2, 3, 4):
        S = po.gen_reachable(npl, 1200, 4242, 120)
        assert (S == pref.gen_reachable(npl, 1200, 4242, 120)).all()
        a, b = po.movegen_batch(npl, S), pref.movegen_batch(npl, S)
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all()
        ra, pa = po.playout_batch(npl, S[:200], 3, 50, 1, 5, 7)
        rb, pb = pref.playout_batch(npl, S[:200], 3, 50, 1, 5, 7)
        assert (ra == rb).all() and pa == pb


# ---- the product's host-callable Pan / LUT code vs the oracle (no GPU) -----------------------------------
def test_product_pan_rules_on_host_vs_oracle(po, hostlib):
    for npl in (2, 3, 4):
        S = po.gen_reachable(npl, 1500, 77, 100)
        c, m = po.movegen_batch(npl, S)
        t, sc, e0, e1 = po.terminal_batch(npl, S)
        for i in range(len(S)):
            mv = (ctypes.c_uint64 * 16)()
            n = hostlib.host_pan_movegen(npl, S[i].ctypes.data, mv)
            assert n == c[i] and list(mv[:n]) == [int(x) for x in m[i, :n]]
            if n:
                k = i % n
                out = np.zeros(5, np.uint64)
                assert hostlib.host_pan_apply(npl, S[i].ctypes.data, mv[k], out.ctypes.data) == 0
                assert (out == po.apply_batch(npl, S[i:i + 1], np.array([mv[k]], dtype=np.uint64))[0]).all()
            s4 = (ctypes.c_int * 4)(); a4 = (ctypes.c_int * 4)(); b4 = (ctypes.c_int * 4)()
            assert hostlib.host_pan_terminal(npl, S[i].ctypes.data, s4, a4, b4) == t[i]
            assert list(s4[:npl]) == list(sc[i, :npl]) and list(a4[:npl]) == list(e0[i, :npl]) and list(b4[:npl]) == list(e1[i, :npl])
    # fuzzy (imperfect-information) states are refused by the product rules
    view = po.player_known_state(3, po.gen_reachable(3, 1, 1, 10)[0], 0)
    mv = (ctypes.c_uint64 * 16)()
    assert hostlib.host_pan_movegen(3, view.ctypes.data, mv) == -1


def test_loa_line_lut_definition(hostlib, oracle):
    """The product's line LUT entry (loa_fast.cuh lut_entry) against an independent statement: place the 8 cells on
    row 0 of a board and ask the oracle's movegen which row moves exist."""
    import random
    rng = random.Random(5)
    for _ in range(1500):
        o = rng.getrandbits(8); e = rng.getrandbits(8) & ~o
        if o == 0:
            continue
        v = hostlib.host_loa_lut_entry(o, e)
        mv = oracle.movegen(o, e, 0)          # white = own on row 0, black = enemy
        exp = 0
        for f, t in mv:
            if t < 8:                           # row moves only
                exp |= (2 if t > f else 1) << (2 * f)
        assert v == exp, (bin(o), bin(e), bin(v), bin(exp))
    # rotation indices are permutations of the 64 squares
    for fn in ("host_loa_pos_c", "host_loa_pos_d1", "host_loa_pos_d2"):
        assert sorted(getattr(hostlib, fn)(s) for s in range(64)) == list(range(64))
