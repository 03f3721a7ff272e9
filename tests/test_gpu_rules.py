"""GPU parity, tier 1: move lists, applied positions, terminal/winner -- bit-exact against the oracle,
through the C ABI (include/gameai_gpu.h), on the committed reference outputs and on the seeded corpus."""
import numpy as np
import pytest
from util import ut_vectors, ref_corpus, case_state, parse_move

pytestmark = pytest.mark.gpu
SEED = 0x10A5EED


def test_reference_ut_vectors_on_gpu(gpu):
    V = ut_vectors()
    for case in V["move_cases"]:
        w, b, cp = case_state(case)
        c, m = gpu.movegen(np.array([[w, b, cp]], dtype=np.uint64))
        got = [tuple(x) for x in m[0, :c[0]].tolist()]
        assert sorted(got) == sorted(parse_move(s) for s in case["moves"]), case["name"]
    for case in V["score_cases"]:
        w, b, _ = case_state(case)
        t, sc = gpu.terminal(np.array([[w, b, 0]], dtype=np.uint64))
        assert sc[0].tolist() == case["score"], case["name"]
    for case in V["terminal_cases"]:
        w, b, _ = case_state(case)
        t, sc = gpu.terminal(np.array([[w, b, 0]], dtype=np.uint64))
        assert bool(t[0]) == case["terminal"], case["name"]
    boards = [int(x, 16) for x in V["calc_if_terminal"]["boards"]]
    # calcIfTerminal(board): put the board on white, leave black with two far-apart pieces (never connected)
    far = (1 << 0) | (1 << 63)
    S = np.array([[bd, 0, 0] for bd in boards], dtype=np.uint64)
    t, sc = gpu.terminal(S)
    assert [bool(x) for x in t] == V["calc_if_terminal"]["expected"]
    w, b, cp = case_state(V["noop_roundtrip"])
    s1 = gpu.apply(np.array([[w, b, cp]], dtype=np.uint64), np.array([[0, 0]], dtype=np.uint8))
    s2 = gpu.apply(s1, np.array([[0, 0]], dtype=np.uint8))
    assert s2[0].tolist() == [w, b, cp] and s1[0].tolist() == [w, b, cp ^ 1]


def test_committed_reference_corpus_on_gpu(gpu):
    G = ref_corpus()
    S = G["states"]
    c, m = gpu.movegen(S)
    assert (c == G["counts"]).all()
    assert (m == G["moves"]).all()                      # ordered (from,to) lists, zero padded
    t, sc = gpu.terminal(S)
    assert (t == G["terminal"]).all() and (sc == G["score"]).all()
    assert (gpu.successor_hash(S) == G["successor_hash"]).all()
    assert (gpu.apply(S, G["mv_first"]) == G["apply_first"]).all()
    assert (gpu.apply(S, G["mv_last"]) == G["apply_last"]).all()


def test_crowded_rows(gpu, oracle):
    """Four own pieces of one square colour in one row, all free to move the same way: the positions that reach the top
    plane of the sweep kernel's carry-save row counter (skipped under a warp vote everywhere else)."""
    from test_sweep_host import crowded_rows
    S = crowded_rows()
    c, m = gpu.movegen(S); oc, om = oracle.movegen_batch(S)
    assert (c == oc).all() and (m == om).all()
    assert c.max() >= 20
    counts, st = gpu.rollout(S, 4, 50000, 0xC0FFEE, 0)
    ocn, opl = oracle.rollout_states(S, 4, 50000, 0xC0FFEE, 0)
    assert (counts == ocn).all() and st["plies"] == opl


def test_empty_and_ragged(gpu, oracle):
    e = np.zeros((0, 3), np.uint64)
    c, m = gpu.movegen(e)
    assert c.shape == (0,) and m.shape == (0, 96, 2)
    t, sc = gpu.terminal(e)
    assert t.shape == (0,)
    for n in (1, 31, 33, 127, 129, 1000):
        S = oracle.gen_uniform(n, 1234 + n)
        c, m = gpu.movegen(S); oc, om = oracle.movegen_batch(S)
        assert (c == oc).all() and (m == om).all()
    # degenerate boards: empty board, one piece each, current_player with junk in the upper 63 bits
    S = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 1], [1, 1 << 63, 0], [0x81, 0x7e, 0xFFFFFFFFFFFFFFFE], [0x81, 0x7e, 0xFFFFFFFFFFFFFFFF]], dtype=np.uint64)
    c, m = gpu.movegen(S); oc, om = oracle.movegen_batch(S)
    assert (c == oc).all() and (m == om).all()
    t, sc = gpu.terminal(S); ot, osc = oracle.terminal_batch(S)
    assert (t == ot).all() and (sc == osc).all()


def test_full_size_sweep_1m_positions(gpu, oracle):
    """Config 2 of BASELINE.json: 1 M synthetic positions (500 k reachable + 500 k uniform), GPU vs oracle.
    The reachable half is generated ON the GPU (gai_loa_gen_reachable) and spot-checked against the oracle's
    generator; the uniform half by the oracle's generator.  Compared: n, ordered move list, every successor
    (hash of white/black/side/connectivity), terminal, score."""
    n_half = 500_000
    reach = gpu.gen_reachable(n_half, SEED, 300)
    probe = np.arange(0, n_half, 997)[:400]
    # oracle's generator is per-index reproducible: regenerate a prefix and compare the probe points inside it
    pref = oracle.gen_reachable(4000, SEED, 300)
    assert (reach[:4000] == pref).all()
    unif = oracle.gen_uniform(n_half, SEED)
    S = np.concatenate([reach, unif])
    c, m = gpu.movegen(S); oc, om = oracle.movegen_batch(S)
    assert (c == oc).all()
    assert (m == om).all()
    t, sc = gpu.terminal(S); ot, osc = oracle.terminal_batch(S)
    assert (t == ot).all() and (sc == osc).all()
    assert (gpu.successor_hash(S) == oracle.successor_hash_batch(S)).all()
    # apply: a pseudo-random legal move per position
    pick = (np.arange(len(S)) * 2654435761 % np.maximum(c.astype(np.int64), 1)).astype(np.int64)
    mv = m[np.arange(len(S)), pick].copy(); mv[c == 0] = 0
    assert (gpu.apply(S, mv) == oracle.apply_batch(S, mv)).all()
    assert t[:n_half].sum() == 0 and 0.15 < t[n_half:].mean() < 0.21
