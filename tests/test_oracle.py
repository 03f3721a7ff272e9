"""CPU tests: pin the oracle restatement (oracle/loa_oracle.c) against
 (a) the reference's own unit-test vectors (tests/golden/loa_ut_vectors.json),
 (b) committed outputs of the reference itself (tests/golden/loa_ref_corpus.npz), and
 (c) the reference compiled live (oracle/_ref), when /root/reference is present."""
import numpy as np
import pytest
from util import ut_vectors, ref_corpus, case_state, parse_move

V = ut_vectors()


def test_masks_ut(oracle):
    for p, m in zip(V["neighbour_mask"]["positions"], V["neighbour_mask"]["masks"]):
        assert oracle.neighbour_mask(p) == int(m, 16)
    for p, m in zip(V["line_positions"]["positions"], V["column_mask"]["masks"]):
        assert oracle.column_mask(p) == int(m, 16)
    for p, m in zip(V["line_positions"]["positions"], V["row_mask"]["masks"]):
        assert oracle.row_mask(p) == int(m, 16)
    for p, m in zip(V["diag_positions"]["positions"], V["left_diagonal_mask"]["masks"]):
        assert oracle.left_diagonal_mask(p) == int(m, 16)
    for p, m in zip(V["diag_positions"]["positions"], V["right_diagonal_mask"]["masks"]):
        assert oracle.right_diagonal_mask(p) == int(m, 16)


def test_calc_if_terminal_ut(oracle):
    for b, e in zip(V["calc_if_terminal"]["boards"], V["calc_if_terminal"]["expected"]):
        assert oracle.calc_if_terminal(int(b, 16)) == e


@pytest.mark.parametrize("case", V["move_cases"], ids=lambda c: c["name"])
def test_move_cases_ut(oracle, case):
    w, b, cp = case_state(case)
    got = oracle.movegen(w, b, cp)
    # the reference test checks set equality (LinesOfActionZasady_ut.cpp:95-124)
    assert sorted(got) == sorted(parse_move(m) for m in case["moves"])
    assert len(got) == len(case["moves"])


def test_noop_roundtrip_ut(oracle):
    w, b, cp = case_state(V["noop_roundtrip"])
    # not-to-move player gets the noop; applying it twice returns to the same position
    s1 = oracle.apply(w, b, cp, 0, 0)
    s2 = oracle.apply(*s1, 0, 0)
    assert s2 == (w, b, cp)   # compared on (white, black, cp&1): the reference's memcmp reads padding


def test_to_string_ut(oracle):
    assert oracle.to_string(*oracle.initial()) == V["state_to_string"]["expected"]


@pytest.mark.parametrize("case", V["score_cases"], ids=lambda c: c["name"])
def test_score_ut(oracle, case):
    w, b, _ = case_state(case)
    assert list(oracle.score(w, b)) == case["score"]
    if "terminal" in case:
        assert oracle.is_terminal(w, b) == case["terminal"]


@pytest.mark.parametrize("case", V["terminal_cases"], ids=lambda c: c["name"])
def test_terminal_ut(oracle, case):
    w, b, _ = case_state(case)
    assert oracle.is_terminal(w, b) == case["terminal"]


def test_opening(oracle):
    w, b, cp = oracle.initial()
    assert (w, b, cp) == (0x0081818181818100, 0x7e0000000000007e, 1)
    mv = oracle.movegen(w, b, cp)
    assert len(mv) == 36
    name = lambda s: "ABCDEFGH"[s % 8] + str(s // 8 + 1)
    assert [name(f) + "->" + name(t) for f, t in mv[:6]] == ["B1->H1", "B1->B3", "B1->D3", "C1->C3", "C1->E3", "C1->A3"]


def test_edge_cases(oracle):
    assert not oracle.calc_if_terminal(0)                  # empty board: 0 groups, popcount 0
    assert oracle.movegen(0, 0, 0) == []
    assert oracle.score(0, 0) == (0, 0)
    full_white = (1 << 12) - 1
    assert oracle.calc_if_terminal(full_white)


# ---- (b) committed reference outputs ---------------------------------------------------------------
def test_port_vs_committed_reference_outputs(oracle):
    G = ref_corpus()
    S = G["states"]
    # the corpus inputs are reproducible from the seed
    n_r = 3000
    assert (oracle.gen_reachable(n_r, int(G["seed"])) == S[:n_r]).all()
    assert (oracle.gen_uniform(len(S) - n_r, int(G["seed"])) == S[n_r:]).all()
    c, m = oracle.movegen_batch(S)
    assert (c == G["counts"]).all() and (m == G["moves"]).all()
    t, sc = oracle.terminal_batch(S)
    assert (t == G["terminal"]).all() and (sc == G["score"]).all()
    assert (oracle.successor_hash_batch(S) == G["successor_hash"]).all()
    assert (oracle.apply_batch(S, G["mv_first"]) == G["apply_first"]).all()
    assert (oracle.apply_batch(S, G["mv_last"]) == G["apply_last"]).all()
    o, pl, f = oracle.playout_trace_batch(G["play_states"], 50000, int(G["seed"]), 1000)
    assert (o == G["play_outcome"]).all() and (pl == G["play_plies"]).all() and (f == G["play_final"]).all()
    o, pl, f = oracle.playout_trace_batch(G["play_states"], 40, int(G["seed"]), 1000)
    assert (o == G["cap_outcome"]).all() and (pl == G["cap_plies"]).all() and (f == G["cap_final"]).all()
    for i, fn in enumerate(("neighbour_mask", "row_mask", "column_mask", "left_diagonal_mask", "right_diagonal_mask")):
        assert [getattr(oracle, fn)(p) for p in range(64)] == [int(x) for x in G["masks"][i]]


def test_corpus_shape(oracle):
    G = ref_corpus()
    assert G["terminal"][:3000].sum() == 0                    # reachable half is non-terminal by construction
    assert 0.10 < G["terminal"][3000:].mean() < 0.30          # uniform half ~18 % terminal
    assert G["counts"].max() <= 96


# ---- (c) live reference ------------------------------------------------------------------------------
def test_port_vs_live_reference(oracle, ref_oracle):
    S = np.concatenate([oracle.gen_reachable(1500, 99), oracle.gen_uniform(1500, 99)])
    for fn in ("movegen_batch", "terminal_batch", "successor_hash_batch"):
        a, b = getattr(oracle, fn)(S), getattr(ref_oracle, fn)(S)
        a = a if isinstance(a, tuple) else (a,)
        b = b if isinstance(b, tuple) else (b,)
        assert all((x == y).all() for x, y in zip(a, b)), fn
    a = oracle.playout_trace_batch(S[:300], 50000, 5, 0)
    b = ref_oracle.playout_trace_batch(S[:300], 50000, 5, 0)
    assert all((x == y).all() for x, y in zip(a, b))
    ca, pa = oracle.rollout_states(S[:64], 3, 50000, 11, 5)
    cb, pb = ref_oracle.rollout_states(S[:64], 3, 50000, 11, 5)
    assert (ca == cb).all() and pa == pb
    assert ref_oracle.to_string(*ref_oracle.initial()) == oracle.to_string(*oracle.initial())
