"""CPU: the reference's own MC::Player (oracle/_ref/libref_mcts.so) builds, runs and behaves as SURVEY.md 6 measured
(mean path ~214 plies from the opening, 36 root moves).  Skipped where /root/reference was not available at build time."""
import pytest


def test_reference_player_runs_on_the_opening():
    from oracle import mcts
    ref = mcts.ref()
    if ref is None:
        pytest.skip("oracle/_ref/libref_mcts.so not built")
    r = ref.select(0x0081818181818100, 0x7e0000000000007e, 1, 300, seed=1234)
    assert len(r["visits"]) == 36 and sum(r["visits"]) == 300
    assert 150 < r["mean_path"] < 280
    assert r["sims_per_sec"] > 10
