"""GPU: INTEGRATION.md section 2 as a binary -- the reference's own MC::Player, patched (oracle/mcts_gpu_stub.patch) to cut its
simulation loop at the first new state and to run the playouts through gai_loa_rollout of the real libgameai_gpu.so, plays ten
moves from the opening."""
import os
import pytest
from test_integration_stub import SO, play, check

pytestmark = pytest.mark.gpu


def test_patched_reference_player_plays_ten_moves_on_the_gpu():
    if not os.path.exists(SO):
        pytest.skip("oracle/_ref/libref_mcts_gpu.so not built (needs /root/reference at build time)")
    # move_sim_limit counts BATCHES here (one can_continue() per batch of 16 selections x 8 playouts per leaf); the host side
    # is the reference's own tree code (a few hundred selections/s), so the budget stays small: 10 x 128 x 16 selections
    res = play(10, 128, None)
    check(res, 10, 128)
    assert res["playouts"] >= 10 * 128 * 16 * 8 * 0.5
