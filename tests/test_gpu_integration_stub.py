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
    res = play(10, 2048, None)
    check(res, 10, 2048)
    assert res["playouts"] >= 10 * 2048 * 8 * 0.9
