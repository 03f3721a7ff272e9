"""GPU, tier 3: the GPU-backed MCTS player through the reference-shaped plugin chain
(GameLauncher -> createGameRules / createPlayer -> IGamePlayer::selectMove -> gai_loa_rollout)."""
import json
import os
import subprocess
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LAUNCHER = os.path.join(ROOT, "game-ai_b200", "host", "GameLauncher")
XML = os.path.join(ROOT, "run_config_loa.xml")


def run(*args):
    r = subprocess.run([LAUNCHER, "--xml", XML] + list(args), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_gpu_mcts_beats_random_and_shallow_alphabeta():
    # win-rate tolerance: a uniformly random mover (or a 2-ply search with the rules' constant eval, which is
    # what the reference's MinMaxAB sees on LOA) must lose essentially every game to 300 k playouts of MCTS per move
    # (fixed seeds and a simulation budget instead of a time budget: the games are reproducible)
    # A game the controller aborts (repeated position / round limit, GameController.cpp:59-97) is neither a win nor a
    # loss: with best_move_value_eps 0.05 a side that is winning everywhere picks among near-equal moves at random and
    # can shuffle into a repetition against a random mover, exactly as the reference player does.
    res = run("--p1", "mcts move_sim_limit=1000000 random_seed=21 philox_seed=22", "--p2", "random random_seed=23", "--ng", "4", "--seed", "3")
    assert res["players"][0]["lose"] == 0 and res["players"][0]["win"] >= 2, res
    assert res["players"][0]["win"] + res["aborted"] == res["games"], res
    stats = res["players"][0]["stats"]
    assert float(stats["playouts_per_sec"]) > 1e5 and float(stats["gpu_batches"]) > 0
    # (batch-by-expansion spends ~1100 playouts per selection: the same tree depth needs a larger playout budget)
    res = run("--p1", "ab3 search_depth=2 random_seed=31", "--p2", "mcts move_sim_limit=3000000 random_seed=32 philox_seed=33", "--ng", "3", "--seed", "4")
    assert res["players"][1]["lose"] == 0 and res["players"][1]["win"] >= 1, res


def test_fixed_simulation_budget_is_reproducible():
    a = run("--p1", "mcts move_sim_limit=65536 random_seed=7 philox_seed=9", "--p2", "random", "--ng", "1", "--seed", "8")
    b = run("--p1", "mcts move_sim_limit=65536 random_seed=7 philox_seed=9", "--p2", "random", "--ng", "1", "--seed", "8")
    assert a["rounds"] == b["rounds"] and a["players"][0]["pts"] == b["players"][0]["pts"]


def test_root_parallel_two_trees_one_process():
    # two independent trees (two contexts; on a 1-GPU box both on device 0) merged at the root
    res = run("--p1", "mcts move_time_limit=0.1 devices=0,0 random_seed=41 philox_seed=42", "--p2", "random random_seed=43", "--ng", "2", "--seed", "5")
    assert res["players"][0]["win"] >= 1, res


def test_determinized_gpu_pan_player_beats_lowcard():
    """Config 5 end to end (the reference's published experiment, results/mcts_player.log:2): the determinized GPU
    player in the MCTS seat against two `lowcard` players.  Tolerance: it must lose clearly fewer games than a
    lowcard player does on average."""
    r = subprocess.run([LAUNCHER, "--xml", os.path.join(ROOT, "run_config_pan.xml"), "--p1", "gpupan", "--p2", "lowcard", "--p3", "lowcard",
                        "--ng", "40", "--seed", "9"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    gp, lc1, lc2 = res["players"]
    assert float(gp["stats"]["playouts"]) > 1e5
    assert gp["lose"] < 0.5 * (lc1["lose"] + lc2["lose"]), res


def test_tree_reuse_and_gc():
    # a long search per move grows the tree past the GC threshold; the root of the next decision is found in the old
    # tree (reference: findRootNode, MCTSPlayer.cpp:219-251) and unreachable nodes are freed
    res = run("--p1", "mcts move_sim_limit=2000000 gpu_batch=2048 playouts_per_leaf=8 expand_all=0 random_seed=51 philox_seed=52", "--p2", "random random_seed=53", "--ng", "1", "--seed", "12")
    st = res["players"][0]["stats"]
    assert float(st["reused_root_visits"]) > 0 and float(st["tree_gc_runs"]) >= 1 and float(st["tree_gc_freed_nodes"]) > 0
    assert res["players"][0]["lose"] == 0
