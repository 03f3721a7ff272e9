"""GPU, two devices (skipped on a one-GPU box): config 4 -- root-parallel MCTS, one process per GPU, every rank its own
tree and Philox stream, root child statistics summed with NCCL (gai_allreduce_root) once per move decision, after which
every rank applies the reference's final-move rule (selectBestMove, MCTSPlayer.cpp:601-632) to identical numbers with the
same tie-break seed.  The two ranks must therefore play the SAME game move for move."""
import json
import os
import subprocess
import tempfile
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LAUNCHER = os.path.join(ROOT, "game-ai_b200", "host", "GameLauncher")
XML = os.path.join(ROOT, "run_config_loa.xml")


def test_two_ranks_play_the_same_game():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    with tempfile.TemporaryDirectory() as tmp:
        idf = os.path.join(tmp, "nccl_id")
        procs = []
        for rank in range(2):
            env = dict(os.environ, WORLD_SIZE="2", RANK=str(rank), LOCAL_RANK=str(rank), GAI_NCCL_ID_FILE=idf)
            procs.append(subprocess.Popen([LAUNCHER, "--xml", XML, "--p1", "mcts devices= move_sim_limit=131072 random_seed=5 philox_seed=6",
                                           "--p2", "random random_seed=7", "--ng", "1", "--seed", "11", "-v"],
                                          env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
        outs = [p.communicate(timeout=900) for p in procs]
        for p, (o, e) in zip(procs, outs):
            assert p.returncode == 0, e[-2000:]
        moves = [[l for l in o.splitlines() if l.startswith("round ")] for o, _ in outs]
        assert len(moves[0]) > 10 and moves[0] == moves[1]                 # identical move sequence on both ranks
        res = [json.loads(o.strip().splitlines()[-1]) for o, _ in outs]
        assert res[0]["rounds"] == res[1]["rounds"] and res[0]["players"][0]["pts"] == res[1]["players"][0]["pts"]
        # each rank searched its own tree: the per-rank playout counts are equal (same budget), the merged root carries both
        st = [r["players"][0]["stats"] for r in res]
        assert st[0]["last_root_stats"] == st[1]["last_root_stats"]
        assert res[0]["players"][0]["lose"] == 0


def test_one_rank_goes_through_the_nccl_merge_on_a_one_gpu_box():
    """The same merge path with a communicator of ONE rank (GAI_FORCE_COMM): ncclCommInitRank + the per-move
    gai_allreduce_root run on whatever single GPU the box has, and -- the sum over one rank being the identity -- the game is
    move for move the one played without a communicator."""
    def play(extra_env):
        env = dict(os.environ, **extra_env)
        r = subprocess.run([LAUNCHER, "--xml", XML, "--p1", "mcts devices=0 move_sim_limit=65536 random_seed=5 philox_seed=6",
                            "--p2", "random random_seed=7", "--ng", "1", "--rl", "16", "--seed", "11", "-v"], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        return [l for l in r.stdout.splitlines() if l.startswith("round ")], json.loads(r.stdout.strip().splitlines()[-1])
    with tempfile.TemporaryDirectory() as tmp:
        m1, r1 = play({"WORLD_SIZE": "1", "RANK": "0", "LOCAL_RANK": "0", "GAI_FORCE_COMM": "1", "GAI_NCCL_ID_FILE": os.path.join(tmp, "nccl_id")})
    m0, r0 = play({})
    assert len(m1) >= 8 and m1 == m0
    assert float(r1["players"][0]["stats"]["allreduce_us_per_move"]) > 0.0
    assert float(r0["players"][0]["stats"]["allreduce_us_per_move"]) == 0.0
