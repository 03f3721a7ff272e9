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
