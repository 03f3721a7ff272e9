"""CPU tests of the HOST logic of the GPU-backed MCTS player (game-ai_b200/host/gpu_mcts_player.cpp): selection, virtual
loss, batch-by-expansion, pipelining, back-propagation, config keys and statistics, with the device replaced by
tests/helpers/fake_gpu.c -- a TEST-ONLY stand-in for libgameai_gpu.so that plays the batches through the oracle on the CPU
(LD_PRELOADed under GameLauncher).  The product library has no such path (tests/test_abi.py::test_no_cpu_fallback)."""
import json
import os
import subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "game-ai_b200", "host")
LAUNCHER = os.path.join(HOST, "GameLauncher")
XML = os.path.join(ROOT, "run_config_loa.xml")
REF_XML = "/root/reference/c++/run_config_loa.xml"


@pytest.fixture(scope="module")
def fake():
    so = os.path.join(ROOT, "tests", "helpers", "libfakegpu.so")
    src = os.path.join(ROOT, "tests", "helpers", "fake_gpu.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-std=c11", "-O2", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), "-o", so, src,
                               "-L", os.path.join(ROOT, "oracle"), "-lloa_oracle", "-lpan_oracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    return so


def run(fake, *args, xml=XML, check=True):
    env = dict(os.environ, LD_PRELOAD=fake)
    r = subprocess.run([LAUNCHER, "--xml", xml] + list(args), capture_output=True, text=True, timeout=900, env=env)
    if check:
        assert r.returncode == 0, r.stderr[-2000:]
        return json.loads(r.stdout.strip().splitlines()[-1])
    return r


TACTICAL = "0x10808c1082a0010,0x100040408000,1"      # black to move, close to connecting (tests/test_gpu_tier3.py)


@pytest.mark.parametrize("mode", ["expand_all=1 playouts_per_leaf=2 gpu_batch=16", "expand_all=0 playouts_per_leaf=4 gpu_batch=16",
                                  "expand_all=0 playouts_per_leaf=1 gpu_batch=8", "expand_all=1 playouts_per_leaf=2 gpu_batch=16 pipeline_depth=3",
                                  "expand_all=1 playouts_per_leaf=2 gpu_batch=16 pipeline_depth=1"])
def test_every_mode_finds_the_winning_move_and_accounts_for_every_playout(fake, mode):
    res = run(fake, "--p1", "random random_seed=3", "--p2", "mcts move_sim_limit=4000 random_seed=5 philox_seed=6 " + mode,
              "--ng", "1", "--rl", "1", "--start", TACTICAL)
    st = res["players"][1]["stats"]
    root = [tuple(float(x) for x in e.split(":")) for e in st["last_root_stats"].split(",")]
    # every edge: value_white + value_black == visits (each playout backs up 1 in total: win/loss 1+0, draw or cap 0.5+0.5)
    for v, a, b in root:
        assert abs(a + b - v) < 1e-3
    total = sum(v for v, _, _ in root)
    assert 4000 <= total <= 4000 + 16 * 96 * 4                 # the budget, overshot by at most one batch
    best = max(root, key=lambda e: e[2] / (1 + e[0]))
    assert best[2] / best[0] > 0.9                             # a move that wins (nearly) every playout exists and was found
    assert float(st["terminal_node_found_ratio"]) == 1.0


def test_fixed_budget_is_reproducible_with_pipelining(fake):
    a = run(fake, "--p1", "mcts move_sim_limit=1500 playouts_per_leaf=1 gpu_batch=8 random_seed=7 philox_seed=9", "--p2", "random random_seed=4", "--ng", "1", "--rl", "6", "--seed", "8", "-v")
    b = run(fake, "--p1", "mcts move_sim_limit=1500 playouts_per_leaf=1 gpu_batch=8 random_seed=7 philox_seed=9", "--p2", "random random_seed=4", "--ng", "1", "--rl", "6", "--seed", "8", "-v")
    assert a["players"][0]["stats"]["last_root_stats"] == b["players"][0]["stats"]["last_root_stats"]
    assert float(a["players"][0]["stats"]["find_root_node_ratio"]) >= 0.0


def test_expand_size_is_honoured_in_the_reference_like_mode(fake):
    """expansion (MCTSPlayer.cpp:524-562): with expand_size = E the first E positions of the playout below the leaf become
    permanent nodes.  More permanent nodes per simulation = a bigger tree after the same number of simulations."""
    sizes = {}
    for e in (0, 3):
        res = run(fake, "--p1", "random random_seed=3", "--p2",
                  "mcts move_sim_limit=600 playouts_per_leaf=1 expand_all=0 gpu_batch=4 expand_size=%d random_seed=5 philox_seed=6" % e,
                  "--ng", "1", "--rl", "1", "--start", "0x0081818181818100,0x7e0000000000007e,1")
        st = res["players"][1]["stats"]
        assert float(st["expand_size_effective"]) == e
        sizes[e] = sum(int(kv.split(":")[0]) * int(kv.split(":")[1]) for kv in st["node_pool_usage"].split(","))
    assert sizes[3] > 2.5 * max(sizes[0], 1000) or sizes[3] >= sizes[0] + 1000


def test_reference_config_line_loads_unmodified_except_provider(fake, tmp_path):
    """The <player name="mcts"> line of the reference's own c++/run_config_loa.xml, provider swapped, nothing else."""
    if not os.path.exists(REF_XML):
        pytest.skip("needs /root/reference")
    text = open(REF_XML, encoding="utf-8-sig").read()
    assert 'provider="mctsplayer"' in text
    text = text.replace('provider="mctsplayer"', 'provider="gpumctsplayer"')
    text = text.replace('out_dir="c:\\MyData\\Projects\\gra_w_pana\\dbg_logs"', 'out_dir="%s"' % tmp_path)      # a directory that exists here
    xml = tmp_path / "run_config_loa.xml"
    xml.write_text(text)
    res = run(fake, "--p1", "random", "--p2", "mcts move_sim_limit=1200", "--ng", "1", "--rl", "2", xml=str(xml))
    st = res["players"][1]["stats"]
    for k in ("runs_per_move", "branching_factor", "path_size", "node_pool_usage", "find_root_node_ratio", "terminal_node_found_ratio", "best_move_is_most_visited_ratio"):
        assert k in st                                           # getGameStats names of MCTSPlayer.cpp:149-165
    assert (tmp_path / "mcts.log").read_text().count("chosen") == 1   # trace="mcts.log" in out_dir


def test_unknown_and_unsupported_keys_are_refused(fake):
    r = run(fake, "--p1", "random", "--p2", "mcts move_sim_limit=100 exploration=3", "--ng", "1", "--rl", "1", check=False)
    assert r.returncode != 0 and "unknown config key 'exploration'" in r.stderr
    r = run(fake, "--p1", "random", "--p2", "mcts move_sim_limit=100 type=mcrl", "--ng", "1", "--rl", "1", check=False)
    assert r.returncode != 0 and "mcrl" in r.stderr
    # the reference's way of commenting an attribute out is accepted
    res = run(fake, "--p1", "random", "--p2", "mcts move_sim_limit=300 _move_time_limit=5 playouts_per_leaf=1", "--ng", "1", "--rl", "1")
    assert res["games"] == 1


def test_cycle_score_and_terminal_root(fake):
    # a terminal root: the player hands back the rules' list instead of searching (no crash, no move from an empty list)
    r = run(fake, "--p1", "random", "--p2", "mcts move_sim_limit=100", "--ng", "1", "--rl", "1", "--start", "0x3,0x0c00000000000000,1", check=False)
    assert r.returncode == 0, r.stderr[-500:]


PAN_XML = os.path.join(ROOT, "run_config_pan.xml")


@pytest.mark.parametrize("mode", ["tree", "flat"])
def test_pan_determinized_player_host_logic(fake, mode):
    """gpu_pan_player.cpp on the fake device: the shared tree over sampled deals (and round 1's flat evaluation) runs a
    three-player game against two lowcard players, every decision gets its deals (no determinization fallback on fresh
    games), the playouts the device reports are the ones the player asked for, and a fixed seed replays the same game."""
    spec = "gpupan mode=%s move_sim_limit=256 gpu_batch=32 deals=16 playouts_per_leaf=4 random_seed=9 philox_seed=10" % mode
    a = run(fake, "--p1", spec, "--p2", "lowcard", "--p3", "lowcard", "--ng", "1", "--rl", "24", "--seed", "6", "-q", xml=PAN_XML)
    b = run(fake, "--p1", spec, "--p2", "lowcard", "--p3", "lowcard", "--ng", "1", "--rl", "24", "--seed", "6", "-q", xml=PAN_XML)
    st = a["players"][0]["stats"]
    assert float(st["decisions"]) > 0 and float(st["determinization_fallbacks"]) == 0 and float(st["devices"]) == 1
    assert float(st["playouts"]) > 0 and float(st["playouts"]) % 4 == 0
    if mode == "tree":
        assert float(st["iterations"]) >= float(st["decisions"]) and float(st["tree_nodes"]) >= 1
    assert a["rounds"] == b["rounds"] and a["first_move"] == b["first_move"]
    assert [p["pts"] for p in a["players"]] == [p["pts"] for p in b["players"]]
    assert st["playouts"] == b["players"][0]["stats"]["playouts"]


def test_single_rank_communicator_is_the_identity(fake, tmp_path):
    """GAI_FORCE_COMM: one rank through the merge path (communicator init, per-move all-reduce with its position checksum) plays
    the same game as no communicator at all; tests/test_gpu_multi.py runs the same pair against NCCL on a GPU."""
    args = ("--p1", "mcts devices=0 move_sim_limit=1024 random_seed=5 philox_seed=6", "--p2", "random random_seed=7", "--ng", "1", "--rl", "8", "--seed", "11", "-v")
    def play(extra):
        env = dict(os.environ, LD_PRELOAD=fake, **extra)
        r = subprocess.run([LAUNCHER, "--xml", XML] + list(args), capture_output=True, text=True, timeout=600, env=env)
        assert r.returncode == 0, r.stderr[-2000:]
        return [l for l in r.stdout.splitlines() if l.startswith("round ")], json.loads(r.stdout.strip().splitlines()[-1])
    m1, r1 = play({"WORLD_SIZE": "1", "RANK": "0", "GAI_FORCE_COMM": "1", "GAI_NCCL_ID_FILE": str(tmp_path / "id")})
    m0, r0 = play({})
    assert len(m0) >= 4 and m0 == m1
    assert float(r1["players"][0]["stats"]["allreduce_us_per_move"]) > 0 and float(r0["players"][0]["stats"]["allreduce_us_per_move"]) == 0
