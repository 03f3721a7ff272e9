"""CPU tests of the host side above the C ABI: the LOA rules plugin (createGameRules), the launcher/controller and the
CPU opponents.  The GPU-backed player must refuse to play without a GPU (no CPU fallback)."""
import ctypes
import json
import os
import subprocess
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "game-ai_b200", "host")
LAUNCHER = os.path.join(HOST, "GameLauncher")
XML = os.path.join(ROOT, "run_config_loa.xml")


@pytest.fixture(scope="module")
def hostrules():
    L = ctypes.CDLL(os.path.join(HOST, "libLinesOfActionZasady.so"))
    L.loa_host_movegen.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p]
    L.loa_host_terminal_score.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p]
    L.loa_host_next.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    return L


def test_plugins_export_the_reference_entry_points():
    for lib, sym in (("libLinesOfActionZasady.so", "createGameRules"), ("libgpumctsplayer.so", "createPlayer"), ("libSimpleStrategyPlayer.so", "createPlayer")):
        L = ctypes.CDLL(os.path.join(HOST, lib))
        assert hasattr(L, sym), (lib, sym)


def test_host_rules_plugin_vs_oracle(hostrules, oracle):
    S = np.concatenate([oracle.gen_reachable(1500, 3), oracle.gen_uniform(1500, 3)])
    oc, om = oracle.movegen_batch(S); ot, osc = oracle.terminal_batch(S)
    buf = (ctypes.c_uint8 * 256)(); sc = (ctypes.c_int * 2)(); out = (ctypes.c_uint64 * 3)()
    for i, (w, b, cp) in enumerate(S.tolist()):
        n = hostrules.loa_host_movegen(w, b, cp, buf)
        assert n == oc[i] and [buf[k] for k in range(2 * n)] == om[i, :n].reshape(-1).tolist()
        assert hostrules.loa_host_terminal_score(w, b, sc) == ot[i] and [sc[0], sc[1]] == osc[i].tolist()
        if n:
            f, t = om[i, i % n].tolist()
            hostrules.loa_host_next(w, b, cp, f, t, out)                 # Next() = move of the mover + noop of the other
            assert (out[0], out[1], out[2] & 1) == oracle.apply(w, b, cp, f, t)


def run_launcher(*args):
    r = subprocess.run([LAUNCHER, "--xml", XML] + list(args), capture_output=True, text=True, timeout=600)
    return r


def test_launcher_cpu_players_are_deterministic():
    a = run_launcher("--p1", "random", "--p2", "ab3 search_depth=2", "--ng", "4", "--seed", "11")
    b = run_launcher("--p1", "random", "--p2", "ab3 search_depth=2", "--ng", "4", "--seed", "11")
    assert a.returncode == 0, a.stderr
    ra, rb = json.loads(a.stdout.strip().splitlines()[-1]), json.loads(b.stdout.strip().splitlines()[-1])
    assert ra["games"] == 4 and ra["players"][0]["name"] == "random" and ra["players"][1]["name"] == "ab3_search_depth=2"
    assert [p["pts"] for p in ra["players"]] == [p["pts"] for p in rb["players"]] and ra["rounds"] == rb["rounds"]
    assert ra["players"][1]["win"] >= 3                                   # a 2-ply search beats the random mover


def test_gpu_player_refuses_to_play_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = run_launcher("--p1", "mcts", "--p2", "random", "--ng", "1")
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


# ---- host logic of the GPU MCTS player with the GPU library replaced by a TEST stub (LD_PRELOAD, tests/helpers/fake_gpu.c in
# its "hash" mode: no games are played, every leaf gets a pseudo-random split of its playouts; tests/test_player_host.py uses
# the same stub with real (oracle) playouts) ----
@pytest.fixture(scope="module")
def fake_gpu():
    so = os.path.join(ROOT, "tests", "helpers", "libfakegpu.so")
    src = os.path.join(ROOT, "tests", "helpers", "fake_gpu.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-std=c11", "-O2", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), "-o", so, src,
                               "-L", os.path.join(ROOT, "oracle"), "-lloa_oracle", "-lpan_oracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    return so


def run_with_fake_gpu(fake_gpu, *args):
    env = dict(os.environ, LD_PRELOAD=fake_gpu, GAI_FAKE_MODE="hash")
    r = subprocess.run([LAUNCHER, "--xml", XML] + list(args), capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_mcts_host_logic_is_deterministic_and_reuses_its_tree(fake_gpu):
    """Selection with virtual loss, lazy expansion, two pipelined batches, transposition table, tree reuse and GC, final-move
    rule -- everything above the C ABI, with fixed seeds and a simulation budget: two runs must agree move for move."""
    args = ("--p1", "mcts move_sim_limit=400000 gpu_batch=512 playouts_per_leaf=8 expand_all=0 random_seed=3 philox_seed=4", "--p2", "random random_seed=5",
            "--ng", "1", "--rl", "40", "--seed", "6")
    a = run_with_fake_gpu(fake_gpu, *args); b = run_with_fake_gpu(fake_gpu, *args)
    assert a["rounds"] == b["rounds"] and a["first_move"] == b["first_move"]
    sa, sb = a["players"][0]["stats"], b["players"][0]["stats"]
    assert sa["last_root_stats"] == sb["last_root_stats"] and sa["runs_per_move"] == sb["runs_per_move"]
    assert float(sa["gpu_batches"]) > 100 and float(sa["reused_root_visits"]) > 0          # the next root was found in the old tree
    assert float(sa["tree_gc_runs"]) >= 1 and float(sa["tree_gc_freed_nodes"]) > 0         # 50 k selections per move: the tree outgrows the GC threshold
    # the merged root statistics are consistent: visits are whole batches of playouts, values never exceed visits
    for item in sa["last_root_stats"].split(","):
        v, w, bl = item.split(":")
        assert int(v) % 8 == 0 and float(w) + float(bl) <= int(v) + 1e-3


def test_root_parallel_trees_in_one_process(fake_gpu):
    res = run_with_fake_gpu(fake_gpu, "--p1", "mcts move_sim_limit=60000 devices=0,0,0 random_seed=3 philox_seed=4", "--p2", "random random_seed=5",
                            "--ng", "1", "--rl", "10", "--seed", "6")
    st = res["players"][0]["stats"]
    total = sum(int(x.split(":")[0]) for x in st["last_root_stats"].split(","))
    assert total >= 3 * (60000 // 3 // 32) * 32 * 0.9                                    # three trees' visits were merged at the root


# ---- Pan host rules plugin (createGameRules for GraWPanaZasadyV2) -----------------------------------------------
def test_pan_host_rules_plugin_vs_oracle():
    from oracle import pan
    H = ctypes.CDLL(os.path.join(HOST, "libGraWPanaZasadyV2.so"))
    vp = ctypes.c_void_p
    H.pan_host_movegen.argtypes = [ctypes.c_int, vp, vp]; H.pan_host_apply.argtypes = [ctypes.c_int, vp, ctypes.c_uint64, vp]
    H.pan_host_known_state.argtypes = [ctypes.c_int, vp, ctypes.c_int, vp]; H.pan_host_score_eval.argtypes = [ctypes.c_int, vp, vp, vp, vp]
    p = pan.port()
    for npl in (2, 3, 4):
        S = p.gen_reachable(npl, 500, 5, 100)
        t, sc, e0, e1 = p.terminal_batch(npl, S)
        for i in range(len(S)):
            for st in (S[i], p.player_known_state(npl, S[i], i % npl)):              # complete state and a fuzzy view
                mv = (ctypes.c_uint64 * 16)(); n = H.pan_host_movegen(npl, st.ctypes.data, mv)
                oc, om = p.movegen_batch(npl, st[None])
                assert n == oc[0] and list(mv[:n]) == [int(x) for x in om[0, :n]]
                if n:
                    out = np.zeros(5, np.uint64); H.pan_host_apply(npl, st.ctypes.data, mv[i % n], out.ctypes.data)
                    assert (out == p.apply_batch(npl, st[None], np.array([mv[i % n]], dtype=np.uint64))[0]).all()
            k = np.zeros(5, np.uint64); H.pan_host_known_state(npl, S[i].ctypes.data, i % npl, k.ctypes.data)
            assert (k == p.player_known_state(npl, S[i], i % npl)).all()
            s4 = (ctypes.c_int * 4)(); a4 = (ctypes.c_int * 4)(); b4 = (ctypes.c_int * 4)()
            assert H.pan_host_score_eval(npl, S[i].ctypes.data, s4, a4, b4) == t[i]
            assert list(s4[:npl]) == list(sc[i, :npl]) and list(a4[:npl]) == list(e0[i, :npl]) and list(b4[:npl]) == list(e1[i, :npl])


def test_pan_launcher_three_players():
    r = subprocess.run([LAUNCHER, "--xml", os.path.join(ROOT, "run_config_pan.xml"), "--p1", "random", "--p2", "lowcard", "--p3", "lowcard",
                        "--ng", "20", "--seed", "4"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert res["games"] == 20 and len(res["players"]) == 3
    finished = 20 - res["aborted"]
    assert sum(p["lose"] for p in res["players"]) == finished                    # exactly one loser per finished game
    assert sum(p["pts"] for p in res["players"]) == finished * 100 + res["aborted"] * 99       # 50+50 / 3 x 33 on an aborted game
