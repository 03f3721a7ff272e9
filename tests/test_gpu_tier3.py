"""GPU, tier 3: the GPU-backed player against the reference's own MC::Player (c++/MCTSPlayer/MCTSPlayer.cpp compiled from
/root/reference into oracle/_ref/) at an EQUAL simulation budget -- move values, move choice and win rate.

Both searches are stochastic and the reference draws from std::default_random_engine (MCTSPlayer.cpp:472-476), so equality is
statistical.  Stated tolerances:

* values: the reference's per-edge visit counter is one byte (MCTSPlayer.h:51) and wraps at 255, so a single decision cannot be
  given 16 k simulations; instead R independent decisions of 512 simulations each (different seeds) are summed on both sides
  (>= 18 k simulations per position) and, for every root move with >= 200 summed visits on both sides, the win-rate estimates
  value / visits agree within 4 standard errors of their difference (binomial bound, sqrt(0.25/n_ref + 0.25/n_gpu)) + 0.02;
* choice: the move the GPU player chooses most often has a reference value within 0.05 + 4 standard errors (of the reference's
  own estimate for that move) of the reference's best move -- tight where the reference has searched the move deeply, loose in
  positions like the opening where all 36 moves are worth the same to within the noise;
* win rate: in a head-to-head match at 200 simulations per move (both colours) the GPU player's score is not below
  0.5 - 3 sigma - 0.05 (sigma = binomial standard error of the match); profiles/r02_tier3_match.json holds a longer run of the
  same pairing, the reference-vs-reference control and both players' aborted-game rates against a random mover.
The GPU player runs in its reference-like mode (expand_all=0 playouts_per_leaf=1: one playout per simulation, expand_size
honoured)."""
import json
import math
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LAUNCHER = os.path.join(ROOT, "game-ai_b200", "host", "GameLauncher")
XML = os.path.join(ROOT, "tests", "golden", "run_config_tier3.xml")
REFDIR = os.path.join(ROOT, "oracle", "_ref")

POSITIONS = [
    (0x0081818181818100, 0x7e0000000000007e, 1),           # the opening (LinesOfActionZasady.cpp:358-365)
    (0x10808c1082a0010, 0x100040408000, 1),                # black close to connecting: several winning moves
    (0x2088000a06031040, 0x28400010, 1),
    (0xa1200000000, 0x240102002200040, 0),
    (0x0000240018240000, 0x4200810000810042, 1),           # white massed in the centre, black on the rim
    (0x0000001c08000400, 0x0020400000281000, 0),
    (0x0001010000808000, 0x3c0000000000003c, 0),           # early middle game
    (0x0040001000020001, 0x0000000000e00700, 1),
]
SIMS, RUNS = 512, 36


def sq(name):
    return (ord(name[0]) - 65) + 8 * (int(name[1]) - 1)


def gpu_decision(pos, seed):
    w, b, cp = pos
    spec = "gpu move_sim_limit=%d random_seed=%d philox_seed=%d" % (SIMS, seed, 1000 + seed)
    args = [LAUNCHER, "--xml", XML, "--p1", spec if cp == 0 else "random", "--p2", spec if cp == 1 else "random",
            "--ng", "1", "--rl", "1", "--start", "%d,%d,%d" % (w, b, cp), "-q"]
    out = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-1500:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    stats = res["players"][cp]["stats"]["last_root_stats"]
    return [tuple(float(x) for x in e.split(":")) for e in stats.split(",")], res["first_move"]


@pytest.mark.parametrize("pos", POSITIONS, ids=["opening", "tactical1", "tactical2", "tactical3", "centre", "sparse", "early", "late"])
def test_values_and_choice_vs_reference_player(pos, oracle):
    from oracle import mcts
    ref = mcts.ref(fastpool=True)
    if ref is None:
        pytest.skip("oracle/_ref/libref_mcts_fastpool.so not built (needs /root/reference at build time)")
    w, b, cp = pos
    moves = oracle.movegen(w, b, cp)
    nm = len(moves)
    assert nm > 1 and not oracle.is_terminal(w, b)
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        refs = list(ex.map(lambda s: ref.select(w, b, cp, SIMS, seed=100 + s), range(RUNS)))
        gpus = list(ex.map(lambda s: gpu_decision(pos, 100 + s), range(RUNS)))
    rv = [0.0] * nm; rn = [0] * nm; gv = [0.0] * nm; gn = [0] * nm
    used = 0
    for r in refs:
        if len(r["visits"]) != nm or sum(r["visits"]) != SIMS:
            continue                                                     # a byte counter wrapped in this run
        used += 1
        for i in range(nm):
            rn[i] += r["visits"][i]; rv[i] += r["values"][i][cp]
    assert used >= RUNS // 2
    chosen = {}
    for g, fm in gpus:
        assert len(g) == nm
        for i in range(nm):
            gn[i] += int(g[i][0]); gv[i] += g[i][1 + cp]
        c = moves.index((sq(fm[0:2]), sq(fm[4:6])))
        chosen[c] = chosen.get(c, 0) + 1
    checked = 0
    for i in range(nm):
        if rn[i] < 200 or gn[i] < 200:
            continue
        pr, pg = rv[i] / rn[i], gv[i] / gn[i]
        tol = 4.0 * math.sqrt(0.25 / rn[i] + 0.25 / gn[i]) + 0.02
        assert abs(pr - pg) <= tol, (i, moves[i], rn[i], gn[i], pr, pg, tol)
        checked += 1
    assert checked >= min(nm, 4)
    # the move the GPU player prefers is (nearly) as good by the reference's own summed estimates
    fav = max(chosen, key=chosen.get)
    ref_val = [rv[i] / (1.0 + rn[i]) for i in range(nm)]
    assert ref_val[fav] >= max(ref_val) - 0.05 - 4.0 * math.sqrt(0.25 / max(1, rn[fav])), (moves[fav], ref_val[fav], max(ref_val), rn[fav])


def test_win_rate_vs_reference_player():
    if not os.path.exists(os.path.join(REFDIR, "librefmctsplayer.so")):
        pytest.skip("oracle/_ref/librefmctsplayer.so not built (needs /root/reference at build time)")
    games, sims = 20, 200

    def match(p1, p2, seed):
        r = subprocess.run([LAUNCHER, "--xml", XML, "--plugins", REFDIR, "--p1", p1, "--p2", p2, "--ng", str(games), "--threads", str(os.cpu_count() or 4),
                            "--seed", str(seed), "-q"], capture_output=True, text=True, timeout=1500)
        assert r.returncode == 0, r.stderr[-1500:]
        return json.loads(r.stdout.strip().splitlines()[-1])
    S = "move_sim_limit=%d" % sims
    a = match("gpu " + S, "ref " + S, 31)
    b = match("ref " + S, "gpu " + S, 32)
    gpu_w = a["players"][0]["win"] + b["players"][1]["win"]; ref_w = a["players"][1]["win"] + b["players"][0]["win"]
    n = 2 * games
    score = (gpu_w + 0.5 * (n - gpu_w - ref_w)) / n
    sigma = math.sqrt(0.25 / n)
    assert score >= 0.5 - 3 * sigma - 0.05, (gpu_w, ref_w, n, score)
