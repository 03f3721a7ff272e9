#!/usr/bin/env python
"""Tier 3, win rate: the GPU-backed player in its reference-like mode against the reference's own MC::Player (test-only plugin
oracle/_ref/librefmctsplayer.so) at the same simulation budget, both colours, with reference-vs-reference as the control and both
against a random mover (aborted-game rates).  MEASUREMENT SCRIPT (test infrastructure: it loads oracle/_ref).
usage: python tests/tools/tier3_match.py [games_per_pairing] [sims] [threads] [--fake]   (--fake: CPU box, device replaced by the oracle-backed stub)"""
import json, os, subprocess, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
L = os.path.join(ROOT, "game-ai_b200", "host", "GameLauncher")
XML = os.path.join(ROOT, "tests", "golden", "run_config_tier3.xml")
args = [a for a in sys.argv[1:] if not a.startswith("--")]
games = int(args[0]) if len(args) > 0 else 64
sims = int(args[1]) if len(args) > 1 else 200
threads = int(args[2]) if len(args) > 2 else (os.cpu_count() or 1)
env = dict(os.environ)
if "--fake" in sys.argv:
    env["LD_PRELOAD"] = os.path.join(ROOT, "tests", "helpers", "libfakegpu.so")


def _limit():
    # the reference's MC::Player (test-only plugin build) occasionally runs away in memory in the middle of a match (seen against a
    # random mover and against the GPU player: +75 MB/s without end, 47 GB in eight minutes with six controller threads); an
    # address-space limit and a timeout per chunk turn that into a failed chunk instead of a dead box
    import resource
    resource.setrlimit(resource.RLIMIT_AS, (24 << 30, 24 << 30))


def match(p1, p2, seed):
    """`games` games in chunks of `threads` (a fresh launcher process per chunk, one game per controller thread)"""
    tot = {"p1": p1, "p2": p2, "games": 0, "aborted": 0, "p1_win": 0, "p2_win": 0, "p1_pts": 0.0, "p2_pts": 0.0, "seconds": 0.0, "failed_chunks": 0}
    done = 0; chunk = 0
    while done < games:
        n = min(threads, games - done)
        try:
            r = subprocess.run([L, "--xml", XML, "--plugins", os.path.join(ROOT, "oracle", "_ref"), "--p1", p1, "--p2", p2, "--ng", str(n),
                                "--threads", str(n), "--seed", str(seed * 1000 + chunk), "-q"], capture_output=True, text=True, env=env, timeout=900, preexec_fn=_limit)
            ok = r.returncode == 0
        except subprocess.TimeoutExpired:
            ok = False
        done += n; chunk += 1
        if not ok:
            tot["failed_chunks"] += 1
            continue
        d = json.loads(r.stdout.strip().splitlines()[-1])
        tot["games"] += d["games"]; tot["aborted"] += d["aborted"]; tot["seconds"] += d["seconds"]
        tot["p1_win"] += d["players"][0]["win"]; tot["p2_win"] += d["players"][1]["win"]
        tot["p1_pts"] += d["players"][0]["pts"]; tot["p2_pts"] += d["players"][1]["pts"]
    return tot


S = "move_sim_limit=%d" % sims
out = {"sims_per_move": sims, "games_per_pairing": games, "device": "fake (oracle playouts)" if "--fake" in sys.argv else "gpu", "pairings": []}
for p1, p2, seed in (("gpu " + S, "ref " + S, 101), ("ref " + S, "gpu " + S, 102), ("ref " + S, "ref " + S + " random_seed=977", 103),
                     ("gpu " + S, "random", 104), ("ref " + S, "random", 105)):
    m = match(p1, p2, seed)
    out["pairings"].append(m)
    print(json.dumps(m), file=sys.stderr, flush=True)
P = out["pairings"]
# score of a side = (wins + half of the draws and aborted games) / games
def score(wins, other_wins, n): return (wins + 0.5 * (n - wins - other_wins)) / n
ng = P[0]["games"] + P[1]["games"]
g = score(P[0]["p1_win"] + P[1]["p2_win"], P[0]["p2_win"] + P[1]["p1_win"], max(1, ng))
c = score(P[2]["p1_win"], P[2]["p2_win"], max(1, P[2]["games"]))
out["gpu_vs_ref_score"] = g
out["gpu_vs_ref_games"] = ng
out["gpu_vs_ref_sigma"] = math.sqrt(0.25 / max(1, ng))
out["control_ref_as_white_score"] = c
out["aborted_rate_vs_random"] = {"gpu": P[3]["aborted"] / max(1, P[3]["games"]), "ref": P[4]["aborted"] / max(1, P[4]["games"])}
out["win_rate_vs_random_of_finished"] = {"gpu": P[3]["p1_win"] / max(1, P[3]["games"] - P[3]["aborted"]), "ref": P[4]["p1_win"] / max(1, P[4]["games"] - P[4]["aborted"])}
print(json.dumps(out, indent=1))
