#!/usr/bin/env python
"""bench.py -- LOA random playouts/s on N B200s (BASELINE.json metric), one JSON line on rank 0.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--leaves L] [--ppl P] [--workload loa|pan]

A "step" is one pass of the hot path over one synthetic leaf batch: L reachable LOA positions (SURVEY.md 8d
generator, made on the device), P uniformly random playouts per leaf to the end of the game
(max_plies = 50 000, run_config_loa.xml:6).  N > 1: launched by torchrun, one rank per GPU, every rank plays its
own L-leaf shard (leaf parallelism has no data-path collective).  value = all ranks' playouts / max-over-ranks time.

  value        device-resident: leaves already in HBM, timed with CUDA events on the launching stream
  e2e          the plugin-facing C-ABI call gai_loa_rollout() with HOST buffers (H2D + kernel + D2H inside)
  roofline     issue-slot roofline of the rollout kernel (the path is integer/branch bound);
               an "hbm" sub-object carries the byte roofline for completeness
  player       config 4: the root-parallel GPU MCTS player (host tree + leaf batches through the same C ABI) from the
               opening at 1 s per move, one rank per GPU, root statistics merged with NCCL (gai_allreduce_root) once per
               move decision -- the collective sits inside this measured loop; every rank must play the same moves
  pan          config 5: determinized Pan rollouts (262 144 sampled deals x 16 playouts, 3 players, playout_depth 50),
               deals sharded over the ranks like the LOA leaves
  cpu_baseline (N = 1 only) the reference's own rules (oracle/_ref, compiled from /root/reference in the build
               container) or the plain-C port, driven by the same Philox stream on the host cores, and the reference's
               own MC::Player (selection_playOut) at 500 / 2 000 simulations and at 1 s per move
  --impl reference: the CPU arm alone, on the same config keys
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

SEED = 0x10A5EED
MAX_PLIES = 50000
METRIC = "loa_random_playouts_per_sec"
UNIT = "playouts/s"
TARGET = 1e9                       # BASELINE.json north_star: >= 1e9 LOA playouts/s per B200
LAUNCHER = os.path.join(ROOT, "game-ai_b200", "host", "GameLauncher")
OPENING = (0x0081818181818100, 0x7e0000000000007e, 1)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="loa", choices=["loa", "pan"])
    ap.add_argument("--leaves", type=int, default=1 << 20)
    ap.add_argument("--ppl", type=int, default=16)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-player", action="store_true", help="skip the config-4 player leg")
    ap.add_argument("--no-pan", action="store_true", help="skip the config-5 Pan leg")
    ap.add_argument("--player-moves", type=int, default=12, help="move decisions of the GPU player in the player leg")
    ap.add_argument("--player-trees", type=int, default=2, help="root-parallel trees per GPU in the player leg (one host thread, context and stream each)")
    ap.add_argument("--player-args", default="", help="extra key=value tokens for the player leg's mcts player")
    ap.add_argument("--threads-per-block", type=int, default=0)
    ap.add_argument("--blocks-per-sm", type=int, default=0)
    ap.add_argument("--rules-impl", type=int, default=2, help="0 = bitboard baseline kernel, 1 = rotated boards + line LUT per piece, 2 = static line sweep")
    return ap.parse_args()


def workload_config(args, world):
    """The workload, in keys both arms print identically."""
    return {"workload": "LOA leaf-parallel rollouts: %d reachable leaves x %d playouts/leaf per GPU to game end (max_plies %d)" % (args.leaves, args.ppl, MAX_PLIES),
            "leaves_per_gpu": args.leaves, "playouts_per_leaf": args.ppl, "max_plies": MAX_PLIES, "philox_key": SEED,
            "l2": "flushed between timed iterations (256 MiB write)",
            "parallelism": "leaf-parallel x%d, no data-path collective" % world}


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region (B200_PROFILING.md "clocks line")
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], 0, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "samples": len(sm),
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons)}


def physical_gpu_index(local):
    cvd = os.environ.get("CUDA_VISIBLE_DEVICES", "").strip()
    if cvd:
        try:
            return int(cvd.split(",")[local])
        except (ValueError, IndexError):
            return local
    return local


# ------------------------------------------------------------------------------------------------
# CPU baselines: the reference's rules / the reference's MC::Player on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_playouts(states, ppl, first_id, threads, budget_s):
    """Times Philox-driven playouts of the reference rules on `threads` host threads (ctypes drops the GIL).
    Returns dict(value playouts/s, playouts, plies, seconds, cores, kind, sample)."""
    from oracle import loa
    orc = loa.best()
    # probe one small chunk to size the sample for ~budget_s of wall time
    probe = states[:64]
    t0 = time.perf_counter(); _, pl = orc.rollout_states(probe, 1, MAX_PLIES, SEED, first_id); dt = time.perf_counter() - t0
    per_thread_rate = 64 / max(dt, 1e-6)
    want = int(per_thread_rate * threads * budget_s)
    n_leaves = max(threads * 8, min(len(states), max(1, want // ppl)))
    chunks = np.array_split(np.arange(n_leaves), threads * 4)
    res = [None] * len(chunks)

    def work(ci):
        idx = chunks[ci]
        if len(idx) == 0:
            res[ci] = (0, 0); return
        c, plies = orc.rollout_states(states[idx[0]:idx[-1] + 1], ppl, MAX_PLIES, SEED, first_id + int(idx[0]) * ppl)
        res[ci] = (int(c.sum()), plies)

    from concurrent.futures import ThreadPoolExecutor
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(work, range(len(chunks))))
    dt = time.perf_counter() - t0
    playouts = sum(r[0] for r in res); plies = sum(r[1] for r in res)
    return {"value": playouts / dt, "unit": UNIT, "cores": threads, "kind": orc.kind, "playouts": playouts, "plies": plies, "seconds": dt,
            "plies_per_sec": plies / dt, "sample_leaves": n_leaves,
            "sample": "first %d leaves of the same corpus x %d playouts/leaf, same Philox key and ids, %d host threads" % (n_leaves, ppl, threads)}


_REF_PLAYER_SNIPPET = r"""
import json, sys
sys.path.insert(0, %r)
from oracle import mcts
r = mcts.ref().select(%d, %d, %d, %d, seed=1234 + int(sys.argv[1]))
print(json.dumps({"sims_per_sec": r["sims_per_sec"], "seconds": r["seconds"], "mean_path": r["mean_path"]}))
"""


def reference_mctsplayer_baseline():
    """The reference's own MC::Player through selection_playOut (MCTSPlayer.cpp:387-470), run_config_loa.xml parameters, LOA
    opening: 500 and 2 000 simulations and 1 s per move on one core, and 500 simulations as `nproc` independent instances
    (the reference's own scaling model: whole games on more threads, GameController.cpp:167).  SURVEY.md 8d (i)."""
    from oracle import mcts
    rm = mcts.ref()
    if rm is None:
        return {"unavailable": "oracle/_ref/libref_mcts.so was not built (needs /root/reference at build time)"}
    w, b, cp = OPENING
    out = {"what": "reference MC::Player (oracle/_ref/libref_mcts.so: MCTSPlayer.cpp + MCplayer_factory.cpp + LinesOfActionZasady.cpp), run_config_loa.xml parameters, opening position",
           "runs": []}
    for sims in (500, 2000):
        r = rm.select(w, b, cp, sims, seed=1234)
        out["runs"].append({"limit": "move_sim_limit=%d" % sims, "cores": 1, "simulations_per_sec": r["sims_per_sec"], "mean_path_plies": r["mean_path"],
                            "plies_per_sec": r["sims_per_sec"] * r["mean_path"], "seconds": r["seconds"]})
    r = rm.select_timed(w, b, cp, 1.0, seed=1234)
    out["runs"].append({"limit": "move_time_limit=1", "cores": 1, "simulations": r["sims"], "simulations_per_sec": r["sims_per_sec"], "mean_path_plies": r["mean_path"],
                        "plies_per_sec": r["sims_per_sec"] * r["mean_path"], "seconds": r["seconds"]})
    nproc = os.cpu_count() or 1
    t0 = time.perf_counter()
    procs = [subprocess.Popen([sys.executable, "-c", _REF_PLAYER_SNIPPET % (ROOT, w, b, cp, 500), str(i)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
             for i in range(nproc)]
    rs = []
    for p in procs:
        o, _ = p.communicate(timeout=600)
        try:
            rs.append(json.loads(o.strip().splitlines()[-1]))
        except Exception:
            pass
    wall = time.perf_counter() - t0
    if rs:
        busy = max(x["seconds"] for x in rs)
        out["runs"].append({"limit": "move_sim_limit=500", "cores": nproc, "instances": len(rs), "simulations_per_sec": 500 * len(rs) / busy,
                            "plies_per_sec": 500 * len(rs) / busy * float(np.mean([x["mean_path"] for x in rs])), "seconds": busy, "wall_seconds_with_process_start": wall})
    out["simulations_per_sec"] = out["runs"][0]["simulations_per_sec"]           # the 500-simulation, one-core figure (round-1 key)
    return out


def run_reference(args, out):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from oracle import loa
    port = loa.port()
    threads = os.cpu_count() or 1
    # the corpus is defined by the oracle's generator (identical to the device generator, see tests); a bounded sample of it
    n_gen = min(args.leaves, 4096)
    states = port.gen_reachable(n_gen, SEED, 300)
    per_step = max(1.0, min(args.cpu_seconds, 60.0 / max(1, args.steps + args.warmup)))
    for _ in range(min(args.warmup, 1)):
        cpu_playouts(states, args.ppl, 0, threads, 0.5)
    tot_p = tot_pl = 0; tot_t = 0.0; last = None
    for _ in range(args.steps):
        last = cpu_playouts(states, args.ppl, 0, threads, per_step)
        tot_p += last["playouts"]; tot_pl += last["plies"]; tot_t += last["seconds"]
    value = tot_p / tot_t
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic", "config": workload_config(args, world),
            "sample_leaves_per_step": last["sample_leaves"],
            "note": "each step times a bounded sample of the workload (the first sample_leaves_per_step leaves of the corpus) on the host cores",
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["cores"], "kind": last["kind"], "sample": last["sample"],
                             "plies_per_sec": tot_pl / tot_t, "mean_plies_per_playout": tot_pl / max(1, tot_p)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "target_frac": value / TARGET, "gpu_launches": 0}
    print(json.dumps(line), file=out, flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# config 4: the root-parallel GPU MCTS player
# ------------------------------------------------------------------------------------------------
def player_leg(args, rank, world, local, dist):
    """Every rank runs the GPU-backed player plugin through GameLauncher (the reference's plugin chain) from the opening
    against a fixed-seed random mover: `player_moves` move decisions at 1 s each; ranks > 1 merge root statistics with NCCL
    once per move (inside this loop) and must all play the same game."""
    import tempfile
    idf = [os.path.join(tempfile.gettempdir(), "gai_nccl_id_%d_%d" % (os.getpid(), int(time.time() * 1e3)))]
    if dist:
        dist.broadcast_object_list(idf, src=0)
    env = dict(os.environ, WORLD_SIZE=str(world), RANK=str(rank), LOCAL_RANK=str(local), GAI_NCCL_ID_FILE=idf[0])
    spec = "mcts devices= trees_per_gpu=%d move_time_limit=1 random_seed=5 philox_seed=6 " % args.player_trees + args.player_args
    cmd = [LAUNCHER, "--xml", os.path.join(ROOT, "run_config_loa.xml"), "--p1", "random random_seed=7", "--p2", spec.strip(),
           "--ng", "1", "--rl", str(2 * args.player_moves), "--seed", "11", "-v"]
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=120 + 3 * args.player_moves)
    wall = time.perf_counter() - t0
    mine = {"rank": rank, "rc": r.returncode, "wall_s": wall}
    if r.returncode == 0:
        res = json.loads(r.stdout.strip().splitlines()[-1])
        st = res["players"][1]["stats"]
        mine.update({"moves": [l.split(": ", 1)[1] for l in r.stdout.splitlines() if l.startswith("round ") and "player 1" in l],
                     "playouts_per_sec": float(st["playouts_per_sec"]), "gpu_busy_fraction": float(st["gpu_busy_fraction"]),
                     "selections_per_sec": float(st["selections_per_sec"]), "plies_per_playout": float(st["plies_per_playout"]),
                     "allreduce_us_per_move": float(st["allreduce_us_per_move"]), "playouts_per_move_histogram": st["runs_per_move"],
                     "pending_collisions": float(st["pending_collisions"]), "gpu_batches": float(st["gpu_batches"])})
    else:
        mine["error"] = r.stderr[-400:]
    allr = [mine]
    if dist:
        allr = [None] * world
        dist.all_gather_object(allr, mine)
    if rank != 0:
        return None
    ok = all(x["rc"] == 0 for x in allr)
    out = {"what": "config 4: root-parallel gpumctsplayer (run_config_loa.xml 'mcts' parameters, provider swapped), Blacks, from the opening vs a fixed-seed random mover; "
                   "%d move decisions at move_time_limit=1; one rank per GPU, %d root-parallel tree(s) per rank (own host thread, context, stream and Philox stream each, summed at the root), "
                   "gai_allreduce_root (NCCL) across ranks once per move decision" % (args.player_moves, args.player_trees), "trees_per_gpu": args.player_trees,
           "ranks": allr, "n_ranks": world, "ok": ok}
    if ok:
        out["total_playouts_per_sec"] = sum(x["playouts_per_sec"] for x in allr)
        out["mean_gpu_busy_fraction"] = float(np.mean([x["gpu_busy_fraction"] for x in allr]))
        out["same_move_sequence"] = all(x["moves"] == allr[0]["moves"] for x in allr)
        out["move_decisions"] = len(allr[0]["moves"])
        assert out["same_move_sequence"], "root-parallel ranks played different games: " + json.dumps([x["moves"] for x in allr])
    return out


# ------------------------------------------------------------------------------------------------
# config 5: determinized Pan rollouts
# ------------------------------------------------------------------------------------------------
PAN_DEALS, PAN_PPL, PAN_PLAYERS, PAN_DEPTH, PAN_EVAL = 1 << 18, 16, 3, 50, 1


def pan_root_view():
    """What CreatePlayerKnownState (GraWPanaZasadyV2.cpp:214-243) gives player 0 at the start of a 3-player game: eight own
    cards (certain), an empty stack, eight cards in each other hand; the 9 of hearts is not ours, so player 1 (who holds
    it, :127-130) is to move.  gai_pan_determinize only reads the stack, the own hand, the public counts and the mover."""
    own = 0
    for c in (1, 4, 7, 10, 13, 16, 19, 22):
        own |= 3 << (2 * c)
    k = np.zeros(5, np.uint64)
    k[0] = np.uint64(1 << 48)                       # stack empty, current_player = 1
    k[1] = np.uint64(8 | (own << 4))
    k[2] = np.uint64(8); k[3] = np.uint64(8)
    return k


def pan_leg(args, eng, rank, world, dist, steps, warmup, with_cpu):
    import gameai_b200 as g
    import torch
    n = PAN_DEALS
    deals = g.pan_determinize(pan_root_view(), PAN_PLAYERS, 0, n, SEED + 5, rank * n)       # every rank samples its own deals
    first = rank * n * PAN_PPL
    # value: deals resident in HBM, results stay there (gai_pan_rollout_device, one pass)
    d_s = eng.dev_alloc(n * 40); d_r = eng.dev_alloc(n * 48)
    eng.h2d(d_s, deals)
    for _ in range(max(1, warmup)):
        eng.pan_rollout_device(PAN_PLAYERS, d_s, n, d_r, PAN_PPL, PAN_DEPTH, PAN_EVAL, SEED, first)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    k_ms = 0.0; plies = playouts = 0
    for _ in range(steps):
        st = eng.pan_rollout_device(PAN_PLAYERS, d_s, n, d_r, PAN_PPL, PAN_DEPTH, PAN_EVAL, SEED, first)
        k_ms += st["kernel_ms"]; plies += st["plies"]; playouts += st["playouts"]
    res_dev = np.zeros((n, 12), np.uint32); eng.d2h(res_dev, d_r)
    eng.dev_free(d_s); eng.dev_free(d_r)
    # e2e: gai_pan_rollout with PINNED host buffers (upload / play out / download in overlapped pipeline stages inside the call)
    pin_in = torch.from_numpy(deals.view(np.int64)).pin_memory().numpy().view(np.uint64)
    res = torch.zeros((n, 12), dtype=torch.int32).pin_memory().numpy().view(np.uint32)
    eng.pan_rollout_into(PAN_PLAYERS, pin_in, res, PAN_PPL, PAN_DEPTH, PAN_EVAL, SEED, first)       # warm-up (allocates the device buffers)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    e_ms = 0.0
    t0 = time.perf_counter()
    for _ in range(steps):
        st = eng.pan_rollout_into(PAN_PLAYERS, pin_in, res, PAN_PPL, PAN_DEPTH, PAN_EVAL, SEED, first)
        e_ms += st["total_ms"]
    wall = time.perf_counter() - t0
    assert (res == res_dev).all(), "Pan: host-buffer path and device-resident path disagree"
    tot = np.array([playouts, plies], np.uint64)
    if dist:
        tot = eng.allreduce_u64(tot)
        t = torch.tensor([k_ms, e_ms, wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        k_ms, e_ms, wall = t.tolist()
    if rank != 0:
        return None
    out = {"metric": "pan_determinized_playouts_per_sec", "unit": UNIT, "value": int(tot[0]) / (k_ms * 1e-3), "n_gpus": world, "steps": steps,
           "config": {"workload": "Pan determinized rollouts: %d sampled deals x %d playouts/deal per GPU, %d players, playout_depth %d, eval num_cards_weighted" % (n, PAN_PPL, PAN_PLAYERS, PAN_DEPTH),
                      "deals_per_gpu": n, "playouts_per_deal": PAN_PPL, "players": PAN_PLAYERS, "max_plies": PAN_DEPTH,
                      "deals": "gai_pan_determinize samples of player 0's view of a fresh 3-player game", "parallelism": "deal-parallel x%d, no data-path collective" % world},
           "kernel_ms_per_step": k_ms / steps, "mean_plies_per_playout": int(tot[1]) / max(1, int(tot[0])), "plies_per_sec": int(tot[1]) / (k_ms * 1e-3),
           "e2e": {"value": int(tot[0]) / wall, "unit": UNIT, "h2d_bytes_per_step": n * 40, "d2h_bytes_per_step": n * 48, "ms_per_step": 1e3 * wall / steps,
                   "device_ms_per_step": e_ms / steps, "api": "gai_pan_rollout (pinned HOST gai_pan_state[] in, pinned HOST gai_pan_result[] out; upload, kernels and download in overlapped stages of 65536 deals inside the call)"},
           "api": "gai_pan_rollout_device (deals resident in HBM, one pass)",
           "dtype": "u32"}
    p = os.path.join(ROOT, "profiles", "pan_kernel.json")
    if os.path.exists(p):
        prof = json.load(open(p))
        info = eng.device_info()
        peak = info["sm_count"] * 4 * info["clock_khz"] * 1e3
        ach = out["plies_per_sec"] / world * prof["warp_inst_per_thread_ply"]
        out["roofline"] = {"bound": "issue", "unit": "Gwarp-inst/s", "achieved": ach / 1e9, "peak": peak / 1e9, "frac": ach / peak,
                           "peak_how": "%d SMs x 4 sub-partitions x max SM clock %.0f MHz" % (info["sm_count"], info["clock_khz"] / 1e3),
                           "warp_inst_per_thread_ply": prof["warp_inst_per_thread_ply"], "traffic": prof.get("dram_bytes_per_launch"), "source": prof.get("source")}
    if with_cpu:
        try:
            from oracle import pan
            po = pan.best()
            m = 2048
            t0 = time.perf_counter(); orr, opl = po.playout_batch(PAN_PLAYERS, deals[:m], PAN_PPL, PAN_DEPTH, PAN_EVAL, SEED, first); dt = time.perf_counter() - t0
            assert (orr == res[:m]).all(), "Pan GPU results differ from the reference rules on the checked deals"
            out["cpu_baseline"] = {"value": m * PAN_PPL / dt, "unit": UNIT, "cores": 1, "kind": po.kind,
                                   "sample": "first %d deals x %d playouts, same Philox key and ids; GPU results on them are bit-identical" % (m, PAN_PPL)}
        except Exception as e:
            out["cpu_baseline"] = {"unavailable": str(e)[:200]}
    return out


# ------------------------------------------------------------------------------------------------
def profile_constants():
    """warp-instructions per ply of the rollout kernel, from the committed ncu capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "rollout_kernel.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return None


def _claim_stdout():
    """Everything any library prints (NCCL writes its version banner to fd 1) goes to stderr; the ONE JSON line is
    written to the real stdout through the returned file object."""
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    return real


def main():
    args = parse()
    real_stdout = _claim_stdout()
    if args.impl == "reference":
        return run_reference(args, real_stdout)

    import torch
    import gameai_b200 as g
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    eng = g.GpuRollout(local)            # raises without a GPU / library: there is no fallback path
    eng.set_rules_impl(args.rules_impl)
    if args.threads_per_block or args.blocks_per_sm:
        eng.set_launch_config(args.threads_per_block, args.blocks_per_sm)
    info = eng.device_info()
    n, ppl = args.leaves, args.ppl

    # synthetic corpus, generated on the device straight into the rollout layout (resident in HBM)
    d_b = eng.dev_alloc(n * 16); d_s = eng.dev_alloc(((n + 31) // 32) * 4); d_o = eng.dev_alloc(n * 16)
    import importlib
    plan = importlib.import_module("game-ai_b200.rootpar").shard_plan(rank, world, n, ppl, SEED)
    corpus_key = plan["corpus_key"]               # every rank plays its own shard
    host_states = eng.gen_reachable(n, corpus_key, 300, d_boards=d_b, d_stm=d_s, host=True)
    first_id = plan["first_playout_id"]

    # CPU legs: rank 0 at N = 1 only, and before anything collective exists (no GPU spins in a barrier meanwhile)
    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:
        cb = cpu_playouts(host_states, ppl, first_id, os.cpu_count() or 1, args.cpu_seconds)
        cpu_base = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "plies_per_sec")}
        try:
            cpu_base["reference_mctsplayer"] = reference_mctsplayer_baseline()
        except Exception as e:      # reported baseline only; never fatal
            cpu_base["reference_mctsplayer"] = {"unavailable": str(e)[:200]}

    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # library-owned NCCL communicator (the root-parallel merge path); id shipped through torch.distributed
        ids = [g.GpuRollout.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        eng.comm_init(rank, world, ids[0])

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    if args.workload == "pan":
        l0 = eng.launch_count()
        pan = pan_leg(args, eng, rank, world, dist, args.steps, args.warmup, with_cpu=(world == 1 and not args.no_cpu_baseline))
        pan_launches = eng.launch_count() - l0
        if rank == 0:
            pan.update({"warmup": args.warmup, "ms_per_step": pan["kernel_ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                        "data": "synthetic", "gpu_launches": pan_launches})
            print(json.dumps(pan), file=real_stdout, flush=True)
        eng.dev_free(d_b); eng.dev_free(d_s); eng.dev_free(d_o); eng.close()
        if dist:
            dist.barrier(); dist.destroy_process_group()
        return 0

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2
    for _ in range(args.warmup):
        eng.rollout_device(d_b, d_s, n, d_o, ppl, MAX_PLIES, SEED, first_id)
    launches0 = eng.launch_count()
    sampler = ClockSampler(physical_gpu_index(local))
    barrier()
    t_wall0 = time.perf_counter()
    dev_ms = 0.0; kern_ms = 0.0; plies = 0; playouts = 0
    for _ in range(args.steps):
        flush.zero_(); torch.cuda.synchronize()          # L2 flush between timed iterations (not inside the event bracket)
        st = eng.rollout_device(d_b, d_s, n, d_o, ppl, MAX_PLIES, SEED, first_id)
        dev_ms += st["total_ms"]; kern_ms += st["kernel_ms"]; plies += st["plies"]; playouts += st["playouts"]
    barrier()
    wall_s = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    launches = eng.launch_count() - launches0

    # outcome counters of the last step (device -> host), summed over ranks through the library's NCCL path
    out = np.zeros((n, 4), np.uint32); eng.d2h(out, d_o)
    tot = np.concatenate([out.sum(axis=0).astype(np.uint64), np.array([playouts, plies], np.uint64)])
    if world > 1:
        tot = eng.allreduce_u64(tot)
        t = torch.tensor([dev_ms, kern_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, kern_ms = t.tolist()
    all_playouts, all_plies = int(tot[4]), int(tot[5])
    value = all_playouts / (dev_ms * 1e-3)

    # ---- e2e: plugin-facing host-buffer call ------------------------------------------------------
    e2e = None
    if args.e2e_steps > 0:
        pin_in = torch.from_numpy(host_states).pin_memory().numpy()
        pin_out = torch.zeros((n, 4), dtype=torch.int32).pin_memory().numpy().view(np.uint32)
        eng.rollout_into(pin_in, pin_out, ppl, MAX_PLIES, SEED, first_id)         # warm-up (allocates staging)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            eng.rollout_into(pin_in, pin_out, ppl, MAX_PLIES, SEED, first_id)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if dist:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = t.item()
        assert (pin_out == out).all(), "host-buffer path and device-resident path disagree"
        e2e = {"value": world * n * ppl * args.e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": n * 24,
               "d2h_bytes_per_step": n * 16, "ms_per_step": 1e3 * dt / args.e2e_steps, "steps": args.e2e_steps,
               "api": "gai_loa_rollout (HOST gai_loa_state[] in, HOST gai_loa_result[] out; H2D of the 24-byte states + device split + kernel + D2H inside)"}
        del pin_in, pin_out
    del flush
    torch.cuda.empty_cache()

    # ---- config 5 (Pan) and config 4 (the player) ---------------------------------------------------
    pan = None
    if not args.no_pan:
        try:
            pan = pan_leg(args, eng, rank, world, dist, max(2, min(args.steps, 8)), 2, with_cpu=(world == 1 and not args.no_cpu_baseline))
        except Exception as e:
            pan = {"error": str(e)[:300]}
    player = None
    if not args.no_player:
        try:
            player = player_leg(args, rank, world, local, dist)
        except AssertionError:
            raise
        except Exception as e:
            player = {"error": str(e)[:300]}

    if rank == 0:
        mean_plies = all_plies / max(1, all_playouts)
        plies_per_s = all_plies / (kern_ms * 1e-3)
        f_sm = (clocks.get("sm_mhz") or info["clock_khz"] / 1e3) * 1e6
        peak_issue = info["sm_count"] * 4 * f_sm                      # warp-instructions/s: 1 per SMSP per cycle
        prof = profile_constants()
        roof = {"bound": "issue", "unit": "Gwarp-inst/s", "peak": peak_issue / 1e9,
                "peak_how": "%d SMs x 4 sub-partitions x %.0f MHz (median SM clock sampled during the timed region)" % (info["sm_count"], f_sm / 1e6),
                "achieved": None, "frac": None, "traffic": None}
        if prof:
            ach = plies_per_s / world * prof["warp_inst_per_thread_ply"]       # per GPU
            roof.update({"achieved": ach / 1e9, "frac": ach / peak_issue, "warp_inst_per_thread_ply": prof["warp_inst_per_thread_ply"],
                         "warp_inst_per_warp_ply": prof.get("warp_inst_per_warp_ply_if_full"),
                         "thread_inst_per_ply": prof.get("thread_inst_per_ply"),
                         # ncu dram bytes of the captured launch (262144 leaves), scaled to this launch's leaf count
                         "traffic": prof.get("dram_bytes_per_launch") * n / prof["leaves_per_launch"] if prof.get("dram_bytes_per_launch") and prof.get("leaves_per_launch") else prof.get("dram_bytes_per_launch"),
                         "traffic_how": "dram__bytes_read.sum + dram__bytes_write.sum of the ncu capture (%s leaves per launch), scaled by leaves" % prof.get("leaves_per_launch"),
                         "source": prof.get("source")})
        alg_bytes = n * (16 + 16) + ((n + 31) // 32) * 4                  # boards in, results out (per launch)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_ach = alg_bytes / (kern_ms / args.steps * 1e-3) / 1e9
        roof["hbm"] = {"bound": "hbm", "achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                       "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)",
                       "algorithmic_bytes_per_launch": alg_bytes}
        lc = eng.launch_config()
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u64", "data": "synthetic", "config": workload_config(args, world),
                "launch": {"threads_per_block": lc["threads_per_block"], "plies_per_loop_trip": lc["plies_per_trip"],
                           "kernel": ["k_rollout (bitboard baseline)", "k_rollout_fast<per-piece> (rotated boards + line LUT)", "k_rollout_fast<sweep> (static line sweep + CSA row counts)"][args.rules_impl],
                           "device": info["name"]},
                "target_frac": value / world / TARGET, "target": "%.0e playouts/s per B200 (BASELINE.json north_star)" % TARGET,
                "kernel_ms_per_step": kern_ms / args.steps, "wall_ms_per_step": 1e3 * wall_s / args.steps,
                "playouts_per_step": all_playouts // args.steps, "mean_plies_per_playout": mean_plies, "plies_per_sec": all_plies / (dev_ms * 1e-3),
                "outcomes_last_step": {"white": int(tot[0]), "black": int(tot[1]), "draw": int(tot[2]), "capped": int(tot[3])},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof, "player": player, "pan": pan}
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        print(json.dumps(line), file=real_stdout, flush=True)
    eng.dev_free(d_b); eng.dev_free(d_s); eng.dev_free(d_o)
    eng.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
