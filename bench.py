#!/usr/bin/env python
"""bench.py -- LOA random playouts/s on N B200s (BASELINE.json metric), one JSON line on rank 0.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--leaves L] [--ppl P]

A "step" is one pass of the hot path over one synthetic leaf batch: L reachable LOA positions (SURVEY.md 8d
generator, made on the device), P uniformly random playouts per leaf to the end of the game
(max_plies = 50 000, run_config_loa.xml:6).  N > 1: launched by torchrun, one rank per GPU, every rank plays its
own L-leaf shard (root/leaf parallelism has no data-path collective; the four outcome counters are summed
once with NCCL through the library's own communicator).  value = all ranks' playouts / max-over-ranks time.

  value     device-resident: leaves already in HBM, timed with CUDA events on the launching stream
  e2e       the plugin-facing C-ABI call gai_loa_rollout() with HOST buffers (H2D + kernel + D2H inside)
  roofline  issue-slot roofline of the rollout kernel (the path is integer/branch bound);
            an "hbm" sub-object carries the byte roofline for completeness
  cpu_baseline / --impl reference: the reference's own rules (oracle/_ref, compiled from /root/reference in the
            build container) or the plain-C port, driven by the same Philox stream on the host cores
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--leaves", type=int, default=1 << 20)
    ap.add_argument("--ppl", type=int, default=16)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--threads-per-block", type=int, default=0)
    ap.add_argument("--blocks-per-sm", type=int, default=0)
    ap.add_argument("--rules-impl", type=int, default=2, help="0 = bitboard baseline kernel, 1 = rotated boards + line LUT per piece, 2 = static line sweep")
    return ap.parse_args()


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


# ------------------------------------------------------------------------------------------------
# CPU baseline: the reference's rules on the host cores, same Philox stream
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
            "plies_per_sec": plies / dt,
            "sample": "first %d leaves of the same corpus x %d playouts/leaf, same Philox key and ids, %d host threads" % (n_leaves, ppl, threads)}


def run_reference(args, out):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import loa
    port = loa.port()
    threads = os.cpu_count() or 1
    # the corpus is defined by the oracle's generator (identical to the device generator, see tests)
    n_gen = 4096
    states = port.gen_reachable(n_gen, SEED, 300)
    # warm-up + K bounded steps; each step is a sample of the workload sized to a few seconds
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
            "data": "synthetic",
            "config": {"workload": "LOA leaf-parallel rollouts: %d reachable leaves x %d playouts/leaf to game end (max_plies %d)" % (args.leaves, args.ppl, MAX_PLIES),
                       "leaves": args.leaves, "playouts_per_leaf": args.ppl, "max_plies": MAX_PLIES, "note": "each step is a bounded sample of this workload on the host cores"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["cores"], "kind": last["kind"], "sample": last["sample"],
                             "plies_per_sec": tot_pl / tot_t, "mean_plies_per_playout": tot_pl / max(1, tot_p)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), file=out, flush=True)
    return 0


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
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    eng = g.GpuRollout(local)            # raises without a GPU / library: there is no fallback path
    eng.set_rules_impl(args.rules_impl)
    if args.threads_per_block or args.blocks_per_sm:
        eng.set_launch_config(args.threads_per_block, args.blocks_per_sm)
    info = eng.device_info()
    n, ppl = args.leaves, args.ppl

    # library-owned NCCL communicator (the root-parallel merge path); id shipped through torch.distributed
    if world > 1:
        ids = [g.GpuRollout.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        eng.comm_init(rank, world, ids[0])

    # synthetic corpus, generated on the device straight into the rollout layout (resident in HBM)
    d_b = eng.dev_alloc(n * 16); d_s = eng.dev_alloc(((n + 31) // 32) * 4); d_o = eng.dev_alloc(n * 16)
    import importlib
    plan = importlib.import_module("game-ai_b200.rootpar").shard_plan(rank, world, n, ppl, SEED)
    corpus_key = plan["corpus_key"]               # every rank plays its own shard
    host_states = eng.gen_reachable(n, corpus_key, 300, d_boards=d_b, d_stm=d_s, host=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    first_id = plan["first_playout_id"]
    for _ in range(args.warmup):
        eng.rollout_device(d_b, d_s, n, d_o, ppl, MAX_PLIES, SEED, first_id)
    launches0 = eng.launch_count()
    sampler = ClockSampler(local if "CUDA_VISIBLE_DEVICES" not in os.environ else int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local]) if os.environ["CUDA_VISIBLE_DEVICES"].strip() else local)
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
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u64", "data": "synthetic",
                "config": {"workload": "LOA leaf-parallel rollouts: %d reachable leaves x %d playouts/leaf per GPU to game end (max_plies %d)" % (n, ppl, MAX_PLIES),
                           "leaves_per_gpu": n, "playouts_per_leaf": ppl, "max_plies": MAX_PLIES, "philox_key": SEED,
                           "l2": "flushed between timed iterations (256 MiB write)", "parallelism": "leaf/root-parallel x%d, no data-path collective" % world,
                           "threads_per_block": eng.launch_config()["threads_per_block"], "plies_per_loop_trip": eng.launch_config()["plies_per_trip"],
                           "kernel": ["k_rollout (bitboard baseline)", "k_rollout_fast<per-piece> (rotated boards + line LUT)", "k_rollout_fast<sweep> (static line sweep + CSA row counts)"][args.rules_impl], "device": info["name"]},
                "kernel_ms_per_step": kern_ms / args.steps, "wall_ms_per_step": 1e3 * wall_s / args.steps,
                "playouts_per_step": all_playouts // args.steps, "mean_plies_per_playout": mean_plies, "plies_per_sec": all_plies / (dev_ms * 1e-3),
                "outcomes_last_step": {"white": int(tot[0]), "black": int(tot[1]), "draw": int(tot[2]), "capped": int(tot[3])},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof}
        if not args.no_cpu_baseline:
            cb = cpu_playouts(host_states, ppl, first_id, os.cpu_count() or 1, args.cpu_seconds)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "plies_per_sec")}
            try:        # the reference's own MC::Player through selection_playOut (MCTSPlayer.cpp:387-470), 1 core, LOA opening
                from oracle import mcts
                rm = mcts.ref()
                if rm is not None:
                    r = rm.select(0x0081818181818100, 0x7e0000000000007e, 1, 500, seed=1234)
                    line["cpu_baseline"]["reference_mctsplayer"] = {
                        "simulations_per_sec": r["sims_per_sec"], "move_sim_limit": 500, "mean_path_plies": r["mean_path"], "cores": 1,
                        "what": "reference MC::Player (oracle/_ref/libref_mcts.so), run_config_loa.xml parameters, opening position"}
            except Exception as e:      # reported baseline only; never fatal
                line["cpu_baseline"]["reference_mctsplayer"] = {"unavailable": str(e)[:200]}
        print(json.dumps(line), file=real_stdout, flush=True)
    eng.dev_free(d_b); eng.dev_free(d_s); eng.dev_free(d_o)
    eng.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
