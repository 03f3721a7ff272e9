#!/usr/bin/env python
"""Per-STAGE dynamic budget of the sweep rollout kernel: joins the per-instruction counters of an ncu capture (source page:
instructions executed, shared-memory wavefronts) with the source lines of the same build (nvdisasm --print-line-info-inline of
the cubin) and rolls them up by stage of wstep_sweep (loa_fast_kernels.cu) / loa_sweep.cuh.
usage: python profiles/sass_stages.py rep.ncu-rep annotated.sass total_thread_plies
       annotated.sass: cuobjdump -xelf all build/loa_fast_kernels.o; nvdisasm --print-line-info-inline *.cubin, cut to the kernel"""
import csv, re, subprocess, sys, collections

rep, sass, plies = sys.argv[1], sys.argv[2], float(sys.argv[3])
out = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv"], stderr=subprocess.DEVNULL).decode()
rows = list(csv.reader(out.splitlines()))
h = rows[1]
ie, wf, src = h.index('Instructions Executed'), h.index('L1 Wavefronts Shared'), h.index('Source')
dyn = [(float(r[ie]), float(r[wf] or 0), r[src].strip()) for r in rows[2:] if len(r) > wf]

cur, pending, chains = [], [], []
for l in open(sass):
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        pending.append((m.group(1).split('/')[-1], int(m.group(2)))); continue
    if re.match(r'\s*/\*[0-9a-f]{4,5}\*/\s+\S', l):
        if pending: cur, pending = pending, []
        chains.append(cur)
assert len(chains) == len(dyn), (len(chains), len(dyn))

import os
CS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "game-ai_b200", "csrc")
def line_of(fn, pat, after=0):
    for i, l in enumerate(open(os.path.join(CS, fn)), 1):
        if i > after and pat in l: return i
    raise SystemExit("marker not found: " + pat)
SW_ROWS = line_of("loa_sweep.cuh", "GAI_HDI RowCounts sweep_rows")
SW_CSA = line_of("loa_sweep.cuh", "positional popcount of the 12 registers", SW_ROWS)
SW_SEL = line_of("loa_sweep.cuh", "// step 3: the idx-th move", SW_CSA)
SW_FA = line_of("loa_sweep.cuh", "GAI_HDI u32 rotl16")
K_STEP = line_of("loa_fast_kernels.cu", "__device__ __forceinline__ int wstep_sweep")
K_FILL1 = line_of("loa_fast_kernels.cu", "const bool c_en = wfill", K_STEP); K_FILL2 = line_of("loa_fast_kernels.cu", "c_my = wfill", K_STEP)
K_ROWS = line_of("loa_fast_kernels.cu", "const RowCounts rc = sweep_rows", K_STEP)
K_PHX = line_of("loa_fast_kernels.cu", "rnd = gai::choice_block", K_STEP); K_SEL = line_of("loa_fast_kernels.cu", "sweep_select(rc, idx", K_STEP)
K_TGT = line_of("loa_fast_kernels.cu", "move_target_tab(", K_STEP); K_END = line_of("loa_fast_kernels.cu", "return result;", K_TGT)
K_WFILL = line_of("loa_fast_kernels.cu", "__device__ __forceinline__ bool wfill")
K_REFILL = line_of("loa_fast_kernels.cu", "const unsigned needmask = (it & 3u)"); K_ALL = line_of("loa_fast_kernels.cu", "if (__all_sync(FULL, exhausted)) break;", K_REFILL)
K_LOOPEND = line_of("loa_fast_kernels.cu", "u64 wp = my_plies", K_ALL)

def stage(ch):
    """stage by the frame of wstep_sweep the instruction was inlined into, refined by the innermost frame of loa_sweep.cuh"""
    kern = [ln for f, ln in ch if f == 'loa_fast_kernels.cu']
    sw = [ln for f, ln in ch if f == 'loa_sweep.cuh']
    files = [f for f, _ in ch]
    phase = next((ln for ln in kern if K_STEP <= ln <= K_END), None)
    if phase in (K_FILL1, K_FILL2, K_FILL2 - 1) or (phase is None and any(K_WFILL <= ln <= K_WFILL + 14 for ln in kern)): return '1 connectivity: chi gate + flood fill'
    if phase == K_ROWS:
        if any(SW_CSA <= ln < SW_SEL for ln in sw): return '3 row totals: CSA planes + field-count table'
        if any(SW_ROWS <= ln < SW_CSA for ln in sw) and not any(ln < SW_FA for ln in sw): return '2b sweep: R popcounts, pack16'
        return '2a sweep: 42 line lookups (address, LDS, diagonal pre-masks)'
    if phase in (K_PHX, K_PHX + 1) or 'philox.h' in files: return '4 Philox block + choice index'
    if phase == K_SEL: return '5 select: row, column, cell, direction'
    if phase is not None and K_TGT <= phase <= K_TGT + 1: return '6 move target + apply (4 rotations, chi update)'
    if phase is not None: return '7 step bookkeeping (terminal / capped / outcome)'
    if any(K_REFILL <= ln < K_ALL for ln in kern): return '8 refill (work counter, leaf loads)'
    if any(K_ALL <= ln < K_LOOPEND for ln in kern): return '9 loop, per-leaf result reduction'
    return 'x prologue / epilogue (LUT stage-in, final reduce)'

ins, wfs = collections.Counter(), collections.Counter()
for (e, w, s), ch in zip(dyn, chains):
    st = stage(ch); ins[st] += e; wfs[st] += w
wp = plies / 32.0
ti, tw = sum(ins.values()), sum(wfs.values())
print("thread-plies %.4e -> warp-plies (32 lanes) %.4e; warp-inst %.4e = %.1f per warp-ply; shared wavefronts %.4e = %.1f per warp-ply" % (plies, wp, ti, ti / wp, tw, tw / wp))
print("%-62s %10s %7s %12s %7s" % ("stage", "inst/w-ply", "share", "wavefr/w-ply", "share"))
for k in sorted(ins):
    print("%-62s %10.1f %6.1f%% %12.1f %6.1f%%" % (k, ins[k] / wp, 100 * ins[k] / ti, wfs[k] / wp, 100 * wfs[k] / max(tw, 1)))
