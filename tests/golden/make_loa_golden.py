"""Generate tests/golden/loa_ref_corpus.npz from the REFERENCE ITSELF (oracle/_ref/libref_loa.so =
/root/reference/c++/LinesOfActionZasady/LinesOfActionZasady.cpp compiled verbatim by `make -C oracle ref`).

Run in the build container only (needs /root/reference):   python tests/golden/make_loa_golden.py
The committed .npz is what pins the oracle restatement (and through it the CUDA path) on machines where
the reference is absent (the GPU box).  Inputs come from the seeded synthetic corpus of SURVEY.md 8d.
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import loa  # noqa: E402

SEED = 0x10A5EED
N_REACH, N_UNIF, N_PLAY = 3000, 3000, 600


def main():
    loa.build(ref=True)
    ref, port = loa.ref(), loa.port()
    assert ref is not None, "oracle/_ref not built"
    # inputs: seeded corpus (the generator is ours -- any inputs would do -- the OUTPUTS are the reference's)
    states = np.concatenate([port.gen_reachable(N_REACH, SEED), port.gen_uniform(N_UNIF, SEED)])
    counts, moves = ref.movegen_batch(states)
    term, score = ref.terminal_batch(states)
    succ = ref.successor_hash_batch(states)
    # apply: the first and the last legal move of each position (positions without moves get the noop)
    first = moves[:, 0, :].copy()
    last = moves[np.arange(len(states)), np.maximum(counts.astype(int) - 1, 0), :].copy()
    first[counts == 0] = 0
    last[counts == 0] = 0
    ap_first = ref.apply_batch(states, first)
    ap_last = ref.apply_batch(states, last)
    pstates = np.concatenate([states[:N_PLAY // 2], states[N_REACH:N_REACH + N_PLAY // 2]])
    o, pl, fin = ref.playout_trace_batch(pstates, 50000, SEED, 1000)
    o_cap, pl_cap, fin_cap = ref.playout_trace_batch(pstates, 40, SEED, 1000)
    masks = np.array([[getattr(ref, f)(p) for p in range(64)] for f in
                      ("neighbour_mask", "row_mask", "column_mask", "left_diagonal_mask", "right_diagonal_mask")], dtype=np.uint64)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "loa_ref_corpus.npz")
    np.savez_compressed(out, seed=np.uint64(SEED), states=states, counts=counts, moves=moves, terminal=term, score=score,
                        successor_hash=succ, mv_first=first, mv_last=last, apply_first=ap_first, apply_last=ap_last,
                        play_states=pstates, play_outcome=o, play_plies=pl, play_final=fin,
                        cap_outcome=o_cap, cap_plies=pl_cap, cap_final=fin_cap, masks=masks)
    print("wrote", out, os.path.getsize(out), "bytes;", len(states), "positions,", len(pstates), "playouts; mean plies", pl.mean())


if __name__ == "__main__":
    main()
