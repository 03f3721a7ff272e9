"""Generate tests/golden/pan_ref_corpus.npz from the REFERENCE ITSELF (oracle/_ref/libref_pan.so =
/root/reference/c++/GraWPanaZasadyV2/GraWPanaZasadyV2.cpp compiled verbatim by `make -C oracle ref`).
Run in the build container only:   python tests/golden/make_pan_golden.py"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pan  # noqa: E402

SEED = 0x10A5EED
N = 1500


def main():
    ref = pan.ref()
    assert ref is not None, "oracle/_ref/libref_pan.so not built (make -C oracle ref)"
    out = {"seed": np.uint64(SEED)}
    for npl in (2, 3, 4):
        S = ref.gen_reachable(npl, N, SEED, 100)
        c, m = ref.movegen_batch(npl, S)
        pick = (np.arange(N) * 2654435761 % np.maximum(c.astype(np.int64), 1)).astype(np.int64)
        mv = m[np.arange(N), pick]
        t, sc, e0, e1 = ref.terminal_batch(npl, S)
        k = "p%d_" % npl
        out.update({k + "states": S, k + "counts": c, k + "moves": m, k + "pick": mv, k + "applied": ref.apply_batch(npl, S, mv),
                    k + "terminal": t, k + "score": sc, k + "eval0": e0, k + "eval1": e1})
        for tag, cap, ek in (("cap50w", 50, 1), ("cap50n", 50, 0), ("full", 100000, 1)):
            code, pl, v, f = ref.playout_trace_batch(npl, S[:400], cap, ek, SEED, 500)
            out.update({k + tag + "_code": code, k + tag + "_plies": pl, k + tag + "_value": v, k + tag + "_final": f})
        # imperfect-information views (fuzzy cards): movegen through the reference for every seat
        views = np.stack([ref.player_known_state(npl, S[i], i % npl) for i in range(300)])
        vc, vm = ref.movegen_batch(npl, views)
        out.update({k + "views": views, k + "view_counts": vc, k + "view_moves": vm})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pan_ref_corpus.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
