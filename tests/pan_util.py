import json
import os
import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
OPS = {"noop": 0, "take": 1, "play": 2}
PROB = {0: 0.0, 1: 1.0 / 3.0, 2: 0.5, 3: 1.0}


def mkmb(lst):
    """game_rules_v2.cpp:27-39 -- highest card first; 11 -> 0b11, 10 -> 0b10, 1 -> 0b01"""
    h = 0
    shift = (len(lst) - 1) * 2
    for v in lst:
        h |= (2 * (v // 10) + v % 10) << shift
        shift -= 2
    return h


def ones(pos):
    return (((1 << 64) - 1) << (2 * pos)) & ((1 << 48) - 1)


def state_words(stack, hands, counts, cp, terminal=0):
    w = [stack | (cp << 48) | (terminal << 51)]
    for p in range(4):
        cards = hands[p] if p < len(hands) else 0
        cnt = counts[p] if p < len(counts) else 0
        w.append((cnt & 15) | (cards << 4))
    return np.array(w, dtype=np.uint64)


def move_fields(w):
    w = int(w)
    return {"cards": w & ((1 << 48) - 1), "count": (w >> 48) & 7, "op": (w >> 51) & 7, "prob": (w >> 54) & 3}


def ut_vectors():
    with open(os.path.join(GOLDEN, "pan_ut_vectors.json")) as f:
        return json.load(f)


def ref_corpus():
    return np.load(os.path.join(GOLDEN, "pan_ref_corpus.npz"))


def case_state(c):
    hands = [mkmb(c["hand0"])] if "hand0" in c else [mkmb(h) for h in c["hands"]]
    counts = [c["count0"]] if "count0" in c else c["counts"]
    s = state_words(mkmb(c["stack"]), hands, counts, c["cp"])
    return s
