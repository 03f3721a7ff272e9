import json
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def sq(name):
    return (ord(name[0]) - ord("A")) + 8 * (ord(name[1]) - ord("1"))


def board(names):
    b = 0
    for n in names:
        b |= 1 << sq(n)
    return b


def ut_vectors():
    with open(os.path.join(GOLDEN, "loa_ut_vectors.json")) as f:
        return json.load(f)


def ref_corpus():
    return np.load(os.path.join(GOLDEN, "loa_ref_corpus.npz"))


def case_state(c):
    if c.get("opening"):
        return 0x0081818181818100, 0x7e0000000000007e, 1
    w = int(c["white_hex"], 16) if "white_hex" in c else board(c.get("white", []))
    b = int(c["black_hex"], 16) if "black_hex" in c else board(c.get("black", []))
    return w, b, c.get("cp", 0)


def parse_move(s):
    return sq(s[0:2]), sq(s[4:6])
