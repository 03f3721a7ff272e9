"""CPU: the static-sweep move generation of the rollout kernel (game-ai_b200/csrc/loa_sweep.cuh), compiled for the host
through its __host__ __device__ twin, against the oracle -- the same checks the GPU tier-1/tier-2 tests make, before any
GPU time is spent (LinesOfActionZasady.cpp:167-232 move order, :233-252 apply, :52-58 terminal)."""
import ctypes
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def corpus(oracle, n, seed):
    return np.ascontiguousarray(np.concatenate([oracle.gen_reachable(n, seed), oracle.gen_uniform(n, seed)]))


def crowded_rows():
    """Positions where one row holds four own pieces on squares of one colour, all free to move the same way along three line
    types: the only ones that reach the 'eights' plane of the carry-save row counter (loa_sweep.cuh), which every other
    position skips."""
    out = []
    for r in range(8):
        for pat in (0x55, 0xAA, 0xFF, 0x77):
            own = pat << (8 * r)
            far = 1 << (8 * ((r + 4) % 8) + 3)
            out += [[own, far, 0], [far, own, 1], [own | far << 1 if (far << 1) & ~own else own, far, 0]]
    return np.array(out, dtype=np.uint64)


def test_sweep_movegen_on_host_vs_oracle(hostlib, oracle):
    S = np.ascontiguousarray(np.concatenate([corpus(oracle, 20000, 0x51EE9), crowded_rows()]))
    n = len(S)
    counts = np.zeros(n, np.uint8); moves = np.zeros((n, 96, 2), np.uint8)
    hostlib.host_loa_sweep_movegen_batch(ctypes.c_void_p(S.ctypes.data), ctypes.c_uint(n), ctypes.c_void_p(counts.ctypes.data), ctypes.c_void_p(moves.ctypes.data))
    oc, om = oracle.movegen_batch(S)
    assert (counts == oc).all()
    assert (moves == om).all()


def test_sweep_playouts_on_host_vs_oracle(hostlib, oracle):
    S = corpus(oracle, 6000, 0xA11CE)
    n = len(S)
    for cap in (50000, 37):
        o = np.zeros(n, np.uint8); pl = np.zeros(n, np.uint32); f = np.zeros((n, 3), np.uint64)
        hostlib.host_loa_sweep_playout_batch(ctypes.c_void_p(S.ctypes.data), ctypes.c_uint(n), ctypes.c_uint(cap), ctypes.c_uint64(0xBEEF), ctypes.c_uint64(1000),
                                             ctypes.c_void_p(o.ctypes.data), ctypes.c_void_p(pl.ctypes.data), ctypes.c_void_p(f.ctypes.data))
        oo, opl, of = oracle.playout_trace_batch(S, cap, 0xBEEF, 1000)
        assert (o == oo).all() and (pl == opl).all()
        assert (f[:, :2] == of[:, :2]).all() and ((f[:, 2] & 1) == (of[:, 2] & 1)).all()


def test_small_tables_by_definition(hostlib):
    """The two per-lane tables of the sweep kernel (loa_fast.cuh slt_entry / tpc_entry) against their definitions: the
    short-line table is the line LUT restricted to four cells (cells 4-7 walls), the field-count table the popcount of
    each 2-bit field in its own byte lane."""
    hostlib.host_loa_lut_entry.restype = ctypes.c_int
    hostlib.host_loa_slt_entry.restype = ctypes.c_uint
    hostlib.host_loa_tpc_entry.restype = ctypes.c_uint
    for idx in range(256):
        o4, e4 = idx & 15, idx >> 4
        assert hostlib.host_loa_slt_entry(idx) == hostlib.host_loa_lut_entry(o4 | 0xf0, e4 | 0xf0) & 0xff
        t = hostlib.host_loa_tpc_entry(idx)
        assert [(t >> (8 * j)) & 0xff for j in range(4)] == [bin((idx >> (2 * j)) & 3).count("1") for j in range(4)]
    # a 4-cell line with no phantoms equals the 8-cell line whose upper four cells are walls: same moves as an edge row
    for o4 in range(16):
        for e4 in range(16):
            if o4 & e4:
                continue
            full = hostlib.host_loa_lut_entry(o4, e4)       # cells 4-7 EMPTY: pieces may run past cell 3
            short = hostlib.host_loa_slt_entry(o4 | (e4 << 4))
            n = bin(o4 | e4).count("1")
            for p in range(4):
                up_full = (full >> (2 * p + 1)) & 1
                up_short = (short >> (2 * p + 1)) & 1
                assert up_short == (up_full if p + n < 4 else 0)
                assert ((short >> (2 * p)) & 1) == ((full >> (2 * p)) & 1)


def _components_and_holes(b):
    """8-connected components of the set squares and holes = 4-connected regions of empty squares that do not reach the
    border (brute force, by definition)."""
    cells = {(i // 8, i % 8) for i in range(64) if (b >> i) & 1}
    seen, comps = set(), 0
    for c in cells:
        if c in seen:
            continue
        comps += 1
        st = [c]; seen.add(c)
        while st:
            r, k = st.pop()
            for dr in (-1, 0, 1):
                for dc in (-1, 0, 1):
                    n = (r + dr, k + dc)
                    if n in cells and n not in seen:
                        seen.add(n); st.append(n)
    # background on a 10 x 10 frame: everything 4-connected to the frame is "outside"
    empty = {(r, k) for r in range(-1, 9) for k in range(-1, 9)} - cells
    seen, regions = set(), 0
    for c in empty:
        if c in seen:
            continue
        regions += 1
        st = [c]; seen.add(c)
        while st:
            r, k = st.pop()
            for n in ((r + 1, k), (r - 1, k), (r, k + 1), (r, k - 1)):
                if n in empty and n not in seen:
                    seen.add(n); st.append(n)
    return comps, regions - 1


def test_euler_characteristic_is_components_minus_holes(hostlib, oracle):
    """loa_euler.cuh: chi = components - holes for arbitrary boards (so chi != 1 proves "not one group"), and the
    component count agrees with the reference's calcNumGroups (LinesOfActionZasady.cpp:271-305)."""
    hostlib.host_loa_chi.argtypes = [ctypes.c_uint64]
    hostlib.host_loa_chi.restype = ctypes.c_int
    rng = np.random.RandomState(5)
    boards = [0, 1, 0xFFFFFFFFFFFFFFFF, 0x0000001c141c0000, 0x00003c24243c0000, 0x0000001008140800 | 0x8000000000000001, 0x7e0000000000007e, 0x0081818181818100]
    for dens in (0.1, 0.2, 0.35, 0.5, 0.7):
        for _ in range(200):
            boards.append(int(sum(1 << i for i in range(64) if rng.rand() < dens)))
    S = oracle.gen_uniform(500, 9)
    boards += [int(x) for x in S[:, 0]] + [int(x) for x in S[:, 1]]
    holes_seen = 0
    for b in boards:
        c, h = _components_and_holes(b)
        assert hostlib.host_loa_chi(b) == c - h, hex(b)
        assert oracle.num_groups(b) == c
        holes_seen += h > 0
    assert holes_seen > 50
