"""CPU: the static-sweep move generation of the rollout kernel (game-ai_b200/csrc/loa_sweep.cuh), compiled for the host
through its __host__ __device__ twin, against the oracle -- the same checks the GPU tier-1/tier-2 tests make, before any
GPU time is spent (LinesOfActionZasady.cpp:167-232 move order, :233-252 apply, :52-58 terminal)."""
import ctypes
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def corpus(oracle, n, seed):
    return np.ascontiguousarray(np.concatenate([oracle.gen_reachable(n, seed), oracle.gen_uniform(n, seed)]))


def test_sweep_movegen_on_host_vs_oracle(hostlib, oracle):
    S = corpus(oracle, 20000, 0x51EE9)
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
