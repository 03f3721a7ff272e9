"""ctypes binding of the reference's own MC::Player compiled behind a C API (oracle/_ref/libref_mcts.so,
ref_mcts_capi.cpp) -- TEST INFRASTRUCTURE ONLY.  `ref()` is None when oracle/_ref has not been built."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_cache = {}


class RefMcts:
    def __init__(self, path):
        self.lib = ctypes.CDLL(path)
        self.lib.ref_mcts_select.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_int, ctypes.c_uint, ctypes.c_double,
                                             ctypes.c_int, ctypes.c_int, ctypes.c_double] + [ctypes.c_void_p] * 5
        self.lib.ref_mcts_select_timed.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_double, ctypes.c_uint, ctypes.c_double,
                                                   ctypes.c_int, ctypes.c_int, ctypes.c_double] + [ctypes.c_void_p] * 6

    def select(self, w, b, cp, sims, seed=1234, C=2.0, playout_depth=50000, expand_size=3, eps=0.05):
        """One move decision of the reference player (run_config_loa.xml parameters by default).
        -> dict(move=(from,to), visits[], values[][2], seconds, mean_path)"""
        mv = (ctypes.c_uint8 * 2)(); vis = (ctypes.c_uint32 * 128)(); val = (ctypes.c_float * 256)()
        sec = ctypes.c_double(); mp = ctypes.c_double()
        n = self.lib.ref_mcts_select(w, b, cp, sims, seed, C, playout_depth, expand_size, eps, mv, vis, val, ctypes.byref(sec), ctypes.byref(mp))
        return {"move": (mv[0], mv[1]), "visits": list(vis[:n]), "values": [(val[2 * i], val[2 * i + 1]) for i in range(n)],
                "seconds": sec.value, "mean_path": mp.value, "sims_per_sec": sims / sec.value if sec.value > 0 else 0.0}


    def select_timed(self, w, b, cp, time_limit, seed=1234, C=2.0, playout_depth=50000, expand_size=3, eps=0.05):
        """The same decision under move_time_limit (seconds); `sims` = simulations the reference player managed."""
        mv = (ctypes.c_uint8 * 2)(); vis = (ctypes.c_uint32 * 128)(); val = (ctypes.c_float * 256)()
        sec = ctypes.c_double(); mp = ctypes.c_double(); runs = ctypes.c_long()
        n = self.lib.ref_mcts_select_timed(w, b, cp, time_limit, seed, C, playout_depth, expand_size, eps, mv, vis, val,
                                           ctypes.byref(sec), ctypes.byref(mp), ctypes.byref(runs))
        return {"move": (mv[0], mv[1]), "visits": list(vis[:n]), "values": [(val[2 * i], val[2 * i + 1]) for i in range(n)],
                "seconds": sec.value, "mean_path": mp.value, "sims": runs.value, "sims_per_sec": runs.value / sec.value if sec.value > 0 else 0.0}


def ref(fastpool=False):
    """fastpool=True: the build with the stand-in pool allocator (same search, no super-linear slow-down; not for timing)."""
    k = "f" if fastpool else "r"
    if k not in _cache:
        p = os.path.join(_HERE, "_ref", "libref_mcts_fastpool.so" if fastpool else "libref_mcts.so")
        _cache[k] = RefMcts(p) if os.path.exists(p) else None
    return _cache[k]
