"""ctypes bindings for the Pan oracle -- TEST INFRASTRUCTURE ONLY.
`port()` = oracle/libpan_oracle.so (plain-C restatement), `ref()` = oracle/_ref/libref_pan.so (the reference's
GraWPanaZasadyV2.cpp compiled verbatim) or None.  States are uint64 arrays of shape (n, 5) -- the reference's raw
40-byte GameState -- moves raw uint64."""
import ctypes
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_u64, _u32, _i32, _vp = ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int, ctypes.c_void_p
MAX_MOVES = 16


class PanOracle:
    def __init__(self, path, prefix, kind):
        self.lib = ctypes.CDLL(path); self.kind = kind
        g = lambda n: getattr(self.lib, prefix + n)
        self._deal = g("deal"); self._deal.argtypes = [_i32, _u64, _u64, _vp]
        self._pks = g("player_known_state"); self._pks.argtypes = [_i32, _vp, _i32, _vp]
        self._make11 = g("make11"); self._make11.argtypes = [_u64]; self._make11.restype = _u64
        self._msm = g("make_stack_masks"); self._msm.argtypes = [_u64, _vp]
        self._mg = g("movegen_batch"); self._mg.argtypes = [_i32, _vp, _u32, _vp, _vp]
        self._ap = g("apply_batch"); self._ap.argtypes = [_i32, _vp, _vp, _u32, _vp]
        self._tb = g("terminal_batch"); self._tb.argtypes = [_i32, _vp, _u32, _vp, _vp, _vp, _vp]
        self._pt = g("playout_trace_batch"); self._pt.argtypes = [_i32, _vp, _u32, _u32, _i32, _u64, _u64, _vp, _vp, _vp, _vp]
        self._pb = g("playout_batch"); self._pb.argtypes = [_i32, _vp, _u32, _u32, _u32, _i32, _u64, _u64, _vp]; self._pb.restype = _u64
        self._gen = g("gen_reachable"); self._gen.argtypes = [_i32, _u32, _u64, _u32, _vp]

    @staticmethod
    def _st(states):
        s = np.ascontiguousarray(states, dtype=np.uint64)
        assert s.ndim == 2 and s.shape[1] == 5
        return s

    def deal(self, np_, key, id_):
        s = np.zeros(5, np.uint64); self._deal(np_, key, id_, s.ctypes.data); return s

    def player_known_state(self, np_, state, player):
        s = np.ascontiguousarray(state, dtype=np.uint64); o = np.zeros(5, np.uint64)
        self._pks(np_, s.ctypes.data, player, o.ctypes.data); return o

    def make11(self, cards):
        return int(self._make11(cards))

    def make_stack_masks(self, stack):
        o = (ctypes.c_uint64 * 3)(); self._msm(stack, o); return tuple(int(x) for x in o)

    def movegen_batch(self, np_, states):
        s = self._st(states); n = s.shape[0]
        c = np.zeros(n, np.uint8); m = np.zeros((n, MAX_MOVES), np.uint64)
        self._mg(np_, s.ctypes.data, n, c.ctypes.data, m.ctypes.data); return c, m

    def apply_batch(self, np_, states, moves):
        s = self._st(states); n = s.shape[0]
        mv = np.ascontiguousarray(moves, dtype=np.uint64); o = np.zeros((n, 5), np.uint64)
        self._ap(np_, s.ctypes.data, mv.ctypes.data, n, o.ctypes.data); return o

    def terminal_batch(self, np_, states):
        """-> (terminal[n], score[n,4], eval_num_cards[n,4], eval_num_cards_weighted[n,4])"""
        s = self._st(states); n = s.shape[0]
        t = np.zeros(n, np.uint8); sc = np.zeros((n, 4), np.int32); e0 = np.zeros((n, 4), np.int32); e1 = np.zeros((n, 4), np.int32)
        self._tb(np_, s.ctypes.data, n, t.ctypes.data, sc.ctypes.data, e0.ctypes.data, e1.ctypes.data); return t, sc, e0, e1

    def playout_trace_batch(self, np_, states, max_plies, eval_kind, key, first_id):
        s = self._st(states); n = s.shape[0]
        c = np.zeros(n, np.uint8); pl = np.zeros(n, np.uint32); v = np.zeros((n, 4), np.int32); f = np.zeros((n, 5), np.uint64)
        self._pt(np_, s.ctypes.data, n, max_plies, eval_kind, key, first_id, c.ctypes.data, pl.ctypes.data, v.ctypes.data, f.ctypes.data)
        return c, pl, v, f

    def playout_batch(self, np_, states, ppl, max_plies, eval_kind, key, first_id):
        s = self._st(states); n = s.shape[0]
        r = np.zeros((n, 12), np.uint32)
        tot = self._pb(np_, s.ctypes.data, n, ppl, max_plies, eval_kind, key, first_id, r.ctypes.data)
        return r, int(tot)

    def gen_reachable(self, np_, n, key, max_k=60):
        s = np.zeros((n, 5), np.uint64); self._gen(np_, n, key, max_k, s.ctypes.data); return s


_cache = {}


def port():
    if "p" not in _cache:
        _cache["p"] = PanOracle(os.path.join(_HERE, "libpan_oracle.so"), "pan_oracle_", "port")
    return _cache["p"]


def ref():
    if "r" not in _cache:
        p = os.path.join(_HERE, "_ref", "libref_pan.so")
        _cache["r"] = PanOracle(p, "ref_pan_", "reference") if os.path.exists(p) else None
    return _cache["r"]


def best():
    return ref() or port()
