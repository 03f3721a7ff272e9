"""ctypes bindings for the LOA oracle -- TEST INFRASTRUCTURE ONLY.

`port()` loads oracle/libloa_oracle.so (plain-C restatement, loa_oracle.c); `ref()` loads
oracle/_ref/libref_loa.so (the reference's LinesOfActionZasady.cpp compiled verbatim, see
ref_loa_capi.cpp) or returns None when it has not been built.  Both expose the same methods.
"""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_u64, _u32, _i32 = ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int
_vp = ctypes.c_void_p


def build(ref=True):
    """Compile the restatement and (when /root/reference exists) the reference itself."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "port"])
    if ref and os.path.isdir("/root/reference/c++"):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


class LoaOracle:
    def __init__(self, path, prefix, kind):
        self.lib = ctypes.CDLL(path)
        self.kind = kind
        self.path = path
        g = lambda name: getattr(self.lib, prefix + name)
        self._initial = g("initial")
        self._movegen = g("movegen"); self._movegen.argtypes = [_u64, _u64, _i32, _vp]; self._movegen.restype = _i32
        self._apply = g("apply"); self._apply.argtypes = [_u64, _u64, _i32, _i32, _i32, _vp, _vp, _vp]
        self._is_terminal = g("is_terminal"); self._is_terminal.argtypes = [_u64, _u64]; self._is_terminal.restype = _i32
        self._score = g("score"); self._score.argtypes = [_u64, _u64, _vp]
        for nm in ("neighbour_mask", "row_mask", "column_mask", "left_diagonal_mask", "right_diagonal_mask"):
            f = g(nm); f.argtypes = [_i32]; f.restype = _u64
            setattr(self, nm, f)
        self._cit = g("calc_if_terminal"); self._cit.argtypes = [_u64]; self._cit.restype = _i32
        self._ng = g("num_groups"); self._ng.argtypes = [_u64]; self._ng.restype = _i32
        self._tostr = g("state_to_string"); self._tostr.argtypes = [_u64, _u64, _i32, _vp, _i32]; self._tostr.restype = _i32
        self._playout = g("playout")
        self._playout.argtypes = [_u64, _u64, _i32, _u64, _u64, _u32, _vp, _vp, _vp]; self._playout.restype = _i32
        self._batch = g("playout_batch")
        self._batch.argtypes = [_vp, _vp, _vp, _u32, _u32, _u32, _u64, _u64, _vp]; self._batch.restype = _u64

        self._mgb = g("movegen_batch"); self._mgb.argtypes = [_vp, _u32, _vp, _vp]
        self._tb = g("terminal_batch"); self._tb.argtypes = [_vp, _u32, _vp, _vp]
        self._ab = g("apply_batch"); self._ab.argtypes = [_vp, _vp, _u32, _vp]
        self._shb = g("successor_hash_batch"); self._shb.argtypes = [_vp, _u32, _vp]
        self._ptb = g("playout_trace_batch"); self._ptb.argtypes = [_vp, _u32, _u32, _u64, _u64, _vp, _vp, _vp]
        if kind == "port":
            self._genr = g("gen_reachable"); self._genr.argtypes = [_u32, _u64, _u32, _vp]
            self._genu = g("gen_uniform"); self._genu.argtypes = [_u32, _u64, _vp]
            self._philox = g("philox"); self._philox.argtypes = [_vp, _vp, _vp]

    # ---- batch forms: states = uint64 array of shape (n, 3) = white, black, current_player ----
    @staticmethod
    def _st(states):
        s = np.ascontiguousarray(states, dtype=np.uint64)
        assert s.ndim == 2 and s.shape[1] == 3
        return s

    def movegen_batch(self, states):
        """-> (counts[n] uint8, moves[n,96,2] uint8)"""
        s = self._st(states); n = s.shape[0]
        counts = np.zeros(n, np.uint8); moves = np.zeros((n, 96, 2), np.uint8)
        self._mgb(s.ctypes.data, n, counts.ctypes.data, moves.ctypes.data)
        return counts, moves

    def terminal_batch(self, states):
        s = self._st(states); n = s.shape[0]
        t = np.zeros(n, np.uint8); sc = np.zeros((n, 2), np.uint8)
        self._tb(s.ctypes.data, n, t.ctypes.data, sc.ctypes.data)
        return t, sc

    def apply_batch(self, states, mv):
        s = self._st(states); n = s.shape[0]
        mv = np.ascontiguousarray(mv, dtype=np.uint8); assert mv.shape == (n, 2)
        out = np.zeros((n, 3), np.uint64)
        self._ab(s.ctypes.data, mv.ctypes.data, n, out.ctypes.data)
        return out

    def successor_hash_batch(self, states):
        s = self._st(states); n = s.shape[0]
        h = np.zeros(n, np.uint64)
        self._shb(s.ctypes.data, n, h.ctypes.data)
        return h

    def playout_trace_batch(self, states, max_plies, key, first_id):
        """-> (outcome[n] uint8, plies[n] uint32, final[n,3] uint64)"""
        s = self._st(states); n = s.shape[0]
        o = np.zeros(n, np.uint8); pl = np.zeros(n, np.uint32); f = np.zeros((n, 3), np.uint64)
        self._ptb(s.ctypes.data, n, max_plies, key, first_id, o.ctypes.data, pl.ctypes.data, f.ctypes.data)
        return o, pl, f

    def gen_reachable(self, n, key, max_k=300):
        s = np.zeros((n, 3), np.uint64)
        self._genr(n, key, max_k, s.ctypes.data)
        return s

    def gen_uniform(self, n, key):
        s = np.zeros((n, 3), np.uint64)
        self._genu(n, key, s.ctypes.data)
        return s

    def philox(self, ctr, key):
        c = (ctypes.c_uint32 * 4)(*ctr); k = (ctypes.c_uint32 * 2)(*key); o = (ctypes.c_uint32 * 4)()
        self._philox(c, k, o)
        return tuple(o)

    def rollout_states(self, states, playouts_per_leaf, max_plies, key, first_id):
        s = self._st(states)
        return self.playout_batch(s[:, 0].copy(), s[:, 1].copy(), (s[:, 2] & 1).astype(np.uint8), playouts_per_leaf, max_plies, key, first_id)

    def initial(self):
        w, b, cp = _u64(), _u64(), _i32()
        self._initial(ctypes.byref(w), ctypes.byref(b), ctypes.byref(cp))
        return w.value, b.value, cp.value

    def movegen(self, w, b, cp):
        buf = (ctypes.c_uint8 * 192)()
        n = self._movegen(w, b, cp, buf)
        return [(buf[2 * i], buf[2 * i + 1]) for i in range(n)]

    def apply(self, w, b, cp, frm, to):
        ow, ob, ocp = _u64(), _u64(), _i32()
        self._apply(w, b, cp, frm, to, ctypes.byref(ow), ctypes.byref(ob), ctypes.byref(ocp))
        return ow.value, ob.value, ocp.value

    def is_terminal(self, w, b):
        return bool(self._is_terminal(w, b))

    def score(self, w, b):
        s = (_i32 * 2)()
        self._score(w, b, s)
        return s[0], s[1]

    def calc_if_terminal(self, board):
        return bool(self._cit(board))

    def num_groups(self, board):
        return self._ng(board)

    def to_string(self, w, b, cp):
        buf = ctypes.create_string_buffer(128)
        self._tostr(w, b, cp, buf, 128)
        return buf.value.decode()

    def playout(self, w, b, cp, key, playout_id, max_plies):
        """-> (outcome, plies, final_white, final_black)"""
        plies, fw, fb = _u32(), _u64(), _u64()
        o = self._playout(w, b, cp, key, playout_id, max_plies, ctypes.byref(plies), ctypes.byref(fw), ctypes.byref(fb))
        return o, plies.value, fw.value, fb.value

    def playout_batch(self, whites, blacks, cps, playouts_per_leaf, max_plies, key, first_id):
        """-> (counts[n,4] uint32, total_plies)"""
        whites = np.ascontiguousarray(whites, dtype=np.uint64)
        blacks = np.ascontiguousarray(blacks, dtype=np.uint64)
        cps = np.ascontiguousarray(cps, dtype=np.uint8)
        n = whites.shape[0]
        counts = np.zeros((n, 4), dtype=np.uint32)
        total = self._batch(whites.ctypes.data, blacks.ctypes.data, cps.ctypes.data, n, playouts_per_leaf,
                            max_plies, key, first_id, counts.ctypes.data)
        return counts, int(total)


_cache = {}


def port():
    if "port" not in _cache:
        p = os.path.join(_HERE, "libloa_oracle.so")
        if not os.path.exists(p):
            build(ref=False)
        _cache["port"] = LoaOracle(p, "loa_oracle_", "port")
    return _cache["port"]


def ref():
    """The reference's own code, or None if oracle/_ref has not been built (e.g. no /root/reference)."""
    if "ref" not in _cache:
        p = os.path.join(_HERE, "_ref", "libref_loa.so")
        _cache["ref"] = LoaOracle(p, "ref_loa_", "reference") if os.path.exists(p) else None
    return _cache["ref"]


def best():
    """Reference build when present, else the restatement (used as the CPU baseline arm)."""
    return ref() or port()
