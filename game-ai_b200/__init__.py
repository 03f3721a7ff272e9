"""game-ai_b200 -- B200 (sm_100a) rollout engine for the TomekBarabasz/Game-AI MCTS player.

Python-side binding of the C ABI in include/gameai_gpu.h (libgameai_gpu.so, built from csrc/ by
`make -C game-ai_b200/csrc`).  Used by tests/ and bench.py; the production caller is the C++
GPU-backed player in host/ (same ABI).  There is no CPU fallback: a missing library or a missing GPU
raises.
"""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgameai_gpu.so")
INCLUDE_DIR = os.path.join(os.path.dirname(_HERE), "include")

_u64, _u32, _i32, _vp = ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int, ctypes.c_void_p

GAI_OK, GAI_ERR_INVALID, GAI_ERR_CUDA, GAI_ERR_NO_DEVICE, GAI_ERR_NCCL, GAI_ERR_NOMEM = range(6)
MAX_MOVES = 96
OUT_WHITE, OUT_BLACK, OUT_DRAW, OUT_CAPPED = range(4)


class GaiError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("gameai_gpu error %d: %s" % (code, msg))
        self.code = code


class RolloutStats(ctypes.Structure):
    _fields_ = [("playouts", _u64), ("plies", _u64), ("kernel_ms", ctypes.c_float), ("total_ms", ctypes.c_float)]


class DeviceInfo(ctypes.Structure):
    _fields_ = [("device", _i32), ("sm_count", _i32), ("cc_major", _i32), ("cc_minor", _i32), ("clock_khz", _i32),
                ("total_mem", _u64), ("name", ctypes.c_char * 64)]


def build(verbose=False):
    """Compile csrc/ for sm_100a into libgameai_gpu.so (in-tree, next to this file)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc")] + ([] if verbose else ["-s"])
    subprocess.check_call(cmd)
    return LIB_PATH


_lib = None


def lib():
    """The loaded C-ABI library.  Raises if it has not been built -- there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(LIB_PATH + " is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                    "(or make -C game-ai_b200/csrc).  The rollout engine has no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        L.gai_last_error.restype = ctypes.c_char_p
        L.gai_last_error.argtypes = [_vp]
        L.gai_launch_count.restype = _u64
        L.gai_launch_count.argtypes = [_vp]
        L.gai_create.argtypes = [_i32, ctypes.POINTER(_vp)]
        L.gai_destroy.argtypes = [_vp]
        L.gai_sync.argtypes = [_vp]
        L.gai_get_device_info.argtypes = [_vp, _vp]
        L.gai_set_launch_config.argtypes = [_vp, _i32, _i32]
        L.gai_get_launch_config.argtypes = [_vp, _vp, _vp]
        L.gai_set_rules_impl.argtypes = [_vp, _i32]
        L.gai_loa_movegen.argtypes = [_vp, _vp, _u32, _vp, _vp]
        L.gai_loa_apply.argtypes = [_vp, _vp, _vp, _u32, _vp]
        L.gai_loa_terminal.argtypes = [_vp, _vp, _u32, _vp, _vp]
        L.gai_loa_successor_hash.argtypes = [_vp, _vp, _u32, _vp]
        L.gai_loa_rollout.argtypes = [_vp, _vp, _u32, _u32, _u32, _u64, _u64, _vp, _vp]
        L.gai_loa_rollout_async.argtypes = [_vp, _vp, _u32, _u32, _u32, _u64, _u64, _vp]
        L.gai_last_rollout_stats.argtypes = [_vp, _vp]
        L.gai_loa_rollout_submit.argtypes = [_vp, _vp, _u32, _u32, _u32, _u64, _u64, _vp, ctypes.POINTER(_i32)]
        L.gai_loa_rollout_children_submit.argtypes = [_vp, _vp, _u32, _vp, _u32, _u32, _u64, _u64, _vp, ctypes.POINTER(_i32)]
        L.gai_wait.argtypes = [_vp, _i32, _vp]
        L.gai_inflight.argtypes = [_vp]
        L.gai_loa_rollout_device.argtypes = [_vp, _vp, _vp, _u32, _u32, _u32, _u64, _u64, _vp, _vp]
        L.gai_loa_playout_trace.argtypes = [_vp, _vp, _u32, _u32, _u64, _u64, _vp, _vp, _vp]
        L.gai_loa_gen_reachable.argtypes = [_vp, _u32, _u64, _u32, _vp, _vp, _vp]
        L.gai_pan_movegen.argtypes = [_vp, _vp, _u32, _i32, _vp, _vp]
        L.gai_pan_apply.argtypes = [_vp, _vp, _vp, _u32, _i32, _vp]
        L.gai_pan_terminal.argtypes = [_vp, _vp, _u32, _i32, _vp, _vp, _vp, _vp]
        L.gai_pan_rollout.argtypes = [_vp, _vp, _u32, _i32, _u32, _u32, _i32, _u64, _u64, _vp, _vp]
        L.gai_pan_rollout_device.argtypes = [_vp, _vp, _u32, _i32, _u32, _u32, _i32, _u64, _u64, _vp, _vp]
        L.gai_pan_playout_trace.argtypes = [_vp, _vp, _u32, _i32, _u32, _i32, _u64, _u64, _vp, _vp, _vp, _vp]
        L.gai_pan_determinize.argtypes = [_vp, _i32, _i32, _u32, _u64, _u64, _vp]
        L.gai_dev_alloc.argtypes = [_vp, _u64, ctypes.POINTER(_vp)]
        L.gai_dev_free.argtypes = [_vp, _vp]
        L.gai_memcpy_h2d.argtypes = [_vp, _vp, _vp, _u64]
        L.gai_memcpy_d2h.argtypes = [_vp, _vp, _vp, _u64]
        L.gai_loa_pack_states.argtypes = [_vp, _u32, _vp, _vp]
        L.gai_philox4x32_10.argtypes = [_vp, _vp, _vp]
        L.gai_comm_unique_id.argtypes = [_vp]
        L.gai_comm_init.argtypes = [_vp, _i32, _i32, _vp]
        L.gai_allreduce_root.argtypes = [_vp, _vp, _vp, _u32]
        L.gai_allreduce_u64.argtypes = [_vp, _vp, _u32]
        _lib = L
    return _lib


def philox4x32_10(ctr, key):
    """Host evaluation of the product's Philox (include/gameai/philox.h) -- no GPU needed."""
    c = (ctypes.c_uint32 * 4)(*ctr); k = (ctypes.c_uint32 * 2)(*key); o = (ctypes.c_uint32 * 4)()
    lib().gai_philox4x32_10(c, k, o)
    return tuple(o)


def pan_determinize(known_state, num_players, player, n_samples, key=0x10A5EED, first_id=0):
    """Sample complete-information deals consistent with a player's view (host only, no GPU). -> (n_samples, 5) uint64"""
    k = np.ascontiguousarray(known_state, dtype=np.uint64).reshape(5)
    out = np.zeros((n_samples, 5), np.uint64)
    r = lib().gai_pan_determinize(k.ctypes.data, num_players, player, n_samples, key, first_id, out.ctypes.data)
    if r:
        raise GaiError(r, "gai_pan_determinize: own hand / stack not certain, or the public hand sizes do not add up to the unseen cards")
    return out


def pack_states(states):
    """(n,3) uint64 reference-layout states -> (boards (n,2) uint64, stm_bits (ceil(n/32),) uint32). Host only."""
    s = _states(states); n = s.shape[0]
    boards = np.zeros((n, 2), np.uint64); stm = np.zeros((n + 31) // 32, np.uint32)
    r = lib().gai_loa_pack_states(s.ctypes.data, n, boards.ctypes.data, stm.ctypes.data)
    if r:
        raise GaiError(r, "gai_loa_pack_states")
    return boards, stm


def _states(states):
    s = np.ascontiguousarray(states, dtype=np.uint64)
    if s.ndim != 2 or s.shape[1] != 3:
        raise ValueError("states must have shape (n, 3): white, black, current_player")
    return s


class GpuRollout:
    """One context (device + stream + scratch) of the rollout engine."""

    def __init__(self, device=0):
        self.L = lib()
        h = _vp()
        r = self.L.gai_create(device, ctypes.byref(h))
        if r:
            raise GaiError(r, self.L.gai_last_error(None).decode())
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.L.gai_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, r):
        if r:
            raise GaiError(r, self.L.gai_last_error(self.h).decode())

    # -- info -------------------------------------------------------------------------------------
    def device_info(self):
        d = DeviceInfo()
        self._ck(self.L.gai_get_device_info(self.h, ctypes.byref(d)))
        return {"device": d.device, "sm_count": d.sm_count, "cc": (d.cc_major, d.cc_minor), "clock_khz": d.clock_khz,
                "total_mem": d.total_mem, "name": d.name.decode()}

    def launch_count(self):
        return int(self.L.gai_launch_count(self.h))

    def set_launch_config(self, threads_per_block=0, blocks_per_sm=0):
        self._ck(self.L.gai_set_launch_config(self.h, threads_per_block, blocks_per_sm))

    def launch_config(self):
        t, u = ctypes.c_int(0), ctypes.c_int(0)
        self._ck(self.L.gai_get_launch_config(self.h, ctypes.byref(t), ctypes.byref(u)))
        return {"threads_per_block": t.value, "plies_per_trip": u.value}

    def set_rules_impl(self, impl):
        """0 = bitboard baseline kernels, 1 = rotated boards + line LUT per piece, 2 = static line sweep."""
        self._ck(self.L.gai_set_rules_impl(self.h, impl))

    def sync(self):
        self._ck(self.L.gai_sync(self.h))

    # -- tier-1 rules sweep -----------------------------------------------------------------------
    def movegen(self, states):
        s = _states(states); n = s.shape[0]
        counts = np.zeros(n, np.uint8); moves = np.zeros((n, MAX_MOVES, 2), np.uint8)
        self._ck(self.L.gai_loa_movegen(self.h, s.ctypes.data, n, counts.ctypes.data, moves.ctypes.data))
        return counts, moves

    def apply(self, states, mv):
        s = _states(states); n = s.shape[0]
        mv = np.ascontiguousarray(mv, dtype=np.uint8)
        assert mv.shape == (n, 2)
        out = np.zeros((n, 3), np.uint64)
        self._ck(self.L.gai_loa_apply(self.h, s.ctypes.data, mv.ctypes.data, n, out.ctypes.data))
        return out

    def terminal(self, states):
        s = _states(states); n = s.shape[0]
        t = np.zeros(n, np.uint8); sc = np.zeros((n, 2), np.uint8)
        self._ck(self.L.gai_loa_terminal(self.h, s.ctypes.data, n, t.ctypes.data, sc.ctypes.data))
        return t, sc

    def successor_hash(self, states):
        s = _states(states); n = s.shape[0]
        h = np.zeros(n, np.uint64)
        self._ck(self.L.gai_loa_successor_hash(self.h, s.ctypes.data, n, h.ctypes.data))
        return h

    # -- rollouts ---------------------------------------------------------------------------------
    def rollout(self, states, playouts_per_leaf=1, max_plies=50000, key=0x10A5EED, first_id=0):
        """HOST buffers in, HOST counts out (the plugin-facing call). -> (counts (n,4) uint32, stats dict)"""
        s = _states(states); n = s.shape[0]
        out = np.zeros((n, 4), np.uint32)
        st = RolloutStats()
        self._ck(self.L.gai_loa_rollout(self.h, s.ctypes.data, n, playouts_per_leaf, max_plies, key, first_id,
                                        out.ctypes.data, ctypes.byref(st)))
        return out, _stats(st)

    def rollout_into(self, states, out, playouts_per_leaf, max_plies, key, first_id):
        """Same call with caller-owned (e.g. pinned) numpy buffers; returns the stats dict."""
        st = RolloutStats()
        self._ck(self.L.gai_loa_rollout(self.h, states.ctypes.data, states.shape[0], playouts_per_leaf, max_plies, key, first_id,
                                        out.ctypes.data, ctypes.byref(st)))
        return _stats(st)

    def rollout_async(self, states, out, playouts_per_leaf, max_plies, key, first_id):
        """Enqueue and return; `out` (n,4) uint32 is valid after sync().  One batch in flight per context."""
        self._ck(self.L.gai_loa_rollout_async(self.h, states.ctypes.data, states.shape[0], playouts_per_leaf, max_plies, key, first_id, out.ctypes.data))

    # ticketed batches: up to GAI_MAX_INFLIGHT queued per context
    def submit(self, states, out, playouts_per_leaf, max_plies, key, first_id):
        """Queue a leaf batch; `states` (n,3) uint64 and `out` (n,4) uint32 must stay alive until wait(ticket)."""
        t = _i32(-1)
        self._ck(self.L.gai_loa_rollout_submit(self.h, states.ctypes.data, states.shape[0], playouts_per_leaf, max_plies, key, first_id,
                                               out.ctypes.data, ctypes.byref(t)))
        return t.value

    def submit_children(self, parents, child_offsets, out, playouts_per_leaf, max_plies, key, first_id):
        """Queue a batch-by-expansion: every successor of every parent is a leaf (out: (child_offsets[-1], 4) uint32)."""
        t = _i32(-1)
        off = np.ascontiguousarray(child_offsets, dtype=np.uint32)
        assert off.shape[0] == parents.shape[0] + 1
        self._ck(self.L.gai_loa_rollout_children_submit(self.h, parents.ctypes.data, parents.shape[0], off.ctypes.data, playouts_per_leaf, max_plies,
                                                        key, first_id, out.ctypes.data, ctypes.byref(t)))
        return t.value

    def wait(self, ticket):
        st = RolloutStats()
        self._ck(self.L.gai_wait(self.h, ticket, ctypes.byref(st)))
        return _stats(st)

    def inflight(self):
        return int(self.L.gai_inflight(self.h))

    def last_rollout_stats(self):
        st = RolloutStats()
        self._ck(self.L.gai_last_rollout_stats(self.h, ctypes.byref(st)))
        return _stats(st)

    def rollout_device(self, d_boards, d_stm, n_leaves, d_out, playouts_per_leaf=1, max_plies=50000, key=0x10A5EED, first_id=0):
        """Inputs resident in HBM (raw device pointers as ints). -> stats dict"""
        st = RolloutStats()
        self._ck(self.L.gai_loa_rollout_device(self.h, d_boards, d_stm, n_leaves, playouts_per_leaf, max_plies, key, first_id,
                                               d_out, ctypes.byref(st)))
        return _stats(st)

    def playout_trace(self, states, max_plies=50000, key=0x10A5EED, first_id=0):
        s = _states(states); n = s.shape[0]
        o = np.zeros(n, np.uint8); pl = np.zeros(n, np.uint32); f = np.zeros((n, 3), np.uint64)
        self._ck(self.L.gai_loa_playout_trace(self.h, s.ctypes.data, n, max_plies, key, first_id,
                                              o.ctypes.data, pl.ctypes.data, f.ctypes.data))
        return o, pl, f

    def gen_reachable(self, n, key=0x10A5EED, max_k=300, d_boards=None, d_stm=None, host=True):
        """Synthetic reachable corpus; optionally written straight to device buffers (rollout layout)."""
        s = np.zeros((n, 3), np.uint64) if host else None
        self._ck(self.L.gai_loa_gen_reachable(self.h, n, key, max_k, d_boards, d_stm, s.ctypes.data if host else None))
        return s

    # -- Pan (determinized) ---------------------------------------------------------------------
    @staticmethod
    def _pstates(states):
        s = np.ascontiguousarray(states, dtype=np.uint64)
        if s.ndim != 2 or s.shape[1] != 5:
            raise ValueError("Pan states must have shape (n, 5): the reference's raw 40-byte GameState")
        return s

    def pan_movegen(self, num_players, states):
        s = self._pstates(states); n = s.shape[0]
        c = np.zeros(n, np.uint8); m = np.zeros((n, 16), np.uint64)
        self._ck(self.L.gai_pan_movegen(self.h, s.ctypes.data, n, num_players, c.ctypes.data, m.ctypes.data))
        return c, m

    def pan_apply(self, num_players, states, moves):
        s = self._pstates(states); n = s.shape[0]
        mv = np.ascontiguousarray(moves, dtype=np.uint64); o = np.zeros((n, 5), np.uint64)
        self._ck(self.L.gai_pan_apply(self.h, s.ctypes.data, mv.ctypes.data, n, num_players, o.ctypes.data))
        return o

    def pan_terminal(self, num_players, states):
        s = self._pstates(states); n = s.shape[0]
        t = np.zeros(n, np.uint8); sc = np.zeros((n, 4), np.int32); e0 = np.zeros((n, 4), np.int32); e1 = np.zeros((n, 4), np.int32)
        self._ck(self.L.gai_pan_terminal(self.h, s.ctypes.data, n, num_players, t.ctypes.data, sc.ctypes.data, e0.ctypes.data, e1.ctypes.data))
        return t, sc, e0, e1

    def pan_rollout(self, num_players, states, playouts_per_leaf=1, max_plies=50, eval_kind=1, key=0x10A5EED, first_id=0):
        """-> (res (n,12) uint32: losses[4], eval_sum[4], capped, pad[3]; stats dict)"""
        s = self._pstates(states); n = s.shape[0]
        r = np.zeros((n, 12), np.uint32); st = RolloutStats()
        self._ck(self.L.gai_pan_rollout(self.h, s.ctypes.data, n, num_players, playouts_per_leaf, max_plies, eval_kind, key, first_id,
                                        r.ctypes.data, ctypes.byref(st)))
        return r, _stats(st)

    def pan_rollout_into(self, num_players, states, out, playouts_per_leaf=1, max_plies=50, eval_kind=1, key=0x10A5EED, first_id=0):
        """gai_pan_rollout with caller-owned buffers (pinned ones are DMA-ed in place): states (n,5) uint64, out (n,12) uint32"""
        assert states.dtype == np.uint64 and states.flags.c_contiguous and out.dtype == np.uint32 and out.flags.c_contiguous
        n = states.shape[0]; st = RolloutStats()
        assert out.size == 12 * n
        self._ck(self.L.gai_pan_rollout(self.h, states.ctypes.data, n, num_players, playouts_per_leaf, max_plies, eval_kind, key, first_id,
                                        out.ctypes.data, ctypes.byref(st)))
        return _stats(st)

    def pan_rollout_device(self, num_players, d_states, n, d_out, playouts_per_leaf=1, max_plies=50, eval_kind=1, key=0x10A5EED, first_id=0):
        """Inputs resident in HBM (raw device pointers as ints): n x 40-byte states in, n x 48-byte results out. -> stats dict"""
        st = RolloutStats()
        self._ck(self.L.gai_pan_rollout_device(self.h, d_states, n, num_players, playouts_per_leaf, max_plies, eval_kind, key, first_id,
                                               d_out, ctypes.byref(st)))
        return _stats(st)

    def pan_playout_trace(self, num_players, states, max_plies=50, eval_kind=1, key=0x10A5EED, first_id=0):
        s = self._pstates(states); n = s.shape[0]
        c = np.zeros(n, np.uint8); pl = np.zeros(n, np.uint32); v = np.zeros((n, 4), np.int32); f = np.zeros((n, 5), np.uint64)
        self._ck(self.L.gai_pan_playout_trace(self.h, s.ctypes.data, n, num_players, max_plies, eval_kind, key, first_id,
                                              c.ctypes.data, pl.ctypes.data, v.ctypes.data, f.ctypes.data))
        return c, pl, v, f

    # -- device memory ----------------------------------------------------------------------------
    def dev_alloc(self, nbytes):
        p = _vp()
        self._ck(self.L.gai_dev_alloc(self.h, nbytes, ctypes.byref(p)))
        return p.value

    def dev_free(self, p):
        self._ck(self.L.gai_dev_free(self.h, p))

    def h2d(self, d_dst, arr):
        arr = np.ascontiguousarray(arr)
        self._ck(self.L.gai_memcpy_h2d(self.h, d_dst, arr.ctypes.data, arr.nbytes))

    def d2h(self, arr, d_src):
        self._ck(self.L.gai_memcpy_d2h(self.h, arr.ctypes.data, d_src, arr.nbytes))

    # -- multi-GPU --------------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id():
        buf = (ctypes.c_uint8 * 128)()
        r = lib().gai_comm_unique_id(buf)
        if r:
            raise GaiError(r, lib().gai_last_error(None).decode())
        return bytes(buf)

    def comm_init(self, rank, world, unique_id):
        buf = (ctypes.c_uint8 * 128).from_buffer_copy(unique_id)
        self._ck(self.L.gai_comm_init(self.h, rank, world, buf))

    def allreduce_root(self, visits, values):
        visits = np.ascontiguousarray(visits, dtype=np.uint32); values = np.ascontiguousarray(values, dtype=np.float32)
        assert values.shape == (visits.shape[0], 2)
        self._ck(self.L.gai_allreduce_root(self.h, visits.ctypes.data, values.ctypes.data, visits.shape[0]))
        return visits, values

    def allreduce_u64(self, v):
        v = np.ascontiguousarray(v, dtype=np.uint64)
        self._ck(self.L.gai_allreduce_u64(self.h, v.ctypes.data, v.shape[0]))
        return v


def _stats(st):
    return {"playouts": int(st.playouts), "plies": int(st.plies), "kernel_ms": float(st.kernel_ms), "total_ms": float(st.total_ms)}
