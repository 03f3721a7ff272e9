"""INTEGRATION.md section 2 as a binary: the reference's own MC::Player with oracle/mcts_gpu_stub.patch applied (simulation
loop cut at the first new state, playouts through gai_loa_rollout) -- oracle/_ref/libref_mcts_gpu.so, built by
`make -C oracle ref` from /root/reference.  On the CPU it runs against the test-only fake device (oracle playouts); the GPU
variant of this test (tests/test_gpu_integration_stub.py) runs the same binary against the real library."""
import ctypes
import os
import subprocess
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_ref", "libref_mcts_gpu.so")

SNIPPET = r"""
import ctypes, json, sys
L = ctypes.CDLL(%r)
L.ref_mcts_gpu_play.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_uint] + [ctypes.c_void_p] * 4 + [ctypes.c_int]
mv = (ctypes.c_uint8 * 64)(); po = ctypes.c_ulonglong(); sec = ctypes.c_double(); err = ctypes.create_string_buffer(256)
n = L.ref_mcts_gpu_play(int(sys.argv[1]), int(sys.argv[2]), 1234, mv, ctypes.byref(po), ctypes.byref(sec), err, 256)
print(json.dumps({"n": n, "moves": [(mv[2 * i], mv[2 * i + 1]) for i in range(max(n, 0))], "playouts": po.value, "seconds": sec.value, "err": err.value.decode()}))
"""


def play(n_moves, sims, fake):
    import json
    env = dict(os.environ)
    if fake:
        env["LD_PRELOAD"] = fake
    r = subprocess.run([sys.executable, "-c", SNIPPET % SO, str(n_moves), str(sims)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-1500:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def check(res, n_moves, sims):
    from oracle import loa
    assert res["n"] == n_moves, res
    # every move the patched reference player chose is legal in the position it was made in... replaying needs the random
    # opponent's moves too, so check what is observable: distinct squares, on the board, and the GPU did the playouts
    for f, t in res["moves"]:
        assert 0 <= f < 64 and 0 <= t < 64 and f != t
    assert res["playouts"] >= n_moves * (sims // 16) * 8 * 0.5           # batches of 16 selections x 8 playouts per leaf
    first = tuple(res["moves"][0])
    assert first in [tuple(m) for m in loa.port().movegen(0x0081818181818100, 0x7e0000000000007e, 1)]


def test_patched_reference_player_plays_through_the_c_abi_on_the_fake_device():
    if not os.path.exists(SO):
        pytest.skip("oracle/_ref/libref_mcts_gpu.so not built (needs /root/reference at build time)")
    fake = os.path.join(ROOT, "tests", "helpers", "libfakegpu.so")
    src = os.path.join(ROOT, "tests", "helpers", "fake_gpu.c")
    if not os.path.exists(fake) or os.path.getmtime(fake) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-std=c11", "-O2", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), "-o", fake, src,
                               "-L", os.path.join(ROOT, "oracle"), "-lloa_oracle", "-lpan_oracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    res = play(3, 64, fake)
    check(res, 3, 64)
