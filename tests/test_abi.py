"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/gameai_gpu.h
declares, and fails loudly (no CPU fallback) when there is no GPU.  No compute calls here."""
import ctypes
import os
import re
import numpy as np
import pytest
import gameai_b200 as g


def declared_symbols():
    src = open(os.path.join(g.INCLUDE_DIR, "gameai_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gai_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = g.lib()
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), "libgameai_gpu.so does not export " + s


def test_abi_version():
    assert g.lib().gai_abi_version() == 2


def test_struct_layouts_match_reference():
    # gai_loa_state mirrors the reference's 24-byte GameState (LinesOfActionZasady.cpp:38-46)
    assert ctypes.sizeof(g.RolloutStats) == 24
    s = np.zeros((2, 3), np.uint64)
    assert s.strides == (24, 8)


def test_pack_states_host_only():
    s = np.array([[1, 2, 1], [3, 4, 0], [5, 6, 0xFFFFFFFFFFFFFFFF]], dtype=np.uint64)   # upper bits of cp are ignored
    boards, stm = g.pack_states(s)
    assert boards.tolist() == [[1, 2], [3, 4], [5, 6]]
    assert stm.tolist() == [0b101]
    s = np.zeros((70, 3), np.uint64); s[::3, 2] = 1
    boards, stm = g.pack_states(s)
    bits = [(int(stm[i >> 5]) >> (i & 31)) & 1 for i in range(70)]
    assert bits == [1 if i % 3 == 0 else 0 for i in range(70)]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(g.GaiError) as e:
        g.GpuRollout(0)
    assert e.value.code == g.GAI_ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_link_the_oracle():
    # the shipped library must not depend on anything under oracle/
    import subprocess
    out = subprocess.check_output(["ldd", g.LIB_PATH]).decode()
    assert "oracle" not in out and "libref" not in out
    for root, _, files in os.walk(os.path.dirname(g.LIB_PATH)):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".cpp", ".py")):
                txt = open(os.path.join(root, f)).read()
                assert not re.search(r'#include\s*[<"][^>"]*oracle', txt), f
                assert not re.search(r'^\s*(from|import)\s+oracle', txt, flags=re.M), f
                assert "dlopen(\"oracle" not in txt and "libloa_oracle" not in txt, f
