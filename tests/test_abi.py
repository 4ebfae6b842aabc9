"""The C-ABI library loads on a CPU-only box, exports every symbol include/legkilo_b200.h declares,
its POD layouts match the Python mirrors, and — with no GPU — it fails loudly instead of falling back."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import conftest
from legkilo_b200 import HEADER_PATH, LIB_PATH, Engine, LkError, abi, lib


def _declared():
    src = open(HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lk_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _declared()
    assert len(names) >= 25, names
    L = C.CDLL(LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert lib().lk_abi_version() == 1


def test_struct_layouts_match_header():
    assert C.sizeof(abi.LkState) == 36 * 8 == abi.STATE_DTYPE.itemsize
    assert C.sizeof(abi.LkEskfCfg) == 14 * 8
    assert C.sizeof(abi.LkMapCfg) == 6 * 8 + 12 * 4
    assert C.sizeof(abi.LkStreamClock) == 16 == abi.CLOCK_DTYPE.itemsize
    assert C.sizeof(abi.LkImuMeas) == 56 == abi.IMU_DTYPE.itemsize
    assert C.sizeof(abi.LkKinImuMeas) == abi.KINIMU_DTYPE.itemsize == 8 + 96 + 96 + 16 + 24 + 24
    hdr = open(HEADER_PATH).read()
    for tag, size in (("lk_map_blob_header; /* 32 B */", 32), ("lk_map_root; /* 16 B */", 16), ("lk_map_node; /* 256 B */", 256),
                      ("lk_map_aux; /* 64 B */", 64), ("lk_map_point; /* 72 B */", 72)):
        assert tag in hdr
    assert abi.MAP_NODE_DTYPE.itemsize == 256 and abi.MAP_AUX_DTYPE.itemsize == 64


def test_host_helpers_without_device():
    Q = np.zeros(900)
    ec = abi.eskf_cfg(abi.CONFIGS["leg_fusion"])
    assert lib().lk_init_process_cov(C.byref(ec), Q.ctypes.data_as(C.c_void_p)) == 0
    np.testing.assert_array_equal(Q, abi.process_cov_Q(abi.CONFIGS["leg_fusion"]))  # eskf.cc:47-62
    x = np.zeros(1, abi.STATE_DTYPE)
    assert lib().lk_state_default(x.ctypes.data_as(C.c_void_p)) == 0
    assert x.tobytes() == abi.default_states(1).tobytes()
    assert lib().lk_create(None, None, None, None, 0, None) == -1  # LK_ERR_INVALID_ARG, never a crash


@pytest.mark.skipif(conftest._has_gpu(), reason="needs a box WITHOUT a GPU")
def test_no_gpu_means_loud_failure_not_fallback():
    with pytest.raises(LkError) as e:
        Engine(abi.CONFIGS["leg_fusion"])
    assert e.value.code == -3 and "no CPU fallback" in str(e.value)  # LK_ERR_NO_DEVICE


def test_product_never_touches_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use oracle/."""
    root = os.path.dirname(os.path.dirname(HEADER_PATH))
    pkg = os.path.join(root, "leg-kilo_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in txt.lower().replace("oracle's", "").replace("the cpu oracle", "").replace("cpu oracle", "") or f in (
                    "synth.py", "__init__.py", "abi.py", "shard.py"), (dp, f)
                assert "import lko" not in txt and "liblko" not in txt, (dp, f)
