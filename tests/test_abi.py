"""The C ABI without a GPU: libhdp_b200.so builds / loads, exports every function include/hdp_b200.h declares, the ctypes
binding covers exactly that set, and the host-only entry points (version, error strings, workspace sizes) answer."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "hdp_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)            # comments mention entry points too
    return sorted(set(re.findall(r"\b(hdp_b200_\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from hdp_b200 import _lib
    return _lib.lib()                                              # builds with nvcc first if the library is missing or stale


def test_header_declares_the_two_paths_and_their_helpers():
    names = declared_functions()
    for must in ("hdp_b200_thresholds", "hdp_b200_metrics", "hdp_b200_thresholds_host", "hdp_b200_metrics_host",
                 "hdp_b200_thresholds_workspace_bytes", "hdp_b200_metrics_workspace_bytes", "hdp_b200_hot_days",
                 "hdp_b200_heat_index", "hdp_b200_strerror", "hdp_b200_abi_version"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    raw = ctypes.CDLL(lib._name)
    missing = [n for n in declared_functions() if not hasattr(raw, n)]
    assert not missing, f"declared in include/hdp_b200.h but not exported: {missing}"


def test_ctypes_binding_matches_the_header():
    from hdp_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_functions()


def test_host_only_entry_points(lib):
    text = open(HEADER).read()
    assert lib.hdp_b200_abi_version() == int(re.search(r"#define HDP_B200_ABI_VERSION\s+(\d+)", text).group(1))
    codes = {name: int(v) for name, v in re.findall(r"#define (HDP_B200_ERR_\w+)\s+\((-\d+)\)", text)}
    assert len(codes) >= 5
    seen = set()
    for name, code in codes.items():
        msg = lib.hdp_b200_strerror(code).decode()
        assert msg and msg not in seen, name
        seen.add(msg)
    assert lib.hdp_b200_strerror(0).decode()
    # workspace sizes are pure host arithmetic: positive, monotone in the number of cells, 0 for invalid shapes
    a = lib.hdp_b200_thresholds_workspace_bytes(64, 10950, 64, 1, 365, 30, 15, 10, 0)
    b = lib.hdp_b200_thresholds_workspace_bytes(64, 10950, 1, 10950, 365, 30, 15, 10, 0)     # time-contiguous: + the transposed copy
    assert lib.hdp_b200_thresholds_workspace_bytes(64, 10950, 64, 1, 365, 30, 15, 10, 1) == b     # Kelvin input: room for a converted copy
    assert 0 < a < b and b - a >= 64 * 10950 * 4
    assert lib.hdp_b200_thresholds_workspace_bytes(-1, 10950, 64, 1, 365, 30, 15, 10, 0) == 0
    m1 = lib.hdp_b200_metrics_workspace_bytes(64, 31390, 64, 1, 365, 10, 6, 86, None)
    m2 = lib.hdp_b200_metrics_workspace_bytes(128, 31390, 128, 1, 365, 10, 6, 86, None)
    assert 0 < m1 < m2
    assert lib.hdp_b200_launch_count() >= 0
