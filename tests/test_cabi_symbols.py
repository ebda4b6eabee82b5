"""The C-ABI library loads on a CPU-only box and exports every symbol include/nanovs.h declares
(no compute calls here: there is no GPU)."""
import ctypes
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(REPO, "include", "nanovs.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nvs_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    from nano_vs_slam_b200 import _cabi, build

    if not os.path.exists(_cabi.LIB_PATH):
        build.build(verbose=False)
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in nanovs.h but not exported by libnanovs.so"
        assert n in _cabi.SIGNATURES, f"{n} has no ctypes signature in _cabi.SIGNATURES"
    assert set(_cabi.SIGNATURES) == set(names)


def test_pure_host_entry_points():
    from nano_vs_slam_b200 import _cabi

    lib = _cabi.lib()
    assert lib.nvs_abi_version() == 3
    assert [lib.nvs_conv_cout_tile(c) for c in (1, 3, 16, 19, 28, 32, 48, 64, 96, 128)] == \
        [8, 8, 16, 24, 32, 32, 48, 64, 48, 64]
    assert lib.nvs_conv_cin_chunk(3) == 4 and lib.nvs_conv_cin_chunk(96) == 8
    assert lib.nvs_flat_padded_dim(4096) == 4096 and lib.nvs_flat_padded_dim(100) == 128
    assert lib.nvs_flat_search_workspace_bytes(125000, 10000, 4096, 25) > 0
    assert lib.nvs_netvlad_workspace_bytes(256, 64, 64, 4800) == 4 * 256 * 5 * (64 * 64 + 64)  # 5 slices per frame
    # without a GPU the device probe must say so (and never crash)
    import torch
    if not torch.cuda.is_available():
        assert lib.nvs_device_ok() == -4


def test_ops_refuse_cpu_tensors():
    import torch
    from nano_vs_slam_b200 import ops
    from nano_vs_slam_b200._cabi import NanovsError

    with pytest.raises(NanovsError):
        ops.softmax_channels(torch.zeros(1, 4, 2, 2))
    with pytest.raises(NanovsError):
        ops.decode(torch.zeros(1, 1, 4, 4), torch.zeros(1, 2, 4, 4), None, 16, 16, 4)
