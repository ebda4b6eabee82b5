"""Static checks on the product tree: the oracle is test infrastructure only, and the hot path has no compiler /
multi-backend layer."""
import os
import re

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "nano_vs_slam_b200")


def _sources(exts):
    for root, _, files in os.walk(PKG):
        if os.sep + "lib" in root:
            continue
        for f in files:
            if f.endswith(exts):
                yield os.path.join(root, f)


def test_product_never_imports_the_oracle_or_the_reference():
    pat = re.compile(r"^\s*(from|import)\s+(oracle|src\.|baseline)", re.M)
    for path in _sources((".py",)):
        if os.sep + "compat" + os.sep in path:
            continue  # the import-path shim re-exports nano_vs_slam_b200 under the reference's module name
        text = open(path).read()
        assert not pat.search(text), path
        assert "/root/reference" not in text, path


def test_no_compiler_or_multi_backend_layer_in_the_hot_path():
    banned = ("import triton", "torch.compile", "tilelang", "torch.jit.script")
    for path in _sources((".py",)):
        text = open(path).read()
        for b in banned:
            assert b not in text, (path, b)


def test_every_kernel_file_is_sm100a_cuda_and_the_build_targets_it():
    cu = list(_sources((".cu",)))
    assert len(cu) >= 9
    build = open(os.path.join(PKG, "build.py")).read()
    assert "arch=compute_100a,code=sm_100a" in build and "-lineinfo" in build
    tc = open(os.path.join(PKG, "csrc", "conv_tc.cu")).read() + open(os.path.join(PKG, "csrc", "retrieval.cu")).read()
    for needle in ("tcgen05.mma", "tcgen05.alloc", "tcgen05.commit", "cp.async.bulk.tensor", "mbarrier.try_wait"):
        assert needle in tc, needle


def test_ops_refuse_cpu_tensors_instead_of_falling_back():
    """No CPU fallback: every wrapper that takes tensors raises on host tensors before touching the library."""
    import pytest
    import torch

    from nano_vs_slam_b200 import _cabi, ops

    i32 = torch.zeros(1, dtype=torch.int32)
    with pytest.raises(_cabi.NanovsError):
        ops.pose_batch(torch.zeros(2, 8, 2), i32, i32 + 1, i32 + 8)
    with pytest.raises(_cabi.NanovsError):
        ops.match(torch.zeros(4, 32), torch.zeros(4, 32))
    with pytest.raises(_cabi.NanovsError):
        ops.select_keypoints(torch.zeros(1, 1, 4, 4), torch.zeros(1, 2, 4, 4), torch.zeros(1, 32, 4, 4), 0.5, 4)


def test_host_harness_is_test_infrastructure_only():
    """tests/host/pose_host.cpp compiles the product's algebra header for CPU checks; nothing in the product tree
    refers to it, and the library sources contain no host implementation of the pose pipeline."""
    for path in _sources((".py", ".cu", ".cuh", ".h")):
        text = open(path).read()
        if path.endswith("pose_math.h"):
            continue  # its header comment says who else compiles it
        assert "pose_host" not in text and "nvs_host_" not in text, path
