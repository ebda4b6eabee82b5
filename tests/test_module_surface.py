"""Drop-in surface: config letters, constructor kwargs, attributes and state_dict names/shapes equal the
reference's (checked against the live reference package when it is present, else against the golden files)."""
import contextlib
import io
import os
import sys

import pytest
import torch

from util import golden_cases, load_golden

REF = "/root/reference"


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def test_config_tables_and_factory():
    from nano_vs_slam_b200 import KP2DTINY_CONFIGS, KP2DTINYV3_CONFIGS, get_config, tiny_factory

    assert set(KP2DTINY_CONFIGS) == {"S", "S_A", "N", "N_A", "D", "F", "GEM_N", "GEM_S_A", "CONVAP_S_A"}
    assert set(KP2DTINYV3_CONFIGS) == {"S", "S_A", "N", "N_A", "D", "D_A", "CONVAP_S_A"}
    with pytest.raises(ValueError):
        get_config("nope")
    c = get_config("S", to_export=True)
    assert c["remove_netvlad"] is True and "remove_netvlad" not in KP2DTINY_CONFIGS["S"]  # no shared-dict mutation
    m = _quiet(tiny_factory, "N", 28, v3=True)
    assert m.get_global_desc_dim() == 3072 and m.cell == 4 and m.nfeatures == 32 and m.training is True
    m2 = _quiet(tiny_factory, "N", 28, v3=False)
    assert m2.get_global_desc_dim() == 1536  # V2-N: num_clusters 32 x 48
    assert sum(p.numel() for p in m.parameters()) == 404639  # V3-N
    assert sum(p.numel() for p in _quiet(tiny_factory, "S", 28).parameters()) == 928079  # V2-S
    info = m.gather_info()
    assert info["netvlad_dim"] == 3072 and info["total_params"] == 404639
    assert _quiet(tiny_factory, "D", 28).get_global_desc_dim() == 128 * 16  # ConvAP, attention head_dim 64
    f = _quiet(tiny_factory, "F", 28)  # cell 8
    assert f.cell == 8 and f.encoder_dim == 128 and f.get_global_desc_dim() == 64 * 128
    assert _quiet(tiny_factory, "GEM_N", 28).get_global_desc_dim() == 48 * 16        # vpr.py:72
    assert _quiet(tiny_factory, "GEM_S_A", 28).get_global_desc_dim() == 64 * 16
    assert _quiet(tiny_factory, "CONVAP_S_A", 28, v3=True).get_global_desc_dim() == 64 * 16  # vpr.py:76


@pytest.mark.skipif(not os.path.isdir(REF), reason="live reference not present")
@pytest.mark.parametrize("letter,v3", [(l, v) for v in (False, True) for l in ("S", "S_A", "N", "N_A")] +
                         [("GEM_N", False), ("GEM_S_A", False), ("CONVAP_S_A", False), ("CONVAP_S_A", True), ("D", True),
                          ("D", False), ("D_A", True), ("F", False)])
def test_state_dict_keys_equal_reference(letter, v3):
    sys.path[:0] = [REF, os.path.join(REF, "src")]
    sys.dont_write_bytecode = True
    from src.kp2dtiny.models.kp2dtiny import tiny_factory as ref_factory  # type: ignore
    from nano_vs_slam_b200 import tiny_factory

    ref = _quiet(ref_factory, letter, 19, v3=v3).state_dict()
    ours = _quiet(tiny_factory, letter, 19, v3=v3)
    sd = ours.state_dict()
    assert list(sd.keys()) == list(ref.keys())
    for k in ref:
        assert tuple(sd[k].shape) == tuple(ref[k].shape), k
    ours.load_state_dict(ref, strict=True)  # a reference checkpoint loads unchanged


@pytest.mark.skipif(not os.path.isdir(REF), reason="live reference not present")
@pytest.mark.parametrize("letter,v3", [("S", False), ("S_A", False), ("N", True), ("S_A", True)])
def test_state_dict_keys_equal_reference_with_depth(letter, v3):
    """depth=True: V2 grows a second segmentation head, V3 a third trunk slice + featD (kp2dtiny.py:402-437)."""
    sys.path[:0] = [REF, os.path.join(REF, "src")]
    sys.dont_write_bytecode = True
    from src.kp2dtiny.models import kp2dtiny as ref  # type: ignore
    from util import build_model

    cls = ref.KP2DTinyV3 if v3 else ref.KP2DTinyV2
    rsd = _quiet(lambda: cls(**dict(ref.get_config(letter, v3=v3)), nClasses=19, depth=True)).state_dict()
    ours = build_model(letter, 19, v3, depth=True)
    sd = ours.state_dict()
    assert list(sd.keys()) == list(rsd.keys())
    for k in rsd:
        assert tuple(sd[k].shape) == tuple(rsd[k].shape), k


@pytest.mark.skipif(not os.path.isdir(REF), reason="live reference not present")
@pytest.mark.parametrize("letter,v3", [("S", False), ("N_A", False), ("N", True), ("S_A", True)])
def test_state_dict_keys_equal_reference_to_mcu(letter, v3):
    """to_mcu=True: ConvTranspose upsampling modules appear (desc_head.upsample first, seg_head.upsample{,2} last)."""
    import copy

    sys.path[:0] = [REF, os.path.join(REF, "src")]
    sys.dont_write_bytecode = True
    from src.kp2dtiny.models import kp2dtiny as ref  # type: ignore
    from nano_vs_slam_b200 import tiny_factory

    saved = copy.deepcopy((ref.KP2DTINY_CONFIGS, ref.KP2DTINYV3_CONFIGS))
    try:  # the reference's get_config mutates its shared dicts (kp2dtiny.py:271-274): restore them afterwards
        rsd = _quiet(ref.tiny_factory, letter, 19, to_mcu=True, v3=v3).state_dict()
    finally:
        for dst, src in zip((ref.KP2DTINY_CONFIGS, ref.KP2DTINYV3_CONFIGS), saved):
            dst.clear()
            dst.update(src)
    ours = _quiet(tiny_factory, letter, 19, to_mcu=True, v3=v3)
    assert ours.upscale_method == "convtranspose" and ours.leaky_relu is False
    sd = ours.state_dict()
    assert list(sd.keys()) == list(rsd.keys())
    for k in rsd:
        assert tuple(sd[k].shape) == tuple(rsd[k].shape), k


def test_load_state_dict_invalidates_packed_weights():
    from nano_vs_slam_b200 import tiny_factory
    from nano_vs_slam_b200.synthetic import spread_init

    m = _quiet(tiny_factory, "S", 28)
    m._packed = {"stale": True}
    m.load_state_dict(spread_init(m.state_dict(), 1))
    assert m._packed is None
    m._packed = {"stale": True}
    m.float()
    assert m._packed is None


def test_reference_import_path_shim():
    import importlib
    import subprocess

    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("from src.kp2dtiny.models.kp2dtiny import KP2DTinyV2, KP2DTinyV3, get_config, tiny_factory; "
            "import nano_vs_slam_b200 as n; assert KP2DTinyV2 is n.KP2DTinyV2; print('ok')")
    env = dict(os.environ, PYTHONPATH=os.path.join(repo, "nano_vs_slam_b200", "compat") + os.pathsep + repo)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd="/tmp")
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr
