"""The oracle (CPU restatement) against golden vectors produced by the REAL reference package
(oracle/gen_golden.py), and against the live reference when /root/reference is present."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest
import torch

from oracle import glue_ref, kp2dtiny_ref as R
from nano_vs_slam_b200.synthetic import spread_init, synthetic_frames
from util import golden_cases, load_golden, rel_err

REF = "/root/reference"


def _state_dict_for(case):
    # key names + shapes come from the product module tree (must equal the reference's)
    from util import build_model

    m = build_model(case["letter"], case["n_classes"], case["v3"], case["depth"], case["to_mcu"])
    return spread_init(m.state_dict(), case["wseed"])


@pytest.mark.parametrize("path", golden_cases(), ids=lambda p: os.path.basename(p)[6:-4])
def test_oracle_matches_reference_golden(path):
    c = load_golden(path)
    a = R.arch_for(c["letter"], c["v3"], c["n_classes"], depth=c["depth"], to_mcu=c["to_mcu"])
    sd = _state_dict_for(c)
    x = synthetic_frames(c["B"], c["H"], c["W"], c["xseed"])
    out = R.forward(x, sd, a)
    assert ("depth" in out) == c["depth"] == ("depth" in c["fwd"])
    for k in ("score", "coord", "feat", "vlad", "seg") + (("depth",) if c["depth"] else ()):
        assert out[k].shape == c["fwd"][k].shape, k
        assert rel_err(out[k], c["fwd"][k]) < 2e-5, (k, rel_err(out[k], c["fwd"][k]))
    post = R.post_processing(dict(out), c["H"], c["W"], a)
    for k in ("score", "coord", "feat"):
        assert rel_err(post[k], c["post"][k]) < 2e-5, k
    assert post["seg"].dtype == torch.int64 and post["seg"].shape == c["post"]["seg"].shape
    assert (post["seg"] == c["post"]["seg"]).float().mean() >= 0.999


def test_netvlad_factored_equals_literal():
    torch.manual_seed(0)
    x = torch.randn(2, 48, 9, 11)
    sd = {"vlad_head.netvlad.centroids": torch.rand(32, 48),
          "vlad_head.netvlad.conv.weight": torch.randn(32, 48, 1, 1)}
    assert rel_err(R.netvlad(x, sd), R.netvlad_literal(x, sd)) < 1e-5


@pytest.mark.skipif(not os.path.isdir(REF), reason="live reference not present (GPU box)")
def test_oracle_matches_live_reference_kitti_shape():
    sys.path[:0] = [REF, os.path.join(REF, "src")]
    sys.dont_write_bytecode = True
    from src.kp2dtiny.models.kp2dtiny import tiny_factory as ref_factory  # type: ignore

    with contextlib.redirect_stdout(io.StringIO()):
        m = ref_factory("S", 19, v3=False)
    sd = spread_init(m.state_dict(), 99)
    m.load_state_dict(sd)
    m.eval()
    m.training = False
    x = synthetic_frames(1, 72, 153, 5)  # odd W (floor pool) and odd W/8 = 19, like KITTI 376x1241
    with torch.no_grad():
        ref = m(x)
        ref_post = m.post_processing(dict(ref), 72, 153)
    a = R.arch_for("S", False, 19)
    out = R.forward(x, sd, a)
    post = R.post_processing(dict(out), 72, 153, a)
    for k in ("score", "coord", "feat", "vlad", "seg"):
        assert rel_err(out[k], ref[k]) < 2e-5, k
    for k in ("score", "coord", "feat"):
        assert rel_err(post[k], ref_post[k]) < 2e-5, k
    assert (post["seg"] == ref_post["seg"]).float().mean() >= 0.999


def test_matcher_oracle_against_cv2_golden():
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "matcher_cv2.npz"))
    idx, dist = glue_ref.knn2_l2(z["des1"], z["des2"])
    assert (idx == z["idx"]).mean() == 1.0
    np.testing.assert_allclose(dist, z["dist"], rtol=2e-5, atol=1e-6)
    qi, ti, _ = glue_ref.mutual_nn(z["des1"], z["des2"])
    assert sorted(zip(qi.tolist(), ti.tolist())) == sorted(map(tuple, z["cross"].tolist()))
    i1, i2, sc = glue_ref.good_matches_one_to_one(z["idx"], z["dist"], 0.7)
    assert len(i1) == len(set(i2)) and len(i1) > 20


def test_flat_l2_oracle_planted():
    from nano_vs_slam_b200.synthetic import planted_retrieval_set

    db, q, planted = planted_retrieval_set(3000, 40, 256, 10, seed=3)
    D, I = glue_ref.flat_l2_search(db, q, 10)
    assert torch.equal(I, planted)
    assert bool((D[:, 1:] - D[:, :-1] > 1e-3).all())
    # sharded == flat
    parts = [glue_ref.flat_l2_search(db[s:s + 1000], q, 10, id_offset=s) for s in range(0, 3000, 1000)]
    Dm, Im = glue_ref.merge_shard_topk([p[0] for p in parts], [p[1] for p in parts], 10)
    assert torch.equal(Im, I)
