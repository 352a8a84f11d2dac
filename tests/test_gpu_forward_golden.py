"""GPU: eval-mode `forward(imgs, K, R, t, depth_min, depth_max, ...)` of the three drop-in models, VALUES against what the
unmodified reference's forward returned for the same weights and images (tests/golden/forward.npz, written by
tests/golden/make_golden_forward.py).  This pins the glue of SURVEY.md 8-a7 / a13 / the CVP frontend -- K/4 and
build_proj_matrices, fill_cam_array and interval/128, depth hypotheses from the reference view's range, the cascade and
pyramid hand-over, the up-sampled confidence maps, the `map_views` batching of the 2-D extractors -- not just shapes."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_linf

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
DEPTH_TOL = 1e-3   # north_star: depth maps within 1e-3 relative L-inf of the reference PyTorch path

_spec = importlib.util.spec_from_file_location("make_golden_forward", os.path.join(GOLDEN, "make_golden_forward.py"))


def _build(kind):
    """Same construction as the generator (imported from it, minus its reference import at call time)."""
    import sys
    saved = sys.modules.get("oracle.ref_import")
    mod = importlib.util.module_from_spec(_spec)
    _spec.loader.exec_module(mod)          # imports oracle.ref_import (a module of test infrastructure) but never calls it
    if saved is None:
        sys.modules.pop("oracle.ref_import", None)
    return mod.build(kind)


@pytest.fixture(scope="module")
def fwd():
    return dict(np.load(os.path.join(GOLDEN, "forward.npz")))


@pytest.fixture(autouse=True)
def fp32_modules():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False   # the PyTorch 2-D extractors of Vis / CVP
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _run(kind):
    net, s, kw = _build(kind)
    net = net.to(DEV)
    s = {k: v.to(DEV) for k, v in s.items()}
    with torch.no_grad():
        return net(s["imgs"], s["K"], s["R"], s["t"], s["depth_min"], s["depth_max"], **kw)


def test_mvsnet_forward_values(fwd):
    out = _run("mvsnet")
    assert rel_linf(out["depth"].cpu().numpy(), fwd["mvsnet_depth"]) < DEPTH_TOL
    assert rel_linf(out["depth_est_list"][0].cpu().numpy(), fwd["mvsnet_est0"]) < DEPTH_TOL
    bad = np.abs(out["photometric_confidence"].cpu().numpy() - fwd["mvsnet_conf"]) > 1e-3
    assert bad.mean() < 0.01, bad.mean()      # .long() truncation boundaries only (tests/test_gpu_mvsnet.py asserts them per pixel)


def test_vis_forward_values(fwd):
    out = _run("vis")
    assert rel_linf(out["depth"].cpu().numpy(), fwd["vis_depth"]) < DEPTH_TOL
    for k in range(3):
        assert rel_linf(out["depth_est_list"][k].cpu().numpy(), fwd["vis_est%d" % k]) < DEPTH_TOL
        for v in range(2):
            est, heads = out["depth_pair_list"][k][v]
            assert rel_linf(est.cpu().numpy(), fwd["vis_pair%d_%d" % (k, v)]) < DEPTH_TOL
            assert np.abs(heads[0].cpu().numpy() - fwd["vis_uncert%d_%d" % (k, v)]).max() < 5e-3
    bad = np.abs(out["photometric_confidence"].cpu().numpy() - fwd["vis_conf"]) > 1e-3
    assert bad.mean() < 0.02, bad.mean()


def test_cvp_forward_values(fwd):
    out = _run("cvp")
    assert rel_linf(out["depth"].cpu().numpy(), fwd["cvp_depth"]) < DEPTH_TOL
    for k in range(2):
        assert rel_linf(out["depth_est_list"][k].cpu().numpy(), fwd["cvp_est%d" % k]) < DEPTH_TOL
    bad = np.abs(out["photometric_confidence"].cpu().numpy() - fwd["cvp_conf"]) > 1e-3
    assert bad.mean() < 0.02, bad.mean()
