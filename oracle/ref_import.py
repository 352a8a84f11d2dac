"""Import the UNMODIFIED reference (fdarmon/wild_deep_mvs) on a CPU-only host.

TEST INFRASTRUCTURE ONLY.  Used by tests/golden/make_golden.py (build container, where
/root/reference exists) to produce the golden tensors that pin the oracle.  Nothing that
runs on the GPU box imports this.

Two shims: the reference imports matplotlib at module import time
(utils/utils_3D.py:22-23, models/VisMVSNet/model_cas.py:8) and hard-codes .cuda() on the hot
path (VisMVSNet/homography.py:78-79, CVP_MVSNet/models/modules.py:71,91,...).
"""
import sys
import types

import torch

REFERENCE_ROOT = "/root/reference"


def import_reference(root=REFERENCE_ROOT):
    if root not in sys.path:
        sys.path.insert(0, root)
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        mpl, cm, plt = (types.ModuleType(n) for n in ("matplotlib", "matplotlib.cm", "matplotlib.pyplot"))
        cm.get_cmap = lambda name: None
        mpl.cm, mpl.pyplot = cm, plt
        sys.modules.update({"matplotlib": mpl, "matplotlib.cm": cm, "matplotlib.pyplot": plt})
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.cuda.empty_cache = lambda: None
    from models.MVSNet.model import MVSNet
    from models.MVSNet import module as mvs_module
    from models.VisMVSNet.frontend import Frontend as VisFrontend
    from models.VisMVSNet import homography as vis_homography
    from models.VisMVSNet import nn_utils as vis_nn
    from models.VisMVSNet import preproc as vis_preproc
    from models.CVP_MVSNet.frontend import Frontend as CVPFrontend
    from models.CVP_MVSNet.models import modules as cvp_modules
    return types.SimpleNamespace(MVSNet=MVSNet, mvs_module=mvs_module, VisFrontend=VisFrontend,
                                 vis_homography=vis_homography, vis_nn=vis_nn, vis_preproc=vis_preproc,
                                 CVPFrontend=CVPFrontend, cvp_modules=cvp_modules)
