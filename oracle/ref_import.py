"""Import the UNMODIFIED reference (fdarmon/wild_deep_mvs) on a CPU-only host.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (tests/, bench.py's reference legs).  In the build container the
reference is imported from /root/reference (tests/golden/make_golden*.py produce the golden tensors that pin the
oracle); on the GPU box, where that tree does not exist, from oracle/_ref/ref_hotpath.zip -- the same files, byte
for byte, packed by oracle/make_ref.py (git-ignored, travels with the gpurun snapshot).  The product never imports
this module.

Two shims: the reference imports matplotlib at module import time
(utils/utils_3D.py:22-23, models/VisMVSNet/model_cas.py:8) and hard-codes .cuda() on the hot
path (VisMVSNet/homography.py:78-79, CVP_MVSNet/models/modules.py:71,91,...).
"""
import os
import sys
import types

import torch

REFERENCE_ROOT = "/root/reference"
ARCHIVE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "ref_hotpath.zip")


def reference_location():
    """Where the unmodified reference can be imported from here: the tree, the packed archive, or None."""
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "models")):
        return REFERENCE_ROOT
    if os.path.exists(ARCHIVE):
        return ARCHIVE
    return None


def import_reference(root=None, cuda_shim=None):
    """Import the reference's model classes.  `cuda_shim`: make Tensor.cuda() a no-op (the reference hard-codes .cuda() on
    the Vis / CVP paths); default = only when no CUDA device exists."""
    root = root or reference_location()
    if root is None:
        raise ImportError("reference not available: neither %s nor %s exists (run python oracle/make_ref.py in the build "
                          "container)" % (REFERENCE_ROOT, ARCHIVE))
    if root not in sys.path:
        sys.path.insert(0, root)
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        mpl, cm, plt = (types.ModuleType(n) for n in ("matplotlib", "matplotlib.cm", "matplotlib.pyplot"))
        cm.get_cmap = lambda name: None
        mpl.cm, mpl.pyplot = cm, plt
        sys.modules.update({"matplotlib": mpl, "matplotlib.cm": cm, "matplotlib.pyplot": plt})
    if cuda_shim is None:
        cuda_shim = not torch.cuda.is_available()
    if cuda_shim:
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
    from utils.utils_3D import build_proj_matrices
    return types.SimpleNamespace(root=root, build_proj_matrices=build_proj_matrices, MVSNet=MVSNet, mvs_module=mvs_module, VisFrontend=VisFrontend,
                                 vis_homography=vis_homography, vis_nn=vis_nn, vis_preproc=vis_preproc,
                                 CVPFrontend=CVPFrontend, cvp_modules=cvp_modules)
