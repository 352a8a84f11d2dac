"""Golden tensors for the 2-D feature extractors of MVSNet and Vis-MVSNet (models/MVSNet/model.py:21-41,
models/VisMVSNet/model_cas.py:18-35).

Run in the BUILD container only (needs /root/reference):   python tests/golden/make_golden_features.py

Imports the unmodified reference (oracle/ref_import.py), runs FeatureNet on CPU on a seeded image batch with
non-trivial BN statistics, and stores the image, the weights and the output features (tests/golden/featurenet.npz,
tests/golden/featext.npz).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.ref_import import import_reference  # noqa: E402
from wild_deep_mvs_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    ref = import_reference()
    torch.manual_seed(0)
    net = ref.MVSNet("variance").eval()
    synth.randomize_norm_stats(net, seed=7)
    g = torch.Generator().manual_seed(11)
    img = torch.rand(2, 3, 44, 60, generator=g)            # sizes that are not multiples of the CUDA tile
    with torch.no_grad():
        feat = net.feature(img)
    out = {"img": img.numpy(), "feat": feat.numpy()}
    out.update({"feature." + k: v.detach().cpu().numpy() for k, v in net.feature.state_dict().items()})
    np.savez(os.path.join(OUT, "featurenet.npz"), **out)
    print("featurenet", feat.shape, float(feat.abs().max()), float(feat.std()))

    # Vis-MVSNet FeatExt (models/VisMVSNet/model_cas.py:18-35): three scales of 32-channel features
    torch.manual_seed(1)
    vis = ref.VisFrontend().eval()
    synth.randomize_norm_stats(vis, seed=8)
    img = torch.rand(2, 3, 40, 56, generator=g)
    with torch.no_grad():
        f1, f2, f3 = vis.model.feat_ext(img)
    out = {"img": img.numpy(), "feat_s1": f1.numpy(), "feat_s2": f2.numpy(), "feat_s3": f3.numpy()}
    out.update({"model.feat_ext." + k: v.detach().cpu().numpy() for k, v in vis.model.feat_ext.state_dict().items()})
    np.savez(os.path.join(OUT, "featext.npz"), **out)
    print("featext", f1.shape, f2.shape, f3.shape, float(f3.abs().max()), float(f3.std()))


if __name__ == "__main__":
    main()
