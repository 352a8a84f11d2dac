"""Golden tensors for the 2-D feature extractor of MVSNet (models/MVSNet/model.py:21-41).

Run in the BUILD container only (needs /root/reference):   python tests/golden/make_golden_features.py

Imports the unmodified reference (oracle/ref_import.py), runs FeatureNet on CPU on a seeded image batch with
non-trivial BN statistics, and stores the image, the weights and the output features (tests/golden/featurenet.npz).
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


if __name__ == "__main__":
    main()
