"""Golden outputs for the consumer of the all-gathered depth maps (`masked_photometricloss`,
models/trainer.py:240-278 of the reference).

Run in the BUILD container only (needs /root/reference):   python tests/golden/make_golden_gathered.py

Executes the UNMODIFIED method `models.trainer.Trainer.masked_photometricloss` (with the unmodified
`get_flow_from_depthmap` it calls) on a synthetic scene and stores its inputs, the masks it returns and the warped images
it records (tests/golden/gathered_masks.npz).  Three things the method takes from its environment are supplied by the
script, none of them on the path under test: `dist.all_gather` / `dist.get_rank` (single process here: the "gathered"
maps are handed over, the rank is the reference view's index), `self.ssim` (the SSIM term of the loss is out of scope:
a stand-in that returns zeros) and `self.args.geom_clamping` (the reference's default 0.05, train.py:278).
Scene: a slanted plane seen by 4 pinhole cameras; every view's depth map is rendered analytically, the rank's own map gets a
noisy band and the gathered map of one source a corrupted block (so the re-projection test fails there), one camera looks
away far enough for part of the image to fall outside.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.ref_import import import_reference  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def plane_depth(K, R, t, h, w, n, c):
    ys, xs = np.meshgrid(np.arange(h, dtype=np.float64), np.arange(w, dtype=np.float64), indexing="ij")
    pix = np.stack([xs, ys, np.ones_like(xs)], -1) @ np.linalg.inv(K).T
    d = pix @ R
    o = -R.T @ t.reshape(3)
    return ((c - n @ o) / (d @ n)).astype(np.float32)


def scene(seed=0):
    rng = np.random.default_rng(seed)
    b, N, h, w = 2, 4, 40, 56
    K = np.zeros((b, N, 3, 3)); R = np.zeros((b, N, 3, 3)); t = np.zeros((b, N, 3, 1))
    gathered = np.zeros((b, N, h, w), np.float32)
    base = [0.0, -20.0, 16.0, 4.0]
    for bi in range(b):
        n, c = np.array([0.12 + 0.05 * bi, -0.08, 1.0]), 620.0 + 40 * bi
        for v in range(N):
            K[bi, v] = [[430.0, 0, w / 2.0], [0, 425.0, h / 2.0], [0, 0, 1]]
            a = 0.02 * v * (1 if v % 2 else -1) + (0.01 if v == 3 else 0.0)      # view 3 looks away: part of the image falls outside
            R[bi, v] = [[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]
            t[bi, v] = [[base[v]], [1.5 * v], [0.5 * v]]
            gathered[bi, v] = plane_depth(K[bi, v], R[bi, v], t[bi, v], h, w, n, c)
    gathered[:, 1, 10:18, 20:34] *= 1.2                                  # a corrupted block in one source's map
    gathered += (rng.standard_normal(gathered.shape) * 2.0).astype(np.float32)   # noise of about a tenth of the threshold
    imgs = rng.random((b, N, 3, h, w)).astype(np.float32)
    return K, R, t, gathered, imgs


def main():
    ref = import_reference(cuda_shim=True)
    import torch.distributed as dist
    sys.modules.setdefault("tensorboardX", types.ModuleType("tensorboardX"))
    from models.trainer import Trainer
    from utils.utils_3D import build_proj_matrices
    K, R, t, gathered, imgs = scene()
    i_ref = 1
    proj = build_proj_matrices(torch.from_numpy(K).float(), torch.from_numpy(R).float(), torch.from_numpy(t).float())
    g = torch.from_numpy(gathered)
    depth_est = g[:, i_ref].clone()
    depth_est[:, 25:30] += torch.linspace(0, 60, 5).view(1, 5, 1)        # this rank's own estimate: a band that drifts off the surface

    def all_gather(lst, tensor):                                          # what NCCL would deliver: the maps of all views
        for v in range(len(lst)):
            lst[v] = g[:, v].clone() if v != i_ref else tensor.clone()
    real = (dist.all_gather, dist.get_rank)
    dist.all_gather, dist.get_rank = all_gather, (lambda: i_ref)
    try:
        self = types.SimpleNamespace(args=types.SimpleNamespace(geom_clamping=0.05), ims={},
                                     ssim=lambda a, c: torch.zeros_like(a))
        self.get_flow_from_depthmap = types.MethodType(Trainer.get_flow_from_depthmap, self)
        with torch.no_grad():
            ssims, masks = Trainer.masked_photometricloss(self, torch.from_numpy(imgs), depth_est, proj)
            flows, depth_src = self.get_flow_from_depthmap(depth_est, proj, imgs.shape[-2:], i_ref)
    finally:
        dist.all_gather, dist.get_rank = real
    src_idx = [v for v in range(g.shape[1]) if v != i_ref]
    masked = np.stack([self.ims["warped_ref_%dsrc_%d_masked" % (i_ref, s)].numpy() for s in src_idx], 1)
    gathered_in = gathered.copy()
    gathered_in[:, i_ref] = depth_est.numpy()
    np.savez_compressed(os.path.join(OUT, "gathered_masks.npz"), proj=proj.numpy(), gathered=gathered_in, ref_depth=depth_est.numpy(),
                        imgs=imgs, ref_idx=np.int32(i_ref), geom_clamping=np.float32(0.05), masks=masks.numpy(),
                        flows=flows.numpy(), depth_src=depth_src.numpy(), warped_masked=masked)
    print("masks kept %.3f of the pixels; per source %s" % (masks.float().mean(), masks.float().mean(dim=(0, 2, 3)).tolist()))


if __name__ == "__main__":
    main()
