"""Golden files for the on-disk formats, written and read by the UNMODIFIED reference functions.

Run in the BUILD container only (needs /root/reference):   python tests/golden/make_golden_formats.py

Writes tests/golden/formats/: the bytes `evaluation.fusibile.write_gipuma_dmb / write_gipuma_cam / fake_gipuma_normal`
and `utils.colmap_utils.write_array` produce for small seeded arrays, a `_out.npz` written the way
`evaluation/run_depthmaps.py:66-67` does, PFM files (the reference has no PFM writer: they are written here, then decoded
by the reference's `data.MVSDataset.read_pfm`), and `expected.npz` with the source arrays plus what the reference's own
readers return for each file.
"""
import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_import import import_reference, REFERENCE_ROOT  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "formats")


def load_file(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REFERENCE_ROOT, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    os.makedirs(OUT, exist_ok=True)
    # evaluation/pipeline_utils.py imports the dataset modules (h5py): not needed for the format functions
    for name in ("data", "data.dtu_yao_eval", "data.yfcc_scene"):
        sys.modules.setdefault(name, types.ModuleType(name))
    import_reference()
    from evaluation import fusibile
    from utils import colmap_utils
    mvsdataset = load_file("ref_mvsdataset", "data/MVSDataset.py")

    rng = np.random.default_rng(0)
    depth = (rng.uniform(400, 900, (5, 7))).astype(np.float32)
    depth[1, 2] = 0.0
    depth[4, 6] = 0.0
    conf = rng.uniform(0, 1, (5, 7)).astype(np.float32)
    conf3 = rng.uniform(0, 1, (3, 5, 7)).astype(np.float32)
    rgb = rng.uniform(0, 1, (5, 7, 3)).astype(np.float32)
    P = (rng.normal(0, 100, (3, 4))).astype(np.float32).astype(np.float64)
    P[2] = [1e-3, -2.5e-4, 0.999999, 1234.5678]
    K = np.array([[723.0825, 0, 160.0], [0, 720.795, 128.0], [0, 0, 1]], np.float32)
    a = 0.07
    R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]], np.float32)
    t = np.array([[-40.0], [1.5], [0.25]], np.float32)

    exp = {"depth": depth, "conf": conf, "conf3": conf3, "rgb": rgb, "P": P, "K": K, "R": R, "t": t}

    # Gipuma
    fusibile.write_gipuma_dmb(os.path.join(OUT, "disp.dmb"), depth)
    fusibile.fake_gipuma_normal(os.path.join(OUT, "disp.dmb"), os.path.join(OUT, "normals.dmb"))
    fusibile.write_gipuma_dmb(os.path.join(OUT, "rgb.dmb"), rgb)
    fusibile.write_gipuma_cam(P, os.path.join(OUT, "cam.P"))
    exp["dmb_disp_read"] = fusibile.read_gipuma_dmb(os.path.join(OUT, "disp.dmb"))
    exp["dmb_normals_read"] = fusibile.read_gipuma_dmb(os.path.join(OUT, "normals.dmb"))
    exp["dmb_rgb_read"] = fusibile.read_gipuma_dmb(os.path.join(OUT, "rgb.dmb"))
    # the projection matrix mvsnet_to_gipuma writes (fusibile.py:112-124), downscale 2
    import torch
    from utils.utils_3D import build_proj_matrices
    pm = build_proj_matrices(torch.from_numpy(K)[None], torch.from_numpy(R)[None], torch.from_numpy(t)[None])[0]
    pm[:2] /= 2
    exp["gipuma_P_down2"] = pm[:3].double().numpy()
    fusibile.write_gipuma_cam(exp["gipuma_P_down2"], os.path.join(OUT, "cam_down2.P"))

    # COLMAP
    colmap_utils.write_array(depth, os.path.join(OUT, "depth.geometric.bin"))
    colmap_utils.write_array(rgb, os.path.join(OUT, "normal.geometric.bin"))
    exp["colmap_depth_read"] = colmap_utils.read_array(os.path.join(OUT, "depth.geometric.bin"))
    exp["colmap_rgb_read"] = colmap_utils.read_array(os.path.join(OUT, "normal.geometric.bin"))

    # npz exactly as run_depthmaps.py:66-67
    np.savez_compressed(os.path.join(OUT, "view0_out.npz"), probability=conf3, depthmap=depth)

    # PFM: little- and big-endian grey, little-endian colour; decoded by the reference's reader
    def pfm(path, img, endian):
        with open(path, "wb") as f:
            f.write(b"PF\n" if img.ndim == 3 else b"Pf\n")
            f.write(("%d %d\n" % (img.shape[1], img.shape[0])).encode())
            f.write(b"-1.000000\n" if endian == "<" else b"2.500000\n")
            f.write(np.flipud(img).astype(endian + "f4").tobytes())
    pfm(os.path.join(OUT, "depth_le.pfm"), depth, "<")
    pfm(os.path.join(OUT, "depth_be.pfm"), depth, ">")
    pfm(os.path.join(OUT, "rgb_le.pfm"), rgb, "<")
    for n in ("depth_le", "depth_be", "rgb_le"):
        d, s = mvsdataset.read_pfm(os.path.join(OUT, n + ".pfm"))
        exp["pfm_%s_read" % n] = np.ascontiguousarray(d).astype(np.float32)
        exp["pfm_%s_scale" % n] = np.float64(s)

    np.savez_compressed(os.path.join(OUT, "expected.npz"), **exp)
    print(sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
