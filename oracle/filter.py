"""CPU restatement (numpy, fp32) of the reference's geometric-consistency filter -- TEST INFRASTRUCTURE ONLY.

Follows evaluation/filtering.py:59-84 of fdarmon/wild_deep_mvs and the helpers it calls in utils/utils_3D.py
(unproject :116-132, project_all :64-73, normalize :243-273, unproj_all :162-178, project :98-113,
compute_triangulation_angles :296-311); `F.grid_sample(bilinear, zeros, align_corners=False)` is written out.
Pinned against masks written by the unmodified reference (tests/golden/geo_filter.npz, make_golden_filter.py).
"""
import numpy as np

f32 = np.float32


def _grid_sample_bilinear_zeros(img, gx, gy):
    """F.grid_sample(img[None,None], grid, mode='bilinear', padding_mode='zeros', align_corners=False) for one map."""
    H, W = img.shape
    ix = ((gx + f32(1)) * f32(W) - f32(1)) / f32(2)
    iy = ((gy + f32(1)) * f32(H) - f32(1)) / f32(2)
    x0 = np.floor(ix)
    y0 = np.floor(iy)
    out = np.zeros(ix.shape, f32)
    for dy in (0, 1):
        for dx in (0, 1):
            xi, yi = x0 + dx, y0 + dy
            wgt = (f32(1) - np.abs(ix - xi)) * (f32(1) - np.abs(iy - yi))
            ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H) & np.isfinite(ix) & np.isfinite(iy)
            xs = np.clip(np.nan_to_num(xi), 0, W - 1).astype(np.int64)
            ys = np.clip(np.nan_to_num(yi), 0, H - 1).astype(np.int64)
            out += np.where(ok, img[ys, xs] * wgt.astype(f32), f32(0)).astype(f32)
    return out


def geometric_filter(depth, src_depths, K, R, t, depth_threshold=0.01, max_reproj_error=1.0, min_tri_angle=1.0,
                     num_consistent=3):
    """depth [h,w]; src_depths list of [hi,wi]; K, R [V,3,3]; t [V,3,1] (view 0 = reference, x_cam = R X + t).
    Returns dict(mask_depth, mask_disp, geo_mask) of bool [h,w] (filtering.py:79-84) and the per-source float seams."""
    depth = np.asarray(depth, f32)
    K, R, t = (np.asarray(a, f32) for a in (K, R, t))
    h, w = depth.shape
    N = len(src_depths)
    ys, xs = np.meshgrid(np.arange(h, dtype=f32), np.arange(w, dtype=f32), indexing="ij")
    grid = np.stack([xs, ys], -1)                                                   # build_grid(h, w, dev, False): (x, y)
    hom = np.concatenate([grid, np.ones((h, w, 1), f32)], -1)
    # unproject (utils_3D.py:131): ((hom * D) @ inv(K0)^T - t0^T) @ R0
    pc = (((hom * depth[..., None]).reshape(-1, 3) @ np.linalg.inv(K[0]).T.astype(f32) - t[0].T) @ R[0]).astype(f32)
    dep_masks, disp_masks, geo_masks = [], [], []
    seams = []
    ray1 = pc + (R[0].T @ t[0]).T                                                    # :304
    n1 = np.maximum(np.linalg.norm(ray1, axis=1), f32(1e-12))
    for i in range(N):
        Ki, Ri, ti = K[1 + i], R[1 + i], t[1 + i]
        sd = np.asarray(src_depths[i], f32)
        hi, wi = sd.shape
        u = ((pc @ Ri.T + ti.T) @ Ki.T).astype(f32)                                  # project_all :70
        dsrc = u[:, 2]
        proj = u[:, :2] / np.maximum(dsrc, f32(1e-6))[:, None]
        gx = f32(2) * proj[:, 0] / f32(wi - 1) - f32(1)                              # normalize :267-268 (size - 1) ...
        gy = f32(2) * proj[:, 1] / f32(hi - 1) - f32(1)
        wd = _grid_sample_bilinear_zeros(sd, gx, gy)                                 # ... sampled with align_corners=False (:69)
        homs = np.concatenate([proj, np.ones((h * w, 1), f32)], -1)
        ps = (((homs * wd[:, None]) @ np.linalg.inv(Ki).T.astype(f32) - ti.T) @ Ri).astype(f32)   # unproj_all :178
        v = ((ps @ R[0].T + t[0].T) @ K[0].T).astype(f32)                            # project :104-107
        drep = v[:, 2] + f32(1e-6)
        rep = v[:, :2] / drep[:, None]
        err = rep - grid.reshape(-1, 2)
        valid_disp = np.linalg.norm(err, axis=1) < f32(max_reproj_error)             # :73
        d0 = depth.reshape(-1)
        mask_depth = (np.abs(drep - d0) < np.maximum(drep, d0) * f32(depth_threshold)) & (drep > 0) & (dsrc > 0)   # :75-76
        ray2 = pc + (Ri.T @ ti).T                                                    # :305
        cos = np.clip((ray1 * ray2).sum(1) / n1 / np.maximum(np.linalg.norm(ray2, axis=1), f32(1e-12)), -1, 1)
        tri = np.arccos(cos) / np.pi * 180 > min_tri_angle                           # :78, :311
        dep_masks.append(mask_depth)
        disp_masks.append(valid_disp)
        geo_masks.append(mask_depth & valid_disp & tri)
        seams.append({"proj_depth_in_src": dsrc.reshape(h, w), "warp_depth_in_src": wd.reshape(h, w),
                      "depth_reproj": drep.reshape(h, w), "reproj_error": np.linalg.norm(err, axis=1).reshape(h, w)})
    k = num_consistent - 1
    cnt = lambda ms: (np.sum(ms, axis=0) >= k).reshape(h, w)
    return {"mask_depth": cnt(dep_masks), "mask_disp": cnt(disp_masks), "geo_mask": cnt(geo_masks), "seams": seams}


def gathered_masks(ref_depth, gathered, proj_mat, ref_idx, geom_clamping=0.05, imgs=None):
    """The geometric half of `masked_photometricloss` (models/trainer.py:240-278 of the reference) in numpy fp32:
    `flows_from_single_depthmap` (utils/utils_3D.py:190-211: points3D = add_hom(add_hom(grid) * depth) @ inv(P_ref)^T,
    reprojected = points3D @ P_src^T, flow = xy / clamp(z, 1e-6)), `normalize` (:243-268: 2 f / (size - 1) - 1),
    `get_flow_from_depthmap` (trainer.py:209-219: z <= 0 -> -10, clamp +-10), the inside mask (:258), the bilinear /
    zeros / align_corners=False samples of the gathered depth maps and images (:261-262) and the re-projection mask
    (:264-267).  ref_depth [b,h,w], gathered [b,N,h,w], proj_mat [b,N,4,4], imgs [b,N,c,h,w] or None.
    Returns dict(masks, inside [b,N-1,h,w] bool; flows [b,N-1,h,w,2]; depth_src, warped_depth, reproj_diff [b,N-1,h,w];
    warped [b,N-1,c,h,w] or None).  Pinned by tests/golden/gathered_masks.npz (the unmodified reference function)."""
    ref_depth = np.asarray(ref_depth, f32)
    gathered = np.asarray(gathered, f32)
    proj_mat = np.asarray(proj_mat, f32)
    b, N, h, w = gathered.shape
    src_idx = [i for i in range(N) if i != ref_idx]
    ys, xs = np.meshgrid(np.arange(h, dtype=f32), np.arange(w, dtype=f32), indexing="ij")
    S = N - 1
    flows = np.zeros((b, S, h, w, 2), f32)
    depth_src = np.zeros((b, S, h, w), f32)
    warped_depth = np.zeros((b, S, h, w), f32)
    warped = None if imgs is None else np.zeros((b, S, imgs.shape[2], h, w), f32)
    for bi in range(b):
        inv = np.linalg.inv(proj_mat[bi, ref_idx].astype(np.float64)).astype(f32)
        d = ref_depth[bi]
        hom = np.stack([xs * d, ys * d, d, np.ones_like(d)], -1).astype(f32)           # [h,w,4]
        X = (hom @ inv.T).astype(f32)
        for j, s in enumerate(src_idx):
            q = (X @ proj_mat[bi, s].T).astype(f32)
            z = q[..., 2]
            zc = np.maximum(z, f32(1e-6))
            gx = (f32(2) * (q[..., 0] / zc) / f32(w - 1) - f32(1)).astype(f32)
            gy = (f32(2) * (q[..., 1] / zc) / f32(h - 1) - f32(1)).astype(f32)
            behind = z <= 0
            gx = np.clip(np.where(behind, f32(-10), gx), f32(-10), f32(10)).astype(f32)
            gy = np.clip(np.where(behind, f32(-10), gy), f32(-10), f32(10)).astype(f32)
            flows[bi, j, ..., 0], flows[bi, j, ..., 1] = gx, gy
            depth_src[bi, j] = z
            warped_depth[bi, j] = _grid_sample_bilinear_zeros(gathered[bi, s], gx, gy)
            if imgs is not None:
                for c in range(imgs.shape[2]):
                    warped[bi, j, c] = _grid_sample_bilinear_zeros(np.asarray(imgs[bi, s, c], f32), gx, gy)
    inside = (flows < 1).all(-1) & (flows > -1).all(-1)
    reproj_diff = (np.abs(depth_src - warped_depth) / np.maximum(warped_depth, f32(1e-8))).astype(f32)
    masks = inside & (reproj_diff < f32(geom_clamping))
    return {"masks": masks, "inside": inside, "flows": flows, "depth_src": depth_src, "warped_depth": warped_depth,
            "reproj_diff": reproj_diff, "warped": warped}
