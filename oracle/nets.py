"""Oracle networks: the three regularisers and the features->depth forwards, composed from
the C primitives in oracle/mvs_oracle.c and a dict of numpy weights keyed like the
reference's state_dict.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

All tensors are single-sample numpy arrays in the reference's layouts ([C,D,H,W] etc.).
"""
import numpy as np

import oracle as orc

EPS = 1e-5  # nn.BatchNorm default


def _bn(sd, p, x, relu):
    return orc.bn_relu(x, sd[p + ".weight"], sd[p + ".bias"], sd[p + ".running_mean"], sd[p + ".running_var"],
                       EPS, relu)


# ---------------------------------------------------------------------------------------------
# MVSNet CostRegNet -- models/MVSNet/model.py:43-84
# ---------------------------------------------------------------------------------------------
def mvsnet_costreg(sd, x, prefix="cost_regularization."):
    def cbr(name, x, stride=1):  # ConvBnReLU3D, module.py:41-48
        y = orc.conv3d(x, sd[prefix + name + ".conv.weight"], None, stride)
        return _bn(sd, prefix + name + ".bn", y, True)

    def dbr(name, x):  # ConvTranspose3d + BN + ReLU, model.py:58-71
        y = orc.deconv3d(x, sd[prefix + name + ".0.weight"], None, 2, 1, 1)
        return _bn(sd, prefix + name + ".1", y, True)

    c0 = cbr("conv0", x)
    c2 = cbr("conv2", cbr("conv1", c0, 2))
    c4 = cbr("conv4", cbr("conv3", c2, 2))
    y = cbr("conv6", cbr("conv5", c4, 2))
    y = c4 + dbr("conv7", y)      # model.py:79 (no ReLU after the add)
    y = c2 + dbr("conv9", y)
    y = c0 + dbr("conv11", y)
    return orc.conv3d(y, sd[prefix + "prob.weight"], sd[prefix + "prob.bias"], 1)  # [1,D,H,W]


def mvsnet_from_features(sd, feats, rel_projs, depth_values, aggregation="variance", temp=1.0):
    """feats: list of [C,H,W] (index 0 = reference); rel_projs: list of (rot, trans) per source;
    models/MVSNet/model.py:202-215."""
    H, W = feats[0].shape[1:]
    warped = [orc.homo_warp_mvs(f, r, t, depth_values, (H, W)) for f, (r, t) in zip(feats[1:], rel_projs)]
    cost = orc.variance(feats[0], warped, 0) if aggregation == "variance" else orc.softmin(feats[0], warped, temp)
    reg = mvsnet_costreg(sd, cost)[0]
    out = orc.softmax_regress(reg, 0, depth_values, conf_mode=1, want_prob=True)
    return {"cost_volume": cost, "cost_reg": reg, "prob": out["prob"], "depth": out["depth"], "conf": out["conf"]}


# ---------------------------------------------------------------------------------------------
# Vis-MVSNet Reg / RegFuse UNet -- models/VisMVSNet/model_cas.py:38-74, nn_utils.py:123-278
# ---------------------------------------------------------------------------------------------
def mvsnet_featurenet(sd, img, prefix="feature."):
    """FeatureNet, models/MVSNet/model.py:21-41: seven ConvBnReLU (module.py:9-17; k5 s2 at layers 2 and 5) and a
    biased 3x3 conv.  img [3,H,W] -> [32,H/4,W/4]."""
    spec = [(3, 1), (3, 1), (5, 2), (3, 1), (3, 1), (5, 2), (3, 1)]
    x = img
    for i, (k, s) in enumerate(spec):
        w = sd[prefix + "conv%d.conv.weight" % i]
        x = orc.conv3d(x[:, None], w[:, :, None], None, s, (0, k // 2, k // 2))[:, 0]
        # a stride also applies to the dummy depth axis: D = 1, kd = 1, pad 0 -> one output plane
        x = _bn(sd, prefix + "conv%d.bn" % i, x, True)
    w = sd[prefix + "feature.weight"]
    return orc.conv3d(x[:, None], w[:, :, None], sd[prefix + "feature.bias"], 1, (0, 1, 1))[:, 0]


def _c2d(x, w, stride=1, pad=None):
    """nn.Conv2d on [C,H,W] through the 3-D oracle conv (a one-plane volume, kd = 1)."""
    k = w.shape[2]
    pad = k // 2 if pad is None else pad
    return orc.conv3d(x[:, None], w[:, :, None], None, stride, (0, pad, pad))[:, 0]


def _basic_block2d(sd, p, x, stride):  # nn_utils.py:123-171 with dim = 2
    y = _bn(sd, p + ".bn1", _c2d(x, sd[p + ".conv1.weight"], stride), True)
    y = _bn(sd, p + ".bn2", _c2d(y, sd[p + ".conv2.weight"]), False)
    if (p + ".downsample.0.weight") in sd:
        r = _bn(sd, p + ".downsample.1", _c2d(x, sd[p + ".downsample.0.weight"], stride, 0), False)
    else:
        r = x
    return np.maximum(y + r, 0.0).astype(np.float32)


def _deconv2d(x, w):
    """nn.ConvTranspose2d(k=3, stride=2, padding=1, output_padding=1) on [C,H,W]: the 3-D oracle deconv on a one-plane
    volume with the 2-D kernel as the middle depth tap; output plane 0 is the 2-D result."""
    w3 = np.zeros(w.shape[:2] + (3, 3, 3), np.float32)
    w3[:, :, 1] = w
    return orc.deconv3d(x[:, None], w3, None, 2, 1, 1)[:, 0]


def vis_featext(sd, img, prefix="model.feat_ext."):
    """FeatExt, models/VisMVSNet/model_cas.py:18-35 over UNet(16, 2, 1, 2, [], [32, 64, 128], [], '2d', 2)
    (nn_utils.py:194-278).  img [3,H,W] -> features at 1/8, 1/4, 1/2 resolution, each [32,h,w]."""
    x = _bn(sd, prefix + "init_conv.1", _c2d(img, sd[prefix + "init_conv.0.weight"], 2), True)
    u = prefix + "unet."
    enc = []
    for name, stride in (("2d2_0", 1), ("2d4_1", 2), ("2d8_2", 2)):
        x = _basic_block2d(sd, u + "enc_blocks.%s.0" % name, x, stride)
        x = _basic_block2d(sd, u + "enc_blocks.%s.1" % name, x, 1)
        enc.append(x)
    outs = [x]
    for i, name in enumerate(("2d16_3", "2d8_4")):
        d = u + "dec_blocks.%s." % name
        x = _deconv2d(x, sd[d + "0.weight"])
        x = _c2d(np.concatenate([x, enc[-2 - i]], 0), sd[d + "1.weight"])
        x = _basic_block2d(sd, d + "2.0", x, 1)
        outs.append(x)
    return [_c2d(o, sd[prefix + "final_conv_%d.weight" % (k + 1)]) for k, o in enumerate(outs)]


def _basic_block(sd, p, x, stride):  # nn_utils.py:123-171
    y = orc.conv3d(x, sd[p + ".conv1.weight"], None, stride)
    y = _bn(sd, p + ".bn1", y, True)
    y = orc.conv3d(y, sd[p + ".conv2.weight"], None, 1)
    y = _bn(sd, p + ".bn2", y, False)
    if (p + ".downsample.0.weight") in sd:
        r = orc.conv3d(x, sd[p + ".downsample.0.weight"], None, stride, 0)
        r = _bn(sd, p + ".downsample.1", r, False)
    else:
        r = x
    return np.maximum(y + r, 0.0).astype(np.float32)


def vis_unet(sd, prefix, tag, x):
    """UNet(8,1,0,4,[],[8,16],[],tag,dim=3).  Key names: enc_blocks.<tag>4_0.0, enc_blocks.<tag>8_1.0,
    dec_blocks.<tag>16_2.{0,1}."""
    e0 = _basic_block(sd, prefix + "enc_blocks.%s4_0.0" % tag, x, 1)
    e1 = _basic_block(sd, prefix + "enc_blocks.%s8_1.0" % tag, e0, 2)
    up = orc.deconv3d(e1, sd[prefix + "dec_blocks.%s16_2.0.weight" % tag], None, 2, 1, 1)
    cat = np.concatenate([up, e0], 0)  # nn_utils.py:268
    return orc.conv3d(cat, sd[prefix + "dec_blocks.%s16_2.1.weight" % tag], None, 1)


def _conv2d(x, w, relu_bn=None, sd=None):
    y = orc.conv3d(x[:, None], w[:, :, None], None, 1, (0, 1, 1))[:, 0]
    if relu_bn:
        y = _bn(sd, relu_bn, y, True)
    return y


def vis_uncert_net(sd, p, ent):  # model_cas.py:77-98, ent [1,H,W]
    y = _conv2d(ent, sd[p + "conv1.0.weight"], p + "conv1.1", sd)
    y = _conv2d(y, sd[p + "conv2.0.weight"], p + "conv2.1", sd)
    y = y + ent
    return _conv2d(y, sd[p + "head_convs.0.weight"])  # [1,H,W]


def scale_cam(cam, s):  # preproc.py:63-92 (torch branch)
    c = cam.copy()
    c[1, 0, 0] *= s
    c[1, 1, 1] *= s
    c[1, 0, 2] *= s
    c[1, 1, 2] *= s
    return c


def vis_stage(sd, prefix, ref_feat, ref_cam, src_feats, src_cams, D, depth_start, depth_interval, s_scale):
    """SingleStage.forward (mode='soft'), models/VisMVSNet/model_cas.py:303-420.
    depth_start: [1] or [H,W]; returns est_depth [H,W], prob_map [H,W], pair list, seams."""
    H, W = ref_feat.shape[1:]
    rc = scale_cam(ref_cam, 1.0 / s_scale)
    interms, uncerts, pairs, seams = [], [], [], {}
    for i, (sf, sc) in enumerate(zip(src_feats, src_cams)):
        warped = orc.vis_warp(sf, rc, scale_cam(sc, 1.0 / s_scale), depth_start, depth_interval, D, (H, W))
        cost = orc.groupcorr(ref_feat, warped, 8)
        interm = vis_unet(sd, prefix + "reg.unet.", "reg1", cost)
        score = orc.conv3d(interm, sd[prefix + "reg_pair.final_conv.weight"], None, 1)[0]
        r = orc.softmax_regress(score, 2, depth_start, depth_interval, 0, want_entropy=True)
        u = vis_uncert_net(sd, prefix + "uncert_net.", r["entropy"][None])
        interms.append(interm)
        uncerts.append(u[0])
        pairs.append((r["depth"], u[0]))
        if i == 0:
            seams.update(warped0=warped, groupcorr0=cost, pair_score0=score)
    fused = orc.vis_fuse(interms, uncerts)
    fu = vis_unet(sd, prefix + "reg_fuse.unet.", "reg2", fused)
    score = orc.conv3d(fu, sd[prefix + "reg_fuse.final_conv.weight"], None, 1)[0]
    r = orc.softmax_regress(score, 2, depth_start, depth_interval, 2)
    seams.update(interms=interms, fuse_unet=fu, fuse_score=score)
    return r["depth"], r["conf"], pairs, seams


def upsample_bilinear(x, out_hw):
    """F.interpolate(mode='bilinear', align_corners=False) on [H,W] (frontend.py:74-91)."""
    H, W = x.shape
    Ho, Wo = out_hw

    def axis(n_in, n_out):
        src = (np.arange(n_out, dtype=np.float32) + 0.5) * np.float32(n_in / n_out) - 0.5
        src = np.maximum(src, 0.0).astype(np.float32)
        i0 = np.minimum(np.floor(src).astype(np.int64), n_in - 1)
        i1 = np.minimum(i0 + 1, n_in - 1)
        l1 = (src - i0).astype(np.float32)
        return i0, i1, l1

    y0, y1, ly = axis(H, Ho)
    x0, x1, lx = axis(W, Wo)
    top = x[y0][:, x0] * (1 - lx)[None] + x[y0][:, x1] * lx[None]
    bot = x[y1][:, x0] * (1 - lx)[None] + x[y1][:, x1] * lx[None]
    return (top * (1 - ly)[:, None] + bot * ly[:, None]).astype(np.float32)


def vis_from_features(sd, feats, ref_cam, src_cams, depth_min, depth_max, depth_nums, interval_scales,
                      self_interval_scales=None):
    """Frontend.forward cascade, models/VisMVSNet/frontend.py:26-109.
    feats[v][k]: view v (0=ref), scale k (0: 1/8, 1: 1/4, 2: 1/2), each [32,h,w]."""
    sis = self_interval_scales or interval_scales  # frontend.py:76-78 uses self.interval_scales
    interval = np.float32((depth_max - depth_min) / np.float32(128))
    ests, probs, pairs, seams = [], [], [], []
    start = np.array([depth_min], np.float32)
    for k, (name, s_scale) in enumerate((("stage1", 8), ("stage2", 4), ("stage3", 2))):
        rf = feats[0][k]
        if k > 0:
            up = upsample_bilinear(ests[-1], rf.shape[1:])
            start = (up - np.float32(depth_nums[k]) * interval * np.float32(sis[k]) / np.float32(2)).astype(np.float32)
        d, p, pr, sm = vis_stage(sd, "model.%s." % name, rf, ref_cam, [f[k] for f in feats[1:]], src_cams,
                                 depth_nums[k], start, np.float32(interval * np.float32(interval_scales[k])), s_scale)
        ests.append(d)
        probs.append(p)
        pairs.append(pr)
        seams.append(sm)
    H3, W3 = ests[2].shape
    conf = np.stack([upsample_bilinear(probs[0], (H3, W3)), upsample_bilinear(probs[1], (H3, W3)), probs[2]], 0)
    return {"depth": ests[2], "depth_est_list": ests[::-1], "depth_pair_list": pairs[::-1], "conf": conf,
            "seams": seams}


# ---------------------------------------------------------------------------------------------
# CVP-MVSNet CostRegNet -- models/CVP_MVSNet/models/net.py:50-85
# ---------------------------------------------------------------------------------------------
def cvp_costreg(sd, x, prefix="model.cost_reg_refine."):
    def cbr(name, x, stride=1):
        y = orc.conv3d(x, sd[prefix + name + ".conv.weight"], None, stride)
        return _bn(sd, prefix + name + ".bn", y, True)

    c0 = cbr("conv0a", cbr("conv0", x))
    c2 = cbr("conv2a", cbr("conv2", cbr("conv1", c0, 2)))
    c4 = cbr("conv4a", cbr("conv4", cbr("conv3", c2)))
    y = orc.deconv3d(c4, sd[prefix + "conv5.0.weight"], None, 1, 1, 0)  # stride-1 transposed conv
    c5 = c2 + _bn(sd, prefix + "conv5.1", y, True)
    y = orc.deconv3d(c5, sd[prefix + "conv6.0.weight"], None, 2, 1, 1)
    c6 = c0 + _bn(sd, prefix + "conv6.1", y, True)
    return orc.conv3d(c6, sd[prefix + "prob0.weight"], sd[prefix + "prob0.bias"], 1)[0]  # [D,H,W]


def upsample_bicubic2x(x):
    """F.interpolate(scale_factor=2, mode='bicubic', align_corners=None) on [H,W] (CVP net.py:169-170).
    Keys' cubic convolution, A=-0.75, border-replicated taps (ATen upsample_bicubic2d)."""
    A = np.float32(-0.75)

    def coeffs(t):
        t = t.astype(np.float32)

        def c1(x):  # |x| <= 1
            return ((A + 2) * x - (A + 3)) * x * x + 1

        def c2(x):  # 1 < |x| < 2
            return ((A * x - 5 * A) * x + 8 * A) * x - 4 * A

        return np.stack([c2(t + 1), c1(t), c1(1 - t), c2(2 - t)], 0).astype(np.float32)

    def axis(n):
        o = np.arange(2 * n, dtype=np.float32)
        src = (o + np.float32(0.5)) * np.float32(0.5) - np.float32(0.5)
        i = np.floor(src)
        t = src - i
        idx = np.stack([np.clip(i.astype(np.int64) + k, 0, n - 1) for k in (-1, 0, 1, 2)], 0)
        return idx, coeffs(t)

    H, W = x.shape
    iy, cy = axis(H)
    ix, cx = axis(W)
    rows = sum(x[:, ix[k]] * cx[k][None] for k in range(4))           # [H, 2W]
    out = sum(rows[iy[k]] * cy[k][:, None] for k in range(4))         # [2H, 2W]
    return out.astype(np.float32)


def cvp_depth_hypos(depth_up, K_ref, K_src0, E_ref, E_src0, depth_min, depth_max, d=4):
    """calDepthHypo (eval path), models/CVP_MVSNet/models/modules.py:131-226, fp64.
    depth_up [H,W]; intrinsics 3x3 (level-conditioned); extrinsics 4x4. Returns [8,H,W] fp32."""
    H, W = depth_up.shape
    Kr, Ks, Er, Es = (np.asarray(a, np.float64) for a in (K_ref, K_src0, E_ref, E_src0))
    xx, yy = np.meshgrid(np.arange(W), np.arange(H), indexing="ij")  # :151 (x-major ordering)
    X = np.stack([xx.reshape(-1), yy.reshape(-1), np.ones(H * W)], 0).astype(np.float64)
    D1 = depth_up.T.reshape(-1).astype(np.float64)  # :161 still in the INPUT dtype (fp32) before promotion
    D2 = (depth_up.T.reshape(-1) + np.float32(1)).astype(np.float64)

    def to_src(Dv):
        ray = np.linalg.inv(Kr) @ (X * Dv)
        P = np.linalg.inv(Er) @ np.concatenate([ray, np.ones((1, H * W))], 0)
        P = (Es @ P)[:3]
        P = Ks @ P
        dd = P[2].copy()
        return P / dd, dd

    X1, X1d = to_src(D1)
    X2, X2d = to_src(D2)
    dirv = X2 - X1
    nrm = np.linalg.norm(dirv, axis=0)
    dirv = dirv / np.maximum(nrm, 1e-8)
    X3 = X1 + dirv
    A = (Kr @ Er[:3, :3]) @ np.linalg.inv(Ks @ Es[:3, :3])
    tmp1 = X1d * (A @ X1)
    tmp2 = A @ X3
    M1 = np.stack([X.T, tmp2.T], 2)[:, 1:, :]  # [N,2,2]
    M2 = tmp1.T[:, 1:]                         # [N,2]
    det = M1[:, 0, 0] * M1[:, 1, 1] - M1[:, 0, 1] * M1[:, 1, 0]
    valid = (nrm > 1e-8) & (X1d > 1e-8) & (X2d > 1e-8) & (np.abs(det) > 1e-8)
    if valid.sum() > 0:
        ans = np.linalg.solve(M1[valid], M2[valid][:, :, None])
        delta = np.abs(ans[:, 0, 0])
        # torch.median returns the LOWER of the two middle values for even counts
        interval = np.sort(delta)[(delta.size - 1) // 2]
    else:
        interval = (depth_max - depth_min) / 128
    base = depth_up.astype(np.float32)
    # depth_hypos stays fp32 in the reference (in-place += of an fp64 map into an fp32 tensor)
    return np.stack([(base + np.float32(k * interval)).astype(np.float32) for k in range(-d, d)], 0)


def cvp_from_features(sd, fps, Ks, Es, depth_min, depth_max, nhyp0=96):
    """network.forward (eval), models/CVP_MVSNet/models/net.py:96-229.
    fps[v][level]: [16,h,w] with level 0 = finest; Ks[v]: 3x3 full-res intrinsics; Es[v]: 4x4."""
    nscale = len(fps[0])
    img_h = fps[0][0].shape[1]
    Kl = [[None] * nscale for _ in fps]
    for v in range(len(fps)):
        for l in range(nscale):
            ratio = np.float32(img_h / fps[v][l].shape[1])  # modules.py:31-46 (uses the HEIGHT ratio)
            k = np.asarray(Ks[v], np.float32).copy()
            k[:2, :] = k[:2, :] / ratio
            Kl[v][l] = k

    def rel(v, l):
        def proj(K, E):
            P = np.zeros((4, 4), np.float32)
            P[:3] = (K @ np.asarray(E, np.float32)[:3]).astype(np.float32)
            P[3, 3] = 1
            return P
        return orc.relative_proj(proj(Kl[v][l], Es[v]), proj(Kl[0][l], Es[0]))

    S = len(fps) - 1
    lc = nscale - 1
    hyp = (np.float32(depth_min) + np.arange(nhyp0, dtype=np.float32) *
           np.float32((np.float32(depth_max) - np.float32(depth_min)) / np.float32(nhyp0))).astype(np.float32)
    H, W = fps[0][lc].shape[1:]
    warped = [orc.homo_warp_mvs(fps[v][lc], *rel(v, lc), hyp, (H, W)) for v in range(1, S + 1)]
    cost = orc.variance(fps[0][lc], warped, 1)
    reg = cvp_costreg(sd, cost)
    r = orc.softmax_regress(reg, 0, hyp, conf_mode=1)
    depth = r["depth"]
    ests = [depth]
    seams = {"reg_in_coarse": cost, "reg_out_coarse": reg}
    for level in range(nscale - 2, -1, -1):
        up = upsample_bicubic2x(depth)
        hy = cvp_depth_hypos(up, Kl[0][level], Kl[1][level], Es[0], Es[1], depth_min, depth_max)
        H, W = fps[0][level].shape[1:]
        warped = [orc.homo_warp_mvs(fps[v][level], *rel(v, level), hy, (H, W)) for v in range(1, S + 1)]
        cost = orc.variance(fps[0][level], warped, 1)
        reg = cvp_costreg(sd, cost)
        r = orc.softmax_regress(reg, 1, hy, conf_mode=1)
        depth = r["depth"]
        ests.append(depth)
        seams["hypos_l%d" % level] = hy
        seams["reg_in_l%d" % level] = cost
        seams["reg_out_l%d" % level] = reg
    return {"depth": depth, "depth_est_list": ests[::-1], "conf": r["conf"], "seams": seams}
