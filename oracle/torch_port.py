"""The reference's hot path restated with the ATen calls the reference itself makes
(grid_sample, conv3d, conv_transpose3d, batch_norm, softmax, ...), as pure functions of a
state_dict.  TEST INFRASTRUCTURE ONLY.

Two uses:
  * bench.py's `cpu_baseline` / `--impl reference` leg: the reference is Python and cannot travel to
    the GPU box, so this port -- same library calls, same materialised intermediates, all host
    threads -- is what is timed as "the reference's own CPU path" (kind = "port");
  * full-size parity in tests/: run on the GPU it is the "plain PyTorch fp32 reference of the same op".

Pinned against the reference by tests/test_oracle_golden.py::test_torch_port_*.
"""
import torch
import torch.nn.functional as F

EPS = 1e-5


# ---------------------------------------------------------------------------------------------
# a1: homo_warping -- models/MVSNet/module.py:111-169
# ---------------------------------------------------------------------------------------------
def homo_warp(src, src_proj, ref_proj, depth, ref_hw):
    B, C, Hs, Ws = src.shape
    H, W = ref_hw
    D = depth.shape[1]
    rel = src_proj @ torch.inverse(ref_proj)
    rot, trans = rel[:, :3, :3], rel[:, :3, 3:4]
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=src.device),
                            torch.arange(W, dtype=torch.float32, device=src.device), indexing="ij")
    pix = torch.stack((xs.reshape(-1), ys.reshape(-1), torch.ones(H * W, device=src.device)))  # [3,HW]
    rays = rot @ pix.unsqueeze(0)                                                               # [B,3,HW]
    dv = depth.reshape(B, 1, D, -1)  # [B,1,D,1] or [B,1,D,HW]
    q = rays.unsqueeze(2) * dv + trans.view(B, 3, 1, 1)                                         # [B,3,D,HW]
    xy = q[:, :2] / q[:, 2:3]
    xy = torch.where((q[:, 2:3] <= 0).expand_as(xy), torch.full_like(xy, -10.0), xy)
    gx = xy[:, 0] / ((Ws - 1) / 2) - 1
    gy = xy[:, 1] / ((Hs - 1) / 2) - 1
    grid = torch.stack((gx, gy), dim=3).clamp(-10, 10)
    out = F.grid_sample(src, grid.view(B, D * H, W, 2), mode="bilinear", padding_mode="zeros", align_corners=True)
    return out.view(B, C, D, H, W)


# ---------------------------------------------------------------------------------------------
# a2 / a3: build_cost_volume -- models/MVSNet/model.py:109-176 (eval-mode, in-place flavour)
# ---------------------------------------------------------------------------------------------
def mvsnet_cost_volume(ref, srcs, ref_proj, src_projs, depth, aggregation="variance", temp=None):
    B, C, H, W = ref.shape
    D = depth.shape[1]
    V = len(srcs) + 1
    if aggregation == "variance":
        m1 = ref.unsqueeze(2).repeat(1, 1, D, 1, 1)
        m2 = m1 ** 2
        for s, p in zip(srcs, src_projs):
            w = homo_warp(s, p, ref_proj, depth, (H, W))
            m1 += w
            m2 += w.pow_(2)
            del w
        return m2.div_(V).sub_(m1.pow_(2).div_(V ** 2))
    if aggregation == "softmin":
        r = ref.unsqueeze(2)
        sum_e = torch.zeros(B, 1, D, H, W, device=ref.device)
        sum_v = torch.zeros(B, C, D, H, W, device=ref.device)
        for s, p in zip(srcs, src_projs):
            w = homo_warp(s, p, ref_proj, depth, (H, W))
            w.sub_(r).pow_(2)
            e = torch.exp(-temp * w.sum(dim=1, keepdim=True))
            sum_e.add_(e)
            sum_v.add_(w.mul_(e))
            del w, e
        return sum_v.div_(sum_e + 1e-6)
    raise NotImplementedError(aggregation)


def mvsnet_cost_volume_train(ref, srcs, ref_proj, src_projs, depth, aggregation="variance", temp=None):
    """The TRAINING flavour of build_cost_volume (models/MVSNet/model.py:124-126,151-156: out-of-place updates), the
    graph torch.autograd differentiates in the reference.  Autograd through this function is the oracle of the K1
    backward kernel; the sampling grid inside homo_warp carries no gradient (module.py:127 builds it under no_grad)."""
    B, C, H, W = ref.shape
    D = depth.shape[1]
    V = len(srcs) + 1
    if aggregation == "variance":
        m1 = ref.unsqueeze(2).repeat(1, 1, D, 1, 1)
        m2 = m1 ** 2
        for s, p in zip(srcs, src_projs):
            w = homo_warp(s, p, ref_proj, depth, (H, W))
            m1 = m1 + w
            m2 = m2 + w ** 2
        return m2 / V - m1 ** 2 / (V ** 2)
    if aggregation == "softmin":
        r = ref.unsqueeze(2)
        sum_e = torch.zeros(B, 1, D, H, W, device=ref.device)
        sum_v = torch.zeros(B, C, D, H, W, device=ref.device)
        for s, p in zip(srcs, src_projs):
            diff = (r - homo_warp(s, p, ref_proj, depth, (H, W))) ** 2
            e = torch.exp(-temp * diff.sum(dim=1, keepdim=True))
            sum_e = sum_e + e
            sum_v = sum_v + e * diff
        return sum_v / (sum_e + 1e-6)
    raise NotImplementedError(aggregation)


def mvsnet_regress_train(reg, depth):
    """softmax + depth_regression as differentiated in training (models/MVSNet/model.py:207-209, module.py:174-178)."""
    return torch.sum(F.softmax(reg, dim=1) * depth.view(*depth.shape, *([1] * (reg.dim() - depth.dim()))), 1)


# ---------------------------------------------------------------------------------------------
# a4: CostRegNet -- models/MVSNet/model.py:43-84
# ---------------------------------------------------------------------------------------------
def _bn(sd, p, x, relu=True):
    y = F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.0, EPS)
    return F.relu(y, inplace=True) if relu else y


def mvsnet_costreg(sd, x, prefix="cost_regularization."):
    def cbr(n, x, stride=1):
        return _bn(sd, prefix + n + ".bn", F.conv3d(x, sd[prefix + n + ".conv.weight"], None, stride, 1))

    def dbr(n, x):
        return _bn(sd, prefix + n + ".1", F.conv_transpose3d(x, sd[prefix + n + ".0.weight"], None, 2, 1, 1))

    c0 = cbr("conv0", x)
    c2 = cbr("conv2", cbr("conv1", c0, 2))
    c4 = cbr("conv4", cbr("conv3", c2, 2))
    y = cbr("conv6", cbr("conv5", c4, 2))
    y = c4 + dbr("conv7", y)
    y = c2 + dbr("conv9", y)
    y = c0 + dbr("conv11", y)
    return F.conv3d(y, sd[prefix + "prob.weight"], sd[prefix + "prob.bias"], 1, 1)


# ---------------------------------------------------------------------------------------------
# a5 / a6: softmax + regression + confidence -- models/MVSNet/model.py:207-215
# ---------------------------------------------------------------------------------------------
def mvsnet_head(reg, depth):
    """reg [B,D,H,W]; depth [B,D] -> depth [B,H,W], conf [B,H,W]."""
    D = reg.shape[1]
    prob = F.softmax(reg, dim=1)
    est = torch.sum(prob * depth.view(*depth.shape, 1, 1), 1)
    sum4 = 4 * F.avg_pool3d(F.pad(prob.unsqueeze(1), pad=(0, 0, 0, 0, 1, 2)), (4, 1, 1), stride=1, padding=0).squeeze(1)
    idx = torch.sum(prob * torch.arange(D, device=reg.device, dtype=torch.float).view(1, D, 1, 1), 1).long()
    conf = torch.gather(sum4, 1, idx.unsqueeze(1)).squeeze(1)
    return est, conf


def mvsnet_hot_path(sd, feats, projs, depth, aggregation="variance", temp=None, timings=None):
    """feats: list of [B,C,H,W] (0 = reference); projs: list of [B,4,4]; depth [B,D]."""
    import time
    t0 = time.perf_counter()
    cost = mvsnet_cost_volume(feats[0], feats[1:], projs[0], projs[1:], depth, aggregation, temp)
    t1 = time.perf_counter()
    reg = mvsnet_costreg(sd, cost).squeeze(1)
    del cost
    t2 = time.perf_counter()
    est, conf = mvsnet_head(reg, depth)
    t3 = time.perf_counter()
    if timings is not None:
        timings.update(build=t1 - t0, regularise=t2 - t1, regress=t3 - t2)
    return est, conf
