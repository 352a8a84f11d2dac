"""CPU: host-side logic of the drop-in package that needs no GPU -- engine selection predicates of the C library
(pure host code), view batching of the feature extractors, abs-max scalar pool."""
import ctypes

import torch

from wild_deep_mvs_b200 import _lib, ops


def _desc(cin, cout, stride=1, transposed=0, cin2=0, k=3, dims=(8, 16, 32)):
    d = _lib.Conv3dDesc()
    d.B, (d.D, d.H, d.W) = 1, dims
    d.Cin, d.Cin2, d.Cout = cin, cin2, cout
    d.kd = d.kh = d.kw = k
    d.stride, d.transposed = stride, transposed
    return d


def test_engine_support_predicates():
    lib = _lib.load()
    zm = lambda *a, **k: lib.mvsb200_conv3d_zm_supported(ctypes.byref(_desc(*a, **k)))
    c1 = lambda *a, **k: lib.mvsb200_conv3d_c1_supported(ctypes.byref(_desc(*a, **k)))
    tc = lambda *a, **k: lib.mvsb200_conv3d_tc_supported(ctypes.byref(_desc(*a, **k)))
    # every 3x3x3 layer of the three regularisers with resident-size weights runs on the z-march engine
    for cin, cout, stride, tr in ((32, 8, 1, 0), (8, 16, 2, 0), (16, 16, 1, 0), (16, 32, 2, 0), (32, 32, 1, 0), (32, 64, 2, 0),
                                  (64, 32, 2, 1), (32, 16, 2, 1), (16, 8, 2, 1), (8, 8, 1, 0), (32, 64, 1, 0)):
        assert zm(cin, cout, stride, tr) == 1, (cin, cout, stride, tr)
    assert zm(8, 8, cin2=8) == 1                      # torch.cat([up, enc], 1) of the Vis U-Net
    assert zm(64, 64) == 0 and tc(64, 64) == 1        # weights too large to stay resident: tile engine
    assert zm(8, 1) == 0 and c1(8, 1) == 1            # single-output-channel heads: CUDA cores
    assert c1(16, 1) == 1 and c1(12, 1) == 0 and c1(8, 1, stride=2) == 0
    assert zm(8, 8, k=1) == 0 and zm(12, 8) == 0 and zm(8, 12) == 0
    # packed size: 16-byte header + [N blocks][variants][chunks] regions of 9 blocks of 2*NC*16 bytes
    n = lib.mvsb200_conv3d_zm_packed_bytes(ctypes.byref(_desc(32, 8)))
    assert n == 16 + 1 * 1 * 4 * 9 * (2 * 48 * 16)
    n = lib.mvsb200_conv3d_zm_packed_bytes(ctypes.byref(_desc(16, 32, 2)))
    assert n == 16 + 2 * 2 * 2 * 9 * (2 * 64 * 16)


def test_map_views_batches_equal_sizes_and_keeps_ragged_lists():
    calls = []

    def extract(x):
        calls.append(tuple(x.shape))
        return [x * 2, x[:, :, ::2, ::2] + 1]

    a, b, c = torch.rand(2, 3, 8, 8), torch.rand(2, 3, 8, 8), torch.rand(2, 3, 8, 8)
    out = ops.map_views(extract, [a, b, c])
    assert calls == [(6, 3, 8, 8)] and len(out) == 3
    assert torch.equal(out[1][0], b * 2) and torch.equal(out[2][1], c[:, :, ::2, ::2] + 1)
    calls.clear()
    d = torch.rand(2, 3, 6, 10)
    out = ops.map_views(lambda x: x + 1, [a, d])          # different sizes (MegaDepth / YFCC test mode): one call per view
    assert len(out) == 2 and torch.equal(out[1], d + 1) and out[0].shape == a.shape
    out = ops.map_views(lambda x: x + 1, [a, b])          # tensor-returning extractor
    assert torch.equal(out[0], a + 1) and torch.equal(out[1], b + 1)


def test_amax_pool_hands_out_distinct_zeroed_scalars():
    pool = ops.AmaxPool("cpu", 4)
    got = [pool.take() for _ in range(9)]
    assert all(g.shape == (1,) and float(g) == 0.0 for g in got)
    ptrs = {g.data_ptr() for g in got}
    assert len(ptrs) == 9
