"""On-disk formats either side of the depth path (row f4 of SURVEY.md 8): what the reference writes after `forward`
and reads before it.  Host code (numpy), byte-compatible with the reference's own readers / writers:

* `<name>_out.npz` with `depthmap` / `probability`          evaluation/run_depthmaps.py:63-68 (written),
                                                            evaluation/filtering.py:53-57, fusibile.py:136-138 (read)
* Gipuma `.dmb` images and `.P` cameras (fusibile input)    evaluation/fusibile.py:27-72
* COLMAP dense arrays (`*.geometric.bin`, `*.photometric.bin`)  utils/colmap_utils.py:233-279
* PFM depth maps (DTU / BlendedMVS ground truth)            data/MVSDataset.py:152-187

The quirks of the reference are part of the format and are kept (each noted where it happens).  Pinned by files the
UNMODIFIED reference functions wrote (tests/golden/make_golden_formats.py -> tests/golden/formats/).
"""
import struct

import numpy as np


# ------------------------------------------------------------------------------------------------
# depth-map archives
# ------------------------------------------------------------------------------------------------
def save_depth_npz(path, depth, confidence):
    """One view's result as the reference stores it (run_depthmaps.py:63-68): keys `probability`, `depthmap`,
    deflate-compressed.  depth [h,w]; confidence [h,w] (MVSNet), [3,h,w] (Vis-MVSNet) or [1,h,w] (CVP-MVSNet)."""
    np.savez_compressed(path, probability=_np32(confidence), depthmap=_np32(depth))


def load_depth_npz(path):
    """-> (depthmap, probability); `probability` is None for archives that hold only a depth map (the files the
    geometric filter is fed in the reference's tests of it)."""
    with np.load(path) as z:
        return z["depthmap"], (z["probability"] if "probability" in z.files else None)


def _np32(a):
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return np.asarray(a, dtype=np.float32)


# ------------------------------------------------------------------------------------------------
# Gipuma / fusibile
# ------------------------------------------------------------------------------------------------
def write_gipuma_dmb(path, image):
    """Header <i type=1, height, width, channels> + float32 payload (fusibile.py:41-62).  A 2-D image is written
    row-major.  A 3-D [h,w,c] image is written channel-plane by channel-plane, exactly as the reference does (its
    reader assumes pixel-interleaved data, so only constant-per-pixel images such as the fake normals survive the
    round trip -- kept, because fusibile is fed these very bytes)."""
    image = np.asarray(image)
    h, w = image.shape[0], image.shape[1]
    c = image.shape[2] if image.ndim == 3 else 1
    if image.ndim == 3:
        image = np.transpose(image, (2, 0, 1)).squeeze()
    with open(path, "wb") as f:
        f.write(struct.pack("<iiii", 1, h, w, c))
        f.write(np.ascontiguousarray(image).tobytes())


def read_gipuma_dmb(path):
    """-> [h,w] or [h,w,c] float32 (fusibile.py:27-38): payload read as pixel-interleaved, x fastest."""
    with open(path, "rb") as f:
        _type, h, w, c = struct.unpack("<iiii", f.read(16))
        data = np.frombuffer(f.read(), dtype=np.float32)
    return np.transpose(data.reshape((w, h, c), order="F"), (1, 0, 2)).squeeze()


def write_gipuma_cam(projection_matrix, path):
    """3x4 projection matrix as text, one row per line, every entry followed by a blank, then an empty line
    (fusibile.py:65-72).  Entries are printed with Python's shortest round-trip repr of the value's own dtype."""
    with open(path, "w") as f:
        for i in range(3):
            for j in range(4):
                f.write(str(projection_matrix[i][j]) + " ")
            f.write("\n")
        f.write("\n")


def gipuma_projection(K, R, t, downscale=1):
    """The matrix `mvsnet_to_gipuma` writes for a view (fusibile.py:112-124): [K R | K t] with its first two rows
    divided by `downscale`, in float64.  K, R [3,3]; t [3,1] or [3] (numpy, float32 like the reference's batches)."""
    K, R = np.asarray(K, dtype=np.float32), np.asarray(R, dtype=np.float32)
    t = np.asarray(t, dtype=np.float32).reshape(3, 1)
    P = np.concatenate([K @ R, K @ t], axis=1)
    P[:2] /= np.float32(downscale)
    return P.astype(np.float64)


def fake_gipuma_normal(depth_dmb_path, normal_dmb_path):
    """Unit normals (1,1,1)/1.732050808 wherever the depth is positive (fusibile.py:75-93)."""
    depth = read_gipuma_dmb(depth_dmb_path)
    h, w = depth.shape
    normal = np.tile(np.ones_like(depth).reshape(h, w, 1), [1, 1, 3]) / 1.732050808
    mask = np.float32(np.tile(np.where(depth > 0, 1, 0).reshape(h, w, 1), [1, 1, 3]))
    write_gipuma_dmb(normal_dmb_path, np.float32(np.multiply(normal, mask)))


def export_gipuma_view(folder, depth, invalid_mask):
    """`disp.dmb` + `normals.dmb` of one view the way `mvsnet_to_gipuma` produces them (fusibile.py:128-158):
    filtered-out pixels get depth 0.  `folder` must exist."""
    import os
    depth = np.array(depth, dtype=np.float32)
    depth[np.asarray(invalid_mask, dtype=bool)] = 0
    write_gipuma_dmb(os.path.join(folder, "disp.dmb"), depth)
    fake_gipuma_normal(os.path.join(folder, "disp.dmb"), os.path.join(folder, "normals.dmb"))


# ------------------------------------------------------------------------------------------------
# COLMAP dense arrays
# ------------------------------------------------------------------------------------------------
def write_colmap_array(array, path):
    """COLMAP `Mat<float>::Write`: ASCII header `width&height&channels&` then float32 little-endian, x fastest, then
    y, then channel (colmap_utils.py:249-279)."""
    array = np.asarray(array)
    if array.dtype != np.float32:
        raise ValueError("COLMAP arrays are float32 (got %s)" % array.dtype)
    if array.ndim == 2:
        h, w = array.shape
        c = 1
        data = np.transpose(array, (1, 0))
    elif array.ndim == 3:
        h, w, c = array.shape
        data = np.transpose(array, (1, 0, 2))
    else:
        raise ValueError("COLMAP arrays are [h,w] or [h,w,c]")
    with open(path, "wb") as f:
        f.write(("%d&%d&%d&" % (w, h, c)).encode("ascii"))
        f.write(data.reshape(-1, order="F").astype("<f4").tobytes())


def read_colmap_array(path):
    """-> [h,w] or [h,w,c] float32 (colmap_utils.py:233-247)."""
    with open(path, "rb") as f:
        head = b""
        while head.count(b"&") < 3:
            ch = f.read(1)
            if not ch:
                raise ValueError("%s: truncated COLMAP array header" % path)
            head += ch
        w, h, c = (int(v) for v in head.split(b"&")[:3])
        data = np.frombuffer(f.read(), dtype="<f4")
    return np.transpose(data.reshape((w, h, c), order="F"), (1, 0, 2)).squeeze()


# ------------------------------------------------------------------------------------------------
# PFM
# ------------------------------------------------------------------------------------------------
def read_pfm(path):
    """-> (data [h,w] or [h,w,3] float32 top row first, scale) (data/MVSDataset.py:152-187).  Header: magic `PF` (colour)
    or `Pf` (grey), `width height` separated by ONE whitespace character, then a signed scale whose sign gives the byte
    order (negative = little endian); the rows follow bottom-up.  Same exceptions as the reference's reader."""
    with open(path, "rb") as f:
        magic = f.readline().decode("utf-8").rstrip()
        channels = {"PF": 3, "Pf": 1}.get(magic)
        if channels is None:
            raise Exception("Not a PFM file.")
        # the reference's header rule is re.match(r'^(\d+)\s(\d+)\s$', line) (data/MVSDataset.py:152-187): digits, ONE whitespace
        # character, digits, ONE whitespace character -- where `$` also matches just before a trailing newline, so
        # "640 480\r\n" and "640 480 \n" are accepted as well as "640 480\n"
        dims = f.readline().decode("utf-8")
        body = dims[:-1] if dims.endswith("\n") and len(dims) >= 2 and dims[-2].isspace() else dims
        i = 0
        while i < len(body) and body[i].isdigit():
            i += 1
        j = i + 1
        while j < len(body) and body[j].isdigit():
            j += 1
        if not (0 < i < len(body) and body[i].isspace() and j > i + 1 and j == len(body) - 1 and body[j].isspace()):
            raise Exception("Malformed PFM header.")
        parts = (body[:i], body[i + 1:j])
        w, h = int(parts[0]), int(parts[1])
        scale = float(f.readline().rstrip())
        data = np.frombuffer(f.read(), dtype=("<" if scale < 0 else ">") + "f4")
    shape = (h, w, 3) if channels == 3 else (h, w)
    return np.flipud(data.reshape(shape)), abs(scale)


def write_pfm(path, image, scale=1.0):
    """Inverse of read_pfm (little endian).  The reference only reads PFM; the writer exists for round trips."""
    image = np.asarray(image, dtype="<f4")
    if image.ndim == 3 and image.shape[2] == 3:
        tag = b"PF\n"
    elif image.ndim == 2:
        tag = b"Pf\n"
    else:
        raise ValueError("PFM images are [h,w] or [h,w,3]")
    with open(path, "wb") as f:
        f.write(tag)
        f.write(("%d %d\n" % (image.shape[1], image.shape[0])).encode("ascii"))
        f.write(("%f\n" % -abs(scale)).encode("ascii"))
        f.write(np.flipud(image).tobytes())
