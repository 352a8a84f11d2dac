"""Pack the UNMODIFIED hot-path sources of the reference into oracle/_ref/ref_hotpath.zip.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (the product never reads it).

    python oracle/make_ref.py            # build container only: needs /root/reference

/root/reference does not exist on the GPU box, and the reference is Python, so "compiling it into oracle/_ref" means
archiving the files of SURVEY.md section 8(a) byte for byte -- nothing is edited, the licence text of the reference's
README travels with them -- into ONE git-ignored archive that the gpurun snapshot carries to the box (like the built
libmvsb200.so).  No reference source enters the repository's history.  Python imports straight from the archive
(zipimport; oracle/ref_import.py puts it on sys.path), which gives, on the GPU box:
  * `bench.py --impl reference`: the reference's own CPU implementation of the path (`kind: "reference"`),
  * the `gpu_reference` leg of bench.py: the same code through stock PyTorch / cuDNN on the B200,
  * the full-size parity tests (tests/test_gpu_fullsize.py): the reference executed on the same GPU in fp32.
"""
import hashlib
import os
import sys
import zipfile

REFERENCE_ROOT = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
ARCHIVE = os.path.join(OUT_DIR, "ref_hotpath.zip")

# the files the hot path lives in (+ the package markers they are imported through)
FILES = [
    "models/__init__.py",
    "models/MVSNet/__init__.py", "models/MVSNet/model.py", "models/MVSNet/module.py",
    "models/VisMVSNet/__init__.py", "models/VisMVSNet/frontend.py", "models/VisMVSNet/homography.py",
    "models/VisMVSNet/model_cas.py", "models/VisMVSNet/nn_utils.py", "models/VisMVSNet/preproc.py",
    "models/CVP_MVSNet/__init__.py", "models/CVP_MVSNet/frontend.py",
    "models/CVP_MVSNet/models/__init__.py", "models/CVP_MVSNet/models/modules.py", "models/CVP_MVSNet/models/net.py",
    "utils/__init__.py", "utils/utils_3D.py",
    "README.md",   # carries the copyright notice / conditions that must accompany any redistribution
]


def build(root=REFERENCE_ROOT, force=False):
    """Returns the archive path, or None when the reference tree is not present (GPU box: the prebuilt archive is used)."""
    if not os.path.isdir(root):
        return ARCHIVE if os.path.exists(ARCHIVE) else None
    srcs = [os.path.join(root, f) for f in FILES]
    if not force and os.path.exists(ARCHIVE) and all(os.path.getmtime(s) <= os.path.getmtime(ARCHIVE) for s in srcs) \
            and os.path.getmtime(__file__) <= os.path.getmtime(ARCHIVE):
        return ARCHIVE
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = ARCHIVE + ".tmp"
    digest = hashlib.sha256()
    with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
        for rel, src in zip(FILES, srcs):
            data = open(src, "rb").read()
            digest.update(rel.encode() + b"\0" + data)
            info = zipfile.ZipInfo(rel, date_time=(2020, 1, 1, 0, 0, 0))   # fixed stamp: the archive is reproducible
            info.compress_type = zipfile.ZIP_DEFLATED
            z.writestr(info, data)
        z.writestr(zipfile.ZipInfo("MANIFEST.txt", date_time=(2020, 1, 1, 0, 0, 0)),
                   "unmodified files of fdarmon/wild_deep_mvs (see README.md inside for the licence), packed by oracle/make_ref.py\n"
                   "sha256 over (name, bytes): %s\n%s\n" % (digest.hexdigest(), "\n".join(FILES)))
    os.replace(tmp, ARCHIVE)
    return ARCHIVE


if __name__ == "__main__":
    p = build(force="--force" in sys.argv)
    print(p if p else "reference tree not found and no prebuilt archive")
