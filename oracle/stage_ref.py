"""ORACLE / TEST INFRASTRUCTURE ONLY.

Stages the UNMODIFIED reference package for timing on the GPU box's host cores (``bench.py --impl reference``,
``cpu_baseline.kind = "reference"``): byte-for-byte copies of ``/root/reference/src/pathway_explanations`` (the
reference is pure Python: there is nothing to compile), its two checkpoints and ``config/configs.json`` into
``oracle/_ref/`` together with a manifest of SHA-256 digests.  ``oracle/_ref/`` is git-ignored (no reference source
enters the history) but not gpurun-ignored, so it travels to the GPU box, where ``/root/reference`` does not exist.

    python -m oracle.stage_ref          # also run by __graft_entry__.build() when /root/reference is present
"""
import hashlib
import json
import os
import shutil

REFERENCE_ROOT = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref")
PARTS = [("src/pathway_explanations", ".py"), ("test_data", ".tar"), ("config", ".json")]


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()


def stage():
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "pathway_explanations")):
        return None
    manifest = {}
    for rel, ext in PARTS:
        src_dir, dst_dir = os.path.join(REFERENCE_ROOT, rel), os.path.join(STAGED, rel)
        os.makedirs(dst_dir, exist_ok=True)
        for name in sorted(os.listdir(src_dir)):
            if not name.endswith(ext):
                continue
            shutil.copyfile(os.path.join(src_dir, name), os.path.join(dst_dir, name))
            manifest[os.path.join(rel, name)] = sha256(os.path.join(dst_dir, name))
    with open(os.path.join(STAGED, "MANIFEST.json"), "w") as fh:
        json.dump({"source": REFERENCE_ROOT, "files": manifest}, fh, indent=1, sort_keys=True)
    return STAGED


def verify():
    """True when every staged file still has the digest recorded at staging time."""
    try:
        with open(os.path.join(STAGED, "MANIFEST.json")) as fh:
            files = json.load(fh)["files"]
    except OSError:
        return False
    return all(os.path.exists(os.path.join(STAGED, k)) and sha256(os.path.join(STAGED, k)) == v for k, v in files.items())


if __name__ == "__main__":
    print(stage(), "verified" if verify() else "NOT verified")
