"""TEST / BASELINE INFRASTRUCTURE — recipe that places the UNMODIFIED reference's Python sources where the GPU box can
import them: /root/reference exists only in the build container, `oracle/_ref/` travels with the snapshot (it is git-ignored,
never committed: the reference's sources are not part of this repository's history).

    python oracle/build_ref.py        # copies /root/reference/{src,experiments/utils.py,LICENSE} -> oracle/_ref/, byte for byte

The reference is pure Python (BSD-3, no build system, nothing to compile), so "building" it is a verbatim copy plus a
MANIFEST.json of SHA-256 digests; `verify()` re-checks the digests before the copy is used, so the baseline that
`bench.py --impl reference` times and the modules the `_ref` parity tests drive are provably the reference's own files.
The torch-1.7 -> 2.x name shim lives in oracle/ref_harness.py and patches torch's namespaces, never these files."""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC_ROOT = os.environ.get("QBN_REFERENCE_SRC", "/root/reference")
WANTED = ("src", os.path.join("experiments", "utils.py"), "LICENSE", "requirements.txt")


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def _walk(root):
    for base, _, files in os.walk(root):
        for name in sorted(files):
            if name.endswith((".pyc",)) or "__pycache__" in base:
                continue
            yield os.path.join(base, name)


def build(verbose=True):
    if not os.path.isdir(os.path.join(SRC_ROOT, "src")):
        raise RuntimeError("reference sources not found at %s (this recipe runs in the build container only)" % SRC_ROOT)
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    manifest = {}
    for item in WANTED:
        src = os.path.join(SRC_ROOT, item)
        if os.path.isdir(src):
            files = list(_walk(src))
        elif os.path.isfile(src):
            files = [src]
        else:
            continue
        for f in files:
            rel = os.path.relpath(f, SRC_ROOT)
            dst = os.path.join(DEST, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(f, dst)
            manifest[rel] = _sha(dst)
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC_ROOT, "files": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print("oracle/_ref: %d files of the unmodified reference (%s)" % (len(manifest), SRC_ROOT))
    return DEST


def available():
    return os.path.isfile(os.path.join(DEST, "MANIFEST.json")) and os.path.isdir(os.path.join(DEST, "src", "models", "stochastic"))


def verify():
    """True iff every file of oracle/_ref still has the digest recorded when it was copied from the reference."""
    if not available():
        return False
    with open(os.path.join(DEST, "MANIFEST.json")) as f:
        files = json.load(f)["files"]
    return all(os.path.isfile(os.path.join(DEST, rel)) and _sha(os.path.join(DEST, rel)) == h for rel, h in files.items())


if __name__ == "__main__":
    build()
    sys.exit(0 if verify() else 1)
