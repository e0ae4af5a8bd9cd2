"""Recipe for `oracle/_ref/`: the UNMODIFIED reference packages the hot path needs, copied from /root/reference so that the
GPU box (where /root/reference does not exist) can time the reference's own CPU code next to the CUDA path
(`bench.py --impl reference`, `cpu_baseline.kind == "reference"`).

TEST / MEASUREMENT INFRASTRUCTURE — never imported by the product path.  `oracle/_ref/` is git-ignored (no reference source
enters the history) but not gpurun-ignored, so it travels with the working tree like the built `.so` files.  Run by
`__graft_entry__.build()` whenever /root/reference is present; `MANIFEST.json` records the SHA-256 of every copied file so a
reader can check that nothing was edited.  Copied: the four Python packages the harness imports (oracle/ref_harness.py) —
utils/, data_process/, archs/, losses/ — `*.py` only (~420 KB).
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
PACKAGES = ("utils", "data_process", "archs", "losses")


def build(src="/root/reference", dest=DEST, quiet=False):
    if not os.path.isdir(os.path.join(src, "data_process")):
        return None
    manifest = {}
    for pkg in PACKAGES:
        for root, _dirs, files in os.walk(os.path.join(src, pkg)):
            for name in sorted(files):
                if not name.endswith(".py"):
                    continue
                s = os.path.join(root, name)
                rel = os.path.relpath(s, src)
                d = os.path.join(dest, rel)
                os.makedirs(os.path.dirname(d), exist_ok=True)
                data = open(s, "rb").read()
                if not os.path.exists(d) or open(d, "rb").read() != data:
                    shutil.copyfile(s, d)
                manifest[rel] = hashlib.sha256(data).hexdigest()
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "files": manifest}, f, indent=1, sort_keys=True)
    if not quiet:
        print(f"oracle/_ref: {len(manifest)} unmodified reference files from {src}")
    return dest


def verify(dest=DEST):
    """True when every file listed in the manifest is present with the recorded hash."""
    try:
        manifest = json.load(open(os.path.join(dest, "MANIFEST.json")))["files"]
    except (OSError, ValueError, KeyError):
        return False
    for rel, digest in manifest.items():
        try:
            if hashlib.sha256(open(os.path.join(dest, rel), "rb").read()).hexdigest() != digest:
                return False
        except OSError:
            return False
    return bool(manifest)


if __name__ == "__main__":
    build()
