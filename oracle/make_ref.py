"""Recipe for oracle/_ref: the UNMODIFIED reference package, copied byte for byte from /root/reference at build time.

    python oracle/make_ref.py            (also run by __graft_entry__.build() when /root/reference is present)

Test infrastructure, not product: nothing under superconductor_vae_b200/ imports it.  The reference is pure
Python/PyTorch with no setup.py / pyproject.toml (so `pip install --target` has nothing to build); its "build" is a copy
of the source tree `src/superconductor/` plus the two vocabulary files its tokenizer reads (`data/fraction_vocab.json`,
`data/isotope_vocab.json`).  The copy lives in oracle/_ref/, which is git-ignored (reference sources never enter this
repository's history) but NOT gpurun-ignored, so it travels to the GPU box where /root/reference does not exist.
Users: `bench.py --impl reference` (times the reference's own `generate_with_kv_cache` on the box's host cores),
`bench.py`'s cpu_baseline leg, and tests that compare the oracle restatement with the live reference module.
A MANIFEST (relative path, size, sha256 of every copied file) is written next to the copy."""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("SCV_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
COPY = [("src/superconductor", "src/superconductor"),
        ("data/fraction_vocab.json", "data/fraction_vocab.json"),
        ("data/isotope_vocab.json", "data/isotope_vocab.json")]


def build(verbose: bool = True) -> str:
    if not os.path.isdir(REF_ROOT):
        if os.path.isdir(os.path.join(DST, "src", "superconductor")):
            return DST                      # GPU box: use the copy that travelled with the snapshot
        raise FileNotFoundError(f"{REF_ROOT} is absent and oracle/_ref has not been built")
    manifest = {}
    for src_rel, dst_rel in COPY:
        src, dst = os.path.join(REF_ROOT, src_rel), os.path.join(DST, dst_rel)
        if os.path.isdir(src):
            if os.path.isdir(dst):
                shutil.rmtree(dst)
            shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        else:
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
    for base, _, files in os.walk(DST):
        for f in sorted(files):
            if f == "MANIFEST.json" or f.endswith(".pyc"):
                continue
            p = os.path.join(base, f)
            with open(p, "rb") as fh:
                blob = fh.read()
            manifest[os.path.relpath(p, DST)] = {"bytes": len(blob), "sha256": hashlib.sha256(blob).hexdigest()}
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": REF_ROOT, "files": manifest}, fh, indent=1, sort_keys=True)
    if verbose:
        print(f"oracle/_ref: {len(manifest)} files copied unmodified from {REF_ROOT}")
    return DST


if __name__ == "__main__":
    build()
    sys.exit(0)
