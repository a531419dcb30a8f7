"""Installs the reference's OWN Python glue for the hot path, unmodified, under baseline/_ref/ (git-ignored; it
travels to the GPU box with gpurun like the built .so files, /root/reference itself does not exist there).

    python baseline/install_ref.py          # no-op when /root/reference is absent and baseline/_ref is populated

What is installed (byte-for-byte copies; baseline/_ref/MANIFEST.json records source path + sha256 of every file and
`verify()` re-hashes them, so "unmodified" is checkable on the GPU box without the source tree):
    models/{__init__,custom_functions,rendering,rendering_noCUDA,networks,global_var,ref_util,implicit_mask}.py  losses.py  metrics.py
(models/networks_noCUDA.py is not importable: it needs the non-existent models/rendering_old)
Nothing under baseline/_ref is ever committed or imported by the product package; it is the reference arm of
bench.py (`reference_gpu`) and of tests/test_reference_glue_gpu.py, with `vren` bound to the reference's own
kernels (oracle/_ref/vren_ref.so) or to ours, and `tinycudann` to baseline/tcnn_standin.py or to ours.
`pip install /root/reference` is not applicable: the reference is a script tree, not a package (no setup.py / pyproject
at its root; its only build artefact is models/csrc's `vren` extension, built by oracle/build_ref.py).
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref")
FILES = ["models/__init__.py", "models/custom_functions.py", "models/rendering.py", "models/networks.py",
         "models/global_var.py", "models/ref_util.py", "models/implicit_mask.py", "models/rendering_noCUDA.py", "losses.py", "metrics.py"]


def _sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def install(verbose=False):
    """-> baseline/_ref path, or None when neither the reference tree nor a previous install exists."""
    man_path = os.path.join(OUT, "MANIFEST.json")
    if not os.path.isdir(REF):
        return OUT if os.path.exists(man_path) else None
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = {"source": src, "sha256": _sha(src)}
        if verbose:
            print("[install_ref]", rel, manifest[rel]["sha256"][:12])
    json.dump(manifest, open(man_path, "w"), indent=1)
    return OUT


def verify():
    """True iff every installed file still hashes to what was recorded at install time (and, when the reference tree
    is mounted, to the file it was copied from)."""
    man_path = os.path.join(OUT, "MANIFEST.json")
    if not os.path.exists(man_path):
        return False
    for rel, rec in json.load(open(man_path)).items():
        if _sha(os.path.join(OUT, rel)) != rec["sha256"]:
            return False
        if os.path.exists(rec["source"]) and _sha(rec["source"]) != rec["sha256"]:
            return False
    return True


if __name__ == "__main__":
    print("reference glue installed at:", install(verbose="-v" in sys.argv), "verified:", verify())
