"""Stage the UNMODIFIED reference package under oracle/_ref/ (TEST INFRASTRUCTURE; git-ignored output).

    python oracle/build_ref.py            # authoring container only: needs /root/reference

The reference is pure Python with no setup.py / pyproject.toml, so `pip install --target` has nothing to
build; this recipe does what such an install would do: it copies the `itr` package's *.py files, byte for
byte, from where they lie under /root/reference into oracle/_ref/itr/ (vocabularies and figures are not
needed by the arithmetic and stay behind).  oracle/_ref/ is listed in .gitignore (reference sources never
enter the history) but not in .gpurunignore, so it travels to the GPU box like a built .so: there
`bench.py --impl reference`, the `cpu_baseline` leg and the same-GPU baseline time the reference's own
xattn_score_* / i2t / t2i / cal_sims instead of the oracle's port (`cpu_baseline.kind` = "reference").
Nothing in the product imports it.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("ITR_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")


def build(verbose=True) -> bool:
    pkg = os.path.join(SRC, "itr")
    if not os.path.isdir(pkg):
        if verbose:
            print("oracle/build_ref.py: no reference tree at", SRC, "- keeping whatever oracle/_ref holds")
        return False
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    manifest = []
    for root, dirs, files in os.walk(pkg):
        dirs[:] = [d for d in dirs if d not in ("__pycache__", "vocab")]
        for f in sorted(files):
            if not f.endswith(".py"):
                continue
            src = os.path.join(root, f)
            rel = os.path.relpath(src, SRC)
            dst = os.path.join(DST, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
            manifest.append("{}  {}".format(hashlib.sha256(open(src, "rb").read()).hexdigest(), rel))
    with open(os.path.join(DST, "MANIFEST.sha256"), "w") as fh:
        fh.write("\n".join(manifest) + "\n")
    if verbose:
        print("oracle/_ref: {} files staged from {}".format(len(manifest), SRC))
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
