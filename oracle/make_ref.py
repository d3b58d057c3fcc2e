#!/usr/bin/env python
"""Build ``oracle/_ref``: the UNMODIFIED reference files of the hot path, so that the reference arm of ``bench.py`` and
the GPU box (where ``/root/reference`` does not exist) can import the reference itself.  TEST / BASELINE INFRASTRUCTURE.

    python oracle/make_ref.py [REFERENCE_ROOT]        (default /root/reference)

The reference is a pure-Python script tree (no C / C++ / CUDA, no setup.py): "building" it means copying, byte for
byte, the four files SURVEY.md Appendix D names - ``models/__init__.py``, ``models/MVS4Net.py``,
``models/mvs4net_utils.py`` (stage loop, schedules, ``stagenet``, ``homo_warping``, ``sinkhorn``) and ``test_mvs4.py``
(the fusion filter, lifted with ``ast`` because the file parses ``sys.argv`` at import).  ``oracle/_ref/`` is
git-ignored (reference sources never enter this repository's history) but travels to the GPU box with ``gpurun``.
A ``MANIFEST.json`` with the SHA-256 of every copied file records what the baseline was.
"""
import hashlib
import json
import os
import shutil
import sys

FILES = ["models/__init__.py", "models/MVS4Net.py", "models/mvs4net_utils.py", "test_mvs4.py"]
HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")


def make_ref(src_root: str = "/root/reference", quiet: bool = False) -> bool:
    """Copy the reference files; returns False (and leaves an existing copy alone) when the reference is absent."""
    if not all(os.path.exists(os.path.join(src_root, f)) for f in FILES):
        return False
    manifest = {}
    for f in FILES:
        dst = os.path.join(DEST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(src_root, f), dst)
        with open(dst, "rb") as fh:
            manifest[f] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": src_root, "sha256": manifest}, fh, indent=1)
    if not quiet:
        print("[make_ref] %d reference files -> %s" % (len(FILES), DEST))
    return True


if __name__ == "__main__":
    ok = make_ref(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    sys.exit(0 if ok else 1)
