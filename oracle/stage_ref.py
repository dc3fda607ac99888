"""Stage the UNMODIFIED reference tree into oracle/_ref/ (git-ignored, NOT gpurun-ignored).

TEST / MEASUREMENT INFRASTRUCTURE.  The reference is pure Python: there is nothing to compile, "building" it means
making its own files importable on the GPU box, where /root/reference does not exist.  ``__graft_entry__.build()`` runs
this in the dev container; the copy then travels with the gpurun snapshot exactly like libirs_b200.so does.  Nothing is
edited: bench.py --impl reference / --impl torch_gpu import the staged files through oracle/ref_shim.py (shims D1-D3 are
applied at import time, in memory).  The staged tree is never tracked by git and never imported by the product package.

    python oracle/stage_ref.py            # copies /root/reference -> oracle/_ref (no-op if the source is absent)
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC_CANDIDATES = [os.environ.get("IRS_REF_SRC", ""), "/root/reference"]


def stage(verbose: bool = False) -> str | None:
    """Returns the staged path, or None when there is neither a source tree nor an earlier copy."""
    for src in SRC_CANDIDATES:
        if src and os.path.isfile(os.path.join(src, "model", "influentialRS.py")):
            shutil.copytree(src, DST, dirs_exist_ok=True,
                            ignore=shutil.ignore_patterns(".git", "__pycache__", "*.pyc"))
            os.chmod(DST, 0o755)
            for root, dirs, files in os.walk(DST):            # the source tree is read-only; the copy must be replaceable
                for n in dirs + files:
                    try:
                        os.chmod(os.path.join(root, n), 0o755 if n in dirs else 0o644)
                    except OSError:
                        pass
            if verbose:
                print(f"staged {src} -> {DST}")
            return DST
    return DST if os.path.isfile(os.path.join(DST, "model", "influentialRS.py")) else None


if __name__ == "__main__":
    p = stage(verbose=True)
    sys.exit(0 if p else 1)
