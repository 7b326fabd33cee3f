"""Recipe for oracle/_ref: the UNMODIFIED reference implementation of the path, staged so that it travels to the GPU box.

The reference is pure Python on top of torch; the files of the hot path (models/superglue/*.py,
models/utils.py) and of the two neighbouring steps that import with torch alone (utils/losses.py, utils/misc.py,
models/gt_matches_generation.py) are copied byte for byte from where they lie under /root/reference into oracle/_ref/
(git-ignored build output, like a compiled .so; NOT gpurun-ignored).  Nothing here is product code: only tests/, smoke() and
bench.py's CPU legs import it - as the checker and as the CPU baseline (`cpu_baseline.kind = "reference"`).

    python oracle/build_ref.py            # (re)stage; no-op with a message when /root/reference is absent (GPU box)
"""
from __future__ import annotations

import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get('OG_REFERENCE_ROOT', '/root/reference')
OUT = os.path.join(HERE, '_ref')
FILES = [
    'models/__init__.py', 'models/utils.py', 'models/gt_matches_generation.py',
    'models/superglue/__init__.py', 'models/superglue/superglue.py', 'models/superglue/attention_gnn.py',
    'models/superglue/attention.py', 'models/superglue/optimal_transport.py', 'models/superglue/positional_encoding.py',
    'utils/__init__.py', 'utils/losses.py', 'utils/misc.py',
]


def available() -> bool:
    return all(os.path.exists(os.path.join(OUT, f)) for f in FILES)


def build_ref(verbose: bool = False) -> bool:
    """Stage the reference files; returns True when oracle/_ref is complete afterwards."""
    if not os.path.isdir(REF_ROOT):
        if verbose:
            print(f'oracle/_ref: {REF_ROOT} not present (GPU box?): using the staged copy' if available() else
                  f'oracle/_ref: {REF_ROOT} not present and nothing staged: the oracle port stands in')
        return available()
    for f in FILES:
        src, dst = os.path.join(REF_ROOT, f), os.path.join(OUT, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
    if verbose:
        print(f'oracle/_ref: staged {len(FILES)} files from {REF_ROOT}')
    return available()


def import_reference():
    """(SuperGlue class, criterion, generate_gt_matches) of the staged reference, or None when it is not staged.
    The reference's top-level package names (`models`, `utils`) are only put on sys.path here, on request."""
    if not available():
        return None
    if OUT not in sys.path:
        sys.path.insert(0, OUT)
    from models.superglue.superglue import SuperGlue          # noqa: E402  (the reference's own module)
    from models.gt_matches_generation import generate_gt_matches
    from utils.losses import criterion
    return SuperGlue, criterion, generate_gt_matches


if __name__ == '__main__':
    ok = build_ref(verbose=True)
    sys.exit(0 if ok or not os.path.isdir(REF_ROOT) else 1)
