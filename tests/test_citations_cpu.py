"""CPU: every reference citation (file:line) in the header, the docs and the product docstrings points at an
existing line of the reference tree.  Skipped where /root/reference is absent (the GPU box)."""
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
# e.g. tf_ops/tf_sampling.cu:111-176, tf_grouping.cpp:80-87, pointnet_util.py:18-60, model.py:22
CITE = re.compile(r"\b((?:[\w./]+/)?(?:tf_\w+|pointnet_util|tf_util|model|train|predict|semantic_dataset|"
                  r"test_tf_ops|test_interpolate|interpolate|kitti_predict)\.(?:cu|cpp|py)):(\d+)(?:-(\d+))?")


def _ref_files():
    out = {}
    for p in glob.glob(os.path.join(REF, "**", "*"), recursive=True):
        if os.path.isfile(p) and p.endswith((".cu", ".cpp", ".py", ".json", ".cmake")):
            out.setdefault(os.path.basename(p), []).append(p)
    return out


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_reference_citations_resolve():
    files = _ref_files()
    lengths = {}
    sources = [os.path.join(ROOT, "include", "pn2_b200.h"), os.path.join(ROOT, "DESIGN.md"),
               os.path.join(ROOT, "INTEGRATION.md")]
    sources += glob.glob(os.path.join(ROOT, "open3d-pointnet2-semantic3d_b200", "**", "*.py"), recursive=True)
    sources += glob.glob(os.path.join(ROOT, "open3d-pointnet2-semantic3d_b200", "csrc", "*.cu*"))
    sources += [os.path.join(ROOT, "oracle", f) for f in ("pn2_oracle.c", "layers_ref.py", "ref_shim.cu")]
    bad, checked = [], 0
    for src in sources:
        for m in CITE.finditer(open(src, errors="replace").read()):
            name, lo, hi = os.path.basename(m.group(1)), int(m.group(2)), int(m.group(3) or m.group(2))
            cands = files.get(name)
            if not cands:
                bad.append("%s cites %s: no such file in the reference" % (os.path.relpath(src, ROOT), m.group(0)))
                continue
            n = max(lengths.setdefault(c, sum(1 for _ in open(c, errors="replace"))) for c in cands)
            checked += 1
            if not (1 <= lo <= hi <= n):
                bad.append("%s cites %s but %s has %d lines" % (os.path.relpath(src, ROOT), m.group(0), name, n))
    assert checked > 100, checked
    assert not bad, "\n".join(bad[:20])
