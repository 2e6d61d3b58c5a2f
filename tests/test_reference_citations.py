"""Every `path:line` citation of the reference in include/tqgpu.h, DESIGN.md and INTEGRATION.md points into an existing file of
/root/reference with at least that many lines.  Runs only where the reference tree is present (not on the GPU box)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")

CITE = re.compile(r"((?:c\+\+/triqs|python/triqs|test/c\+\+|test/python|doc|benchmarks)/[A-Za-z0-9_+./-]+\.(?:hpp|cpp|py|rst)):(\d+)")


def _line_count(path, cache={}):
    if path not in cache:
        with open(path, "rb") as f:
            cache[path] = sum(1 for _ in f)
    return cache[path]


DOCS = ["include/tqgpu.h", "DESIGN.md", "INTEGRATION.md", "oracle/triqs_oracle.py", "triqs_b200/gf.py", "triqs_b200/cpp/triqs_b200/gfs.hpp"] + \
    sorted(os.path.join("triqs_b200/csrc", f) for f in os.listdir(os.path.join(ROOT, "triqs_b200/csrc")) if f.endswith((".cu", ".cuh")))


@pytest.mark.parametrize("doc", DOCS)
def test_citations_resolve(doc):
    text = open(os.path.join(ROOT, doc)).read()
    cites = CITE.findall(text)
    if doc in ("include/tqgpu.h", "DESIGN.md", "INTEGRATION.md", "oracle/triqs_oracle.py"):
        assert cites, doc
    bad = []
    for rel, line in cites:
        path = os.path.join(REF, rel)
        if not os.path.isfile(path) or int(line) > _line_count(path):
            bad.append(f"{rel}:{line}")
    assert not bad, bad
