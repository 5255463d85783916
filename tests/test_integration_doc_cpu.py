"""INTEGRATION.md cites reference symbols as `Name`, `Name` (`path.py:line,line`): where the reference is present (build container:
/root/reference; GPU box: oracle/_ref) every cited line must define the symbol named at the same position of the row."""
import os
import re

import pytest

from oracle import ref_harness

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not ref_harness.reference_available(), reason="reference sources not present")
def test_every_cited_reference_symbol_exists_at_the_cited_line():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    table = text[text.index("| reference symbol (file:line) | drop-in |"):text.index("Stock-quantised families")]
    checked = 0
    for row in table.splitlines()[2:]:
        if not row.startswith("|"):
            continue
        left = row.split("|")[1]
        # groups:  `A`, `B` (`file.py:1,2`)
        for names, path, lines in re.findall(r"((?:`[A-Za-z_][A-Za-z_0-9]*`(?:, )?)+) \(`([^`]+?\.py):([0-9,]+)`\)", left):
            syms = re.findall(r"`([A-Za-z_][A-Za-z_0-9]*)`", names)
            nums = [int(n) for n in lines.split(",")]
            assert len(syms) == len(nums), "row cites %d symbols but %d lines: %s" % (len(syms), len(nums), row[:120])
            path = path.replace("…/bbb", "src/models/stochastic/bbb")
            src = open(os.path.join(ref_harness.REFERENCE_ROOT, path)).read().splitlines()
            for sym, n in zip(syms, nums):
                line = src[n - 1]
                assert re.match(r"\s*(class|def)\s+%s\b|\s*%s\s*=" % (sym, sym), line), "%s:%d is %r, not the definition of %s" % (path, n, line, sym)
                checked += 1
    assert checked >= 40
