"""The ctypes stub printed in INTEGRATION.md must actually run (GPU box)."""
import os
import re

import pytest

from conftest import ROOT, has_cuda

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")]


def test_integration_md_ctypes_stub_runs():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stub = [b for b in blocks if "ovl_overlap_dp" in b and "ctypes.CDLL" in b]
    assert len(stub) == 1
    cwd = os.getcwd()
    os.chdir(ROOT)
    try:
        exec(compile(stub[0], "INTEGRATION.md:ctypes-stub", "exec"), {"__name__": "integration_stub"})
    finally:
        os.chdir(cwd)
