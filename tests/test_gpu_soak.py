"""Short randomised soak (scripts/soak.py): single-GPU, batch and 3-rank distributed paths against the oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_soak_20s():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "soak.py"), "20", "11"], capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "soak ok" in out.stdout
