"""CPU: the numpy emulations of kernel logic (scripts/emulation/) still agree with the oracle.  They are the
only evidence for the experimental kernels until those run on a GPU, and the design record of the prob_sample
scan, so they are kept running."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("script", ["sim_scan.py", "sim_grid.py", "sim_select.py"])
def test_emulation_agrees_with_oracle(script):
    if script == "sim_grid.py":  # needs pn2_ball_threshold from the built library
        import __graft_entry__ as g
        if not os.path.exists(os.path.join(ROOT, "open3d-pointnet2-semantic3d_b200", "lib", "libpn2_b200.so")):
            g.build()
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "emulation", script)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]
    assert "False" not in out.stdout
