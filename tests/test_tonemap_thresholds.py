"""The threshold form of the tone-map quantiser (vulkanpbrt_b200/csrc/common.cuh: tonemap_code) rests on one fact:
unorm8(clamp(vk_pow(x, .454545), 0, 1)) is monotone in x.  tests/tools/pow_unorm8_sweep.c walks every non-negative
binary32 through the ORACLE's vk_pow / f32_to_unorm8, checks monotonicity and regenerates the threshold table; the
committed table (csrc/tonemap_thresholds.inc) must be exactly what it produces."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.slow
def test_tonemap_quantiser_is_monotone_and_table_is_current(tmp_path):
    exe, out = tmp_path / "pow_sweep", tmp_path / "thr.inc"
    subprocess.run(["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-o", str(exe), str(ROOT / "tests" / "tools" / "pow_unorm8_sweep.c"), "-lm"],
                   check=True)
    r = subprocess.run([str(exe), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "monotonicity violations: 0, codes never produced: 0" in r.stdout
    assert out.read_text() == (ROOT / "vulkanpbrt_b200" / "csrc" / "tonemap_thresholds.inc").read_text()
