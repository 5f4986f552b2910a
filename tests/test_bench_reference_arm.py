"""bench.py --impl reference on the CPU: the JSON contract of the reference arm, the `config` object it shares with the GPU
arm, and that the arm loads nothing of this repo's product (only oracle/_ref, the reference compiled in place)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import json, runpy, sys
sys.argv = ["bench.py", "--impl", "reference", "--steps", "1", "--warmup", "1", "--workload", "c2", "--gpus", "4"]
try:
    runpy.run_path(%r, run_name="__main__")
except SystemExit:
    pass
libs = sorted({l.split()[-1] for l in open("/proc/self/maps") if %r in l and ".so" in l})
print("LIBS " + json.dumps(libs))
"""


def test_reference_arm_line_and_libraries(po):
    if po.ref() is None:
        pytest.skip("oracle/_ref/libref_oracle.so not present")
    bench = os.path.join(ROOT, "bench.py")
    out = subprocess.run([sys.executable, "-c", SCRIPT % (bench, ROOT)], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    line = json.loads([l for l in lines if l.startswith("{")][-1])
    libs = json.loads([l for l in lines if l.startswith("LIBS ")][-1][5:])
    assert line["impl"] == "reference" and line["unit"] == "Mrays/s" and line["higher_is_better"] is True
    assert line["n_gpus"] == 4 and line["value"] > 0 and line["cpu_baseline"]["kind"] == "reference"
    assert line["e2e"] == {"value": line["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["cores"] >= 1 and "every 16-th row" in line["cpu_baseline"]["sample"]
    # the same object the GPU arm prints for the same command line
    sys.path.insert(0, ROOT)
    import bench as b
    width, height, nss, _, desc = b.WORKLOADS["c2"]
    want = b.make_config(desc, "sibenik_standin", line["config"]["triangles"], width * 2, height * 2, 4, "auto")
    assert line["config"] == want and "reference_arm_sample" in want
    # nothing of the product on this path
    assert libs and all("oracle/_ref" in p for p in libs), libs
