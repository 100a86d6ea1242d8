"""CPU: the REAL reference engine (OfflineTrackingEngine + TrackerState + Pipeline.validate, unmodified, from /root/reference or
the staged oracle/_ref/) drives the drop-in ByteTrack module — its device tracker swapped for the NumPy oracle behind the same
interface so that no GPU is needed — and reproduces tests/golden/engine_bytetrack_2videos.npz, which the real engine produced
with the reference's own wrapper. Checks the module protocol (first base class = the real ImageLevelModule, level, datapipe /
dataloader overrides, process() row contract, the process-global id counter across two videos) on the real engine."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_reference():
    return os.path.isdir("/root/reference/tracklab") or os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "tracklab"))


def run_driver(*args, timeout=900):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "real_engine_driver.py"), *args], capture_output=True, text=True,
                       timeout=timeout, cwd=ROOT)
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("RESULT ")]
    assert p.returncode == 0 and lines, (p.stdout[-2000:], p.stderr[-4000:])
    return json.loads(lines[-1][7:])


@pytest.mark.skipif(not _have_reference(), reason="neither /root/reference nor the staged oracle/_ref is present")
def test_real_engine_drives_the_dropin_module_protocol():
    r = run_driver("bytetrack", "--cpu-standin")
    assert r["real_base_class"] and r["index_equal"] and r["ids_equal"] and r["max_box_err"] == 0.0 and r["with_track"] > 1000
