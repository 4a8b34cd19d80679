"""bench.py's supervisor (the native arm runs in a child with a time budget and one conservative retry) exercised
without a GPU: a child that hangs is killed and retried, a child that fails is retried, only the JSON line is printed."""
import json
import os
import subprocess
import sys
import time
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def _run(mode, budget="3"):
    env = dict(os.environ, TT_BENCH_TEST_CHILD=mode, TT_BENCH_BUDGET_S=budget)
    env.pop("TT_BENCH_CHILD", None)
    t0 = time.time()
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "1"], env=env, capture_output=True,
                       text=True, timeout=120)
    return r, time.time() - t0


@pytest.mark.parametrize("mode", ["ok", "hang_once", "fail_once"])
def test_supervisor_retries_and_prints_one_json_line(mode):
    r, dt = _run(mode)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert len(lines) == 1, r.stdout  # the child's non-JSON stdout noise is dropped
    d = json.loads(lines[0])
    assert d["fake"] is True
    assert d["retry"] == (mode != "ok")
    if mode != "ok":
        assert d["slots"] == "1" and "retrying" in r.stderr  # the retry runs the conservative paths on one slot
    if mode == "hang_once":
        assert dt >= 3.0  # the first child was given its budget, then killed
