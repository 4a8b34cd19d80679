"""The C++ example callers (examples/resume.cpp, examples/table.cpp: the reference's examples/*.cpp on the drop-in
header) build against include/tuatara.h + the C-ABI library, fail softly like the reference on a missing image, and
on a GPU print exactly what the engine returns through the Python binding."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
EX = ROOT / "examples"


@pytest.fixture(scope="module")
def examples(native_lib):
    r = subprocess.run(["make", "-C", str(EX), "-B"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return EX / "resume", EX / "table"


def _write_ppm(path, rgb):
    with open(path, "wb") as f:
        f.write(b"P6\n# synthetic page\n%d %d\n255\n" % (rgb.shape[1], rgb.shape[0]))
        f.write(np.ascontiguousarray(rgb).tobytes())


def test_examples_build_and_fail_softly(examples, tmp_path):
    resume, table = examples
    r = subprocess.run([str(resume)], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr
    r = subprocess.run([str(resume), str(tmp_path / "missing.ppm"), "w", str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 1 and "Error reading image from file" in r.stderr  # tuatara.cpp:344-347
    r = subprocess.run([str(table), str(tmp_path / "missing.ppm")], capture_output=True, text=True)
    assert r.returncode == 1


@pytest.mark.gpu
def test_resume_example_matches_python_binding(examples, engine, weights_dir, tmp_path):
    from tuatara_b200 import synth

    resume, _ = examples
    rgb = np.ascontiguousarray(synth.synth_page(3)[:640, :768])
    _write_ppm(tmp_path / "page.ppm", rgb)
    r = subprocess.run([str(resume), str(tmp_path / "page.ppm"), weights_dir, str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = [ln.split("\t") for ln in r.stdout.splitlines()]
    bgr = np.ascontiguousarray(rgb[..., ::-1])  # the example hands the page over as BGR, like cv::imread
    ref = engine.ocr_pages([bgr])[0]
    assert len(got) == len(ref)
    for g, x in zip(got, ref):
        assert g[0] == x["text"] and [float(v) for v in g[1:]] == list(x["bbox"])
    assert (tmp_path / "annotated.ppm").exists()
