"""Docs stay honest: every test, source file and profile that DESIGN.md / INTEGRATION.md / README.md name exists."""
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
DOCS = ["DESIGN.md", "INTEGRATION.md", "README.md"]


def _all_test_names():
    names = set()
    for f in (ROOT / "tests").glob("test_*.py"):
        names.add(f.name)
        names.update(re.findall(r"^def (test_\w+)", f.read_text(), flags=re.M))
    return names


def test_named_tests_exist():
    have = _all_test_names()
    for doc in DOCS:
        text = (ROOT / doc).read_text()
        for name in set(re.findall(r"`(?:tests/)?(test_\w+?(?:\.py)?)(?:::\w+)?`", text)) | set(re.findall(r"::(test_\w+)", text)):
            if name.endswith("_*") or name.endswith("_"):
                continue
            assert name in have or any(h.startswith(name.rstrip("*")) for h in have), f"{doc} names a test that does not exist: {name}"


def test_named_files_exist():
    pat = re.compile(r"`((?:profiles|tools|examples|oracle|tests|include|tuatara_b200)/[\w./\-]+\.(?:md|json|py|cpp|cu|cuh|h|sh|npz|gz))`")
    for doc in DOCS:
        for rel in set(pat.findall((ROOT / doc).read_text())):
            assert (ROOT / rel).exists(), f"{doc} names a file that does not exist: {rel}"
