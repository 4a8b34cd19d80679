"""TEST INFRASTRUCTURE ONLY -- builds oracle/_ref/tokenizer_ref from the REFERENCE's own source.

The only part of /root/reference that compiles in this image is `class Tokenizer` (tuatara.cpp:25-117): it needs nothing
but LibTorch, and the torch wheel ships LibTorch's headers and libraries.  Everything else needs the OpenCV C++ SDK
(CMakeLists.txt:9), which the image does not have.  The recipe cuts those lines out of the read-only reference tree where
they lie into oracle/_ref/ (git-ignored: no reference source enters the history) and compiles oracle/tokenizer_ref_main.cpp
around them.  `python -m oracle.build_ref` ; __graft_entry__.build() calls build() when /root/reference exists.
"""
from __future__ import annotations

import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent
REF_SRC = Path("/root/reference/tuatara.cpp")
OUT = ROOT / "_ref"
FIRST, LAST = 25, 117  # class Tokenizer { ... };


def binary() -> Path:
    return OUT / "tokenizer_ref"


def build(force: bool = False) -> Path | None:
    """Returns the binary, or None when the reference tree is absent (the GPU box) and nothing was prebuilt."""
    exe = binary()
    if not REF_SRC.exists():
        return exe if exe.exists() else None
    main = ROOT / "tokenizer_ref_main.cpp"
    if exe.exists() and not force and exe.stat().st_mtime > max(main.stat().st_mtime, Path(__file__).stat().st_mtime):
        return exe
    OUT.mkdir(exist_ok=True)
    lines = REF_SRC.read_text().splitlines(keepends=True)
    cut = "".join(lines[FIRST - 1:LAST])
    assert cut.lstrip().startswith("class Tokenizer") and cut.rstrip().endswith("};"), "reference layout changed"
    (OUT / "tokenizer_class.inc").write_text(cut)
    import torch
    from torch.utils import cpp_extension

    tdir = Path(torch.__file__).resolve().parent
    cmd = ["g++", "-O1", "-std=c++17", f"-I{ROOT}", *[f"-I{p}" for p in cpp_extension.include_paths()],
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}", str(main), "-o", str(exe),
           f"-L{tdir / 'lib'}", "-ltorch", "-ltorch_cpu", "-lc10", f"-Wl,-rpath,{tdir / 'lib'}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("tokenizer_ref build failed:\n" + r.stderr[-4000:])
    return exe


def table() -> tuple[str, int, int, int]:
    out = subprocess.run([str(binary()), "table"], capture_output=True, text=True, check=True).stdout.split()
    return bytes.fromhex(out[4]).decode("latin-1"), int(out[1]), int(out[2]), int(out[3])


def decode(logits) -> list[str]:
    """logits: float32 [N, L, C] -> the reference's strings (softmax :486, Tokenizer::decode :492, cut at ']' :495-502)."""
    import numpy as np

    a = np.ascontiguousarray(logits, np.float32)
    n, l, c = a.shape
    out = subprocess.run([str(binary()), "decode", str(n), str(l), str(c)], input=a.tobytes(), capture_output=True, check=True).stdout
    return [bytes.fromhex(x).decode("latin-1") for x in out.decode().split("\n")[:n]]


if __name__ == "__main__":
    print(build(force=True))
