"""Generates tests/golden/tokenizer_ref.json with the REFERENCE's own tokenizer (oracle/_ref/tokenizer_ref, compiled from
/root/reference/tuatara.cpp:25-117 by oracle/build_ref.py).  Run here (the reference tree is not on the GPU box):
    python tests/golden/make_golden_tokenizer.py
Inputs are seeded, so only the seed, the planted argmax ids and the reference's answers are stored."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import build_ref  # noqa: E402


def make_logits(seed: int, n: int = 96, L: int = 26, C: int = 95):
    """Peaky logits with every class present, EOS (class 0) planted at random positions, class 88 (the reference's
    `eos_id`, silently dropped) and the shifted punctuation classes 69..94 over-represented."""
    rng = np.random.default_rng(seed)
    ids = rng.integers(0, C, size=(n, L))
    special = rng.random((n, L))
    ids = np.where(special < 0.10, 88, ids)
    ids = np.where((special >= 0.10) & (special < 0.25), rng.integers(69, 95, size=(n, L)), ids)
    for i in range(n):
        if i % 4 != 3:  # three in four items end with an EOS somewhere
            ids[i, rng.integers(0, L)] = 0
    ids[0, :] = np.arange(L) % C            # no EOS at all
    ids[1, 0] = 0                           # EOS first -> empty string
    ids[2, :] = 88                          # everything filtered -> empty string
    logits = rng.standard_normal((n, L, C)).astype(np.float32)
    np.put_along_axis(logits, ids[..., None], 9.0, axis=-1)
    return logits, ids.astype(np.int64)


if __name__ == "__main__":
    exe = build_ref.build()
    assert exe is not None, "needs /root/reference"
    itos, eos, bos, pad = build_ref.table()
    seed = 20261017
    logits, ids = make_logits(seed)
    assert (logits.argmax(-1) == ids).all()
    strings = build_ref.decode(logits)
    out = dict(source="/root/reference/tuatara.cpp:25-117 compiled by oracle/build_ref.py", seed=seed, shape=list(logits.shape),
               itos_hex=itos.encode("latin-1").hex(), eos_id=eos, bos_id=bos, pad_id=pad, ids=ids.tolist(),
               strings_hex=[s.encode("latin-1").hex() for s in strings])
    (ROOT / "tests" / "golden" / "tokenizer_ref.json").write_text(json.dumps(out))
    print(len(strings), "strings;", sum(1 for s in strings if not s), "empty; e.g.", strings[:6])
