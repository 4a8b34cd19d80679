"""Offline converter: the reference's weight files -> the engine's flat ``craft.ttw`` / ``parseq.ttw``.

The reference's ``weights_dir`` holds two TorchScript archives, ``craft_traced_torchscript_model.pt`` and
``parseq_torchscript.bin`` (tuatara.cpp:333, :423; fetched from HuggingFace by setup.sh:6).  LibTorch is only needed
here, once, never on the OCR path:

    python -m tuatara_b200.convert --weights-dir ../weights            # converts in place, next to the .pt/.bin
    python -m tuatara_b200.convert --craft a.pt --parseq b.bin --out d

Plain ``torch.save``d checkpoints / state_dicts work as well.  Parameter names are normalised to the upstream ones
the exporter expects (clovaai/CRAFT-pytorch ``craft.py``, baudm/parseq ``PARSeq``): wrapper prefixes such as
``module.`` (DataParallel) and ``model.`` (the Lightning system around PARSeq) are stripped.
"""
from __future__ import annotations

import argparse
from pathlib import Path

import torch

from . import weights

CRAFT_FILE = "craft_traced_torchscript_model.pt"   # tuatara.cpp:333
PARSEQ_FILE = "parseq_torchscript.bin"              # tuatara.cpp:423


def load_state_dict(path: str | Path, unsafe_pickle: bool = False) -> dict:
    """state_dict of a TorchScript archive or a pickled (possibly nested) state_dict.  Plain pickles are read with
    ``weights_only=True`` (tensors and containers only); a file that needs arbitrary unpickling -- a pickled nn.Module,
    i.e. code execution on load -- is refused unless ``unsafe_pickle`` (CLI: --unsafe-pickle) says the file is trusted."""
    path = str(path)
    try:
        return dict(torch.jit.load(path, map_location="cpu").state_dict())
    except (RuntimeError, ValueError):
        pass
    try:
        obj = torch.load(path, map_location="cpu", weights_only=True)
    except Exception as exc:  # noqa: BLE001  (pickle.UnpicklingError and friends)
        if not unsafe_pickle:
            raise ValueError(f"{path}: not a TorchScript archive or a tensors-only pickle ({exc}); "
                             "pass unsafe_pickle=True / --unsafe-pickle only for files you trust") from exc
        obj = torch.load(path, map_location="cpu", weights_only=False)
    if hasattr(obj, "state_dict"):
        return dict(obj.state_dict())
    for key in ("state_dict", "model", "net"):
        if isinstance(obj, dict) and isinstance(obj.get(key), dict):
            obj = obj[key]
    if not isinstance(obj, dict):
        raise ValueError(f"{path}: neither a TorchScript archive nor a state_dict")
    return dict(obj)


def normalise(sd: dict, anchor: str) -> dict:
    """Strips whatever wrapper prefix precedes the key `anchor` (e.g. 'module.', 'model.') from every key."""
    hit = next((k for k in sd if k.endswith(anchor)), None)
    if hit is None:
        raise KeyError(f"no parameter named *{anchor}: not the expected architecture")
    prefix = hit[: len(hit) - len(anchor)]
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def convert(craft_path, parseq_path, out_dir, unsafe_pickle: bool = False) -> str:
    out = Path(out_dir)
    out.mkdir(parents=True, exist_ok=True)
    weights.export_craft(normalise(load_state_dict(craft_path, unsafe_pickle), "basenet.slice1.0.weight"), out / "craft.ttw")
    weights.export_parseq(normalise(load_state_dict(parseq_path, unsafe_pickle), "encoder.patch_embed.proj.weight"),
                          out / "parseq.ttw")
    return str(out)


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--weights-dir", help=f"directory holding {CRAFT_FILE} and {PARSEQ_FILE}; output goes there too")
    ap.add_argument("--craft")
    ap.add_argument("--parseq")
    ap.add_argument("--out")
    ap.add_argument("--unsafe-pickle", action="store_true",
                    help="allow full unpickling (executes code in the file): only for checkpoints you trust")
    a = ap.parse_args(argv)
    if a.weights_dir:
        a.craft = a.craft or str(Path(a.weights_dir) / CRAFT_FILE)
        a.parseq = a.parseq or str(Path(a.weights_dir) / PARSEQ_FILE)
        a.out = a.out or a.weights_dir
    if not (a.craft and a.parseq and a.out):
        ap.error("give --weights-dir, or --craft, --parseq and --out")
    print(convert(a.craft, a.parseq, a.out, a.unsafe_pickle))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
