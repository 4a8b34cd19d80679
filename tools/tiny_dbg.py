import os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import tuatara_b200 as tb
from oracle.models import make_parseq, make_craft
from tuatara_b200 import weights
d = ROOT / "tests" / "_cache" / "weights_tiny_seed0"
d.mkdir(parents=True, exist_ok=True)
if not (d / "craft.ttw").exists():
    weights.export_craft(make_craft(0).state_dict(), d / "craft.ttw")
if not (d / "parseq.ttw").exists():
    weights.export_parseq(make_parseq("tiny", 0).state_dict(), d / "parseq.ttw")
eng = tb.Engine(str(d), devices=[0])
crops = np.random.default_rng(0).integers(0, 256, (60, 32, 128, 3), dtype=np.uint8)
for lnf in ("1", "0"):
    for dec in ("1", "0"):
        for ar in ("0", "1"):
            os.environ.update(TT_ENC_LNFUSE=lnf, TT_DEC_FUSED=dec, TT_PARSEQ_AR_LOGITS=ar)
            l, ids = eng.parseq_forward(crops)
            print(f"lnfuse={lnf} decfused={dec} ar={ar}: finite {np.isfinite(l).mean():.3f} absmax {np.nanmax(np.abs(l)):.3f}")
