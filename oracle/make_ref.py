"""Recipe for oracle/_ref/: a runnable copy of the UNMODIFIED reference (BASELINE.json configs[0]).

Test infrastructure only.  oracle/_ref/ is git-ignored (reference sources never enter the history) but travels to the
GPU box with the gpurun snapshot, where /root/reference does not exist.  Run in the build container:

    python oracle/make_ref.py          # also called by __graft_entry__.build() when /root/reference is present

What it does
  1. copies /root/reference/IntEL/src (*.py only) and IntEL/data/Tmall_toy to oracle/_ref/IntEL/{src,data};
  2. synthesises the one input file the reference ships without (`.MISSING_LARGE_BLOBS`: data/Tmall_toy/intents_multi.json,
     format of BaseReader._read_intent, BaseReader.py:102-109): key = str(c_id_c), value = I floats, I = 3 * n_class
     (IntEL.py:226 needs I divisible by model_num with I / model_num > the largest class id); the vector of a session is
     the normalised histogram of behaviour * n_class + i_class_c over its positive items (pay, fav, click in list order,
     BaseModel.py:177-185); a session without positives gets the uniform vector.  Deterministic: no RNG involved;
  3. drops the reference-side stubs of INTEGRATION.md (integration/ref_stubs/, OUR files) next to the classes they stand
     in for, so that `main.py --model_name IntEL_b200 --loss_name IntListloss_b200 --runner_name BaseRunner_b200` resolves.
"""
from __future__ import annotations

import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/IntEL"
OUT = os.path.join(ROOT, "oracle", "_ref", "IntEL")
STUBS = os.path.join(ROOT, "integration", "ref_stubs")
BEHAVIOURS = 3


def synthesise_intents(data_dir: str) -> int:
    import pandas as pd
    items = json.load(open(os.path.join(data_dir, "item_metadata.json")))
    cls = {int(k): int(v["i_class_c"]) for k, v in items.items()}
    n_class = max(max(cls.values()) + 1, len(set(cls.values()) | {0}))       # BaseReader._read_meta's itemfnum
    I = BEHAVIOURS * n_class
    out = {}
    for phase in ("train", "dev", "test"):
        df = pd.read_csv(os.path.join(data_dir, phase + ".csv"), sep="\t")
        for cid, iids, pay, fav, click in zip(df["c_id_c"], df["i_id_s"], df["c_paynum_i"], df["c_favnum_i"], df["c_clicknum_i"]):
            ids = json.loads(iids)
            beh = [2] * int(pay) + [1] * int(fav) + [0] * int(click)
            v = [0.0] * I
            for b, iid in zip(beh, ids):
                v[b * n_class + cls[int(iid)]] += 1.0
            s = sum(v)
            out[str(int(cid))] = [x / s for x in v] if s > 0 else [1.0 / I] * I
    with open(os.path.join(data_dir, "intents_multi.json"), "w") as f:
        json.dump(out, f)
    return I


def make_ref(force: bool = False) -> str | None:
    if not os.path.isdir(REF):
        return OUT if os.path.isdir(os.path.join(OUT, "src")) else None
    stamp = os.path.join(OUT, ".complete")
    if os.path.exists(stamp) and not force:
        _copy_stubs()
        return OUT
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    shutil.copytree(os.path.join(REF, "src"), os.path.join(OUT, "src"),
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc", ".DS_Store"))
    shutil.copytree(os.path.join(REF, "data", "Tmall_toy"), os.path.join(OUT, "data", "Tmall_toy"))
    for dp, _, fs in os.walk(OUT):                                           # the reference tree is read-only
        os.chmod(dp, 0o755)
        for f in fs:
            os.chmod(os.path.join(dp, f), 0o644)
    I = synthesise_intents(os.path.join(OUT, "data", "Tmall_toy"))
    _copy_stubs()
    with open(stamp, "w") as f:
        f.write(f"intent_num={I}\n")
    return OUT


def _copy_stubs() -> None:
    for dp, _, fs in os.walk(STUBS):
        rel = os.path.relpath(dp, STUBS)
        for f in fs:
            if f.endswith(".py"):
                dst = os.path.join(OUT, "src", rel, f)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(os.path.join(dp, f), dst)


if __name__ == "__main__":
    print(make_ref(force="-f" in sys.argv))
