"""Reader ingest for the path's inputs: the interaction CSVs of the reference (`BaseReader._read_inter`,
helpers/BaseReader.py:44-69) hold every per-session list (`*_s` columns: item ids, basic-model scores) as a Python list
literal that `utils.df2dict` (utils/utils.py:15-30) turns back into lists with one `eval()` per cell.  Here a list column
is parsed in one pass into CSR form (`values`, `offsets`) and the whole table is kept as flat numpy columns that can be
written to / memory-mapped from a columnar directory, so start-up cost is one `np.load(mmap_mode='r')` per column.

Host code only (numpy / pandas); nothing here touches the GPU library."""
from __future__ import annotations

import json
import os
from typing import Dict, Iterable, List, Tuple

import numpy as np


def parse_list_column(cells: Iterable[str], dtype=np.float64, max_len: int = -1) -> Tuple[np.ndarray, np.ndarray]:
    """["[1, 2, 3]", "[]", "[4.5]"] -> (values [nnz], offsets int64 [n+1]); same numbers as `eval(str(x))` per cell
    (decimal literals are parsed correctly rounded on both sides).  `max_len` > -1 keeps the first `max_len` entries of each
    list, like `eval(str(x))[:max_session_len]` (utils.py:24)."""
    cells = [str(c).strip() for c in cells]
    inner = []
    for c in cells:
        if not (c.startswith("[") and c.endswith("]")):
            raise ValueError(f"not a list literal: {c[:40]!r}")
        inner.append(c[1:-1].strip())
    counts = np.fromiter((0 if not s else s.count(",") + 1 for s in inner), dtype=np.int64, count=len(inner))
    flat = ",".join(s for s in inner if s)
    values = np.array(flat.split(","), dtype=np.float64) if flat else np.zeros(0, dtype=np.float64)
    if values.size != int(counts.sum()):
        raise ValueError("nested or malformed list literal")
    offsets = np.zeros(len(inner) + 1, dtype=np.int64)
    np.cumsum(counts, out=offsets[1:])
    if max_len > -1:
        keep = np.minimum(counts, max_len)
        idx = np.repeat(offsets[:-1], keep) + (np.arange(int(keep.sum())) - np.repeat(np.cumsum(keep) - keep, keep))
        values = values[idx]
        offsets = np.zeros(len(inner) + 1, dtype=np.int64)
        np.cumsum(keep, out=offsets[1:])
    if np.issubdtype(np.dtype(dtype), np.integer):
        as_int = values.astype(dtype)
        if not np.array_equal(as_int, values):
            raise ValueError("non-integer entries in an integer list column")
        values = as_int
    return values.astype(dtype, copy=False), offsets


def csr_to_lists(values: np.ndarray, offsets: np.ndarray) -> List[list]:
    """the reference's in-memory form (list of Python lists) of a CSR column"""
    v = values.tolist()
    return [v[offsets[i]:offsets[i + 1]] for i in range(len(offsets) - 1)]


def read_inter(path: str, sep: str = "\t", max_session_len: int = -1) -> Dict[str, np.ndarray]:
    """One phase file (train/dev/test .csv) -> flat columns.  Rows are ordered like `_read_inter` (sorted by user, then time;
    same pandas calls, BaseReader.py:54-55).  Scalar columns keep df2dict's dtypes (`*_c` -> int, others as pandas read
    them); a list column `<name>_s` becomes `<name>_s.values` + `<name>_s.offsets` (item-id lists int64, the rest
    float64); `session_len` is added as in BaseReader.py:61-66 (length before any `max_session_len` cut)."""
    import pandas as pd
    df = pd.read_csv(path, sep=sep)
    df.sort_values(by=["u_id_c", "c_time_i"], inplace=True)
    df.reset_index(drop=True, inplace=True)
    out: Dict[str, np.ndarray] = {}
    for key in df.columns:
        if key.endswith("_s"):
            dt = np.int64 if key == "i_id_s" else np.float64
            full_v, full_o = parse_list_column(df[key].tolist(), dtype=dt)
            if key == "i_id_s":
                out["session_len"] = np.diff(full_o)
            if max_session_len > -1:
                full_v, full_o = parse_list_column(df[key].tolist(), dtype=dt, max_len=max_session_len)
            out[key + ".values"], out[key + ".offsets"] = full_v, full_o
        else:
            col = df[key].to_numpy()
            out[key] = col.astype(int) if key.endswith("_c") else col
    return out


def save_columnar(columns: Dict[str, np.ndarray], directory: str) -> None:
    """one .npy per column + a manifest; `load_columnar` memory-maps them"""
    os.makedirs(directory, exist_ok=True)
    for k, v in columns.items():
        np.save(os.path.join(directory, k + ".npy"), np.ascontiguousarray(v))
    with open(os.path.join(directory, "manifest.json"), "w") as f:
        json.dump({k: [str(v.dtype), list(v.shape)] for k, v in columns.items()}, f)


def load_columnar(directory: str, mmap: bool = True) -> Dict[str, np.ndarray]:
    with open(os.path.join(directory, "manifest.json")) as f:
        manifest = json.load(f)
    return {k: np.load(os.path.join(directory, k + ".npy"), mmap_mode="r" if mmap else None) for k in manifest}
