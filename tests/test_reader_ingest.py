"""Reader ingest: CSR parsing of the `*_s` list columns and the columnar store against the reference's
own `eval()` path - a restatement on a synthetic file in the Tmall-toy schema everywhere, and the unmodified
`utils.df2dict` on the bundled toy CSVs where the reference tree is present (the build container)."""
import importlib
import os
import sys
import types

import numpy as np
import pytest

from intel_sigir2023_b200 import reader

REF_SRC = "/root/reference/IntEL/src"
TOY = "/root/reference/IntEL/data/Tmall_toy"


def _write_synthetic(path, n=57, seed=0):
    rng = np.random.default_rng(seed)
    cols = ["u_id_c", "c_time_i", "c_pCTR_s", "c_pCVR_s", "c_pFVR_s", "i_id_s", "c_paynum_i", "c_favnum_i", "c_clicknum_i",
            "c_trueneg_i", "pos_num", "c_id_c"]
    with open(path, "w") as f:
        f.write("\t".join(cols) + "\n")
        for r in range(n):
            L = int(rng.integers(1, 30))
            lists = [[round(float(x), 6) for x in rng.normal(3, 6, L)] for _ in range(3)]
            items = [int(x) for x in rng.integers(1, 5000, L)]
            row = [int(rng.integers(1, 9)), int(rng.integers(901, 931))] + [str(l) for l in lists] + [str(items)] + \
                  [int(rng.integers(0, 3)) for _ in range(4)] + [int(rng.integers(1, 5)), r + 1]
            f.write("\t".join(str(x) for x in row) + "\n")


def _df2dict_restated(path, max_session_len=-1):
    """utils.df2dict (utils/utils.py:15-30) after BaseReader._read_inter's sort (BaseReader.py:53-55)"""
    import pandas as pd
    df = pd.read_csv(path, sep="\t")
    df.sort_values(by=["u_id_c", "c_time_i"], inplace=True)
    df.reset_index(drop=True, inplace=True)
    res = df.to_dict("list")
    for key in res:
        if key[-2:] == "_s":
            res[key] = [eval(str(x)) if max_session_len == -1 else eval(str(x))[:max_session_len] for x in res[key]]
        else:
            res[key] = np.array(res[key])
            if key[-2:] == "_c":
                res[key] = res[key].astype(int)
    return res


def _assert_same(cols, ref):
    for key, v in ref.items():
        if key.endswith("_s"):
            got = reader.csr_to_lists(cols[key + ".values"], cols[key + ".offsets"])
            assert got == v, key                 # exact: same doubles, same ints, same lengths
        elif key != "session_len":
            assert cols[key].dtype == v.dtype and np.array_equal(cols[key], v), key


@pytest.mark.parametrize("cut", [-1, 7])
def test_read_inter_matches_the_eval_path(tmp_path, cut):
    p = str(tmp_path / "train.csv")
    _write_synthetic(p)
    cols, ref = reader.read_inter(p, max_session_len=cut), _df2dict_restated(p, cut)
    _assert_same(cols, ref)
    assert np.array_equal(cols["session_len"], [len(eval(str(x))) for x in _df2dict_restated(p)["i_id_s"]])
    assert cols["i_id_s.values"].dtype == np.int64 and cols["c_pCTR_s.values"].dtype == np.float64


def test_parse_list_column_edge_cases():
    v, o = reader.parse_list_column(["[1, 2, 3]", "[]", " [4.5] ", "[-1e-3,7]"])
    assert v.tolist() == [1.0, 2.0, 3.0, 4.5, -0.001, 7.0] and o.tolist() == [0, 3, 3, 4, 6]
    v, o = reader.parse_list_column(["[1, 2, 3]", "[]", "[4, 5]"], dtype=np.int64, max_len=2)
    assert v.tolist() == [1, 2, 4, 5] and o.tolist() == [0, 2, 2, 4]
    v, o = reader.parse_list_column([])
    assert v.size == 0 and o.tolist() == [0]
    with pytest.raises(ValueError):
        reader.parse_list_column(["1, 2"])
    with pytest.raises(ValueError):
        reader.parse_list_column(["[1.5]"], dtype=np.int64)


def test_columnar_round_trip(tmp_path):
    p = str(tmp_path / "dev.csv")
    _write_synthetic(p, n=23, seed=3)
    cols = reader.read_inter(p)
    reader.save_columnar(cols, str(tmp_path / "dev.col"))
    back = reader.load_columnar(str(tmp_path / "dev.col"))
    assert set(back) == set(cols)
    for k, v in cols.items():
        assert back[k].dtype == v.dtype and np.array_equal(back[k], v), k


@pytest.mark.skipif(not os.path.isdir(TOY), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("phase", ["dev", "test"])
def test_toy_files_match_the_unmodified_df2dict(phase):
    import pandas as pd
    np.object, np.float, np.int = object, float, int
    imp = types.ModuleType("imp")
    imp.reload = importlib.reload
    sys.modules.setdefault("imp", imp)
    sys.path.insert(0, REF_SRC)
    try:
        from utils import utils as ref_utils
    finally:
        sys.path.remove(REF_SRC)
    path = os.path.join(TOY, phase + ".csv")
    df = pd.read_csv(path, sep="\t")
    df.sort_values(by=["u_id_c", "c_time_i"], inplace=True)
    df.reset_index(drop=True, inplace=True)
    ref = ref_utils.df2dict(df)
    _assert_same(reader.read_inter(path), ref)
