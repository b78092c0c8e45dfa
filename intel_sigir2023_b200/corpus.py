"""Columnar corpus and device-side batch construction.

The reference keeps its corpus as Python dicts of lists (`BaseReader` / `SeqReader`, helpers/BaseReader.py:44-110,
helpers/SeqReader.py:18-60) and builds every batch on the host, one `_get_feed_dict` per session (BaseModel.py:158-197,
GeneralSeq.py:35-54, IntEL.py:220-239) followed by `collate_batch` (BaseModel.py:121-142): the dense float64
`his_intents` / `his_item_int` tensors alone are 2*H*I*8 bytes per session.  Here

* `ColumnarCorpus` is the Reader: it reads the same files with `reader.read_inter` (one-pass CSR parsing of the list
  columns), applies the same sort / merge / history rules, and exposes the attributes the model constructor reads
  (`max_uid`, `max_iid`, `itemfnum`, `contextfnum`, `zero_int`, ...) next to flat numpy columns: CSR item lists with
  precomputed rankings and min-max normalised scores, per-user session / item histories, CSR intent vectors;
* `DeviceCorpus` uploads those columns once and builds each batch with one kernel launch (`intel_batch_build`) from
  the batch's row indices (and, optionally, the per-session list permutation): the host never materialises a batch,
  the history intents are emitted in the compact (index, value) layout the model kernels read directly.

A device-built batch equals the reference's `collate_batch` dict field by field (tests/test_device_corpus.py).
"""
from __future__ import annotations

import ctypes as C
import json
import os
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib, reader

PHASES = ("train", "dev", "test")
POS_TYPES = ["c_paynum_i", "c_favnum_i", "c_clicknum_i"]          # BaseReader.py:36
BASIC_SCORES = ["c_pCTR_s", "c_pCVR_s", "c_pFVR_s"]                # BaseReader.py:37


class ColumnarCorpus:
    """Host-side columnar corpus.  `phase[p]` holds the per-row and CSR columns of one phase, in the reference's row
    order (sorted by user, then time: BaseReader.py:54-55)."""

    cfeatures = ["c_time_i"]
    ifeatures = ["i_class_c"]
    ufeatures = ["u_age_c", "u_gender_c"]
    pos_types = POS_TYPES
    basic_scores = BASIC_SCORES

    def __init__(self, datapath: str, dataset: str, max_session_len: int = 40, intent_note: str = "", history_max: int = 20,
                 model_num: int = 3, sep: str = "\t"):
        root = os.path.join(datapath, dataset)
        self.max_session_len, self.history_max, self.model_num = max_session_len, history_max, model_num
        raw = {p: reader.read_inter(os.path.join(root, p + ".csv"), sep=sep) for p in PHASES}
        # ---- BaseReader._read_inter: table sizes over all phases, full (uncut) lists ----
        self.max_uid = int(max(r["u_id_c"].max() for r in raw.values()))
        self.max_iid = int(max(r["i_id_s.values"].max() for r in raw.values()))
        ctx_vals = set([0])
        for r in raw.values():
            ctx_vals |= set(int(x) for x in np.unique(r["c_time_i"]))
        self.contextfnum = [max(len(ctx_vals), max(ctx_vals) + 1)]
        # ---- BaseReader._read_meta ----
        items = json.load(open(os.path.join(root, "item_metadata.json")))
        self.item_class = np.zeros(self.max_iid + 1, dtype=np.int64)
        fs = set([0])
        for k, v in items.items():
            if int(k) <= self.max_iid:
                self.item_class[int(k)] = int(v["i_class_c"])
            fs.add(int(v["i_class_c"]) - 1)
        self.itemfnum = [max(len(fs), max(fs) + 1)]
        users = json.load(open(os.path.join(root, "user_metadata.json")))
        ufs = [set([0]) for _ in self.ufeatures]
        for v in users.values():
            for i, f in enumerate(self.ufeatures):
                ufs[i].add(int(v[f]))
        self.userfnum = [max(len(f), max(f) + 1) for f in ufs]
        self.user_mh = np.zeros(self.max_uid + 1, dtype=np.int64)
        for k, v in users.items():
            if int(k) <= self.max_uid:
                mh = 0
                for i, f in enumerate(self.ufeatures):
                    mh = mh * self.userfnum[i] + int(v[f])
                self.user_mh[int(k)] = mh
        # ---- BaseReader._read_intent: CSR over the sessions of the file; row 0 is the all-zero vector ----
        intents = json.load(open(os.path.join(root, "intents%s.json" % intent_note)))
        self.intent_num = len(next(iter(intents.values())))
        self.zero_int = np.zeros(self.intent_num)
        cid_row: Dict[int, int] = {}
        off, idx, val = [0, 0], [], []
        for key, vec in intents.items():
            v = np.asarray(vec, dtype=np.float64)
            nz = np.nonzero(v)[0]
            cid_row[int(float(key))] = len(off) - 1
            idx.append(nz.astype(np.int32))
            val.append(v[nz])
            off.append(off[-1] + len(nz))
        self.int_off = np.asarray(off, dtype=np.int64)
        self.int_idx = np.concatenate(idx) if idx else np.zeros(0, np.int32)
        self.int_val = np.concatenate(val) if val else np.zeros(0, np.float64)
        self.nz1 = int(max(1, np.diff(self.int_off).max()))
        # ---- SeqReader._append_his_info: histories in (time, user) order, stable ----
        cat = {k: np.concatenate([raw[p][k] for p in PHASES]) for k in ("u_id_c", "c_time_i", "c_id_c", "c_clicknum_i", "c_paynum_i",
                                                                        "c_favnum_i")}
        full_off = np.concatenate([[0], np.cumsum(np.concatenate([np.diff(raw[p]["i_id_s.offsets"]) for p in PHASES]))])
        full_items = np.concatenate([raw[p]["i_id_s.values"] for p in PHASES])
        order = np.lexsort((cat["u_id_c"], cat["c_time_i"]))          # primary c_time_i, then u_id_c; lexsort is stable
        G = len(order)
        position = np.zeros(G, dtype=np.int64)
        item_position = np.zeros(G, dtype=np.int64)
        his_rows: List[List[int]] = [[] for _ in range(self.max_uid + 1)]
        his_ctx: List[List[int]] = [[] for _ in range(self.max_uid + 1)]
        it_ids: List[List[int]] = [[] for _ in range(self.max_uid + 1)]
        it_int: List[List[int]] = [[] for _ in range(self.max_uid + 1)]
        per_beh = self.intent_num / self.model_num                     # IntEL.py:226 (float division, then int())
        for g in order:
            uid, cid = int(cat["u_id_c"][g]), int(cat["c_id_c"][g])
            click, pay, fav = int(cat["c_clicknum_i"][g]), int(cat["c_paynum_i"][g]), int(cat["c_favnum_i"][g])
            position[g] = len(his_rows[uid])
            item_position[g] = len(it_ids[uid])
            his_rows[uid].append(cid_row.get(cid, -1))
            his_ctx[uid].append(int(cat["c_time_i"][g]))
            pos_items = full_items[full_off[g]:full_off[g] + click + pay + fav]
            beh = [0] * click + [1] * fav + [2] * pay
            it_ids[uid].extend(int(x) for x in pos_items)
            it_int[uid].extend(int(b * per_beh + self.item_class[int(i)]) for b, i in zip(beh, pos_items))
        self.uhis_off = np.concatenate([[0], np.cumsum([len(x) for x in his_rows])]).astype(np.int64)
        self.uhis_row = np.asarray([x for l in his_rows for x in l], dtype=np.int64)
        self.uhis_ctx = np.asarray([x for l in his_ctx for x in l], dtype=np.int64)
        self.uitem_off = np.concatenate([[0], np.cumsum([len(x) for x in it_ids])]).astype(np.int64)
        self.uitem_id = np.asarray([x for l in it_ids for x in l], dtype=np.int64)
        self.uitem_int = np.asarray([x for l in it_int for x in l], dtype=np.int32)
        if (self.uhis_row < 0).any():
            raise KeyError("a session that appears in a user history has no intent vector (GeneralSeq.py:43 indexes "
                           "corpus.intents directly)")
        # ---- per phase: cut lists (train only, BaseReader.py:72-76), rankings, normalised scores ----
        self.phase: Dict[str, Dict[str, np.ndarray]] = {}
        base = 0
        for p in PHASES:
            r = raw[p]
            n_rows = len(r["u_id_c"])
            cut = max_session_len if p == "train" else -1
            offs = r["i_id_s.offsets"]
            full_n = np.diff(offs)
            n = np.minimum(full_n, cut) if cut > -1 else full_n
            new_off = np.concatenate([[0], np.cumsum(n)]).astype(np.int64)
            take = np.repeat(offs[:-1], n) + (np.arange(int(n.sum())) - np.repeat(new_off[:-1], n))
            item_id = r["i_id_s.values"][take]
            scores = np.stack([r[s + ".values"][take] for s in BASIC_SCORES], axis=1)          # [nnz, K] raw
            seg = np.repeat(np.arange(n_rows), n)
            lo = np.full((n_rows, scores.shape[1]), np.inf)
            hi = np.full((n_rows, scores.shape[1]), -np.inf)
            np.minimum.at(lo, seg, scores)
            np.maximum.at(hi, seg, scores)
            scores = (scores - lo[seg]) / (hi[seg] - lo[seg] + 1e-6)                           # BaseModel.py:173
            # ranking: [3]*pay + [2]*fav + [1]*click + [0]*trueneg, -1 behind, cut to the list (BaseModel.py:177-185)
            slot = np.arange(int(n.sum())) - np.repeat(new_off[:-1], n)
            pay, fav, clk, neg = (r[k] for k in ("c_paynum_i", "c_favnum_i", "c_clicknum_i", "c_trueneg_i"))
            c1, c2, c3, c4 = pay[seg], (pay + fav)[seg], (pay + fav + clk)[seg], (pay + fav + clk + neg)[seg]
            ranking = np.where(slot < c1, 3, np.where(slot < c2, 2, np.where(slot < c3, 1, np.where(slot < c4, 0, -1)))).astype(np.int64)
            g = base + np.arange(n_rows)
            self.phase[p] = {
                "u_id_c": r["u_id_c"].astype(np.int64), "c_id_c": r["c_id_c"].astype(np.int64),
                "context_mh": r["c_time_i"].astype(np.int64), "user_mh": self.user_mh[r["u_id_c"]],
                "c_paynum_i": pay.astype(np.int64), "c_favnum_i": fav.astype(np.int64), "c_clicknum_i": clk.astype(np.int64),
                "session_len": n.astype(np.int64), "full_session_len": full_n.astype(np.int64),
                "position": position[g], "item_position": item_position[g],
                "intent_row": np.asarray([cid_row.get(int(c), 0) for c in r["c_id_c"]], dtype=np.int64),
                "item_off": new_off, "item_id": item_id.astype(np.int64), "item_class": self.item_class[item_id],
                "ranking": ranking, "scores": np.ascontiguousarray(scores, dtype=np.float64),
            }
            base += n_rows

    def __len__(self):
        return sum(len(v["u_id_c"]) for v in self.phase.values())


class _Corpus(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("N", "K", "I", "max_his")] + [("nz1", C.c_int32)] + \
               [(n, C.c_void_p) for n in ("u_id", "c_id", "context_mh", "user_mh", "pay", "fav", "click", "session_len", "position",
                                          "item_position", "intent_row", "item_off", "item_id", "item_class", "ranking", "scores",
                                          "uhis_off", "uhis_row", "uhis_ctx", "uitem_off", "uitem_id", "uitem_int",
                                          "int_off", "int_idx", "int_val")]


class _Built(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("u_id", "c_id", "context_mh", "user_mh", "pay", "fav", "click", "session_len", "position",
                                          "history_len", "history_item_len", "i_id", "i_class", "ranking", "scores", "intents",
                                          "his_context", "his_intents_idx", "his_intents_val", "his_item_id", "his_item_int_idx",
                                          "his_item_int_val")]


class DeviceCorpus:
    """The columns of a ColumnarCorpus (or of `synthetic_columns`) resident on one device + the batch builder."""

    def __init__(self, columns: Dict[str, np.ndarray], shared: Dict[str, np.ndarray], K: int, I: int, max_his: int, nz1: int,
                 device, phase_name: str = "train"):
        self.device = torch.device(device)
        self.K, self.I, self.max_his, self.nz1, self.phase_name = int(K), int(I), int(max_his), int(nz1), phase_name
        self.host = {k: np.ascontiguousarray(columns[k]) for k in ("session_len", "position", "item_position")}
        self.N = len(self.host["session_len"])
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
        self.col = {k: up(v) for k, v in columns.items() if k != "full_session_len"}
        self.shared = {k: up(v) for k, v in shared.items()}
        c = _Corpus()
        c.N, c.K, c.I, c.max_his, c.nz1 = self.N, self.K, self.I, self.max_his, self.nz1
        names = {"u_id": "u_id_c", "c_id": "c_id_c", "pay": "c_paynum_i", "fav": "c_favnum_i", "click": "c_clicknum_i"}
        for f in ("u_id", "c_id", "context_mh", "user_mh", "pay", "fav", "click", "session_len", "position", "item_position",
                  "intent_row", "item_off", "item_id", "item_class", "ranking", "scores"):
            setattr(c, f, _lib.ptr(self.col[names.get(f, f)]))
        for f in ("uhis_off", "uhis_row", "uhis_ctx", "uitem_off", "uitem_id", "uitem_int", "int_off", "int_idx", "int_val"):
            setattr(c, f, _lib.ptr(self.shared[f]))
        self._c = c

    @staticmethod
    def from_columnar(corpus: ColumnarCorpus, phase: str, device) -> "DeviceCorpus":
        shared = {k: getattr(corpus, k) for k in ("uhis_off", "uhis_row", "uhis_ctx", "uitem_off", "uitem_id", "uitem_int", "int_off",
                                                   "int_idx", "int_val")}
        return DeviceCorpus(corpus.phase[phase], shared, len(BASIC_SCORES), corpus.intent_num, corpus.history_max, corpus.nz1, device,
                            phase)

    def __len__(self):
        return self.N

    def shape_of(self, rows: np.ndarray):
        """(L, H1, H2) of the batch made of these rows: host lookups in three int64 columns"""
        n = self.host["session_len"][rows]
        h1 = np.clip(self.host["position"][rows], 1, self.max_his) if self.max_his > 0 else np.maximum(self.host["position"][rows], 1)
        h2 = np.clip(self.host["item_position"][rows], 1, self.max_his) if self.max_his > 0 else np.maximum(self.host["item_position"][rows], 1)
        return int(n.max()), int(h1.max()), int(h2.max())

    def random_perm(self, rows_dev: torch.Tensor, L: int, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        """a uniformly random order of the first n_b slots of every list (BaseModel.py:194-196), int32 [B, L]"""
        n = self.col["session_len"][rows_dev]
        key = torch.rand(rows_dev.numel(), L, device=self.device, generator=generator)
        key.masked_fill_(torch.arange(L, device=self.device)[None, :] >= n[:, None], 2.0)
        return key.argsort(dim=1).to(torch.int32)

    def batch(self, rows, perm: Optional[torch.Tensor] = None, shuffle: bool = True, shape=None) -> Dict[str, object]:
        """rows: int64 indices into the phase (numpy / CPU tensor: what a sampler yields).  perm: int32 [B, L] source slot
        of every output slot (shared with the reference in the parity test); None + shuffle -> drawn on the device."""
        rows_np = rows.numpy() if torch.is_tensor(rows) else np.asarray(rows, dtype=np.int64)
        B = len(rows_np)
        L, H1, H2 = shape if shape is not None else self.shape_of(rows_np)
        dev = self.device
        rows_host = torch.from_numpy(np.ascontiguousarray(rows_np, dtype=np.int64))
        rows_dev = (rows_host.pin_memory() if dev.type == "cuda" else rows_host).to(dev, non_blocking=True)
        if perm is None and shuffle:
            perm = self.random_perm(rows_dev, L)
        i64 = lambda *s: torch.empty(*s, dtype=torch.int64, device=dev)
        out = {
            "u_id_c": i64(B), "c_id_c": i64(B), "context_mh": i64(B), "user_mh": i64(B), "c_paynum_i": i64(B), "c_favnum_i": i64(B),
            "c_clicknum_i": i64(B), "session_len": i64(B), "position": i64(B), "history_len": i64(B), "history_item_len": i64(B),
            "i_id_s": i64(B, L), "i_class_c": i64(B, L), "ranking": i64(B, L),
            "scores": torch.empty(B, L, self.K, dtype=torch.float64, device=dev),
            "intents": torch.empty(B, self.I, dtype=torch.float64, device=dev),
            "his_context_mh": i64(B, H1),
            "his_intents_idx": torch.empty(B, H1, self.nz1, dtype=torch.int32, device=dev),
            "his_intents_val": torch.empty(B, H1, self.nz1, dtype=torch.float32, device=dev),
            "his_item_id": i64(B, H2),
            "his_item_int_idx": torch.empty(B, H2, 1, dtype=torch.int32, device=dev),
            "his_item_int_val": torch.empty(B, H2, 1, dtype=torch.float32, device=dev),
        }
        b = _Built()
        names = {"u_id": "u_id_c", "c_id": "c_id_c", "pay": "c_paynum_i", "fav": "c_favnum_i", "click": "c_clicknum_i", "i_id": "i_id_s",
                 "i_class": "i_class_c", "his_context": "his_context_mh"}
        for f, _ in _Built._fields_:
            setattr(b, f, _lib.ptr(out[names.get(f, f)]))
        _lib.check(_lib.load().intel_batch_build(C.byref(self._c), B, _lib.ptr(rows_dev), _lib.ptr(perm) if perm is not None else None,
                                                 L, H1, H2, C.byref(b), _lib.stream_ptr(dev)))
        out["batch_size"], out["phase"] = B, self.phase_name
        return out


def synthetic_columns(n_sessions: int, list_len: int, n_item: int, n_class: int, n_user: int, n_ctx: int, K: int, I: int, max_his: int = 20,
                      max_nnz: int = 8, seed: int = 0):
    """A synthetic Tmall-schema corpus in the columnar form, same marginals as synthetic.make_batch: every
    session has `list_len` items; a user owns `max_his` consecutive sessions and one positive item of each enters the item
    history, so both history lengths are uniform on 1..max_his-1 like the batches of the resident benchmark."""
    rng = np.random.default_rng(seed)
    N, L = n_sessions, list_len
    per_user = max(max_his, 2)
    u_id = (1 + (np.arange(N) // per_user) % n_user).astype(np.int64)
    u = rng.random((N, L))
    item_id = np.clip(1 + (u * u * n_item).astype(np.int64), 1, n_item)
    item_class = 1 + (item_id * 2654435761 % max(n_class - 1, 1))
    mu = np.array([(10.1, 1.7, 0.5, 0.0)[k % 4] for k in range(K)])
    sd = np.array([(6.1, 3.1, 3.0, 3.0)[k % 4] for k in range(K)])
    raw = rng.standard_normal((N, L, K)) * sd + mu
    lo, hi = raw.min(axis=1, keepdims=True), raw.max(axis=1, keepdims=True)
    scores = (raw - lo) / (hi - lo + 1e-6)
    cap = max(L // 2, 1)
    pay = np.minimum(rng.poisson(0.26, N), cap)
    fav = np.minimum(rng.poisson(0.94, N), cap - pay)
    clk = np.minimum(1 + rng.poisson(2.8, N), cap - pay - fav)
    pos = pay + fav + clk
    neg = L - pos - ((L - pos) * 0.1).astype(np.int64)
    slot = np.arange(L)[None, :]
    ranking = np.where(slot < pay[:, None], 3, np.where(slot < (pay + fav)[:, None], 2, np.where(slot < pos[:, None], 1,
                       np.where(slot < (pos + neg)[:, None], 0, -1)))).astype(np.int64)
    nnz = rng.integers(1, max_nnz + 1, N)
    int_off = np.concatenate([[0, 0], np.cumsum(nnz)]).astype(np.int64)               # row 0 = the zero vector
    int_idx = rng.integers(0, I, int(nnz.sum())).astype(np.int32)
    v = rng.random(int(nnz.sum())) + 0.05
    int_val = v / np.repeat(np.add.reduceat(v, int_off[1:-1]), nnz)
    ctx = rng.integers(1, n_ctx, N).astype(np.int64)
    users = n_user + 1
    uhis_off = np.zeros(users + 1, dtype=np.int64)
    uhis_off[1:] = np.cumsum(np.bincount(u_id, minlength=users))
    order = np.argsort(u_id, kind="stable")                                           # a user's sessions in row order
    position = np.empty(N, dtype=np.int64)
    position[order] = np.arange(N) - uhis_off[u_id[order]]
    first_cls = item_class[order, 0]
    beh = np.minimum(np.where(pay[order] > 0, 2, np.where(fav[order] > 0, 1, 0)), K - 1)      # behaviour block of IntEL.py:226
    columns = {
        "u_id_c": u_id, "c_id_c": np.arange(1, N + 1, dtype=np.int64), "context_mh": ctx,
        "user_mh": np.zeros(N, dtype=np.int64), "c_paynum_i": pay.astype(np.int64), "c_favnum_i": fav.astype(np.int64),
        "c_clicknum_i": clk.astype(np.int64), "session_len": np.full(N, L, dtype=np.int64), "position": position,
        "item_position": position.copy(), "intent_row": np.arange(1, N + 1, dtype=np.int64),
        "item_off": (np.arange(N + 1) * L).astype(np.int64), "item_id": item_id.reshape(-1), "item_class": item_class.reshape(-1),
        "ranking": ranking.reshape(-1), "scores": np.ascontiguousarray(scores.reshape(N * L, K)),
    }
    shared = {"uhis_off": uhis_off, "uhis_row": (order + 1).astype(np.int64), "uhis_ctx": ctx[order], "uitem_off": uhis_off.copy(),
              "uitem_id": item_id[order, 0].astype(np.int64),
              "uitem_int": (beh * (I // K) + first_cls % max(I // K, 1)).astype(np.int32),
              "int_off": int_off, "int_idx": int_idx, "int_val": int_val}
    return columns, shared, int(nnz.max())
