"""Drive the UNMODIFIED reference the way its main.py does (IntEL/src/main.py:44-139), class names resolved by name,
from the copy under oracle/_ref/ (oracle/make_ref.py) or from /root/reference.  Test infrastructure only (tests/,
bench.py's reference arm).

main.py itself cannot be imported here: its `from models.supervise import *` pulls in ERA.py, which needs `pygad`
(not in this image).  `wire()` repeats main.py's wiring with the same parsers, the same construction order and the same
seeding, importing only the modules the selected names need.
"""
from __future__ import annotations

import argparse
import importlib
import logging
import os
from typing import Dict, List, Optional

import numpy as np
import torch

from . import ref_shims

# script/IntEL.sh:21 (IntEL-PL), :15 (IntEL-BPR), :9 (IntEL-MSE) without the seed / gpu / save_anno flags
COMMON = ["--batch_size", "512", "--topk", "3,1,5,10", "--max_session_len", "100", "--intent_note", "_multi", "--model_num", "3",
          "--kl_weight", "0.5", "--main_metric", "NDCG@3", "--cal_diversity", "1"]
SCRIPT_FLAGS = {
    "pl": COMMON + ["--test_epoch", "5", "--intent_weight", "0.1", "--lr", "2e-3", "--l2", "1e-4", "--dropout", "0", "--decay_lr", "0",
                    "--context_emb_size", "32", "--intent_emb_size", "32", "--encoder", "GRU4Rec", "--i_emb_size", "16",
                    "--im_emb_size", "16", "--u_emb_size", "32", "--s_emb_size", "32", "--cross_attn_qsize", "64", "--num_heads", "2",
                    "--num_layers", "2", "--diversity_alpha", "1e-4"],
    "bpr": COMMON + ["--test_epoch", "3", "--intent_weight", "0.01", "--lr", "1e-4", "--l2", "1e-4", "--dropout", "0",
                     "--context_emb_size", "64", "--intent_emb_size", "32", "--encoder", "GRU4Rec", "--i_emb_size", "16",
                     "--im_emb_size", "16", "--u_emb_size", "32", "--s_emb_size", "32", "--diversity_alpha", "1e-5",
                     "--cross_attn_qsize", "32", "--num_heads", "2", "--num_layers", "2"],
    "mse": COMMON + ["--test_epoch", "3", "--intent_weight", "0.003", "--encoder", "BERT4Rec", "--lr", "1e-3", "--l2", "1e-6",
                     "--dropout", "0.5", "--diversity_alpha", "1e-5"],
}
LOSS_NAME = {"pl": "IntListloss", "bpr": "IntBPRloss", "mse": "IntMSEloss"}


def available() -> bool:
    return ref_shims.ref_root() is not None


def _cls(package: str, name: str):
    return getattr(importlib.import_module(f"{package}.{name}"), name)


def _global_args(parser):            # main.py:23-41
    parser.add_argument('--gpu', type=str, default='')
    parser.add_argument('--verbose', type=int, default=logging.INFO)
    parser.add_argument('--log_file', type=str, default='')
    parser.add_argument('--random_seed', type=int, default=0)
    parser.add_argument('--load', type=int, default=0)
    parser.add_argument('--train', type=int, default=1)
    parser.add_argument('--regenerate', type=int, default=0)
    parser.add_argument('--save_anno', type=str, default='test')
    parser.add_argument('--test_train', type=int, default=0)
    return parser


class Wired:
    """args, corpus, model, criterion, runner, data_dict of one main.py run"""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def wire(model_name: str, loss_name: str, runner_name: str, flags: List[str], device: torch.device, work_dir: str,
         share_from: Optional["Wired"] = None, seed: int = 0, phases=("train", "dev", "test")) -> Wired:
    """share_from: a previous wiring whose corpus (and buffered dev / test feed dicts) this one re-uses: Dataset.prepare()
    pops the list columns out of the corpus once the feed dicts are buffered (BaseModel.py:112-119), so a corpus can be
    prepared only once."""
    root = ref_shims.ref_root()
    assert root is not None, "no reference tree (run oracle/make_ref.py in the build container)"
    ref_shims.install(os.path.join(root, "src"))
    model_cls = _cls("models.IntEL", model_name)
    reader_cls = _cls("helpers", model_cls.reader)
    runner_cls = _cls("helpers", runner_name)
    loss_cls = _cls("loss", loss_name)
    ref_shims.patch_predict_ragged(_cls("helpers", "BaseRunner"))
    parser = argparse.ArgumentParser(description='')
    parser = _global_args(parser)
    parser = reader_cls.parse_data_args(parser)
    parser = model_cls.parse_model_args(parser)
    parser = runner_cls.parse_runner_args(parser)
    parser = loss_cls.parse_loss_args(parser)
    args, _ = parser.parse_known_args(["--datapath", os.path.join(root, "data"), "--dataset", "Tmall_toy", "--random_seed", str(seed),
                                       "--num_workers", "0"] + list(flags))
    os.makedirs(work_dir, exist_ok=True)
    args.model_path = os.path.join(work_dir, f"{model_name}.pt")
    args.device = device
    np.random.seed(args.random_seed)                   # main.py:51-54
    torch.manual_seed(args.random_seed)
    corpus = share_from.corpus if share_from is not None else reader_cls(args)
    model = model_cls(args, corpus).to(device)
    criterion = loss_cls(args)
    runner = runner_cls(args)
    data_dict: Dict[str, object] = {}
    for phase in phases:
        data_dict[phase] = model_cls.Dataset(model, corpus, phase)
        if share_from is not None and phase != "train" and model.buffer:
            data_dict[phase].buffer_dict = share_from.data[phase].buffer_dict
        else:
            data_dict[phase].prepare()
    return Wired(args=args, corpus=corpus, model=model, criterion=criterion, runner=runner, data=data_dict)


def first_batch(w: Wired, phase: str = "train", n: int = 512, seed: int = 0) -> Dict[str, object]:
    """the first `n` sessions of a phase through the reference's own Dataset.collate_batch (BaseModel.py:121-142); the
    per-session list shuffle (BaseModel.py:194-196) draws from numpy's global RNG, seeded here"""
    ds = w.data[phase]
    np.random.seed(seed)
    return ds.collate_batch([ds[i] for i in range(min(n, len(ds)))])


def to_device(batch: Dict[str, object], device) -> Dict[str, object]:
    return {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in batch.items()}


# ---- the reference classes on a synthetic corpus (bench.py's reference arm, BASELINE.json configs[1..]) ----
class _SyntheticCorpus:
    """the attributes IntEL.__init__ reads from a reader (IntEL.py:36-49, BaseModel.py:152-156)"""

    def __init__(self, cfg):
        self.itemfnum = [cfg.class_rows]
        self.contextfnum = [cfg.ctx_rows]
        self.zero_int = np.zeros(cfg.intent_num)
        self.max_uid = cfg.user_rows - 1
        self.max_iid = cfg.item_rows - 1


def reference_on_synthetic(cfg, loss_kind: str, loss_kw: Dict[str, float], device=torch.device("cpu")):
    """(model, criterion) = the UNMODIFIED reference IntEL + Int{List,BPR,MSE}loss for an IntelConfig"""
    root = ref_shims.ref_root()
    assert root is not None
    ref_shims.install(os.path.join(root, "src"))
    model_cls = _cls("models.IntEL", "IntEL")
    loss_cls = _cls("loss", {"list": "IntListloss", "bpr": "IntBPRloss", "mse": "IntMSEloss"}[loss_kind])
    a = argparse.Namespace(**cfg.to_dict())
    a.device, a.model_path, a.buffer = device, "/tmp/intel_ref/model.pt", 1
    a.cal_diversity, a.ensemble_weight, a.kl_temp, a.kl_weight = 1, 1.0, 2.0, 0.5
    for k, v in loss_kw.items():
        setattr(a, k, v)
    model = model_cls(a, _SyntheticCorpus(cfg)).to(device)
    return model, loss_cls(a)
