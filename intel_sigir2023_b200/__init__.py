"""B200-native IntEL hot path (drop-in for the reference's model / loss / evaluation interface)."""
from .config import IntelConfig  # noqa: F401

__all__ = ["IntelConfig", "IntEL", "IntListloss", "IntBPRloss", "IntMSEloss", "Listloss", "BPRloss", "MSEloss",
           "evaluate_method", "evaluate_intents", "SingleSort", "Borda", "RandomFusion", "aWELv", "aWELv_Int", "aWELv_IntEL"]


def __getattr__(name):
    # lazy: importing the package must not require torch/CUDA (config and synthetic data are pure python)
    if name == "IntEL":
        from .IntEL import IntEL
        return IntEL
    if name in ("IntListloss", "IntBPRloss", "IntMSEloss", "Listloss", "BPRloss", "MSEloss"):
        from . import losses
        return getattr(losses, name)
    if name in ("evaluate_method", "evaluate_intents"):
        from . import evaluate
        return getattr(evaluate, name)
    if name in ("SingleSort", "Borda", "RandomFusion", "aWELv", "aWELv_Int", "aWELv_IntEL"):
        from . import baselines
        return getattr(baselines, name)
    raise AttributeError(name)
