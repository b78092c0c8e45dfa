# Stub for IntEL/src/loss/: `--loss_name IntBPRloss_b200` (criterion(out_dict, batch) -> (loss, ensemble_loss, intent_loss))
from intel_sigir2023_b200.losses import IntBPRloss as IntBPRloss_b200  # noqa: F401
