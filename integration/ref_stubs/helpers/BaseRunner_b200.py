# Stub for IntEL/src/helpers/: `--runner_name BaseRunner_b200` keeps BaseRunner.train / fit / predict (BaseRunner.py:190-355)
# and replaces the numpy evaluation (BaseRunner.py:57-150) and the optimizer (BaseRunner.py:182-188) by the B200 kernels.
import torch

from helpers.BaseRunner import BaseRunner
from intel_sigir2023_b200 import evaluate
from intel_sigir2023_b200.optim import Adam


class BaseRunner_b200(BaseRunner):
    evaluate_method = staticmethod(evaluate.evaluate_method)

    def evaluate_intents(self, true_intents, predict_intents, topk=[1, 5, 10, 30]):
        return evaluate.evaluate_intents(true_intents, predict_intents, topk=topk)

    def _build_optimizer(self, model):
        if self.optimizer_name != 'Adam':
            return super()._build_optimizer(model)
        optimizer = Adam(model.customize_parameters(), lr=self.learning_rate, weight_decay=self.l2)
        return optimizer, torch.optim.lr_scheduler.StepLR(optimizer, step_size=self.decay_step, gamma=self.decay_lr)
