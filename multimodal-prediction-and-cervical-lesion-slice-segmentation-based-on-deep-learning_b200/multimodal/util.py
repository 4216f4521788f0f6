"""Host helpers the reference's multimodal train scripts import from ``util.py`` (MM/*/util.py): the step learning-rate
schedule (:79-82), the tee logger (:50-67) and the clinical-graph topology (:69-77).  Pure host code."""
import os
import sys
import time

from .my_mae_model import get_edge_index_full  # noqa: F401  (util.py:69-77)


def adjust_learning_rate(optimizer, lr, epoch, lr_step=20, lr_gamma=0.5):
    """lr * gamma^(epoch // step) written into every parameter group."""
    new_lr = lr * (lr_gamma ** (epoch // lr_step))
    for group in optimizer.param_groups:
        group["lr"] = new_lr


class Logger(object):
    """Duplicates everything written to ``stream`` into ``log/<timestamp>.log``."""

    def __init__(self, stream=sys.stdout, output_dir="log"):
        os.makedirs(output_dir, exist_ok=True)
        self.terminal = stream
        self.log = open(os.path.join(output_dir, "%s.log" % time.strftime("%Y-%m-%d-%H-%M")), "a+")

    def write(self, message):
        self.terminal.write(message)
        self.log.write(message)

    def flush(self):
        pass
