"""Host-side learning-rate schedule of the reference's ``sgd`` optimizer (training/sgd.py:79-185): a
fixed-rate step decay or a cosine decay with warm restarts, evaluated once per batch; the value is the
``lr`` of that batch's ``kge_train_step``.  adam / adagrad / momentum use a constant rate
(training/adam.py:31-48, adagrad.py:30-46, momentum.py:51-69)."""
from __future__ import annotations

import math

DEFAULT_LR = 0.0005  # training/_optimizer_constants.py:9
DEFAULT_DECAY_CYCLE = 0  # :12
DEFAULT_DECAY_CYCLE_MULTIPLE = 1  # :13
DEFAULT_LR_DECAY_FACTOR = 2  # :14
DEFAULT_END_LR = 1e-8  # :15
DEFAULT_SINE = False  # :16


class SGDSchedule:
    """lr(batch_num, epoch_num), both 1-based, called in training order (the schedule is stateful)."""

    def __init__(self, optimizer_params, batches_count):
        p = optimizer_params or {}
        self.batches_count = int(batches_count)
        self.peak = p.get("lr", DEFAULT_LR)          # start rate of the running cycle
        self.lr = self.peak                          # rate handed to the current batch
        self.floor = p.get("end_lr", DEFAULT_END_LR)
        self.cycle0 = p.get("decay_cycle", DEFAULT_DECAY_CYCLE)
        self.cosine = bool(p.get("cosine_decay", DEFAULT_SINE))
        self.expand = p.get("expand_factor", DEFAULT_DECAY_CYCLE_MULTIPLE)
        self.shrink = p.get("decay_lr_rate", DEFAULT_LR_DECAY_FACTOR)
        self.cycle_len = self.cycle0                 # epochs in the running cosine cycle
        self.cycle_start = 0                         # epoch at which the running cosine cycle began
        self.next_epoch = self.cycle0 + 1            # first epoch of the next cycle

    def __call__(self, batch_num, epoch_num):
        if self.cosine:
            # position inside the running cycle in [0,1): half a cosine from peak down to floor
            done = (epoch_num - 1 - self.cycle_start) * self.batches_count + (batch_num - 1)
            frac = done / (self.cycle_len * self.batches_count)
            self.lr = self.floor + (self.peak - self.floor) * 0.5 * (1 + math.cos(math.pi * frac))
            if epoch_num % (self.next_epoch - 1) == 0 and batch_num == self.batches_count:
                # last batch of the cycle: restart with a longer cycle and a lower peak
                self.cycle_len = self.cycle_len * self.expand
                self.next_epoch = self.next_epoch + self.cycle_len
                self.cycle_start = epoch_num
                self.peak = self.peak / self.shrink
            self.lr = max(self.lr, self.floor)
        elif self.cycle0 > 0:
            if epoch_num % self.next_epoch == 0 and batch_num == 1 and self.lr > self.floor:
                self.next_epoch = self.cycle0 + (self.next_epoch - 1) * self.expand + 1
                self.lr = max(self.lr / self.shrink, self.floor)
        return self.lr
