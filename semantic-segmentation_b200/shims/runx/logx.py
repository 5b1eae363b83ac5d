import os
import sys

import torch


class _LogX:
    def __init__(self):
        self.logdir = None
        self.rank0 = True

    def initialize(self, logdir=None, tensorboard=False, hparams=None, global_rank=0, coolname=False, eager_flush=False,
                   **kwargs):
        self.logdir = logdir
        self.rank0 = global_rank == 0
        if logdir and self.rank0:
            os.makedirs(logdir, exist_ok=True)

    def msg(self, text):
        if self.rank0:
            print(text, file=sys.stdout, flush=True)

    def metric(self, phase, metrics, epoch):
        if self.rank0:
            print("[%s %s] %s" % (phase, epoch, metrics), flush=True)

    def add_image(self, *a, **k):
        pass

    def add_scalar(self, *a, **k):
        pass

    def save_model(self, save_dict, metric, epoch, higher_better=True, delete_old=True):
        if self.rank0 and self.logdir:
            torch.save(save_dict, os.path.join(self.logdir, "last_checkpoint_ep%d.pth" % epoch))

    def get_logroot(self):
        return self.logdir or "."


logx = _LogX()
