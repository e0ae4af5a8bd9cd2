"""Small host-side helpers with the reference's names and behaviour (utils/utils.py:73-192,221-224)."""
import os
import time

import torch


def log(string, log=None, str=False, end='\n', notime=False):
    """utils/utils.py:73-86 — print and append to a log file; timestamp format kept so outputs diff
    against the reference's logs/*.log."""
    log_string = f'{time.strftime("%Y-%m-%d %H:%M:%S")} >>  {string}' if not notime else string
    print(log_string)
    if log is not None:
        os.makedirs(os.path.dirname(log) or ".", exist_ok=True)
        with open(log, 'a+') as f:
            f.write(log_string + '\n')
    if str:
        return string + end


class AverageMeter(object):
    """utils/utils.py:88-139 without the matplotlib history plot: running sum / count; `reset` files a positive average into
    `history` first; printing follows the reference's '{name}:{val}({avg})' format."""

    def __init__(self, name, fmt=':f', log=True, last_epoch=0):
        self.name, self.fmt, self.log, self.last_epoch = name, fmt, log, last_epoch
        self.history = []
        self.val = self.avg = self.sum = self.count = 0

    def reset(self):
        if self.log and self.avg > 0:
            self.history.append(self.avg)
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val, self.sum, self.count = val, self.sum + val * n, self.count + n
        self.avg = self.sum / self.count

    def __str__(self):
        return ('{name}:{val' + self.fmt + '}({avg' + self.fmt + '})').format(name=self.name, val=self.val, avg=self.avg)


def load_weights(model, pretrained_dict, multi_gpu=False, by_name=False):
    """utils/utils.py:148-192: by_name drops keys that are absent or shape-mismatched, then
    load_state_dict on the merged dict."""
    target = model.module if multi_gpu else model
    model_dict = target.state_dict()
    pretrained_dict = dict(pretrained_dict)
    if by_name:
        for k in list(pretrained_dict):
            if k not in model_dict:
                log(f'Warning:  "{k}" is not exist and has been deleted!!')
                del pretrained_dict[k]
            elif model_dict[k].shape != pretrained_dict[k].shape:
                log(f'Warning:  "{k}":{pretrained_dict[k].shape}->{model_dict[k].shape}')
                del pretrained_dict[k]
    model_dict.update(pretrained_dict)
    target.load_state_dict(model_dict)
    return model


def tensor_dim5to4(tensor):
    """utils/utils.py:194-197 — DataLoader adds a batch dim in front of the crop dim."""
    b, crops, c, h, w = tensor.shape
    return tensor.reshape(b * crops, c, h, w)


def get_cos_lr(step, period=1000, peak=20, lr=1e-4, ratio=0.2):
    """base_trainer.py:141-149 — WarmUpCosine (SGDR): cycle k of `period` steps runs at lr / 2^k; from the second cycle on the
    first `peak` steps ramp up linearly; otherwise a half cosine from lr down to ratio * lr (same operation order as the
    reference, so the values are equal to the last bit)."""
    import math
    cycle, pos = divmod(step, period)
    if cycle > 0 and pos <= peak:
        shape = pos / peak
    else:
        half_cos = math.cos((pos - peak) / (period - peak) * math.pi) * 0.5 + 0.5
        shape = (1 - ratio) * half_cos + ratio
    return lr * shape / 2 ** cycle


def get_multistep_lr(step, period=1000, lr=1e-4, milestone=[500, 900], gamma=[0.5, 0.1], decay_base=1):
    """base_trainer.py:151-160: inside a cycle the factor of the LAST milestone already passed (strictly), 1 before the first;
    cycle k divides by decay_base^k."""
    cycle, pos = divmod(step, period)
    passed = [g for m, g in zip(milestone, gamma) if pos > m]
    return lr * (passed[-1] if passed else 1) / decay_base ** cycle


def lr_lambda_from_hyper(hyper):
    """base_trainer.py:33-43 (get_lr_lambda_func): the schedule the YAML `hyper` block selects, as a function of the epoch."""
    num_of_epochs = hyper['stop_epoch'] - hyper['last_epoch']
    step_size = hyper['step_size']
    T = hyper['T'] if 'T' in hyper else 1
    if 'cos' in hyper['lr_scheduler'].lower():
        return lambda x: get_cos_lr(x, period=num_of_epochs // T, lr=hyper['learning_rate'], peak=step_size)
    if 'multi' in hyper['lr_scheduler'].lower():
        return lambda x: get_multistep_lr(x, period=num_of_epochs // T, decay_base=1, milestone=[step_size, step_size * 9 // 5],
                                          gamma=[0.5, 0.1], lr=hyper['learning_rate'])
    raise KeyError(f"lr_scheduler {hyper['lr_scheduler']!r}: the reference knows 'cos' and 'multi' schedules")


def lr_for_epoch(lr_lambda, epoch, last_epoch):
    """Learning rate of training epoch `epoch` (the loop runs last_epoch+1 .. stop_epoch).  The reference's LambdaScheduler
    (base_trainer.py:131-138) is built with torch's default last_epoch = -1 regardless of hyper['last_epoch'], stepped once by
    its constructor, once at the top of train() (trainer_SID.py:75) and once after every trained epoch (:127), and returns
    lmbda(last_epoch) itself: the k-th trained epoch runs at lr_lambda(k), k = epoch - hyper['last_epoch']."""
    return lr_lambda(epoch - last_epoch)


class PhaseTimer:
    """The reference's per-step split `dataloader / preprocess / net / bp` (trainer_SID.py:81-123: wall-clock stamps without a
    synchronize, so its GPU phases are launch-side only) with device time: every phase is an NVTX range (visible in Nsight Systems /
    `ncu --nvtx`) bracketed by CUDA events on the current stream; `summary()` resolves the events once per epoch, so the loop itself
    never synchronises.  Host-only phases (no kernels in between) read as ~0 device time and are reported by wall clock too."""

    def __init__(self, device=None, enabled=True):
        import torch
        self.torch, self.device, self.enabled = torch, device, enabled and torch.cuda.is_available()
        self.events, self.wall, self._open = {}, {}, None

    def start(self, name):
        import time as _t
        self.stop()
        if not self.enabled:
            return
        t = self.torch
        t.cuda.nvtx.range_push(name)
        e0 = t.cuda.Event(enable_timing=True)
        e0.record()
        self._open = (name, e0, _t.perf_counter())

    def stop(self):
        import time as _t
        if self._open is None:
            return
        name, e0, w0 = self._open
        e1 = self.torch.cuda.Event(enable_timing=True)
        e1.record()
        self.torch.cuda.nvtx.range_pop()
        self.events.setdefault(name, []).append((e0, e1))
        self.wall[name] = self.wall.get(name, 0.0) + (_t.perf_counter() - w0)
        self._open = None

    def summary(self):
        """{phase: (device ms, host wall ms)} accumulated since the last call; clears the accumulators."""
        self.stop()
        if not self.enabled:
            return {}
        self.torch.cuda.synchronize(self.device)
        out = {k: (sum(a.elapsed_time(b) for a, b in v), self.wall.get(k, 0.0) * 1e3) for k, v in self.events.items()}
        self.events, self.wall = {}, {}
        return out
