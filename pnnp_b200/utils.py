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
    """utils/utils.py:88-139 without the matplotlib history plot: avg = sum / count."""

    def __init__(self, name, fmt=':f', log=True, last_epoch=0):
        self.name, self.fmt, self.log, self.history, self.last_epoch = name, fmt, log, [], last_epoch
        self.reset()

    def reset(self):
        if self.log and getattr(self, "avg", 0) > 0:
            self.history.append(self.avg)
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count

    def __str__(self):
        fmtstr = '{name}:{val' + self.fmt + '}({avg' + self.fmt + '})'
        return fmtstr.format(**self.__dict__)


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
    """base_trainer.py:141-149 — WarmUpCosine (SGDR): the rate halves every period; linear warm-up over `peak` steps from the
    second period on; cosine from lr down to ratio * lr inside a period."""
    import math
    T = step // period
    decay = 2 ** T
    step = step % period
    if step <= peak and T > 0:
        mul = step / peak
    else:
        mul = (1 - ratio) * (math.cos((step - peak) / (period - peak) * math.pi) * 0.5 + 0.5) + ratio
    return lr * mul / decay


def get_multistep_lr(step, period=1000, lr=1e-4, milestone=[500, 900], gamma=[0.5, 0.1], decay_base=1):
    """base_trainer.py:151-160."""
    decay = decay_base ** (step // period)
    step = step % period
    mul = 1
    for i in range(len(milestone), 0, -1):
        if step > milestone[i - 1]:
            mul = gamma[i - 1]
            break
    return lr * mul / decay


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
