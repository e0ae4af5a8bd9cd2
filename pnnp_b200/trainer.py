"""Eval entry points with the reference's CLI, YAML schema, loop structure and log format
(base_trainer.py:6-129, trainer_SID.py:11-72,181-315,421-486,519-563, trainer_LRID.py:195-319,474-505),
running the hot path on the B200 kernels:

    python trainer_SID.py  -f runfiles/SonyA7S2/PNNP.yml --mode evaltest
    python trainer_LRID.py -f runfiles/IMX686/PNNP.yml   --mode evaltest
    torchrun --nproc-per-node 8 trainer_SID.py -f ... --mode evaltest     (frames sharded by rank)

Per frame: synthetic clean frame -> fused noise synthesis at the dataset's ratio/ISO (device) ->
[reflect-pad 4 if W % 16] -> UNet / ResUnet forward (tcgen05) -> crop -> x ratio if `ori` -> clamp ->
IlluminanceCorrect (Sony, final eval) -> PSNR / SSIM partial sums (device) -> AverageMeter.
Across ranks: one all-reduce of [sum PSNR, sum SSIM, count] per sweep.
`--mode train` runs the synthetic-pair training loop (trainer_SID.py:74-180) on the explicit training step of train.py.
Out of scope: plotting, ProcessPool rendering, real-dataset loaders.
"""
import argparse
import os
import pickle as pkl
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F
import yaml

from . import _lib, distributed as D
from .archs import ResUnet, UNetSeeInDark, initialize_weights  # noqa: F401  (resolved by name from YAML)
from .datasets import Raw_Dataset, Synthetic_ELD_Dataset, Synthetic_IMX686_Dataset, Synthetic_SID_Dataset  # noqa: F401
from .metrics import eval_partial_sums, finish_metrics
from .noise import synthesize_batch
from .noise_params import HALF_CLIP, sample_params_max
from .utils import AverageMeter, PhaseTimer, load_weights, log, lr_for_epoch, lr_lambda_from_hyper, tensor_dim5to4


class BaseParser():
    """base_trainer.py:6-17 / trainer_SID.py:504-517: same six flags."""

    def __init__(self, default_runfile="runfiles/SonyA7S2/PNNP.yml"):
        self.parser = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
        self.default_runfile = default_runfile

    def parse(self, argv=None):
        a = self.parser.add_argument
        a('--runfile', '-f', default=self.default_runfile, type=Path, help="path to config")
        a('--mode', '-m', default='evaltest', type=str, help="train or test")
        a('--debug', action='store_true', default=False, help="debug or not")
        a('--nofig', action='store_true', default=True, help="don't save_plot")
        a('--nohost', action='store_true', default=True, help="no host-specific data roots")
        a('--gpu', default="0", help="device index when not launched by torchrun")
        return self.parser.parse_args(argv)


class Base_Trainer():
    def initialization(self):
        """base_trainer.py:45-81 (host-specific data roots are not applied: --nohost semantics)."""
        with open(self.parser.runfile, 'r', encoding="utf-8") as f:
            self.args = yaml.load(f.read(), Loader=yaml.FullLoader)
        self.mode = self.args['mode'] if self.parser.mode is None else self.parser.mode
        if self.parser.debug:
            self.args['num_workers'] = 0
        self.args['dst'].setdefault('clip', False)
        self.save_plot = False
        self.dst, self.hyper, self.arch = self.args['dst'], self.args['hyper'], self.args['arch']
        self.rank, self.world_size = D.init_from_env()
        index = int(os.environ.get("LOCAL_RANK", self.parser.gpu.split(",")[0] if self.world_size == 1 else 0))
        self.device = _lib.cuda_device(index, set_current=True)         # raises without a CUDA device: there is no CPU fallback
        self.model_name, self.fast_ckpt = self.args['model_name'], self.args['fast_ckpt']
        self.model_dir = self.args['checkpoint']
        self.sample_dir = os.path.join(self.args['result_dir'], f"samples-{self.model_name}")
        for d in (self.model_dir, './logs', self.fast_ckpt, './metrics'):
            os.makedirs(d, exist_ok=True)
        self.logfile = f'./logs/log_{self.model_name}.log' if self.rank == 0 else None

    def print_model_log(self):
        """base_trainer.py:83-120: header lines in the reference's format."""
        self.best_psnr = self.hyper.get('best_psnr', 0)
        self.eval_psnr, self.eval_ssim = AverageMeter('PSNR', ':2f'), AverageMeter('SSIM', ':4f')
        self.eval_psnr_lr, self.eval_ssim_lr = AverageMeter('PSNR', ':2f'), AverageMeter('SSIM', ':4f')
        self.eval_psnr_dn, self.eval_ssim_dn = AverageMeter('PSNR', ':2f'), AverageMeter('SSIM', ':4f')
        if self.rank != 0:
            return
        rows = (("Model Name:\t", self.model_name), ("Architecture:\t", self.arch["name"]),
                ("TrainDataset:\t", self.args["dst_train"]["dataset"]), ("EvalDataset:\t", self.args["dst_eval"]["dataset"]),
                ("CameraType:\t", self.dst["camera_type"]), ("num_channels:\t", self.arch["nf"]),
                ("BatchSize:\t", self.hyper["batch_size"]), ("PatchSize:\t", self.dst["patch_size"]),
                ("LearningRate:\t", self.hyper["learning_rate"]), ("Epoch:\t\t", self.hyper["stop_epoch"]),
                ("num_workers:\t", self.args["num_workers"]), ("Command:\t", self.dst["command"]))
        for k, v in rows:
            log(f'{k}{v}', log=self.logfile, notime=True)
        log(f"Let's use {self.world_size} GPUs!", log=self.logfile, notime=True)

    def metrics_reset(self):
        for m in (self.eval_psnr, self.eval_ssim, self.eval_psnr_lr, self.eval_ssim_lr, self.eval_psnr_dn, self.eval_ssim_dn):
            m.reset()


class SID_Trainer(Base_Trainer):
    default_runfile = "runfiles/SonyA7S2/PNNP.yml"

    def __init__(self, argv=None):
        self.parser = BaseParser(self.default_runfile).parse(argv)
        self.initialization()
        self.net = globals()[self.arch['name']](self.arch)            # trainer_SID.py:17
        if self.hyper['last_epoch']:
            try:                                                      # trainer_SID.py:19-28
                path = f'{self.fast_ckpt}/{self.model_name}_best_model.pth'
                if not os.path.exists(path):
                    path = f'{self.fast_ckpt}/{self.model_name}_last_model.pth'
                self.net = load_weights(self.net, torch.load(path, map_location="cpu"), by_name=True)
            except Exception:
                log('No checkpoint file!!!')
        else:
            log(f'Initializing {self.arch["name"]}...')
            initialize_weights(self.net)
        self.infos = None
        self.net = self.net.to(self.device)
        self.multi_gpu = False
        self.print_model_log()

    def change_eval_dst(self, mode='eval'):
        self.dst = self.args[f'dst_{mode}']
        self.dstname = self.dst['dstname']
        self.dst_eval = globals()[self.dst['dataset']](self.dst)      # trainer_SID.py:69
        self.cache_dir = f'/data/cache/{self.dstname}'

    # -- preprocess: the GPU route of trainer_SID.py:421-486 (noise synthesis + clamps) in one launch
    def preprocess(self, data, mode='eval', preprocess=True):
        hr = tensor_dim5to4(data['hr']).float().to(self.device, non_blocking=True)
        ratios = data['ratio'].float().view(-1).tolist()
        if preprocess:
            params = data['param_list']
            # Philox (seed, offset) = (fixed seed, sweep number) and crop id = dataset index: every rank count gives
            # bit-identical noisy frames, so sharded and single-GPU sweeps report the same metrics
            lr = synthesize_batch(hr.contiguous(), params, self.dst['noise_code'], _lib.CHAIN_NUMPY, ori=self.dst['ori'],
                                  clip=False, crop_id0=int(data['index'][0]),
                                  seed_offset=(getattr(self, 'noise_seed', 1997), getattr(self, 'sweep_id', 0)))
        else:
            lr = tensor_dim5to4(data['lr']).float().to(self.device, non_blocking=True)
        ratio = torch.tensor(ratios, device=self.device).view(-1, 1, 1, 1)
        if self.dst['clip']:
            lb = -np.inf if self.dst['clip'] == HALF_CLIP else 0     # trainer_SID.py:481-485
            lr, hr = lr.clamp(lb, 1), hr.clamp(0, 1)
        return lr, hr, ratio

    def forward_frame(self, imgs_lr):
        """trainer_SID.py:221-228 / trainer_LRID.py:224-231."""
        if imgs_lr.shape[-1] % 16 != 0 or imgs_lr.shape[-2] % 16 != 0:
            padded = F.pad(imgs_lr, (4, 4, 4, 4), mode='reflect')
            return self.net(padded)[..., 4:-4, 4:-4].contiguous()
        return self.net(imgs_lr)

    use_corrector = True

    def eval(self, epoch=-1):
        self.net.eval()
        self.metrics_reset()
        self.sweep_id = getattr(self, 'sweep_id', 0) + 1
        metrics, metrics_path = {}, f'./metrics/{self.model_name}_metrics.pkl'
        if self.rank == 0 and os.path.exists(metrics_path):
            with open(metrics_path, 'rb') as f:
                metrics = pkl.load(f)
        correct = bool(self.args.get('brightness_correct')) and epoch < 0 and self.use_corrector
        psnr_sum = ssim_sum = 0.0
        psnr_lr_sum = ssim_lr_sum = 0.0
        mine = D.shard_range(len(self.dst_eval), self.rank, self.world_size)
        pending = []
        with torch.no_grad():
            for k in mine:
                item = self.dst_eval[k]
                data = {"hr": item["hr"][None], "lr": item["lr"][None], "ratio": torch.tensor([float(item["ratio"])]),
                        "param_list": [item["param"]], "index": [item["index"]]}
                imgs_lr, imgs_hr, ratio = self.preprocess(data, mode='eval', preprocess=True)
                imgs_dn = self.forward_frame(imgs_lr)
                scale = float(item["ratio"]) if self.dst['ori'] else 1.0
                n, c, h, w = imgs_dn.shape
                pending.append((item["name"], eval_partial_sums(imgs_dn, imgs_hr.contiguous(), scale, correct),
                                eval_partial_sums(imgs_lr.contiguous(), imgs_hr.contiguous(), scale, False), (c, h, w)))
        for name, s_dn, s_lr, (c, h, w) in pending:                  # one small D2H per frame, after the loop
            r_dn, r_lr = finish_metrics(s_dn, c, h, w)[0], finish_metrics(s_lr, c, h, w)[0]
            metrics[name] = [r_dn['PSNR'], r_dn['SSIM']]
            psnr_sum += r_dn['PSNR']; ssim_sum += r_dn['SSIM']
            psnr_lr_sum += r_lr['PSNR']; ssim_lr_sum += r_lr['SSIM']
        cnt = len(pending)
        if self.world_size > 1 and epoch < 0:                            # per-frame entries of every rank's shard -> rank 0's pickle
            mine_named = {name: metrics[name] for name, *_ in pending}
            gathered = [None] * self.world_size if self.rank == 0 else None
            torch.distributed.gather_object(mine_named, gathered, dst=0)
            if self.rank == 0:
                for part in gathered:
                    metrics.update(part)
        p_dn, s_dn, total = D.reduce_metric_sums(psnr_sum, ssim_sum, cnt, self.device)
        p_lr, s_lr, _ = D.reduce_metric_sums(psnr_lr_sum, ssim_lr_sum, cnt, self.device)
        self.eval_psnr.update(p_dn, total); self.eval_ssim.update(s_dn, total)
        self.eval_psnr_lr.update(p_lr, total); self.eval_ssim_lr.update(s_lr, total)
        self.eval_psnr_dn, self.eval_ssim_dn = self.eval_psnr, self.eval_ssim
        if self.eval_psnr_dn.avg >= self.best_psnr and epoch > 0:        # trainer_SID.py:302-307 (the all-reduced average: every
            self.best_psnr = self.eval_psnr_dn.avg                       # rank tracks the same record, rank 0 writes the file)
            if self.rank == 0:
                log(f"Best PSNR is {self.best_psnr} now!!")
                torch.save(_detached_state(self.net), f'{self.fast_ckpt}/{self.model_name}_best_model.pth')
        if self.rank == 0:
            log(f"Epoch {epoch}: PSNR={self.eval_psnr.avg:.2f}\n"
                + f"psnrs_lr={self.eval_psnr_lr.avg:.2f}, psnrs_dn={self.eval_psnr_dn.avg:.2f}"
                + f"\nssims_lr={self.eval_ssim_lr.avg:.4f}, ssims_dn={self.eval_ssim_dn.avg:.4f}", log=self.logfile)
            if epoch < 0:
                with open(metrics_path, 'wb') as f:
                    pkl.dump(metrics, f)
        return {"PSNR": self.eval_psnr.avg, "SSIM": self.eval_ssim.avg, "frames": total}

    def predict(self, raw, name='ds', tile_batch=8):
        """trainer_SID.py:345-360: tile-wise inference of a full RAW frame — raw2bayer(raw + bl) (the reference passes no
        wp / bl here, i.e. raw2bayer's defaults) -> eval_crop (overlapped tiles) -> network per tile -> eval_merge ->
        `<name>.npy`.  Returns the merged (c, h, w) array."""
        from .isp_ops import raw2bayer
        self.net.eval()
        if not hasattr(self, 'dst_eval'):
            self.change_eval_dst('eval')
        raw = torch.as_tensor(np.asarray(raw)).to(self.device)
        img_lr = raw2bayer((raw.float() + self.dst["bl"]).contiguous())[None]
        with torch.no_grad():
            tiles = self.dst_eval.eval_crop(img_lr)
            outs = [self.net(tiles[s:s + tile_batch].contiguous()) for s in range(0, tiles.shape[0], tile_batch)]
            img_dn = self.dst_eval.eval_merge(torch.cat(outs))[0].cpu().numpy()
        np.save(f'{name}.npy', img_dn)
        return img_dn

    # -- T1: the synthetic-pair training loop of trainer_SID.py:74-180 on the explicit B200 training step
    def get_lr_lambda_func(self):
        """base_trainer.py:33-43."""
        self.lr_lambda = lr_lambda_from_hyper(self.hyper)
        return self.lr_lambda

    def preprocess_train(self, imgs_lr, imgs_hr, dst_args):
        """The `gpu_preprocess: True` route of trainer_SID.py:449-462 + :481-485 for Raw_Dataset batches: the dataset hands over
        clean crops, every crop gets `sample_params_max(camera_type, ratio=None)` and the float32 (torch) noise chain —
        generate_noisy_torch's arithmetic, one fused launch for the batch — then lr.clamp(lb, 1), hr.clamp(0, 1).  As in the
        reference only the codes 'p', 'pr', 'prq' exist on this route ('g' / 'd' raise)."""
        n = imgs_lr.shape[0]
        params = [sample_params_max(camera_type=dst_args['camera_type'], ratio=None) if dst_args.get('params') is None
                  else dst_args['params'] for _ in range(n)]
        imgs_lr = synthesize_batch(imgs_lr, params, dst_args['noise_code'], _lib.CHAIN_TORCH, ori=dst_args['ori'],
                                   clip=bool(dst_args['clip']))
        if dst_args['clip']:
            lb = -np.inf if dst_args['clip'] == HALF_CLIP else 0
            imgs_lr, imgs_hr = imgs_lr.clamp(lb, 1), imgs_hr.clamp(0, 1)
        return imgs_lr, imgs_hr

    def train(self):
        """trainer_SID.py:74-180.  Per step: `batch_size` dataset items (each `crop_per_image` noisy/clean crop pairs built on
        the device by Raw_Dataset: P1 -> D2 -> S2 -> fused N1-N3) -> UNetTrainStep (forward, L1 on pred.clamp(0,1), explicit
        backward, DDP gradient all-reduce, Adam).  Ranks take disjoint items of every epoch (same permutation on every rank)."""
        from .train import UNetTrainStep
        from .train_resunet import ResUnetTrainStep
        if not isinstance(self.net, (UNetSeeInDark, ResUnet)):
            raise RuntimeError("pnnp_b200: the explicit training step is built for UNetSeeInDark and ResUnet")
        TrainStep = UNetTrainStep if isinstance(self.net, UNetSeeInDark) else ResUnetTrainStep
        dst_args = dict(self.args['dst_train'])
        if dst_args.get('ori'):
            raise RuntimeError("pnnp_b200: training with ori=True (pred * ratio inside the loss) is not built")
        self.dst_train = globals()[dst_args['dataset']](dst_args)       # trainer_SID.py:48
        self.change_eval_dst('eval')
        lr_lambda = self.get_lr_lambda_func()
        step = TrainStep(self.net, lr=lr_lambda(1))
        self.train_psnr = AverageMeter('PSNR', ':2f')
        bs = int(self.hyper['batch_size'])
        phases = PhaseTimer(self.device)          # dataloader / preprocess / net+bp, as trainer_SID.py:81-123 splits a step
        for epoch in range(self.hyper['last_epoch'] + 1, self.hyper['stop_epoch'] + 1):
            self.net.train()
            self.train_psnr.reset()
            # LambdaScheduler starts at last_epoch = -1 whatever hyper['last_epoch'] is, is stepped once by its constructor, once at
            # the top of train() and once per trained epoch (trainer_SID.py:57,75,127): the k-th trained epoch runs at lr_lambda(k)
            step.lr = lr_for_epoch(lr_lambda, epoch, self.hyper['last_epoch'])
            order = np.random.RandomState(1997 + epoch).permutation(len(self.dst_train))   # DataLoader(shuffle=True)
            batches = [order[i:i + bs] for i in range(0, len(order), bs)]
            losses = []
            per_rank = -(-len(batches) // self.world_size)             # every rank takes the same number of steps (the
            for j in range(per_rank):                                   # gradient all-reduce is collective); wrap around
                k = (self.rank * per_rank + j) % len(batches)
                phases.start('dataloader')                              # device-built items: pack + crop/aug (+ CPU-route synthesis)
                items = [self.dst_train[int(i)] for i in batches[k]]
                imgs_lr = torch.cat([it['lr'] for it in items]).contiguous()
                imgs_hr = torch.cat([it['hr'] for it in items]).contiguous()
                phases.start('preprocess')                              # GPU-route synthesis + clamps (trainer_SID.py:449-485)
                if dst_args['gpu_preprocess'] is not False:
                    imgs_lr, imgs_hr = self.preprocess_train(imgs_lr, imgs_hr, dst_args)
                phases.start('net+bp')                                  # forward, L1, backward, all-reduce, Adam: one graph replay
                losses.append(step.step(imgs_lr, imgs_hr))
                phases.stop()
            if losses:
                with torch.no_grad():                                   # PSNR of the last batch, as the progress bar shows
                    mse = (step.scr.bufs['pred'].clamp(0, 1) - imgs_hr.clamp(0, 1)).pow(2).mean()   # both steps keep `pred` there
                    self.train_psnr.update(float(-10.0 * torch.log10(mse)))
            if self.rank == 0:
                mean_loss = float(torch.stack(losses).mean()) if losses else float('nan')
                log(f"Epoch {epoch}: lr={step.lr:.2e}, L1={mean_loss:.5f}, PSNR={self.train_psnr.avg:.2f}", log=self.logfile)
            rt = phases.summary()                                       # every rank: summary() synchronises the device
            if self.rank == 0 and rt:
                log("runtime (device ms | host ms): " + ", ".join(f"{k}={d:.1f}|{h:.1f}" for k, (d, h) in rt.items()),
                    log=self.logfile)
            if epoch % self.hyper['save_freq'] == 0 and self.rank == 0:
                epoch_id = epoch // self.hyper['plot_freq'] * self.hyper['plot_freq']
                torch.save(_detached_state(self.net), os.path.join(self.model_dir, '%s_e%04d.pth' % (self.model_name, epoch_id)))
            if epoch % self.hyper['plot_freq'] == 0:                    # fast eval + last-model checkpoint
                if self.rank == 0:
                    log(f"learning_rate: {step.lr:.3e}", log=self.logfile)
                self.eval(epoch=epoch)
                if self.rank == 0:
                    torch.save(_detached_state(self.net), f'{self.fast_ckpt}/{self.model_name}_last_model.pth')
            # reload best model each period (trainer_SID.py:170-179); optimiser moments are kept, as there
            period = (self.hyper['stop_epoch'] - self.hyper['last_epoch']) // (self.hyper['T'] if 'T' in self.hyper else 1)
            if period > 0 and (self.hyper['last_epoch'] + epoch) % period == 0:
                if self.world_size > 1:
                    torch.distributed.barrier()                         # rank 0 may just have written the file
                model_path = f'{self.fast_ckpt}/{self.model_name}_best_model.pth'
                if os.path.exists(model_path):
                    self.net = load_weights(self.net, torch.load(model_path, map_location=self.device), by_name=True)
                    step.sync_parameters()                              # the parameters are views of the step's flat buffer
                    step.refresh_packed()
                    if self.rank == 0:
                        log(f'Successfully reload best model (Eval PSNR:{self.best_psnr})', log=self.logfile)
        if self.rank == 0:
            torch.save(_detached_state(self.net), f'{self.fast_ckpt}/{self.model_name}_last_model.pth')
        step.close()                                                    # captured all-reduces must not outlive the process group
        if self.world_size > 1:
            torch.distributed.barrier()
        return step


def _detached_state(net):
    """state_dict with private storage per tensor (the training step keeps all parameters as views of one flat buffer)."""
    return {k: v.detach().clone() for k, v in net.state_dict().items()}


class IMX686_Trainer(SID_Trainer):
    """trainer_LRID.py: same loop; always takes the reflect-pad branch (2312 % 16 = 8) and has no
    IlluminanceCorrect call in eval (trainer_LRID.py:195-319)."""
    default_runfile = "runfiles/IMX686/PNNP.yml"
    use_corrector = False


def _load_best_or_make_checkpoint(trainer):
    """trainer_SID.py:527-533 loads `<fast_ckpt>/<model>_best_model.pth` (else `_last_model.pth`) unguarded.
    No released weights exist in this environment, so a missing file is created once from the reference
    initialiser (seed 1997) with the reference's state_dict keys and then loaded through load_weights."""
    best = os.path.join(trainer.fast_ckpt, f'{trainer.model_name}_best_model.pth')
    if not os.path.exists(best):
        best = os.path.join(trainer.fast_ckpt, f'{trainer.model_name}_last_model.pth')
    if not os.path.exists(best):
        if trainer.rank == 0:
            log(f'No checkpoint at {best}: writing a random-init one (initialize_weights, seed 1997)')
            torch.manual_seed(1997)
            fresh = globals()[trainer.arch['name']](trainer.arch)
            initialize_weights(fresh)
            torch.save(fresh.state_dict(), best)
        if trainer.world_size > 1:
            torch.distributed.barrier()
    state = torch.load(best, map_location=trainer.device)
    trainer.net = load_weights(trainer.net, state, multi_gpu=trainer.multi_gpu)


def main_sid(argv=None):
    """trainer_SID.py:519-563."""
    trainer = SID_Trainer(argv)
    if trainer.mode == 'train':
        trainer.train()
        trainer.mode = 'evaltest'                                       # trainer_SID.py:527: training ends with the full sweeps
    _load_best_or_make_checkpoint(trainer)
    results = {}
    if 'eval' in trainer.mode:
        trainer.change_eval_dst('eval')
        for dgain in trainer.args['dst_eval']['ratio_list']:
            if trainer.rank == 0:
                log(f'ELD Datasets: Dgain={dgain}', log=trainer.logfile)
            trainer.dst_eval.ratio_list = [dgain]
            trainer.dst_eval.recheck_length()
            results[f'eval_x{dgain}'] = trainer.eval(-1)
    if 'test' in trainer.mode:
        trainer.change_eval_dst('test')
        for dgain in [100, 250, 300]:
            if trainer.rank == 0:
                log(f'SID Datasets: Dgain={dgain}', log=trainer.logfile)
            trainer.dst_eval.change_eval_ratio(ratio=dgain)
            results[f'test_x{dgain}'] = trainer.eval(-1)
    if trainer.rank == 0:
        log(f'Metrics have been saved in ./metrics/{trainer.model_name}_metrics.pkl')
    _shutdown(trainer)
    return results


def _shutdown(trainer):
    if trainer.world_size > 1 and torch.distributed.is_initialized():
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def main_lrid(argv=None):
    """trainer_LRID.py:474-505."""
    trainer = IMX686_Trainer(argv)
    if trainer.mode == 'train':
        trainer.train()
        trainer.mode = 'evaltest'                                       # trainer_LRID.py:481
    _load_best_or_make_checkpoint(trainer)
    results = {}
    for mode in ('eval', 'test'):
        if mode in trainer.mode:
            trainer.change_eval_dst(mode)
            for dgain in trainer.dst_eval.args['ratio_list'][:]:
                if trainer.rank == 0:
                    log(f'{trainer.dstname} Datasets: Dgain={dgain}', log=trainer.logfile)
                trainer.dst_eval.change_eval_ratio(ratio=dgain)
                results[f'{mode}_x{dgain}'] = trainer.eval(-1)
    if trainer.rank == 0:
        log(f'Metrics have been saved in ./metrics/{trainer.model_name}_metrics.pkl')
    _shutdown(trainer)
    return results
