"""Synthetic stand-ins for the reference's eval datasets (real_datasets.py ELD_Dataset / SID_Dataset,
phone_datasets.py IMX686_Dataset), which need 560 GB of RAW files that are out of scope here.

Same constructor (`Dataset(yaml_dict)`), same item dict keys (lr hr ratio wb ccm name ISO
ExposureTime) and the same sweep controls (`ratio_list`, `recheck_length`, `change_eval_ratio`).
The clean frame is a seeded dark-scene frame; the noisy frame is produced ON THE DEVICE by the fused
synthesis kernel from per-ISO calibrated parameters (sample_params_max(camera, ratio, iso)), i.e. the
P1 -> S3 -> N1-N3 chain of the hot path.  Items carry host tensors for hr only; `lr` is filled by the
trainer's preprocess on the GPU."""
import numpy as np
import torch

from .noise_params import sample_params_max


class _SyntheticEvalBase(torch.utils.data.Dataset):
    default_frames = 10

    def __init__(self, args=None):
        self.args = dict(args)
        self.H, self.W = self.args['H'], self.args['W']
        self.h, self.w, self.c = self.H // 2, self.W // 2, 4
        self.ratio_list = list(self.args.get('ratio_list', [100]))
        self.iso_list = list(self.args.get('iso_list', [self.default_iso]))
        self.n_scenes = int(self.args.get('synthetic_frames', self.default_frames))
        self.recheck_length()

    def recheck_length(self):
        self.items = [(s, iso, r) for r in self.ratio_list for s in range(self.n_scenes) for iso in self.iso_list]
        self.length = len(self.items)

    def change_eval_ratio(self, ratio):
        self.ratio_list = [ratio]
        self.recheck_length()

    def __len__(self):
        return self.length

    def eval_crop(self, data, base=64):
        """SynBase_Dataset.eval_crop (syn_datasets.py:109-133): overlapped `patch_size` tiles of a (1,c,h,w) CUDA frame."""
        from . import crops
        return crops.eval_crop(data, self.args["patch_size"], base)

    def eval_merge(self, croped_data, base=64):
        """SynBase_Dataset.eval_merge (syn_datasets.py:135-159)."""
        from . import crops
        return crops.eval_merge(croped_data, self.h, self.w, base)

    def clean_frame(self, scene):
        g = torch.Generator().manual_seed(1997 + scene)
        return torch.rand((self.c, self.h, self.w), generator=g) ** 2

    def __getitem__(self, idx):
        scene, iso, ratio = self.items[idx]
        rs = np.random.RandomState(7919 * scene + iso)
        state = np.random.get_state()
        np.random.set_state(rs.get_state())
        try:
            iso_arg = iso if self.args['camera_type'] in ("SonyA7S2", "IMX686") and iso in self.legal_iso else None
            param = sample_params_max(self.args['camera_type'], ratio=ratio, iso=iso_arg)
        finally:
            np.random.set_state(state)
        hr = self.clean_frame(scene)
        return {"hr": hr[None], "lr": hr[None].clone(), "ratio": np.float32(ratio), "wb": np.ones(4, np.float32),
                "ccm": np.eye(3, dtype=np.float32), "name": f"{self.args.get('dstname', 'syn')}_s{scene:03d}_iso{iso}_x{ratio}",
                "ISO": iso, "ExposureTime": 1.0, "param": param, "index": idx}


class Synthetic_ELD_Dataset(_SyntheticEvalBase):
    """ELD-shaped sweep: 10 scenes x iso_list, ratio_list [100, 200] (real_datasets.py:333-338)."""
    default_frames, default_iso = 10, 1600
    legal_iso = (50, 64, 80, 100, 125, 160, 200, 250, 320, 400, 500, 640, 800, 1000, 1250, 1600, 2000, 2500, 3200,
                 4000, 5000, 6400, 8000, 10000, 12800, 16000, 20000, 25600)


class Synthetic_SID_Dataset(Synthetic_ELD_Dataset):
    """SID-shaped sweep: 40 frames per ratio (real_datasets.py:324)."""
    default_frames = 40


class Synthetic_IMX686_Dataset(_SyntheticEvalBase):
    """LRID-shaped sweep: 9 scenes (phone_datasets.py:245), 4x1736x2312 frames."""
    default_frames, default_iso = 9, 6400
    legal_iso = (100, 6400)


class Raw_Dataset(torch.utils.data.Dataset):
    """Raw_Dataset.__getitem__ (data_process/syn_datasets.py:285-347) with the whole item built on the
    device: synthetic uint16 RAW frame -> raw2bayer(norm, clip) -> random_crop + 8-mode aug ->
    per-crop sample_params -> ONE fused synthesis launch -> lr.clip(lb, 1), hr.clip(0, 1).
    Returns CUDA tensors (lr, hr: crop_per_image x 4 x patch x patch; ratio: crop_per_image).
    The clean long-exposure RAW files of the reference are replaced by seeded synthetic frames."""

    def __init__(self, args=None):
        self.args = dict(args)
        self.H, self.W = self.args['H'], self.args['W']
        self.h, self.w, self.c = self.H // 2, self.W // 2, 4
        self.length = int(self.args.get('synthetic_frames', 16))
        self.args.setdefault('params', None)
        self.args.setdefault('mode', 'train')

    def __len__(self):
        return self.length

    @staticmethod
    def device():
        """Items are built on the current CUDA device (there is no CPU path)."""
        from . import _lib
        return _lib.cuda_device()

    def synthetic_raw(self, idx, device):
        g = torch.Generator(device=device).manual_seed(4242 + idx)
        wp, bl = self.args['wp'], self.args['bl']
        u = torch.rand((self.H, self.W), device=device, generator=g)
        return (bl + (u * u) * (wp - bl)).to(torch.int16)          # bit pattern of the uint16 sensor codes (< 2^15)

    def __getitem__(self, idx):
        from . import crops
        from .isp_ops import raw2bayer
        from .noise import synthesize_batch
        from .noise_params import HALF_CLIP, sample_params
        device = self.device()
        a = self.args
        hr_imgs = raw2bayer(self.synthetic_raw(idx, device), wp=a['wp'], bl=a['bl'], norm=True, clip=True)
        if a['mode'] == 'train':
            hs, ws, aug = crops.init_random_crop_point(self.h, self.w, a['patch_size'], a['crop_per_image'], a['croptype'])
            hr_crops = crops.random_crop(hr_imgs, hs, ws, aug, a['patch_size'])
        else:
            hr_crops = hr_imgs[None]
        n = hr_crops.shape[0]
        wb = np.ones(4, np.float32)
        if a["lock_wb"] is False and np.random.randint(2):          # syn_datasets.py:313-319: white-balance jitter, one item in two
            from .unprocess import random_gains
            crops.wb_jitter(hr_crops, wb, random_gains())
        post = None
        if a['clip']:
            post = (-float("inf") if a['clip'] == HALF_CLIP else 0.0, 1.0)
        if a['gpu_preprocess'] is False:                            # syn_datasets.py:325-337: the noise is added by the dataset
            params = [sample_params(camera_type=a['camera_type']) if a['params'] is None else a['params'] for _ in range(n)]
            # synthesised from the crops as they are (a jittered crop may exceed 1); hr is clipped afterwards (:339-342)
            lr_crops = synthesize_batch(hr_crops, params, a['noise_code'], ori=a['ori'], post_clip=post)
            ratio = torch.tensor([float(p['ratio']) for p in params], dtype=torch.float32, device=device)
        else:                                                       # noise left to the trainer's preprocess (trainer_SID.py:449-462)
            lr_crops = hr_crops.clone()
            if post is not None:
                lr_crops.clamp_(post[0], post[1])
            ratio = torch.ones(n, dtype=torch.float32, device=device)
        if a['clip']:
            hr_crops = hr_crops.clamp_(0, 1)
        return {"lr": lr_crops, "hr": hr_crops, "ratio": ratio, "wb": wb,
                "ccm": np.eye(3, dtype=np.float32), "name": f"syn_{idx:04d}"}
