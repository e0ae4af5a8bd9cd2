"""The one function of data_process/unprocess.py that lies on the synthetic-pair path: `random_gains`
(unprocess.py:60-77), the white-balance jitter drawn in Raw_Dataset.__getitem__ (syn_datasets.py:313-319).
Host scalars only; the products run on the device (crops.wb_jitter -> csrc/crop_aug.cu)."""
import numpy as np
import torch

# camera -> (range of the red gain, coefficients c0 + c1 r + c2 r^2 of the blue gain as a function of the red gain r)
_WB_FITS = {
    'SonyA7S2': ((1.75, 2.65), (14.65, -9.63942308, 1.80288462)),
    'IMX686': ((1.4, 2.3), (6.14381188, -3.65620261, 0.70205967)),
}


def random_gains(camera_type='SonyA7S2'):
    """-> (rgb_gain, red_gain, blue_gain): float32 tensors of shape (1,).

    Draw order and arithmetic of the reference, so that equal seeds give equal gains (tests/golden/wb_jitter.npz):
      1. rgb_gain = 1 / g with g ~ Normal(0.8, 0.1), ONE sample from torch's global CPU generator, shape (1,) float32;
      2. red_gain ~ np.random.uniform over the camera's range (NumPy's global RandomState, float64);
      3. blue_gain = c0 + c1 * red + c2 * red**2 in float64, evaluated left to right;
      4. red / blue rounded to float32 once.
    Any camera without a fit raises NotImplementedError — after the Normal sample has been consumed, as in the reference."""
    g = torch.distributions.Normal(loc=torch.tensor([0.8]), scale=torch.tensor([0.1])).sample()
    rgb = 1.0 / g
    if camera_type not in _WB_FITS:
        raise NotImplementedError
    (lo, hi), (c0, c1, c2) = _WB_FITS[camera_type]
    red = np.random.uniform(lo, hi)
    blue = c0 + c1 * red + c2 * red ** 2
    as_f32 = lambda v: torch.from_numpy(np.array([v]).astype(np.float32))
    return rgb, as_f32(red), as_f32(blue)
