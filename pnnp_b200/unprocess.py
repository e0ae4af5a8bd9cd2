"""The one function of data_process/unprocess.py that lies on the synthetic-pair path: `random_gains`
(unprocess.py:60-77), the white-balance jitter drawn in Raw_Dataset.__getitem__ (syn_datasets.py:313-319).
Host scalars only; the products run on the device (crops.wb_jitter -> csrc/crop_aug.cu)."""
import numpy as np
import torch
import torch.distributions as tdist


def random_gains(camera_type='SonyA7S2'):
    """-> (rgb_gain, red_gain, blue_gain), float32 tensors of shape (1,), with the reference's draw order: one
    Normal(0.8, 0.1) sample from torch's global CPU generator, then one np.random.uniform for the red gain; the blue gain is
    the camera's quadratic fit of the red gain."""
    n = tdist.Normal(loc=torch.tensor([0.8]), scale=torch.tensor([0.1]))
    rgb_gain = 1.0 / n.sample()
    if camera_type == 'SonyA7S2':
        red_gain = np.random.uniform(1.75, 2.65)
        fit = (14.65, -9.63942308, 1.80288462)
    elif camera_type == 'IMX686':
        red_gain = np.random.uniform(1.4, 2.3)
        fit = (6.14381188, -3.65620261, 0.70205967)
    else:
        raise NotImplementedError
    blue_gain = fit[0] + fit[1] * red_gain + fit[2] * red_gain ** 2
    red_gain = torch.FloatTensor(np.array([red_gain])).view(1)
    blue_gain = torch.FloatTensor(np.array([blue_gain])).view(1)
    return rgb_gain, red_gain, blue_gain
