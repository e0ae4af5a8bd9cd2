"""Runs one 3x3 conv layer through a host build of conv_tc.cu given on the command line and prints `max-abs-error pipeline-error`
(tests/test_device_tc_on_cpu.py::test_model_catches_a_removed_wait runs it in a subprocess: a build with a barrier wait removed
either aborts inside the tensor-core model or computes garbage).  Test infrastructure only."""
import ctypes as C
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pnnp_b200 import _lib, archs  # noqa: E402

lib = C.CDLL(sys.argv[1])
lib.emul_tc_last_error.restype = C.c_char_p


class Fake:
    def pnnp_conv2d_tc_ex(self, d, s):
        return lib.emul_conv2d_tc_ex(C.byref(d))

    def pnnp_last_error(self):
        return lib.emul_tc_last_error()


_lib.lib = lambda: Fake()
_lib.stream_ptr = lambda device=None: None
g = torch.Generator().manual_seed(1)
x = torch.randn((2, 64, 40, 48), generator=g)
wt = torch.randn((64, 64, 3, 3), generator=g) / 24
b = torch.randn((64,), generator=g) * 0.1
bf = lambda t: t.to(torch.bfloat16).float()
m = type("M", (), {})()
m.weight, m.bias = wt, None
out = torch.zeros((2, 40, 48, 64), dtype=torch.bfloat16)
archs._conv(_lib.CONV3, x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16), archs._PackedLayer(m, "conv").get(wt.device)[0], b, out, 64, 1)
ref = F.leaky_relu(F.conv2d(bf(x), bf(wt), b, padding=1), 0.2)
err = (out.float().permute(0, 3, 1, 2) - ref).abs().max().item()
print(f"RESULT {err} {lib.emul_conv_pipeline_error()}")
