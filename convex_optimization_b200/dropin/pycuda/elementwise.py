# -*- coding: utf-8 -*-
"""``pycuda.elementwise.ElementwiseKernel`` (lasso.py:16,392-398).  PyCUDA compiles the C
snippet with nvcc at run time; the only kernel the reference defines is the soft threshold
``zsoft_t`` (lasso.py:392-398), which lives inside the fused kernel here
(csrc/b200lasso.cu, "prox").  The shim recognises that snippet and runs it on the device
vectors; any other snippet is refused with a message, not silently ignored."""
import re


class ElementwiseKernel:
    def __init__(self, arguments, operation, name="kernel", **_):
        self.arguments = arguments
        self.operation = operation
        self.name = name
        text = re.sub(r"[\s\\]+", "", operation)
        self._soft = text == "soft_t[i]=copysign(1.0,tensor[i])*fmax(fabs(tensor[i])-thres,0.0);"

    def __call__(self, *args, **_):
        if not self._soft:
            raise RuntimeError("pycuda shim: ElementwiseKernel %r is not the reference's zsoft_t snippet; "
                               "run-time compilation of arbitrary C is not part of the B200 drop-in" % self.name)
        import torch
        soft_t, tensor, thres = args
        t = tensor.tensor
        soft_t.tensor.copy_(torch.copysign(torch.ones_like(t), t) * torch.clamp(t.abs() - float(thres), min=0.0))
