"""Operator API of the reference's `utils/stylegan2/op` package (op/__init__.py:1-2):
`fused_leaky_relu`, `FusedLeakyReLU`, `upfirdn2d` — backed by librw_b200.so instead of the
two JIT-compiled pybind11 extensions."""
from .fused_act import FusedLeakyReLU, fused_leaky_relu, fused
from .upfirdn2d import upfirdn2d, upfirdn2d_op
