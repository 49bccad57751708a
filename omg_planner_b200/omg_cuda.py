"""Drop-in for the reference's PyTorch extension module `omg_cuda` (layers/omg_layers.cpp:47-49,
built by layers/setup.py:7-13): one function, identical signature and tensor contract."""
from .engine import sdf_loss_forward  # noqa: F401
