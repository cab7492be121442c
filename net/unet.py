"""Shim for the reference module path net/unet.py -> rpnet_b200.nn.unet."""
from rpnet_b200.nn.unet import *  # noqa: F401,F403
from rpnet_b200.nn import unet as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith('__')})
