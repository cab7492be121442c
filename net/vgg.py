"""Shim for the reference module path net/vgg.py -> rpnet_b200.nn.vgg."""
from rpnet_b200.nn.vgg import *  # noqa: F401,F403
from rpnet_b200.nn import vgg as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith('__')})
