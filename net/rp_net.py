"""Shim for the reference module path net/rp_net.py -> rpnet_b200.nn.rp_net."""
from rpnet_b200.nn.rp_net import *  # noqa: F401,F403
from rpnet_b200.nn import rp_net as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith('__')})
