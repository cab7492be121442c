"""Shim for the reference module path net/modules.py -> rpnet_b200.nn.modules."""
from rpnet_b200.nn.modules import *  # noqa: F401,F403
from rpnet_b200.nn import modules as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith('__')})
