"""Shim for the reference module path net/model.py -> rpnet_b200.nn.model."""
from rpnet_b200.nn.model import *  # noqa: F401,F403
from rpnet_b200.nn import model as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith('__')})
