"""Shim for the reference module path net/registration.py (test_rpnet.py:32 imports NCC, MSE) -> rpnet_b200.registration."""
from rpnet_b200.registration import *  # noqa: F401,F403
from rpnet_b200.registration import MSE, NCC  # noqa: F401
