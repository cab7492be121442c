"""Shim for the reference module path utils/util.py -> rpnet_b200.utils.util."""
from rpnet_b200.utils.util import *  # noqa: F401,F403
from rpnet_b200.utils.util import Logger, dice_score_seperate, load_yaml  # noqa: F401
