"""net/model.py:4-7 — name -> class.  'LGCANet_V3' is a different model outside the hot path."""
from .rp_net import RP_Net

model_factory = {
    'RP_Net': RP_Net,
}
