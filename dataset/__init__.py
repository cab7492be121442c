"""Drop-in shim: the reference imports `dataset.few_shot_reader` (test_rpnet.py).  Everything lives in rpnet_b200.dataset."""
