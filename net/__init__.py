"""Drop-in shim: the reference imports `net.model`, `net.rp_net`, ... (test_rpnet.py:11,32).
Everything lives in rpnet_b200.nn."""
