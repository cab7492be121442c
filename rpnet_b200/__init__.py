"""rpnet_b200 — B200-native (sm_100a) implementation of the RP-Net hot path.

Host side is Python/PyTorch (device memory, streams, torch.distributed); all arithmetic on the hot path
runs in hand-written CUDA kernels reached through the C ABI declared in include/rpnet_b200.h
(rpnet_b200/lib/librpnet_sm100.so).  There is no CPU fallback: importing the ops without the library
raises.  The reference interface (net/rp_net.py, net/unet.py, net/vgg.py, net/model.py of
uci-cbcl/RP-Net) is mirrored in rpnet_b200.nn and re-exported by the top-level `net` shim package; together with the
`dataset/` and `utils/` shims the reference's own test_rpnet.py runs unmodified on top of this package
(tools/run_reference_driver.py, profiles/r02_reference_driver_test_rpnet.log).
"""
__version__ = '0.2.0'
