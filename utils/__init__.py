"""Drop-in shim: the reference's eval driver imports `utils.util` (test_rpnet.py:15,27,29)."""
