"""Episode builder and NRRD container I/O ("next" row N4 of SURVEY §8f): dataset/few_shot_reader.py for `mode='eval'`."""
from . import nrrd_io                                                   # noqa: F401
from .few_shot_reader import (FewshotRegReader, FewshotSliceReader, FewshotVolumeReader, train_collate)   # noqa: F401
