"""Drop-in shim for dataset/few_shot_reader.py of the reference (eval readers): see rpnet_b200/dataset/few_shot_reader.py."""
from rpnet_b200.dataset.few_shot_reader import (FewshotRegReader, FewshotSliceReader, FewshotVolumeReader,   # noqa: F401
                                                crop, keep_only_annotation_z_slices, make_support_query_same_size,
                                                train_collate)
from rpnet_b200.registration import get_registration_field                                                   # noqa: F401
