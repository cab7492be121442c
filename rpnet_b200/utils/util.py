"""The helpers of the reference's utils/util.py that its eval driver imports (test_rpnet.py:15,27,29): Logger, load_yaml and
dice_score_seperate, with the reference's signatures and behaviour (utils/util.py:63-88, 379-390).  Host-side plumbing only;
the DICOM / plotting / NMS utilities of that file are out of scope."""
import sys

import yaml


class Logger(object):
    """utils/util.py:63-76: tee for sys.stdout (test_rpnet.py:105 `sys.stdout = Logger(logfile)`)."""

    def __init__(self, logfile):
        self.terminal = sys.stdout
        self.log = open(logfile, 'a')

    def write(self, message):
        self.terminal.write(message)
        self.log.write(message)

    def flush(self):
        pass


def load_yaml(path):
    """utils/util.py:79-88: (dict, attribute-style view of the same dict)."""
    class Struct:
        def __init__(self, **entries):
            self.__dict__.update(entries)
    with open(path) as f:
        data_dict = yaml.load(f, Loader=yaml.FullLoader)
    return data_dict, Struct(**data_dict)


def dice_score_seperate(y_pred, y_true, num_class=1, decimal=4):
    """utils/util.py:379-390 on numpy arrays (the host-side rule; the device form is rpnet_b200.volume.dice_sums):
    per class 2 * sum(t * p) / (sum(t) + sum(p)) rounded to `decimal` places, None when the target is empty."""
    res = []
    for i in range(num_class):
        target, pred = y_true[i], y_pred[i]
        if target.sum():
            res.append(round(2 * (target * pred).sum() / float(target.sum() + pred.sum()), decimal))
        else:
            res.append(None)
    return res
