"""reference src/utils/Util.py:5-16."""
import numpy as np


def split_row(row_content):
    for sep in (',', ';'):
        if sep in row_content:
            return row_content.strip().split(sep)
    return row_content.strip().split()


def matBinarize(sR, r_threshold):
    """``(sR > thr).astype(float32)`` on a scipy sparse matrix (Util.py:15-16)."""
    return (sR > r_threshold).astype(np.float32)
