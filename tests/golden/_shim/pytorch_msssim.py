"""Stand-in for the third-party `pytorch_msssim==0.2.1` (reference requirements.txt:3), which is not
installed in the build container.  Only used by make_golden.py so that the UNMODIFIED reference utils.py
imports; it forwards to the oracle's restatement of the published algorithm."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "..")))
from oracle import nerv_oracle as _o  # noqa: E402


def ssim(X, Y, data_range=1, size_average=True, **kw):
    assert data_range == 1 and size_average
    return _o.ssim(X, Y)


def ms_ssim(X, Y, data_range=1, size_average=True, **kw):
    assert data_range == 1 and size_average
    return _o.ms_ssim(X, Y)
