"""Import-time stand-in for the `mc3` package (absent in this image, no network).

Test infrastructure only: it lets the unmodified reference package under
/root/reference be imported to generate golden vectors (tests/golden/make_golden.py).
Only `mc3.utils.Log` is exercised on the tli/opacity paths.
"""
from . import utils, plots, stats  # noqa: F401


def sample(*args, **kwargs):
    raise NotImplementedError("mc3 stub: sampling is not available")
