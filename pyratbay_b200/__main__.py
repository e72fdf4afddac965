"""Command-line entry: `python -m pyratbay_b200 -c opacity.cfg` (the `pbay -c` of the
reference, pyratbay/__main__.py:11-127, restricted to runmode = tli and runmode = opacity)."""
import argparse
import sys

from . import __version__
from .pyrat import run


def main():
    parser = argparse.ArgumentParser(
        prog="pyratbay_b200",
        description="B200-native cross-section table builder (runmode = opacity) and TLI "
                    "writer (runmode = tli).")
    parser.add_argument("-c", dest="cfile", required=True,
                        help="Pyrat Bay configuration file ([pyrat] section)")
    parser.add_argument("--device", type=int, default=0, help="CUDA device index")
    parser.add_argument("-v", "--version", action="version",
                        version=f"pyratbay_b200 {__version__}")
    args = parser.parse_args()
    pyrat = run(args.cfile, device=args.device)
    if pyrat is not None and pyrat.inputs.runmode != "opacity":
        sys.exit(f"runmode '{pyrat.inputs.runmode}' is outside the scope of pyratbay_b200 "
                 "(only 'tli' and 'opacity' are implemented)")


if __name__ == "__main__":
    main()
