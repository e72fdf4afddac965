"""Drop-in modules with the call signatures of the reference's compiled extensions
`pyratbay.lib._extcoeff` (src_c/_extcoeff.c:114-123, 375-380, 429-434) and
`pyratbay.lib.vprofile` (src_c/vprofile.c:53-57), so that the UNMODIFIED reference Python
package runs its opacity stage on the GPU engine.

The reference parallelises with fork() (pyrat/extinction.py:108-119, line_by_line.py:231-246)
and a forked child cannot use a CUDA context of its parent, so the shim never touches CUDA in
the calling process: the calls are forwarded to ONE server process per box
(`python -m pyratbay_b200.shim.server`, started on first use) that owns the engines.  Static
arguments (grids, Voigt table, line list) are uploaded once and recognised by content key on
later calls; a call moves only the per-layer scalars in and one spectrum out.

Install: place two one-line modules in the reference's `pyratbay/lib/`:
    _extcoeff.py :  from pyratbay_b200.shim._extcoeff import *
    vprofile.py  :  from pyratbay_b200.shim.vprofile import *
(`install_into(lib_dir)` writes them).
"""
import os


def install_into(lib_dir):
    """Write the two forwarding modules into a reference package's lib/ directory (removing
    compiled _extcoeff / vprofile extensions there, which would take precedence)."""
    for name in ("_extcoeff", "vprofile"):
        for f in os.listdir(lib_dir):
            if f.startswith(name + ".") and f.endswith((".so", ".pyd")):
                os.remove(os.path.join(lib_dir, f))
        with open(os.path.join(lib_dir, name + ".py"), "w") as out:
            out.write(f"from pyratbay_b200.shim.{name} import *  # noqa: F401,F403\n")
