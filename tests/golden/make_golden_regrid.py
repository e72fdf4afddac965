#!/usr/bin/env python
"""Golden vectors for the p/T re-gridding of cross-section tables: the UNMODIFIED reference's
pyratbay.tools.interpolate_opacity (tools/tools.py:1026-1107) applied to the reference-written
table tests/golden/mock_opacity_file.npz.  Runs only in the build container (needs
/root/reference; same scratch set-up as make_golden.py).  Writes tests/golden/mock_regrid.npz."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden  # noqa: E402


def main():
    make_golden.build_scratch()
    import pyratbay.tools as pt
    import pyratbay.io as io
    path = os.path.join(HERE, "mock_opacity_file.npz")
    _, temp, press, wn = io.read_opacity(path, extract='arrays')
    out = {}
    # both axes: in range, out of range on either side, exact nodes
    t_both = np.array([250.0, 300.0, 450.0, 1234.5, 2999.0, 3000.0, 3500.0])
    p_both = np.array([1e-8, 1e-6, 3.3e-5, 2e-3, 0.7, 50.0, 100.0])
    out["t_both"], out["p_both"] = t_both, p_both
    out["cs_both"] = pt.interpolate_opacity(path, t_both, p_both)
    # temperature only (pressure = table), with a wavenumber mask and thinning
    mask = (wn > wn[10]) & (wn < wn[-5])
    t_only = np.linspace(310.0, 2950.0, 13)
    out["t_only"], out["mask"] = t_only, mask
    out["cs_t_only"] = pt.interpolate_opacity(path, t_only, press, mask, 3)
    # pressure only
    p_only = np.logspace(-5.5, 1.7, 23)
    out["p_only"] = p_only
    out["cs_p_only"] = pt.interpolate_opacity(path, temp, p_only)
    # a table with zeros and denormal-small entries: the -230 floor
    _, species, t0, p0, w0, tab = io.read_opacity(path, extract='all')
    tab = tab.copy()
    tab[::2, 5:9, 20:40] = 0.0
    tab[1, :, 60:70] *= 1e-290
    zpath = "/tmp/pbref/run/zero_table.npz"
    io.write_opacity(zpath, species, t0, p0, w0, tab)
    out["zero_table"] = tab
    out["cs_zero"] = pt.interpolate_opacity(zpath, t_both, p_both)
    # Line_Sample built on re-gridded axes + its text form (line_sampling.py:466-498)
    import pyratbay.opacity as op
    ls = op.Line_Sample(path, pressure=p_both[1:6], temperature=t_both[1:6])
    out["ls_str"] = str(ls).replace(path, "FILE")
    out["ls_cs_table"] = ls.cs_table
    dens = np.full((5, 1), 1.0e12)
    out["ls_temp"] = np.array([320.0, 800.0, 1234.5, 2000.0, 2998.0])
    out["ls_ec"] = ls.calc_extinction_coefficient(out["ls_temp"], dens)
    np.savez_compressed(os.path.join(HERE, "mock_regrid.npz"), **out)
    print({k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
