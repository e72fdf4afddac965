"""The drop-in shown live: the UNMODIFIED reference Python package (a scratch copy beside
oracle/_ref, built by oracle/build_ref.sh) runs `pb.run(opacity.cfg)`, its forked
Line_By_Line.calc_extinction_coefficient and its Line_Sample with lib/_extcoeff and
lib/vprofile replaced by pyratbay_b200/shim, i.e. on the GPU engine.  Outputs are compared with
the goldens the same package produced with its own C modules (tests/golden/make_golden.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu

SHIMMED = os.path.join(helpers.ROOT, "oracle", "_ref", "shimmed")


def _peak_err(got, want):
    peak = np.max(np.abs(want), axis=-1, keepdims=True)
    peak[peak == 0] = 1.0
    return np.max(np.abs(got - want) / peak)


def test_unmodified_reference_runs_on_the_engine(tmp_path):
    if not os.path.isdir(os.path.join(SHIMMED, "pyratbay")):
        pytest.skip("oracle/_ref/shimmed not built (oracle/build_ref.sh needs /root/reference)")
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join(
        [os.path.join(helpers.ROOT, "oracle", "refstubs"), SHIMMED, helpers.ROOT])
    env.pop("PB200_SHIM_ADDR", None)
    res = subprocess.run([sys.executable, os.path.join(helpers.ROOT, "tests", "dropin_worker.py"),
                          helpers.GOLDEN, str(tmp_path)], env=env, capture_output=True, text=True,
                         timeout=900)
    assert res.returncode == 0 and "DROPIN OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
    got = np.load(tmp_path / "dropin_results.npz", allow_pickle=True)

    table = helpers.golden("mock_opacity_table.npz")
    assert got["etable"].shape == table["etable"].shape == (10, 51, 100)
    assert _peak_err(got["etable"], table["etable"]) < 1e-10
    assert np.array_equal(got["file_etable"], got["etable"])
    assert _peak_err(got["etable_R"], helpers.golden("mock_opacity_table_R.npz")["etable"]) < 1e-10

    voigt = helpers.golden("mock_voigt.npz")
    assert np.array_equal(got["voigt_size"], voigt["size"])
    assert np.array_equal(got["voigt_index"], voigt["index"])
    np.testing.assert_allclose(got["profile_strided"], voigt["profile_strided"], rtol=1e-12)

    fwd = helpers.golden("mock_forward.npz")
    assert _peak_err(got["ec_all"], fwd["ec_all"]) < 1e-10
    assert _peak_err(got["ec_layer31"], fwd["ec_layer31"]) < 1e-10
    assert np.all(got["ec_skip"] == 0.0)

    ls = helpers.golden("mock_line_sample.npz")
    # the table under the reference's Line_Sample is the engine's (<= 1e-10 of the peak from the
    # golden one), the interpolation itself is bit-exact (tests/test_gpu_parity.py)
    for key, name in (("ls_cs", "cs"), ("ls_cs_per_mol", "cs_per_mol"), ("ls_ec", "ec"),
                      ("ls_ec_layer", "ec_layer")):
        assert _peak_err(got[key], ls[name]) < 1e-10
    print(f"drop-in: table of {int(got['n_units'])} units in {float(got['table_s']):.2f} s, "
          f"forward model {float(got['forward_s']):.2f} s, "
          f"{float(got['per_call_s']) * 1e3:.2f} ms per ec.extinction call; "
          f"server {got['server_stats']}")
