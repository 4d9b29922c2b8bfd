"""Host-side pieces of the CLI drop-in: MGF IO and the settings parser."""
import io

import numpy as np
import pytest

from falcon_b200 import config as fconfig
from falcon_b200 import synth
from falcon_b200.ms_io import mgf_io


def test_mgf_round_trip(tmp_path):
    sp = synth.generate(300, 3)
    dicts = sp.as_dicts()
    dicts[5]["precursor_charge"] = None  # absent charge is allowed (mgf_io.py:55-58)
    path = str(tmp_path / "s.mgf")
    mgf_io.write_spectra(path, dicts)
    back = list(mgf_io.get_spectra(path))
    assert len(back) == 300
    for a, b in zip(dicts, back):
        assert b["identifier"] == a["identifier"] and b["filename"] == path
        assert b["precursor_charge"] == a["precursor_charge"]
        assert b["precursor_mz"] == pytest.approx(a["precursor_mz"], abs=0) and b["retention_time"] == a["retention_time"]
        assert np.array_equal(b["mz"], a["mz"]) and np.array_equal(b["intensity"], a["intensity"])
    ss, ids, files = mgf_io.read_mgf(path)
    assert len(ss) == 300 and ids[7] == "synth:7" and ss.precursor_charge[5] == 0
    assert np.array_equal(ss.mz, sp.mz)


def test_mgf_parser_tolerates_real_world_files():
    text = """# comment
BEGIN IONS
TITLE=spec one
PEPMASS=500.25 12345.6
CHARGE=2+
RTINSECONDS=12.5
SCANS=17
300.5 10
100.25 5.5
200.0\t7
END IONS

BEGIN IONS
TITLE=broken
PEPMASS=abc
100 1
END IONS
BEGIN IONS
TITLE=no charge
PEPMASS=400.0
END IONS
"""
    spectra = list(mgf_io.get_spectra(io.StringIO(text)))
    assert [s["identifier"] for s in spectra] == ["spec one", "no charge"]  # the unparsable one is skipped
    s = spectra[0]
    assert s["precursor_mz"] == 500.25 and s["precursor_charge"] == 2 and s["retention_time"] == 12.5
    assert s["mz"].tolist() == [100.25, 200.0, 300.5] and s["intensity"].tolist() == [5.5, 7.0, 10.0]  # sorted by m/z
    assert spectra[1]["precursor_charge"] is None and spectra[1]["retention_time"] == -1 and len(spectra[1]["mz"]) == 0


def test_config_defaults_ini_and_command_line(tmp_path):
    cfg = fconfig.Config()
    with pytest.raises(RuntimeError):
        cfg.eps
    cfg.parse(["in.mgf", "out"])
    # falcon's defaults (/root/reference/falcon/config.py:77-183)
    assert (cfg.eps, cfg.low_dim, cfg.n_neighbors, cfg.n_neighbors_ann, cfg.n_probe) == (0.1, 400, 64, 128, 32)
    assert cfg.precursor_tol == [20.0, "ppm"] and cfg.fragment_tol == 0.05 and cfg.rt_tol is None
    assert (cfg.min_peaks, cfg.min_mz_range, cfg.min_mz, cfg.max_mz) == (5, 250.0, 101.0, 1500.0)
    assert (cfg.remove_precursor_tol, cfg.min_intensity, cfg.max_peaks_used, cfg.scaling) == (1.5, 0.01, 50, "off")
    ini = tmp_path / "my.ini"
    ini.write_text("# settings\neps = 0.25\nprecursor_tol = 0.5 Da\nexport_representatives = true\nscaling = rank\n")
    cfg.parse(["-c", str(ini), "a.mgf", "b.mgf", "out", "--eps", "0.3"])
    assert cfg.eps == 0.3 and cfg.precursor_tol == [0.5, "Da"] and cfg.export_representatives and cfg.scaling == "rank"
    assert cfg.input_filenames == ["a.mgf", "b.mgf"] and cfg["output_filename"] == "out"
    with pytest.raises(ValueError):
        cfg.parse(["in.mgf", "out", "--precursor_tol", "20", "Th"])
    with pytest.raises(ValueError):
        cfg.parse(["in.mgf", "out", "--n_neighbors", "64", "--n_neighbors_ann", "32"])
    cfg.parse(["in.mgf", "out", "--distance_threshold", "0.15"])
    assert cfg.eps == 0.15


def test_config_refuses_the_development_heads_flags_and_maps_its_alias():
    cfg = fconfig.Config()
    cfg.parse(["a.mgf", "out", "--distance_threshold", "0.25", "--singletons_as_clusters"])
    assert cfg.eps == 0.25 and cfg.singletons_as_clusters
    for flag, val in (("--linkage", "complete"), ("--min_matched_peaks", "6")):
        with pytest.raises(ValueError, match="hierarchical-clustering pipeline"):
            fconfig.Config().parse(["a.mgf", "out", flag, val])


def test_cli_refuses_existing_outputs_before_doing_any_work(tmp_path):
    """falcon.py:90-122: an existing .csv (or .mgf with --export_representatives) without --overwrite is an
    error (return code 1), checked before anything is read or computed -- no GPU is needed to get there."""
    from falcon_b200 import falcon as fmain

    out = tmp_path / "res"
    (tmp_path / "res.csv").write_text("old")
    assert fmain.main([str(tmp_path / "missing.mgf"), str(out)]) == 1
    (tmp_path / "res.csv").unlink()
    (tmp_path / "res.mgf").write_text("old")
    assert fmain.main([str(tmp_path / "missing.mgf"), str(out), "--export_representatives"]) == 1
    assert (tmp_path / "res.mgf").read_text() == "old"


def test_console_script_is_declared_like_the_reference():
    """setup.cfg:43-45 of the reference declares `falcon = falcon.falcon:main`; pyproject.toml declares the same
    command on this package's main, and that object exists and takes an argument list."""
    import inspect
    import os
    import tomllib

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "pyproject.toml"), "rb") as fh:
        meta = tomllib.load(fh)
    assert meta["project"]["scripts"]["falcon"] == "falcon_b200.falcon:main"
    from falcon_b200 import falcon as fmain

    assert list(inspect.signature(fmain.main).parameters) == ["args"]
