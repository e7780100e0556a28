"""The reference's .bin wire format (big-endian 4-byte, stream, (i,k,j) order, its file names)."""
import os

import numpy as np
import pytest

import wrf_model_cuda_sample_b200 as wrf
from wrf_model_cuda_sample_b200 import binio
from oracle import loader
from tests import cases


def _make_case(tmp_path, variant="specified"):
    g = cases.grid(30, 22, 8, halo=3, variant=variant)
    f = cases.random_fields(g, seed=42)
    inp, gold = str(tmp_path / "in"), str(tmp_path / "gold")
    binio.write_case(inp, g, cases.SCALARS_12KM, f)
    want = cases.copy_fields(f)
    loader.oracle_c(want, g, cases.SCALARS_12KM)
    binio.write_case(gold, g, cases.SCALARS_12KM, want, suffix="_output", names=binio.GOLDEN)
    return g, f, want, inp, gold


def test_wire_format_is_big_endian_fortran_order(tmp_path):
    g, f, _, inp, _ = _make_case(tmp_path)
    raw = open(os.path.join(inp, "grid_u_2.bin"), "rb").read()
    assert len(raw) == 4 * f["u"].size
    # element (i=ims+2, k=kms+1, j=jms+3) sits at ((j*kdim + k)*idim + i)*4, big-endian
    nj, nk, ni = g.shape3
    off = ((3 * nk + 1) * ni + 2) * 4
    assert np.frombuffer(raw[off:off + 4], dtype=">f4")[0] == f["u"][3, 1, 2]
    assert np.frombuffer(open(os.path.join(inp, "ime.bin"), "rb").read(), dtype=">i4")[0] == g.ime
    assert set(os.listdir(inp)) >= {"grid_ww.bin", "ww1.bin", "grid_u_save.bin", "t_tend.bin", "t_2save.bin",
                                    "grid_msfvx_inv.bin", "dts_rk.bin", "config_flags_periodic_x.bin", "kds.bin"}


def test_round_trip_and_report(tmp_path):
    g, f, want, inp, gold = _make_case(tmp_path, "periodic_specified")
    g2, scalars, f2 = binio.read_case(inp)
    assert g2 == g and all(a == b for a, b in zip(scalars, cases.SCALARS_12KM))
    for n in wrf.FIELDS:
        assert np.array_equal(cases.bits(f2[n]), cases.bits(f[n])), n
    rep = binio.compare_with_golden(g, want, gold)
    assert set(rep) == set(binio.GOLDEN) and all(r["n_different"] == 0 and r["max_ulp"] == 0 for r in rep.values())
    rep = binio.compare_with_golden(g, f, gold)                # inputs vs outputs: t must differ
    assert rep["t"]["n_different"] > 0 and rep["ww_1"]["n_different"] == 0


def test_nan_input_is_rejected(tmp_path):
    g, f, _, inp, _ = _make_case(tmp_path)
    bad = f["u"].copy(); bad[0, 0, 0] = np.nan
    binio.write_field(os.path.join(inp, "grid_u_2.bin"), bad)
    with pytest.raises(ValueError):
        binio.read_case(inp)


@pytest.mark.gpu
def test_cli_reproduces_golden_dump(tmp_path, capsys):
    from wrf_model_cuda_sample_b200 import cli
    g, f, want, inp, gold = _make_case(tmp_path)
    out = str(tmp_path / "out")
    assert cli.main([inp, gold, "--write", out]) == 0
    text = capsys.readouterr().out
    assert "worst max-ulp over all fields: 0" in text and "# of non-equal values: 0" in text
    got = binio.read_field(os.path.join(out, "grid_t_2_output.bin"), g.shape3)
    assert np.array_equal(cases.bits(got), cases.bits(want["t"]))
