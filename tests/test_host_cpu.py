"""CPU: the C++ host driver's pieces that need no GPU — INI reader + normalisation (vs the reference's
own parse of the same file), the loader (vs reference golden particles, bit for bit), and the HDF5
writer (read back by the independent reader in tests/h5mini.py)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import picsp_b200
from picsp_b200 import host
from tests import h5mini
from tests.helpers import GOLDEN, load_golden

INI = os.path.join(GOLDEN, "input_ini_shipped.ini")


def test_ini_matches_reference_parse_bit_for_bit():
    want = json.load(open(os.path.join(GOLDEN, "input_ini_parsed.json")))   # produced by the reference's parse_ini_file
    got = host.config_dict(host.parse_ini(INI))
    for k, v in want.items():
        assert got[k] == v, (k, got[k], v)


def test_ini_rules(tmp_path):
    """iniparser's rules: lower-cased keys, value cut at ';' or '#', quotes, strtol base 0 (1e9 -> 1), -1 defaults."""
    p = tmp_path / "t.ini"
    p.write_text("""
# comment
[Time]
NTIMESTEPS = 0x10 ; hex through strtol base 0
timeStep = 1E-10
[grid]
stepSize = "1.2E-4"
numxCells = 1e9;   strtol stops at 'e'
numyCells=010# octal
[population]
nParticlesI = 10
nParticlesE = 10
chargeE=1.602e-19
massE = 9.109e-31
massI = 1.673e-27;
density = 1e12
vthE = 0.9
vthI = 0.026
driftE = 0.2
driftI = 0
loadType = 1
[solver]
solverType = 2
""")
    c = host.parse_ini(str(p))
    assert c.nTimeSteps == 16 and c.numxCells == 1 and c.numyCells == 8
    assert c.dumpPeriod == -1                      # missing key -> the caller's default
    assert abs(c.stepSize - 0.017013464085752668) < 1e-18


def test_ini_sanity_gates_fail_like_the_reference(tmp_path):
    txt = open(INI).read()
    for bad in (txt.replace("solverType = 2", "solverType = 3"), txt.replace("loadType    = 2", "loadType = 0"),
                txt.replace("timeStep = 1E-10", "timeStep = 1E-9"), txt.replace("stepSize = 1.2E-4", "stepSize = 1")):
        p = tmp_path / "bad.ini"
        p.write_text(bad)
        with pytest.raises(picsp_b200.PicspError):
            host.parse_ini(str(p))
    with pytest.raises(picsp_b200.PicspError):
        host.parse_ini(str(tmp_path / "missing.ini"))


@pytest.mark.parametrize("name", ["loop_sor_65_load2_O0", "loop_sor_33_load1", "loop_spectral_48x_load1"])
def test_loader_bit_identical_to_reference(name):
    g = load_golden(name)
    numx, n, solver, load_type, _ = (int(v) for v in g["meta"])
    cfg = host.parse_ini(INI)
    cfg.numxCells = cfg.numyCells = numx
    cfg.nParticlesI = cfg.nParticlesE = n
    cfg.loadType = load_type
    cfg.driftE = float(g["params"][5])
    ions, electrons = host.load_species(cfg, seed=0)
    assert np.array_equal(np.stack(ions), g["loaded/part_i"])
    assert np.array_equal(np.stack(electrons), g["loaded/part_e"])


def test_h5_writer_roundtrip(tmp_path):
    L = picsp_b200.load_library()
    path = str(tmp_path / "t.h5")
    h = L.picsp_host_h5_open(path.encode())
    assert h
    rng = np.random.default_rng(0)
    data = {}
    for g in ("/particle.e", "/particle.i", "/timedata", "/phi", "/den.e", "/den.i"):
        assert L.picsp_host_h5_group(h, g.encode()) == 0
    for ts in range(0, 10001, 50):                      # 201 datasets per group: names sort as strings, not numbers
        a = rng.standard_normal((5, 3))
        data[f"/phi/{ts}"] = a
        assert L.picsp_host_h5_dataset_f64(h, f"/phi/{ts}".encode(), a.ctypes.data_as(C.POINTER(C.c_double)), 5, 3) == 0
    e = rng.standard_normal((201, 2)); data["/timedata/energy"] = e
    assert L.picsp_host_h5_dataset_f64(h, b"/timedata/energy", e.ctypes.data_as(C.POINTER(C.c_double)), 201, 2) == 0
    assert L.picsp_host_h5_attr_f64(h, b"Lx", 1.0888617) == 0
    assert L.picsp_host_h5_attr_i32(h, b"Nx", 65) == 0
    assert L.picsp_host_h5_close(h) == 0

    f = h5mini.File(path)
    assert f.attrs()["Lx"] == 1.0888617 and f.attrs()["Nx"] == 65 and f.attrs()["Nx"].dtype == np.int32
    assert sorted(f.groups()) == sorted(["particle.e", "particle.i", "timedata", "phi", "den.e", "den.i"])
    assert len(f.datasets("phi")) == 201 and f.datasets("den.e") == []
    for k, a in data.items():
        assert np.array_equal(f.read(k), a), k
    assert open(path, "rb").read(8) == b"\x89HDF\r\n\x1a\n"


TWISTED = "/root/reference/lib/iniparser/test/twisted.ini"


@pytest.mark.skipif(not os.path.isfile(TWISTED), reason="iniparser's own fixture lives in the reference tree (build container only)")
def test_ini_reader_agrees_with_iniparser_on_its_own_twisted_fixture():
    """The reference's vendored iniparser 3.1 ships `twisted.ini`, 131 lines of blank / comment / quote / trailing-';'
    edge cases.  Every `section:key` that the reference's parser (compiled from the reference tree into
    oracle/_ref/libpicsp_ref.so) reports must come out of the product's own reader with the same value."""
    from oracle.oracle import REF_SO
    if not os.path.isfile(REF_SO):
        pytest.skip("oracle/_ref not built")
    R = C.CDLL(REF_SO)
    R.iniparser_load.argtypes = [C.c_char_p]; R.iniparser_load.restype = C.c_void_p
    R.iniparser_getstring.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]; R.iniparser_getstring.restype = C.c_char_p
    R.iniparser_freedict.argtypes = [C.c_void_p]
    R.iniparser_getnsec.argtypes = [C.c_void_p]; R.iniparser_getnsec.restype = C.c_int
    R.iniparser_getsecname.argtypes = [C.c_void_p, C.c_int]; R.iniparser_getsecname.restype = C.c_char_p
    R.iniparser_getsecnkeys.argtypes = [C.c_void_p, C.c_char_p]; R.iniparser_getsecnkeys.restype = C.c_int
    R.iniparser_getseckeys.argtypes = [C.c_void_p, C.c_char_p]; R.iniparser_getseckeys.restype = C.POINTER(C.c_char_p)
    d = R.iniparser_load(TWISTED.encode())
    assert d
    want = {}
    for s in range(R.iniparser_getnsec(d)):
        sec = R.iniparser_getsecname(d, s)
        n = R.iniparser_getsecnkeys(d, sec)
        keys = R.iniparser_getseckeys(d, sec)
        for k in range(n):
            key = keys[k]
            want[key.decode()] = R.iniparser_getstring(d, key, b"<missing>").decode()
    R.iniparser_freedict(d)
    assert len(want) >= 30
    L = picsp_b200.load_library()
    need = L.picsp_host_ini_dump(TWISTED.encode(), None, 0)
    assert need > 0
    buf = C.create_string_buffer(need)
    assert L.picsp_host_ini_dump(TWISTED.encode(), buf, need) == need
    got = dict(line.split("\t", 1) for line in buf.value.decode().split("\n") if "\t" in line)
    bad = {k: (got.get(k), v) for k, v in want.items() if got.get(k) != v}
    assert not bad, f"reader differs from iniparser on {len(bad)} of {len(want)} keys: {dict(list(bad.items())[:8])}"
