"""CPU: the C-ABI library loads, exports every symbol include/picsp_b200.h declares, and fails
loudly (no CPU fallback) when there is no CUDA device.  No compute calls are made here."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import picsp_b200
from picsp_b200 import lib as pl


def test_library_is_in_tree_and_loads():
    if not os.path.isfile(pl.LIB_PATH):
        from picsp_b200 import build
        build.build()
    L = picsp_b200.load_library()
    assert L.picsp_abi_version() == 1


def test_every_declared_symbol_is_exported_and_bound():
    L = picsp_b200.load_library()
    declared = picsp_b200.abi_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/picsp_b200.h but not exported"
    assert set(L._picsp_signatures) == set(declared), "python binding out of sync with the header"
    out = subprocess.run(["nm", "-D", "--defined-only", pl.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert set(declared) <= exported


def test_header_has_no_foreign_types():
    import re
    text = re.sub(r"/\*.*?\*/", "", open(pl.HEADER).read(), flags=re.S)   # declarations only, comments stripped
    for bad in ("torch", "at::", "cudaStream_t", "#include <cuda"):
        assert bad not in text


def test_params_struct_layout_matches_header():
    # int32 x2, f64 x2, int32 x2, f64[2] x3, int64[2], int32 x2  ->  104 bytes, natural alignment
    assert C.sizeof(pl.CParams) == 104
    assert pl.CParams.stepSize.offset == 8 and pl.CParams.charge.offset == 32 and pl.CParams.capacity.offset == 80


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    with pytest.raises(picsp_b200.PicspError) as ei:
        picsp_b200.Simulation(picsp_b200.Params(16, 16, 0.017, 0.005, 1836.0, 10, 10))
    assert ei.value.code == -2          # PICSP_ERR_NO_DEVICE


def test_product_never_touches_the_oracle():
    """The product tree must not import, link or mention oracle/ (it is test infrastructure)."""
    pkg = os.path.dirname(pl.LIB_PATH)
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                src = open(os.path.join(d, f), errors="replace").read()
                assert "oracle" not in src.lower(), os.path.join(d, f)
    out = subprocess.run(["ldd", pl.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "picsp_ref" not in out


@pytest.mark.parametrize("define", ["-DPICSP_BULK_PIPE=0", "-DPICSP_REPL=4"])
def test_ab_build_switches_still_compile(define, tmp_path):
    """The A/B switches documented in DESIGN.md section 3.1 (register-prefetch mover, replicated window) must keep
    compiling for sm_100a, or the recorded sweeps cannot be repeated."""
    import subprocess
    from picsp_b200 import build as b
    out = tmp_path / "abi.o"
    cmd = [b.NVCC, *b.ARCH, "-O1", "-std=c++17", "-Xcompiler", "-fPIC", "-I", os.path.join(b.ROOT, "include"), define,
           "-c", os.path.join(b.CSRC, "abi.cu"), "-o", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]


# ---- the boundary proof (oracle/gpu_patch.py + oracle/ref_gpu_harness.cpp) -------------------------------------
REFERENCE_MAIN = "/root/reference/src/main.cpp"


@pytest.mark.skipif(not os.path.isfile(REFERENCE_MAIN), reason="needs the reference tree (build container only)")
def test_integration_patch_applies_to_the_reference_and_binds_only_the_abi(tmp_path):
    """INTEGRATION.md section 2 as code: every anchor of the patch is found in the reference's main(), the patched TU
    compiles against the shims + include/picsp_b200.h, and the only picsp_* symbols it needs are exported by the
    product library."""
    from oracle import gpu_patch, oracle as orc
    patched = gpu_patch.patch(open(REFERENCE_MAIN).read())
    for call in ("picsp_create", "picsp_species_upload", "picsp_deposit(gpu, 0)", "picsp_compute_rho", "picsp_solve_spectral",
                 "picsp_solve_sor", "picsp_compute_ef", "picsp_push(gpu, 1)", "picsp_rewind(gpu, 0)", "picsp_step(gpu, 1)",
                 "picsp_species_download_rows", "picsp_grid_download", "picsp_compute_ke", "picsp_destroy"):
        assert call in patched, call
    body = patched[patched.index("int main(int argc"):patched.index("void init(Species")]
    for gone in ("scatterSpecies(&ions);", "pushSpecies(&ions, efx, efy);", "computeRho(rho, &ions, &electrons);"):
        assert gone not in body.replace("/* ", "").replace("//", ""), gone
    orc.build()
    for so in (orc.REF_GPU_SO, orc.REF_GPU_FUSED_SO):
        assert os.path.isfile(so)
        out = subprocess.run(["nm", "-D", "--undefined-only", so], capture_output=True, text=True, check=True).stdout
        needed = {ln.split()[-1] for ln in out.splitlines() if "picsp_" in ln}
        assert needed and needed <= set(picsp_b200.abi_symbols()), needed
    fused = subprocess.run(["nm", "-D", "--undefined-only", orc.REF_GPU_FUSED_SO], capture_output=True, text=True).stdout
    assert "picsp_step" in fused


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a GPU")
def test_patched_reference_fails_loudly_without_a_gpu(tmp_path):
    from oracle import oracle as orc
    if not os.path.isfile(orc.REF_GPU_SO):
        pytest.skip("oracle/_ref/libpicsp_ref_gpu.so not built")
    ini = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "input_ini_shipped.ini")
    code = ("import ctypes as C, os, sys; os.makedirs('output', exist_ok=True); L = C.CDLL(sys.argv[1]); "
            "L.picsp_refgpu_main.argtypes = [C.c_char_p]; sys.exit(L.picsp_refgpu_main(os.fsencode(sys.argv[2])))")
    r = subprocess.run([os.sys.executable, "-c", code, orc.REF_GPU_SO, ini], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode != 0 and "no CUDA device" in r.stderr


def test_own_dft_plan_selection():
    """Host logic of the library's own DFT (no device needed): how a node count is split.  kind 1 = two direct coprime
    factors <= 64, kind 0 = small direct factor x Bluestein on a power of two >= 2Q - 1, kind -1 = too long for shared memory."""
    import ctypes as C
    L = picsp_b200.load_library()

    def plan(M):
        out = (C.c_int32 * 5)()
        assert L.picsp_fft_plan_query(M, out) == 0
        return tuple(out)

    assert plan(1025) == (1, 25, 41, 0, 41)            # the bench grid: 25 x 41, both direct
    assert plan(513) == (1, 19, 27, 0, 19)
    assert plan(65) == (1, 5, 13, 0, 13)
    assert plan(48) == (1, 3, 16, 0, 3)
    assert plan(2049) == (0, 3, 683, 2048, 683)        # BASELINE config 5: 3 x Bluestein(683) on 2048 points
    assert plan(257) == (0, 1, 257, 1024, 257)         # prime
    assert plan(49) == (0, 1, 49, 128, 7)              # 7^2: no coprime split
    assert plan(4097) == (0, 17, 241, 512, 241)
    assert plan(8193)[0] == -1                         # 3 x 2731: 3 x 8192 complex do not fit
    for M in (33, 34, 97, 129, 131, 201, 256, 1000, 1537):
        kind, P, Q, Lp, lpf = plan(M)
        assert kind in (0, 1) and P * Q == M and np.gcd(P, Q) == 1
        if kind == 0:
            assert Lp >= 2 * Q - 1 and Lp & (Lp - 1) == 0 and P <= 32
        else:
            assert P <= 64 and Q <= 64
    assert L.picsp_fft_plan_query(1, (C.c_int32 * 5)()) != 0
