"""The C-ABI shared library loads without a GPU and exports every entry point that include/dspeed_b200.h
declares (the drop-in boundary: what a maintainer's ctypes binding would resolve).  No compute calls here."""
import ctypes
import os
import re
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "include", "dspeed_b200.h")


def declared_symbols():
    # run the header through the C preprocessor so that the DSPB_DECLARE*(_f32 / _f64) macro lists expand
    src = subprocess.run(["gcc", "-E", "-P", HEADER], capture_output=True, text=True, check=True).stdout
    names = set(re.findall(r"\b(dspb_\w+)\s*\(", src))
    return sorted(n for n in names if not n.startswith("dspb_chain") or n in (
        "dspb_chain_create", "dspb_chain_launch", "dspb_chain_smem_bytes", "dspb_chain_profile", "dspb_chain_destroy"))


def test_header_declares_both_type_loops():
    names = declared_symbols()
    assert len(names) > 80
    f32 = {n[:-4] for n in names if n.endswith("_f32")}
    f64 = {n[:-4] for n in names if n.endswith("_f64")}
    assert f32 - {"dspb_convolve_tc"} == f64 and "dspb_histogram" in f32 and "dspb_pole_zero" in f32


def test_library_exports_every_declared_symbol():
    from dspeed_b200 import _lib

    lib = ctypes.CDLL(_lib.LIB_PATH)       # loads on a machine without a GPU
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.dspb_version() >= 100


def test_generated_chain_kernels_export_the_documented_entry_points():
    """the specialised chain kernels (dspeed_b200/_chains/*.so) export chain_launch / chain_smem_bytes / chain_n_nodes
    with the signature documented in include/dspeed_b200_chain.h"""
    import glob

    hdr = open(os.path.join(REPO, "include", "dspeed_b200_chain.h")).read()
    for name in ("chain_launch", "chain_smem_bytes", "chain_n_nodes"):
        assert re.search(rf"\b{name}\s*\(", hdr)
    libs = glob.glob(os.path.join(REPO, "dspeed_b200", "_chains", "*.so"))
    assert libs, "build() prebuilds the shipped configurations"
    lib = ctypes.CDLL(sorted(libs, key=os.path.getmtime)[-1])
    for name in ("chain_launch", "chain_smem_bytes", "chain_n_nodes"):
        assert hasattr(lib, name)
    assert 0 < lib.chain_smem_bytes() <= 227 * 1024
