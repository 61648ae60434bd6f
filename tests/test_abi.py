"""The C-ABI library builds for sm_100a without a GPU, loads, and exports every symbol include/pssgpu.h declares."""
import ctypes
import importlib
import os
import re
import subprocess

pkg = importlib.import_module("pss-bam_b200")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "pssgpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(pssgpu_[a-z0-9_]+)\s*\(", hdr)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(pkg.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = pkg.build_library()
    h = ctypes.CDLL(lib)
    for sym in _declared_symbols():
        assert hasattr(h, sym), sym
    assert h.pssgpu_abi_version() == 1


def test_library_holds_sm100a_code_with_bulk_copies():
    lib = pkg.build_library()
    out = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass            # cp.async.bulk (TMA engine) stages the SAM tiles
    assert "IDP.4A" in sass            # byte dot products gather the separator masks


def test_no_cpu_fallback_without_a_device():
    """Without a CUDA device the product path must fail loudly, not fall back."""
    h = pkg.load_library()
    if h.pssgpu_device_count() > 0:
        return
    try:
        pkg.Context(0)
    except pkg.PssGpuError as e:
        assert e.code != 0
    else:
        raise AssertionError("Context(0) succeeded without a GPU")


def test_product_does_not_reference_the_oracle():
    """Nothing under pss-bam_b200/ or include/ may import, link or execute oracle/ (or the test harnesses)."""
    bad = []
    for base in ("pss-bam_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            for fn in fns:
                if fn.endswith((".so", ".o", ".pyc")) or "__pycache__" in dp:
                    continue
                txt = open(os.path.join(dp, fn), errors="replace").read()
                for needle in ("liboracle", "oracle_pss", "ora_pss", "ora_fragkon", "ora_kmer", "pss_emul", "libpssemul"):
                    if needle in txt:
                        bad.append((os.path.join(dp, fn), needle))
    assert not bad, bad
