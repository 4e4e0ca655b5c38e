"""The C-ABI library: builds for sm_100a, loads without a GPU, exports every symbol that
include/caissa_b200.h declares, and refuses to run without a device (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(REPO, "include", "caissa_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(cb):
    L = cb.load_library()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), "missing export " + n
    assert set(names) == set(cb.SIGNATURES), "ctypes binding and header disagree"


def test_header_is_plain_c():
    src = '#include "caissa_b200.h"\nint main(void){cb_board b; return sizeof(b)==128 && sizeof(cb_timing)==40 ? 0 : 1;}\n'
    exe = os.path.join(REPO, "tests", "_abi_c_check")
    p = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(REPO, "include"), "-x", "c", "-",
                        "-o", exe], input=src.encode(), capture_output=True)
    assert p.returncode == 0, p.stderr.decode()
    try:
        assert subprocess.run([exe]).returncode == 0
    finally:
        os.remove(exe)


def test_weight_count_matches_reference_blob(cb):
    L = cb.load_library()
    from oracle import weights as W
    assert L.cb_weight_count(6, 128) == 1982672 == W.blob_size(6, 128)   # 7,930,688 B (BENCHMARKS.md:203)
    # = the model's 1,979,086 trainable parameters (python/test_architectures.py:262-266) + the running mean / variance of
    #   its BatchNorm layers (14 of 128 channels, v_bn of 1), which the export carries and the parameter count does not
    assert L.cb_weight_count(6, 128) - 2 * (14 * 128 + 1) == 1979086
    assert L.cb_weight_count(10, 256) == W.blob_size(10, 256)
    assert L.cb_version().startswith(b"caissa_b200")


def test_no_cpu_fallback(cb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(cb.CaissaError):
        cb.Evaluator(6, 128, "bf16")
    nn = cb.NeuralNetPolicy()
    assert isinstance(nn.load([0.0]), str)          # Err(String), like load() on a bad file
    assert nn.is_available() is False
    assert nn.predict_batch([bytes(128)], [True], [0.0]) == [None]   # None per item, never raises


def test_sass_contains_blackwell_tensor_and_tma_ops(cb):
    out = subprocess.run(["cuobjdump", "-sass", cb.lib_path()], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM"):
        assert mnemonic in out, mnemonic + " missing: the conv kernel is not the tcgen05/TMA one"
    assert "HMMA." not in out.replace("UTCHMMA", ""), "legacy mma.sync path present"
