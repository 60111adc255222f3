"""CPU tests of the C-ABI boundary: the library loads without a GPU, exports every symbol
include/asr.h declares, and rejects bad calls with the documented codes before touching CUDA."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from deeplabv3plus_augmented_superresolution_b200 import _build, _lib


@pytest.fixture(scope="module")
def L():
    _build.build()          # cross-compiles for sm_100a without a GPU
    return _lib.lib()


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "asr.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(asr_[a-z_0-9]+)\s*\(", hdr)))


def test_exports_every_declared_symbol(L):
    names = declared_symbols()
    assert len(names) == len(_lib.SIGNATURES) >= 17
    for n in names:
        assert hasattr(L, n), f"libasr.so does not export {n}"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert L.asr_version() == 200


def test_struct_layout_matches_header():
    # 4 floats, 2 ints, 4 floats, int, 2 floats, 2 ints, 2 floats, (pad) int64, 2 ints
    assert C.sizeof(_lib.AsrSolveParams) == 88
    assert _lib.AsrSolveParams.step_offset.offset == 72


def test_workspace_query_and_error_codes(L):
    need = C.c_size_t()
    assert L.asr_solve_workspace_bytes(2, 100, 128, 128, 512, 512, 300, C.byref(need)) == 0
    lo = 2 * (5 * 512 * 512 * 4 + 100 * 128 * 128 * 4)
    assert lo <= need.value <= lo + (1 << 22)
    assert L.asr_solve_workspace_bytes(2, 100, 128, 128, 512, 512, 300, None) == -2
    assert L.asr_solve_workspace_bytes(1, 4, 64, 64, 512, 512, 10, C.byref(need)) == 0      # x8: Superresolution's default feature_size
    assert L.asr_solve_workspace_bytes(1, 4, 64, 64, 192, 192, 10, C.byref(need)) == -3     # odd ratio
    assert b"even integer" in L.asr_last_error()
    assert L.asr_solve_workspace_bytes(1, 4, 64, 64, 512, 256, 10, C.byref(need)) == -3     # anisotropic
    assert L.asr_solve_workspace_bytes(0, 4, 16, 16, 64, 64, 10, C.byref(need)) == -1

    p = _lib.SolveParams(num_iter=3).to_c()
    ang = np.zeros((1, 4), np.float32); shf = np.zeros((1, 4, 2), np.float32)
    fake = C.c_void_p(0x1000)           # never dereferenced: every check below fails before any CUDA call
    args = lambda prm, n, B=1, h=16, H=64, ws=1 << 30: (C.byref(prm), n, fake, ang.ctypes.data_as(_lib._fp), shf.ctypes.data_as(_lib._fp),
                                                        None, B, 4, h, h, H, H, fake, None, fake, ws, None)
    assert L.asr_solve_batched(*args(p, 1, ws=16)) == -5                                     # workspace too small
    assert L.asr_solve_batched(*args(p, 2)) == -1                                           # n_params must be 1 or B
    assert L.asr_solve_batched(*args(p, 1, H=48)) == -3
    pb = _lib.SolveParams(optimizer="adam").to_c(); pb.optimizer = 9
    assert L.asr_solve_batched(*args(pb, 1)) == -1 and b"optimizer" in L.asr_last_error()
    assert L.asr_solve_batched(None, 1, fake, ang.ctypes.data_as(_lib._fp), shf.ctypes.data_as(_lib._fp), None,
                               1, 4, 16, 16, 64, 64, fake, None, fake, 1 << 30, None) == -2
    # the traced form validates its extra arguments before anything else touches the device
    tr = lambda every, buf, cols: L.asr_solve_batched_traced(C.byref(p), 1, fake, ang.ctypes.data_as(_lib._fp), shf.ctypes.data_as(_lib._fp), None,
                                                             1, 4, 16, 16, 64, 64, fake, None, every, buf, cols, fake, 1 << 30, None)
    assert tr(10, None, 3) == -1 and b"loss trace" in L.asr_last_error()
    assert tr(-1, fake, 3) == -1 and tr(10, fake, 0) == -1
    # workspace queries of the aux calls
    need2 = C.c_size_t()
    assert L.asr_warp_affine_workspace_bytes(100, 512, 512, 3, C.byref(need2)) == 0 and need2.value >= 512 * 512 * 16 + 100 * 32
    assert L.asr_warp_affine_workspace_bytes(100, 512, 512, 5, C.byref(need2)) == -1
    assert L.asr_backproject_workspace_bytes(8, 100, C.byref(need2)) == 0 and need2.value >= 8 * 100 * 32
    assert L.asr_backproject_workspace_bytes(0, 100, C.byref(need2)) == -1 and L.asr_backproject_workspace_bytes(8, 100, None) == -2
    assert L.asr_l2_read_probe(None, 1 << 20, 2, fake, None) == -2 and L.asr_l2_read_probe(fake, 8, 2, fake, None) == -1
    with pytest.raises(NotImplementedError):
        _lib.check(-3)
    with pytest.raises(_lib.AsrError):
        _lib.check(-1)


def test_aux_entry_points_validate_arguments(L):
    fake = C.c_void_p(0x1000)
    a = np.zeros(4, np.float32)
    fp = a.ctypes.data_as(_lib._fp)
    assert L.asr_warp_affine(fake, fp, fp, 2, 8, 8, 9, 1, fake, None) == -1            # C > 4
    assert L.asr_warp_affine(fake, fp, fp, 2, 8, 8, 3, 7, fake, None) == -1            # unknown interpolation
    assert L.asr_opm_extract(fake, 2, 8, 8, 21, 21, 0, fake, None, None, None) == -1   # class_id out of range
    assert L.asr_opm_extract(fake, 2, 8, 8, 21, 8, 2, fake, None, None, None) == -2    # slice_max needs max_out
    assert L.asr_backproject_batched(5, fake, fp, fp, 1, 2, 8, 8, 32, 32, fake, None) == -1
    assert L.asr_threshold(fake, 1, 64, 8, 0.2, None, fake, None, None) == -2          # th_factor path needs workspace
    assert L.asr_minmax_normalize(fake, 0, 0.0, 1.0, fake, fake, None) == -1


def test_plain_c_consumer(L, tmp_path):
    """include/asr.h compiles as C99 and a C program links and calls the library (no Python, no torch, no GPU)."""
    import subprocess
    exe = str(tmp_path / "abi_consumer")
    libdir = os.path.dirname(_lib.LIB_PATH)
    env = dict(os.environ); env.pop("CC", None)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c", "abi_consumer.c"), "-o", exe, "-L", libdir, "-lasr",
                           "-Wl,-rpath," + libdir], env=env)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "c abi ok" in out.stdout


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libasr.so"))
    with pytest.raises(_lib.AsrError, match="no CPU fallback"):
        _lib.lib()


def test_no_cuda_means_error_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(_lib.AsrError, match="no CPU path"):
        _lib.solve_batched(torch.zeros(1, 2, 8, 8), np.zeros((1, 2)), np.zeros((1, 2, 2)), _lib.SolveParams())
