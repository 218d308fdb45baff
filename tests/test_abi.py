"""The C-ABI library: loads without a GPU, exports every symbol the headers declare, fails loudly without a device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from helpers import ROOT, cmoracle, cmref, get_map, have_ref
from threecore_b200 import engine

HEADERS = [os.path.join(ROOT, "include", h) for h in ("cm_gpu.h", "cm_public_batch.h", "sv_world_batch.h")]


def declared_functions():
    names = []
    for h in HEADERS:
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        for m in re.finditer(r"^\s*(?:const\s+)?[A-Za-z_][A-Za-z0-9_ ]*?\**\s*\b((?:cmgpu_|CM_|SV_)[A-Za-z0-9_]+)\s*\(", src, flags=re.M):
            names.append(m.group(1))
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    lib = engine.load_library()
    names = declared_functions()
    assert len(names) >= 40, names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/*.h but not exported by libcmgpu.so: {missing}"


def test_only_c_abi_symbols_are_exported():
    out = subprocess.run(["nm", "-D", "--defined-only", engine.LIB_PATH], capture_output=True, text=True).stdout
    syms = [l.split()[-1] for l in out.splitlines() if " T " in l]
    stray = [s for s in syms if not re.match(r"^(cmgpu_|CM_|SV_)", s)]
    assert not stray, stray


def test_trace_record_layout():
    assert engine.TRACE_DTYPE.itemsize == 56
    offs = {n: engine.TRACE_DTYPE.fields[n][1] for n in engine.TRACE_DTYPE.names}
    # q_shared.h:976-986 measured with offsetof against the reference headers
    assert offs == dict(allsolid=0, startsolid=4, fraction=8, endpos=12, plane_normal=24, plane_dist=36,
                        plane_type=40, plane_signbits=41, plane_pad=42, surfaceFlags=44, contents=48, entityNum=52)


def test_rotation_matrices_match_reference_angle_vectors():
    """host-side AngleVectors/CreateRotationMatrix in the product (cm_gpu.cu) vs the reference's q_math.c"""
    rng = np.random.default_rng(3)
    angles = np.concatenate([rng.uniform(-720, 720, size=(2000, 3)),
                             np.array([[0, 0, 0], [0, 90, 0], [90, 0, 0], [0, 0, 180], [0, 1e-30, 0], [359.99, 0.01, -0.01]])]
                            ).astype(np.float32)
    mats, idx = engine.rotation_matrices(angles)
    assert idx[-6] == -1 and (idx[:-6] >= 0).all()
    chk = cmref.RefCM() if have_ref() else cmoracle.OracleCM(get_map_flat())
    for i in list(range(0, 2000, 7)) + list(range(2001, 2006)):
        f, r, u = chk.angle_vectors(angles[i])
        want = np.concatenate([f, -r, u])
        assert mats[i].tobytes() == want.tobytes(), (i, angles[i], mats[i], want)


def get_map_flat():
    return engine.FlatMap.from_bsp(get_map("tiny").data).to_dict()


def test_no_device_is_a_loud_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = engine.load_library()
    assert lib.cmgpu_device_count() == 0
    with pytest.raises(engine.CMGPUError) as e:
        engine.Engine(0)
    assert e.value.code == 2          # CMGPU_ERR_NO_DEVICE
    assert "no CPU path" in str(e.value)
