"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports exactly what
include/hdpgpc_b200.h declares; the product refuses to run without a GPU; the product never imports
the oracle."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "hdpgpc_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(hgp_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_and_exports_every_declared_symbol():
    import hdpgpc_b200
    from hdpgpc_b200 import _lib
    path = hdpgpc_b200.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    syms = header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/hdpgpc_b200.h but not exported"
    assert sorted(_lib.PROTOTYPES) == syms, "ctypes prototypes out of sync with the header"
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r"\bT (hgp_[a-z0-9_]+)", out)))
    assert exported == syms, "library exports symbols the header does not declare (or vice versa)"


def test_chain_desc_layout_matches_the_header(tmp_path):
    """The ctypes mirror of hgp_chain_desc (copied byte for byte to the device by ops.chain_run) against the header as a C
    compiler lays it out: size and the offset of every field -- a header edit that moves a pointer must fail here, not
    make the chain kernel follow a corrupted pointer."""
    from hdpgpc_b200 import _lib
    fields = [n for n, _ in _lib.ChainDesc._fields_]
    src = tmp_path / "offsets.c"
    prints = "\n".join(f'    printf("{n} %zu\\n", offsetof(hgp_chain_desc, {n}));' for n in fields)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "hdpgpc_b200.h"\nint main(void) {\n'
                   + prints + '\n    printf("sizeof %zu\\n", sizeof(hgp_chain_desc));\n    return 0;\n}\n')
    exe = tmp_path / "offsets"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    assert int(out["sizeof"]) == ctypes.sizeof(_lib.ChainDesc)
    for n in fields:
        assert int(out[n]) == getattr(_lib.ChainDesc, n).offset, n
    assert len(out) == len(fields) + 1        # no field of the header is missing from the mirror (sizes agree as well)


def test_library_is_sm100a_with_tensor_and_tma_sass():
    import hdpgpc_b200
    path = hdpgpc_b200.build()
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "DMMA" in sass            # FP64 tensor-core instruction of the tile kernel
    assert "UBLKCP" in sass          # bulk async copy (TMA) of the factor stream
    assert "USETMAXREG" in sass      # warpgroup register re-balancing


def test_no_cpu_fallback():
    import torch
    import hdpgpc_b200
    from hdpgpc_b200 import ops
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(hdpgpc_b200.HgpError):
        ops.chol_batched(torch.eye(4, dtype=torch.float64).reshape(1, 4, 4))
    with pytest.raises(hdpgpc_b200.HgpError):
        ops.pack_leads(torch.zeros(2, 4, 2, dtype=torch.float64))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "hdpgpc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
                assert "/root/reference" not in txt, f
    code = "import sys; import hdpgpc_b200, hdpgpc_b200.synthetic; assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)


def test_host_operands_match_oracle(golden):
    """The K-sized HMM operands and the state index maps are host logic of the product; they must
    agree with the oracle's restatement bit for bit."""
    import numpy as np
    from hdpgpc_b200 import hdp, model
    from oracle import hdpgpc_oracle as O
    z = golden("hmm_synth")
    for c in range(int(z["n_cases"])):
        tt, st = z[f"c{c}_transTheta"], z[f"c{c}_startTheta"]
        K = z[f"c{c}_q"].shape[1]
        sp, tp = hdp.expected_log_pi(tt, st, K)
        assert np.allclose(sp, z[f"c{c}_startPi"], rtol=1e-13, atol=0)
        assert np.allclose(tp, z[f"c{c}_transPi"], rtol=1e-13, atol=0)
        assert np.allclose(hdp.compute_trans_A(tt, K), z[f"c{c}_trans_A"], rtol=1e-13, atol=0)
        osp, otp = O.start_trans_pi(tt, st, K)
        assert np.array_equal(sp, osp) and np.array_equal(tp, otp)
        mine = hdp.hmm_operands(tt, sp, K)
        ref = O.hmm_operands(tt, sp, K)
        for a, b in zip(mine, ref):
            assert np.array_equal(a, b)
    idx = [3, 4, 9, 17]
    gp = type("G", (), {"indexes": idx})()
    i1, f1 = model.state_index_map(idx, 25)
    og = O.OracleGP.__new__(O.OracleGP)
    og.indexes = idx
    i2, f2 = og.state_index_map(25)
    assert np.array_equal(i1, i2) and np.array_equal(f1, f2)
    assert np.array_equal(model.snr_state_index(idx, 5, 25), O.snr_state_index(idx, 5, 25))


def test_slice_bounds_cover_the_beats_tile_aligned():
    """EStepEngine.slice_bounds (host logic of sweep_from_host): the slices partition [0, N), every cut is a multiple of the
    tile, sizes never shrink along a geometric schedule, degenerate inputs stay sane."""
    from hdpgpc_b200.hdp import EStepEngine
    for N in (0, 1, 63, 64, 65, 1000, 100000, 2000000 // 8):
        for n_slices, growth in ((1, 1.0), (8, 1.0), (4, 4.0), (3, 2.0), (40, 3.0), (10 ** 6, 1.0)):
            b = EStepEngine.slice_bounds(N, n_slices, growth)
            if N == 0:
                assert b == []
                continue
            assert b[0][0] == 0 and b[-1][1] == N and len(b) <= max(1, n_slices)
            assert all(x[1] == y[0] for x, y in zip(b[:-1], b[1:]))
            assert all(lo % 64 == 0 and hi > lo for lo, hi in b)
            if growth > 1.0 and len(b) > 2:
                sizes = [hi - lo for lo, hi in b[:-1]]
                assert all(s2 >= s1 for s1, s2 in zip(sizes[:-1], sizes[1:]))
    b = EStepEngine.slice_bounds(100000, 4, 4.0)
    assert b[0][1] - b[0][0] < 100000 // 50       # the exposed first copy is a small fraction of the batch


def test_q_lat_stale_marks():
    """Which members compute_q_lat_all re-scores after each online seam call (index bookkeeping only, no device):
    member j reads smoothed states j, j+1 and parameter set min(j+1, last) (GPI_model.py:288-306)."""
    from hdpgpc_b200.model import GPI_model
    gp = GPI_model.__new__(GPI_model)
    gp._qlat_stable = 12                          # a chain of 12 members, all cached
    gp._qlat_dirty(10)                            # backwards_pair with N = 12: states 11, 12 -> members 10, 11
    assert gp._qlat_stable == 10
    gp._qlat_dirty(11)                            # a later, weaker mark never raises the level
    assert gp._qlat_stable == 10
    gp._qlat_dirty(-3)
    assert gp._qlat_stable == 0
    gp.invalidate_caches()
    assert gp._qlat_stable == 0 and gp._tables is None
