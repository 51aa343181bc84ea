"""The C-ABI library: loads, exports every symbol include/tm_lattice.h declares, and fails loudly
(no CPU fallback) when there is no CUDA device.  No compute calls here."""
import ctypes as C
import os
import re

import pytest

import trophic_maelstorm_b200 as tm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "tm_lattice.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = tm.lib()
    names = header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libtm_lattice.so does not export %s" % n
    assert sorted(tm.lattice.ABI_SYMBOLS) == names, "python binding list out of sync with the header"
    assert lib.tm_abi_version() == 1
    assert lib.tm_halo_rows() == 4


def test_struct_layouts_match_header():
    assert C.sizeof(tm.Params) == 8 + 21 * 8
    assert C.sizeof(tm.Stats) == 3 * 8 + 3 * 8 + 8 + 8


def test_no_cpu_fallback():
    lib = tm.lib()
    if lib.tm_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(tm.TmError, match="no CUDA device"):
        tm.Lattice(tm.rietkerk_params(64))


def test_bad_arguments_rejected_before_touching_the_device():
    lib = tm.lib()
    ctx = C.c_void_p()
    p = tm.rietkerk_params(4)        # L < 8
    assert lib.tm_create(C.byref(p), 1, 0, C.byref(ctx)) == -1
    assert b"L must be" in lib.tm_last_error(None)
    p = tm.rietkerk_params(64)
    assert lib.tm_create(C.byref(p), 0, 0, C.byref(ctx)) == -1
    p.model = 9
    assert lib.tm_create(C.byref(p), 1, 0, C.byref(ctx)) == -1
    assert lib.tm_step(None, 1) == -1
    assert lib.tm_destroy(None) == 0


def test_product_never_references_the_oracle():
    """The product tree must not import, link or mention the oracle (it is the checker only)."""
    pkg = os.path.join(ROOT, "trophic_maelstorm_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(d, f), errors="ignore").read()
                assert "tm_oracle" not in txt and "oracle_lib" not in txt and "libtm_oracle" not in txt, os.path.join(d, f)


def _header_names(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tm_[a-z0-9_]+)\s*\(", src)))


def test_fke_and_consumer_headers_are_exported_too():
    """include/tm_fke.h and include/tm_consumer.h: every declared entry point is exported and bound by the ctypes views."""
    lib = tm.lib()
    for header, bound in (("tm_fke.h", tm.FKE_ABI_SYMBOLS), ("tm_consumer.h", tm.CONS_ABI_SYMBOLS), ("tm_consumer3.h", tm.CONS3_ABI_SYMBOLS)):
        names = _header_names(header)
        assert names, header
        for n in names:
            assert hasattr(lib, n), "libtm_lattice.so does not export %s (%s)" % (n, header)
        assert sorted(bound) == names, "python binding list out of sync with %s" % header


def test_other_entry_points_have_no_cpu_fallback_either():
    """The consumer-side stepper, the FKE mover and the noise-off DP path fail loudly without a device."""
    lib = tm.lib()
    if lib.tm_device_count() > 0:
        pytest.skip("a CUDA device is present")
    import numpy as np
    with pytest.raises(tm.TmError, match="no CUDA device"):
        tm.ConsumerOps(32)
    with pytest.raises(tm.TmError, match="no CUDA device"):
        tm.FkeMover(32, 1, device=0)
    with pytest.raises(tm.TmError, match="no CUDA device"):
        tm.dp_noise_off(8, 1.0, 1.0, 0.1, 0.1, 1.0, np.ones(64), 1)


def test_slab_library_exports_every_declared_symbol():
    """include/tm_slab.h -> libtm_slab.so (links libtm_lattice.so and libnccl.so.2); creating a ring without a GPU fails loudly."""
    from trophic_maelstorm_b200 import slab as S
    lib = S.lib()
    names = [n for n in _header_names("tm_slab.h")]
    assert names == sorted(S.SLAB_ABI_SYMBOLS), (names, sorted(S.SLAB_ABI_SYMBOLS))
    for n in names:
        assert hasattr(lib, n), "libtm_slab.so does not export %s" % n
    assert S.partition(16384, 8, 7) == (14336, 2048) and S.partition(10, 3, 0) == (0, 4) and S.partition(10, 3, 2) == (7, 3)
    if tm.lib().tm_device_count() == 0:
        with pytest.raises(tm.TmError, match="no CUDA device"):
            S.LocalSlabRing(tm.rietkerk_params(64), 1, [0, 0], transport=S.PEER_COPY)
