"""CPU: libevs.so loads, exports every symbol include/evs.h declares, and refuses to compute without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

import evo_ssearch_b200 as evs
from evo_ssearch_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "evs.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"EVS_API\s+[\w\s\*]+?\b(evs_\w+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    names = _declared_symbols()
    assert len(names) >= 30
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/evs.h but not exported by libevs.so"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    assert _lib.lib().evs_version() == 200


def test_header_constants_match_binding():
    text = open(os.path.join(ROOT, "include", "evs.h")).read()
    consts = dict(re.findall(r"#define\s+(EVS_\w+)\s+\(?(-?\d+)\)?", text))
    for name in ("EVS_OK", "EVS_EINVAL", "EVS_ENODEV", "EVS_ECUDA", "EVS_ENOMEM", "EVS_EIO", "EVS_EFORMAT", "EVS_ELIMIT", "EVS_ETIMEOUT",
                 "EVS_F32", "EVS_F16", "EVS_BF16", "EVS_STORE_F32", "EVS_STORE_BF16_F32", "EVS_MAX_K"):
        assert int(consts[name]) == getattr(_lib, name), name


def test_options_roundtrip_and_validation():
    old = evs.get_option("scan_variant")
    evs.set_option("scan_variant", 2)
    assert evs.get_option("scan_variant") == 2
    evs.set_option("scan_variant", old)
    with pytest.raises(evs.EvsError):
        evs.set_option("scan_variant", 7)
    with pytest.raises(evs.EvsError):
        evs.set_option("no_such_option", 1)
    # the options of the small-shard kernel and of the threaded loader: defaults, round trip, range checks
    for name, default, good, bad in (("small_max_rows", 32768, 0, -1), ("small_fast_cap", 2048, 1, 0), ("small_fast_cap", 2048, 7, 4096),
                                     ("io_threads", 0, 5, 65)):
        assert evs.get_option(name) == default, name
        evs.set_option(name, good)
        assert evs.get_option(name) == good, name
        with pytest.raises(evs.EvsError):
            evs.set_option(name, bad)
        assert evs.get_option(name) == good, name
        evs.set_option(name, default)


def test_host_ptr_is_the_array_address():
    """The search path takes numpy addresses through the buffer protocol (ndarray.ctypes.data_as costs ~5 us per call, three
    per search); read-only and empty arrays fall back to ndarray.ctypes.data."""
    for a in (np.zeros((1, 512), np.float32), np.zeros((3, 48), np.int64), np.zeros((5, 8), np.float32)[2:], np.zeros((0, 12), np.float32)):
        assert _lib.host_ptr(a) == a.ctypes.data
    ro = np.arange(12, dtype=np.float32).reshape(3, 4)
    ro.flags.writeable = False
    assert _lib.host_ptr(ro) == ro.ctypes.data


@pytest.mark.skipif(evs.device_count() > 0, reason="checks the no-GPU behaviour")
def test_no_gpu_means_loud_failure_not_fallback():
    assert evs.device_count() == 0
    with pytest.raises(evs.EvsError) as e:
        evs.IndexFlatIP(512)
    assert e.value.code == _lib.EVS_ENODEV and "no CPU fallback" in str(e.value)
    x = np.ones((2, 8), np.float32)
    with pytest.raises(evs.EvsError) as e:
        evs.normalize_L2(x)
    assert e.value.code == _lib.EVS_ENODEV
    assert (x == 1).all()  # untouched
    with pytest.raises(evs.EvsError):
        evs.read_index(os.path.join(ROOT, "tests", "golden", "index_flat_3x4.faiss"))


def test_read_index_errors_without_touching_the_gpu(tmp_path):
    # format errors are detected before any device work
    h = ctypes.c_void_p()
    L = _lib.lib()
    assert L.evs_index_read(str(tmp_path / "missing.faiss").encode(), 0, 0, ctypes.byref(h)) == _lib.EVS_EIO
    bad = tmp_path / "bad.faiss"
    bad.write_bytes(b"IxHN" + b"\0" * 60)
    assert L.evs_index_read(str(bad).encode(), 0, 0, ctypes.byref(h)) == _lib.EVS_EFORMAT
    blob = open(os.path.join(ROOT, "tests", "golden", "index_flat_3x4.faiss"), "rb").read()
    trunc = tmp_path / "trunc.faiss"
    trunc.write_bytes(blob[:-4])
    assert L.evs_index_read(str(trunc).encode(), 0, 0, ctypes.byref(h)) == _lib.EVS_EFORMAT
    assert b"truncated" in L.evs_last_error()
    l2 = tmp_path / "l2.faiss"
    l2.write_bytes(b"IxF2" + blob[4:33] + (1).to_bytes(4, "little") + blob[37:])
    assert L.evs_index_read(str(l2).encode(), 0, 0, ctypes.byref(h)) == _lib.EVS_EFORMAT
    assert not h.value


def test_null_and_range_arguments():
    L = _lib.lib()
    assert L.evs_index_create(0, 0, 0, ctypes.byref(ctypes.c_void_p())) == _lib.EVS_EINVAL
    assert L.evs_index_create(8, 0, 9, ctypes.byref(ctypes.c_void_p())) == _lib.EVS_EINVAL
    assert L.evs_index_free(None) == 0
    n = ctypes.c_int64()
    assert L.evs_index_ntotal(None, ctypes.byref(n)) == _lib.EVS_EINVAL
    assert L.evs_index_search(None, 1, None, 1, None, None) == _lib.EVS_EINVAL
    assert L.evs_merge_partials_dev(0, 0, 1, 1, None, None, 0, None, None, None) == _lib.EVS_EINVAL
    ex = ctypes.c_void_p()
    assert L.evs_exchange_create(0, 0, 9, 16, 48, ctypes.byref(ex)) == _lib.EVS_ELIMIT  # one box: world <= 8
    assert L.evs_exchange_create(0, 2, 2, 16, 48, ctypes.byref(ex)) == _lib.EVS_EINVAL
    assert L.evs_exchange_create(0, 0, 1, 16, 4096, ctypes.byref(ex)) == _lib.EVS_EINVAL
    assert L.evs_exchange_free(None) == 0
    assert L.evs_index_search_exchange_dev(None, None, 1, None, 1, None, None, None) == _lib.EVS_EINVAL
    assert not ex.value
