"""not gpu: the C-ABI shared library loads and exports every symbol include/mgb.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from materngabo_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "mgb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mgb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_typed():
    names = _declared()
    assert len(names) >= 19
    lib = ctypes.CDLL(build.build())
    for n in names:
        assert hasattr(lib, n), "libmgb.so does not export %s" % n
        assert n in _lib.PROTOTYPES, "no ctypes prototype for %s" % n
    assert sorted(_lib.PROTOTYPES) == names


def test_loader_and_version():
    lib = _lib.load()
    assert lib.mgb_version() == _lib.MGB_VERSION
    assert lib.mgb_launch_count() >= 0
    assert lib.mgb_posterior_pack_bytes(5000) == 5120 * 5008 * 8
    assert lib.mgb_posterior_workspace_bytes(128, 100) == (128 * 112 + 128) * 8 + lib.mgb_posterior_reduce_workspace_bytes(128, 100)
    assert lib.mgb_posterior_reduce_workspace_bytes(0, 100) == 0


def test_series_desc_layout_matches_header():
    d = _lib.SeriesDesc(1, 3, 10, 0, 1.0, 0.0, -1.0, 1.0)
    assert ctypes.sizeof(d) == 4 * 4 + 4 * 8 + 8             # 4 ints, 4 doubles, pair_square + tail padding
    assert (d.n_factors, d.factor_width, d.n_terms, d.basis, d.pair_square) == (1, 3, 10, 0, 0)
    assert _lib.SeriesDesc.pair_square.offset == 48


def test_product_has_no_oracle_dependency():
    """The product package must never import oracle/ (SURVEY scope rule: the oracle is the checker only)."""
    pkg = os.path.join(ROOT, "materngabo_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def test_int8_path_size_queries_need_no_gpu():
    """Host-side size arithmetic of the opt-in INT8 posterior path (csrc/int8_posterior.cu): five column blocks from
    2048 training points on, operands padded to 256 * blocks, compacted copies of (J + 1) / 5 of the columns."""
    from materngabo_b200 import _lib

    lib = _lib.load()
    assert lib.mgb_int8_available() in (0, 1)
    assert lib.mgb_posterior_int8_pack_bytes(0) == 0
    # n = 5000 -> padded 5120, five blocks of 1024 rows; block J holds 1024 rows x 7 slices x (J + 1) * 1024 columns
    assert lib.mgb_posterior_int8_pack_bytes(5000) == 4 * 5120 + 7 * 1024 * 1024 * (1 + 2 + 3 + 4 + 5)
    # n = 1000 -> one block, padded 1024
    assert lib.mgb_posterior_int8_pack_bytes(1000) == 4 * 1024 + 7 * 1024 * 1024
    m = 32768
    ws = lib.mgb_posterior_int8_workspace_bytes(m, 5000)
    assert ws == 4 * m + m * 7 * 1024 * 15 + 7 * m * 5120 * 4 + (1 << 20)
    assert lib.mgb_posterior_int8_workspace_bytes(0, 5000) == 0


def test_int8_gram_path_size_queries_need_no_gpu():
    """Host-side arithmetic of the opt-in INT8 tensor-core series Gram kernel (csrc/series_i8.cu): one stacked operand
    tile (7 planes x 32 columns x 128 bytes) per 32 training points plus one exponent word per (padded) column."""
    import ctypes as C

    from materngabo_b200 import _lib, _ops

    lib = _lib.load()
    assert lib.mgb_series_gram_i8_workspace_bytes(0) == 0
    assert lib.mgb_series_gram_i8_workspace_bytes(2000) == 63 * 7 * 32 * 128 + 2016 * 4 + 128      # exponent words padded to 256 bytes
    assert lib.mgb_series_gram_i8_workspace_bytes(32) == 7 * 32 * 128 + 256
    sphere = _ops.SeriesSpec(1, 11, 0, 1.0, 0.0, -1.0, 1.0).desc(10)
    torus = _ops.SeriesSpec(10, 2, 1, 1.0, 0.0, -1.0, 1.0).desc(40)
    wide = _ops.SeriesSpec(1, 17, 0, 1.0, 0.0, -1.0, 1.0).desc(10)
    long_series = _ops.SeriesSpec(1, 3, 0, 1.0, 0.0, -1.0, 1.0).desc(13)
    assert lib.mgb_series_gram_i8_supported(C.byref(sphere)) == 1
    assert [lib.mgb_series_gram_i8_supported(C.byref(d)) for d in (torus, wide, long_series)] == [0, 0, 0]


def test_gram_path_knob_is_validated(monkeypatch):
    from materngabo_b200 import _ops

    monkeypatch.setenv("MGB_GRAM_PATH", "int8")
    assert _ops._gram_path() == "int8"
    monkeypatch.setenv("MGB_GRAM_PATH", "fast")
    with pytest.raises(ValueError):
        _ops._gram_path()
    monkeypatch.delenv("MGB_GRAM_PATH")
    assert _ops._gram_path() == "fp64"
