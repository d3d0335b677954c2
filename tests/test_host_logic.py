"""Host-side logic that needs no GPU: operand collapsing, layout detection,
sharding, and that the product path FAILS LOUDLY without CUDA (no CPU or
oracle fallback)."""
import sys

import pytest
import torch

import nitorch_fastmath_b200 as nfm
from nitorch_fastmath_b200 import _dispatch as D
from nitorch_fastmath_b200 import _lib
from nitorch_fastmath_b200.shard import shard_bounds, shard_sizes


def test_collapse_dense_broadcast_and_strided():
    t = torch.zeros(4, 5, 6)
    op = D.as_operand(t, (4, 5), 1, torch.float32)
    assert op.stride == 6 and op.ptr == t.data_ptr()
    # fully broadcast operand: stride 0, no copy
    v = torch.zeros(3)
    op = D.as_operand(v, (4, 5), 1, torch.float32)
    assert op.stride == 0 and op.ptr == v.data_ptr()
    # leading singleton batch dims broadcast as well
    op = D.as_operand(torch.zeros(1, 1, 3), (4, 5), 1, torch.float32)
    assert op.stride == 0
    # every second matrix: a single non-dense stride, still no copy
    s = torch.zeros(10, 6)[::2]
    op = D.as_operand(s, (5,), 1, torch.float32)
    assert op.stride == 12 and op.ptr == s.data_ptr()
    # partially broadcast / transposed batch dims do not collapse -> materialised
    p = torch.zeros(1, 5, 3)
    op = D.as_operand(p, (4, 5), 1, torch.float32)
    assert op.stride == 3 and op.tensor.is_contiguous() and op.tensor.shape == (4, 5, 3)
    tr = torch.zeros(5, 4, 6).transpose(0, 1)
    op = D.as_operand(tr, (4, 5), 1, torch.float32)
    assert op.stride == 6 and op.tensor.is_contiguous()
    # coefficient dim not unit-stride -> made contiguous
    cf = torch.zeros(6, 7).t()
    op = D.as_operand(cf, (7,), 1, torch.float32)
    assert op.stride == 6 and op.tensor.is_contiguous()
    # dense 2-D records
    a = torch.zeros(3, 4, 4)
    assert D.as_operand(a, (3,), 2, torch.float32).stride == 16
    assert D.as_operand(a.transpose(-1, -2), (3,), 2, torch.float32).tensor.is_contiguous()
    # dtype conversion
    assert D.as_operand(torch.zeros(2, 3), (2,), 1, torch.float64).tensor.dtype == torch.float64
    # single-element batch
    assert D.as_operand(torch.zeros(6), (), 1, torch.float32).stride == 6


def test_out_operand():
    o, res, back = D.out_operand(None, (4, 3), 1, torch.float32, torch.device("cpu"))
    assert res.shape == (4, 3) and o.stride == 3 and not back
    given = torch.zeros(4, 3)
    o, res, back = D.out_operand(given, (4, 3), 1, torch.float32, torch.device("cpu"))
    assert res is given and o.ptr == given.data_ptr() and not back
    tr = torch.zeros(3, 4).t()                      # coefficient-first storage: staged + copied back
    o, res, back = D.out_operand(tr, (4, 3), 1, torch.float32, torch.device("cpu"))
    assert res is tr and back and o.ptr != tr.data_ptr()
    with pytest.raises(RuntimeError):
        D.out_operand(torch.zeros(4, 2), (4, 3), 1, torch.float32, torch.device("cpu"))


def test_layout_and_order():
    assert D.packed_order(6) == 3 and D.packed_order(55) == 10 and D.packed_order(1) == 1
    with pytest.raises(ValueError):
        D.packed_order(7)
    assert D.detect_layout(6, 3) == _lib.LAYOUT_SYM
    assert D.detect_layout(1, 3) == _lib.LAYOUT_SCALED_IDENTITY
    assert D.detect_layout(4, 4) == _lib.LAYOUT_DIAG
    assert D.detect_layout(16, 4) == _lib.LAYOUT_FULL
    assert D.detect_layout(1, 1) == _lib.LAYOUT_SYM
    with pytest.raises(ValueError):
        D.detect_layout(5, 3)


def test_compute_dtype():
    f32, f64 = torch.zeros(1), torch.zeros(1, dtype=torch.float64)
    assert D.compute_dtype(f32, f64) == torch.float64
    assert D.compute_dtype(f32, f32) == torch.float32
    assert D.compute_dtype(torch.zeros(1, dtype=torch.float16)) == torch.float32
    with pytest.raises(TypeError):
        D.compute_dtype(torch.zeros(1, dtype=torch.int64))


@pytest.mark.parametrize("batch,world", [(256 ** 3, 8), (192 ** 3, 4), (160 ** 3, 8), (1_000_003, 3), (5, 8), (0, 2)])
def test_shard_bounds_partition(batch, world):
    bounds = [shard_bounds(batch, world, r) for r in range(world)]
    assert bounds[0][0] == 0 and bounds[-1][1] == batch
    for (b0, e0), (b1, e1) in zip(bounds[:-1], bounds[1:]):
        assert e0 == b1 and b0 <= e0
    sizes = shard_sizes(batch, world)
    assert sum(sizes) == batch and max(sizes) - min(sizes) < 2 * 1024
    for b, e in bounds[:-1]:
        assert b % 1024 == 0 or b == batch   # every non-empty slab keeps the TMA alignment
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_product_fails_loudly_without_cuda():
    """No CPU fallback: CPU tensors are streamed through a CUDA device, and when
    there is none every entry point raises instead of computing elsewhere."""
    m, v = torch.rand(8, 6) + 3, torch.rand(8, 3)
    a = torch.rand(8, 3, 3) + 3 * torch.eye(3)
    calls = [
        lambda: nfm.sym_solve(m, v), lambda: nfm.sym_matvec(m, v), lambda: nfm.sym_addmatvec(v, m, v),
        lambda: nfm.sym_invert(m), lambda: nfm.sym_det(m), lambda: nfm.sym_to_full(m), lambda: nfm.sym_outer(v),
        lambda: nfm.batchinv(a), lambda: nfm.batchdet(a), lambda: nfm.batchmatvec(a, v),
        lambda: nfm.solvevec(a, v), lambda: nfm.lmdiv(a, a), lambda: nfm.inv(a),
    ]
    for call in calls:
        with pytest.raises(RuntimeError):
            call()


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: importing the product must not load it."""
    import subprocess
    code = ("import sys; import nitorch_fastmath_b200; "
            "bad=[m for m in sys.modules if m == 'oracle' or m.startswith('oracle.')]; "
            "sys.exit(1 if bad else 0)")
    from conftest import ROOT
    assert subprocess.run([sys.executable, "-c", code], cwd=ROOT).returncode == 0
    import glob, os
    for path in glob.glob(os.path.join(ROOT, "nitorch_fastmath_b200", "**", "*.py"), recursive=True):
        text = open(path).read()
        assert "import oracle" not in text and "from oracle" not in text, path


def test_missing_library_is_an_error(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.NfmError):
        _lib.load()


def test_outer_split_finds_the_leading_dims_to_loop_over():
    """Partially broadcast operands are not materialised: the Python layer loops over the
    leading batch dims and launches on views (_dispatch.outer_split / outer_views)."""
    import torch
    from nitorch_fastmath_b200 import _dispatch as D
    vec = torch.zeros(3, 4, 5, 3)
    assert D.outer_split((3, 4, 5), [(torch.zeros(3, 4, 5, 6), 1), (vec, 1)]) == 0        # dense: nothing to split
    assert D.outer_split((3, 4, 5), [(torch.zeros(6), 1), (vec, 1), (None, 1)]) == 0        # fully broadcast
    assert D.outer_split((3, 4, 5), [(torch.zeros(1, 4, 5, 6), 1), (vec, 1)]) == 1          # one field for the batch
    assert D.outer_split((3, 4, 5), [(torch.zeros(3, 1, 1, 6), 1), (vec, 1)]) == 1          # one matrix per image
    assert D.outer_split((3, 4, 5), [(torch.zeros(3, 1, 5, 6), 1), (vec, 1)]) == 2          # broadcast in the middle
    assert D.outer_split((3, 4, 5), [(torch.zeros(1, 4, 5, 6), 1), (vec, 1)], max_outer=2) == 0   # too many launches
    v = D.outer_views(torch.arange(120.).view(1, 4, 5, 6), (3, 4, 5), 1, (2,))
    assert v.shape == (4, 5, 6) and v.stride() == (30, 6, 1)          # a view of the one shared field, no copy
    assert D.outer_views(None, (3, 4, 5), 1, (0,)) is None


def test_empty_like_phased_matches_the_misalignment_of_its_reference():
    import torch
    from nitorch_fastmath_b200 import _dispatch as D
    base = torch.zeros(1000 * 3 + 3)
    for start in (0, 1, 2, 3):
        ref = base[start:start + 999].view(333, 3)
        out = D.empty_like_phased(ref)
        assert out.shape == ref.shape and out.dtype == ref.dtype and out.is_contiguous()
        assert out.data_ptr() % 16 == ref.data_ptr() % 16
    out = D.empty_like_phased(base[1:7], shape=(2, 5))
    assert out.shape == (2, 5) and out.data_ptr() % 16 == base[1:7].data_ptr() % 16
