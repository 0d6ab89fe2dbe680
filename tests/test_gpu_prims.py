"""Device-wide primitives (comprox_b200/csrc/cr_sort.cuh): the hand-written stable LSD radix sort and exclusive scan
against numpy, at ragged sizes, partial bit ranges and both key widths."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def handle():
    from comprox_b200 import api
    return api.Handle(api.ROLZ)


@pytest.mark.parametrize("n", [1, 31, 4096, 4097, 100003, 5_000_001])
@pytest.mark.parametrize("bits", [(0, 3), (0, 8), (0, 11), (0, 21), (0, 32), (5, 22)])
def test_sort_u32(handle, n, bits):
    rng = np.random.default_rng(n * 131 + bits[1])
    # skewed keys: most positions share a few buckets, like text contexts
    keys = (rng.zipf(1.3, n).astype(np.uint64) * 2654435761 % (1 << 32)).astype(np.uint32)
    vals = np.arange(n, dtype=np.uint32)
    ko, vo = handle.debug_sort(keys, vals, *bits)
    field = (keys >> np.uint32(bits[0])) & np.uint32((1 << (bits[1] - bits[0])) - 1) if bits[1] - bits[0] < 32 else keys
    order = np.argsort(field, kind="stable")
    assert np.array_equal(vo, vals[order])
    assert np.array_equal(ko, keys[order])


@pytest.mark.parametrize("n", [7, 65536, 1_234_567])
@pytest.mark.parametrize("end_bit", [32, 40])
def test_sort_u64(handle, n, end_bit):
    rng = np.random.default_rng(n + end_bit)
    keys = rng.integers(0, 1 << 40, n, dtype=np.uint64) | (rng.integers(0, 1 << 20, n, dtype=np.uint64) << np.uint64(44))
    vals = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    ko, vo = handle.debug_sort(keys, vals, 0, end_bit)
    order = np.argsort(keys & np.uint64((1 << end_bit) - 1), kind="stable")
    assert np.array_equal(vo, vals[order])
    assert np.array_equal(ko, keys[order])


@pytest.mark.parametrize("n", [1, 1023, 1024, 1025, 1 << 20, (1 << 20) + 1, 40_000_003])
def test_scan(handle, n):
    rng = np.random.default_rng(n)
    v = rng.integers(0, 1 << 12, n, dtype=np.uint32)
    got = handle.debug_scan(v)
    want = np.concatenate([[0], np.cumsum(v[:-1], dtype=np.uint64)]).astype(np.uint64) & np.uint64(0xFFFFFFFF)
    assert np.array_equal(got, want.astype(np.uint32))
