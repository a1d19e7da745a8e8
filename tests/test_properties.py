"""Property-based (hypothesis) tests of the pure-logic pieces: container encodings, checksums,
TF-SAME geometry, the shuffle-buffer order and the wave-aware split-K chooser."""
import numpy as np
from hypothesis import given, settings, strategies as st

from tf2_cyclegan_b200.data.pipeline import tf_shuffle_order
from tf2_cyclegan_b200.engine.native_ops import pick_ksplit
from tf2_cyclegan_b200.ops import hostlib
from tf2_cyclegan_b200.ops import oracle as O
from tf2_cyclegan_b200.utils import tensorbundle as tb

FAST = settings(max_examples=60, deadline=None)


@FAST
@given(st.integers(min_value=0, max_value=2**63 - 1))
def test_varint_roundtrip(n):
  enc = tb._varint(n)
  assert tb._read_varint(enc + b"\x01", 0) == (n, len(enc))
  assert len(enc) == max(1, (n.bit_length() + 6) // 7)


@FAST
@given(st.binary(max_size=300), st.binary(max_size=300))
def test_crc32c_is_chainable_and_matches_python(a, b):
  assert hostlib.crc32c(a + b) == hostlib.crc32c(b, hostlib.crc32c(a)) == hostlib._crc32c_py(a + b)
  assert hostlib.unmask_crc32c(hostlib.masked_crc32c(a)) == hostlib.crc32c(a)


@FAST
@given(st.dictionaries(st.binary(min_size=1, max_size=40), st.binary(max_size=80), min_size=1, max_size=60))
def test_sstable_roundtrip(tmp_path_factory, d):
  items = [(b"", b"header")] + sorted(d.items())
  path = str(tmp_path_factory.mktemp("tb") / "t.index")
  tb.write_table(path, items)
  assert tb.read_table(path) == items


@FAST
@given(st.lists(st.integers(min_value=0, max_value=5), min_size=0, max_size=4),
       st.sampled_from(["float32", "int64", "float64", "int32"]))
def test_bundle_entry_roundtrip(shape, dtype):
  e = tb.encode_entry(tb._NP_TO_DT[np.dtype(dtype)], shape, 12345, 678, 0xABCDEF01)
  d = tb.decode_entry(e)
  assert d["shape"] == tuple(shape) and d["offset"] == 12345 and d["size"] == 678 and d["crc32c"] == 0xABCDEF01


@FAST
@given(st.integers(1, 300), st.integers(1, 7), st.integers(1, 3))
def test_tf_same_padding_geometry(size, k, s):
  before, after = O.same_pad_amounts(size, k, s)
  out = -(-size // s)
  total = max((out - 1) * s + k - size, 0)
  assert before + after == total and before == total // 2 and after - before in (0, 1)
  # the padded extent holds exactly `out` windows
  assert (size + total - k) // s + 1 == out or size + total < k


@FAST
@given(st.integers(1, 400), st.integers(1, 300), st.integers(0, 2**31 - 1))
def test_shuffle_order_is_a_permutation_with_bounded_displacement(n, buf, seed):
  order = tf_shuffle_order(n, buf, np.random.default_rng(seed))
  assert sorted(order) == list(range(n))
  # an element cannot be emitted before it entered the buffer: position p only sees indices < p + buf
  assert all(v < p + buf for p, v in enumerate(order))


@FAST
@given(st.integers(1, 64), st.integers(1, 4096))
def test_pick_ksplit_minimises_wave_cost(base, total_kt):
  ks = pick_ksplit(base, total_kt)
  assert 1 <= ks <= max(1, total_kt)

  def cost(k):
    return -(-base * k // 148) * (-(-total_kt // k) + 12)

  hi = max(1, min(total_kt, 4 * 148 // base + 1))
  assert cost(ks) == min(cost(k) for k in range(1, hi + 1))
