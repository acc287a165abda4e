"""Pins the CPU oracle (and host helpers) to every reference expectation that does not need MuJoCo at run time
(tests/golden/reference_vectors.json: math_test.py closest-point cases, triangular index maps, io_test padding sizes)."""

import ctypes
import json
import os

import numpy as np
import pytest

from tests import util

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")))


@pytest.fixture(scope="module")
def lib(built):
  from oracle import orc

  return orc._lib(8)


def _v(x):
  return (ctypes.c_double * 3)(*[float(t) for t in x])


@pytest.mark.parametrize("case", G["closest_segment_to_segment_points"], ids=lambda c: c["ref"])
def test_closest_segment_to_segment_points(lib, case):
  oa, ob = (ctypes.c_double * 3)(), (ctypes.c_double * 3)()
  lib.orc_closest_segment_to_segment_points(_v(case["a0"]), _v(case["a1"]), _v(case["b0"]), _v(case["b1"]), oa, ob)
  tol = 0.5 * 10 ** (-case["places"])  # assertSequenceAlmostEqual(places)
  np.testing.assert_allclose(list(oa), case["best_a"], atol=tol, rtol=0)
  np.testing.assert_allclose(list(ob), case["best_b"], atol=tol, rtol=0)
  assert np.isfinite(list(oa) + list(ob)).all()


def test_upper_tri_index(lib):
  for n in G["upper_tri_index"]["sizes"]:
    arr = [lib.orc_upper_tri_index(n, i, j) for i in range(n) for j in range(i + 1, n)]
    assert arr == list(range(n * (n - 1) // 2))
  for n in G["upper_trid_index"]["sizes"]:
    arr = [lib.orc_upper_trid_index(n, i, j) for i in range(n) for j in range(i, n)]
    assert arr == list(range(n * (n + 1) // 2))
  n, i, j = G["upper_trid_index"]["symmetric"]
  assert lib.orc_upper_trid_index(n, i, j) == lib.orc_upper_trid_index(n, j, i)


def test_nxn_pairs_follow_upper_tri_index(lib):
  """The filtered NXN pair list enumerates np.triu_indices order == upper_tri_index order (collision_driver_test.py:689)."""
  from mujoco_warp_b200._src import io as mio
  from mujoco_warp_b200._src import mjcf

  mjm = mjcf.load_any(util.HUMANOID)
  t = mio.derive_tables(mjm)
  ids = [lib.orc_upper_tri_index(mjm.ngeom, int(a), int(b)) for a, b in t["nxn_geom_pair"]]
  assert ids == list(range(len(ids)))


def test_padded_sizes():
  from mujoco_warp_b200._src import io as mio

  for nv, want in G["padded_sizes_augmented"]:
    assert mio._get_padded_sizes(nv, 0, False, 16, augment_cholesky=True)[1] == want
  # dense models: round_up(nv, 4); humanoid 27 -> 28
  assert mio._get_padded_sizes(27, 64, False) == (64, 28)


def test_halton_matches_definition(lib):
  """util_misc.py:61-76: radical inverse in the given (possibly non-prime) base, fp32 accumulation."""
  from oracle import orc

  def halton(index, base):
    f, hn, n0 = np.float32(1.0) / np.float32(base), np.float32(0), index
    while n0 > 0:
      n1 = n0 // base
      hn = np.float32(hn + f * np.float32(n0 - n1 * base))
      f = np.float32(f / np.float32(base))
      n0 = n1
    return float(hn)

  for idx in (1, 2, 7, 100, 8193, 65536 * 3 + 5):
    for base in (2, 3, 4, 10, 22):
      assert orc.halton(idx, base) == pytest.approx(halton(idx, base), abs=1e-7)


def test_humanoid_facts():
  from mujoco_warp_b200._src import mjcf

  mjm = mjcf.load_any(util.HUMANOID)
  f = G["humanoid_facts"]
  assert (mjm.nbody, mjm.nv, mjm.nu, mjm.ngeom) == (f["nbody"], f["nv"], f["nu"], f["ngeom"])
