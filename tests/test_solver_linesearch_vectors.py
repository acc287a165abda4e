"""The reference's known-answer test of the elliptic line-search evaluation (solver_test.py:296-350
`test_elliptic_shifted_cost_preserves_small_delta`: five transitions between the cone's zones, chosen with powers of two so that the exact
cost delta lies below the ulp of the absolute cost) transcribed as data, and run through
 * the oracle (fp32 and fp64 builds): `orc_elliptic_eval_pt` = `_eval_elliptic_reference` + `_eval_elliptic_shifted`, and
 * the CUDA solver's own source: `mujoco_warp_b200/csrc/mjb_linesearch.cuh` compiled as host C++ (tests/host_harness/linesearch_host.cpp),
   the functions `k_solver` calls per contact in every line-search evaluation.
Plus a randomised cross-check of the device source against the oracle for elliptic contacts and for the three plain row kinds."""

import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import orc

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_harness", "linesearch_host.cpp")
OUT = os.path.join(HERE, "host_harness", "_build", "liblinesearch_host.so")
CSRC = os.path.join(HERE, "..", "mujoco_warp_b200", "csrc")

# name, alpha, quad, quad1, quad2, expected (cost(alpha) - cost(0), grad, hess); friction mu = 1, impratio_invsqrt = 1  (solver_test.py:296-330)
CASES = [
  ("middle_zone", 1.0e-4, (0.0, 0.0, 0.0), (0.5, 1.0e-4, 1.0), (0.0, 0.0, 1.0e8), (-0.5, -5000.0, 1.0)),
  ("quadratic_zone", 1.0e-4, (1.25e7, -5000.0, 0.5), (-2.0, 0.0, 1.0), (0.0, 0.0, 1.0), (-0.5, -5000.0, 1.0)),
  ("cone_to_quadratic", 2.44140625e-4, (33546241.0, -8192.0, 33554432.0), (-0.999755859375, -1.0, 1.0), (-1.0, 1.0, 16777216.0), (0.5, 8192.0, 67108864.0)),
  ("quadratic_to_cone", 2.44140625e-4, (33546241.0, -8192.0, 33554432.0), (-1.0, 1.0, 0.9995117783546448), (0.999755859375, 1.0, 16777216.0), (-0.5, 0.0, 0.0)),
  ("satisfied_to_quadratic", 2.44140625e-4, (1.0, -16384.0, 67108864.0), (2.44140625e-4, -2.0, 0.0), (0.0, 0.0, 16777216.0), (1.0, 16384.0, 134217728.0)),
]


@pytest.fixture(scope="module")
def hlib():
  deps = [SRC] + [os.path.join(CSRC, f) for f in ("mjb_linesearch.cuh", "mjb_math.cuh", "mjb_types.cuh")]
  if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-w", "-x", "c++", "-ffp-contract=off", f"-I{cuda_inc}", SRC, "-o", OUT], check=True)
  lib = ctypes.CDLL(OUT)
  V, F, I = ctypes.c_void_p, ctypes.c_float, ctypes.c_int
  lib.hls_elliptic_eval_pt.argtypes = [F, V, V, V, F, V]
  lib.hls_elliptic_zero.argtypes = [V, V, V, F, V]
  lib.hls_eval_row.argtypes = [I, F, I, I, F, F, F, F, V]
  lib.hls_eval_row_zero.argtypes = [I, I, I, F, F, F, F, V]
  return lib


def _ptr(a):
  return a.ctypes.data_as(ctypes.c_void_p)


def oracle_elliptic(real_bytes, alpha, quad, quad1, quad2, mu=1.0, zero=False):
  lib = orc._lib(real_bytes)
  dt = np.float32 if real_bytes == 4 else np.float64
  q, q1, q2, out = (np.asarray(x, dtype=dt) for x in (quad, quad1, quad2, np.zeros(3)))
  if zero:
    lib.orc_elliptic_zero.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_double, ctypes.c_void_p]
    lib.orc_elliptic_zero(_ptr(q), _ptr(q1), _ptr(q2), mu, _ptr(out))
  else:
    lib.orc_elliptic_eval_pt.argtypes = [ctypes.c_double] + [ctypes.c_void_p] * 3 + [ctypes.c_double, ctypes.c_void_p]
    lib.orc_elliptic_eval_pt(alpha, _ptr(q), _ptr(q1), _ptr(q2), mu, _ptr(out))
  return out.astype(np.float64)


def device_elliptic(lib, alpha, quad, quad1, quad2, mu=1.0, zero=False):
  q, q1, q2, out = (np.asarray(x, dtype=np.float32) for x in (quad, quad1, quad2, np.zeros(3)))
  if zero:
    lib.hls_elliptic_zero(_ptr(q), _ptr(q1), _ptr(q2), mu, _ptr(out))
  else:
    lib.hls_elliptic_eval_pt(alpha, _ptr(q), _ptr(q1), _ptr(q2), mu, _ptr(out))
  return out.astype(np.float64)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_reference_elliptic_shifted_cost_vectors(hlib, case):
  _, alpha, quad, quad1, quad2, expected = case
  # the reference asserts its fp32 kernel at rtol = atol = 1e-6
  np.testing.assert_allclose(oracle_elliptic(4, alpha, quad, quad1, quad2), expected, rtol=1e-6, atol=1e-6, err_msg="oracle fp32")
  np.testing.assert_allclose(device_elliptic(hlib, alpha, quad, quad1, quad2), expected, rtol=1e-6, atol=1e-6, err_msg="device source on the host")
  # in double the same formulas give the exact small delta as well
  np.testing.assert_allclose(oracle_elliptic(8, alpha, quad, quad1, quad2), expected, rtol=1e-6, atol=1e-6, err_msg="oracle fp64")


def test_device_elliptic_evaluation_matches_the_oracle_on_random_contacts(hlib):
  rng = np.random.default_rng(5)
  zones = set()
  for _ in range(4000):
    mu = float(rng.uniform(0.2, 1.5))
    # a contact with normal residual u0, tangential residual (norm sqrt(uu)), search components v0, (uv, vv), stiffness dm
    u0, v0 = float(rng.normal(0, 1.0)), float(rng.normal(0, 1.0))
    t, s = rng.normal(0, 1.0, 2), rng.normal(0, 1.0, 2)
    if rng.random() < 0.1:
      t[:] = 0  # degenerate tangent
    uu, uv, vv = float(t @ t), float(t @ s), float(s @ s)
    dm = float(rng.uniform(10.0, 1e4))
    D = dm * mu * mu * (1 + mu * mu)
    q0 = 0.5 * D * (u0 * u0 / (mu * mu) + uu)
    q1 = D * (u0 * v0 / (mu * mu) + uv)
    q2 = 0.5 * D * (v0 * v0 / (mu * mu) + vv)
    alpha = float(rng.uniform(-0.5, 1.5))
    quad, quad1, quad2 = (q0, q1, q2), (u0, v0, uu), (uv, vv, dm)
    f32 = lambda x: tuple(float(np.float32(v)) for v in x)
    quad, quad1, quad2, alpha, mu = f32(quad), f32(quad1), f32(quad2), float(np.float32(alpha)), float(np.float32(mu))
    for zero in (False, True):
      want = oracle_elliptic(8, alpha, quad, quad1, quad2, mu, zero)
      got = device_elliptic(hlib, alpha, quad, quad1, quad2, mu, zero)
      o32 = oracle_elliptic(4, alpha, quad, quad1, quad2, mu, zero)
      scale = max(1.0, abs(quad[0]), abs(want).max())
      np.testing.assert_allclose(got, o32, rtol=2e-5, atol=2e-5 * scale)   # same formulas in fp32: rounding-level agreement
      np.testing.assert_allclose(got, want, rtol=2e-3, atol=2e-4 * scale)  # and both track the double evaluation
    zones.add((want[1] == 0.0, want[2] == 2 * np.float32(quad[2])))
  assert len(zones) >= 3  # satisfied (zero gradient), quadratic (hess = 2 q2) and cone zone were all visited


def test_device_row_evaluation_matches_closed_forms(hlib):
  """eval_row / eval_row_zero for the three row kinds against the piecewise-quadratic cost written out directly (solver.py:425-517)."""
  rng = np.random.default_rng(9)
  out, out0 = np.zeros(3, np.float32), np.zeros(3, np.float32)
  for _ in range(3000):
    D, f = float(rng.uniform(1.0, 1e3)), float(rng.uniform(0.01, 2.0))
    ja, jv, alpha = float(rng.normal(0, 0.5)), float(rng.normal(0, 1.0)), float(rng.uniform(-0.5, 1.5))
    ne, nf = 2, 3
    for r, kind in ((0, "eq"), (3, "fric"), (7, "ineq")):
      hlib.hls_eval_row(r, alpha, ne, nf, D, f, ja, jv, _ptr(out))
      hlib.hls_eval_row_zero(r, ne, nf, D, f, ja, jv, _ptr(out0))

      def cost(x):
        if kind == "eq":
          return 0.5 * D * x * x, D * x, D
        if kind == "ineq":
          return (0.5 * D * x * x, D * x, D) if x < 0 else (0.0, 0.0, 0.0)
        rf = f / D
        if -rf < x < rf:
          return 0.5 * D * x * x, D * x, D
        return (f * (-0.5 * rf - x), -f, 0.0) if x <= -rf else (f * (-0.5 * rf + x), f, 0.0)

      c0, g0, h0 = cost(ja)
      c1, g1, h1 = cost(ja + alpha * jv)
      scale = max(1.0, abs(c0), abs(c1))
      np.testing.assert_allclose(out0, (c0, g0 * jv, h0 * jv * jv), rtol=1e-4, atol=1e-4 * scale)
      np.testing.assert_allclose(out, (c1 - c0, g1 * jv, h1 * jv * jv), rtol=1e-4, atol=2e-4 * scale)
