"""The reference's own known-answer tests for the convex path (collision_gjk_test.py:307-1002), transcribed as data:
geoms, poses and the expected distance / contact count / witness points are the reference's; the harness calls the oracle's
convex pair routine the way `_geom_dist` (collision_gjk_test.py:35-303) calls `ccd` + `multicontact`.  The mesh cases live in
tests/test_mesh_gjk_vectors.py; height-field cases are left out (no height fields in this build).  `tests/test_gpu_gjk_vectors.py` runs the penetrating cases on the GPU."""
import numpy as np
import pytest

from mujoco_warp_b200._src import mjcf
from oracle import orc

BOX025 = '<geom size=".025 .025 .025" type="box"/><geom size=".025 .025 .025" type="box"/>'

# name -> dict(xml=worldbody content, pos1/mat1/pos2/mat2 overrides, multiccd, iterations, expectations)
CASES = {
  "spheres_distance": dict(  # :307
    xml='<geom type="sphere" pos="-1.5 0 0" size="1"/><geom type="sphere" pos="1.5 0 0" size="1"/>', dist=1.0, places=12, x1_0=-0.5, x2_0=0.5),
  "spheres_touching": dict(xml='<geom type="sphere" pos="-1 0 0" size="1"/><geom type="sphere" pos="1 0 0" size="1"/>', dist=0.0, places=12),  # :325
  "sphere_sphere_contact": dict(xml='<geom type="sphere" pos="-1 0 0" size="3"/><geom type="sphere" pos=" 3 0 0" size="3"/>', dist=-2.0),  # :368
  "box_box_contact": dict(  # :384
    xml='<geom type="box" pos="-1 0 0" size="2.5 2.5 2.5"/><geom type="box" pos="1.5 0 0" size="1 1 1"/>', dist=-1.0, normal=(1.0, 0.0, 0.0)),
  "cylinder_cylinder_contact": dict(  # :467
    xml='<geom pos="0 0 0" type="cylinder" size="1 .5"/><geom pos="1.999 0 0" type="cylinder" size="1 .5"/>', dist=-0.001),
  "box_box_shallow_penetration": dict(  # :483
    xml='<geom type="box" size="0.2 0.2 0.2" pos="0 0 0.19972974"/><geom type="box" size="0.1 0.1 0.1" pos="0 0 0.49947918"/>',
    multiccd=True, ncon=4, dist=-0.00025054812),
  "box_edge": dict(  # :499
    xml='<geom pos="0 0 2" type="box" size="1 1 1"/><geom pos="0 0 4.4" euler="0 90 40" type="box" size="1 1 1"/>', multiccd=True, ncon=2),
  "box_box_ccd": dict(xml='<geom type="box" pos="0 0 1.9" size="1 1 1"/><geom type="box" pos="0 0 0" size="10 10 1"/>', multiccd=True, ncon=4),  # :513
  "box_box_ccd2": dict(  # :548
    xml='<geom size="1 1 1" pos="0 0 2" type="box"/><geom size="1 1 1" pos="0 1 3.99" euler="0 0 40" type="box"/>', multiccd=True, ncon=4),
  "box_box_early": dict(  # :564 EPA terminates in its first iteration
    xml=BOX025, dist=-0.00396448, places=6,
    pos1=(0.07524700462818145752, -0.13524700701236724854, 0.12491077929735183716),
    mat1=(1.0, 0.00000000006837434091, 0.00000000080494955146, -0.00000000006837435479, 1.0, 0.00000002552030764491, -0.00000000080494955146, -0.00000002552030764491, 1.0),
    pos2=(0.07524700462818145752, -0.13524700701236724854, 0.17094630002975463867),
    mat2=(1.0, 0.00000000006837435479, -0.00000000018000903546, -0.00000000006837434091, 1.0, 0.00000004174798817758, 0.00000000018000903546, -0.00000004174798817758, 1.0)),
  "box_box_early2": dict(  # :606
    xml=BOX025, dist=-2.515156e-06,
    pos1=(0.07122065126895904541, -0.19126638770103454590, 0.29129269719123840332),
    mat1=(0.99999558925628662109, 0.00258362153545022011, 0.00148368685040622950, -0.00258197076618671417, 0.99999606609344482422, -0.00111339206341654062,
          -0.00148655765224248171, 0.00110955617856234312, 0.99999833106994628906),
    pos2=(0.07183132320642471313, -0.13260576128959655762, 0.30987158417701721191),
    mat2=(0.99827724695205688477, 0.02493947930634021759, 0.05311207473278045654, 0.00605074502527713776, 0.85659545660018920898, -0.51595354080200195312,
          -0.05836316198110580444, 0.51538598537445068359, 0.85496860742568969727)),
  "cylinder_box": dict(  # :668
    xml='<geom type="box" size="1 1 0.1"/><geom type="cylinder" size=".1 .2 .3"/>', iterations=50, dist=-0.0016624178339902445,
    pos2=(0.00015228791744448245, -0.00074981129728257656, 0.29839199781417846680),
    mat2=(0.99996972084045410156, 0.00776371126994490623, -0.00043433305108919740, -0.00776385562494397163, 0.99996984004974365234, -0.00033095158869400620,
          0.00043175052269361913, 0.00033431366318836808, 0.99999988079071044922)),
  "cylinder_capsule": dict(xml='<geom type="cylinder" size="2 4" pos="0 0 0"/><geom type="capsule" size="1 1" pos="0 0 5"/>', dist=-1.0, places=6),  # :698
  "box_box_float": dict(  # :713
    xml=BOX025, ncon=1, dist_less=0.0001,
    pos1=(-0.17624500393867492676, -0.12375499308109283447, 0.12499777972698211670),
    mat1=(1.0, -0.00000000184385418045, -0.00000025833372774287, 0.00000000184391857339, 1.0, 0.00000024928382913458, 0.00000025833372774287, -0.00000024928382913458, 1.0),
    pos2=(-0.17624500393867492676, -0.12375499308109283447, 0.17499557137489318848),
    mat2=(1.0, -0.00000000184292525685, 0.00000012980596864054, 0.00000000184294413064, 1.0, -0.00000014602545661546, -0.00000012980596864054, 0.00000014602545661546, 1.0)),
  "box_box_horizon": dict(  # :756 EPA horizon with 13 edges
    xml=BOX025, dist=-0.00011578822, places=6,
    pos1=(0.065118454396725, -0.125125020742416, 0.124963559210300),
    mat1=(0.996357858181000, 0.085266821086407, -0.000942531623878, -0.085266284644604, 0.996358215808868, 0.000591202871874, 0.000989508931525, -0.000508683384396, 0.999999582767487),
    pos2=(0.065104484558105, -0.124979749321938, 0.174992129206657),
    mat2=(0.996556758880615, -0.082913912832737, -0.000453041866422, 0.082915119826794, 0.996536433696747, 0.006357696373016, -0.000075668765930, -0.006373368669301, 0.999979794025421)),
  "box_box_rotation": dict(  # :816
    xml=BOX025, multiccd=True, ncon=4,
    pos1=(0.015344001352787, -0.195344015955925, 0.174637570977211),
    mat1=(1.0, 0.000000000029901, 0.000004057303613, -0.000000000062404, 1.0, 0.000008010840247, -0.000004057303613, -0.000008010840247, 1.0),
    pos2=(0.015344001352787, -0.195344015955925, 0.224056228995323),
    mat2=(1.0, 0.000000000029692, -0.000003355821491, -0.000000000057016, 1.0, -0.000008142159459, 0.000003355821491, 0.000008142159459, 1.0)),
  "box_box_diagonal": dict(  # :866 multiccd sees a face diagonal as its feature
    xml='<geom size="0.50 0.50 0.10" type="box"/><geom size=".025 .025 .025" type="box"/>', multiccd=True, ncon=4, dist=-1.5778851595232846e-05,
    pos2=(0.135535001754761, -0.195535004138947, 0.124984227120876),
    mat2=(1.0, 0.000000000048563, -0.000000135524601, -0.000000000048577, 1.0, -0.000000103374248, 0.000000135524601, 0.000000103374248, 1.0)),
  "box_box_max": dict(  # :900 needs 16 EPA iterations.  The expected value was produced in fp32 with poses ~20 m from the origin
    # (fp32 spacing 2e-6 there); the fp32 build meets it to 7 places, the double build lands 1.6e-7 away
    xml='<geom type="box" size=".018 .018 .01"/><geom type="box" size=".020 .020 .04"/>', multiccd=True, dist=-0.03636224, fp64_slack=5.0,
    mat1=(0.8378595710, 0.3184406757, -0.4433811009, 0.5328434706, -0.3006005287, 0.7910227776, 0.1186132580, -0.8990187645, -0.4215400815),
    pos1=(6.0405082703, 21.4734001160, 0.036854844),
    mat2=(-0.6420212388, -0.0727036372, -0.7632319927, 0.3801730871, -0.8946756721, -0.2345722020, -0.6657907367, -0.4407605529, 0.6020406485),
    pos2=(6.0641078949, 21.4842395782, 0.0212156791)),
  "box_box_max2": dict(  # :952 GJK converges very slowly
    xml='<geom type="box" size=".5 .5 .1"/><geom type="box" size=".025 .025 .025"/>', multiccd=True, dist=-4.9374998e-05,
    mat2=(0.9999999404, 0.0004342802, -0.0001755831, -0.0004346797, 0.9999973178, -0.0022819033, 0.0001745916, 0.0022819792, 0.9999974370),
    pos2=(0.0885666460, 0.0911745951, 0.1250119805)),
  "box_edge_flipped": dict(  # :986
    xml='<geom pos="1.10164554 -0.11389316 0.74" quat="-0.348312918 0 0 0.937378318" type="box" size="0.65 0.48 0.04"/>'
        '<geom type="box" size="0.1 1.2 1.4" pos="1.4 0 1.425"/>',
    multiccd=True, ncon=2, x1=(1.907368, -0.052973, 0.700000), x2=(1.30000, -0.052973, 0.700000), xtol=1e-4),
}


def posed_geoms(case):
  """(type, size, pos, mat) of the two geoms: from the XML, with the test's explicit pose overrides applied."""
  mjm = mjcf.load_string(f"<mujoco><worldbody>{case['xml']}</worldbody></mujoco>")
  kin = mjcf.kinematics_np(mjm, mjm.qpos0)
  xpos, xmat = np.asarray(kin.geom_xpos, dtype=np.float64).reshape(-1, 3), np.asarray(kin.geom_xmat, dtype=np.float64).reshape(-1, 9)
  out = []
  for g in (0, 1):
    pos = np.asarray(case.get(f"pos{g + 1}", xpos[g]), dtype=np.float64)
    mat = np.asarray(case.get(f"mat{g + 1}", xmat[g]), dtype=np.float64)
    out.append((int(mjm.geom_type[g]), np.asarray(mjm.geom_size[g], dtype=np.float64), pos, mat))
  return out


def check(case, dist, ncon, w1, w2, slack=1.0):
  """The reference's assertions (assertAlmostEqual(places) rounds the difference to `places` decimals)."""
  if "dist" in case:
    assert abs(dist - case["dist"]) < slack * 0.5 * 10.0 ** -case.get("places", 7), (dist, case["dist"])
  if "dist_less" in case:
    assert dist < case["dist_less"]
  if "ncon" in case:
    assert ncon == case["ncon"], (ncon, case["ncon"])
  if "x1_0" in case:
    assert abs(w1[0][0] - case["x1_0"]) < 1e-12 and abs(w2[0][0] - case["x2_0"]) < 1e-12
  if "normal" in case:
    n = (w1[0] - w2[0]) / np.linalg.norm(w1[0] - w2[0])
    np.testing.assert_allclose(n, case["normal"], atol=5e-8)
  if "x1" in case:
    np.testing.assert_allclose(w1[0], case["x1"], atol=case["xtol"])
    np.testing.assert_allclose(w2[0], case["x2"], atol=case["xtol"])


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_oracle_reproduces_reference_gjk_vectors(name, dtype):
  """fp64 build and the fp32 build (the reference's own arithmetic width) both meet the reference's assertions."""
  case = CASES[name]
  (t1, s1, p1, m1), (t2, s2, p2, m2) = posed_geoms(case)
  dist, ncon, w1, w2, ovf = orc.ccd(t1, s1, p1, m1, t2, s2, p2, m2, iterations=case.get("iterations", 35), multiccd=case.get("multiccd", False), dtype=dtype)
  assert ovf == 0
  check(case, dist, ncon, w1.astype(np.float64), w2.astype(np.float64), slack=case.get("fp64_slack", 1.0) if dtype is np.float64 else 1.0)
