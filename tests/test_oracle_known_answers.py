"""The oracle against analytically derived known-answer vectors (SURVEY.md Appendix E — derived from the
pvtrace formulas, not from reference output: the reference pins no per-ray values) and the Philox generator
against the published Random123 vectors."""
import ctypes as C
import math
from collections import Counter

import numpy as np
import pytest

import pvtrace_oracle as po
from helpers import standard_scene


def test_critical_angles_and_normal_incidence():
    n_pmma, n_pfa, n_acn = 1.48, 1.34, 1.344
    for (n1, n2), deg in {(n_pmma, 1.0): 42.506642, (n_pmma, n_pfa): 64.877940, (n_pfa, 1.0): 48.268183,
                          (n_acn, n_pfa): 85.578438}.items():
        assert math.degrees(math.asin(n2 / n1)) == pytest.approx(deg, abs=1e-5)
        assert po.fresnel_reflectivity(math.radians(deg + 1e-3), n1, n2) == 1.0
        assert po.fresnel_reflectivity(math.radians(deg - 1e-3), n1, n2) < 1.0
    assert po.fresnel_reflectivity(0.0, 1.0, n_pmma) == pytest.approx(0.0374609781, rel=1e-8)
    assert po.fresnel_reflectivity(0.0, n_pmma, n_pfa) == pytest.approx(0.00246466476, rel=1e-8)
    assert po.fresnel_reflectivity(0.0, n_pfa, n_acn) == pytest.approx(2.22103274e-6, rel=1e-7)


@pytest.mark.parametrize("deg,r,dx,dz", [(40, 0.0430463296, 0.434315952, -0.900760597),
                                         (60, 0.0858444380, 0.585152300, -0.810923416),
                                         (80, 0.384207216, 0.665410644, -0.746477511)])
def test_air_to_pmma(deg, r, dx, dz):
    th = math.radians(deg)
    d = np.array([math.sin(th), 0.0, -math.cos(th)])
    n = po.facing_normal(d, np.array([0.0, 0.0, 1.0]))
    assert po.fresnel_reflectivity(math.acos(float(np.dot(n, d))), 1.0, 1.48) == pytest.approx(r, rel=1e-8)
    t = po.fresnel_refraction(d, n, 1.0, 1.48)
    assert t[0] == pytest.approx(dx, rel=1e-8) and t[2] == pytest.approx(dz, rel=1e-8)
    assert np.linalg.norm(t) == pytest.approx(1.0, abs=1e-12)
    refl = po.specular_reflection(d, n)
    assert refl == pytest.approx([d[0], 0.0, -d[2]])


@pytest.mark.parametrize("n1,n2,deg,r", [(1.48, 1.0, 30, 0.0509874714), (1.48, 1.0, 42, 0.461546411), (1.48, 1.0, 43, 1.0),
                                         (1.48, 1.34, 60, 0.0710313390), (1.48, 1.34, 64, 0.323151121), (1.48, 1.34, 65, 1.0)])
def test_internal_reflection(n1, n2, deg, r):
    assert po.fresnel_reflectivity(math.radians(deg), n1, n2) == pytest.approx(r, rel=1e-8)


def test_capillary_intersections():
    """E6-E8 in world coordinates of the tilt-0 scene (plate top face at z = 0)."""
    scene = po.OracleScene(standard_scene(tilt=0, elevation=90, wavelength=585.0))
    names = [n.name for n in scene.nodes]
    hits = scene.intersections(np.array([-0.1055, 0.0, 0.0]), np.array([0.0, 0.0, -1.0]))
    got = {(names[i], round(t, 11)) for t, i, _, _ in hits}
    for name, t in (("Capillary_PFA_4", 0.00249329623), ("Capillary_PFA_4", 0.00550670377),
                    ("Reaction_mixture_4", 0.00338352692), ("Reaction_mixture_4", 0.00461647308),
                    ("LSC-PM Reactor 47x47 cm^2", 0.008)):
        assert (name, t) in got
    hits = scene.intersections(np.array([-0.1065, 0.0, 0.0]), np.array([0.0, 0.0, -1.0]))
    got = {(names[i], round(t, 11)) for t, i, _, _ in hits}
    assert ("Capillary_PFA_4", 0.00348023443) in got and ("Capillary_PFA_4", 0.00451976557) in got
    assert not any(n == "Reaction_mixture_4" for n, _ in got)
    d = np.array([0.434315952, 0.0, -0.900760597])
    nh = scene.next_hit(np.array([-0.1, 0.05, -1e-9]), d / np.linalg.norm(d))
    assert names[nh[0]] == "LSC-PM Reactor 47x47 cm^2" and nh[5] == 4            # bottom face
    assert nh[4] == pytest.approx(0.00888138316, rel=1e-6)


def test_spectra_tables_and_emission():
    """E9-E11: header-dropped tables, kT shift, emission CDF.  App. E11 was derived with a cumulative trapezoid; the
    reference's committed results pin pvtrace's Distribution to cumsum(y)/max instead, which
    shifts the emission quantiles by half a knot (0.5 nm)."""
    from lscpm_b200.miniplant import scene_creator as sc

    spectra = sc._spectra()
    assert spectra["lr305_abs"].shape == (451, 2) and spectra["lr305_abs"][0, 0] == 251
    assert spectra["lr305_ems"].shape == (450, 2) and spectra["lr305_ems"][0, 0] == 351
    assert spectra["mb_abs"].shape == (1300, 2) and spectra["mb_abs"][0, 0] == 250.5
    dye = lambda nm: float(np.interp(nm, spectra["lr305_abs"][:, 0], spectra["lr305_abs"][:, 1]))
    mb = lambda nm: float(np.interp(nm, spectra["mb_abs"][:, 0], spectra["mb_abs"][:, 1]))
    assert dye(555) == pytest.approx(411.5233333, rel=1e-8) and dye(585) == pytest.approx(431.5633333, rel=1e-8)
    assert dye(650) == pytest.approx(10.97666667, rel=1e-8)
    assert mb(555) == pytest.approx(922700) and mb(585) == pytest.approx(2563600) and mb(650) == pytest.approx(10727200)
    assert po.kT_shifted_wavelength(555) == pytest.approx(545.537601, abs=2e-6)
    assert po.kT_shifted_wavelength(585) == pytest.approx(574.496671, abs=2e-6)
    assert po.kT_shifted_wavelength(650) == pytest.approx(637.058744, abs=2e-6)
    from lscpm_b200.compat.pvtrace_api import Distribution

    ems = Distribution(spectra["lr305_ems"][:, 0], spectra["lr305_ems"][:, 1])
    x, y = spectra["lr305_ems"][:, 0], spectra["lr305_ems"][:, 1]
    cdf = np.cumsum(y) / np.sum(y)
    assert np.allclose(ems._cdf, cdf) and ems._cdf[0] == pytest.approx(1e-3 / np.sum(y))
    assert float(ems.lookup(po.kT_shifted_wavelength(555))) == pytest.approx(0.00366148065, rel=1e-7)
    assert float(ems.lookup(po.kT_shifted_wavelength(585))) == pytest.approx(0.00495664298, rel=1e-7)
    assert float(ems.lookup(po.kT_shifted_wavelength(650))) == pytest.approx(0.274291580, rel=1e-7)
    trap = np.concatenate([[0.0], np.cumsum(0.5 * (y[1:] + y[:-1]) * np.diff(x))])
    trap /= trap[-1]
    for q, nm in ((0.1, 621.6216), (0.25, 635.1308), (0.5, 654.5908), (0.75, 680.9917), (0.9, 718.8948)):
        assert float(ems.sample(q)) == pytest.approx(nm, abs=1e-3)
        assert abs(float(ems.sample(q)) - float(np.interp(q, trap, x))) < 0.51      # App. E11 values, half a knot away


def test_distribution_rejects_negative_and_matches_numpy():
    from lscpm_b200.compat.pvtrace_api import Distribution

    with pytest.raises(ValueError):
        Distribution(np.array([1.0, 2.0]), np.array([1.0, -1.0]))
    x = np.array([360.0, 370.0, 400.0, 460.0])
    y = np.array([1.0, 3.0, 0.0, 2.0])
    d = Distribution(x, y)
    cdf = np.array([1.0, 4.0, 4.0, 6.0]) / 6.0          # cumsum(y) / max: knot spacing does not enter
    assert np.allclose(d._cdf, cdf)
    assert float(d.sample(0.5)) == pytest.approx(float(np.interp(0.5, cdf, x)))


def test_philox_known_answers(built_library):
    """Random123 kat_vectors for philox4x32-10, evaluated by the very function the kernels inline."""
    def philox(counter, key):
        c = (C.c_uint32 * 4)(*counter)
        k = (C.c_uint32 * 2)(*key)
        out = (C.c_uint32 * 4)()
        built_library.lscpm_philox4x32_10(c, k, out)
        return list(out)

    assert philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_container_and_adjacent_bookkeeping():
    """pvtrace ``next_hit`` / ``find_container``: container = nearest node hit exactly once; adjacent = the hit node, or
    - when the photon leaves its container - the container of what lies beyond the surface; coincident faces in random
    order.  Pinned on the reference's committed results; these are the consequences in the scenes
    of the path."""
    scene = po.OracleScene(standard_scene(tilt=0, elevation=90, wavelength=585.0))
    names = [n.name for n in scene.nodes]
    plate, world = "LSC-PM Reactor 47x47 cm^2", "World (air)"
    x4, zc, r_o, r_i = -0.105, -0.004, 0.0015875, 0.00079375

    def hit(p, d):
        d = np.asarray(d, dtype=float)
        h, c, a = scene.next_hit(np.asarray(p, dtype=float), d / np.linalg.norm(d))[:3]
        return names[h], names[c], None if a is None else names[a]

    # in the tube wall of capillary 4, leaving sideways with capillary 5 straight ahead (30 mm on, same height): the
    # adjacent node is the plate, not the sibling
    assert hit([x4 + 0.5 * (r_o + r_i), 0.0, zc], [1.0, 0.0, 0.0]) == ("Capillary_PFA_4", "Capillary_PFA_4", plate)
    # in the tube wall heading for the mixture, and in the mixture heading out
    assert hit([x4 + 0.5 * (r_o + r_i), 0.0, zc], [-1.0, 0.0, 0.0]) == ("Reaction_mixture_4", "Capillary_PFA_4", "Reaction_mixture_4")
    assert hit([x4, 0.0, zc], [1.0, 0.0, 0.2]) == ("Reaction_mixture_4", "Reaction_mixture_4", "Capillary_PFA_4")
    # in the plate: capillary ahead / top face ahead
    assert hit([x4 + 0.01, 0.0, zc], [-1.0, 0.0, 0.0]) == ("Capillary_PFA_4", plate, "Capillary_PFA_4")
    assert hit([x4 + 0.01, 0.0, zc], [0.0, 0.0, 1.0]) == (plate, plate, world)
    # along the tube wall / the mixture to the end cap, which coincides with the plate's +y face: without a generator the
    # tied faces keep scene-graph order (plate, tube, mixture) ...
    assert hit([x4 + 0.5 * (r_o + r_i), 0.2, zc], [0.0, 1.0, 0.0]) == (plate, plate, "Capillary_PFA_4")
    assert hit([x4, 0.2, zc], [0.0, 1.0, 0.0]) == (plate, plate, "Capillary_PFA_4")
    # ... and with one (the photon loop) every order occurs: the first of the tied nodes is hit node and container, the
    # second is adjacent
    rng = np.random.default_rng(5)
    seen = Counter()
    for _ in range(600):
        h, c, a = scene.next_hit(np.array([x4, 0.2, zc]), np.array([0.0, 1.0, 0.0]), rng)[:3]
        assert h == c and a != h
        seen[(names[h], names[a])] += 1
    assert len(seen) == 6 and set(n for pair in seen for n in pair) == {plate, "Capillary_PFA_4", "Reaction_mixture_4"}
    assert min(seen.values()) > 60 and max(seen.values()) < 140           # 100 expected each
    seen = Counter(names[scene.next_hit(np.array([x4 + 0.5 * (r_o + r_i), 0.2, zc]), np.array([0.0, 1.0, 0.0]), rng)[1]]
                   for _ in range(400))
    assert set(seen) == {plate, "Capillary_PFA_4"} and min(seen.values()) > 150
    # away from the caps the generator changes nothing
    assert scene.next_hit(np.array([x4, 0.0, zc]), np.array([1.0, 0.0, 0.2]), rng)[:3] == \
        scene.next_hit(np.array([x4, 0.0, zc]), np.array([1.0, 0.0, 0.2]))[:3]

    # photovoltaic box under the plate (air gap): a photon leaving the plate meets air.  Side boxes share the plate's edge
    # faces: in scene-graph order the plate's face comes first and the photon meets air as well (with the box's face
    # first - half of the time in the photon loop - it enters the box directly)
    pv = po.OracleScene(standard_scene(tilt=0, elevation=90, wavelength=585.0, add_bottom_PV=True, add_side_PV=True))
    names = [n.name for n in pv.nodes]
    scene = pv
    assert hit([x4 + 0.01, 0.0, zc], [0.0, 0.0, -1.0]) == (plate, plate, world)
    assert hit([0.2, 0.0, zc], [1.0, 0.0, 0.1]) == (plate, plate, world)
    # ... and in the air gap the box ahead is entered from the world
    assert hit([0.0, 0.0, -0.0085], [0.0, 0.0, -1.0])[1:] == (world, "bottomPV")
