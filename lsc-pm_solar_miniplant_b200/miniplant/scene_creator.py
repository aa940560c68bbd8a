"""Scene construction for the LSC-PM photomicroreactor — mirror of reference ``miniplant/scene_creator.py``.

Public names, arguments and the resulting scene graph are the reference's
(``_create_scene_common`` ``:52-274``, ``create_direct_scene`` ``:277-311``,
``create_diffuse_scene`` ``:314-331``, constants ``:33-49``).  New here: ``flatten(scene)``
turns the graph into the SoA arrays the CUDA library consumes (see ``../flatten.py``).

Scene (reference ``:58-272``): world sphere r=10 m (air) > PMMA plate 0.47 x 0.47 x 0.008 m
(n=1.48, background absorber 0.1 m^-1, optional Lumogen F Red 305 luminophore QY 0.95) > 16 PFA
capillaries (1/8" OD, n=1.34, 0.1 m^-1) each holding one reaction-mixture cylinder (1/16" ID,
n=1.344, methylene-blue ``Reactor``).  Capillaries run along the plate's y axis at
x_k = -0.225 + 0.03 k in the mid-plane; the plate is tilted about +y and shifted so that its top
face passes through the world origin.
"""
from __future__ import annotations

import io
import logging
import pkgutil
from functools import lru_cache
from typing import Callable

import numpy as np

from ..compat import pvtrace_namespace
from .utils import IsotropicPhotonGenerator, LightPosition, MyLight, VectorInverter

_pv = pvtrace_namespace()
Node, Box, Sphere, Cylinder = _pv.Node, _pv.Box, _pv.Sphere, _pv.Cylinder
Material, Luminophore, Absorber, Reactor = _pv.Material, _pv.Luminophore, _pv.Absorber, _pv.Reactor
Light, Scene, isotropic, spherical_to_cart = _pv.Light, _pv.Scene, _pv.isotropic, _pv.spherical_to_cart

REACTOR_AREA_IN_M2 = 0.47 * 0.47

MB_ABS_DATAFILE = pkgutil.get_data(__name__, "reactor_data/MB_1M_1m_ACN.tsv")
LR305_ABS_DATAFILE = pkgutil.get_data(__name__, "reactor_data/Evonik_lr305_normalized_to_1m.tsv")
LR305_EMS_DATAFILE = pkgutil.get_data(__name__, "reactor_data/Evonik_lr305_normalized_to_1m_ems.tsv")

# Refractive indexes
PMMA_RI = 1.48
PFA_RI = 1.34
ACN_RI = 1.344

# Units
INCH = 0.0254  # meters

# Geometry of the device (reference :89-90, :219-246, :231, :262)
PLATE_SIZE = (0.47, 0.47, 0.008)
N_CAPILLARIES = 16
CAPILLARY_PITCH = 0.03
CAPILLARY_FIRST_X = -0.47 / 2 + 0.01
PV_RI = 3.4
PV_ABSORPTION = 1e10


def _headerless_table(blob: bytes) -> np.ndarray:
    """Parse a two-column TSV the way the reference does: ``pd.read_csv(..., sep="\\t").values``
    without ``header=None`` (reference :75-80, :213-215), which consumes the FIRST DATA ROW as
    column names.  The tables the reference traces with therefore start at the second row
    (251 nm / 351 nm / 250.5 nm); parity requires dropping it here too."""
    rows = [line.split("\t") for line in blob.decode("utf8").splitlines() if line.strip()]
    return np.array([[float(a), float(b)] for a, b in rows[1:]], dtype=float)


@lru_cache(maxsize=None)
def _spectra() -> dict:
    tables = {
        "lr305_abs": _headerless_table(LR305_ABS_DATAFILE),
        "lr305_ems": _headerless_table(LR305_EMS_DATAFILE),
        "mb_abs": _headerless_table(MB_ABS_DATAFILE),
    }
    for table in tables.values():
        table.setflags(write=False)
    return tables


def _pv_material():
    return Material(
        refractive_index=PV_RI,
        components=[Absorber(coefficient=PV_ABSORPTION)],
        color=0x000000,
        transparent=False,
        opacity=0.5,
    )


def _create_scene_common(tilt_angle, light_source, include_dye=None, **kwargs) -> Scene:
    logger = logging.getLogger("pvtrace").getChild("miniplant")
    logger.debug(f"Creating simulation scene w/ angle={tilt_angle}deg...")
    spectra = _spectra()

    world = Node(
        name="World (air)",
        geometry=Sphere(radius=10.0, material=Material(refractive_index=1.0)),
    )
    light_source.parent = world

    # PMMA background absorption, plus the dye unless explicitly disabled (None means True, :69-70)
    matrix_component = [Absorber(coefficient=0.1)]
    if include_dye is None:
        include_dye = True
    if include_dye:
        matrix_component.append(
            Luminophore(
                coefficient=np.array(spectra["lr305_abs"]),
                emission=np.array(spectra["lr305_ems"]),
                quantum_yield=0.95,
                phase_function=isotropic,
            )
        )

    reactor = Node(
        name="LSC-PM Reactor 47x47 cm^2",
        geometry=Box(
            size=PLATE_SIZE,
            material=Material(refractive_index=PMMA_RI, components=matrix_component),
        ),
        parent=world,
    )

    # Optional photovoltaic absorbers around the plate (reference :99-206): one under the plate
    # behind an air gap, four strips sharing a face with the plate edges.
    if kwargs.get("add_bottom_PV", False):
        bottom = Node(name="bottomPV", geometry=Box(size=PLATE_SIZE, material=_pv_material()), parent=reactor)
        bottom.translate((0, 0, -0.025))

    if kwargs.get("add_side_PV", False):
        half = 0.47 / 2 + 0.005
        strips = (
            ("sidePV1", (0.47, 0.01, 0.008), (0, half, 0)),
            ("sidePV2", (0.47, 0.01, 0.008), (0, -half, 0)),
            ("sidePV3", (0.01, 0.47, 0.008), (half, 0, 0)),
            ("sidePV4", (0.01, 0.47, 0.008), (-half, 0, 0)),
        )
        for name, size, offset in strips:
            strip = Node(name=name, geometry=Box(size=size, material=_pv_material()), parent=reactor)
            strip.translate(offset)

    # Capillaries: ONE shared PFA cylinder geometry for all 16 tubes, a fresh reaction-mixture
    # cylinder per tube, all sharing one Reactor component (reference :216-257).
    reaction_mixture_material = Reactor(np.array(spectra["mb_abs"]))
    pfa_cil = Cylinder(
        length=0.47,
        radius=(1 / 8 * INCH) / 2,
        material=Material(refractive_index=PFA_RI, components=[Absorber(coefficient=0.1)]),
    )
    pfa_cil.color, pfa_cil.transparency, pfa_cil.opacity = 0xEEEEEE, True, 0.5

    # Extension for the scaled mini-plant (BASELINE config 5, dense channel grids): `capillary_count` / `capillary_pitch`
    # keyword arguments; the defaults reproduce the reference's 16 capillaries at 30 mm starting 10 mm from the edge.
    capillary_count = int(kwargs.get("capillary_count", N_CAPILLARIES))
    capillary_pitch = float(kwargs.get("capillary_pitch", CAPILLARY_PITCH))
    first_x = CAPILLARY_FIRST_X if (capillary_count, capillary_pitch) == (N_CAPILLARIES, CAPILLARY_PITCH) \
        else -0.5 * capillary_pitch * (capillary_count - 1)

    for capillary_num in range(capillary_count):
        tube = Node(name=f"Capillary_PFA_{capillary_num}", geometry=pfa_cil, parent=reactor)
        reaction_cil = Cylinder(
            length=0.47,
            radius=(1 / 16 * INCH) / 2,
            material=Material(refractive_index=ACN_RI, components=[reaction_mixture_material]),
        )
        reaction_cil.color, reaction_cil.transparency, reaction_cil.opacity = 0x0000FF, False, 1
        Node(name=f"Reaction_mixture_{capillary_num}", geometry=reaction_cil, parent=tube)

        # cylinder axis z -> y, then slide to its slot in the plate (reference :259-262)
        tube.rotate(np.radians(90), (1, 0, 0))
        tube.translate((first_x + capillary_pitch * capillary_num, 0, 0))

    # Tilt about +y, then drop by half the thickness along the tilted normal so that the top face
    # contains the world origin (reference :264-272)
    reactor.rotate(np.radians(tilt_angle), (0, 1, 0))
    half_thickness = 0.5 * PLATE_SIZE[2]
    reactor.translate(
        (
            -np.sin(np.deg2rad(tilt_angle)) * half_thickness,
            0,
            -np.cos(np.deg2rad(tilt_angle)) * half_thickness,
        )
    )
    return Scene(world)


def create_direct_scene(
    tilt_angle: float = 30,
    solar_elevation: float = 30,
    solar_azimuth: float = 180,
    solar_spectrum_function: Callable = lambda: 555,
    include_dye: bool = None,
    **kwargs,
) -> Scene:
    """Scene lit by a collimated beam from the sun position (reference :277-311): rays start on the
    plate's top face shifted 1 m towards the sun and travel along the negated sun vector."""
    solar_light_vector = spherical_to_cart(np.deg2rad(-solar_elevation + 90), np.deg2rad(-solar_azimuth + 180))
    solar_light = Node(
        name="Solar Light",
        light=Light(
            wavelength=solar_spectrum_function,
            direction=VectorInverter(solar_light_vector),
            position=LightPosition(tilt_angle=tilt_angle),
        ),
        parent=None,
    )
    solar_light.translate(solar_light_vector)
    return _create_scene_common(tilt_angle=tilt_angle, light_source=solar_light, include_dye=include_dye, **kwargs)


def create_diffuse_scene(
    tilt_angle: float = 30,
    solar_spectrum_function: Callable = lambda: 555,
    include_dye: bool = None,
):
    """Scene lit by sky-diffuse light: position and direction drawn jointly (reference :314-331)."""
    solar_light = Node(
        name="Solar Light",
        light=MyLight(
            wavelength=solar_spectrum_function,
            position_and_direction=IsotropicPhotonGenerator(tilt_angle),
        ),
        parent=None,
    )
    return _create_scene_common(tilt_angle=tilt_angle, light_source=solar_light, include_dye=include_dye)


def flatten(scene: Scene):
    """Flatten the scene graph into the SoA description of ``include/lscpm.h`` (``LscpmSceneDesc``)."""
    from ..flatten import flatten_scene

    return flatten_scene(scene)
