"""Scene construction through the mirrored ``scene_creator`` with the reference's own imports and checks (reference
tests/test_scene_creator.py:19-87): one light node carrying a ``Light``; the first ``Box`` has a ``Luminophore`` exactly
when the dye is requested (default: yes); the diffuse scene's light is a ``MyLight``; a plain ``anytree.Node`` works
as the light holder and ``scene.light_nodes`` is a list."""
import numpy as np
import pytest
from anytree import LevelOrderIter, Node
from miniplant.scene_creator import _create_scene_common, create_diffuse_scene, create_direct_scene
from miniplant.utils import MyLight
from pvtrace import Box, Light, Luminophore, Scene


def plain_anytree_light() -> Node:
    return Node(name="Test Light", light=Light(), parent=None)


def check_valid(scene):
    assert isinstance(scene, Scene)
    lights = scene.light_nodes
    assert len(lights) == 1 and isinstance(lights[0].light, Light)


def has_dye(scene) -> bool:
    plate = next(n for n in LevelOrderIter(scene.root) if isinstance(getattr(n, "geometry", None), Box))
    return any(isinstance(component, Luminophore) for component in plate.geometry.material.components)


@pytest.mark.parametrize("include_dye,expected", [(None, True), (True, True), (False, False)])
def test__create_scene_common(include_dye, expected):
    kwargs = {} if include_dye is None else {"include_dye": include_dye}
    scene = _create_scene_common(tilt_angle=0, light_source=plain_anytree_light(), **kwargs)
    check_valid(scene)
    assert has_dye(scene) is expected


@pytest.mark.parametrize("kwargs,expected", [({}, True), ({"include_dye": False}, False)])
def test_create_direct_scene_defaults(kwargs, expected):
    scene = create_direct_scene(**kwargs)
    check_valid(scene)
    assert has_dye(scene) is expected


def test_create_diffuse_scene_light_is_mylight():
    scene = create_diffuse_scene()
    assert isinstance(scene, Scene)
    assert len(scene.light_nodes) == 1
    assert isinstance(scene.light_nodes.pop().light, MyLight)     # light_nodes is a fresh list


def test_scene_graph_matches_reference_layout():
    """34 geometry nodes, shared PFA geometry, per-tube reaction cylinders sharing one Reactor component,
    capillaries at x_k = -0.225 + 0.03 k (reference scene_creator.py:216-262)."""
    scene = create_direct_scene(tilt_angle=0)
    nodes = [n for n in LevelOrderIter(scene.root) if getattr(n, "geometry", None) is not None]
    assert len(nodes) == 34
    tubes = [n for n in nodes if n.name.startswith("Capillary_PFA_")]
    mixes = [n for n in nodes if n.name.startswith("Reaction_mixture_")]
    assert len(tubes) == len(mixes) == 16
    assert len({id(t.geometry) for t in tubes}) == 1
    assert len({id(m.geometry) for m in mixes}) == 16
    assert len({id(m.geometry.material.components[0]) for m in mixes}) == 1
    xs = sorted(t.world_pose()[0, 3] for t in tubes)
    assert np.allclose(xs, -0.225 + 0.03 * np.arange(16))
    plate = next(n for n in nodes if n.name.startswith("LSC-PM"))
    assert np.allclose(plate.world_pose()[:3, 3], [0, 0, -0.004])     # top face through the world origin


def test_pv_kwargs_add_the_reference_boxes():
    scene = create_direct_scene(tilt_angle=0, add_bottom_PV=True, add_side_PV=True)
    names = [n.name for n in LevelOrderIter(scene.root)]
    assert "bottomPV" in names and {"sidePV1", "sidePV2", "sidePV3", "sidePV4"} <= set(names)
    bottom = next(n for n in LevelOrderIter(scene.root) if n.name == "bottomPV")
    assert bottom.geometry.material.refractive_index == 3.4
    assert np.allclose(bottom.world_pose()[:3, 3], [0, 0, -0.004 - 0.025])
