"""The reference's scene-construction tests (reference tests/test_scene_creator.py:19-87), against the mirrored
scene_creator: same imports, same assertions (including the plain anytree.Node light and light_nodes.pop())."""
from anytree import LevelOrderIter, Node
from miniplant.scene_creator import _create_scene_common, create_diffuse_scene, create_direct_scene
from miniplant.utils import MyLight
from pvtrace import Box, Light, Luminophore, Scene


def get_light_node() -> Node:
    return Node(name="Test Light", light=Light(), parent=None)


def scene_is_valid(scene: Scene):
    assert isinstance(scene, Scene)
    assert len(scene.light_nodes) == 1
    light = scene.light_nodes[0]
    assert isinstance(light.light, Light)


def scene_has_dye(scene: Scene):
    for node in LevelOrderIter(scene.root):
        if hasattr(node, "geometry") and isinstance(node.geometry, Box):
            return any(isinstance(x, Luminophore) for x in node.geometry.material.components)


def test__create_scene_common():
    scene = _create_scene_common(tilt_angle=0, light_source=get_light_node())
    scene_is_valid(scene)
    assert scene_has_dye(scene)


def test__create_scene_common_w_dye():
    assert scene_has_dye(_create_scene_common(tilt_angle=0, light_source=get_light_node(), include_dye=True))


def test__create_scene_common_wo_dye():
    assert not scene_has_dye(_create_scene_common(tilt_angle=0, light_source=get_light_node(), include_dye=False))


def test_create_diffuse_scene():
    scene = create_direct_scene()
    scene_is_valid(scene)
    assert scene_has_dye(scene)


def test_create_diffuse_scene_wo_dye():
    scene = create_direct_scene(include_dye=False)
    scene_is_valid(scene)
    assert not scene_has_dye(scene)


def test_create_direct_scene():
    scene = create_diffuse_scene()
    assert isinstance(scene, Scene)
    assert len(scene.light_nodes) == 1
    light = scene.light_nodes.pop()
    assert isinstance(light.light, MyLight)


def test_scene_graph_matches_reference_layout():
    """34 geometry nodes, shared PFA geometry, per-tube reaction cylinders sharing one Reactor component,
    capillaries at x_k = -0.225 + 0.03 k (reference scene_creator.py:216-262)."""
    import numpy as np

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
