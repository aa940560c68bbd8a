"""Naming contract of the per-node tally bins (host logic; mirrors ``lscpm_desc_bin_name`` in
``csrc/bins.cpp`` and the classification in ``oracle/``).

The reference reduces every photon history to its final event and reports ``count[REACT] / N``
(``miniplant/simulation_runner.py:40-41,63,65-66``).  The same mechanism defines every bin:

``kill``                      step limit reached (pvtrace ``maxsteps=1000``)
``exit/miss``                 left the scene without touching any surface
``exit/entry_reflect``        Fresnel-reflected at first contact, never entered
``exit/<node>/<face>``        left the scene; <node>/<face> is the last surface it interacted with
``absorb/<node>/<c>``         absorbed in <node> by material component <c> (background absorber, or a
                              luminophore whose quantum-yield draw failed)
``react/<node>/<c>``          absorbed by a ``Reactor`` component (the reference's ``Event.REACT``)

Layout (G geometry nodes in pre-order, node 0 = world): ``0 kill, 1 exit/miss, 2 exit/entry_reflect``,
then six face slots per non-world node, then one slot per (node, component).  Every row carries
``N_COUNTERS`` trailing event counters (interaction events, emissions) that are not photon bins.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence

import numpy as np

N_COUNTERS = 2
BIN_KILL, BIN_MISS, BIN_ENTRY_REFLECT, BIN_FIRST_EXIT = 0, 1, 2, 3
FACES_PER_NODE = 6
BOX_FACES = ("-x", "+x", "-y", "+y", "-z", "+z")
CYL_FACES = ("side", "-cap", "+cap")
GEOM_SPHERE, GEOM_BOX, GEOM_CYLINDER = 0, 1, 2
COMP_ABSORBER, COMP_LUMINOPHORE, COMP_REACTOR = 0, 1, 2


def _face_name(geom: int, face: int) -> str:
    if geom == GEOM_BOX:
        return BOX_FACES[face]
    if geom == GEOM_CYLINDER and face < 3:
        return CYL_FACES[face]
    if geom == GEOM_SPHERE and face == 0:
        return "surface"
    return f"unused{face}"


def bin_names(node_names: Sequence[str], node_geom: Sequence[int], node_ncomp: Sequence[int],
              node_comp_kinds: Sequence[Sequence[int]]) -> list:
    names = ["kill", "exit/miss", "exit/entry_reflect"]
    for i in range(1, len(node_names)):
        for f in range(FACES_PER_NODE):
            names.append(f"exit/{node_names[i]}/{_face_name(int(node_geom[i]), f)}")
    for i, name in enumerate(node_names):
        for c in range(int(node_ncomp[i])):
            prefix = "react" if int(node_comp_kinds[i][c]) == COMP_REACTOR else "absorb"
            names.append(f"{prefix}/{name}/{c}")
    return names


@dataclass
class TallyResult:
    """Tallies of one batch: integer photon bins whose sum is the photon count, plus event counters."""

    names: list
    counts: np.ndarray  # int64 [n_bins]
    events: int
    emissions: int

    @property
    def num_photons(self) -> int:
        return int(self.counts.sum())

    def fraction(self, prefix: str) -> float:
        n = self.num_photons
        total = sum(int(c) for name, c in zip(self.names, self.counts) if name.startswith(prefix))
        return total / n if n else 0.0

    @property
    def reacted_fraction(self) -> float:
        """The reference's scalar: ``Counter(finals)[Event.REACT] / num_photons``."""
        return self.fraction("react/")

    def nonzero(self) -> dict:
        return {name: int(c) for name, c in zip(self.names, self.counts) if c}
