"""``pvtrace.photon_tracer`` / ``pvtrace.algorithm.photon_tracer`` surface, CUDA-backed.

The reference calls ``photon_tracer.follow(scene, ray)`` per photon and ``next_hit(scene, ray)`` on
the last ray (``miniplant/simulation_runner.py:38-48``).  Here both go to the CUDA library; there is
no CPU implementation in the product (the CPU restatement lives in ``oracle/`` and is test-only).
"""
from __future__ import annotations


def follow(scene, ray, maxsteps=1000, **kwargs):
    """Trace ONE ray to termination on the GPU and return its ``[(Ray, Event), ...]`` history."""
    from .backend import follow_ray

    return follow_ray(scene, ray, maxsteps=maxsteps, **kwargs)


def next_hit(scene, ray):
    """Next interface along the ray: ``(hit_node, (container, adjacent), point, distance)`` or None."""
    from .backend import next_hit_ray

    return next_hit_ray(scene, ray)
