"""The s-t plugin interface kept verbatim on the host (reference lib/st_selection.py:15-35).

A selector is any object with

    get_source_target(agent_type=None, current_location=None, trip_count=0) -> (source, target)

working on node *labels*.  Inside the DRIFT application the reference's own five models
(RandomSelection, GravitySelection, HubAndSpokeSelection, ActivityBasedSelection,
ZoneBasedSelection) are passed in unchanged; ``make_selector`` imports them from ``lib.st_selection``
when that package is importable.  Only the trivial random model is restated here so that the
engine can run without the reference tree (GPU box, tests).  Draws come from the global ``random``
module exactly like the reference, so a run seeded the same way consumes the same stream.
"""
from __future__ import annotations

import random


class SourceTargetSelection:
    """Base class: same constructor and method as the reference's (lib/st_selection.py:15-35)."""

    def __init__(self, graph):
        self.graph = graph
        self.nodes = list(graph.nodes)

    def get_source_target(self, agent_type=None, current_location=None, trip_count=0):
        raise NotImplementedError


class RandomSelection(SourceTargetSelection):
    """Uniform target different from the source (behaviour of lib/st_selection.py:37-51)."""

    def get_source_target(self, agent_type=None, current_location=None, trip_count=0):
        source = random.choice(self.nodes) if current_location is None else current_location
        target = random.choice(self.nodes)
        while target == source:
            target = random.choice(self.nodes)
        return source, target


def make_selector(mode, graph, settings=None, fast=False):
    """Selector for ``selection_mode`` (config.py:174), built with the same arguments the reference
    passes at lib/simulation_thread.py:201-261.  Non-random models come from the reference package;
    ``fast=True`` takes the vectorised restatements of drift_b200/fast_selection.py (gravity, hub, zone:
    same sequences for the same seed, no O(V) Python work per call, no O(V^2) set-up) where they exist."""
    if fast:
        from .fast_selection import make_fast_selector

        sel = make_fast_selector(mode, graph, settings)
        if sel is not None:
            return sel
    if mode not in ("activity", "zone", "gravity", "hub"):
        try:
            from lib.st_selection import RandomSelection as RefRandom  # inside the DRIFT app

            return RefRandom(graph)
        except Exception:
            return RandomSelection(graph)
    try:
        from lib import st_selection as ref
    except Exception:
        # outside the DRIFT application (headless box): the restatements of drift_b200/fast_selection.py give the
        # same (source, target) sequences for the same seed (tests/test_fast_selection.py)
        from .fast_selection import make_fast_selector

        return make_fast_selector(mode, graph, settings)
    s = settings or {}
    if mode == "activity":
        if "agent_type_distributions" in s:
            return ref.ActivityBasedSelection(graph, agent_type_distributions=s["agent_type_distributions"])
        return ref.ActivityBasedSelection(graph)
    if mode == "zone":
        if "zone_intra_probability" in s:
            return ref.ZoneBasedSelection(graph, intra_zone_probability=s["zone_intra_probability"])
        return ref.ZoneBasedSelection(graph)
    if mode == "gravity":
        if settings:
            return ref.GravitySelection(graph, alpha=s.get("gravity_alpha", 1.0), beta=s.get("gravity_beta", 2.0))
        return ref.GravitySelection(graph)
    if settings:
        return ref.HubAndSpokeSelection(graph, hub_trip_probability=s.get("hub_trip_probability", 0.3),
                                        hub_percentage=s.get("hub_percentage", 0.1))
    return ref.HubAndSpokeSelection(graph)
