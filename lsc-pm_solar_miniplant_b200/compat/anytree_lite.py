"""Minimal anytree-compatible tree (host logic, no tracing arithmetic).

The reference's scene graph is an ``anytree`` tree: ``pvtrace.Node`` derives from
``anytree.NodeMixin`` and the reference's own tests mix a *plain* ``anytree.Node``
(carrying a ``light=`` attribute) into it (reference ``tests/test_scene_creator.py:8-16``,
then ``miniplant/scene_creator.py:64`` assigns ``.parent``).  ``anytree`` is not
installed in the build image, so this module provides the small subset that path
needs: ``NodeMixin`` (``parent``/``children``/``root``/``ancestors``/``path``),
``Node(name, parent=None, children=None, **attrs)`` and the two iterators
``LevelOrderIter``/``PreOrderIter``.  Written from the anytree public API, not
from its source.
"""
from __future__ import annotations

from collections import deque
from typing import Iterator


class LoopError(RuntimeError):
    """Raised when re-parenting would create a cycle."""


class NodeMixin:
    """Adds ``parent``/``children`` bookkeeping to any class."""

    @property
    def parent(self):
        return getattr(self, "_nm_parent", None)

    @parent.setter
    def parent(self, value):
        if value is not None and not isinstance(value, NodeMixin):
            raise TypeError(f"Parent node {value!r} is not of type 'NodeMixin'.")
        old = getattr(self, "_nm_parent", None)
        if old is value:
            return
        if value is not None:
            if value is self:
                raise LoopError(f"Cannot set parent. {self!r} cannot be parent of itself.")
            if any(anc is self for anc in value.path):
                raise LoopError(f"Cannot set parent. {self!r} is parent of {value!r}.")
        if old is not None:
            old._nm_children_list().remove(self)
        self._nm_parent = value
        if value is not None:
            value._nm_children_list().append(self)

    def _nm_children_list(self) -> list:
        if not hasattr(self, "_nm_children"):
            self._nm_children = []
        return self._nm_children

    @property
    def children(self) -> tuple:
        return tuple(self._nm_children_list())

    @children.setter
    def children(self, nodes):
        for child in self.children:
            child.parent = None
        for child in nodes:
            child.parent = self

    @property
    def path(self) -> tuple:
        out = []
        node = self
        while node is not None:
            out.append(node)
            node = node.parent
        return tuple(reversed(out))

    @property
    def ancestors(self) -> tuple:
        return self.path[:-1]

    @property
    def root(self):
        node = self
        while node.parent is not None:
            node = node.parent
        return node

    @property
    def is_root(self) -> bool:
        return self.parent is None

    @property
    def is_leaf(self) -> bool:
        return len(self._nm_children_list()) == 0

    @property
    def descendants(self) -> tuple:
        return tuple(PreOrderIter(self))[1:]

    @property
    def depth(self) -> int:
        return len(self.path) - 1


class Node(NodeMixin):
    """``anytree.Node``: a named node with arbitrary keyword attributes."""

    def __init__(self, name, parent=None, children=None, **kwargs):
        self.__dict__.update(kwargs)
        self.name = name
        self.parent = parent
        if children:
            self.children = children

    def __repr__(self):
        return f"Node({'/'.join([''] + [str(n.name) for n in self.path])!r})"


class PreOrderIter:
    """Depth-first, parents before children, children in insertion order."""

    def __init__(self, node):
        self._node = node

    def __iter__(self) -> Iterator:
        stack = [self._node]
        while stack:
            node = stack.pop()
            yield node
            stack.extend(reversed(node.children))


class LevelOrderIter:
    """Breadth-first iteration."""

    def __init__(self, node):
        self._node = node

    def __iter__(self) -> Iterator:
        queue = deque([self._node])
        while queue:
            node = queue.popleft()
            yield node
            queue.extend(node.children)
