"""QuadTree / QuadNode drop-in (src/jpeg/quadtree.py:41-165).  The tree is built on the device
(csrc/quadtree.cu, aeaj_quadtree); node objects are materialised lazily on the host for callers that
walk them (test/analysis/quad_tree.py:60-86 reads .x/.y/.size of the leaves)."""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np

from aeaj.codec import get_stages
from .utils import largest_power_of_2


class QuadNode:
    def __init__(self, x: int, y: int, size: int) -> None:
        self.x = x
        self.y = y
        self.size = size
        self.children: List[Optional["QuadNode"]] = []

    def is_leaf(self) -> bool:
        return len(self.children) == 0


class QuadTree:
    def __init__(self, edge_image: np.ndarray, max_size: int = 64, min_size: int = 4) -> None:
        if not isinstance(edge_image, np.ndarray):
            raise TypeError("Input must be a numpy array.")
        if edge_image.ndim != 2:
            raise ValueError("Input array must be a 2D with a single channel.")
        self.image = edge_image
        self.max_size = max_size
        self.min_size = min_size
        self.root_size = largest_power_of_2(max(edge_image.shape)) * 2
        self._leaves, self._states, root = get_stages().quadtree(edge_image, max_size, min_size)
        assert root == self.root_size
        self._root: Optional[QuadNode] = None

    @property
    def root(self) -> QuadNode:
        """Root node; the child links are rebuilt from the state stream on first access."""
        if self._root is None:
            it = iter(self._states.tolist())

            def build(x, y, size):
                s = next(it)
                if s == 2:
                    return None
                node = QuadNode(x, y, size)
                if s == 1:
                    h = size // 2
                    node.children = [build(x, y, h), build(x + h, y, h), build(x, y + h, h), build(x + h, y + h, h)]
                return node

            import sys
            sys.setrecursionlimit(max(sys.getrecursionlimit(), 10000))
            self._root = build(0, 0, self.root_size)
        return self._root

    def get_leaves_and_states(self) -> Tuple[List[QuadNode], List[str]]:
        """Leaves in DFS order and the '00' / '01' / '10' state strings (quadtree.py:136-165)."""
        names = ("00", "01", "10")
        leaves = [QuadNode(int(x), int(y), int(s)) for x, y, s, _ in self._leaves]
        return leaves, [names[s] for s in self._states.tolist()]
