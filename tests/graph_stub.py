"""Minimal stand-in for ``igraph.Graph`` (igraph is not installed in this image).

Exposes exactly the API ``BNLearnWrapper.score`` touches in the reference
(``src/problem/bn/bnlearn.py:29-42``): ``vcount()``, ``vs()[key]``, iteration over ``vs``
yielding vertices with ``.index`` and ``[key]``, ``get_edgelist()`` — plus the constructor the
reference's ``LabeledDag.from_dict_to_graph`` (``src/toolkit/labeled.py:132-154``) performs.
"""
from typing import Dict, List, Tuple


class Vertex:
    def __init__(self, index: int, attrs: Dict):
        self.index = index
        self._attrs = attrs

    def __getitem__(self, key):
        return self._attrs[key]


class VertexSeq:
    def __init__(self, vertices: List[Vertex]):
        self._v = vertices

    def __call__(self):
        return self

    def __iter__(self):
        return iter(self._v)

    def __len__(self):
        return len(self._v)

    def __getitem__(self, key):
        if isinstance(key, str):
            return [v[key] for v in self._v]
        return self._v[key]


class Graph:
    def __init__(self, n: int, edges: List[Tuple[int, int]], labels: List[int], label_key: str = "type"):
        self._vs = VertexSeq([Vertex(i, {label_key: labels[i]}) for i in range(n)])
        self._edges = list(edges)

    def vcount(self) -> int:
        return len(self._vs)

    @property
    def vs(self) -> VertexSeq:
        return self._vs

    def get_edgelist(self) -> List[Tuple[int, int]]:
        return list(self._edges)


def from_dict_to_graph(d: Dict, n: int, label_key: str = "type") -> Graph:
    """What ``LabeledDag(n, n).from_dict_to_graph`` builds (``labeled.py:132-154``)."""
    edges = [(u, i) for i in range(n) for u in range(i) if int(d[f"e{i}"][u]) == 1]
    return Graph(n, edges, [int(d[f"l{i}"]) for i in range(n)], label_key)
