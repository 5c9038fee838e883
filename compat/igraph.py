"""Minimal igraph stand-in: the subset of `igraph.Graph` the reference's learning package uses
(learning/spg.py:104-128,134-176; learning/ecc/GraphConvInfo.py:48-58).  Directed multigraph on numpy
arrays; vertex/edge attributes are per-element Python lists (as igraph returns them)."""
import numpy as np


class _Seq(object):
    def __init__(self, attrs, n, index=None):
        self._attrs, self._n, self._index = attrs, n, index

    def attributes(self):
        return list(self._attrs.keys())

    def __len__(self):
        return self._n if self._index is None else len(self._index)

    def __getitem__(self, key):
        if isinstance(key, str):  # whole column, e.g. G.vs['s']
            col = self._attrs[key]
            return list(col) if self._index is None else [col[i] for i in self._index]
        if isinstance(key, (int, np.integer)):  # one element, e.g. G.vs[3]['v']
            i = int(key) if self._index is None else self._index[int(key)]
            return {a: col[i] for a, col in self._attrs.items()}
        idx = list(key)  # sub-sequence, e.g. G.es[[2, 0, 1]]
        if self._index is not None:
            idx = [self._index[i] for i in idx]
        return _Seq(self._attrs, self._n, idx)

    def get_attribute_values(self, name):
        return self[name]

    def __iter__(self):
        return iter(range(len(self)))


class Graph(object):
    def __init__(self, n=0, edges=None, directed=False, edge_attrs=None, vertex_attrs=None):
        self._n = int(n)
        self._edges = np.asarray(edges if edges is not None else [], dtype=np.int64).reshape(-1, 2)
        self._directed = directed
        self._vattrs = {k: list(v) for k, v in (vertex_attrs or {}).items()}
        self._eattrs = {k: list(v) for k, v in (edge_attrs or {}).items()}

    def vcount(self):
        return self._n

    def ecount(self):
        return int(self._edges.shape[0])

    def get_edgelist(self):
        return [tuple(e) for e in self._edges.tolist()]

    @property
    def vs(self):
        return _Seq(self._vattrs, self._n)

    @property
    def es(self):
        return _Seq(self._eattrs, self.ecount())

    def indegree(self, vertices=None, loops=True):
        deg = np.bincount(self._edges[:, 1], minlength=self._n) if self.ecount() else np.zeros(self._n, np.int64)
        return deg.tolist()

    def _induced(self, keep):
        """Induced subgraph on the sorted vertex ids `keep` (igraph renumbers them 0..k-1 in order)."""
        keep = np.asarray(sorted(set(int(v) for v in keep)), dtype=np.int64)
        remap = -np.ones(self._n, dtype=np.int64)
        remap[keep] = np.arange(keep.shape[0])
        if self.ecount():
            sel = (remap[self._edges[:, 0]] >= 0) & (remap[self._edges[:, 1]] >= 0)
            eidx = np.nonzero(sel)[0]
            edges = remap[self._edges[eidx]]
        else:
            eidx, edges = np.zeros(0, np.int64), np.zeros((0, 2), np.int64)
        g = Graph(n=keep.shape[0], edges=edges, directed=self._directed)
        g._vattrs = {k: [col[i] for i in keep] for k, col in self._vattrs.items()}
        g._eattrs = {k: [col[i] for i in eidx] for k, col in self._eattrs.items()}
        return g

    def subgraph(self, vertices):
        return self._induced(list(vertices))

    def permute_vertices(self, perm):
        """Vertex i of this graph becomes vertex perm[i] of the result (igraph semantics)."""
        perm = np.asarray(perm, dtype=np.int64)
        inv = np.empty_like(perm)
        inv[perm] = np.arange(perm.shape[0])
        g = Graph(n=self._n, edges=perm[self._edges] if self.ecount() else self._edges, directed=self._directed)
        g._vattrs = {k: [col[i] for i in inv] for k, col in self._vattrs.items()}
        g._eattrs = {k: list(col) for k, col in self._eattrs.items()}
        return g

    def neighborhood(self, vertices, order=1, mode="all"):
        """For every start vertex: the vertices within `order` hops, ignoring edge directions, start first."""
        nbrs = [[] for _ in range(self._n)]
        for a, b in self._edges.tolist():
            nbrs[a].append(b)
            nbrs[b].append(a)
        out = []
        for v in vertices:
            seen, frontier, res = {int(v)}, [int(v)], [int(v)]
            for _ in range(order):
                nxt = []
                for u in frontier:
                    for x in nbrs[u]:
                        if x not in seen:
                            seen.add(x)
                            nxt.append(x)
                            res.append(x)
                frontier = nxt
            out.append(res)
        return out
