"""NeighborList stand-in: restated semantics of ASE's NewPrimitiveNeighborList
for the one way the reference uses it (descriptor/atoms.py:348-355,366,402)."""
import numpy as np


class NewPrimitiveNeighborList:
    pass


class PrimitiveNeighborList:
    pass


class NeighborList:
    def __init__(self, cutoffs, skin=0.3, sorted=False, self_interaction=True, bothways=False, primitive=None):
        self.cutoffs = np.asarray(cutoffs, dtype=float) + skin
        assert not self_interaction and bothways, "shim only implements the reference's usage"
        assert len(self.cutoffs) == 0 or np.all(self.cutoffs == self.cutoffs[0])

    def update(self, atoms):
        from oracle.sgpr_oracle import neighbor_list, neighbor_list_bruteforce

        rc = 2 * float(self.cutoffs[0]) if len(self.cutoffs) else 0.0
        # same result either way (tests/test_oracle_golden.py::test_neighbor_list_kdtree_equals_bruteforce); the
        # brute-force scan is quadratic, so larger cells (oracle/ref_bench.py) take the k-d tree version
        fn = neighbor_list_bruteforce if len(atoms) <= 300 else neighbor_list
        self.first, self.J, self.S = fn(atoms.positions, atoms.get_cell(complete=True), atoms.pbc, rc)
        return True

    def get_neighbors(self, a):
        sl = slice(self.first[a], self.first[a + 1])
        return self.J[sl].copy(), self.S[sl].copy()
