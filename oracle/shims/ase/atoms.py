import numpy as np


class Atoms:
    def __init__(self, symbols=None, positions=None, numbers=None, cell=None, pbc=None, calculator=None, velocities=None, **kw):
        if numbers is None:
            numbers = []
        numbers = np.array(numbers, dtype=int).reshape(-1)
        n = len(numbers)
        if positions is None:
            positions = np.zeros((n, 3))
        self.arrays = {"numbers": numbers, "positions": np.array(positions, dtype=float).reshape(-1, 3)}
        if cell is None:
            cell = np.zeros((3, 3))
        cell = np.array(cell, dtype=float)
        self._cellobj = np.diag(cell) if cell.ndim == 1 else cell.reshape(3, 3)
        if pbc is None:
            pbc = False
        self._pbc = np.array([pbc] * 3 if np.ndim(pbc) == 0 else pbc, dtype=bool)
        self._calc = calculator
        self._vel = velocities

    # -- arrays
    @property
    def positions(self):
        return self.arrays["positions"]

    @positions.setter
    def positions(self, v):
        self.arrays["positions"] = np.array(v, dtype=float).reshape(-1, 3)

    @property
    def numbers(self):
        return self.arrays["numbers"]

    @property
    def cell(self):
        return self._cellobj

    @cell.setter
    def cell(self, v):
        self._cellobj = np.array(v, dtype=float).reshape(3, 3)

    @property
    def pbc(self):
        return self._pbc

    @pbc.setter
    def pbc(self, v):
        self._pbc = np.array([v] * 3 if np.ndim(v) == 0 else v, dtype=bool)

    @property
    def calc(self):
        return self._calc

    @calc.setter
    def calc(self, c):
        self._calc = c

    def __len__(self):
        return len(self.arrays["numbers"])

    def get_global_number_of_atoms(self):
        return len(self)

    def get_atomic_numbers(self):
        return self.numbers.copy()

    def get_positions(self):
        return self.positions.copy()

    def get_volume(self):
        v = abs(np.linalg.det(self._cellobj))
        if v == 0.0:
            raise ValueError("cell has zero volume")
        return v

    def get_cell(self, complete=False):
        if complete:
            from oracle.sgpr_oracle import complete_cell

            return complete_cell(self._cellobj)
        return self._cellobj.copy()

    def get_pbc(self):
        return self._pbc.copy()

    def get_velocities(self):
        return self._vel

    def set_velocities(self, v):
        self._vel = v

    def set_cell(self, cell, scale_atoms=False):
        self.cell = cell

    def set_positions(self, p):
        self.positions = p

    def get_temperature(self):
        return 0.0

    def copy(self):
        return Atoms(positions=self.positions.copy(), numbers=self.numbers.copy(), cell=self._cellobj.copy(), pbc=self._pbc.copy())

    def _get(self, name):
        if self._calc is None:
            raise RuntimeError("Atoms object has no calculator.")
        return self._calc.get_property(name, self)

    def write(self, path, format="extxyz", append=False):
        """Minimal extended-xyz frame (io/sgprio.py:76-82 appends training data to the tape this way)."""
        assert format == "extxyz"
        sym = {1: "H", 3: "Li", 8: "O", 15: "P", 16: "S", 29: "Cu", 79: "Au"}
        res = getattr(self._calc, "results", {}) if self._calc is not None else {}
        head = 'Lattice="{}" Properties=species:S:1:pos:R:3{} pbc="{}"'.format(
            " ".join(repr(float(v)) for v in self._cellobj.reshape(-1)), ":forces:R:3" if "forces" in res else "",
            " ".join("T" if b else "F" for b in self._pbc))
        if "energy" in res:
            head += " energy={!r}".format(float(res["energy"]))
        with open(path, "a" if append else "w") as f:
            f.write(f"{len(self)}\n{head}\n")
            for k, (z, x) in enumerate(zip(self.numbers, self.positions)):
                line = "{:<2s} {:20.12f} {:20.12f} {:20.12f}".format(sym.get(int(z), f"X{int(z)}"), *x)
                if "forces" in res:
                    line += " {:20.12f} {:20.12f} {:20.12f}".format(*np.asarray(res["forces"])[k])
                f.write(line + "\n")

    def get_potential_energy(self):
        return self._get("energy")

    def get_forces(self):
        return self._get("forces")

    def get_stress(self):
        return self._get("stress")
