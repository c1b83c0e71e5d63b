"""LAMMPS fix-external driver (cl/lmp.py:8-71 mirror): host logic on CPU with a fake lammps object."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from autoforce_b200 import lammps_driver as ld  # noqa: E402


class FakeLammps:
    def __init__(self, cell, pos, types_, pbc=(1, 1, 1)):
        self.cell, self.pos, self.types, self.pbc = np.asarray(cell, float), np.asarray(pos, float), list(types_), list(pbc)
        self.energy = self.virial = None
        self.commands = []

    def extract_box(self):
        c = self.cell
        return [0.0, 0.0, 0.0], [c[0, 0], c[1, 1], c[2, 2]], c[0, 1], c[1, 2], c[0, 2], self.pbc, 0

    def gather_atoms(self, name, kind, count):
        return list(self.pos.reshape(-1)) if name == "x" else list(self.types)

    def fix_external_set_energy_global(self, fix_id, e):
        self.energy = (fix_id, e)

    def fix_external_set_virial_global(self, fix_id, v):
        self.virial = (fix_id, np.array(v))

    def commands_list(self, cmds):
        self.commands.append(list(cmds))

    def set_fix_external_callback(self, fix_id, cb):
        self.cb = (fix_id, cb)


class FakeCalc:
    def calculate(self, atoms, properties=()):
        self.atoms = atoms
        n = len(atoms.numbers)
        self.results = {"energy": np.array(-3.5), "forces": np.arange(3.0 * n).reshape(n, 3),
                        "stress": np.array([1.0, 2.0, 3.0, 4.0, 5.0, 6.0]) * 1e-3}
        return self.results


def test_read_lammps_file(tmp_path):
    f = tmp_path / "in.lammps"
    f.write_text("# a comment\n#autoforce atomic_numbers = {1: 29, 2: 8}\nunits metal   # eV, A\n\nboundary p p p\n"
                 "fix   autoforce all external pf/callback 1 1\nrun 10\n")
    units, mp, fix_id, idx, cmds = ld.read_lammps_file(str(f))
    assert units == "metal" and mp == {1: 29, 2: 8} and fix_id == "autoforce"
    assert cmds == ["units metal", "boundary p p p", "fix autoforce all external pf/callback 1 1", "run 10"] and idx == 2
    g = tmp_path / "bad.lammps"
    g.write_text("#autoforce atomic_numbers = {1: 29}\nunits metal\nrun 1\n")
    with pytest.raises(RuntimeError):
        ld.read_lammps_file(str(g))
    # directives are literals, never code (the reference exec()s this line, cl/lmp.py:14-16)
    h = tmp_path / "evil.lammps"
    h.write_text("#autoforce atomic_numbers = __import__('os').system('true')\nunits metal\nfix autoforce all external pf/callback 1 1\n")
    with pytest.raises(RuntimeError, match="not a Python literal"):
        ld.read_lammps_file(str(h))
    lmp = FakeLammps(np.eye(3) * 5, np.zeros((2, 3)), [1, 2])
    cb = ld.run(str(f), FakeCalc(), lmp=lmp)
    assert lmp.commands == [cmds[:3], cmds[3:]] and lmp.cb == ("autoforce", cb)


@pytest.mark.parametrize("units", ["metal", "real"])
def test_callback_units_and_virial_order(units):
    cell = np.array([[4.0, 0.5, 0.25], [0.0, 5.0, 0.125], [0.0, 0.0, 6.0]])
    pos = np.array([[0.1, 0.2, 0.3], [1.0, 2.0, 3.0], [2.0, 1.0, 0.5]])
    lmp = FakeLammps(cell, pos, [2, 1, 2])
    calc = FakeCalc()
    cb = ld.FixExternalCallback(lmp, calc, units, {1: 29, 2: 8})
    tag = np.array([3, 1, 2])
    fext = np.zeros((3, 3))
    cb(None, 0, 3, tag, None, fext)
    assert list(calc.atoms.numbers) == [8, 29, 8] and np.allclose(calc.atoms.positions, pos) and np.allclose(calc.atoms.cell, cell)
    ev = 1.0 if units == "metal" else 4184.0 / 6.022140857e23 / 1.6021766208e-19
    assert np.allclose(fext, calc.results["forces"][tag - 1] / ev, rtol=1e-12)
    assert lmp.energy[0] == "autoforce" and np.isclose(lmp.energy[1], -3.5 / ev, rtol=1e-12)
    # virial = -stress * volume in energy units (to the accuracy of LAMMPS' nktv2p constant), xy and yz swapped
    vol = 4.0 * 5.0 * 6.0
    expect = -calc.results["stress"][[0, 1, 2, 5, 4, 3]] * vol / ev
    assert np.allclose(lmp.virial[1], expect, rtol=2e-5 if units == "real" else 1e-6)
    with pytest.raises(ValueError):
        ld.FixExternalCallback(lmp, calc, "lj", {})
