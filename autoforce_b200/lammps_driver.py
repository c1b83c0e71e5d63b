"""LAMMPS ``fix external pf/callback`` driver for the B200 prediction path.

Mirrors the reference's ``theforce/cl/lmp.py:8-71`` (input-script conventions, the callback's unit handling and
virial ordering) without importing ASE or the reference: the callback gathers positions from LAMMPS, calls a
calculator with the reference's ``results`` contract (``autoforce_b200.B200Calculator``) and hands forces,
energy and the global virial back to the fix.

Input script conventions (cl/lmp.py:8-33): a comment line ``#autoforce atomic_numbers = {1: 29, 2: 8}`` maps LAMMPS
types to atomic numbers, and the fix must be called ``autoforce``: ``fix autoforce all external pf/callback 1 1``.
"""
from __future__ import annotations

import ast
import re
import types

import numpy as np

# ase.units (CODATA 2014, ASE's default) as used by ase.calculators.lammps.convert in cl/lmp.py:5,46-66
_E = 1.6021766208e-19
_NA = 6.022140857e23
_KCAL_MOL = 4184.0 / _NA / _E           # eV
_BAR = 1e5 / (_E * 1e30)                # eV / A^3
_ATM = 101325.0 / (_E * 1e30)
_UNITS = {   # LAMMPS unit -> ASE (eV, A) factors
    "metal": dict(distance=1.0, energy=1.0, force=1.0, pressure=_BAR),
    "real": dict(distance=1.0, energy=_KCAL_MOL, force=_KCAL_MOL, pressure=_ATM),
}
NKTV2P = {"real": 68568.415, "metal": 1.6021765e6}   # cl/lmp.py:74-83 (the styles supported here)


def convert(value, quantity, fromunits, tounits):
    """ase.calculators.lammps.convert for the unit styles "metal", "real" and "ASE"."""
    value = np.asarray(value, dtype=float)
    if fromunits != "ASE":
        value = value * _UNITS[fromunits][quantity]
    if tounits != "ASE":
        value = value / _UNITS[tounits][quantity]
    return value


_DIRECTIVE = re.compile(r"^#\s*autoforce\b(.*)$", re.IGNORECASE)
_ASSIGN = re.compile(r"^\s*([A-Za-z_]\w*)\s*=\s*(.+?)\s*$")


class LammpsInput:
    """A LAMMPS input script as the AutoForce driver sees it (conventions of cl/lmp.py:8-33).

    * ``#autoforce name = <python literal>`` comment lines are directives; ``atomic_numbers = {type: Z, ...}`` is
      mandatory.  The value is parsed with ``ast.literal_eval`` (the reference ``exec``s the line; a literal is all
      the convention needs, and an input deck should not be able to run code).
    * everything after ``#`` is a comment; blank lines are dropped; whitespace is normalised.
    * ``units <style>`` selects the unit style; ``fix autoforce <group> external pf/callback 1 1`` is the fix the
      callback is attached to -- the commands up to and including it run first, then the callback is set.
    """

    def __init__(self, text):
        self.directives, self.commands = {}, []
        self.units, self.fix_id, self.fix_index = None, None, None
        for raw in text.splitlines():
            m = _DIRECTIVE.match(raw.strip())
            if m:
                self._directive(m.group(1))
                continue
            words = raw.split("#", 1)[0].split()
            if not words:
                continue
            if words[0] == "units" and len(words) > 1:
                self.units = words[1]
            if len(words) > 1 and words[0].lower() == "fix" and words[1].lower() == "autoforce":
                self.fix_id, self.fix_index = words[1], len(self.commands)
            self.commands.append(" ".join(words))
        if "atomic_numbers" not in self.directives:
            raise RuntimeError("no '#autoforce atomic_numbers = {...}' line!")
        if self.fix_id is None:
            raise RuntimeError("no fix autoforce!")

    def _directive(self, body):
        m = _ASSIGN.match(body)
        if not m:
            raise RuntimeError(f"cannot parse '#autoforce{body}': expected 'name = literal'")
        try:
            self.directives[m.group(1)] = ast.literal_eval(m.group(2))
        except (ValueError, SyntaxError) as ex:
            raise RuntimeError(f"'#autoforce {m.group(1)} = ...' is not a Python literal: {ex}") from None

    @property
    def map_numbers(self):
        return {int(t): int(z) for t, z in dict(self.directives["atomic_numbers"]).items()}


def read_lammps_file(file):
    """-> (units, map_numbers, fixID, fixIndex, commands), the tuple cl/lmp.py:8-33 returns."""
    with open(file) as f:
        inp = LammpsInput(f.read())
    return inp.units, inp.map_numbers, inp.fix_id, inp.fix_index, inp.commands


class FixExternalCallback:
    """The ``callback(caller, ntimestep, nlocal, tag, pos, fext)`` of cl/lmp.py:42-71 as an object.

    ``lmp``: a ``lammps.lammps`` instance (extract_box, gather_atoms, fix_external_set_energy_global,
    fix_external_set_virial_global); ``calc``: calculator with ``calculate(atoms)`` and ``results``."""

    def __init__(self, lmp, calc, units, map_numbers, fix_id="autoforce"):
        if units not in _UNITS:
            raise ValueError(f"LAMMPS units {units!r} not supported (metal, real)")
        self.lmp, self.calc, self.units, self.map_numbers, self.fix_id = lmp, calc, units, dict(map_numbers), fix_id
        self.numbers = None
        self.steps = 0

    def get_cell(self):
        boxlo, (xhi, yhi, zhi), xy, yz, xz, pbc, box_change = self.lmp.extract_box()
        # cl/lmp.py:36-40 (box taken with its lower corner at the origin, as the reference does)
        cell = np.array([[xhi, xy, xz], [0.0, yhi, yz], [0.0, 0.0, zhi]])
        return cell, pbc

    def __call__(self, caller, ntimestep, nlocal, tag, pos, fext):
        cell, pbc = self.get_cell()
        cell = convert(cell, "distance", self.units, "ASE")
        xyz = np.array(self.lmp.gather_atoms("x", 1, 3), dtype=float).reshape(-1, 3)
        positions = convert(xyz, "distance", self.units, "ASE")
        if self.numbers is None:
            types_ = np.array(self.lmp.gather_atoms("type", 0, 1))
            self.numbers = np.array([self.map_numbers[int(t)] for t in types_], dtype=np.int64)
        atoms = types.SimpleNamespace(positions=positions, cell=cell, pbc=np.array([bool(p) for p in pbc]), numbers=self.numbers)
        self.calc.calculate(atoms, properties=("energy", "forces", "stress"))
        res = self.calc.results
        f = np.asarray(res["forces"])[np.asarray(tag) - 1]
        fext[:] = convert(f, "force", "ASE", self.units)
        self.lmp.fix_external_set_energy_global(self.fix_id, float(convert(res["energy"], "energy", "ASE", self.units)))
        if "stress" in res:
            v = convert(res["stress"], "pressure", "ASE", self.units)
            vol = abs(np.linalg.det(cell))
            v = -v / (NKTV2P[self.units] / vol)
            v[3:] = v[3:][::-1]          # ASE Voigt (xx,yy,zz,yz,xz,xy) -> LAMMPS (xx,yy,zz,xy,xz,yz)
            self.lmp.fix_external_set_virial_global(self.fix_id, v)
        self.steps += 1


def run(input_file, calc, lmp=None):
    """``python -m theforce.cl.lmp -i in.lammps`` (cl/lmp.py:86-118) with a given calculator."""
    units, map_numbers, fix_id, fix_index, commands = read_lammps_file(input_file)
    if lmp is None:
        from lammps import lammps   # not a dependency of this package

        lmp = lammps()
    cb = FixExternalCallback(lmp, calc, units, map_numbers, fix_id)
    lmp.commands_list(commands[: fix_index + 1])
    lmp.set_fix_external_callback(fix_id, cb)
    lmp.commands_list(commands[fix_index + 1:])
    return cb
