from .calculator import Calculator


class SinglePointCalculator(Calculator):
    def __init__(self, atoms, **results):
        super().__init__()
        self.atoms = atoms.copy() if hasattr(atoms, "copy") else atoms
        self.results = dict(results)

    def get_property(self, name, atoms=None):
        return self.results[name]
