class Filter:
    def __init__(self, atoms, indices=None, mask=None):
        self.atoms = atoms
        self.indices = indices
