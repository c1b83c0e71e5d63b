def read(*args, **kwargs):
    raise NotImplementedError("ase.io shim: read() is not available")


class Trajectory:
    def __init__(self, name, mode="r"):
        self.name, self.mode, self.frames = name, mode, []

    def write(self, atoms):
        self.frames.append(atoms)

    def close(self):
        pass
