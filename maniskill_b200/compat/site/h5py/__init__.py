"""Import-level stand-in: trajectory recording to HDF5 needs the real h5py (the recorder of this repo writes .npz instead)."""


class _Missing:
    def __init__(self, *a, **kw):
        raise ImportError("h5py is not installed; this is an import-level stand-in")


File = Group = Dataset = _Missing
