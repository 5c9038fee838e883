"""Minimal h5py stand-in over `.npz` archives (numpy's zip container), enough for the reference's
readers and writers on the learning path (learning/spg.py:66-103,198-205; learning/main.py:379-381;
learning/s3dis_dataset.py:70).  Dataset names with `/` act as groups."""
import os

import numpy as np


class _Dataset(object):
    def __init__(self, arr):
        self._a = arr

    def __getitem__(self, key):
        return self._a[key]

    @property
    def shape(self):
        return self._a.shape

    @property
    def size(self):
        return self._a.size

    @property
    def dtype(self):
        return self._a.dtype

    @property
    def value(self):
        return self._a

    def __len__(self):
        return len(self._a)

    def __array__(self, dtype=None):
        return np.asarray(self._a, dtype=dtype)


class _Group(object):
    def __init__(self, store, prefix):
        self._store, self._prefix = store, prefix

    def keys(self):
        n = len(self._prefix)
        return sorted({k[n:].split("/")[0] for k in self._store if k.startswith(self._prefix)})

    def __getitem__(self, name):
        full = self._prefix + name
        if full in self._store:
            return _Dataset(self._store[full])
        if any(k.startswith(full + "/") for k in self._store):
            return _Group(self._store, full + "/")
        raise KeyError(name)

    def __contains__(self, name):
        full = self._prefix + name
        return full in self._store or any(k.startswith(full + "/") for k in self._store)


_CACHE = {}


class File(_Group):
    def __init__(self, name, mode="r"):
        self._name, self._mode = name, mode
        if mode == "r":
            key = (os.path.abspath(name), os.path.getmtime(name))
            if key not in _CACHE:
                with np.load(name, allow_pickle=False) as z:
                    _CACHE[key] = {k: z[k] for k in z.files}
            store = _CACHE[key]
        else:
            store = {}
        _Group.__init__(self, store, "")

    def create_dataset(self, name, data=None, **kwargs):
        self._store[name] = np.asarray(data)
        return _Dataset(self._store[name])

    def close(self):
        if self._mode != "r":
            with open(self._name, "wb") as f:
                np.savez(f, **self._store)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
