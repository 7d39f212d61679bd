"""Stub of the tiny part of tensorflow the reference's sampling.py touches
(tf.io.gfile.*), mapped onto the local filesystem. Test infrastructure only."""
import glob as _glob
import os as _os


class _GFile:
    def __init__(self, path, mode="r"):
        self._f = open(path, mode)

    def __enter__(self):
        return self._f

    def __exit__(self, *a):
        self._f.close()
        return False


class _gfile:
    GFile = _GFile

    @staticmethod
    def makedirs(p):
        _os.makedirs(p, exist_ok=True)

    @staticmethod
    def exists(p):
        return _os.path.exists(p)

    @staticmethod
    def glob(p):
        return _glob.glob(p)

    @staticmethod
    def isdir(p):
        return _os.path.isdir(p)


class _io:
    gfile = _gfile


io = _io
