"""Small helpers with the reference's names (svirl/parallel/utils.py:7-27)."""
import numpy as np


class Utils(object):

    @staticmethod
    def abs2(c):
        return np.square(c.real) + np.square(c.imag)

    @staticmethod
    def intceil(k, l):
        return int(np.ceil(float(k) / float(l)))

    @staticmethod
    def copy_dtod(dest, src):
        """Device-to-device copy; accepts DeviceArray or GArray on either side."""
        if dest is None or src is None:
            print('Warning! src/dest pointer is null')
            return
        d = dest.get_d_obj() if hasattr(dest, 'get_d_obj') else dest
        s = src.get_d_obj() if hasattr(src, 'get_d_obj') else src
        d.copy_from(s)
