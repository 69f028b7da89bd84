"""Small helpers under the reference's names (svirl/parallel/utils.py:7-27)."""
import numpy as np


class Utils(object):

    @staticmethod
    def abs2(c):
        """|c|^2 without the square root."""
        c = np.asarray(c)
        return c.real * c.real + c.imag * c.imag

    @staticmethod
    def intceil(k, l):
        """ceil(k / l) for the launch-geometry integers the callers pass."""
        return int(-(-int(k) // int(l))) if float(k).is_integer() and float(l).is_integer() else int(np.ceil(float(k) / float(l)))

    @staticmethod
    def copy_dtod(dest, src):
        """Device-to-device copy; either side may be a DeviceArray or a GArray."""
        if dest is None or src is None:
            print('Warning! src/dest pointer is null')
            return
        unwrap = lambda x: x.get_d_obj() if hasattr(x, 'get_d_obj') else x
        unwrap(dest).copy_from(unwrap(src))
