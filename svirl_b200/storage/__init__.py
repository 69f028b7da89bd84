from .arrays import GArray, DeviceArray
