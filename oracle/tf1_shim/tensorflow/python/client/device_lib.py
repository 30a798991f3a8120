"""Part of the TF1 stand-in (see ../../__init__.py): the reference's models pick ``list_local_devices()`` entries by type."""
import collections

_Device = collections.namedtuple('_Device', 'name device_type')


def list_local_devices():
    return [_Device('/device:CPU:0', 'CPU')]
