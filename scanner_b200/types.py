"""Element types of byte columns: how a Python value becomes the bytes of a table row and back.

The surface of the reference's `scannerpy.types` (python/scannerpy/types.py:1-160): a registry
keyed by the Python type used in a kernel's annotations, each entry with the name the column is
stored under (`cpp_name`, the `type_name` of the reference's ColumnDescriptor) and a serialize /
deserialize pair.  Only what the decode -> evaluate -> save path and its tests use is predefined:
bytes, Any (pickle), FrameType, the numpy array types and Histogram; `register_type`,
`VariableList` and `UniformList` build the rest the way the reference does.
"""
import pickle
import struct
from typing import Any

import numpy as np

PYTHON_TYPE_REGISTRY = {}


class ScannerTypeError(Exception):
    pass


class FrameType:
    """Annotation of a frame-valued input or output: the kernel sees an ndarray (H, W, C)."""


BlobType = bytes


class ScannerTypeInfo:
    __slots__ = ("type", "cpp_name", "serialize", "deserialize")

    def __init__(self, type, cpp_name, serialize, deserialize):
        self.type, self.cpp_name, self.serialize, self.deserialize = type, cpp_name, serialize, deserialize


def _register_type(ty, cpp_name, serialize, deserialize):
    PYTHON_TYPE_REGISTRY[ty] = ScannerTypeInfo(ty, cpp_name, serialize, deserialize)


def get_type_info(ty):
    if ty not in PYTHON_TYPE_REGISTRY:
        raise ScannerTypeError("Type `{}` has not been registered with Scanner".format(getattr(ty, "__name__", ty)))
    return PYTHON_TYPE_REGISTRY[ty]


def get_type_info_cpp(cpp_name):
    for info in PYTHON_TYPE_REGISTRY.values():
        if info.cpp_name == cpp_name:
            return info
    raise ScannerTypeError("Type `{}` has not been registered with Scanner".format(cpp_name))


_register_type(bytes, "Bytes", lambda x: x, lambda x: x)
_register_type(Any, "Any", pickle.dumps, pickle.loads)
_register_type(FrameType, "FrameType", lambda x: x, lambda x: x)


def register_type(cls):
    """Class decorator: `cls.serialize(obj) -> bytes`, `cls.deserialize(bytes) -> obj`."""
    _register_type(cls, cls.__name__, cls.serialize, cls.deserialize)
    return cls


def ProtobufType(name, proto):
    """A column type whose elements are messages of the protobuf class `proto` (types.py:57-67):
    anything with SerializeToString / ParseFromString, e.g. a google.protobuf generated class."""

    def serialize(message):
        return message.SerializeToString()

    def deserialize(buf):
        message = proto()
        message.ParseFromString(bytes(buf))
        return message

    return register_type(type(name, (), dict(serialize=staticmethod(serialize), deserialize=staticmethod(deserialize))))


def VariableList(name, typ):
    """u64 count, then per element u64 size + bytes (types.py:70-93)."""

    def serialize(items):
        out = [struct.pack("=Q", len(items))]
        for item in items:
            blob = typ.serialize(item)
            out.append(struct.pack("=Q", len(blob)))
            out.append(blob)
        return b"".join(out)

    def deserialize(buf):
        (n,) = struct.unpack_from("=Q", buf, 0)
        off, items = 8, []
        for _ in range(n):
            (size,) = struct.unpack_from("=Q", buf, off)
            off += 8
            items.append(typ.deserialize(buf[off:off + size]))
            off += size
        return items

    return register_type(type(name, (), dict(serialize=staticmethod(serialize), deserialize=staticmethod(deserialize))))


def UniformList(name, typ, size=None, parts=None):
    """Equal-sized elements back to back: `size` bytes each, or `parts` of them (types.py:95-116)."""
    assert (size is not None) ^ (parts is not None)

    def serialize(items):
        return b"".join(typ.serialize(item) for item in items)

    def deserialize(buf):
        if len(buf) <= 4:
            return []
        step = size if parts is None else len(buf) // parts
        assert len(buf) % step == 0
        return [typ.deserialize(buf[i:i + step]) for i in range(0, len(buf), step)]

    return register_type(type(name, (), dict(serialize=staticmethod(serialize), deserialize=staticmethod(deserialize))))


@register_type
class NumpyArrayFloat32:
    @staticmethod
    def serialize(array):
        return np.ascontiguousarray(array, dtype=np.float32).tobytes()

    @staticmethod
    def deserialize(buf):
        return np.frombuffer(buf, dtype=np.float32)


@register_type
class NumpyArrayInt32:
    @staticmethod
    def serialize(array):
        return np.ascontiguousarray(array, dtype=np.int32).tobytes()

    @staticmethod
    def deserialize(buf):
        return np.frombuffer(buf, dtype=np.int32)


# the Histogram op's output: three int32[16] arrays, channel-major (types.py:125-132)
Histogram = UniformList("Histogram", NumpyArrayInt32, parts=3)


@register_type
class Image:
    """PNG bytes <-> ndarray (the ImageEncoder op's column type; needs cv2 only when used)."""

    @staticmethod
    def serialize(image):
        import cv2
        ok, blob = cv2.imencode(".png", image)
        if not ok:
            raise ScannerTypeError("could not encode the image as PNG")
        return blob.tobytes()

    @staticmethod
    def deserialize(blob):
        import cv2
        return cv2.imdecode(np.frombuffer(blob, dtype=np.uint8), cv2.IMREAD_UNCHANGED)
