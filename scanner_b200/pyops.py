"""Ops whose kernel is Python code: the reference's `@scannerpy.register_python_op`
(python/scannerpy/op.py:317-620) and `scannerpy.Kernel` (kernel.py:17-81).

    @register_python_op()
    def Brightness(config, frame: FrameType) -> bytes: ...

    @register_python_op(batch=8)
    class Scale(Kernel):
        def __init__(self, config, factor=2): ...
        def new_stream(self, offset=0): ...
        def execute(self, frame: Sequence[FrameType]) -> Sequence[FrameType]: ...

The annotations of `execute` define the op's columns exactly as in the reference: FrameType is a
frame column, anything else a byte column whose elements go through the type's serialize /
deserialize pair (scanner_b200/types.py); batched ops take and return `Sequence[T]`, stencilled ops
take `Sequence[T]` (both: `Sequence[Sequence[T]]`); a tuple return type makes several output columns
(`ret0`, `ret1`, ...); `*args` declares variadic inputs.  Arguments of `__init__` after `config` are
the op's init arguments, arguments of `new_stream` are per-stream arguments (`sc.ops.Scale(frame=f,
factor=3, offset=[1, 2])`).

The reference pickles the kernel to worker processes and calls it through an embedded interpreter
(scanner/engine/python_kernel.cpp); here the engine runs inside this interpreter, so registration
hands the engine ONE C callback per op (include/scn_engine.h scn_register_callback_op) and the
kernel object lives in this process.  One kernel object exists per pipeline instance, called only
by that instance's evaluate thread; the GIL serialises the Python part across instances.
An exception in a kernel fails the run with the traceback in the error message.
"""
import collections.abc
import ctypes
import pickle
import traceback
import types as pytypes
from collections import OrderedDict
from inspect import signature
from itertools import islice
from typing import Sequence, Tuple

import numpy as np

from . import engine as E
from . import types as T
from .types import FrameType

PYTHON_OP_REGISTRY = {}

_NP_OF_FRAME_TYPE = {0: np.uint8, 1: np.float32, 2: np.float64, 3: np.uint16}
_FRAME_TYPE_OF_NP = {np.dtype(v): k for k, v in _NP_OF_FRAME_TYPE.items()}

EV_CONSTRUCT, EV_DESTROY, EV_NEW_STREAM, EV_RESET, EV_EXECUTE, EV_FETCH, EV_SETUP = range(7)


class PythonOpError(Exception):
    pass


class KernelConfig:
    """What a kernel's constructor receives (kernel.py:5-14)."""

    def __init__(self, devices, input_columns, input_column_types, output_columns, output_column_types, args,
                 node_id):
        self.devices = devices
        self.input_columns = input_columns
        self.input_column_types = input_column_types
        self.output_columns = output_columns
        self.output_column_types = output_column_types
        self.args = args
        self.node_id = node_id


class Kernel:
    """Base class of class kernels (kernel.py:17-81)."""

    def __init__(self, config):
        self.config = config

    def close(self):
        pass

    def new_stream(self):
        pass

    def reset(self):
        pass

    def fetch_resources(self):
        pass

    def setup_with_resources(self):
        pass

    def execute(self):
        raise NotImplementedError


# ---- ctypes mirror of include/scn_engine.h -------------------------------------------------------
class _Elem(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("size", ctypes.c_uint64), ("shape", ctypes.c_int32 * 3),
                ("frame_type", ctypes.c_int32), ("index", ctypes.c_int64)]


class _Call(ctypes.Structure):
    _fields_ = [("event", ctypes.c_int32), ("node_id", ctypes.c_int32), ("instance", ctypes.c_int64),
                ("device_type", ctypes.c_int32), ("device_id", ctypes.c_int32), ("args", ctypes.c_void_p),
                ("args_size", ctypes.c_uint64), ("n_cols", ctypes.c_int32), ("n_rows", ctypes.c_int32),
                ("n_stencil", ctypes.c_int32), ("reserved", ctypes.c_int32), ("elems", ctypes.POINTER(_Elem)),
                ("out", ctypes.c_void_p)]


class _Desc(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char_p), ("n_inputs", ctypes.c_int32),
                ("input_names", ctypes.POINTER(ctypes.c_char_p)), ("input_is_frame", ctypes.POINTER(ctypes.c_int32)),
                ("variadic_inputs", ctypes.c_int32), ("n_outputs", ctypes.c_int32),
                ("output_names", ctypes.POINTER(ctypes.c_char_p)),
                ("output_is_frame", ctypes.POINTER(ctypes.c_int32)),
                ("output_type_names", ctypes.POINTER(ctypes.c_char_p)), ("stencil", ctypes.POINTER(ctypes.c_int32)),
                ("n_stencil", ctypes.c_int32), ("bounded_state", ctypes.c_int32), ("unbounded_state", ctypes.c_int32),
                ("batch", ctypes.c_int32), ("also_gpu", ctypes.c_int32)]


_CALLBACK = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(_Call), ctypes.c_void_p, ctypes.c_size_t)
_bound = False


def _lib():
    global _bound
    lib = E.lib()
    if not _bound:
        lib.scn_register_callback_op.restype = ctypes.c_int
        lib.scn_register_callback_op.argtypes = [ctypes.POINTER(_Desc), _CALLBACK, ctypes.c_void_p]
        lib.scn_cb_emit_bytes.restype = ctypes.c_int
        lib.scn_cb_emit_bytes.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t]
        lib.scn_cb_emit_frame.restype = ctypes.c_int
        lib.scn_cb_emit_frame.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p] + [ctypes.c_int] * 4
        _bound = True
    return lib


# ---- annotation parsing (op.py:389-520) ----------------------------------------------------------
def _is_sequence(typ):
    origin = getattr(typ, "__origin__", None)
    return origin is not None and origin in (Sequence, collections.abc.Sequence, list)


def _parse_tuple(typ):
    origin = getattr(typ, "__origin__", None)
    if origin in (Tuple, tuple):
        args = getattr(typ, "__args__", ()) or ()
        ellipsis = bool(args) and args[-1] is Ellipsis
        return True, ellipsis, list(args[:-1] if ellipsis else args)
    return False, False, [typ]


class _Column:
    __slots__ = ("name", "is_frame", "info")

    def __init__(self, name, is_frame, info):
        self.name, self.is_frame, self.info = name, is_frame, info


class _PythonOp:
    """Everything known about one registered op + its live kernel objects."""

    def __init__(self, name):
        self.name = name
        self.inputs, self.outputs = [], []
        self.variadic = False
        self.can_batch = self.can_stencil = False
        self.return_is_tuple = False
        self.is_fn = False
        self.target = None
        self.kernel_params, self.stream_params = [], []
        self.kernels = {}
        self.callback = None  # keeps the ctypes thunk alive

    # -- element conversion -------------------------------------------------------------------
    def _load(self, e, col):
        if not e.data:
            return None
        if e.frame_type >= 0:
            dt = np.dtype(_NP_OF_FRAME_TYPE[e.frame_type])
            n = int(e.shape[0]) * int(e.shape[1]) * int(e.shape[2])
            view = np.ctypeslib.as_array(ctypes.cast(e.data, ctypes.POINTER(ctypes.c_uint8)), shape=(int(e.size),))
            # the engine's buffer is only valid during the call: the kernel gets its own copy
            return view[:n * dt.itemsize].view(dt).reshape(int(e.shape[0]), int(e.shape[1]), int(e.shape[2])).copy()
        return col.info.deserialize(ctypes.string_at(e.data, int(e.size)))

    def _column_value(self, call, c, col):
        rows, sten = call.n_rows, call.n_stencil
        base = c * rows * sten

        def window(r):
            return [self._load(call.elems[base + r * sten + s], col) for s in range(sten)]

        if self.can_batch and self.can_stencil:
            return [window(r) for r in range(rows)]
        if self.can_batch:
            return [self._load(call.elems[base + r * sten], col) for r in range(rows)]
        if self.can_stencil:
            return window(0)
        return self._load(call.elems[base], col)

    def _emit(self, lib, out, c, col, value):
        if value is None:
            rc = lib.scn_cb_emit_bytes(out, c, None, 0)
        elif col.is_frame:
            a = np.ascontiguousarray(value)
            if a.ndim == 2:
                a = a[:, :, None]
            if a.ndim != 3 or a.dtype not in _FRAME_TYPE_OF_NP:
                raise PythonOpError(f"Op {self.name}: output column {col.name} must be an (H, W[, C]) array of "
                                    f"uint8, uint16, float32 or float64, got {a.dtype} with shape {a.shape}")
            rc = lib.scn_cb_emit_frame(out, c, a.ctypes.data, a.shape[0], a.shape[1], a.shape[2],
                                       _FRAME_TYPE_OF_NP[a.dtype])
        else:
            blob = col.info.serialize(value)
            if not isinstance(blob, (bytes, bytearray, memoryview)):
                raise PythonOpError(f"Op {self.name}: the serializer of output column {col.name} returned "
                                    f"{type(blob).__name__}, not bytes")
            blob = bytes(blob)
            rc = lib.scn_cb_emit_bytes(out, c, blob, len(blob))
        if rc != 0:
            raise PythonOpError(f"Op {self.name}: the engine refused an element of output column {col.name}")

    # -- kernel events ------------------------------------------------------------------------
    def handle(self, call):
        ev = call.event
        if ev == EV_CONSTRUCT:
            blob = ctypes.string_at(call.args, int(call.args_size)) if call.args_size else b""
            args = pickle.loads(blob) if blob else {}
            from .client import DeviceHandle, DeviceType  # client imports this module
            config = KernelConfig([DeviceHandle(DeviceType(call.device_type), call.device_id)],
                                  [c.name for c in self.inputs],
                                  ["Video" if c.is_frame else "Bytes" for c in self.inputs],
                                  [c.name for c in self.outputs],
                                  ["Video" if c.is_frame else "Bytes" for c in self.outputs], args, call.node_id)
            self.kernels[call.instance] = config if self.is_fn else self.target(config, **(args or {}))
            return
        kernel = self.kernels.get(call.instance)
        if kernel is None:
            raise PythonOpError(f"Op {self.name}: unknown kernel instance {call.instance}")
        if ev == EV_DESTROY:
            del self.kernels[call.instance]
            if not self.is_fn:
                kernel.close()
        elif ev == EV_NEW_STREAM:
            if not self.is_fn:
                blob = ctypes.string_at(call.args, int(call.args_size)) if call.args_size else b""
                kernel.new_stream(**(pickle.loads(blob) if blob else {}))
        elif ev == EV_RESET:
            if not self.is_fn:
                kernel.reset()
        elif ev == EV_FETCH:
            if not self.is_fn:
                kernel.fetch_resources()
        elif ev == EV_SETUP:
            if not self.is_fn:
                kernel.setup_with_resources()
        elif ev == EV_EXECUTE:
            self._execute(kernel, call)

    def _execute(self, kernel, call):
        fn = self.target if self.is_fn else self.exec_fn  # fn(config, ...) / Class.execute(self, ...)
        if self.variadic:
            col = self.inputs[0]
            result = fn(kernel, *[self._column_value(call, c, col) for c in range(call.n_cols)])
        else:
            if call.n_cols != len(self.inputs):
                raise PythonOpError(f"Op {self.name}: got {call.n_cols} input columns, declared {len(self.inputs)}")
            result = fn(kernel, **{col.name: self._column_value(call, c, col) for c, col in enumerate(self.inputs)})
        columns = result if self.return_is_tuple else (result,)
        if not isinstance(columns, (tuple, list)) or len(columns) != len(self.outputs):
            raise PythonOpError(f"Op {self.name}: execute must return {len(self.outputs)} output column(s)")
        lib = _lib()
        for c, (col, value) in enumerate(zip(self.outputs, columns)):
            if self.can_batch:
                if not isinstance(value, (list, tuple)) or len(value) != call.n_rows:
                    got = len(value) if isinstance(value, (list, tuple)) else type(value).__name__
                    raise PythonOpError(f"Op {self.name}: output column {col.name} must be a sequence of "
                                        f"{call.n_rows} elements (one per input row), got {got}")
                for item in value:
                    self._emit(lib, call.out, c, col, item)
            else:
                self._emit(lib, call.out, c, col, value)


def _make_callback(op):
    def thunk(_user, call_p, err, err_cap):
        try:
            op.handle(call_p.contents)
            return 0
        except BaseException:  # nothing may propagate into the engine's thread
            msg = f"Python kernel of op {op.name} raised:\n{traceback.format_exc()}".encode("utf-8", "replace")
            if err and err_cap:
                msg = msg[-(err_cap - 1):]  # keep the end: the exception line
                ctypes.memmove(err, msg + b"\0", len(msg) + 1)
            return 1
    return _CALLBACK(thunk)


def register_python_op(name=None, stencil=None, unbounded_state=False, bounded_state=None, device_type=None,
                       device_sets=None, batch=1, proto_path=None):
    """Class or function decorator registering an Op and its Python kernel (op.py:317-620)."""

    def dec(fn_or_class):
        is_fn = isinstance(fn_or_class, (pytypes.FunctionType, pytypes.BuiltinFunctionType))
        kname = name if name is not None else fn_or_class.__name__
        can_stencil, can_batch = stencil is not None, batch > 1
        exec_fn = fn_or_class if is_fn else getattr(fn_or_class, "execute", None)
        if not callable(exec_fn):
            raise PythonOpError(f'Attempted to register Python Op with name {kname}, but that provided class has '
                                f'no "execute" method.')
        if kname in PYTHON_OP_REGISTRY:
            raise PythonOpError(f"Attempted to register Op with name {kname} twice")
        if unbounded_state and bounded_state is not None:
            raise PythonOpError("unbounded_state and bounded_state are mutually exclusive")
        if device_type is not None and device_sets is not None:
            raise PythonOpError('Must only specify one of "device_type" or "device_sets" for python Op.')

        op = _PythonOp(kname)
        op.is_fn, op.target, op.exec_fn = is_fn, fn_or_class, exec_fn
        op.can_batch, op.can_stencil = can_batch, can_stencil

        def column_type(typ, is_input):
            if can_batch:
                if not _is_sequence(typ):
                    raise PythonOpError('A batched Op must specify a "Sequence" type annotation for each input '
                                        'and output.')
                typ = typ.__args__[0]
            if is_input and can_stencil:
                if not _is_sequence(typ):
                    raise PythonOpError('A stenciled Op must specify a "Sequence" type annotation for each input. '
                                        'If the Op both stencils and batches, then it should have the type '
                                        '"Sequence[Sequence[T]], where T = {bytes, FrameType}.')
                typ = typ.__args__[0]
            return typ is FrameType, T.get_type_info(typ)

        sig = signature(exec_fn)
        params = OrderedDict(islice(sig.parameters.items(), 1, None))  # skip config / self
        for pname, param in params.items():
            if param.kind in (param.POSITIONAL_ONLY, param.VAR_KEYWORD):
                raise PythonOpError('Positional arguments and **kwargs are currently not supported for the '
                                    '"execute" method of kernels')
            if param.kind == param.VAR_POSITIONAL:
                if len(params) > 1:
                    raise PythonOpError("Variadic positional inputs (*args) are not supported when used with "
                                        "other inputs.")
                is_tuple, ellipsis, inner = _parse_tuple(param.annotation)
                if is_tuple and not ellipsis:
                    raise PythonOpError('Variadic positional inputs (*args) must be annotated as '
                                        '"args: Tuple[Type, ...]" or "args: Type"')
                op.variadic = True
                op.inputs.append(_Column(pname, *column_type(inner[0], True)))
                break
            if param.annotation is param.empty:
                raise PythonOpError(f'No type annotation specified for input {pname}. Must specify an annotation '
                                    f'of "bytes" or "FrameType".')
            op.inputs.append(_Column(pname, *column_type(param.annotation, True)))
        if sig.return_annotation is sig.empty:
            raise PythonOpError('Return annotation must be specified for "execute" method.')
        op.return_is_tuple, ellipsis, rets = _parse_tuple(sig.return_annotation)
        if ellipsis:
            raise PythonOpError("Ellipsis tuples not supported for return type.")
        for i, typ in enumerate(rets):
            op.outputs.append(_Column(f"ret{i}", *column_type(typ, False)))

        if not is_fn:
            init_params = list(signature(fn_or_class.__init__).parameters.keys())[1:]
            if not init_params or init_params[0] != "config":
                raise PythonOpError("__init__ first argument (after self) must be `config`")
            op.kernel_params = init_params[1:]
            op.stream_params = list(signature(fn_or_class.new_stream).parameters.keys())[1:]

        # ---- hand the op to the engine
        def c_strings(items):
            return (ctypes.c_char_p * max(1, len(items)))(*[s.encode() for s in items])

        def c_ints(items):
            return (ctypes.c_int32 * max(1, len(items)))(*[int(v) for v in items])

        n_in = 0 if op.variadic else len(op.inputs)
        gpu = device_type is not None and int(device_type) == 1
        if device_sets:
            gpu = gpu or any(int(d[0]) == 1 for d in device_sets)
        desc = _Desc(kname.encode(), n_in, c_strings([c.name for c in op.inputs][:n_in]),
                     c_ints([c.is_frame for c in op.inputs][:n_in]), int(op.variadic), len(op.outputs),
                     c_strings([c.name for c in op.outputs]), c_ints([c.is_frame for c in op.outputs]),
                     c_strings(["" if c.is_frame else c.info.cpp_name for c in op.outputs]),
                     c_ints(stencil or []), len(stencil) if can_stencil else 0,
                     -1 if bounded_state is None else int(bounded_state), int(bool(unbounded_state)), int(batch),
                     int(gpu))
        op.callback = _make_callback(op)
        E.check(_lib().scn_register_callback_op(ctypes.byref(desc), op.callback, None), f"register_python_op({kname})")
        PYTHON_OP_REGISTRY[kname] = op
        return fn_or_class

    return dec
