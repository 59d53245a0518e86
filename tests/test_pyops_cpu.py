"""Python kernels (`@register_python_op`) through the engine, CPU only.

The cases follow the reference's own tests of the feature (tests/py_test.py:557-760, 1042-1060,
1125-1128: TestPy, ResourceTest, TestPyBatch, TestPyStencil, TestPyStencilBatch, TestPyVariadic,
TestPyFail, CacheTest) plus what they leave out: typed columns stored in a database, null rows,
state handling, registration errors.
"""
import os
import pickle
import struct
import subprocess
import sys
from typing import Any, Sequence, Tuple

import numpy as np
import pytest

import scanner_b200 as sp
from scanner_b200 import FrameType, Kernel, register_python_op
from scanner_b200 import pyops
from scanner_b200 import types as T

H, W, N = 12, 16, 30


def frames(n=N):
    return np.stack([np.full((H, W, 3), i, np.uint8) for i in range(n)])


@pytest.fixture()
def sc():
    c = sp.Client(gpus=[], cpu_instances=3)
    yield c
    c.stop()


def video(sc, name="test1", n=N):
    return sp.NamedVideoStream(sc, name, frames=frames(n))


# ------------------------------------------------------------------ ops (registered once per process)
@register_python_op(name="TestPy")
class PyArgs(Kernel):
    def __init__(self, config, kernel_arg):
        assert kernel_arg == 1
        self.x = 20
        self.y = 20

    def new_stream(self, x, y):
        self.x = x
        self.y = y

    def execute(self, frame: FrameType) -> Any:
        return {"x": self.x, "y": self.y, "v": int(frame[0, 0, 0])}


@register_python_op()
class ResourceTest(Kernel):
    def __init__(self, config, path):
        self.path = path

    def fetch_resources(self):
        with open(self.path) as f:
            n = int(f.read())
        with open(self.path, "w") as f:
            f.write(str(n + 1))

    def setup_with_resources(self):
        with open(self.path) as f:
            assert int(f.read()) == 1

    def execute(self, frame: FrameType) -> Any:
        return None


@register_python_op(name="TestPyBatch", batch=50)
class PyBatch(Kernel):
    def execute(self, frame: Sequence[FrameType]) -> Sequence[bytes]:
        return [struct.pack("<ii", int(f[0, 0, 0]), len(frame)) for f in frame]


@register_python_op(name="TestPyStencil", stencil=[0, 1])
class PyStencil(Kernel):
    def execute(self, frame: Sequence[FrameType]) -> bytes:
        assert len(frame) == 2
        return struct.pack("<ii", int(frame[0][0, 0, 0]), int(frame[1][0, 0, 0]))


@register_python_op(name="TestPyStencilBatch", stencil=[0, 1], batch=50)
class PyStencilBatch(Kernel):
    def __init__(self, config):
        pass

    def close(self):
        pass

    def execute(self, frame: Sequence[Sequence[FrameType]]) -> Sequence[bytes]:
        assert len(frame[0]) == 2
        return [struct.pack("<ii", int(w[0][0, 0, 0]), int(w[1][0, 0, 0])) for w in frame]


@register_python_op(name="TestPyVariadic")
class PyVariadic(Kernel):
    def execute(self, *frame: Tuple[FrameType, ...]) -> FrameType:
        assert len(frame) == 3
        return frame[0] + frame[1] + frame[2]


@register_python_op(name="TestPyFail")
class PyFail(Kernel):
    def execute(self, frame: FrameType) -> bytes:
        raise sp.ScannerException("Test")


@register_python_op(name="CacheTest")
def cache_test(config, n: Any) -> Any:
    return n + (config.args or {}).get("step", 1)


@register_python_op()
def EveryThirdIsNull(config, frame: FrameType) -> Tuple[T.NumpyArrayFloat32, FrameType]:
    v = int(frame[0, 0, 0])
    if v % 3 == 0:
        return None, None
    return np.array([v, v * 2], np.float32), (frame[:4, :5, :1].astype(np.float32) / 2)


@register_python_op(unbounded_state=True)
class RunningCount(Kernel):
    """counts the rows seen since the last reset(): needs every earlier row of its task"""

    def __init__(self, config):
        self.n = 0
        self.resets = 0

    def reset(self):
        self.n = 0
        self.resets += 1

    def execute(self, frame: FrameType) -> Any:
        self.n += 1
        return (int(frame[0, 0, 0]), self.n)


@register_python_op(bounded_state=2)
class Warm(Kernel):
    """output = number of consecutive rows seen since reset, saturating at 3 (warmup 2)"""

    def __init__(self, config):
        self.n = 0

    def reset(self):
        self.n = 0

    def execute(self, frame: FrameType) -> Any:
        self.n += 1
        return (int(frame[0, 0, 0]), min(self.n, 3))


def run(sc, col, name, ios=10, wps=5, video_out=False):
    out = (sp.NamedVideoStream if video_out else sp.NamedStream)(sc, name)
    sc.run(sc.io.Output(col, [out]), sp.PerfParams.manual(wps, ios), cache_mode=sp.CacheMode.Overwrite)
    return out


# ------------------------------------------------------------------ the reference's cases
def test_python_kernel_init_and_stream_args(sc):
    frame = sc.io.Input([video(sc)])
    rng = sc.streams.Range(frame, ranges=[{"start": 0, "end": 30}])
    out = run(sc, sc.ops.TestPy(frame=rng, kernel_arg=1, x=[0], y=[0]), "test_hist")
    rows = list(out.load())
    assert len(rows) == 30 and rows[7] == {"x": 0, "y": 0, "v": 7}
    # the wrong init argument fails the constructor: a validation error, not a crash
    with pytest.raises(sp.ScannerException, match="failed validation"):
        run(sc, sc.ops.TestPy(frame=rng, kernel_arg=2, x=[0], y=[0]), "test_hist2")
    # an unknown keyword is an init argument the constructor does not take
    with pytest.raises(sp.ScannerException, match=r"(?s)failed validation.*unexpected keyword argument 'z'"):
        run(sc, sc.ops.TestPy(frame=rng, kernel_arg=1, z=3, x=[0], y=[0]), "test_hist3")
    with pytest.raises(sp.ScannerException, match="must be a list"):
        sc.ops.TestPy(frame=rng, kernel_arg=1, x=0, y=[0])
    with pytest.raises(sp.ScannerException, match="for stream parameters"):
        sc.ops.TestPy(frame=rng, kernel_arg=1)
    with pytest.raises(sp.ScannerException, match="required sequence frame as input"):
        sc.ops.TestPy(kernel_arg=1, x=[0], y=[0])


def test_fetch_resources_runs_once_per_op_and_before_every_setup(sc, tmp_path):
    path = str(tmp_path / "counter")
    open(path, "w").write("0")
    frame = sc.io.Input([video(sc)])
    out = run(sc, sc.ops.ResourceTest(frame=frame, path=path), "test_resource", ios=3, wps=3)
    assert out.len() == N and open(path).read() == "1"  # three pipeline instances, one fetch


def test_python_batch_kernel(sc):
    frame = sc.io.Input([video(sc)])
    out = run(sc, sc.ops.TestPyBatch(frame=frame, batch=50), "test_batch", ios=30, wps=30)
    rows = [struct.unpack("<ii", r) for r in out.load()]
    assert [r[0] for r in rows] == list(range(N)) and {r[1] for r in rows} == {30}
    out = run(sc, sc.ops.TestPyBatch(frame=frame, batch=4), "test_batch4", ios=10, wps=10)
    assert [struct.unpack("<ii", r)[1] for r in out.load()][:10] == [4, 4, 4, 4, 4, 4, 4, 4, 2, 2]


def test_python_stencil_kernel(sc):
    frame = sc.io.Input([video(sc)])
    rows = [struct.unpack("<ii", r) for r in run(sc, sc.ops.TestPyStencil(frame=frame), "test_stencil").load()]
    assert rows == [(i, min(i + 1, N - 1)) for i in range(N)]  # REPEAT_EDGE at the end
    rows = [struct.unpack("<ii", r)
            for r in run(sc, sc.ops.TestPyStencilBatch(frame=frame, batch=50), "test_stencil_batch").load()]
    assert rows == [(i, min(i + 1, N - 1)) for i in range(N)]
    rows = [struct.unpack("<ii", r)
            for r in run(sc, sc.ops.TestPyStencil(frame=frame, stencil=[-2, 0]), "test_stencil2").load()]
    assert rows == [(max(i - 2, 0), i) for i in range(N)]


def test_py_variadic(sc):
    frame = sc.io.Input([video(sc)])
    rng = sc.streams.Range(frame, ranges=[{"start": 0, "end": 30}])
    out = run(sc, sc.ops.TestPyVariadic(rng, rng, rng), "test_variadic", video_out=True)
    got = list(out.load())
    assert len(got) == 30 and all((g == 3 * i).all() and g.shape == (H, W, 3) for i, g in enumerate(got))


def test_python_kernel_exception_fails_the_run_not_the_process(sc):
    frame = sc.io.Input([video(sc)])
    with pytest.raises(sp.ScannerException, match=r"(?s)Op TestPyFail failed.*ScannerException: Test"):
        run(sc, sc.ops.TestPyFail(frame=frame), "test_py_fail")
    # the client is still usable
    assert run(sc, sc.ops.TestPyBatch(frame=frame, batch=8), "after_fail").len() == N


def test_function_op_over_a_byte_stream(sc):
    src = sp.NamedStream(sc, "numbers", rows=[pickle.dumps(i) for i in range(7)])
    col = sc.io.Input([src])
    out = run(sc, sc.ops.CacheTest(n=sc.ops.CacheTest(n=col)), "plus2", ios=2, wps=1)
    assert list(out.load()) == [i + 2 for i in range(7)]
    # function kernels read their init arguments from config.args (tutorial 01_defining_python_ops.py)
    out = run(sc, sc.ops.CacheTest(n=col, step=10), "plus10", ios=2, wps=1)
    assert list(out.load()) == [i + 10 for i in range(7)]


# ------------------------------------------------------------------ beyond the reference's tests
def test_null_rows_and_two_typed_outputs(sc):
    frame = sc.io.Input([video(sc, n=9)])
    vec, img = sc.ops.EveryThirdIsNull(frame=frame)
    o1, o2 = sp.NamedStream(sc, "vec"), sp.NamedVideoStream(sc, "img")
    sc.run([sc.io.Output(vec, [o1]), sc.io.Output(img, [o2])], sp.PerfParams.manual(2, 4))
    v, im = list(o1.load()), list(o2.load())
    for i in range(9):
        if i % 3 == 0:
            assert isinstance(v[i], sp.NullElement) and isinstance(im[i], sp.NullElement)
            assert not v[i] and v[i] == None  # noqa: E711  (falsy, equal to None)
        else:
            assert v[i].dtype == np.float32 and v[i].tolist() == [i, 2 * i]
            assert im[i].dtype == np.float32 and im[i].shape == (4, 5, 1) and (im[i] == i / 2).all()


def test_typed_columns_survive_the_database(tmp_path):
    sc = sp.Client(gpus=[], cpu_instances=2, db_path=str(tmp_path / "db"))
    frame = sc.io.Input([sp.NamedVideoStream(sc, "v", frames=frames(8))])
    sc.run(sc.io.Output(sc.ops.TestPy(frame=frame, kernel_arg=1, x=[4], y=[5]), [sp.NamedStream(sc, "pts")]),
           sp.PerfParams.manual(2, 4))
    sc.stop()
    sc = sp.Client(gpus=[], cpu_instances=1, db_path=str(tmp_path / "db"))
    col = sc.table("pts").column("ret0")
    assert col._desc["type_name"] == "Any"
    assert list(col.load())[3] == {"x": 4, "y": 5, "v": 3}
    assert list(sp.NamedStream(sc, "pts").load(rows=[6])) == [{"x": 4, "y": 5, "v": 6}]
    assert list(sp.NamedStream(sc, "pts").load(ty="Bytes"))[0] == pickle.dumps({"x": 4, "y": 5, "v": 0})
    sc.stop()


def test_storage_hints_and_custom_loaders(sc):
    """`.lossless()` / `.compress_video()` on a frame column (tutorial 06) and `load(fn=...)`,
    `load_bytes()` of the reference's StoredStream."""
    frame = sc.io.Input([video(sc, n=6)])
    neg = sc.ops.Negative(frame=frame)
    for name, col in (("raw", neg.lossless()), ("h264", neg.compress_video(quality=35)),
                      ("dflt", neg.compress("default"))):
        got = list(run(sc, col, name, video_out=True).load())
        assert len(got) == 6 and all((g == 255 - i).all() for i, g in enumerate(got))  # stored exactly
    with pytest.raises(sp.ScannerException, match="not currently supported"):
        neg.compress("av1")
    meta = sc.ops.TestPy(frame=frame, kernel_arg=1, x=[1], y=[2])
    with pytest.raises(sp.ScannerException, match='only supported for sequences of type "video"'):
        meta.lossless()
    out = run(sc, meta, "meta")
    assert list(out.load(fn=lambda b: len(pickle.loads(b))))[:2] == [3, 3]
    assert list(out.load_bytes())[1] == pickle.dumps({"x": 1, "y": 2, "v": 1})


def test_unbounded_state_sees_every_row_of_its_task_after_a_reset(sc):
    frame = sc.io.Input([video(sc)])
    strided = sc.streams.Stride(sc.ops.RunningCount(frame=frame), [5])
    rows = list(run(sc, strided, "running", ios=2, wps=1).load())
    # row r of the output is input row 5r; the kernel was reset at the start of the task and then
    # fed every input row up to it, so its counter equals the input row + 1
    assert rows == [(5 * r, 5 * r + 1) for r in range(6)]


def test_bounded_state_recomputes_the_warmup_rows(sc):
    frame = sc.io.Input([video(sc)])
    rows = list(run(sc, sc.ops.Warm(frame=frame), "warm", ios=6, wps=3).load())
    # every task starts 2 rows early (except at row 0): the third row on is saturated
    assert rows == [(i, min(i + 1, 3)) for i in range(N)]
    rows = list(run(sc, sc.ops.Warm(frame=frame, bounded_state=0), "warm0", ios=10, wps=5).load())
    assert rows == [(i, min(i % 10 + 1, 3)) for i in range(N)]


def test_registration_errors():
    with pytest.raises(pyops.PythonOpError, match="twice"):
        @register_python_op(name="CacheTest")
        def again(config, n: Any) -> Any:
            return n

    with pytest.raises(pyops.PythonOpError, match="No type annotation"):
        @register_python_op()
        def NoAnnotation(config, frame) -> bytes:
            return b""

    with pytest.raises(pyops.PythonOpError, match="Return annotation"):
        @register_python_op()
        def NoReturn(config, frame: FrameType):
            return b""

    with pytest.raises(pyops.PythonOpError, match='"Sequence" type annotation'):
        @register_python_op(batch=4)
        def BatchedNeedsSequences(config, frame: FrameType) -> bytes:
            return b""

    with pytest.raises(pyops.PythonOpError, match="stenciled Op"):
        @register_python_op(stencil=[0, 1])
        def StencilNeedsSequence(config, frame: FrameType) -> bytes:
            return b""

    with pytest.raises(pyops.PythonOpError, match="must be `config`"):
        @register_python_op()
        class BadInit(Kernel):
            def __init__(self, cfg):
                pass

            def execute(self, frame: FrameType) -> bytes:
                return b""

    with pytest.raises(T.ScannerTypeError, match="has not been registered"):
        @register_python_op()
        def UnknownType(config, frame: FrameType) -> int:
            return 1

    for name in ("again", "NoAnnotation", "NoReturn", "BatchedNeedsSequences", "StencilNeedsSequence", "BadInit",
                 "UnknownType"):
        assert name not in pyops.PYTHON_OP_REGISTRY


def test_wrong_output_shape_is_reported(sc):
    @register_python_op(batch=4)
    def ShortBatch(config, frame: Sequence[FrameType]) -> Sequence[bytes]:
        return [b"x"] * (len(frame) - 1)

    @register_python_op()
    def NotAFrame(config, frame: FrameType) -> FrameType:
        return np.zeros((3, 3), np.int64)

    frame = sc.io.Input([video(sc, n=8)])
    with pytest.raises(sp.ScannerException, match="must be a sequence of 4 elements"):
        run(sc, sc.ops.ShortBatch(frame=frame, batch=4), "short")
    with pytest.raises(sp.ScannerException, match="uint8, uint16, float32 or float64"):
        run(sc, sc.ops.NotAFrame(frame=frame), "notaframe", video_out=True)


def test_types_registry_round_trips():
    boxes = T.VariableList("Boxes", T.NumpyArrayFloat32)
    blob = boxes.serialize([np.arange(4, dtype=np.float32), np.zeros(0, np.float32), np.ones(2, np.float32)])
    back = boxes.deserialize(blob)
    assert [b.tolist() for b in back] == [[0, 1, 2, 3], [], [1, 1]]
    assert struct.unpack_from("=QQ", blob) == (3, 16)
    h = T.get_type_info_cpp("Histogram")
    parts = h.deserialize(np.arange(48, dtype=np.int32).tobytes())
    assert len(parts) == 3 and parts[2].tolist() == list(range(32, 48))
    assert h.serialize(parts) == np.arange(48, dtype=np.int32).tobytes()
    assert T.get_type_info(Any).deserialize(T.get_type_info(Any).serialize({"a": 1})) == {"a": 1}
    with pytest.raises(T.ScannerTypeError):
        T.get_type_info_cpp("NoSuchType")
    # a protobuf-backed type and a list of them (the reference's Bbox / BboxList pattern)
    from google.protobuf import descriptor_pb2
    field = T.ProtobufType("FieldProto", descriptor_pb2.FieldDescriptorProto)
    fields = T.VariableList("FieldProtoList", field)
    msgs = [descriptor_pb2.FieldDescriptorProto(name="x", number=1), descriptor_pb2.FieldDescriptorProto(name="y", number=2)]
    back = fields.deserialize(fields.serialize(msgs))
    assert [(m.name, m.number) for m in back] == [("x", 1), ("y", 2)]
    assert T.get_type_info_cpp("FieldProtoList").type is fields


@register_python_op()
def Negative(config, frame: FrameType) -> FrameType:
    return 255 - frame


@pytest.mark.parametrize("with_db", [False, True])
def test_frames_written_by_a_python_op_save_as_mp4(tmp_path, with_db):
    """Tutorials 01/02 end with `stream.save_mp4(...)` on a job's frame output.  Such frames are
    stored uncompressed here; save_mp4 writes them as intra-PCM H.264, which FFmpeg plays back to
    within colour-conversion rounding plus the 2x2 chroma averaging of a gradient."""
    import cv2
    h, w, n = 37, 50, 5  # odd height: padded to 38 in the file
    yy, xx = np.mgrid[0:h, 0:w]
    src = np.stack([np.stack([xx * 4 + i, yy * 5, 200 - xx * 2 + 0 * yy], axis=2).clip(0, 255).astype(np.uint8)
                    for i in range(n)])
    sc = sp.Client(gpus=[], cpu_instances=2, db_path=str(tmp_path / "db") if with_db else None)
    frame = sc.io.Input([sp.NamedVideoStream(sc, "src", frames=src)])
    out = sp.NamedVideoStream(sc, "neg")
    sc.run(sc.io.Output(sc.ops.Negative(frame=frame), [out]), sp.PerfParams.manual(2, 4))
    path = out.save_mp4(str(tmp_path / "neg"), fps=30)
    cap = cv2.VideoCapture(path)
    assert int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == n and abs(cap.get(cv2.CAP_PROP_FPS) - 30) < 1e-6
    for i in range(n):
        ok, bgr = cap.read()
        assert ok and bgr.shape == (h + 1, w, 3)
        err = np.abs(bgr[:h, :, ::-1].astype(np.int32) - (255 - src[i]).astype(np.int32))
        assert err.max() <= 8 and err.mean() < 2.5, (i, err.max(), err.mean())
    if with_db:
        out.delete()
    sc.stop()


@register_python_op()
def FrameSum(config, frame: FrameType) -> Any:
    return int(frame.sum())


@pytest.mark.parametrize("with_db", [False, True])
def test_outputs_of_one_job_are_inputs_of_the_next(tmp_path, with_db):
    """Tables written by a job (byte rows and frames) feed later jobs, in memory or through the
    database, and a re-written stream is re-read rather than served from a stale binding."""
    sc = sp.Client(gpus=[], cpu_instances=2, db_path=str(tmp_path / "db") if with_db else None)
    v = sp.NamedVideoStream(sc, "v", frames=frames(9))
    sums, inv = sp.NamedStream(sc, "sums"), sp.NamedVideoStream(sc, "inv")
    frame = sc.io.Input([v])
    sc.run([sc.io.Output(sc.ops.FrameSum(frame=frame), [sums]), sc.io.Output(sc.ops.Negative(frame=frame), [inv])],
           sp.PerfParams.manual(2, 4))
    out2, out3 = sp.NamedStream(sc, "sums_plus"), sp.NamedStream(sc, "inv_sums")
    sc.run([sc.io.Output(sc.ops.CacheTest(n=sc.io.Input([sums])), [out2]),
            sc.io.Output(sc.ops.FrameSum(frame=sc.io.Input([inv])), [out3])], sp.PerfParams.manual(2, 4))
    px = H * W * 3
    assert list(out2.load()) == [i * px + 1 for i in range(9)]
    assert list(out3.load()) == [(255 - i) * px for i in range(9)]
    # overwrite `sums` with other values and consume it again
    sc.run(sc.io.Output(sc.ops.CacheTest(n=sc.io.Input([out2]), step=100), [sums]), sp.PerfParams.manual(2, 4),
           cache_mode=sp.CacheMode.Overwrite)
    sc.run(sc.io.Output(sc.ops.CacheTest(n=sc.io.Input([sums])), [out2]), sp.PerfParams.manual(2, 4),
           cache_mode=sp.CacheMode.Overwrite)
    assert list(out2.load()) == [i * px + 102 for i in range(9)]
    with pytest.raises(sp.ScannerException, match="does not exist"):
        sc.run(sc.io.Output(sc.ops.CacheTest(n=sc.io.Input([sp.NamedStream(sc, "nope")])), [out3]),
               sp.PerfParams.manual(2, 4), cache_mode=sp.CacheMode.Overwrite)
    sc.stop()


@register_python_op()
class Fragile(Kernel):
    def __init__(self, config, fail_in=""):
        self.fail_in = fail_in

    def fetch_resources(self):
        if self.fail_in == "fetch":
            raise RuntimeError("no network here")

    def setup_with_resources(self):
        if self.fail_in == "setup":
            raise RuntimeError("model file is missing")

    def new_stream(self, mode="ok"):
        if mode == "bad":
            raise ValueError("unknown mode")

    def execute(self, frame: FrameType) -> bytes:
        return b"x"


def test_failures_outside_execute_are_reported_too(sc):
    frame = sc.io.Input([video(sc, n=6)])
    with pytest.raises(sp.ScannerException, match=r"(?s)failed to fetch resources.*no network here"):
        run(sc, sc.ops.Fragile(frame=frame, fail_in="fetch", mode=["ok"]), "f1")
    with pytest.raises(sp.ScannerException, match=r"(?s)failed setup.*model file is missing"):
        run(sc, sc.ops.Fragile(frame=frame, fail_in="setup", mode=["ok"]), "f2")
    with pytest.raises(sp.ScannerException, match=r"(?s)failed in reset / new_stream.*unknown mode"):
        run(sc, sc.ops.Fragile(frame=frame, mode=["bad"]), "f3")
    assert run(sc, sc.ops.Fragile(frame=frame, mode=["ok"]), "f4").len() == 6


def test_the_cpu_example_runs():
    """examples/python_ops_cpu.py (own process: it registers ops by global name)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "examples", "python_ops_cpu.py")],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "smoothed brightness of every third frame: [32.0, 36.0, 42.0" in out.stdout and "tinted.mp4" in out.stdout


@pytest.mark.parametrize("with_db", [False, True])
def test_files_in_files_out(tmp_path, with_db):
    """The reference's tutorial 05_sources_sinks.py shape: image files -> ImageDecoder -> an op ->
    ImageEncoder -> image files, through FilesStream (png: the in-tree encoder's only format)."""
    import cv2
    rng = np.random.default_rng(4)
    imgs = [rng.integers(0, 256, (20 + i, 30, 3), dtype=np.uint8) for i in range(3)]
    src_paths = [str(tmp_path / f"sample-frame-{i + 1}.png") for i in range(3)]
    for p, im in zip(src_paths, imgs):
        cv2.imwrite(p, im[..., ::-1])
    dst_paths = [str(tmp_path / f"negative-{i + 1}.png") for i in range(3)]
    with sp.Client(gpus=[], cpu_instances=2, db_path=str(tmp_path / "db") if with_db else None) as sc:
        image_stream = sp.FilesStream(src_paths)
        frames = sc.ops.ImageDecoder(img=sc.io.Input([image_stream]))
        encoded = sc.ops.ImageEncoder(frame=sc.ops.Negative(frame=frames), format="png")
        out_stream = sp.FilesStream(dst_paths)
        assert not out_stream.exists()
        sc.run(sc.io.Output(encoded, [out_stream]), sp.PerfParams.estimate(), cache_mode=sp.CacheMode.Overwrite)
        assert out_stream.exists() and (not with_db or sc.table_names() == [])
    for p, im in zip(dst_paths, imgs):
        assert (cv2.imread(p)[..., ::-1] == 255 - im).all()
    assert [len(b) for b in out_stream.load()] == [os.path.getsize(p) for p in dst_paths]
    out_stream.delete()
    assert not out_stream.exists()
