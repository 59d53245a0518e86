"""The reference's tutorials, replayed with only the import changed.

examples/tutorials/00_basic.py ... 07_profiling.py are the scripts a Scanner user starts from.  Each
test below is the body of one of them against `scanner_b200` (the downloaded example video replaced
by a synthetic .mp4 written here, table names kept), followed by assertions on the values the
tutorial only prints: decode through NVDEC, the GPU stdlib ops, Python kernels, sampling, slicing,
frame outputs, save_mp4, the profile trace.
"""
import json
from typing import Sequence

import cv2
import numpy as np
import pytest

import oracle
import scanner_b200 as sp
from oracle import synth
from scanner_b200 import engine as E

pytestmark = pytest.mark.gpu

W, H, FRAMES, GOP = 320, 240, 120, 12


def _planes():
    rgb = np.stack([synth.smooth_frame(40 + i, H, W) for i in range(FRAMES)])
    f = rgb.astype(np.float32)
    y = 16 + 0.257 * f[..., 0] + 0.504 * f[..., 1] + 0.098 * f[..., 2]
    cb = 128 - 0.148 * f[..., 0] - 0.291 * f[..., 1] + 0.439 * f[..., 2]
    cr = 128 + 0.439 * f[..., 0] - 0.368 * f[..., 1] - 0.071 * f[..., 2]
    sub = lambda p: p.reshape(FRAMES, H // 2, 2, W // 2, 2).mean(axis=(2, 4))  # noqa: E731
    q = lambda p: np.clip(np.rint(p), 0, 255).astype(np.uint8)  # noqa: E731
    return q(y), q(sub(cb)), q(sub(cr))


@pytest.fixture(scope="module")
def example(tmp_path_factory):
    """-> (path of the .mp4, the frames a decoder shows for it)"""
    y, u, v = _planes()
    yuv = np.concatenate([y.reshape(FRAMES, -1), u.reshape(FRAMES, -1), v.reshape(FRAMES, -1)], axis=1)
    path = str(tmp_path_factory.mktemp("video") / "example.mp4")
    open(path, "wb").write(E.mp4_mux(E.h264_synth(yuv, W, H, gop=GOP), 24, 1))
    shown = []
    for i in range(FRAMES):
        chroma = np.empty((H // 2, W), np.uint8)
        chroma[:, 0::2], chroma[:, 1::2] = u[i], v[i]
        shown.append(oracle.nv12_to_rgb(y[i], chroma))
    return path, np.stack(shown)


@pytest.fixture()
def sc(tmp_path):
    c = sp.Client(db_path=str(tmp_path / "scanner_db"))
    yield c
    c.stop()


def test_00_basic(sc, example):
    path, shown = example
    video_stream1 = sp.NamedVideoStream(sc, "example1", path=path)
    frames = sc.io.Input([video_stream1])
    hists = sc.ops.Histogram(frame=frames)
    named_stream1 = sp.NamedStream(sc, "example1_hist")
    output_op = sc.io.Output(hists, [named_stream1])
    sc.run(output_op, sp.PerfParams.estimate())

    named_stream1.delete(sc)
    video_stream2 = sp.NamedVideoStream(sc, "example2", path=path, inplace=True)  # py_test.py 'test1_inplace'
    frames = sc.io.Input([video_stream1, video_stream2])
    hists = sc.ops.Histogram(frame=frames)
    named_stream2 = sp.NamedStream(sc, "example2_hist")
    output_op = sc.io.Output(hists, [named_stream1, named_stream2])
    sc.run(output_op, sp.PerfParams.estimate())

    num_rows = 0
    for i, hist in enumerate(named_stream1.load()):
        assert len(hist) == 3
        assert hist[0].shape[0] == 16
        assert (np.stack(hist) == oracle.hist16(shown[i])).all()
        num_rows += 1
    assert num_rows == video_stream1.len() == FRAMES
    assert named_stream2.len() == FRAMES
    # py_test.py test_load_video_column / test_gather_video_column: frames of an ingested video
    assert (next(video_stream1.load()) == shown[0]).all()
    rows = [0, 10, 100, 57]
    frames = list(video_stream2.load(rows=rows))
    assert len(frames) == len(rows) and all((f == shown[r]).all() for f, r in zip(frames, rows))
    table = sc.table("example1")  # py_test.py test_table_properties
    assert table.num_rows() == FRAMES and table.column_names() == ["index", "frame"]
    assert (next(table.column("frame").load(rows=[33])) == shown[33]).all()
    for stream in [video_stream1, video_stream2, named_stream1, named_stream2]:
        stream.delete(sc)


@sp.register_python_op()
def resize_fn(config, frame: sp.FrameType) -> sp.FrameType:
    return cv2.resize(frame, (config.args["width"], config.args["height"]))


@sp.register_python_op()
class ResizeClass(sp.Kernel):
    def __init__(self, config, width, height):
        self._width = width
        self._height = height

    def execute(self, frame: sp.FrameType) -> sp.FrameType:
        return cv2.resize(frame, (self._width, self._height))


def test_01_defining_python_ops(sc, example, tmp_path):
    path, shown = example
    video_stream = sp.NamedVideoStream(sc, "example", path=path)
    frames = sc.io.Input([video_stream])
    resized_fn_frames = sc.ops.resize_fn(frame=frames, width=64, height=48)
    resized_class_frames = sc.ops.ResizeClass(frame=frames, width=32, height=24)
    fn_stream = sp.NamedVideoStream(sc, "fn_frames")
    fn_output = sc.io.Output(resized_fn_frames, [fn_stream])
    class_stream = sp.NamedVideoStream(sc, "class_frames")
    class_output = sc.io.Output(resized_class_frames, [class_stream])
    sc.run([fn_output, class_output], sp.PerfParams.estimate())

    for i, (a, b) in enumerate(zip(fn_stream.load(), class_stream.load())):
        assert (a == cv2.resize(shown[i], (64, 48))).all() and (b == cv2.resize(shown[i], (32, 24))).all()
    out = fn_stream.save_mp4(str(tmp_path / "01_resized_fn"))
    class_stream.save_mp4(str(tmp_path / "01_resized_class"))
    cap = cv2.VideoCapture(out)
    assert int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == FRAMES and cap.read()[1].shape == (48, 64, 3)
    for stream in [fn_stream, class_stream]:
        stream.delete(sc)


def test_02_op_attributes(sc, example):
    path, shown = example
    video_stream = sp.NamedVideoStream(sc, "example", path=path)

    @sp.register_python_op(device_type=sp.DeviceType.CPU)
    def device_resize(config, frame: sp.FrameType) -> sp.FrameType:
        return cv2.resize(frame, (config.args["width"], config.args["height"]))

    frames = sc.io.Input([video_stream])
    stream = sp.NamedVideoStream(sc, "example_resize")
    sc.run(sc.io.Output(sc.ops.device_resize(frame=frames, width=64, height=48), [stream]), sp.PerfParams.estimate())
    assert stream.len() == FRAMES

    @sp.register_python_op(batch=10)
    def batch_resize(config, frame: Sequence[sp.FrameType]) -> Sequence[sp.FrameType]:
        return [cv2.resize(fr, (config.args["width"], config.args["height"])) for fr in frame]

    frame = sc.io.Input([video_stream])
    resized_frame = sc.ops.batch_resize(frame=frame, width=64, height=48, batch=10)
    stream = sp.NamedVideoStream(sc, "example_batch_resize")
    sc.run(sc.io.Output(resized_frame, [stream]), sp.PerfParams.estimate())
    assert all((f == cv2.resize(shown[i], (64, 48))).all() for i, f in enumerate(stream.load()))

    # the tutorial's two flow ops (dense Farneback on the stencil pair, then an HSV rendering of the
    # field), written out here rather than copied
    @sp.register_python_op(stencil=[0, 1])
    def optical_flow(config, frame: Sequence[sp.FrameType]) -> sp.FrameType:
        prev, nxt = (cv2.cvtColor(f, cv2.COLOR_BGR2GRAY) for f in frame)
        return cv2.calcOpticalFlowFarneback(prev, nxt, None, pyr_scale=0.5, levels=3, winsize=15, iterations=3,
                                            poly_n=5, poly_sigma=1.2, flags=0)

    @sp.register_python_op()
    def visualize_flow(config, flow: sp.FrameType) -> sp.FrameType:
        magnitude, angle = cv2.cartToPolar(flow[..., 0], flow[..., 1])
        hue = (angle * (90.0 / np.pi)).astype(np.uint8)                      # 0..2pi -> 0..180
        value = cv2.normalize(magnitude, None, 0, 255, cv2.NORM_MINMAX).reshape(magnitude.shape).astype(np.uint8)
        hsv = np.dstack([hue, np.full_like(hue, 255), value])
        return cv2.cvtColor(hsv, cv2.COLOR_HSV2BGR)

    frames = sc.io.Input([video_stream])
    range_frames = sc.streams.Range(frames, [(0, 30)])
    flows = sc.ops.optical_flow(frame=range_frames, stencil=[0, 1])
    flow_viz_frames = sc.ops.visualize_flow(flow=flows)
    stream = sp.NamedVideoStream(sc, "example_flow")
    sc.run(sc.io.Output(flow_viz_frames, [stream]), sp.PerfParams.estimate())
    got = list(stream.load())
    assert len(got) == 30 and got[0].shape == (H, W, 3) and got[0].dtype == np.uint8
    want = visualize_flow(None, optical_flow(None, [shown[3], shown[4]]))
    assert (got[3] == want).all()


def _background_subtraction():
    if "BackgroundSubtraction" in sp.pyops.PYTHON_OP_REGISTRY:
        return

    @sp.register_python_op(bounded_state=60)
    class BackgroundSubtraction(sp.Kernel):
        """Tutorial 02/04's stateful op: an exponential running mean of the frames is the background;
        pixels within `threshold` of it (in any channel) are blanked.  State restarts at reset()."""

        def __init__(self, config, alpha, threshold):
            self.alpha, self.limit = alpha, 255 * threshold
            self.background = None

        def reset(self):
            self.background = None

        def execute(self, frame: sp.FrameType) -> sp.FrameType:
            if self.background is None:
                self.background = frame
            near = (np.abs(frame - self.background) < self.limit).any(axis=2)
            out = frame.copy()
            out[near] = 0
            self.background = self.background * (1.0 - self.alpha) + frame * self.alpha
            return out


def test_03_sampling(sc, example):
    path, shown = example
    video_stream = sp.NamedVideoStream(sc, "example", path=path)
    frames = sc.io.Input([video_stream])
    strided_frames = sc.streams.Stride(frames, [4])
    hists = sc.ops.Histogram(frame=strided_frames)
    hist_stream = sp.NamedVideoStream(sc, "example_hist_strided")
    output = sc.io.Output(hists, [hist_stream])
    sc.run(output, sp.PerfParams.estimate())
    num_rows = 0
    for frame_hists in hist_stream.load():
        assert len(frame_hists) == 3
        assert frame_hists[0].shape[0] == 16
        assert (np.stack(frame_hists) == oracle.hist16(shown[4 * num_rows])).all()
        num_rows += 1
    assert num_rows == round(video_stream.len() / 4)
    video_stream.delete(sc)
    hist_stream.delete(sc)
    sc.streams.Range(frames, [(0, 100)])
    sc.streams.Gather(frames, [[10, 17, 32]])


def test_04_slicing(sc, example, tmp_path):
    path, shown = example
    _background_subtraction()
    video_stream = sp.NamedVideoStream(sc, "example", path=path)
    frame = sc.io.Input([video_stream])
    sc.streams.Slice(frame, partitions=[sc.partitioner.all(50)])

    frame = sc.io.Input([video_stream])
    scene_partitions = sc.partitioner.ranges([(20, 45), (45, 110)])
    sliced_frame = sc.streams.Slice(frame, partitions=[scene_partitions])
    masked_frame = sc.ops.BackgroundSubtraction(frame=sliced_frame, alpha=0.02, threshold=0.05, bounded_state=60)
    unsliced_frame = sc.streams.Unslice(masked_frame)
    stream = sp.NamedVideoStream(sc, "04_masked_video")
    output = sc.io.Output(unsliced_frame, [stream])
    sc.run(output, sp.PerfParams.estimate())
    got = list(stream.load())
    assert len(got) == 90
    # the state of the op restarts at each slice: the first frame of a scene is its own background
    assert not got[0].any() and not got[25].any()
    # sequential replay of the kernel over the first scene
    k = sp.pyops.PYTHON_OP_REGISTRY["BackgroundSubtraction"].target(None, 0.02, 0.05)
    k.reset()
    for i in range(25):
        assert (k.execute(shown[20 + i]) == got[i]).all(), i
    out = stream.save_mp4(str(tmp_path / "04_masked"))
    assert int(cv2.VideoCapture(out).get(cv2.CAP_PROP_FRAME_COUNT)) == 90
    stream.delete(sc)


def test_06_compression(sc, example, tmp_path):
    path, shown = example

    def make_blurred_frame(streams):
        frames = sc.io.Input(streams)
        blurred_frames = sc.ops.Blur(frame=frames, kernel_size=3, sigma=0.5)
        sampled_frames = sc.streams.Range(blurred_frames, [(0, 30)])
        return frames, sampled_frames

    video_stream = sp.NamedVideoStream(sc, "example", path=path)
    frame, blurred_frame = make_blurred_frame([video_stream])
    stream = sp.NamedVideoStream(sc, "output_table_name")
    sc.run(sc.io.Output(blurred_frame, [stream]), sp.PerfParams.estimate())
    got = list(stream.load())
    assert len(got) == 30
    want = oracle.blur(shown[7], 3)
    assert (got[7][1:-1, 1:-1] == want[1:-1, 1:-1]).all()
    stream.delete(sc)

    frame, blurred_frame = make_blurred_frame([video_stream])
    low_quality_stream = sp.NamedVideoStream(sc, "low_quality_video")
    sc.run(sc.io.Output(blurred_frame.compress_video(quality=35), [low_quality_stream]), sp.PerfParams.estimate())
    frame, blurred_frame = make_blurred_frame([video_stream])
    lossless_stream = sp.NamedVideoStream(sc, "lossless_video")
    sc.run(sc.io.Output(blurred_frame.lossless(), [lossless_stream]), sp.PerfParams.estimate())
    assert all((a == b).all() for a, b in zip(lossless_stream.load(), got))
    out = low_quality_stream.save_mp4(str(tmp_path / "low_quality_video"))
    assert int(cv2.VideoCapture(out).get(cv2.CAP_PROP_FRAME_COUNT)) == 30
    low_quality_stream.delete(sc)
    lossless_stream.delete(sc)


def test_07_profiling(sc, example, tmp_path):
    path, shown = example
    video_stream = sp.NamedVideoStream(sc, "example", path=path)
    frames = sc.io.Input([video_stream])
    resized_frames = sc.ops.Resize(frame=frames, width=[64], height=[48])
    output_stream = sp.NamedVideoStream(sc, "example_profile")
    output = sc.io.Output(resized_frames, [output_stream])
    job_id = sc.run(output, sp.PerfParams.estimate())
    profile = sc.get_profile(job_id)
    trace = str(tmp_path / "resize-graph.trace")
    profile.write_trace(trace)
    events = json.load(open(trace))
    events = events["traceEvents"] if isinstance(events, dict) else events
    names = {e.get("name", "") for e in events}
    assert any("Resize" in n for n in names) and len(events) > 10
    got = list(output_stream.load())
    assert len(got) == FRAMES and np.abs(got[5].astype(int) - oracle.resize(shown[5], 64, 48).astype(int)).max() <= 1
    video_stream.delete(sc)
    output_stream.delete(sc)
