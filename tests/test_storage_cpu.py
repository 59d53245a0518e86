"""Tables on disk and .mp4 ingest, CPU only.

Independent readers pin the formats:
  * FFmpeg (cv2.VideoCapture) must open the .mp4 files our muxer writes and decode the same pictures
    it decodes from the raw Annex-B stream -- so the demuxer is tested on files a third party accepts;
  * the real protobuf runtime (google.protobuf, message types declared here with the reference's
    field numbers, scanner/metadata.proto:6-23,55-126) must parse every descriptor we write."""
import os
import struct
import subprocess

import numpy as np
import pytest

import oracle
from oracle import synth
from scanner_b200 import engine as E
from scanner_b200 import protolite

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TEST_ARGS = protolite.parse_proto(open(os.path.join(ROOT, "tests", "cpp", "test_args.proto")).read())


# ------------------------------------------------------------------ protobuf-runtime message types
def _reference_messages():
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    F = descriptor_pb2.FieldDescriptorProto
    fd = descriptor_pb2.FileDescriptorProto(name="ref_metadata_subset.proto", package="refmeta", syntax="proto3")

    def msg(name, fields):
        m = fd.message_type.add(name=name)
        for fname, num, ftype, rep, tname in fields:
            f = m.field.add(name=fname, number=num, type=ftype,
                            label=F.LABEL_REPEATED if rep else F.LABEL_OPTIONAL)
            if tname:
                f.type_name = ".refmeta." + tname
    msg("DbEntry", [("id", 1, F.TYPE_INT32, 0, None), ("name", 2, F.TYPE_STRING, 0, None),
                    ("committed", 3, F.TYPE_BOOL, 0, None)])
    msg("DatabaseDescriptor", [("next_bulk_job_id", 1, F.TYPE_INT32, 0, None), ("next_table_id", 2, F.TYPE_INT32, 0, None),
                               ("bulk_jobs", 3, F.TYPE_MESSAGE, 1, "DbEntry"), ("tables", 4, F.TYPE_MESSAGE, 1, "DbEntry")])
    msg("Column", [("id", 1, F.TYPE_INT32, 0, None), ("name", 2, F.TYPE_STRING, 0, None),
                   ("type", 3, F.TYPE_INT32, 0, None), ("type_name", 4, F.TYPE_STRING, 0, None)])
    msg("TableDescriptor", [("id", 1, F.TYPE_INT32, 0, None), ("name", 2, F.TYPE_STRING, 0, None),
                            ("columns", 3, F.TYPE_MESSAGE, 1, "Column"), ("end_rows", 4, F.TYPE_INT64, 1, None),
                            ("job_id", 6, F.TYPE_INT32, 0, None), ("timestamp", 7, F.TYPE_INT64, 0, None)])
    msg("VideoDescriptor", [
        ("table_id", 1, F.TYPE_INT32, 0, None), ("column_id", 2, F.TYPE_INT32, 0, None),
        ("item_id", 3, F.TYPE_INT32, 0, None), ("frames", 4, F.TYPE_INT64, 0, None),
        ("width", 5, F.TYPE_INT32, 0, None), ("height", 6, F.TYPE_INT32, 0, None),
        ("codec_type", 7, F.TYPE_INT32, 0, None), ("chroma_format", 8, F.TYPE_INT32, 0, None),
        ("sample_offsets", 9, F.TYPE_UINT64, 1, None), ("sample_sizes", 10, F.TYPE_UINT64, 1, None),
        ("keyframe_indices", 11, F.TYPE_UINT64, 1, None), ("metadata_packets", 12, F.TYPE_BYTES, 0, None),
        ("frame_type", 13, F.TYPE_INT32, 0, None), ("channels", 14, F.TYPE_INT32, 0, None),
        ("time_base_num", 15, F.TYPE_INT32, 0, None), ("time_base_denom", 16, F.TYPE_INT32, 0, None),
        ("num_encoded_videos", 17, F.TYPE_INT64, 0, None), ("frames_per_video", 18, F.TYPE_INT64, 1, None),
        ("keyframes_per_video", 19, F.TYPE_INT64, 1, None), ("size_per_video", 20, F.TYPE_INT64, 1, None),
        ("data_path", 21, F.TYPE_STRING, 0, None), ("inplace", 22, F.TYPE_BOOL, 0, None)])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = getattr(message_factory, "GetMessageClass", None)
    out = {}
    for n in ("DatabaseDescriptor", "TableDescriptor", "VideoDescriptor"):
        d = pool.FindMessageTypeByName("refmeta." + n)
        out[n] = get(d) if get else message_factory.MessageFactory(pool).GetPrototype(d)
    return out


REF = _reference_messages()


def parse_ref(kind, path):
    m = REF[kind]()
    m.ParseFromString(open(path, "rb").read())
    return m


# ------------------------------------------------------------------------------------- fixtures
@pytest.fixture(scope="module", autouse=True)
def plugin():
    oracle.build()
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cpp")], stdout=subprocess.DEVNULL)
    so = os.path.join(ROOT, "build", "tests", "libtest_plugin_ops.so")
    if "TestWindow" not in E.list_ops():
        E.load_op_library(so)


def make_stream(seed, n, h, w, gop, non_key="pcm"):
    rng = np.random.default_rng(seed)
    k = n if non_key == "pcm" else (n + gop - 1) // gop
    yuv = rng.integers(0, 256, (k, h * w * 3 // 2), dtype=np.uint8)
    return E.h264_synth(yuv, w, h, gop=gop, non_key=non_key, frames=n), yuv


def cv2_frames(path):
    import cv2
    cap = cv2.VideoCapture(path)
    assert cap.isOpened(), path
    out = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        out.append(f)
    return out


# ------------------------------------------------------------------------------------- container
@pytest.mark.parametrize("h,w,n,gop,mode", [(48, 64, 12, 4, "pcm"), (96, 128, 9, 3, "skip"), (270, 480, 5, 5, "pcm")])
def test_mp4_written_here_is_read_by_ffmpeg_and_demuxes_back(tmp_path, h, w, n, gop, mode):
    stream, _ = make_stream(3, n, h, w, gop, mode)
    mp4 = E.mp4_mux(stream, 30, 1)
    raw_path, mp4_path = str(tmp_path / "a.h264"), str(tmp_path / "a.mp4")
    open(raw_path, "wb").write(stream)
    open(mp4_path, "wb").write(mp4)
    from_raw, from_mp4 = cv2_frames(raw_path), cv2_frames(mp4_path)
    assert len(from_mp4) == n == len(from_raw)
    for a, b in zip(from_raw, from_mp4):
        assert (a == b).all()
    # our demuxer gives back the elementary stream byte for byte (parameter sets re-inserted in
    # front of every sync sample, which is where the synthetic encoder puts them)
    back, info = E.mp4_demux(mp4)
    assert info == {"width": w, "height": h, "timescale": 30, "duration": n, "samples": n,
                    "sync_samples": (n + gop - 1) // gop}
    assert back == stream


@pytest.mark.parametrize("n,gop", [(11, 6), (9, 9), (8, 4), (2, 2)])
def test_b_picture_stream_is_what_ffmpeg_displays(tmp_path, n, gop):
    """non_key="bidir": B pictures coded after their later anchor.  FFmpeg must show the pictures in
    display order with the B pictures equal to the rounded mean of their anchors -- that pins
    bidir_expected(), which the NVDEC parity tests compare against.  Through .mp4 as well: the
    sample table is in coding order there."""
    import cv2
    h, w = 48, 64
    rng = np.random.default_rng(7)
    yuv = rng.integers(0, 256, (n, h * w * 3 // 2), dtype=np.uint8)
    stream = E.h264_synth(yuv, w, h, gop=gop, non_key="bidir")
    shown = E.bidir_expected(yuv, gop)
    if n > 2 and gop > 2:
        assert not (shown[1] == yuv[1]).all()
    paths = [str(tmp_path / "b.h264"), str(tmp_path / "b.mp4")]
    open(paths[0], "wb").write(stream)
    open(paths[1], "wb").write(E.mp4_mux(stream, 30, 1))
    for path in paths:
        cap = cv2.VideoCapture(path)
        cap.set(cv2.CAP_PROP_CONVERT_RGB, 0)  # the raw luma plane
        for i in range(n):
            ok, f = cap.read()
            assert ok, (path, i)
            assert (np.asarray(f).reshape(-1)[:h * w] == shown[i][:h * w]).all(), (path, i)
        assert not cap.read()[0]


def test_ingested_b_picture_stream_keeps_its_reordering_flag(tmp_path):
    """The flag is re-derived from the SPS stored in the VideoDescriptor (no field for it in the
    reference's descriptor); the sample table of an .mp4 with B pictures is in coding order."""
    h, w = 48, 64
    rng = np.random.default_rng(5)
    yuv = rng.integers(0, 256, (7, h * w * 3 // 2), dtype=np.uint8)
    path = str(tmp_path / "b.mp4")
    open(path, "wb").write(E.mp4_mux(E.h264_synth(yuv, w, h, gop=4, non_key="bidir")))
    db = E.Database(str(tmp_path / "db"))
    db.ingest_video("bclip", path)
    eng = E.Engine(gpus=[], cpu_instances=1)
    sid = db.add_video_stream(eng, "bclip")
    assert eng.stream_may_reorder(sid) and eng.stream_rows(sid) == 7
    assert eng.stream_info(sid)["keyframes"] == 2
    eng.close()
    db.close()


class _Bits:
    def __init__(self):
        self.b = []

    def u(self, n, v):
        self.b += [(v >> i) & 1 for i in range(n - 1, -1, -1)]

    def ue(self, v):
        k = v + 1
        n = k.bit_length() - 1
        self.u(n, 0)
        self.u(n + 1, k)

    def rbsp(self):
        bits = self.b + [1]
        bits += [0] * (-len(bits) % 8)
        raw = bytes(int("".join(map(str, bits[i:i + 8])), 2) for i in range(0, len(bits), 8))
        out, zeros = bytearray(), 0
        for c in raw:  # emulation prevention
            if zeros >= 2 and c <= 3:
                out.append(3)
                zeros = 0
            out.append(c)
            zeros = zeros + 1 if c == 0 else 0
        return bytes(out)


@pytest.mark.parametrize("reorder,hrd", [(0, False), (0, True), (1, False), (None, True)])
def test_sps_vui_bitstream_restriction_decides_reordering(tmp_path, reorder, hrd):
    """What encoders such as x264 write: POC type 0 plus a VUI whose bitstream_restriction says
    max_num_reorder_frames.  0 -> display order is coding order and sparse decodes may stop at the
    wanted picture; anything else (or no restriction coded) -> they may not.  The hand-built SPS is
    swapped into a synthetic stream; FFmpeg decoding it unchanged shows the VUI is well-formed."""
    h, w, n = 48, 64, 2
    rng = np.random.default_rng(9)
    yuv = rng.integers(0, 256, (n, h * w * 3 // 2), dtype=np.uint8)
    stream = E.h264_synth(yuv, w, h, gop=2, non_key="bidir")  # I P: no B picture fits in 2 frames
    b = _Bits()
    b.u(8, 77), b.u(8, 0x40), b.u(8, 51), b.ue(0)
    b.ue(0), b.ue(0), b.ue(4), b.ue(2), b.u(1, 0)         # frame_num, POC type 0 + lsb bits, refs, gaps
    b.ue(w // 16 - 1), b.ue(h // 16 - 1), b.u(1, 1), b.u(1, 1), b.u(1, 0)
    b.u(1, 1)                                             # vui_parameters_present_flag
    b.u(1, 1), b.u(8, 255), b.u(16, 4), b.u(16, 3)        # Extended_SAR 4:3
    b.u(1, 0)                                             # overscan
    b.u(1, 1), b.u(3, 5), b.u(1, 0), b.u(1, 1), b.u(8, 1), b.u(8, 1), b.u(8, 1)   # video_signal_type
    b.u(1, 1), b.ue(0), b.ue(0)                           # chroma_loc
    b.u(1, 1), b.u(32, 1001), b.u(32, 60000), b.u(1, 1)   # timing_info
    for present in (hrd, False):                          # nal_hrd, vcl_hrd
        b.u(1, int(present))
        if present:
            b.ue(1), b.u(4, 4), b.u(4, 6)
            for _ in range(2):
                b.ue(999), b.ue(1999), b.u(1, 0)
            b.u(5, 23), b.u(5, 23), b.u(5, 23), b.u(5, 24)
    if hrd:
        b.u(1, 0)                                         # low_delay_hrd_flag
    b.u(1, 0)                                             # pic_struct_present_flag
    b.u(1, int(reorder is not None))                      # bitstream_restriction_flag
    if reorder is not None:
        b.u(1, 1), b.ue(0), b.ue(0), b.ue(10), b.ue(10), b.ue(reorder), b.ue(2)
    sc = b"\x00\x00\x00\x01"
    assert stream.startswith(sc + b"\x67")
    stream = sc + b"\x67" + b.rbsp() + stream[stream.index(sc, 4):]
    path = str(tmp_path / "v.h264")
    open(path, "wb").write(stream)
    got = cv2_frames(path)
    assert len(got) == n
    eng = E.Engine(gpus=[], cpu_instances=1)
    sid = eng.add_h264(stream)
    assert eng.stream_rows(sid) == n and eng.stream_info(sid)["width"] == w
    assert eng.stream_may_reorder(sid) == (reorder != 0)
    eng.close()


@pytest.mark.parametrize("what,match", [("offset", "lies outside"), ("size", "lies outside"),
                                        ("keyframes", "keyframe"), ("first", "first frame"),
                                        ("width", "does not fit"), ("count", "inconsistent")])
def test_corrupt_video_descriptor_is_rejected_when_bound(tmp_path, what, match):
    """The sample table of a stored descriptor becomes pointers handed to the hardware decoder:
    every entry is checked against the data file when the table is bound to an engine."""
    stream, _ = make_stream(8, 9, 48, 64, 3)
    db = E.Database(str(tmp_path / "db"))
    db.ingest_h264("clip", stream)
    path = str(tmp_path / "db/tables/0/1_0_video_metadata.bin")
    vd = parse_ref("VideoDescriptor", path)
    if what == "offset":
        vd.sample_offsets[4] = len(stream) + 5
    elif what == "size":
        vd.sample_sizes[2] = 2 ** 63
    elif what == "keyframes":
        vd.keyframe_indices[1] = 40
    elif what == "first":
        vd.keyframe_indices[0] = 1
        vd.keyframe_indices[1] = 3
    elif what == "width":
        vd.width = 4096
    elif what == "count":
        vd.frames = 7
    open(path, "wb").write(vd.SerializeToString())
    eng = E.Engine(gpus=[], cpu_instances=1)
    with pytest.raises(E.EngineError, match=match):
        db.add_video_stream(eng, "clip")
    eng.close()
    db.close()


def test_parsers_survive_mutated_files_under_asan(tmp_path):
    """Mutation fuzzing (byte flips, truncation, field smashing) of the Annex-B indexer, the mp4
    demuxer/muxer and the descriptor readers, compiled with AddressSanitizer + UBSan: any
    out-of-bounds access, overflow or leak fails the run."""
    exe = os.path.join(ROOT, "build", "tests", "fuzz_parsers")
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cpp"), exe], stdout=subprocess.DEVNULL)
    seeds = []
    for mode in ("pcm", "bidir", "skip"):
        n = 6
        stream, _ = make_stream(10, n, 48, 64, 3, mode) if mode != "bidir" else (
            E.h264_synth(np.random.default_rng(10).integers(0, 256, (n, 48 * 64 * 3 // 2), dtype=np.uint8), 64, 48,
                         gop=3, non_key="bidir"), None)
        for ext, blob in ((".h264", stream), (".mp4", E.mp4_mux(stream, 30, 1))):
            seeds.append(str(tmp_path / (mode + ext)))
            open(seeds[-1], "wb").write(blob)
    db = E.Database(str(tmp_path / "db"))
    db.ingest_video("a", seeds[1])
    db.close()
    seeds += [str(tmp_path / "db" / f) for f in ("db_metadata.bin", "tables/0/descriptor.bin",
                                                 "tables/0/1_0_video_metadata.bin")]
    out = subprocess.run([exe, "1", "400"] + seeds, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    accepted, rejected = (int(x) for x in out.stdout.split()[1::2])
    assert accepted > 100 and rejected > 100  # both outcomes exercised


def test_mp4_errors_are_reported():
    stream, _ = make_stream(4, 4, 48, 64, 2)
    mp4 = E.mp4_mux(stream)
    with pytest.raises(E.EngineError, match="moov"):
        E.mp4_demux(mp4[:len(mp4) // 2])          # index (moov) cut off
    broken = bytearray(mp4)
    i = broken.find(b"avc1", broken.find(b"stsd"))   # the sample entry, not the ftyp brand
    broken[i:i + 4] = b"hev1"
    with pytest.raises(E.EngineError, match="hev1"):
        E.mp4_demux(bytes(broken))
    with pytest.raises(E.EngineError):
        E.mp4_demux(b"\x00\x00\x00\x08free" + b"\x00" * 32)


# ------------------------------------------------------------------------------------- ingest
def test_ingest_mp4_writes_the_reference_table_layout(tmp_path):
    n, h, w, gop = 10, 48, 64, 5
    stream, _ = make_stream(5, n, h, w, gop)
    mp4_path = str(tmp_path / "clip.mp4")
    open(mp4_path, "wb").write(E.mp4_mux(stream, 24, 1))
    db = E.Database(str(tmp_path / "db"))
    db.ingest_video("clip", mp4_path)
    assert db.tables() == ["clip"] and db.has_table("clip")
    info = db.table_info("clip")
    assert info["rows"] == n and info["width"] == w and info["height"] == h and info["keyframes"] == 2
    assert [c["name"] for c in info["columns"]] == ["index", "frame"] and info["columns"][1]["type"] == "Video"

    root = str(tmp_path / "db")
    meta = parse_ref("DatabaseDescriptor", os.path.join(root, "db_metadata.bin"))
    assert meta.next_table_id == 1 and [(t.id, t.name, t.committed) for t in meta.tables] == [(0, "clip", True)]
    td = parse_ref("TableDescriptor", os.path.join(root, "tables/0/descriptor.bin"))
    assert td.name == "clip" and list(td.end_rows) == [n] and td.job_id == -1
    assert [(c.id, c.name, c.type) for c in td.columns] == [(0, "index", 0), (1, "frame", 1)]
    vd = parse_ref("VideoDescriptor", os.path.join(root, "tables/0/1_0_video_metadata.bin"))
    assert (vd.frames, vd.width, vd.height, vd.channels, vd.codec_type, vd.chroma_format) == (n, w, h, 3, 0, 1)
    assert (vd.time_base_num, vd.time_base_denom) == (1, 24)
    assert list(vd.keyframe_indices) == [0, 5] and list(vd.frames_per_video) == [n]
    data = open(os.path.join(root, "tables/0/1_0.bin"), "rb").read()
    assert data == stream and list(vd.size_per_video) == [len(data)]
    # every sample is one access unit of the stored stream; IDR samples start with the SPS
    offs, sizes = list(vd.sample_offsets), list(vd.sample_sizes)
    assert len(offs) == n and offs[0] == 0 and all(offs[i] + sizes[i] == offs[i + 1] for i in range(n - 1))
    assert offs[-1] + sizes[-1] == len(data)
    for k in vd.keyframe_indices:
        assert data[offs[k]:offs[k] + 5] == b"\x00\x00\x00\x01\x67"
    assert vd.metadata_packets.startswith(b"\x00\x00\x00\x01\x67")
    # index column: row i = int64 i, metadata = n then n sizes (reference ingest.cpp:321-345)
    idx = open(os.path.join(root, "tables/0/0_0.bin"), "rb").read()
    assert idx == b"".join(struct.pack("<q", i) for i in range(n))
    m = open(os.path.join(root, "tables/0/0_0_metadata.bin"), "rb").read()
    assert struct.unpack(f"<{n + 1}Q", m) == (n,) + (8,) * n
    assert db.read_rows("clip", "index", [0, 7]) == [struct.pack("<q", 0), struct.pack("<q", 7)]
    with pytest.raises(E.EngineError, match="compressed"):
        db.read_rows("clip", "frame", [0])
    with pytest.raises(E.EngineError, match="already exists"):
        db.ingest_video("clip", mp4_path)
    with pytest.raises(E.EngineError, match="cannot read"):
        db.ingest_video("other", str(tmp_path / "missing.mp4"))

    # the stored table binds to an engine as an H.264 stream without rescanning
    eng = E.Engine(gpus=[], cpu_instances=1)
    sid = db.add_video_stream(eng, "clip")
    si = eng.stream_info(sid)
    assert (si["is_video"], si["width"], si["height"], si["keyframes"]) == (1, w, h, 2) and eng.stream_rows(sid) == n
    assert not eng.stream_may_reorder(sid)  # POC type 2: display order is coding order
    eng.close()
    db.close()

    # a second process / session sees the committed table; raw .h264 files ingest too
    db2 = E.Database(root)
    assert db2.tables() == ["clip"]
    raw_path = str(tmp_path / "clip.h264")
    open(raw_path, "wb").write(stream)
    db2.ingest_video("raw", raw_path)
    assert db2.table_info("raw")["id"] == 1 and sorted(db2.tables()) == ["clip", "raw"]
    db2.delete_table("clip")
    assert db2.tables() == ["raw"] and not os.path.exists(os.path.join(root, "tables/0"))
    db2.close()


# ------------------------------------------------------------------------------------- job output tables
def test_job_outputs_saved_as_tables_and_read_back(tmp_path):
    n, h, w = 11, 24, 32
    frames = np.stack([synth.rand_frame(200 + i, h, w) for i in range(n)])
    eng = E.Engine(gpus=[], cpu_instances=2)
    g = E.Graph()
    src = g.add_source(True)
    hs = g.add_op("TestHistogramOracle", [(src, "frame")])
    rz = g.add_op("TestResizeOracle", [(src, "frame")])
    s_h, s_r = g.add_sink((hs, "histogram")), g.add_sink((rz, "frame"))
    j = E.Job()
    j.bind_source(src, eng.add_raw_frames(frames))
    j.set_stream_args(rz, protolite.encode(TEST_ARGS["TestSizeArgs"], {"width": 16, "height": 12}))
    eng.run(g, [j], 2, 4)
    db = E.Database(str(tmp_path / "db"))
    tid = db.save_job(j, "out", [(s_h, "histogram", "Histogram"), (s_r, "frame", "")], job_id=7)
    assert tid == 0
    info = db.table_info("out")
    assert info["rows"] == n and info["items"] == 3 and info["job_id"] == 7 and info["keyframes"] == -1
    assert [(c["name"], c["type"], c["type_name"]) for c in info["columns"]] == \
        [("index", "Bytes", ""), ("histogram", "Bytes", "Histogram"), ("frame", "Video", "")]
    td = parse_ref("TableDescriptor", str(tmp_path / "db/tables/0/descriptor.bin"))
    assert list(td.end_rows) == [4, 8, 11]
    vd = parse_ref("VideoDescriptor", str(tmp_path / "db/tables/0/2_1_video_metadata.bin"))
    assert (vd.codec_type, vd.frames, vd.height, vd.width, vd.channels, vd.frame_type) == (2, 4, 12, 16, 3, 0)
    hist = db.read_rows("out", "histogram")
    res = db.read_rows("out", "frame", [0, 5, 10])
    for i in range(n):
        assert hist[i] == oracle.hist16(frames[i]).tobytes() == j.output_row(s_h, i)
    for k, i in enumerate([0, 5, 10]):
        assert res[k].shape == (12, 16, 3) and (res[k] == oracle.resize(frames[i], 16, 12)).all()
    assert db.read_rows("out", "index", [9]) == [struct.pack("<q", 9)]
    with pytest.raises(E.EngineError, match="outside"):
        db.read_rows("out", "histogram", [n])
    with pytest.raises(E.EngineError, match="no column"):
        db.read_rows("out", "nope", [0])
    eng.close()
    db.close()


def test_foreign_mp4_with_another_codec_is_rejected_by_name(tmp_path):
    """A file written by FFmpeg's own muxer (MPEG-4 part 2 video: no H.264 encoder is available to it
    here) walks through the box parser up to the sample entry, where the codec is refused by name."""
    import cv2
    path = str(tmp_path / "foreign.mp4")
    w = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), 25, (128, 96))
    if not w.isOpened():
        pytest.skip("this OpenCV build cannot write mp4v")
    for i in range(12):
        w.write(np.full((96, 128, 3), i * 20, np.uint8))
    w.release()
    data = open(path, "rb").read()
    with pytest.raises(E.EngineError, match="mp4v.*not H.264"):
        E.mp4_demux(data)
    db = E.Database(str(tmp_path / "db"))
    with pytest.raises(E.EngineError, match="not H.264"):
        db.ingest_video("foreign", path)
    assert db.tables() == []
    db.close()


def test_save_stage_writes_items_while_the_job_runs_and_frees_the_rows(tmp_path):
    """reference SaveWorker/ColumnSink: every finished task becomes one item of the output table; with
    keep_rows=False the rows are not held in host memory afterwards."""
    n, h, w = 13, 16, 24
    frames = np.stack([synth.rand_frame(300 + i, h, w) for i in range(n)])
    db = E.Database(str(tmp_path / "db"))
    eng = E.Engine(gpus=[], cpu_instances=2)
    g = E.Graph()
    src = g.add_source(True)
    hs = g.add_op("TestHistogramOracle", [(src, "frame")])
    rz = g.add_op("TestResizeOracle", [(src, "frame")])
    s_h, s_r = g.add_sink((hs, "histogram")), g.add_sink((rz, "frame"))
    j = E.Job()
    j.bind_source(src, eng.add_raw_frames(frames))
    j.set_stream_args(rz, protolite.encode(TEST_ARGS["TestSizeArgs"], {"width": 12, "height": 8}))
    t_h = db.new_table("hists", "histogram", False, "Histogram", job_id=3)
    t_r = db.new_table("small", "frame", True, "", job_id=3)
    assert db.tables() == []                      # reserved, not visible until committed
    eng.run(g, [j], 5, 5)                         # rows kept in memory: the allocator holds them
    kept = eng.stats()["counters"]["cpu_bytes_live"]
    j.set_sink_table(s_h, t_h, keep_rows=False)
    j.set_sink_table(s_r, t_r, keep_rows=False)
    eng.run(g, [j], 5, 5, db.path)
    assert eng.stats()["counters"]["cpu_bytes_live"] <= kept   # (process-wide counter: other tests' objects may live)
    with pytest.raises(E.EngineError):
        j.output_row(s_h, 0)                      # the rows live in the table now
    db.commit_job_table(t_h, j)
    db.commit_job_table(t_r, j)
    assert sorted(db.tables()) == ["hists", "small"]
    info = db.table_info("small")
    assert (info["rows"], info["items"], info["job_id"], info["keyframes"]) == (n, 3, 3, -1)
    td = parse_ref("TableDescriptor", str(tmp_path / f"db/tables/{t_h}/descriptor.bin"))
    assert list(td.end_rows) == [5, 10, 13]
    hist, small = db.read_rows("hists", "histogram"), db.read_rows("small", "frame")
    for i in range(n):
        assert hist[i] == oracle.hist16(frames[i]).tobytes()
        assert (small[i] == oracle.resize(frames[i], 12, 8)).all()
    assert db.read_rows("hists", "index", [12]) == [struct.pack("<q", 12)]
    eng.close()
    db.close()


@pytest.mark.parametrize("mode,gop", [("pcm", 4), ("skip", 5)])
def test_ffmpeg_decodes_the_synthetic_streams_to_the_source_luma(tmp_path, mode, gop):
    """Pin of the H.264 writer (and of the expected values of every decode test): an independent decoder
    (FFmpeg through cv2, raw output = the luma plane) reproduces the source planes bit for bit -- the
    streams are lossless (I_PCM) and P_Skip pictures repeat the last key picture."""
    import cv2
    n, h, w = 11, 48, 64
    stream, yuv = make_stream(9, n, h, w, gop, mode)
    path = str(tmp_path / "s.h264")
    open(path, "wb").write(stream)
    cap = cv2.VideoCapture(path)
    if not cap.set(cv2.CAP_PROP_CONVERT_RGB, 0):
        pytest.skip("this OpenCV build cannot return undecorated decoder output")
    for i in range(n):
        ok, y = cap.read()
        assert ok and y.shape == (h, w)
        src = i if mode == "pcm" else i // gop
        assert (y == yuv[src, :h * w].reshape(h, w)).all(), i
    assert not cap.read()[0]


def _ingest_many(args):
    root, prefix, stream = args
    db = E.Database(root)
    for i in range(6):
        db.ingest_h264(f"{prefix}_{i}", stream)
    names = db.tables()
    db.close()
    return len(names)


def test_processes_sharing_one_database_do_not_lose_tables(tmp_path):
    """One rank per GPU may write into the same database directory: catalogue updates are
    read-modify-write under a file lock (the reference serialises them in its master)."""
    import multiprocessing as mp
    stream, _ = make_stream(12, 4, 48, 64, 2)
    root = str(tmp_path / "db")
    E.Database(root).close()
    with mp.get_context("spawn").Pool(3) as pool:
        pool.map(_ingest_many, [(root, p, stream) for p in ("a", "b", "c")])
    db = E.Database(root)
    names = db.tables()
    assert len(names) == 18 and len({db.table_info(n)["id"] for n in names}) == 18
    meta = parse_ref("DatabaseDescriptor", os.path.join(root, "db_metadata.bin"))
    assert meta.next_table_id == 18 and all(t.committed for t in meta.tables)
    db.close()


def test_many_tables_reserved_committed_and_deleted_under_one_catalogue_update(tmp_path):
    """scn_db_new_tables / scn_db_commit_job_tables / scn_db_delete_tables: the per-table calls, batched so a job
    list with many output streams rewrites db_metadata.bin once (reference: the master writes the descriptors of
    all of a bulk job's tables in one go, master.cpp new-job path)."""
    n, h, w = 7, 16, 24
    frames = np.stack([synth.rand_frame(500 + i, h, w) for i in range(n)])
    db = E.Database(str(tmp_path / "db"))
    eng = E.Engine(gpus=[], cpu_instances=1)
    g = E.Graph()
    src = g.add_source(True)
    hs = g.add_op("TestHistogramOracle", [(src, "frame")])
    s_h = g.add_sink((hs, "histogram"))
    ids = db.new_tables([(f"h{k}", "histogram", False, "Histogram", k) for k in range(3)])
    assert len(set(ids)) == 3 and db.tables() == []
    with pytest.raises(E.EngineError):
        db.new_tables([("fresh", "c", False, "", 0), ("h1", "c", False, "", 0)])   # all or nothing
    meta = parse_ref("DatabaseDescriptor", os.path.join(db.path, "db_metadata.bin"))
    assert sorted(t.name for t in meta.tables) == ["h0", "h1", "h2"]
    jobs = []
    for k in range(3):
        j = E.Job()
        j.bind_source(src, eng.add_raw_frames(frames[k:k + 4]))
        j.set_sink_table(s_h, ids[k], keep_rows=False)
        jobs.append(j)
    eng.run(g, jobs, 3, 3, db.path)
    db.commit_job_tables(list(zip(ids, jobs)))
    assert sorted(db.tables()) == ["h0", "h1", "h2"]
    for k in range(3):
        rows = db.read_rows(f"h{k}", "histogram")
        assert [r for r in rows] == [oracle.hist16(frames[k + i]).tobytes() for i in range(4)]
        assert db.table_info(f"h{k}")["job_id"] == k
    with pytest.raises(E.EngineError):
        db.delete_tables(["h0", "missing", "h2"])
    assert db.tables() == ["h1", "h2"]            # tables named before the missing one are gone, as with single calls
    db.delete_tables(["h1", "h2"])
    assert db.tables() == [] and not os.path.exists(os.path.join(db.path, "tables", str(ids[1])))
    eng.close()
    db.close()


def test_committing_a_hundred_tables_writes_every_descriptor(tmp_path):
    """commit_tables writes the descriptors of a long table list from several threads (>= 64 tables)."""
    n_tables, n = 100, 5
    frames = np.stack([synth.rand_frame(700 + i, 8, 12) for i in range(n)])
    db = E.Database(str(tmp_path / "db"))
    eng = E.Engine(gpus=[], cpu_instances=2)
    g = E.Graph()
    src = g.add_source(True)
    s_h = g.add_sink((g.add_op("TestHistogramOracle", [(src, "frame")]), "histogram"))
    ids = db.new_tables([(f"t{k:03d}", "histogram", False, "Histogram", k) for k in range(n_tables)])
    sid = eng.add_raw_frames(frames)
    jobs = []
    for k in range(n_tables):
        j = E.Job()
        j.bind_source(src, sid)
        j.set_sink_table(s_h, ids[k], keep_rows=False)
        jobs.append(j)
    eng.run(g, jobs, 5, 5, db.path)
    db.commit_job_tables(list(zip(ids, jobs)))
    assert len(db.tables()) == n_tables
    for k in (0, 37, 63, 64, 99):
        td = parse_ref("TableDescriptor", str(tmp_path / f"db/tables/{ids[k]}/descriptor.bin"))
        assert td.name == f"t{k:03d}" and list(td.end_rows) == [n] and td.job_id == k
        assert db.read_rows(f"t{k:03d}", "histogram")[n - 1] == oracle.hist16(frames[n - 1]).tobytes()
    eng.close()
    db.close()


# ---- re-layout of an .mp4 in Python: the shapes other muxers produce (several chunks, 64-bit chunk
# offsets, moov in front of mdat = "faststart"), to exercise the demuxer's table walking
def _boxes(buf, start, end):
    out, off = [], start
    while off + 8 <= end:
        size = struct.unpack(">I", buf[off:off + 4])[0]
        hdr = 8
        if size == 1:
            size, hdr = struct.unpack(">Q", buf[off + 8:off + 16])[0], 16
        if size < hdr or off + size > end:
            break
        out.append((buf[off + 4:off + 8], off, hdr, size))
        off += size
    return out


def _box(kind, payload):
    return struct.pack(">I", len(payload) + 8) + kind + payload


def _relayout(mp4, chunk_sizes, use_co64, faststart):
    """-> new file bytes with the samples grouped into chunks of chunk_sizes samples."""
    top = {k: (o, h, s) for k, o, h, s in _boxes(mp4, 0, len(mp4))}
    ftyp = mp4[top[b"ftyp"][0]:top[b"ftyp"][0] + top[b"ftyp"][2]]
    mo, mh, ms = top[b"mdat"]
    mdat_payload = mp4[mo + mh:mo + ms]
    moov_o, moov_h, moov_s = top[b"moov"]
    i = mp4.find(b"stsz", moov_o)
    n = struct.unpack(">I", mp4[i + 12:i + 16])[0]
    sizes = struct.unpack(f">{n}I", mp4[i + 16:i + 16 + 4 * n])
    assert sum(chunk_sizes) == n

    def table(offsets):
        return _box(b"co64" if use_co64 else b"stco", struct.pack(">II", 0, len(offsets)) +
                    b"".join(struct.pack(">Q" if use_co64 else ">I", o) for o in offsets))

    def rebuild(off, hdr, size, chunk_table):
        kind = mp4[off + 4:off + 8]
        if kind in (b"moov", b"trak", b"mdia", b"minf", b"stbl"):
            return _box(kind, b"".join(rebuild(o, h, sz, chunk_table) for _, o, h, sz in _boxes(mp4, off + hdr, off + size)))
        if kind == b"stsc":
            runs, prev = [], None
            for c, cnt in enumerate(chunk_sizes):
                if cnt != prev:
                    runs.append((c + 1, cnt, 1))
                    prev = cnt
            return _box(b"stsc", struct.pack(">II", 0, len(runs)) + b"".join(struct.pack(">III", *r) for r in runs))
        if kind in (b"stco", b"co64"):
            return chunk_table
        return mp4[off:off + size]

    moov_len = len(rebuild(moov_o, moov_h, moov_s, table([0] * len(chunk_sizes))))
    pos, s, offs = len(ftyp) + (moov_len if faststart else 0) + 8, 0, []
    for c in chunk_sizes:
        offs.append(pos)
        pos += sum(sizes[s:s + c])
        s += c
    moov = rebuild(moov_o, moov_h, moov_s, table(offs))
    assert len(moov) == moov_len
    mdat = struct.pack(">I", len(mdat_payload) + 8) + b"mdat" + mdat_payload
    return ftyp + (moov + mdat if faststart else mdat + moov)


@pytest.mark.parametrize("chunks,co64,faststart", [([5, 5, 2], False, False), ([1] * 12, True, False),
                                                   ([4, 4, 4], False, True), ([7, 3, 1, 1], True, True)])
def test_demuxer_walks_the_chunk_tables_other_muxers_write(tmp_path, chunks, co64, faststart):
    stream, _ = make_stream(13, 12, 48, 64, 4)
    variant = _relayout(E.mp4_mux(stream, 25, 1), chunks, co64, faststart)
    back, info = E.mp4_demux(variant)
    assert info["samples"] == 12 and info["sync_samples"] == 3 and back == stream
    path = str(tmp_path / "v.mp4")
    open(path, "wb").write(variant)
    frames = cv2_frames(path)                      # FFmpeg agrees that the re-laid-out file is valid
    raw = str(tmp_path / "v.h264")
    open(raw, "wb").write(stream)
    want = cv2_frames(raw)
    assert len(frames) == 12 and all((a == b).all() for a, b in zip(frames, want))


@pytest.mark.parametrize("container", ["mp4", "h264"])
def test_inplace_ingest_keeps_the_bitstream_where_it_is(tmp_path, container):
    """Reference ingest `inplace` (ingest.cpp:175-215, test tables 'test1_inplace'): no copy of the
    bitstream in the database; the table binds by reading the original file again, notices when that
    file changed, and deleting the table leaves the file alone."""
    stream, _ = make_stream(12, 10, 48, 64, 5)
    path = str(tmp_path / ("clip." + container))
    open(path, "wb").write(E.mp4_mux(stream, 30, 1) if container == "mp4" else stream)
    root = str(tmp_path / "db")
    db = E.Database(root)
    db.ingest_video("copied", path)
    db.ingest_video("inplace", path, inplace=True)
    assert os.path.exists(os.path.join(root, "tables/0/1_0.bin"))
    assert not os.path.exists(os.path.join(root, "tables/1/1_0.bin"))
    a = parse_ref("VideoDescriptor", os.path.join(root, "tables/0/1_0_video_metadata.bin"))
    b = parse_ref("VideoDescriptor", os.path.join(root, "tables/1/1_0_video_metadata.bin"))
    assert b.inplace and not a.inplace and b.data_path == os.path.realpath(path)
    assert list(a.sample_offsets) == list(b.sample_offsets) and list(a.keyframe_indices) == list(b.keyframe_indices)
    assert db.table_info("inplace")["rows"] == 10 and db.table_info("inplace")["keyframes"] == 2
    eng = E.Engine(gpus=[], cpu_instances=1)
    sid = db.add_video_stream(eng, "inplace")
    assert eng.stream_rows(sid) == 10 and eng.stream_info(sid)["bytes"] == len(stream)
    out = str(tmp_path / "back.mp4")
    db.export_mp4("inplace", out, 30, 1)
    assert E.mp4_demux(open(out, "rb").read())[0] == stream
    # the file changes under the table
    other, _ = make_stream(13, 8, 48, 64, 4)
    open(path, "wb").write(E.mp4_mux(other, 30, 1) if container == "mp4" else other)
    with pytest.raises(E.EngineError, match="changed since table inplace was ingested in place"):
        db.add_video_stream(eng, "inplace")
    os.remove(path)
    with pytest.raises(E.EngineError, match="ingested in place"):
        db.add_video_stream(eng, "inplace")
    db.delete_table("inplace")
    assert db.tables() == ["copied"] and eng.stream_rows(db.add_video_stream(eng, "copied")) == 10
    eng.close()
    db.close()


def test_product_stream_writer_equals_the_independent_numpy_writer():
    """scn_h264_synth ("skip" shape: I_PCM IDR + P_Skip) against oracle/h264_writer.py, byte for byte --
    including cropped sizes and payloads that need emulation-prevention bytes.  The benchmark's CPU
    reference arm builds its clips with the numpy writer (it must not load product libraries)."""
    from oracle import h264_writer
    for (w, h, gop, frames, k) in [(64, 48, 4, 10, 3), (30, 22, 3, 7, 3), (128, 96, 1, 3, 3), (320, 240, 5, 11, 3)]:
        rng = np.random.default_rng(w * 7 + h)
        yuv = rng.integers(0, 256, (k, w * h * 3 // 2), dtype=np.uint8)
        yuv[0, :300] = 0            # runs of zeros: 00 00 0x patterns
        yuv[1, 100:140] = [0, 0, 3, 0] * 10
        assert E.h264_synth(yuv, w, h, gop=gop, non_key="skip", frames=frames) == \
            h264_writer.h264_synth_skip(yuv, w, h, gop=gop, frames=frames), (w, h, gop, frames)


@pytest.mark.parametrize("w,h,n,gop,mv", [(64, 48, 9, 4, (2, -2)), (128, 96, 10, 5, (-4, 6)), (1920, 1080, 4, 3, (2, -2)),
                                          (30, 22, 5, 2, (0, 2))])
def test_ffmpeg_decodes_the_cavlc_motion_streams_to_the_writers_model(tmp_path, w, h, n, gop, mv):
    """Pin of scanner_b200/synth_h264.py (Intra16x16 + CAVLC key pictures, motion-compensated P pictures): an
    independent decoder (FFmpeg through cv2, raw output = the luma plane) outputs exactly the pictures the
    writer's integer model predicts; the index built from the stream has the right samples and keyframes."""
    import cv2
    from scanner_b200 import synth_h264
    stream, expect = synth_h264.write(w, h, n, gop=gop, seed=w, mv=mv)
    path = str(tmp_path / "s.h264")
    open(path, "wb").write(stream)
    cap = cv2.VideoCapture(path)
    if not cap.set(cv2.CAP_PROP_CONVERT_RGB, 0):
        pytest.skip("this OpenCV build cannot return undecorated decoder output")
    for i in range(n):
        ok, y = cap.read()
        assert ok and y.shape == (h, w)
        assert (y == expect[i, :h * w].reshape(h, w)).all(), i
    assert not cap.read()[0]
    eng = E.Engine(gpus=[], cpu_instances=1)
    sid = eng.add_h264(stream)
    info = eng.stream_info(sid)
    assert (info["width"], info["height"], info["keyframes"]) == (w, h, (n + gop - 1) // gop) and eng.stream_rows(sid) == n
    eng.close()
    # bitrate of a real encoder, not of PCM: well under a tenth of the raw size
    assert len(stream) < n * w * h * 3 // 2 // 10
