"""ImageEncoder / ImageDecoder (reference scanner/util/image_encoder.cpp:8-131; tutorial
05_sources_sinks.py:41-47) -- CPU kernels like upstream.  OpenCV's PNG codec is the independent
reader and writer: what we encode it must decode to the same pixels, what it encodes we must decode."""
import numpy as np
import pytest

from scanner_b200 import engine as E
from scanner_b200 import protolite


@pytest.fixture(scope="module", autouse=True)
def stdlib():
    E.load_stdlib()


def run_encoder(frames, fmt=None):
    eng = E.Engine(gpus=[], cpu_instances=2)
    g = E.Graph()
    src = g.add_source(True)
    args = protolite.encode(protolite.parse_proto("message A { string format = 1; }")["A"], {"format": fmt}) if fmt else b""
    enc = g.add_op("ImageEncoder", [(src, "frame")], args=args)
    sink = g.add_sink((enc, "img"))
    j = E.Job()
    j.bind_source(src, eng.add_raw_frames(frames))
    eng.run(g, [j], 3, 6)
    out = [j.output_row(sink, i) for i in range(len(frames))]
    eng.close()
    return out


def run_decoder(blobs):
    eng = E.Engine(gpus=[], cpu_instances=2)
    g = E.Graph()
    src = g.add_source(False)
    dec = g.add_op("ImageDecoder", [(src, "column")])
    sink = g.add_sink((dec, "frame"))
    j = E.Job()
    j.bind_source(src, eng.add_bytes(blobs))
    eng.run(g, [j], 2, 4)
    out = [j.output_row(sink, i) for i in range(len(blobs))]
    eng.close()
    return out


@pytest.mark.parametrize("shape,dtype", [((7, 40, 56, 3), np.uint8), ((4, 33, 21, 1), np.uint8), ((3, 16, 24, 4), np.uint8),
                                         ((3, 20, 30, 3), np.uint16), ((2, 9, 5, 2), np.uint8)])
def test_encoder_output_is_read_back_by_opencv_and_by_our_decoder(shape, dtype):
    import cv2
    rng = np.random.default_rng(5)
    smooth = np.linspace(0, np.iinfo(dtype).max, shape[1] * shape[2]).reshape(1, shape[1], shape[2], 1)
    frames = ((smooth + rng.integers(0, 40, shape)) % (np.iinfo(dtype).max + 1)).astype(dtype)
    pngs = run_encoder(frames, "png")
    for f, blob in zip(frames, pngs):
        assert blob[:8] == b"\x89PNG\r\n\x1a\n"
        got = cv2.imdecode(np.frombuffer(blob, np.uint8), cv2.IMREAD_UNCHANGED)
        if shape[3] == 1:
            assert (got == f[:, :, 0]).all()
        elif shape[3] == 3:
            assert (got[:, :, ::-1] == f).all()          # OpenCV hands back BGR
        elif shape[3] == 4:
            assert (got[:, :, [2, 1, 0, 3]] == f).all()
        assert len(blob) < f.nbytes * 1.2 + 200          # compressed, not merely stored (+ chunk overhead)
    back = run_decoder(pngs)
    for f, b in zip(frames, back):
        assert b.dtype == dtype and b.shape == f.shape and (b == f).all()


def test_decoder_reads_opencv_written_png():
    import cv2
    rng = np.random.default_rng(6)
    imgs = [rng.integers(0, 256, (31, 47, 3), dtype=np.uint8), rng.integers(0, 256, (12, 12), dtype=np.uint8),
            rng.integers(0, 65536, (10, 14, 3), dtype=np.uint16)]
    blobs = [cv2.imencode(".png", im)[1].tobytes() for im in imgs]
    out = run_decoder(blobs)
    assert (out[0] == imgs[0][:, :, ::-1]).all()         # file order is RGB
    assert (out[1][:, :, 0] == imgs[1]).all()
    assert (out[2] == imgs[2][:, :, ::-1]).all() and out[2].dtype == np.uint16


def test_encoder_validates_its_format_argument():
    with pytest.raises(E.EngineError, match="Valid types are: png"):
        run_encoder(np.zeros((1, 4, 4, 3), np.uint8), "jpg")
    assert run_encoder(np.zeros((1, 4, 4, 3), np.uint8))[0][:4] == b"\x89PNG"   # default format


def test_undecodable_image_fails_the_run_not_the_process():
    """Bad DATA is reported through scanner::report_kernel_error: the run ends with the reason and
    the row, the engine (and the interpreter it lives in) stays up.  The reference aborts the worker."""
    good = run_encoder(np.full((1, 6, 5, 3), 7, np.uint8))[0]
    broken = bytearray(good)
    broken[40] ^= 0xFF  # inside the IDAT chunk: CRC mismatch
    with pytest.raises(E.EngineError, match=r"CRC mismatch \(row 1\)"):
        run_decoder([good, bytes(broken), good])
    with pytest.raises(E.EngineError, match="not a PNG stream"):
        run_decoder([b"\xff\xd8\xff\xe0 this is a JPEG header, the CPU kernel reads PNG only"])
    out = run_decoder([good, good])  # still working
    assert len(out) == 2 and (out[1] == 7).all()
