"""ImageDecoder on the GPU: JPEG through nvJPEG into device frames, PNG through the same kernel.
JPEG decoders agree only up to IDCT / chroma-upsampling rounding, so JPEG is compared with FFmpeg-free
libjpeg (cv2.imdecode) within a stated tolerance on smooth content; PNG is exact."""
import cv2
import numpy as np
import pytest

import scanner_b200 as sp

pytestmark = pytest.mark.gpu


def _smooth(h, w, seed):
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    rng = np.random.default_rng(seed)
    a, b, c = rng.uniform(0.5, 2.0, 3)
    img = np.stack([128 + 100 * np.sin(xx / (9 * a) + seed), 128 + 100 * np.cos(yy / (7 * b)),
                    128 + 90 * np.sin((xx * c + yy) / 40)], axis=2)
    return np.clip(img, 0, 255).astype(np.uint8)


def test_image_decoder_gpu_jpeg_and_png():
    rgb = [_smooth(48, 64, 1), _smooth(77, 123, 2), _smooth(240, 320, 3)]
    gray = _smooth(40, 56, 4)[..., 0]
    blobs, want = [], []
    full = [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444]
    for k, im in enumerate(rgb):
        # 4:4:4 for the first two (odd size included): only IDCT rounding can differ between decoders;
        # the third is 4:2:0, where chroma upsampling filters differ too (looser bound below)
        ok, enc = cv2.imencode(".jpg", im[..., ::-1], [cv2.IMWRITE_JPEG_QUALITY, 95] + (full if k < 2 else []))
        assert ok
        blobs.append(enc.tobytes())
        want.append(cv2.imdecode(enc, cv2.IMREAD_COLOR)[..., ::-1])
    ok, enc = cv2.imencode(".jpg", gray, [cv2.IMWRITE_JPEG_QUALITY, 95])
    blobs.append(enc.tobytes())
    want.append(cv2.imdecode(enc, cv2.IMREAD_COLOR)[..., ::-1])
    ok, enc = cv2.imencode(".png", rgb[0][..., ::-1])
    blobs.append(enc.tobytes())
    want.append(rgb[0])

    with sp.Client(gpus=[0], instances_per_gpu=2) as sc:
        src = sp.NamedStream(sc, "encoded", rows=blobs)
        frames = sc.ops.ImageDecoder(img=sc.io.Input([src]), device=sp.DeviceType.GPU)
        hists = sc.ops.Histogram(frame=frames, device=sp.DeviceType.GPU)  # consumed on the device too
        out, hout = sp.NamedVideoStream(sc, "decoded"), sp.NamedStream(sc, "hists")
        sc.run([sc.io.Output(frames, [out]), sc.io.Output(hists, [hout])], sp.PerfParams.manual(2, 4))
        got, h = list(out.load()), list(hout.load())
    assert len(got) == len(want) == 5
    for i, (g, w) in enumerate(zip(got, want)):
        assert g.shape == w.shape and g.dtype == np.uint8, i
        assert int(np.stack(h[i]).sum()) == 3 * g.shape[0] * g.shape[1]
        err = np.abs(g.astype(np.int32) - w.astype(np.int32))
        if i == 4:
            assert err.max() == 0  # PNG is lossless
        elif i == 2:
            # 4:2:0: nvJPEG replicates chroma samples, libjpeg interpolates them ("fancy upsampling"):
            # measured mean |diff| 3.6 on this gradient image
            assert err.mean() < 6.0, (i, err.mean(), err.max())
        else:
            assert err.mean() < 1.0 and err.max() <= 4, (i, err.mean(), err.max())
