"""Synthetic H.264 with a realistic bitrate: Intra16x16 + CAVLC key pictures, motion-compensated P pictures.

There is no encoder offline (hard part 1), and the lossless writer of the engine
(I_PCM key pictures + P_Skip, ~105 KB per 1080p frame) exercises neither entropy decoding nor intra
prediction nor motion compensation.  This pure-numpy writer produces streams a real encoder could have
produced, small enough to reason about bit by bit (H.264 (04/2017) clause numbers in the comments):

  * key pictures: every macroblock Intra16x16 with DC prediction (8.3.3.3), its only residual the DC
    coefficient of the luma DC block and of the two chroma DC blocks, CAVLC coded (9.2.1-9.2.3:
    coeff_token tables for nC in [0,2) and nC = -1, level prefix, total_zeros); QP 28, where the
    dequantised DC level is exactly the sample offset (luma: +L, chroma: +2c);
  * the pictures in between: P slices, every macroblock P_L0_16x16 with one global integer motion
    vector (even components, so chroma moves by whole samples) and no residual: a pan of the previous
    picture with the edge samples repeated (8.4.2.2, reference picture padding);
  * Baseline profile, POC type 2, deblocking off (the expected pictures are then exact integer models).

`write()` returns the Annex-B bytes AND the pictures a conformant decoder must output (I420, display size),
computed by the same integer model; tests hold FFmpeg (cv2) and NVDEC to them bit for bit.  ~20 KB per key
picture and ~5 KB per P picture at 1080p: 1-2 Mbit/s at 30 fps.

No native library is involved: the CPU reference arm of bench.py may import this module.
"""
import re

import numpy as np

_EPB = re.compile(b"\x00\x00(?=[\x00-\x03])")
QP = 28  # (QP % 6 == 4, QP / 6 == 4): LevelScale(4,0,0) = 16 * 16 = 256


class _Codes:
    """A growing list of (value, nbits) codes, packed MSB first."""

    def __init__(self):
        self.v, self.n = [], []

    def u(self, n, v):
        self.v.append(v)
        self.n.append(n)

    def ue(self, v):
        k = v + 1
        b = k.bit_length()
        self.u(2 * b - 1, k)

    def se(self, v):
        self.ue(2 * v - 1 if v > 0 else -2 * v)

    def extend(self, vals, lens):
        self.v.extend(vals)
        self.n.extend(lens)

    def tobytes(self, trailing=True):
        v, n = np.asarray(self.v, np.uint64), np.asarray(self.n, np.int64)
        total = int(n.sum())
        idx = np.repeat(np.arange(len(n)), n)                       # code index of every bit
        pos = np.arange(total) - np.repeat(np.cumsum(n) - n, n)     # bit position inside its code
        bits = ((v[idx] >> (n[idx] - 1 - pos).astype(np.uint64)) & np.uint64(1)).astype(np.uint8)
        if trailing:
            pad = (-(total + 1)) % 8
            bits = np.concatenate([bits, np.array([1] + [0] * pad, np.uint8)])
        return np.packbits(bits).tobytes()


def _nal(ref_idc, nal_type, rbsp):
    return b"\x00\x00\x00\x01" + bytes([(ref_idc << 5) | nal_type]) + _EPB.sub(b"\x00\x00\x03", rbsp)


def _sps(width, height):
    wmb, hmb = (width + 15) // 16, (height + 15) // 16
    b = _Codes()
    b.u(8, 66)      # profile_idc Baseline
    b.u(8, 0xC0)    # constraint_set0/1
    b.u(8, 41)      # level_idc 4.1
    b.ue(0)         # seq_parameter_set_id
    b.ue(0)         # log2_max_frame_num_minus4
    b.ue(2)         # pic_order_cnt_type 2
    b.ue(1)         # max_num_ref_frames
    b.u(1, 0)       # gaps_in_frame_num_value_allowed_flag
    b.ue(wmb - 1)
    b.ue(hmb - 1)
    b.u(1, 1)       # frame_mbs_only_flag
    b.u(1, 1)       # direct_8x8_inference_flag
    crop = wmb * 16 != width or hmb * 16 != height
    b.u(1, 1 if crop else 0)
    if crop:
        b.ue(0)
        b.ue((wmb * 16 - width) // 2)
        b.ue(0)
        b.ue((hmb * 16 - height) // 2)
    b.u(1, 0)       # vui_parameters_present_flag
    return b.tobytes()


def _pps():
    b = _Codes()
    b.ue(0)         # pic_parameter_set_id
    b.ue(0)         # seq_parameter_set_id
    b.u(1, 0)       # entropy_coding_mode_flag: CAVLC
    b.u(1, 0)       # bottom_field_pic_order_in_frame_present_flag
    b.ue(0)         # num_slice_groups_minus1
    b.ue(0)         # num_ref_idx_l0_default_active_minus1
    b.ue(0)         # num_ref_idx_l1_default_active_minus1
    b.u(1, 0)       # weighted_pred_flag
    b.u(2, 0)       # weighted_bipred_idc
    b.se(QP - 26)   # pic_init_qp_minus26
    b.se(0)         # pic_init_qs_minus26
    b.se(0)         # chroma_qp_index_offset
    b.u(1, 1)       # deblocking_filter_control_present_flag
    b.u(1, 0)       # constrained_intra_pred_flag
    b.u(1, 0)       # redundant_pic_cnt_present_flag
    return b.tobytes()


def _slice_header(b, key, frame_num, idr_id):
    b.ue(0)                    # first_mb_in_slice
    b.ue(7 if key else 5)      # slice_type I / P (all slices of the picture)
    b.ue(0)                    # pic_parameter_set_id
    b.u(4, frame_num % 16)     # frame_num
    if key:
        b.ue(idr_id & 0xFFFF)  # idr_pic_id
    else:
        b.u(1, 0)              # num_ref_idx_active_override_flag
        b.u(1, 0)              # ref_pic_list_modification_flag_l0
    if key:
        b.u(1, 0)              # no_output_of_prior_pics_flag
        b.u(1, 0)              # long_term_reference_flag
    else:
        b.u(1, 0)              # adaptive_ref_pic_marking_mode_flag
    b.se(0)                    # slice_qp_delta
    b.ue(1)                    # disable_deblocking_filter_idc = 1


# ---- CAVLC for a block whose only possibly non-zero coefficient is the first one (9.2) -------------
def _dc_only_block(b, level, chroma_dc):
    """residual_block_cavlc of a block with coefficients (level, 0, 0, ...); |level| <= 7.
    chroma_dc: the 2x2 chroma DC block (nC = -1, Table 9-5 last column); otherwise nC = 0 (first column)."""
    if level == 0:
        b.u(2, 1) if chroma_dc else b.u(1, 1)           # coeff_token TotalCoeff 0: '01' / '1'
        return
    sign = 1 if level < 0 else 0
    if abs(level) == 1:
        # TotalCoeff 1, TrailingOnes 1: '1' (nC = -1) / '01' (nC 0..1); then trailing_ones_sign_flag
        b.u(1, 1) if chroma_dc else b.u(2, 1)
        b.u(1, sign)
    else:
        # TotalCoeff 1, TrailingOnes 0: '000111' (nC = -1) / '000101' (nC 0..1)
        b.u(6, 7) if chroma_dc else b.u(6, 5)
        # level (9.2.2.1): suffixLength 0; levelCode = 2|l|-2 (+1 if negative), minus 2 because TrailingOnes < 3
        code = 2 * abs(level) - 2 + sign - 2
        assert 0 <= code < 14
        b.u(code + 1, 1)                                # level_prefix: `code` zeros then a one, no suffix
    b.u(1, 1)                                           # total_zeros = 0 (Tables 9-7 / 9-9a, tzVlcIndex 1): '1'


def _dc_pred(top, left):
    """Intra DC prediction from a flat row above / flat column beside (either may be None)."""
    if top is not None and left is not None:
        return (top + left + 1) >> 1
    if top is not None:
        return top
    if left is not None:
        return left
    return 128


def _key_picture(b, wmb, hmb, rng, target):
    """Codes one I picture into `b`; returns the reconstructed planes (coded size).  target: (Y, Cb, Cr)
    smooth fields the DC levels steer the reconstruction towards."""
    ty, tcb, tcr = target
    luma = np.zeros((hmb, wmb), np.int32)           # every luma macroblock is flat
    chroma = np.zeros((2, hmb * 2, wmb * 2), np.int32)  # chroma at 4x4-block granularity (2x2 blocks per MB)
    noise = rng.integers(-2, 3, (hmb, wmb, 3))
    for my in range(hmb):
        for mx in range(wmb):
            top = int(luma[my - 1, mx]) if my else None
            left = int(luma[my, mx - 1]) if mx else None
            pred = _dc_pred(top, left)
            want = int(ty[my, mx]) + int(noise[my, mx, 0])
            lev = max(-7, min(7, want - pred))
            lev = max(-pred, min(255 - pred, lev))   # keep the sample in range: no clipping to model
            luma[my, mx] = pred + lev                # QP 28: dequantised DC == level (module docstring)
            # chroma DC prediction per 4x4 block (8.3.4): block (0,0) and (1,1) use both neighbours, (1,0) prefers
            # the row above, (0,1) the column beside
            levs = []
            for p, tgt in ((0, tcb), (1, tcr)):
                c = chroma[p]
                by, bx = 2 * my, 2 * mx
                t0 = int(c[by - 1, bx]) if my else None       # block above (0,0): the MB above's bottom-left block
                t1 = int(c[by - 1, bx + 1]) if my else None
                l0 = int(c[by, bx - 1]) if mx else None       # block left of (0,0): the left MB's top-right block
                l1 = int(c[by + 1, bx - 1]) if mx else None
                p00 = _dc_pred(t0, l0)
                p10 = t1 if t1 is not None else (l0 if l0 is not None else 128)   # x = 4..7, y = 0..3
                p01 = l1 if l1 is not None else (t0 if t0 is not None else 128)   # x = 0..3, y = 4..7
                p11 = _dc_pred(t1, l1)
                preds = (p00, p10, p01, p11)
                wantc = int(tgt[my, mx]) + int(noise[my, mx, 1 + p])
                cl = max(-3, min(3, (wantc - (sum(preds) >> 2)) // 2))
                cl = max(-(min(preds) // 2), min((255 - max(preds)) // 2, cl))
                c[by, bx], c[by, bx + 1], c[by + 1, bx], c[by + 1, bx + 1] = (q + 2 * cl for q in preds)
                levs.append(cl)
            b.ue(1 + 2 + 4 * 1)      # mb_type I_16x16_2_1_0: DC prediction, chroma DC coded, no luma AC
            b.ue(0)                  # intra_chroma_pred_mode DC
            b.se(0)                  # mb_qp_delta
            _dc_only_block(b, lev, False)       # Intra16x16DCLevel
            _dc_only_block(b, levs[0], True)    # ChromaDCLevel Cb
            _dc_only_block(b, levs[1], True)    # ChromaDCLevel Cr
    y = np.repeat(np.repeat(luma, 16, 0), 16, 1).astype(np.uint8)
    cb = np.repeat(np.repeat(chroma[0], 4, 0), 4, 1).astype(np.uint8)
    cr = np.repeat(np.repeat(chroma[1], 4, 0), 4, 1).astype(np.uint8)
    return y, cb, cr


def _pan(plane, dx, dy):
    """pred[y, x] = ref[clamp(y + dy), clamp(x + dx)] (8.4.2.2.1 with integer vectors)."""
    h, w = plane.shape
    ys = np.clip(np.arange(h) + dy, 0, h - 1)
    xs = np.clip(np.arange(w) + dx, 0, w - 1)
    return plane[ys][:, xs]


def _p_picture(b, wmb, hmb, ref, mv):
    dx, dy = mv
    assert dx % 2 == 0 and dy % 2 == 0, "even vectors: chroma moves by whole samples"
    n = wmb * hmb
    # macroblock 0 carries the vector (no neighbours: predictor 0); every other macroblock predicts it exactly
    # from its neighbours (8.4.1.3: one available neighbour -> that one; otherwise the median, two of three equal)
    b.ue(0)          # mb_skip_run
    b.ue(0)          # mb_type P_L0_16x16
    b.se(4 * dx)     # mvd_l0 x (quarter samples)
    b.se(4 * dy)     # mvd_l0 y
    b.ue(0)          # coded_block_pattern 0 (me(v): codeNum 0 for Inter)
    b.extend([0b11111] * (n - 1), [5] * (n - 1))   # skip_run 0, P_L0_16x16, mvd 0, mvd 0, cbp 0
    y, cb, cr = ref
    return _pan(y, dx, dy), _pan(cb, dx // 2, dy // 2), _pan(cr, dx // 2, dy // 2)


def _smooth_field(rng, hmb, wmb, lo, hi):
    """a low-frequency field at macroblock resolution"""
    gy, gx = np.mgrid[0:hmb, 0:wmb].astype(np.float64)
    f = np.zeros((hmb, wmb))
    for _ in range(4):
        fx, fy, ph = rng.uniform(0.02, 0.25), rng.uniform(0.02, 0.25), rng.uniform(0, 6.28)
        f += np.sin(fx * gx + fy * gy + ph)
    f = (f - f.min()) / max(1e-9, f.max() - f.min())
    return (lo + f * (hi - lo)).astype(np.int32)


def write(width, height, frames, gop=30, seed=0, mv=(2, -2)):
    """-> (annexb bytes, expected (frames, width*height*3/2) uint8 I420 pictures in display order)."""
    assert width % 2 == 0 and height % 2 == 0 and frames >= 1 and gop >= 1
    wmb, hmb = (width + 15) // 16, (height + 15) // 16
    rng = np.random.default_rng(seed)
    sps, pps = _nal(3, 7, _sps(width, height)), _nal(3, 8, _pps())
    out, expect = [], np.empty((frames, width * height * 3 // 2), np.uint8)
    ref, idr_id = None, 0
    for f in range(frames):
        in_gop = f % gop
        key = in_gop == 0
        b = _Codes()
        _slice_header(b, key, in_gop, idr_id)
        if key:
            idr_id += 1
            target = (_smooth_field(rng, hmb, wmb, 24, 232), _smooth_field(rng, hmb, wmb, 64, 192),
                      _smooth_field(rng, hmb, wmb, 64, 192))
            ref = _key_picture(b, wmb, hmb, rng, target)
            out.append(sps + pps + _nal(3, 5, b.tobytes()))
        else:
            ref = _p_picture(b, wmb, hmb, ref, mv)
            out.append(_nal(2, 1, b.tobytes()))
        y, cb, cr = ref
        expect[f] = np.concatenate([y[:height, :width].ravel(), cb[:height // 2, :width // 2].ravel(),
                                    cr[:height // 2, :width // 2].ravel()])
    return b"".join(out), expect
