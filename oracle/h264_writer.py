"""Pure-numpy writer of the synthetic H.264 streams the benchmarks decode -- TEST INFRASTRUCTURE.

The product has its own writer (scanner_b200/csrc/engine/h264.cpp write_ipcm_stream, reached through
scn_h264_synth).  This is an independent statement of the same bitstream (H.264 7.3: SPS, PPS, IDR slices
of I_PCM macroblocks, P slices that are one mb_skip_run) used for two things:
  * `bench.py --impl reference` builds its clips with it, so the CPU reference arm never loads a
    product library;
  * tests/test_storage_cpu.py holds the product writer to it byte for byte.
Only the "skip" shape is written here: an IDR of I_PCM macroblocks every `gop` pictures, the pictures in
between P_Skip (a repeat of the IDR), Baseline profile, POC type 2, deblocking off.
"""
import re

import numpy as np


class _Bits:
    def __init__(self):
        self.bits = []

    def u(self, n, v):
        self.bits.extend((v >> i) & 1 for i in range(n - 1, -1, -1))

    def ue(self, v):
        k = v + 1
        n = k.bit_length() - 1
        self.u(n, 0)
        self.u(n + 1, k)

    def se(self, v):
        self.ue(2 * v - 1 if v > 0 else -2 * v)

    def align_zero(self):
        while len(self.bits) & 7:
            self.bits.append(0)

    def trailing(self):
        self.bits.append(1)
        self.align_zero()

    def tobytes(self):
        assert len(self.bits) % 8 == 0
        return np.packbits(np.array(self.bits, np.uint8)).tobytes()


_EPB = re.compile(b"\x00\x00(?=[\x00-\x03])")


def _nal(ref_idc, nal_type, rbsp):
    """start code + header byte + RBSP with emulation prevention bytes (7.4.1)"""
    return b"\x00\x00\x00\x01" + bytes([(ref_idc << 5) | nal_type]) + _EPB.sub(b"\x00\x00\x03", rbsp)


def _sps(width, height):
    wmb, hmb = (width + 15) // 16, (height + 15) // 16
    b = _Bits()
    b.u(8, 66)      # profile_idc Baseline
    b.u(8, 0xC0)    # constraint_set0/1
    b.u(8, 51)      # level_idc
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
    b.trailing()
    return b.tobytes()


def _pps():
    b = _Bits()
    b.ue(0)         # pic_parameter_set_id
    b.ue(0)         # seq_parameter_set_id
    b.u(1, 0)       # entropy_coding_mode_flag (CAVLC)
    b.u(1, 0)       # bottom_field_pic_order_in_frame_present_flag
    b.ue(0)         # num_slice_groups_minus1
    b.ue(0)         # num_ref_idx_l0_default_active_minus1
    b.ue(0)         # num_ref_idx_l1_default_active_minus1
    b.u(1, 0)       # weighted_pred_flag
    b.u(2, 0)       # weighted_bipred_idc
    b.se(0)         # pic_init_qp_minus26
    b.se(0)         # pic_init_qs_minus26
    b.se(0)         # chroma_qp_index_offset
    b.u(1, 1)       # deblocking_filter_control_present_flag
    b.u(1, 0)       # constrained_intra_pred_flag
    b.u(1, 0)       # redundant_pic_cnt_present_flag
    b.trailing()
    return b.tobytes()


def _slice_header(key, in_gop, idr_id):
    b = _Bits()
    b.ue(0)                    # first_mb_in_slice
    b.ue(7 if key else 5)      # slice_type I / P
    b.ue(0)                    # pic_parameter_set_id
    b.u(4, in_gop % 16)        # frame_num
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
    return b


def _pad(plane, rows, cols):
    h, w = plane.shape
    return np.pad(plane, ((0, rows - h), (0, cols - w)), mode="edge")


def h264_synth_skip(yuv, width, height, gop=30, frames=None):
    """yuv: (k, w*h*3/2) uint8 I420 pictures, one per GOP.  Returns the Annex-B stream (bytes) of `frames`
    pictures: picture f is an IDR carrying yuv[f // gop] when f % gop == 0, a P_Skip repeat otherwise."""
    yuv = np.ascontiguousarray(yuv, np.uint8).reshape(len(yuv), -1)
    assert yuv.shape[1] == width * height * 3 // 2 and width % 2 == 0 and height % 2 == 0
    n = frames if frames is not None else len(yuv) * gop
    wmb, hmb = (width + 15) // 16, (height + 15) // 16
    sps, pps = _nal(3, 7, _sps(width, height)), _nal(3, 8, _pps())
    ysz, csz = width * height, width * height // 4
    out, idr_id = [], 0
    for f in range(n):
        in_gop = f % gop
        key = in_gop == 0
        b = _slice_header(key, in_gop, idr_id)
        if not key:
            b.ue(wmb * hmb)    # mb_skip_run covering the picture
            b.trailing()
            out.append(_nal(2, 1, b.tobytes()))
            continue
        idr_id += 1
        src = yuv[f // gop]
        y = _pad(src[:ysz].reshape(height, width), hmb * 16, wmb * 16)
        u = _pad(src[ysz:ysz + csz].reshape(height // 2, width // 2), hmb * 8, wmb * 8)
        v = _pad(src[ysz + csz:].reshape(height // 2, width // 2), hmb * 8, wmb * 8)
        mbs = np.empty((hmb * wmb, 386), np.uint8)
        # every macroblock after the first starts byte aligned: ue(25) = 000011010, then pcm_alignment_zero_bits
        mbs[:, 0], mbs[:, 1] = 0x0D, 0x00
        mbs[:, 2:258] = y.reshape(hmb, 16, wmb, 16).transpose(0, 2, 1, 3).reshape(-1, 256)
        mbs[:, 258:322] = u.reshape(hmb, 8, wmb, 8).transpose(0, 2, 1, 3).reshape(-1, 64)
        mbs[:, 322:386] = v.reshape(hmb, 8, wmb, 8).transpose(0, 2, 1, 3).reshape(-1, 64)
        b.ue(25)               # mb_type I_PCM of the first macroblock, wherever the header ended
        b.align_zero()
        rbsp = b.tobytes() + mbs.reshape(-1)[2:].tobytes() + b"\x80"   # rbsp_trailing_bits
        out.append(sps + pps + _nal(3, 5, rbsp))
    return b"".join(out)
