"""uint8 frame input: the loader-side chain of the reference moved behind the H2D copy.

The reference decodes a clip to uint8 (T, H, W, C) frames, then on the host: `tensor_normalize` (u8 -> float / 255,
- mean, / std; SlowFast/slowfast/datasets/utils.py:298-315), `permute(3, 0, 1, 2)` (datasets/kinetics.py:231-235),
crop, and `pack_pathway_output` (datasets/utils.py:73-112: optional channel reversal, slow pathway = frames gathered at
`linspace(0, T-1, T // ALPHA).long()`).  The result -- 4 bytes per sample, the slow frames twice -- is what crosses
PCIe.  `model.forward_frames(frames_u8)` takes the byte frames instead and the stem-pack kernel does the chain on the
device: 1/5 of the H2D bytes, bit-identical stem input (the byte -> value table is computed here with the reference's
own FP32 operations).
"""
import ctypes

import torch

from . import runtime as rt


def slow_frame_index(num_frames, alpha):
    """datasets/utils.py:93-102 -- the reference's own expression, so that float rounding of linspace matches."""
    return torch.linspace(0, num_frames - 1, num_frames // alpha).long()


def normalization_table(mean, std, channels):
    """FP32 [C][256]: tensor_normalize (datasets/utils.py:298-315) applied to every byte value, per channel."""
    u = torch.arange(256, dtype=torch.uint8).reshape(256, 1).repeat(1, channels)
    t = u.float()
    t = t / 255.0
    t = t - torch.tensor(list(mean)[:channels], dtype=torch.float32)
    t = t / torch.tensor(list(std)[:channels], dtype=torch.float32)
    return t.t().contiguous()


class FrameInput:
    """Per-(model, device) state of the frame route: look-up tables and frame indices on the device."""

    def __init__(self, cfg, device, adt, channels=3):
        data = cfg.DATA
        mean = data.get("MEAN", [0.45] * channels) if hasattr(data, "get") else data.MEAN
        std = data.get("STD", [0.225] * channels) if hasattr(data, "get") else data.STD
        reverse = bool(data.get("REVERSE_INPUT_CHANNEL", False)) if hasattr(data, "get") else False
        self.channels = channels
        lut = normalization_table(mean, std, channels)
        self.lut32 = lut.to(device)
        self.lut16 = lut.to(adt).to(device)
        src = list(range(channels))
        if reverse:
            assert channels == 3, "DATA.REVERSE_INPUT_CHANNEL needs 3 channels"
            src = [2, 1, 0]
            # the reference reverses AFTER normalising: output channel c holds source channel src[c] normalised with
            # the statistics of the SOURCE channel
            self.lut32 = self.lut32[src].contiguous()
            self.lut16 = self.lut16[src].contiguous()
        self.chan_src = (ctypes.c_int32 * channels)(*src)
        self.device = device
        self._index = {}

    def index(self, t_src, t_out, alpha):
        """device int32 frame gather for a pathway with t_out of t_src frames (None: identity)."""
        if t_out == t_src:
            return None
        key = (t_src, t_out)
        if key not in self._index:
            idx = slow_frame_index(t_src, alpha)
            assert idx.numel() == t_out, "pathway has %d frames, expected %d" % (t_out, idx.numel())
            self._index[key] = idx.to(torch.int32).to(self.device)
        return self._index[key]


def launch_frames(plan, fin, frames, alpha, frame_index=None):
    """Fill the stem inputs of `plan` from uint8 frames (B, T, H, W, C) on the current stream.  `frame_index`: optional
    list (one entry per pathway) of device int32 tensors, source frame of every pathway frame -- replaces the loader's
    rule (slow = linspace gather of the fast frames) for callers that keep frames in a ring (demo.SlidingWindow)."""
    L = rt.lib()
    s = rt.current_stream_ptr()
    B, Tsrc, H, W, C = frames.shape
    for pw, own in enumerate(plan.inputs):
        b, c, T, h, w = own.shape
        assert (b, c, h, w) == (B, C, H, W)
        if frame_index is not None:
            idx = frame_index[pw]
            assert idx.dtype == torch.int32 and idx.numel() == T and idx.device == frames.device
        else:
            idx = fin.index(Tsrc, T, alpha)
        iptr = idx.data_ptr() if idx is not None else None
        route = plan.stem_routes.get(own.data_ptr())
        if route is not None:
            xp, pitch, lpad = route
            rt.check(L.esf_stem_pack_u8(frames.data_ptr(), B, Tsrc, H, W, C, iptr, T, fin.chan_src, fin.lut16.data_ptr(),
                                        pitch, lpad, xp.data_ptr(), s), "esf_stem_pack_u8")
        else:
            rt.check(L.esf_frames_to_clip(frames.data_ptr(), B, Tsrc, H, W, C, iptr, T, fin.chan_src,
                                          fin.lut32.data_ptr(), own.data_ptr(), s), "esf_frames_to_clip")
