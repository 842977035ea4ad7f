"""Sliding-window inference over a frame stream -- the loop of the reference's `tools/demo_net.py:153-305` without its
per-window host work.

The reference keeps the last `NUM_FRAMES * SAMPLING_RATE` frames in a Python list and, once it is full, for EVERY new
frame: stacks the list into one tensor, normalises it on the host (`tensor_normalize`), permutes, picks the fast
pathway's frames with `linspace(0, L-1, NUM_FRAMES).long()` and the slow pathway's with a second `linspace` over those
(demo_net.py:198-224), copies both pathways to the GPU and runs the model; then it drops the oldest frame.

Here the window is a ring of uint8 frames that lives on the device: a new frame costs one H2D copy of H*W*C bytes into
its slot, and the two pathway selections are index arrays into the ring (`model.forward_frames(frame_index=...)`), so
normalisation, layout change and both gathers happen inside the stem-pack kernel and nothing is re-assembled.
"""
import torch

from . import runtime as rt


def window_indices(seq_len, num_frames, alpha=None):
    """demo_net.py:203-224: positions (inside the window, oldest frame = 0) of the fast / single pathway's frames, and of
    the slow pathway's frames when `alpha` is given."""
    fast = torch.linspace(0, seq_len - 1, num_frames).long()
    if alpha is None:
        return [fast]
    slow = fast.index_select(0, torch.linspace(0, fast.shape[0] - 1, fast.shape[0] // alpha).long())
    return [slow, fast]


class SlidingWindow:
    def __init__(self, model, cfg, height, width, channels=3, device=None):
        device = torch.device(device) if device is not None else next(model.parameters()).device
        if device.type != "cuda":
            raise rt.EsfError("SlidingWindow needs a CUDA device; there is no CPU fallback")
        self.model, self.device = model, device
        self.seq_len = int(cfg.DATA.NUM_FRAMES) * int(cfg.DATA.get("SAMPLING_RATE", 1) if hasattr(cfg.DATA, "get")
                                                      else cfg.DATA.SAMPLING_RATE)
        alpha = int(cfg.SLOWFAST.ALPHA) if model.num_pathways > 1 else None
        self.positions = window_indices(self.seq_len, int(cfg.DATA.NUM_FRAMES), alpha)
        self.ring = torch.zeros((1, self.seq_len, height, width, channels), dtype=torch.uint8, device=device)
        self.count = 0          # frames pushed so far

    def push(self, frame):
        """frame: uint8 (H, W, C) tensor, host (pinned for an asynchronous copy) or device.  Returns the model output
        (1, num_classes) for the window that ends with this frame, or None while the window is filling."""
        assert frame.dtype == torch.uint8 and tuple(frame.shape) == tuple(self.ring.shape[2:])
        self.ring[0, self.count % self.seq_len].copy_(frame, non_blocking=True)
        self.count += 1
        if self.count < self.seq_len:
            return None
        oldest = self.count % self.seq_len          # ring slot of window position 0
        index = [((p + oldest) % self.seq_len).to(torch.int32).to(self.device) for p in self.positions]
        return self.model.forward_frames(self.ring, frame_index=index)
