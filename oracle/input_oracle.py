"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's loader-side input chain.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product path
(efficient_slowfast_b200.frames + csrc/esf_simt.cu) never does.

Pinned against the reference's own functions in tests/test_oracle_vs_reference.py (container only) and by the
committed fixture tests/golden/frames_input.npz (made by tests/golden/make_golden_frames.py from the reference).
"""
import torch


def tensor_normalize(tensor, mean, std):
    """SlowFast/slowfast/datasets/utils.py:298-315."""
    if tensor.dtype == torch.uint8:
        tensor = tensor.float()
        tensor = tensor / 255.0
    if isinstance(mean, list):
        mean = torch.tensor(mean)
    if isinstance(std, list):
        std = torch.tensor(std)
    tensor = tensor - mean
    tensor = tensor / std
    return tensor


def pack_pathway_output(frames, alpha, single_pathway=False, reverse_input_channel=False):
    """SlowFast/slowfast/datasets/utils.py:73-112 for one clip (C, T, H, W)."""
    if reverse_input_channel:
        frames = frames[[2, 1, 0], :, :, :]
    if single_pathway:
        return [frames]
    index = torch.linspace(0, frames.shape[1] - 1, frames.shape[1] // alpha).long()
    return [torch.index_select(frames, 1, index), frames]


def frames_to_inputs(frames_u8, mean, std, alpha, single_pathway=False, reverse_input_channel=False):
    """uint8 (B, T, H, W, C) -> the model's input list, clip by clip like the loader
    (SlowFast/slowfast/datasets/kinetics.py:231-248: tensor_normalize, permute(3,0,1,2), pack_pathway_output; the
    spatial crop in between is the identity for frames already at crop size), stacked by the collate function."""
    per_clip = []
    for b in range(frames_u8.shape[0]):
        f = tensor_normalize(frames_u8[b], list(mean), list(std))
        f = f.permute(3, 0, 1, 2)
        per_clip.append(pack_pathway_output(f, alpha, single_pathway, reverse_input_channel))
    return [torch.stack([c[i] for c in per_clip]).contiguous() for i in range(len(per_clip[0]))]


def demo_window_inputs(window_u8, mean, std, num_frames, alpha=None):
    """SlowFast/tools/demo_net.py:198-224: the model inputs the demo builds from its frame buffer.  window_u8: uint8
    (L, H, W, C), oldest frame first.  tensor_normalize -> permute(3,0,1,2) -> unsqueeze(0) -> fast (or single) pathway =
    index_select(linspace(0, L-1, num_frames).long()); slow = index_select(fast, linspace(0, NF-1, NF // alpha).long())."""
    x = tensor_normalize(window_u8, list(mean), list(std)).permute(3, 0, 1, 2).unsqueeze(0)
    fast = torch.index_select(x, 2, torch.linspace(0, x.shape[2] - 1, num_frames).long())
    if alpha is None:
        return [fast]
    slow = torch.index_select(fast, 2, torch.linspace(0, fast.shape[2] - 1, fast.shape[2] // alpha).long())
    return [slow, fast]
