"""Drop-in for SlowFast/slowfast/models/build.py: MODEL_REGISTRY and build_model(cfg)."""
import torch

from .registry import Registry

MODEL_REGISTRY = Registry("MODEL")
MODEL_REGISTRY.__doc__ = "Registry for video models; the registered object is called as obj(cfg) -> nn.Module."


def build_model(cfg):
    """build.py:18-44 of the reference: look the class up by cfg.MODEL.MODEL_NAME, construct it, move it to the
    current CUDA device when NUM_GPUS >= 1 and wrap it in DistributedDataParallel when NUM_GPUS > 1."""
    assert cfg.NUM_GPUS <= torch.cuda.device_count(), "Cannot use more GPU devices than available"
    model = MODEL_REGISTRY.get(cfg.MODEL.MODEL_NAME)(cfg)
    if cfg.NUM_GPUS >= 1:
        cur_device = torch.cuda.current_device()
        model = model.cuda(device=cur_device)
    if cfg.NUM_GPUS > 1:
        model = torch.nn.parallel.DistributedDataParallel(module=model, device_ids=[cur_device],
                                                          output_device=cur_device)
    return model
