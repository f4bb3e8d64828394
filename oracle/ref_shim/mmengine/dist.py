"""mmengine.dist.all_gather as used at cmunet_head.py:19 (test-only shim): list of per-rank tensors."""
import torch
import torch.distributed as td


def all_gather(data, group=None):
    if not (td.is_available() and td.is_initialized()) or td.get_world_size(group) == 1:
        return [data]
    out = [torch.empty_like(data) for _ in range(td.get_world_size(group))]
    td.all_gather(out, data.contiguous(), group=group)
    return out
