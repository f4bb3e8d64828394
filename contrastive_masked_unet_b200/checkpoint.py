"""Checkpoint hand-off between pretraining and fine-tuning.

Writes the mmengine CheckpointHook layout the reference's fine-tuning script sniffs for
(`Finetuning/train.py:262-273`: `{'meta': {'mmengine_version': ...}, 'state_dict': {'backbone.*', 'pixel_decoder.*', ...}}`)
and restates that loader: keys containing `pixel_decoder` / `backbone` are kept with the prefix stripped, `conv_last.*`
is dropped, `load_state_dict(strict=False)`.  Because the drop-in modules keep the reference's key names, a checkpoint
written here loads in the unmodified reference script and vice versa."""
import torch

MMENGINE_VERSION = '0.10.5'


def save_pretrain_checkpoint(model, path, epoch=0, iteration=0, optimizer=None):
    core = model.module if hasattr(model, 'module') else model
    ckpt = {'meta': {'mmengine_version': MMENGINE_VERSION, 'epoch': epoch, 'iter': iteration,
                     'producer': 'contrastive_masked_unet_b200'},
            'state_dict': {k: v.detach().cpu() for k, v in core.state_dict().items()}}
    if optimizer is not None and hasattr(optimizer, 'state'):
        ckpt['optimizer'] = {'step': getattr(optimizer, 'step_count', 0),
                             'state': {k: (m.detach().cpu(), v.detach().cpu()) for k, (m, v) in optimizer.state.items()}}
    torch.save(ckpt, path)
    return ckpt


def finetune_state_dict(checkpoint):
    """The key mapping of Finetuning/train.py:262-273 ("CMAE" branch) and :275-285 ("encoder only")."""
    if 'meta' in checkpoint and 'mmengine_version' in checkpoint['meta']:
        out = {}
        for key, val in checkpoint['state_dict'].items():
            if 'pixel_decoder' in key:
                out[key.replace('pixel_decoder.', '')] = val
            if 'backbone' in key:
                out[key.replace('backbone.', '')] = val
    else:
        sd = checkpoint.get('state_dict', checkpoint)
        out = {k.replace('module.', ''): v for k, v in sd.items()}
        out = {k: v for k, v in out.items() if 'down_conv' in k or 'double_conv' in k}
    out.pop('conv_last.weight', None)
    out.pop('conv_last.bias', None)
    return out


def load_pretrained_into_unet(unet, path_or_checkpoint, map_location='cpu'):
    ckpt = path_or_checkpoint
    if isinstance(ckpt, str):
        ckpt = torch.load(ckpt, map_location=map_location, weights_only=False)
    return unet.load_state_dict(finetune_state_dict(ckpt), strict=False)
