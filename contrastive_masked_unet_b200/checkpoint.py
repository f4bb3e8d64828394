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


def finetune_state_dict(checkpoint, path=None):
    """The key mappings of `load_model` in Finetuning/train.py:240-308, one branch per checkpoint flavour:
      * mmengine layout ("CMAE", :262-273): `pixel_decoder.*` / `backbone.*` with the prefix stripped;
      * MoCo-v2 Lightning checkpoint (".ckpt", :286-296): `encoder_q.*` with the prefix stripped, encoder only --
        detected by the file extension like the reference, or by the presence of `encoder_q.` keys (this package's
        `Moco_v2.state_dict()`); `encoder_k.*` entries never match a UNet key and are dropped here;
      * {'module': ...} (SparK, :250-260): `sparse_encoder.sp_cnn.` / `dense_decoder.` stripped, encoder + decoder;
      * a bare / DataParallel state dict ("encoder only", :275-285): `module.` stripped, encoder keys only.
    `conv_last.*` is always dropped (:271-272)."""
    enc_only = lambda d: {k: v for k, v in d.items() if 'down_conv' in k or 'double_conv' in k}   # noqa: E731
    sd = checkpoint.get('state_dict', None) if isinstance(checkpoint, dict) else None
    is_moco = (path is not None and str(path).endswith('.ckpt')) or \
        (sd is not None and any(k.startswith('encoder_q.') for k in sd))
    if isinstance(checkpoint, dict) and 'module' in checkpoint and isinstance(checkpoint['module'], dict):
        out = {}
        for key, val in checkpoint['module'].items():
            out[key.replace('sparse_encoder.sp_cnn.', '')] = val
            out[key.replace('dense_decoder.', '')] = val
        out = {k: v for k, v in out.items() if 'down_conv' in k or 'double_conv' in k or 'up_conv' in k}
    elif is_moco:
        out = enc_only({k.replace('encoder_q.', ''): v for k, v in sd.items() if not k.startswith('encoder_k.')})
    elif isinstance(checkpoint, dict) and 'meta' in checkpoint and 'mmengine_version' in checkpoint['meta']:
        out = {}
        for key, val in checkpoint['state_dict'].items():
            if 'pixel_decoder' in key:
                out[key.replace('pixel_decoder.', '')] = val
            if 'backbone' in key:
                out[key.replace('backbone.', '')] = val
    else:
        sd = checkpoint.get('state_dict', checkpoint)
        out = enc_only({k.replace('module.', ''): v for k, v in sd.items()})
    out.pop('conv_last.weight', None)
    out.pop('conv_last.bias', None)
    return out


def load_pretrained_into_unet(unet, path_or_checkpoint, map_location='cpu'):
    """`load_state_dict(strict=False)` like the reference -- but a checkpoint from which NOT ONE encoder tensor was
    taken (wrong flavour, unexpected prefix) raises instead of silently fine-tuning from random weights."""
    ckpt, path = path_or_checkpoint, None
    if isinstance(ckpt, str):
        path = ckpt
        ckpt = torch.load(ckpt, map_location=map_location, weights_only=False)
    mapped = finetune_state_dict(ckpt, path)
    own = unet.state_dict()
    enc_keys = [k for k in own if k.startswith('down_conv') or k.startswith('double_conv')]
    if enc_keys and not any(k in mapped for k in enc_keys):
        raise ValueError('load_pretrained_into_unet: the checkpoint holds no tensor for any encoder key of the UNet '
                         f'(first checkpoint keys after mapping: {list(mapped)[:4]}); refusing to continue from '
                         'random weights')
    return unet.load_state_dict(mapped, strict=False)


def save_moco_checkpoint(model, path, epoch=0, global_step=0):
    """Lightning-style `.ckpt` of a `Moco_v2` (what Finetuning/train.py:286-296 reads): {'state_dict': {'encoder_q.*',
    'encoder_k.*', 'queue', 'queue_ptr', ...}, 'epoch', 'global_step'}."""
    core = model.module if hasattr(model, 'module') else model
    ckpt = {'epoch': epoch, 'global_step': global_step, 'producer': 'contrastive_masked_unet_b200',
            'state_dict': {k: v.detach().cpu() for k, v in core.state_dict().items()}}
    torch.save(ckpt, path)
    return ckpt
