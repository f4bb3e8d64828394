def create_transform(*args, **kwargs):
    raise NotImplementedError('timm is not installed; not used on the CM-UNet data path')
