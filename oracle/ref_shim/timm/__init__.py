"""TEST-ONLY stand-in: the reference's processing.py imports `timm.data.create_transform` at module level and never
calls it on the CM-UNet path."""
