"""Import site: reference tts/ldm/unet_1d_condition.py:17."""


class UNet2DConditionLoadersMixin:
    """LoRA/attn-processor loading helpers in diffusers; nothing on the hot path uses them."""
