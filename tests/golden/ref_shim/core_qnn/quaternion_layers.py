from oracle.unet import QuaternionLinear as QuaternionLinearAutograd  # noqa: F401

__all__ = ["QuaternionLinearAutograd"]
