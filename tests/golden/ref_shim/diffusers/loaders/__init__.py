class UNet2DConditionLoadersMixin:
    pass


class PeftAdapterMixin:
    pass


class FromSingleFileMixin:
    pass


class FromOriginalModelMixin:
    pass


FromOriginalControlNetMixin = FromOriginalModelMixin
