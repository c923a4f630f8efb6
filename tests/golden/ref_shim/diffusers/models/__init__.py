class AutoencoderKLTemporalDecoder:      # type annotation only on the hot path
    pass


def __getattr__(name):
    if name == "UNetSpatioTemporalConditionModel":   # stock SVD UNet == the oracle's plain UNet
        from oracle import UNetSpatioTemporalConditionControlNetModel
        return UNetSpatioTemporalConditionControlNetModel
    raise AttributeError(name)
