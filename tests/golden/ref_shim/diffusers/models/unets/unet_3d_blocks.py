"""get_down_block / get_up_block / UNetMidBlockSpatioTemporal: un-vendored block arithmetic, delegated to the
oracle's restatement (the ``resnet_eps`` the reference passes is ignored by the 0.27.2 factories
for the spatio-temporal blocks - recollection U1 - and each block class keeps its own default eps)."""
from oracle.blocks import (CrossAttnDownBlockSpatioTemporal, CrossAttnUpBlockSpatioTemporal, DownBlockSpatioTemporal,
                           UpBlockSpatioTemporal)
from oracle.blocks import UNetMidBlockSpatioTemporal as _Mid


def get_down_block(down_block_type, num_layers, in_channels, out_channels, temb_channels, add_downsample,
                   resnet_eps=None, resnet_act_fn=None, transformer_layers_per_block=1, num_attention_heads=None,
                   cross_attention_dim=None, **kw):
    if down_block_type == "DownBlockSpatioTemporal":
        return DownBlockSpatioTemporal(in_channels, out_channels, temb_channels, num_layers,
                                       add_downsample=add_downsample)
    if down_block_type == "CrossAttnDownBlockSpatioTemporal":
        return CrossAttnDownBlockSpatioTemporal(in_channels, out_channels, temb_channels, num_layers,
                                                transformer_layers_per_block, num_attention_heads,
                                                cross_attention_dim, add_downsample=add_downsample)
    raise ValueError(f"{down_block_type} does not exist.")


def get_up_block(up_block_type, num_layers, in_channels, out_channels, prev_output_channel, temb_channels,
                 add_upsample, resnet_eps=None, resnet_act_fn=None, resolution_idx=None,
                 transformer_layers_per_block=1, num_attention_heads=None, cross_attention_dim=None, **kw):
    if up_block_type == "UpBlockSpatioTemporal":
        return UpBlockSpatioTemporal(in_channels, prev_output_channel, out_channels, temb_channels, num_layers,
                                     add_upsample=add_upsample)
    if up_block_type == "CrossAttnUpBlockSpatioTemporal":
        return CrossAttnUpBlockSpatioTemporal(in_channels, prev_output_channel, out_channels, temb_channels,
                                              num_layers, transformer_layers_per_block, num_attention_heads,
                                              cross_attention_dim, add_upsample=add_upsample)
    raise ValueError(f"{up_block_type} does not exist.")


class UNetMidBlockSpatioTemporal(_Mid):
    def __init__(self, in_channels, temb_channels, num_layers=1, transformer_layers_per_block=1,
                 num_attention_heads=1, cross_attention_dim=1280):
        super().__init__(in_channels, temb_channels, num_layers, transformer_layers_per_block, num_attention_heads,
                         cross_attention_dim)
