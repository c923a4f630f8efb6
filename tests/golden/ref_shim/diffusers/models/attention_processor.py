class AttnProcessor:
    pass


class AttnAddedKVProcessor:
    pass


AttentionProcessor = AttnProcessor
CROSS_ATTENTION_PROCESSORS = (AttnProcessor,)
ADDED_KV_ATTENTION_PROCESSORS = (AttnAddedKVProcessor,)


import torch.nn as nn
import torch.nn.functional as F


class Attention(nn.Module):
    """INDEPENDENT of oracle/: `Attention` with the default `AttnProcessor2_0`, written on the torch primitive that
    processor calls (`F.scaled_dot_product_attention`, default scale, no mask, dropout 0) - used only by
    tests/golden/make_patch_golden.py so that the blocks the reference's own forward code (patch/patch.py) drives do
    not contain oracle arithmetic.  Parameter names as in the reference's dumps (to_q / to_k / to_v / to_out.0)."""

    def __init__(self, query_dim, heads, dim_head, cross_attention_dim=None):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.out_dim = query_dim                 # read by ToMeBlock.initialize_joint_layers (patch/patch.py:147-151)
        kv = query_dim if cross_attention_dim is None else cross_attention_dim
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(kv, inner, bias=False)
        self.to_v = nn.Linear(kv, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kw):
        assert attention_mask is None and not kw
        src = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        b, n, _ = hidden_states.shape
        split = lambda x: x.view(b, -1, self.heads, x.shape[-1] // self.heads).transpose(1, 2)  # noqa: E731
        o = F.scaled_dot_product_attention(split(self.to_q(hidden_states)), split(self.to_k(src)), split(self.to_v(src)),
                                           attn_mask=None, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(b, n, -1)
        return self.to_out[1](self.to_out[0](o))
