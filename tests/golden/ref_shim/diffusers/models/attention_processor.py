class AttnProcessor:
    pass


class AttnAddedKVProcessor:
    pass


AttentionProcessor = AttnProcessor
CROSS_ATTENTION_PROCESSORS = (AttnProcessor,)
ADDED_KV_ATTENTION_PROCESSORS = (AttnAddedKVProcessor,)
