import torch.nn as nn


class AutoregressivePolicy(nn.Module):
    """Holds encoder/decoder and the decode defaults (rl4co ConstructivePolicy.__init__)."""

    def __init__(self, encoder, decoder, env_name="tsp", temperature=1.0, tanh_clipping=0,
                 mask_logits=True, train_decode_type="sampling", val_decode_type="greedy",
                 test_decode_type="greedy", **unused_kw):
        super().__init__()
        self.encoder = encoder
        self.decoder = decoder
        self.env_name = env_name
        self.temperature = temperature
        self.tanh_clipping = tanh_clipping
        self.mask_logits = mask_logits
        self.train_decode_type = train_decode_type
        self.val_decode_type = val_decode_type
        self.test_decode_type = test_decode_type
