"""MLPRefiner: parameter container with the reference's state_dict layout
(mmedited/models/components/refiners/mlp_refiner.py:65-102: ``layers.{0,2,4,..}``).

Inside the head the three refiners are never called as modules: their weights
are packed into the native plan (ciaosr_b200.native.HeadPlan).  ``forward`` is
kept for API completeness (plain torch ops on whatever device the input is on).
"""
import torch.nn as nn


class MLPRefiner(nn.Module):
    def __init__(self, in_dim, out_dim, hidden_list=None, act=None):
        super().__init__()
        if act not in (None, "relu"):
            raise NotImplementedError("only the ReLU MLPRefiner is used by the reference configs")
        layers, last = [], in_dim
        for hidden in hidden_list or []:
            layers += [nn.Linear(last, hidden), nn.ReLU()]
            last = hidden
        layers.append(nn.Linear(last, out_dim))
        self.layers = nn.Sequential(*layers)

    def forward(self, x):
        shape = x.shape[:-1]
        return self.layers(x.reshape(-1, x.shape[-1])).view(*shape, -1)

    def init_weights(self, pretrained=None, strict=True):
        if pretrained is not None and not isinstance(pretrained, str):
            raise TypeError(f'"pretrained" must be a str or None. But received {type(pretrained)}.')
        if isinstance(pretrained, str):
            from .builder import load_checkpoint
            load_checkpoint(self, pretrained, strict=strict)
