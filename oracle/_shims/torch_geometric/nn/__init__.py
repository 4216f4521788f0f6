import torch
import torch.nn as nn


class GlobalAttention(nn.Module):  # imported by the reference, never instantiated
    pass


class SAGEConv(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.lin_l = nn.Linear(in_channels, out_channels, bias=True)
        self.lin_r = nn.Linear(in_channels, out_channels, bias=False)

    def forward(self, x, edge_index):
        src, dst = edge_index[0], edge_index[1]
        agg = torch.zeros_like(x).index_add(0, dst, x[src])
        deg = torch.zeros(x.shape[0], dtype=x.dtype, device=x.device).index_add(0, dst, torch.ones_like(dst, dtype=x.dtype))
        agg = agg / deg.clamp_min(1).unsqueeze(-1)
        return self.lin_l(agg) + self.lin_r(x)


class LayerNorm(nn.Module):
    def __init__(self, in_channels, eps=1e-5, affine=True, mode="graph"):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(in_channels))
        self.bias = nn.Parameter(torch.zeros(in_channels))

    def forward(self, x, batch=None):
        x = x - x.mean()
        out = x / (x.std(unbiased=False) + self.eps)
        return out * self.weight + self.bias
