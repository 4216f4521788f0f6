import torch


def softmax(src, index, ptr=None, num_nodes=None, dim=0):
    n = int(index.max()) + 1 if num_nodes is None else num_nodes
    mx = torch.full((n,) + tuple(src.shape[1:]), float("-inf"), dtype=src.dtype, device=src.device)
    mx = mx.scatter_reduce(0, index.view(-1, *([1] * (src.dim() - 1))).expand_as(src), src, reduce="amax")
    out = (src - mx[index]).exp()
    s = torch.zeros_like(mx).index_add(0, index, out)
    return out / (s[index] + 1e-16)
