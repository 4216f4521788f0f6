import torch


def scatter_add(src, index, dim=0, out=None, dim_size=None):
    assert dim == 0
    size = int(index.max()) + 1 if dim_size is None else dim_size
    res = torch.zeros((size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return res.index_add(0, index, src)
