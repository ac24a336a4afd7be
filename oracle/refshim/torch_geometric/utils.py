import torch


def mask_to_index(mask: torch.Tensor) -> torch.Tensor:
    return mask.nonzero(as_tuple=False).view(-1)


def index_to_mask(index: torch.Tensor, size: int) -> torch.Tensor:
    mask = index.new_zeros(size, dtype=torch.bool)
    mask[index] = True
    return mask


def softmax(src: torch.Tensor, index=None, ptr=None, num_nodes=None, dim: int = 0) -> torch.Tensor:
    assert ptr is not None
    out = torch.empty_like(src)
    for i in range(ptr.numel() - 1):
        lo, hi = int(ptr[i]), int(ptr[i + 1])
        out[lo:hi] = torch.softmax(src[lo:hi], dim)
    return out
