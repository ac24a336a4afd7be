"""Pure-torch stand-in for the one torch_scatter function the reference uses -- TEST INFRASTRUCTURE
ONLY (lets oracle/refrun.py execute the unmodified DecimaScheduler; see ../gymnasium/__init__.py)."""
import torch


def segment_csr(src: torch.Tensor, indptr: torch.Tensor, out=None, reduce: str = "sum") -> torch.Tensor:
    assert reduce == "sum"
    n = indptr.numel() - 1
    res = src.new_zeros((n,) + tuple(src.shape[1:]))
    counts = (indptr[1:] - indptr[:-1]).to(torch.long)
    seg = torch.repeat_interleave(torch.arange(n, device=src.device), counts)
    lo = int(indptr[0])
    res.index_add_(0, seg, src[lo:lo + seg.numel()])
    return res
