"""Pure-torch stand-in for torch_sparse.{SparseTensor, matmul} as used by the reference's NodeEncoder
(schedulers/decima/scheduler.py:219-232, utils.py:67-76) -- TEST INFRASTRUCTURE ONLY."""
import torch


class SparseTensor:
    def __init__(self, row, col, value=None, sparse_sizes=None, is_sorted=False, trust_data=False):
        self.row, self.col, self.sizes = row, col, tuple(sparse_sizes)

    def t(self):
        return SparseTensor(self.col, self.row, sparse_sizes=(self.sizes[1], self.sizes[0]))


def matmul(src: SparseTensor, other: torch.Tensor, reduce: str = "sum") -> torch.Tensor:
    """out[r] = sum over stored entries (r, c) of other[c]  (unweighted adjacency)."""
    assert reduce == "sum"
    out = other.new_zeros((src.sizes[0],) + tuple(other.shape[1:]))
    out.index_add_(0, src.row.to(torch.long), other[src.col.to(torch.long)])
    return out
