"""Pure-torch stand-in for the torch_geometric surface the reference uses (data.Batch as an attribute
bag, utils.{mask_to_index,index_to_mask,softmax}) -- TEST INFRASTRUCTURE ONLY."""
from . import data, utils  # noqa: F401
