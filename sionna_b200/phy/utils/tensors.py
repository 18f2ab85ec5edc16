"""Shape helpers (mirror of /root/reference/src/sionna/phy/utils/tensors.py), torch views only."""
import torch


def expand_to_rank(tensor, target_rank, axis=-1):
    """Insert singleton dims at ``axis`` until ``tensor`` has ``target_rank`` dims (tensors.py:9-50)."""
    tensor = torch.as_tensor(tensor)
    num_dims = max(target_rank - tensor.dim(), 0)
    return insert_dims(tensor, num_dims, axis)


def insert_dims(tensor, num_dims, axis=-1):
    """Insert ``num_dims`` singleton dims starting at ``axis`` (tensors.py:72-115)."""
    tensor = torch.as_tensor(tensor)
    rank = tensor.dim()
    if not -(rank + 1) <= axis <= rank:
        raise ValueError("`axis` is out of range")
    if axis < 0:
        axis = rank + axis + 1
    shape = list(tensor.shape)
    return tensor.reshape(shape[:axis] + [1] * num_dims + shape[axis:])


def flatten_dims(tensor, num_dims, axis):
    """Merge ``num_dims`` dims starting at ``axis`` (tensors.py:52-70)."""
    shape = list(tensor.shape)
    if num_dims < 2:
        raise ValueError("`num_dims` must be >= 2")
    if num_dims > len(shape) or axis > len(shape) or num_dims + axis > len(shape):
        raise ValueError("`num_dims`/`axis` out of range")
    if num_dims == len(shape):
        return tensor.reshape(-1)
    return tensor.reshape(shape[:axis] + [-1] + shape[axis + num_dims:])


def flatten_last_dims(tensor, num_dims=2):
    """Merge the last ``num_dims`` dims (tensors.py:117-147)."""
    shape = list(tensor.shape)
    if num_dims < 2:
        raise ValueError("`num_dims` must be >= 2")
    if num_dims > len(shape):
        raise ValueError("`num_dims` must <= rank(`tensor`)")
    if num_dims == len(shape):
        return tensor.reshape(-1)
    return tensor.reshape(shape[:-num_dims] + [-1])


def split_dim(tensor, shape, axis):
    """Reshape dim ``axis`` into ``shape`` (tensors.py:149-180)."""
    s = list(tensor.shape)
    if axis < 0:
        axis += len(s)
    return tensor.reshape(s[:axis] + list(shape) + s[axis + 1:])
