"""De-normalisation helpers with the reference's signatures (lib/data_normalization.py:29-53).

On the hot path these are fused into CUDA kernels (``rd_loss``, ``rd_blend_accumulate``); the functions here
serve callers outside the path and keep the reference's semantics: per-sample ``data[i] * std[i] + mean[i]``
when ``std`` is a sequence, a plain broadcast otherwise."""
import numpy as np
import torch


def _is_sequence(v):
    return isinstance(v, (torch.Tensor, list, np.ndarray))


def denormalize_torch(data, mean, std):
    if not _is_sequence(std):
        return data * std + mean
    std_t = torch.as_tensor(std, dtype=data.dtype, device=data.device).flatten().view(-1, 1, 1, 1)
    mean_t = torch.as_tensor(mean, dtype=data.dtype, device=data.device).flatten().view(-1, 1, 1, 1)
    return data * std_t + mean_t


def denormalize_numpy(data, mean, std):
    if isinstance(data, torch.Tensor):
        data = data.detach().cpu().numpy()
    if not _is_sequence(std):
        return data * std + mean
    std_a = np.asarray(torch.as_tensor(std).flatten().tolist() if isinstance(std, torch.Tensor) else std,
                       dtype=np.float64).reshape(-1, 1, 1, 1)
    mean_a = np.asarray(torch.as_tensor(mean).flatten().tolist() if isinstance(mean, torch.Tensor) else mean,
                        dtype=np.float64).reshape(-1, 1, 1, 1)
    return (data * std_a.astype(data.dtype) + mean_a.astype(data.dtype)).astype(data.dtype)
