"""Device policy of the variant-2 drop-in: CUDA (B200) or nothing for inference.

API-compatible with shopformer_2/utils/device.py:11-115; the MPS helpers are kept as no-ops so
that the reference's train.py / evaluate.py import and call them unchanged.
"""
import torch


def get_device(preference: str = "auto") -> torch.device:
    pref = (preference or "auto").lower()
    if pref == "cpu":
        return torch.device("cpu")
    if pref in ("auto", "cuda") and torch.cuda.is_available():
        return torch.device("cuda")
    if pref == "mps" and check_mps_availability():
        return torch.device("mps")
    return torch.device("cpu")


def check_mps_availability() -> bool:
    mps = getattr(torch.backends, "mps", None)
    return bool(mps is not None and mps.is_available() and mps.is_built())


def setup_mps_environment():
    """No-op on CUDA hosts."""


def clear_mps_cache():
    """No-op on CUDA hosts."""


def print_device_info():
    print(f"PyTorch {torch.__version__}; CUDA available: {torch.cuda.is_available()}")
    if torch.cuda.is_available():
        print(f"  device 0: {torch.cuda.get_device_name(0)}")


def move_to_device(data, device: torch.device):
    if isinstance(data, torch.Tensor):
        return data.to(device)
    if isinstance(data, (list, tuple)):
        return type(data)(move_to_device(d, device) for d in data)
    if isinstance(data, dict):
        return {k: move_to_device(v, device) for k, v in data.items()}
    return data
