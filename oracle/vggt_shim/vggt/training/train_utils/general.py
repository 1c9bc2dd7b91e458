import torch


def check_and_fix_inf_nan(input_tensor, loss_name="default", hard_max=100):
    if input_tensor is None:
        return input_tensor
    if torch.isnan(input_tensor).any() or torch.isinf(input_tensor).any():
        input_tensor = torch.nan_to_num(input_tensor, nan=0.0, posinf=0.0, neginf=0.0)
    if hard_max is not None:
        input_tensor = torch.clamp(input_tensor, min=-hard_max, max=hard_max)
    return input_tensor
