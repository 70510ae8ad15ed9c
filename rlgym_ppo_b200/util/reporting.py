"""Console / wandb iteration report with the reference's keys and grouping (rlgym_ppo/util/reporting.py:23-75).
A missing key is an error there (KeyError) and here."""
import numpy as np
import torch

_GROUPS = (
    ("Policy Reward", "Policy Entropy", "Value Function Loss"),
    ("Mean KL Divergence", "SB3 Clip Fraction", "Policy Update Magnitude", "Value Function Update Magnitude"),
    ("Collected Steps per Second", "Overall Steps per Second"),
    ("Timestep Collection Time", "Timestep Consumption Time", "PPO Batch Consumption Time", "Total Iteration Time"),
    ("Cumulative Model Updates", "Cumulative Timesteps"),
    ("Timesteps Collected",),
)


def _fmt(val):
    if isinstance(val, torch.Tensor):
        val = val.detach().cpu().item() if val.dim() == 0 else val.detach().cpu().tolist()
    if isinstance(val, (tuple, list, np.ndarray)):
        return "[" + ", ".join(_fmt(v) for v in val) + "]"
    if isinstance(val, (float, np.floating)):
        return f"{float(val):,.5f}"
    if isinstance(val, (int, np.integer)) and not isinstance(val, bool):
        return f"{int(val):,d}"
    return str(val)


def dump_dict_to_debug_string(dictionary):
    return "".join(f"{k}: {_fmt(v)}\n" for k, v in dictionary.items())


def report_metrics(loggable_metrics, debug_metrics, wandb_run=None):
    if wandb_run is not None:
        wandb_run.log(loggable_metrics)
    if debug_metrics is not None:
        print("\nBEGIN DEBUG\n")
        print(dump_dict_to_debug_string(debug_metrics))
        print("\nEND DEBUG\n")
    print("-" * 8 + "BEGIN ITERATION REPORT" + "-" * 8)
    blocks = [dump_dict_to_debug_string({k: loggable_metrics[k] for k in group}) for group in _GROUPS]
    print("\n".join(blocks).rstrip("\n"))
    print("-" * 8 + "END ITERATION REPORT" + "-" * 8 + "\n\n")
