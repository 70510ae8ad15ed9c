"""MetricsLogger: user metrics that ride along with env steps (rlgym_ppo/util/metrics_logger.py).
Wire format per metric array: [ndim, *shape, *values] as float32, concatenated."""
from abc import ABC

import numpy as np


class MetricsLogger(ABC):
    def collect_metrics(self, game_state) -> np.ndarray:
        flat = []
        for arr in self._collect_metrics(game_state):
            a = np.asarray(arr)
            flat.append(float(a.ndim))
            flat.extend(float(d) for d in a.shape)
            flat.extend(np.ravel(a).astype(np.float64).tolist())
        return np.asarray(flat, dtype=np.float32)

    def report_metrics(self, collected_metrics, wandb_run, cumulative_timesteps):
        if wandb_run is None:
            return
        reports = []
        for blob in collected_metrics:
            arrays, i = [], 0
            while i < len(blob):
                ndim = int(blob[i])
                shape = [int(d) for d in blob[i + 1:i + 1 + ndim]]
                count = int(np.prod(shape)) if ndim else 1
                arrays.append(blob[i + 1 + ndim:i + 1 + ndim + count])
                i += 1 + ndim + count
            reports.append(arrays)
        self._report_metrics(reports, wandb_run, cumulative_timesteps)

    def _collect_metrics(self, game_state):
        raise NotImplementedError

    def _report_metrics(self, collected_metrics, wandb_run, cumulative_timesteps):
        raise NotImplementedError
