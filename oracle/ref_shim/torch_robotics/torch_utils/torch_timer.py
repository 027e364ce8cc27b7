"""Stand-in for torch_robotics TimerCUDA (mp_baselines/planners/gpmp2.py:20,309,325)."""
import time


class TimerCUDA:
    def __enter__(self):
        self._t0 = time.perf_counter()
        self.elapsed = 0.0
        return self

    def __exit__(self, *exc):
        self.elapsed = time.perf_counter() - self._t0
        return False
