"""Process-wide path counters.  The reference's scripts are run UNMODIFIED, so a test cannot ask them which of our code paths they
took; with LGR_TRACE=<file> set, the counters below are written to that file (JSON) when the interpreter exits:
    render_fused / render_unfused      gaussian_renderer.render()/count_render() calls through the raw-leaf kernels / the plain API
    render_fused_strided_rest          ... of which with a row-strided _features_rest (the distillation student)
    adamw_steps / adamw_strided_params FusedAdamW.step() calls / row-strided parameters updated in place
    unfused_exchange                   view-parallel backward passes that took the dense all-reduce of the unfused node
Cost when LGR_TRACE is unset: one dict increment per call."""
from __future__ import annotations

import atexit
import json
import os

counters: dict = {}


def bump(name: str, by: int = 1) -> None:
    counters[name] = counters.get(name, 0) + by


def _dump() -> None:
    path = os.environ.get("LGR_TRACE")
    if path:
        try:
            with open(path, "w") as f:
                json.dump(counters, f)
        except OSError:
            pass


atexit.register(_dump)
