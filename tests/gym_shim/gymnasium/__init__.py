"""TEST-HARNESS stand-in for the `gymnasium` package (not installed in the build image, no network): exactly the surface the
reference's main.py / sac/envs.py / sac/agent.py touch -- ``Env`` (reset(seed=) seeding ``np_random``), ``spaces.Box``,
``make``. Put on sys.path by tests only; the product imports the real gymnasium."""
import numpy as np

from . import spaces  # noqa: F401


class Env:
    metadata = {}
    spec = None
    render_mode = None
    np_random = None

    def reset(self, seed=None, options=None):
        if seed is not None or self.np_random is None:
            self.np_random = np.random.default_rng(seed)

    def close(self):
        pass


def make(env_id, **kwargs):
    raise RuntimeError(f"gymnasium shim: no registry (asked for {env_id!r}); the probe environments are constructed by name")
