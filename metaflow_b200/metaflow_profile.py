"""from_start()/profile() timers (same switch as metaflow/metaflow_profile.py:10-34)."""
import time
from contextlib import contextmanager

from .metaflow_config import PROFILE_FROM_START

_t0 = None


def from_start(msg):
    global _t0
    if not PROFILE_FROM_START:
        return
    if _t0 is None:
        _t0 = time.time()
    print("From start: %s took %dms" % (msg, int((time.time() - _t0) * 1000)))


@contextmanager
def profile(label, stats_dict=None):
    if stats_dict is None:
        print("PROFILE: %s starting" % label)
    t = time.time()
    yield
    ms = int((time.time() - t) * 1000)
    if stats_dict is None:
        print("PROFILE: %s completed in %dms" % (label, ms))
    else:
        stats_dict[label] = stats_dict.get(label, 0) + ms
