"""from_start() / profile() timers behind the METAFLOW_PROFILE_FROM_START switch (metaflow/metaflow_profile.py:10-34):
same call sites, same printed lines ("From start: <msg> took <n>ms", "PROFILE: <label> completed in <n>ms")."""
import time

from . import metaflow_config


class _Clock(object):
    """Milliseconds since the first reading."""

    def __init__(self):
        self._origin = None

    def ms(self):
        now = time.perf_counter()
        if self._origin is None:
            self._origin = now
        return int((now - self._origin) * 1000)


_since_start = _Clock()


def from_start(msg):
    if metaflow_config.PROFILE_FROM_START:
        print("From start: %s took %dms" % (msg, _since_start.ms()))


class profile(object):
    """`with profile(label):` prints start/end lines; `with profile(label, stats):` adds the milliseconds to
    stats[label] instead."""

    def __init__(self, label, stats_dict=None):
        self._label, self._stats = label, stats_dict

    def __enter__(self):
        if self._stats is None:
            print("PROFILE: %s starting" % self._label)
        self._t0 = time.perf_counter()
        return self

    def __exit__(self, *exc):
        took = int((time.perf_counter() - self._t0) * 1000)
        if self._stats is None:
            print("PROFILE: %s completed in %dms" % (self._label, took))
        else:
            self._stats[self._label] = self._stats.get(self._label, 0) + took
        return False
