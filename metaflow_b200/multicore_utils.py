"""parallel_map / parallel_imap_unordered with the reference's signatures
(metaflow/multicore_utils.py:82-175): `func` runs in forked children (so closures and lambdas work and there is
no 32 MB queue limit -- results come back through files), at most `max_parallel` at a time, and
parallel_map returns the results in argument order.

What is different is what the children do with ARTIFACTS.  A forked child cannot use its parent's CUDA
context, and one context per child would serialise them on the device.  So before the first fork the
parent starts the batcher (metaflow_b200/batcher.py: CodecServer) on its own GpuCodec and the children reach
the GPUs through it: whatever the siblings save or load at the same time is merged into one wide GPU batch.
Without a usable GPU in the parent the children simply have no GPU codec (the datastore classes then raise
GpuCasUnavailable on first use -- there is no CPU fallback); functions that never touch artifacts are not affected.
"""
import os
import pickle
import sys
import time
import traceback
from itertools import islice
from tempfile import NamedTemporaryFile


class MulticoreException(Exception):
    pass


_server = None


def _ensure_batcher():
    """Start (once) the in-process batcher on the default codec; returns its socket path or None."""
    global _server
    from . import batcher

    if os.environ.get(batcher.SERVER_ENV):
        return os.environ[batcher.SERVER_ENV]               # already inside a served tree
    if _server is not None:
        return _server.socket_path
    try:
        from .codec import default_codec

        codec = default_codec()
    except Exception:  # noqa: BLE001 -- no library / no device: children run without a codec
        return None
    _server = batcher.CodecServer(codec).start()
    import atexit

    atexit.register(_stop_batcher, os.getpid())
    return _server.socket_path


def _stop_batcher(owner_pid):
    global _server
    if _server is not None and os.getpid() == owner_pid:      # (a forked child inherits the object, not the server)
        _server.stop()
    _server = None


def _child_setup(server_path):
    from . import batcher, codec

    codec._default = None                                    # the parent's CUDA context is not usable after fork
    codec._forked_from_gpu_parent = True
    if server_path:
        os.environ[batcher.SERVER_ENV] = server_path


def _spawn(func, arg, dir, server_path):
    with NamedTemporaryFile(prefix="parallel_map_", dir=dir, delete=False) as tmp:
        out_path = tmp.name
    sys.stderr.flush()
    sys.stdout.flush()
    pid = os.fork()
    if pid:
        return pid, out_path
    code = 1
    try:
        _child_setup(server_path)
        result = func(arg)
        with open(out_path, "wb") as f:
            pickle.dump(result, f, protocol=pickle.HIGHEST_PROTOCOL)
        code = 0
    except BaseException:  # noqa: BLE001 -- nothing may escape into the parent's stack frames above the fork
        traceback.print_exc()
    finally:
        sys.stderr.flush()
        sys.stdout.flush()
        os._exit(code)


def parallel_imap_unordered(func, iterable, max_parallel=None, dir=None):
    """Yields func(x) for x in iterable, in completion order."""
    if max_parallel is None:
        from multiprocessing import cpu_count

        max_parallel = cpu_count()
    server_path = _ensure_batcher()
    args = iter(iterable)
    running = []
    try:
        for a in islice(args, max_parallel):
            running.append(_spawn(func, a, dir, server_path))
        yield from _drain(func, args, running, dir, server_path)
    finally:
        # an abandoned or failed iteration leaves neither orphans nor result files behind
        for pid, path in running:
            try:
                os.waitpid(pid, 0)
            except ChildProcessError:
                pass
            _silent_remove(path)


def _drain(func, args, running, dir, server_path):
    delay = 0.001
    while running:
        finished = None
        for k, (pid, out_path) in enumerate(running):
            got, status = os.waitpid(pid, os.WNOHANG)
            if got:
                finished = (k, out_path, status)
                break
        if finished is None:
            time.sleep(delay)                                # the reference polls every 100 ms; back off up to that
            delay = min(0.1, delay * 2)
            continue
        delay = 0.001
        k, out_path, status = finished
        running.pop(k)
        if status:
            _silent_remove(out_path)
            raise MulticoreException("Child failed")             # (the caller's finally reaps the siblings)
        with open(out_path, "rb") as f:
            value = pickle.load(f)
        _silent_remove(out_path)
        yield value
        nxt = list(islice(args, 1))
        if nxt:
            running.insert(0, _spawn(func, nxt[0], dir, server_path))


def _silent_remove(path):
    try:
        os.remove(path)
    except OSError:
        pass


def parallel_map(func, iterable, max_parallel=None, dir=None):
    """[func(x) for x in iterable], computed by forked children, in the order of `iterable`."""

    def indexed(pair):
        return pair[0], func(pair[1])

    done = parallel_imap_unordered(indexed, enumerate(iterable), max_parallel=max_parallel, dir=dir)
    return [v for _, v in sorted(done, key=lambda p: p[0])]
