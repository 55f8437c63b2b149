"""A client with the reference's ClientBase task API (elfi/client.py:195-279) for plugging the
device operators into an existing ELFI installation:

    import elfi
    from elfi_b200.client import Client
    elfi.set_client(Client())          # elfi/client.py:31-48 accepts an instance

Batches are the reference's unit of parallelism (parameter_inference.py:270-305 keeps
``max_parallel_batches`` = ``num_cores`` of them in flight).  This client owns one worker thread
per GPU: task t runs on GPU ``devices[t % len(devices)]`` with that device current in its thread
(the node operations pick their context, stream and allocations from the current device;
ctypes releases the GIL for the duration of every library call), so with N GPUs N batches
execute concurrently and ``num_cores`` reports N.  With a single device the task runs lazily in
the caller's thread when its result is requested, like elfi/clients/native.py:17-95.  Results are
handed out by task id, which is how the reference's BatchHandler restores batch order
(elfi/client.py:129-139, 172-182).  All arithmetic happens in the node operations
(elfi_b200.ops / integration/elfi_b200_ops.py); there is none in the client.
If the reference package is importable, the class also derives from its ``ClientBase`` so that
``compile`` / ``load_data`` / ``submit`` / ``compute`` are inherited unchanged.
"""
import itertools
from concurrent.futures import ThreadPoolExecutor

try:                                    # optional: only when plugging into a reference install
    from elfi.client import ClientBase as _Base
except Exception:                       # standalone use: elfi_b200's own samplers need no client
    _Base = object


def _visible_devices():
    try:
        import torch
        return list(range(torch.cuda.device_count()))
    except Exception:
        return []


def _run_on(device, kallable, args, kwargs):
    import torch
    torch.cuda.set_device(device)
    out = kallable(*args, **kwargs)
    torch.cuda.current_stream().synchronize()    # the result is complete when it is handed over
    return out


class Client(_Base):
    def __init__(self, devices=None, **kwargs):
        """`devices`: GPU ordinals to spread the batches over (default: all visible ones)."""
        self.devices = list(devices) if devices is not None else (_visible_devices() or [0])
        self.tasks = {}
        self._ids = itertools.count()
        # one single-threaded executor per device: tasks of a device run in submission order
        self._workers = ([ThreadPoolExecutor(max_workers=1) for _ in self.devices]
                         if len(self.devices) > 1 else [])

    def apply(self, kallable, *args, **kwargs):
        task_id = next(self._ids)
        if self._workers:
            slot = task_id % len(self.devices)
            self.tasks[task_id] = self._workers[slot].submit(_run_on, self.devices[slot], kallable,
                                                             args, kwargs)
        else:
            self.tasks[task_id] = (kallable, args, kwargs)
        return task_id

    def apply_sync(self, kallable, *args, **kwargs):
        return kallable(*args, **kwargs)

    def get_result(self, task_id):
        task = self.tasks.pop(task_id)
        if self._workers:
            return task.result()
        kallable, args, kwargs = task
        return kallable(*args, **kwargs)

    def is_ready(self, task_id):
        task = self.tasks[task_id]
        return task.done() if self._workers else True

    def remove_task(self, task_id):
        task = self.tasks.pop(task_id, None)
        if self._workers and task is not None:
            task.cancel()

    def reset(self):
        for task in self.tasks.values():
            if self._workers:
                task.cancel()
        self.tasks.clear()

    @property
    def num_cores(self):
        """Batches that really execute concurrently: one per device."""
        return max(1, len(self.devices))
