"""A client with the reference's ClientBase task API (elfi/client.py:195-279) for plugging the
device operators into an existing ELFI installation:

    import elfi
    from elfi_b200.client import Client
    elfi.set_client(Client())          # elfi/client.py:31-48 accepts an instance

ELFI's master loop needs results in batch_index order and one GPU runs one batch at a time, so the
client executes each task lazily in the caller's thread when its result is requested (like
elfi/clients/native.py:17-95) and reports ``num_cores`` = number of visible GPUs, which becomes
the default ``max_parallel_batches`` (parameter_inference.py:93).  All the work happens in the node
operations (elfi_b200.ops) that the executed graph calls; there is no arithmetic in the client.
If the reference package is importable, the class also derives from its ``ClientBase`` so that
``compile`` / ``load_data`` / ``submit`` / ``compute`` are inherited unchanged.
"""
import itertools

try:                                    # optional: only when plugging into a reference install
    from elfi.client import ClientBase as _Base
except Exception:                       # standalone use: elfi_b200's own samplers need no client
    _Base = object


class Client(_Base):
    def __init__(self, devices=None, **kwargs):
        self.tasks = {}
        self._ids = itertools.count()
        self._devices = devices

    def apply(self, kallable, *args, **kwargs):
        task_id = next(self._ids)
        self.tasks[task_id] = (kallable, args, kwargs)
        return task_id

    def apply_sync(self, kallable, *args, **kwargs):
        return kallable(*args, **kwargs)

    def get_result(self, task_id):
        kallable, args, kwargs = self.tasks.pop(task_id)
        return kallable(*args, **kwargs)

    def is_ready(self, task_id):
        return True

    def remove_task(self, task_id):
        self.tasks.pop(task_id, None)

    def reset(self):
        self.tasks.clear()

    @property
    def num_cores(self):
        if self._devices is not None:
            return max(1, len(self._devices))
        try:
            import torch
            return max(1, torch.cuda.device_count())
        except Exception:
            return 1
