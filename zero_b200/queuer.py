"""Background preparation of batches (reference utils/queuer.py:37-140, used by main.py:242-250).

The reference runs the batcher in `process_num` forked worker processes that feed a bounded queue, so that reading,
id conversion, length sorting and padding overlap the training step.  Same interface here (`EnQueuer(reader,
preprocessor, worker_processes_num, input_queue_size, output_queue_size)`, iterable, order-preserving), with one
worker THREAD: a generator cannot be shipped to a spawned process, and the work is small next to a training step —
what matters is that the burst when a sort buffer refills (a `buffer_size`-sample sort every few hundred steps)
happens while the GPU is busy with queued work instead of in front of the next launch.  `worker_processes_num = 0`
runs everything inline, like the reference.  The batcher shuffles with the global numpy RNG; it is the only user
of that RNG while it runs, so the batch order is the same with and without the worker.
"""
from __future__ import annotations

import queue
import threading

_DONE = object()


class EnQueuer(object):
    def __init__(self, reader, preprocessor=None, worker_processes_num=1, input_queue_size=5, output_queue_size=5):
        if worker_processes_num < 0:
            raise ValueError("worker_processes_num must be a non-negative integer.")       # utils/queuer.py:45-47
        self.reader = reader
        self.preprocessor = preprocessor if preprocessor is not None else (lambda x: x)
        self.worker_processes_number = int(worker_processes_num)
        self.input_queue_size = int(input_queue_size)
        self.output_queue_size = max(1, int(output_queue_size))

    def __iter__(self):
        if self.worker_processes_number == 0:
            return (self.preprocessor(chunk) for chunk in self.reader)
        return self._threaded()

    def _threaded(self):
        out = queue.Queue(self.output_queue_size)
        stop = threading.Event()

        def put(item):
            while not stop.is_set():
                try:
                    out.put(item, timeout=0.1)
                    return True
                except queue.Full:
                    continue
            return False

        def work():
            try:
                for chunk in self.reader:
                    if not put(self.preprocessor(chunk)):
                        return
                put(_DONE)
            except BaseException as e:      # noqa: BLE001 — handed to the consumer, which re-raises it
                put(e)

        worker = threading.Thread(target=work, name="zero_b200-batcher", daemon=True)
        worker.start()
        try:
            while True:
                item = out.get()
                if item is _DONE:
                    break
                if isinstance(item, BaseException):
                    raise item
                yield item
        finally:
            stop.set()                      # a consumer that leaves early (early stop) releases the worker
            worker.join(timeout=5.0)
