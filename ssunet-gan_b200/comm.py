"""Master/slave rendezvous API of the reference's SyncBN (comm.py:18-138), kept for drop-in
compatibility.  On the B200 path the exchange itself is an NCCL all-reduce issued by
SynchronizedBatchNorm (one process per GPU), so these classes are only exercised by code that
drives them directly (e.g. thread-level tests); semantics match the reference:
one-to-one FutureResult pipes, a registry of slaves, `run_master` collecting one message per
slave, invoking the callback and distributing the replies, then waiting for every ACK.
"""
import collections
import queue
import threading

__all__ = ["FutureResult", "SlavePipe", "SyncMaster"]


class FutureResult(object):
    """Single-slot, thread-safe mailbox: `put` once, `get` once."""

    def __init__(self):
        self._cond = threading.Condition(threading.Lock())
        self._value = None

    def put(self, result):
        with self._cond:
            assert self._value is None, "Previous result has't been fetched."
            self._value = result
            self._cond.notify()

    def get(self):
        with self._cond:
            while self._value is None:
                self._cond.wait()
            value, self._value = self._value, None
            return value


_Registry = collections.namedtuple("MasterRegistry", ["result"])
_PipeFields = collections.namedtuple("_SlavePipeBase", ["identifier", "queue", "result"])


class SlavePipe(_PipeFields):
    """Slave end: post (identifier, msg), block for the reply, acknowledge."""

    def run_slave(self, msg):
        self.queue.put((self.identifier, msg))
        reply = self.result.get()
        self.queue.put(True)
        return reply


class SyncMaster(object):
    """Collects one message from every registered slave per forward pass and answers each."""

    def __init__(self, master_callback):
        self._master_callback = master_callback
        self._queue = queue.Queue()
        self._registry = collections.OrderedDict()
        self._activated = False

    def __getstate__(self):
        return {"master_callback": self._master_callback}

    def __setstate__(self, state):
        self.__init__(state["master_callback"])

    def register_slave(self, identifier):
        if self._activated:   # first registration after a forward: start a fresh round
            assert self._queue.empty(), "Queue is not clean before next initialization."
            self._activated = False
            self._registry.clear()
        future = FutureResult()
        self._registry[identifier] = _Registry(future)
        return SlavePipe(identifier, self._queue, future)

    def run_master(self, master_msg):
        self._activated = True
        gathered = [(0, master_msg)] + [self._queue.get() for _ in range(self.nr_slaves)]
        replies = self._master_callback(gathered)
        assert replies[0][0] == 0, "The first result should belongs to the master."
        for ident, reply in replies:
            if ident != 0:
                self._registry[ident].result.put(reply)
        for _ in range(self.nr_slaves):
            assert self._queue.get() is True
        return replies[0][1]

    @property
    def nr_slaves(self):
        return len(self._registry)
