"""Host-side communicator for one-process-per-GPU runs (no MPI, no torch needed).

Launched by ``torchrun`` / ``python -m torch.distributed.run`` the ranks only read ``RANK``,
``WORLD_SIZE``, ``LOCAL_RANK``, ``MASTER_ADDR`` and ``MASTER_PORT`` from the environment.  Rank 0
listens on a TCP port derived from ``MASTER_PORT``; the other ranks connect to it.  The channel
carries the 128-byte NCCL unique id and a handful of small host reductions (max of step times,
sums of error norms): everything on the data path goes through NCCL inside ``libb200ipcs.so``.

It offers the few ``mpi4py`` calls the reference's drivers use on ``mesh.comm``
(``allreduce``, ``Barrier``, ``gather``, ``rank``, ``size``; demo/taylor_green.py:205,207,224).
"""
from __future__ import annotations

import os
import pickle
import socket
import struct
import time

_MAGIC = b"B2IPCS01"


def _send(sock: socket.socket, obj):
    data = pickle.dumps(obj)
    sock.sendall(struct.pack("<Q", len(data)) + data)


def _recv(sock: socket.socket):
    hdr = b""
    while len(hdr) < 8:
        chunk = sock.recv(8 - len(hdr))
        if not chunk:
            raise ConnectionError("peer closed")
        hdr += chunk
    (n,) = struct.unpack("<Q", hdr)
    buf = bytearray()
    while len(buf) < n:
        chunk = sock.recv(min(1 << 20, n - len(buf)))
        if not chunk:
            raise ConnectionError("peer closed")
        buf += chunk
    return pickle.loads(bytes(buf))


class HostComm:
    def __init__(self, rank: int, size: int, addr: str = "127.0.0.1", port: int = 29500, timeout: float = 300.0):
        self.rank, self.size = rank, size
        self._peers: dict[int, socket.socket] = {}
        self._root: socket.socket | None = None
        if size == 1:
            return
        ports = [port + 1000 + 7 * k for k in range(8)]
        if rank == 0:
            srv = None
            for p in ports:
                try:
                    srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
                    srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
                    srv.bind((addr, p))
                    break
                except OSError:
                    srv.close()
                    srv = None
            if srv is None:
                raise RuntimeError("HostComm: no free rendezvous port")
            srv.listen(size)
            srv.settimeout(timeout)
            while len(self._peers) < size - 1:
                conn, _ = srv.accept()
                conn.settimeout(timeout)
                hello = conn.recv(len(_MAGIC) + 4)
                if hello[: len(_MAGIC)] != _MAGIC:
                    conn.close()
                    continue
                (r,) = struct.unpack("<i", hello[len(_MAGIC):])
                conn.sendall(_MAGIC + struct.pack("<i", size))
                conn.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
                self._peers[r] = conn
            srv.close()
        else:
            deadline = time.time() + timeout
            while self._root is None:
                for p in ports:
                    try:
                        s = socket.create_connection((addr, p), timeout=2.0)
                        s.sendall(_MAGIC + struct.pack("<i", rank))
                        reply = s.recv(len(_MAGIC) + 4)
                        if reply[: len(_MAGIC)] == _MAGIC and struct.unpack("<i", reply[len(_MAGIC):])[0] == size:
                            s.settimeout(timeout)
                            s.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
                            self._root = s
                            break
                        s.close()
                    except OSError:
                        pass
                if self._root is None:
                    if time.time() > deadline:
                        raise RuntimeError("HostComm: could not reach rank 0")
                    time.sleep(0.2)

    @classmethod
    def from_env(cls) -> "HostComm":
        return cls(int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
                   os.environ.get("MASTER_ADDR", "127.0.0.1"), int(os.environ.get("MASTER_PORT", "29500")))

    # ---- collectives through rank 0 --------------------------------------------------------
    def gather(self, value, root: int = 0):
        assert root == 0
        if self.size == 1:
            return [value]
        if self.rank == 0:
            out = [value] + [None] * (self.size - 1)
            for r, s in self._peers.items():
                out[r] = _recv(s)
            return out
        _send(self._root, value)
        return None

    def bcast(self, value, root: int = 0):
        assert root == 0
        if self.size == 1:
            return value
        if self.rank == 0:
            for s in self._peers.values():
                _send(s, value)
            return value
        return _recv(self._root)

    def allgather(self, value):
        return self.bcast(self.gather(value))

    def allreduce(self, value, op=None):
        """op: None / 'sum' -> sum; 'max' / 'min'.  (mpi4py's MPI.SUM / MPI.MAX objects are accepted by name.)"""
        vals = self.allgather(value)
        name = getattr(op, "__name__", None) or (op if isinstance(op, str) else "sum") or "sum"
        name = str(name).lower()
        if "max" in name:
            return max(vals)
        if "min" in name:
            return min(vals)
        total = vals[0]
        for v in vals[1:]:
            total = total + v
        return total

    def Barrier(self):
        self.allgather(0)

    barrier = Barrier

    def close(self):
        for s in self._peers.values():
            s.close()
        if self._root is not None:
            self._root.close()
        self._peers, self._root = {}, None
