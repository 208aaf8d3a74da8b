"""Host-side communicator for one-process-per-GPU runs (no MPI, no torch needed).

Launched by ``torchrun`` / ``python -m torch.distributed.run`` the ranks only read ``RANK``,
``WORLD_SIZE``, ``LOCAL_RANK``, ``MASTER_ADDR`` and ``MASTER_PORT`` from the environment.  Rank 0
listens on a TCP port derived from ``MASTER_PORT``; the other ranks connect to it.  The channel
carries the 128-byte NCCL unique id and a handful of small host reductions (max of step times,
sums of error norms) as length-prefixed JSON (no pickle: nothing received is ever executed), after an
HMAC challenge-response on a shared secret (B2_COMM_SECRET, else torchrun's run id): everything on the data path goes through NCCL inside ``libb200ipcs.so``.

It offers the few ``mpi4py`` calls the reference's drivers use on ``mesh.comm``
(``allreduce``, ``Barrier``, ``gather``, ``rank``, ``size``; demo/taylor_green.py:205,207,224).
"""
from __future__ import annotations

import base64
import hashlib
import hmac
import json
import os
import socket
import struct
import time

_MAGIC = b"B2IPCS02"
_MAX_MSG = 1 << 26  # nothing on this channel is larger than a few KB; refuse absurd lengths


def _secret() -> bytes:
    """Shared secret of the job's ranks: B2_COMM_SECRET if set, else torchrun's run id (the same string in every
    worker of one launch).  It authenticates the handshake; the payloads are fixed-format JSON, never code."""
    return (os.environ.get("B2_COMM_SECRET") or os.environ.get("TORCHELASTIC_RUN_ID") or "b200ipcs").encode()


def _encode(obj):
    """JSON with two extensions: bytes (the 128-byte NCCL id) and tuples travel as tagged objects."""
    if isinstance(obj, (bytes, bytearray)):
        return {"__bytes__": base64.b64encode(bytes(obj)).decode()}
    if isinstance(obj, tuple):
        return {"__tuple__": [_encode(v) for v in obj]}
    if isinstance(obj, list):
        return [_encode(v) for v in obj]
    if isinstance(obj, dict):
        return {"__dict__": [[_encode(k), _encode(v)] for k, v in obj.items()]}
    if obj is None or isinstance(obj, (bool, int, float, str)):
        return obj
    if hasattr(obj, "tolist"):  # numpy scalars and small arrays
        return _encode(obj.tolist())
    raise TypeError(f"HostComm carries plain data only (numbers, strings, bytes, lists, tuples, dicts), not {type(obj)!r}")


def _decode(obj):
    if isinstance(obj, list):
        return [_decode(v) for v in obj]
    if isinstance(obj, dict):
        if "__bytes__" in obj:
            return base64.b64decode(obj["__bytes__"])
        if "__tuple__" in obj:
            return tuple(_decode(v) for v in obj["__tuple__"])
        if "__dict__" in obj:
            return {(_decode(k) if not isinstance(k, list) else tuple(_decode(k))): _decode(v) for k, v in obj["__dict__"]}
    return obj


def _read_exact(sock: socket.socket, n: int) -> bytes:
    buf = bytearray()
    while len(buf) < n:
        chunk = sock.recv(min(1 << 20, n - len(buf)))
        if not chunk:
            raise ConnectionError("peer closed")
        buf += chunk
    return bytes(buf)


def _send(sock: socket.socket, obj):
    data = json.dumps(_encode(obj), allow_nan=True).encode()
    sock.sendall(struct.pack("<Q", len(data)) + data)


def _recv(sock: socket.socket):
    (n,) = struct.unpack("<Q", _read_exact(sock, 8))
    if n > _MAX_MSG:
        raise ConnectionError(f"HostComm: message of {n} bytes refused")
    return _decode(json.loads(_read_exact(sock, n).decode()))


def _resolve_op(op) -> str:
    """'sum' | 'max' | 'min' from a string, None (sum) or an mpi4py op object (compared by identity)."""
    if op is None:
        return "sum"
    if isinstance(op, str):
        name = op.lower()
        if name in ("sum", "max", "min"):
            return name
        raise ValueError(f"HostComm.allreduce: unknown op {op!r}")
    try:
        from mpi4py import MPI  # only where it exists

        for name, o in (("sum", MPI.SUM), ("max", MPI.MAX), ("min", MPI.MIN)):
            if op is o:
                return name
    except ImportError:
        pass
    raise ValueError(f"HostComm.allreduce: unsupported op {op!r} (use 'sum', 'max' or 'min')")


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
                try:
                    # challenge-response: only a peer that knows the job's secret is admitted
                    nonce = os.urandom(16)
                    conn.sendall(_MAGIC + nonce)
                    hello = _read_exact(conn, len(_MAGIC) + 4 + 32)
                    (r,) = struct.unpack("<i", hello[len(_MAGIC):len(_MAGIC) + 4])
                    want = hmac.new(_secret(), nonce + struct.pack("<i", r), hashlib.sha256).digest()
                    if hello[: len(_MAGIC)] != _MAGIC or not hmac.compare_digest(hello[len(_MAGIC) + 4:], want) \
                            or not (0 < r < size) or r in self._peers:
                        conn.close()
                        continue
                    conn.sendall(_MAGIC + struct.pack("<i", size) + hmac.new(_secret(), want, hashlib.sha256).digest())
                except (OSError, ConnectionError, struct.error):
                    conn.close()
                    continue
                conn.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
                self._peers[r] = conn
            srv.close()
        else:
            deadline = time.time() + timeout
            while self._root is None:
                for p in ports:
                    try:
                        s = socket.create_connection((addr, p), timeout=2.0)
                        s.settimeout(10.0)
                        chal = _read_exact(s, len(_MAGIC) + 16)
                        if chal[: len(_MAGIC)] != _MAGIC:
                            s.close()
                            continue
                        mac = hmac.new(_secret(), chal[len(_MAGIC):] + struct.pack("<i", rank), hashlib.sha256).digest()
                        s.sendall(_MAGIC + struct.pack("<i", rank) + mac)
                        reply = _read_exact(s, len(_MAGIC) + 4 + 32)
                        # the listener proves knowledge of the secret too (workers do not trust any process on the port)
                        ok = hmac.compare_digest(reply[len(_MAGIC) + 4:], hmac.new(_secret(), mac, hashlib.sha256).digest())
                        if ok and reply[: len(_MAGIC)] == _MAGIC and struct.unpack("<i", reply[len(_MAGIC):len(_MAGIC) + 4])[0] == size:
                            s.settimeout(timeout)
                            s.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
                            self._root = s
                            break
                        s.close()
                    except (OSError, ConnectionError):
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
        """op: None / 'sum', 'max', 'min' (or mpi4py's MPI.SUM / MPI.MAX / MPI.MIN); anything else raises."""
        name = _resolve_op(op)
        vals = self.allgather(value)
        if name == "max":
            return max(vals)
        if name == "min":
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
