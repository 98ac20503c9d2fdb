"""Process-group plumbing of the row-sharded oracle, without a third-party communication package.

``SocketStore``  a stdlib TCP rendezvous (rank 0 listens, the others connect): all-gather / broadcast of
                 small byte strings and host-side sums.  It carries what has to be agreed on ONCE - the
                 NCCL unique id, the CUDA IPC handles of the exchange areas, the shard index tables - and
                 the reductions of the CPU tests.  Any object with the same three methods
                 (``rank``, ``world``, ``allgather(bytes) -> [bytes]``) can replace it, e.g. an adapter
                 over an existing ``torch.distributed`` / MPI group owned by the caller.
``Comm``         the device-side communicator (``dnlp_comm`` of include/dnlp_b200.h): exchange areas in
                 HBM mapped into every peer through CUDA IPC, plus an NCCL communicator for large payloads.

The data path itself never comes through here: per evaluation the ranks meet in kernels over NVLink
(csrc/dnlp_shard.cu).
"""
import ctypes as C
import os
import pickle
import socket
import struct
import time

import numpy as np


def _send(sock, payload):
    sock.sendall(struct.pack("<Q", len(payload)) + payload)


def _recv(sock):
    hdr = b""
    while len(hdr) < 8:
        chunk = sock.recv(8 - len(hdr))
        if not chunk:
            raise ConnectionError("peer closed the rendezvous connection")
        hdr += chunk
    (n,) = struct.unpack("<Q", hdr)
    buf = bytearray(n)
    view, got = memoryview(buf), 0
    while got < n:
        k = sock.recv_into(view[got:], n - got)
        if k == 0:
            raise ConnectionError("peer closed the rendezvous connection")
        got += k
    return bytes(buf)


class SocketStore:
    """All-gather of byte strings between ``world`` processes over TCP (127.0.0.1 by default).

    Address and port default to ``MASTER_ADDR`` and ``MASTER_PORT + 731`` (torchrun's own store owns
    ``MASTER_PORT`` itself); rank and world size default to ``RANK`` / ``WORLD_SIZE``."""

    def __init__(self, rank=None, world=None, addr=None, port=None, timeout=300.0):
        self.rank = int(os.environ.get("RANK", "0")) if rank is None else int(rank)
        self.world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else int(world)
        addr = addr or os.environ.get("MASTER_ADDR", "127.0.0.1")
        port = int(port) if port is not None else int(os.environ.get("MASTER_PORT", "29500")) + 731
        self._peers, self._srv, self._sock = {}, None, None
        if self.world == 1:
            return
        if self.rank == 0:
            self._srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
            self._srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            self._srv.bind((addr, port))
            self._srv.listen(self.world)
            self._srv.settimeout(timeout)
            while len(self._peers) < self.world - 1:
                conn, _ = self._srv.accept()
                conn.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
                conn.settimeout(timeout)
                (r,) = struct.unpack("<I", _recv(conn))
                self._peers[r] = conn
        else:
            deadline = time.time() + timeout
            while True:
                try:
                    self._sock = socket.create_connection((addr, port), timeout=timeout)
                    break
                except OSError:
                    if time.time() > deadline:
                        raise
                    time.sleep(0.05)
            self._sock.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
            self._sock.settimeout(timeout)
            _send(self._sock, struct.pack("<I", self.rank))

    def allgather(self, payload):
        """Every rank contributes ``payload`` (bytes); returns the list of all contributions in rank order."""
        payload = bytes(payload)
        if self.world == 1:
            return [payload]
        if self.rank == 0:
            parts = [payload] + [None] * (self.world - 1)
            for r, conn in self._peers.items():
                parts[r] = _recv(conn)
            blob = pickle.dumps(parts, protocol=4)
            for conn in self._peers.values():
                _send(conn, blob)
            return parts
        _send(self._sock, payload)
        return pickle.loads(_recv(self._sock))

    def close(self):
        for conn in self._peers.values():
            conn.close()
        if self._sock:
            self._sock.close()
        if self._srv:
            self._srv.close()
        self._peers, self._sock, self._srv = {}, None, None


# ---- helpers on top of any store -----------------------------------------------------------------
def bcast(store, payload, root=0):
    return store.allgather(payload if store.rank == root else b"")[root]


def barrier(store):
    store.allgather(b"")


def allgather_array(store, arr):
    """All-gather of NumPy arrays of one dtype (lengths may differ); list in rank order."""
    arr = np.ascontiguousarray(arr)
    return [np.frombuffer(b, dtype=arr.dtype) for b in store.allgather(arr.tobytes())]


def allreduce_sum(store, vec):
    """Sum of float64 vectors over the ranks, added in rank order (identical on every rank)."""
    parts = allgather_array(store, np.ascontiguousarray(vec, dtype=np.float64))
    out = np.zeros_like(parts[0])
    for p in parts:
        out = out + p
    return out


def allreduce_max(store, vec):
    return np.max(np.stack(allgather_array(store, np.ascontiguousarray(vec, dtype=np.float64))), axis=0)


class Comm:
    """Device-side communicator of one rank (``dnlp_comm``).  ``nccl``: also create an NCCL
    communicator (needs one distinct GPU per rank); without it only the peer-memory path exists."""

    def __init__(self, store, device=0, nccl=True):
        from . import _cabi
        self._L = L = _cabi.lib()
        _cabi._require_device(L)
        self.store, self.rank, self.world, self.device = store, store.rank, store.world, int(device)
        nid = None
        if nccl and self.world > 1:
            buf = C.create_string_buffer(128)
            if self.rank == 0 and L.dnlp_comm_unique_id(buf) != 0:
                raise RuntimeError("dnlp_comm_unique_id: %s" % L.dnlp_comm_last_error(None).decode())
            nid = bcast(store, buf.raw)
        h = C.c_void_p()
        if L.dnlp_comm_create(nid, self.rank, self.world, self.device, C.byref(h)) != 0:
            raise RuntimeError("dnlp_comm_create: %s" % L.dnlp_comm_last_error(None).decode())
        self.h = h
        hb = C.create_string_buffer(64)
        self.check(L.dnlp_comm_ipc_handle(self.h, hb))
        table = b"".join(store.allgather(hb.raw))
        if self.world > 1:
            self.check(L.dnlp_comm_open_peers(self.h, table))
        barrier(store)
        self.has_nccl = bool(L.dnlp_comm_has_nccl(self.h))

    def check(self, rc):
        if rc != 0:
            raise RuntimeError("dnlp_b200 comm: %s" % self._L.dnlp_comm_last_error(self.h).decode())

    def allreduce_sum(self, vec):
        """Sum of a float64 host vector over the ranks through the peer-memory exchange (NVLink)."""
        v = np.array(vec, dtype=np.float64).reshape(-1)
        self.check(self._L.dnlp_comm_allreduce_host(self.h, v.ctypes.data_as(C.POINTER(C.c_double)), int(v.size)))
        return v

    def close(self):
        if getattr(self, "h", None):
            barrier(self.store)                 # nobody unmaps an area a peer may still write to
            self._L.dnlp_comm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self._L.dnlp_comm_destroy(self.h)
                self.h = None
        except Exception:
            pass
