"""Control plane of a multi-GPU run: a 60-line TCP star instead of torch.distributed.

The data path of the data-parallel step is the library's single ncclAllReduce (include/sbr_b200.h).  The only
things the processes have to tell each other on the host are (a) the 128-byte NCCL id made by rank 0, (b) a barrier
around timed regions, (c) a max over ranks of a timing, and (d) rank 0's decision to stop / validate when that
decision depends on a wall clock (RNNBase.train with --max_time / --time_based_progress).  Rank 0 listens on
MASTER_ADDR:(MASTER_PORT + k) -- the port next to the one the launcher (torchrun) uses for its own store -- and every
other rank keeps one connection to it; every operation is a gather to rank 0 followed by a broadcast.
"""
import os
import pickle
import socket
import struct
import time

_MAGIC = b"SBRB200v1"
_OFFSETS = (17, 29, 43, 61, 83)


def _send(sock, obj):
    data = pickle.dumps(obj, protocol=4)
    sock.sendall(struct.pack("<Q", len(data)) + data)


def _recv(sock):
    hdr = b""
    while len(hdr) < 8:
        chunk = sock.recv(8 - len(hdr))
        if not chunk:
            raise ConnectionError("control-plane peer closed the connection")
        hdr += chunk
    n = struct.unpack("<Q", hdr)[0]
    buf = bytearray()
    while len(buf) < n:
        chunk = sock.recv(min(1 << 20, n - len(buf)))
        if not chunk:
            raise ConnectionError("control-plane peer closed the connection")
        buf += chunk
    return pickle.loads(bytes(buf))


class Control(object):
    """rank / world come from the launcher's environment (RANK, WORLD_SIZE, MASTER_ADDR, MASTER_PORT)."""

    def __init__(self, rank=None, world=None, addr=None, port=None, timeout=300.0):
        self.rank = int(os.environ.get("RANK", "0")) if rank is None else int(rank)
        self.world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else int(world)
        self.addr = addr or os.environ.get("MASTER_ADDR", "127.0.0.1")
        self.port = int(port if port is not None else os.environ.get("MASTER_PORT", "29500"))
        self.peers = []      # rank 0: sockets of ranks 1..world-1 (index = rank - 1)
        self.sock = None     # other ranks: the socket to rank 0
        self._listener = None
        if self.world > 1:
            self._connect(timeout)

    def _connect(self, timeout):
        deadline = time.time() + timeout
        if self.rank == 0:
            srv = None
            for off in _OFFSETS:
                try:
                    s = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
                    s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
                    s.bind((self.addr, self.port + off))
                    s.listen(self.world)
                    srv = s
                    break
                except OSError:
                    s.close()
            if srv is None:
                raise RuntimeError("control plane: no free port next to MASTER_PORT %d" % self.port)
            self._listener = srv
            slots = [None] * (self.world - 1)
            while any(p is None for p in slots):
                srv.settimeout(max(0.1, deadline - time.time()))
                conn, _ = srv.accept()
                conn.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
                hello = _recv(conn)
                if not (isinstance(hello, tuple) and hello[0] == _MAGIC and 1 <= hello[1] < self.world):
                    conn.close()
                    continue
                _send(conn, (_MAGIC, self.world))
                slots[hello[1] - 1] = conn
            self.peers = slots
        else:
            while True:
                for off in _OFFSETS:
                    try:
                        s = socket.create_connection((self.addr, self.port + off), timeout=2.0)
                        s.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
                        _send(s, (_MAGIC, self.rank))
                        s.settimeout(10.0)
                        ack = _recv(s)
                        if isinstance(ack, tuple) and ack[0] == _MAGIC and ack[1] == self.world:
                            s.settimeout(None)
                            self.sock = s
                            return
                        s.close()
                    except (OSError, ConnectionError, pickle.UnpicklingError, struct.error, EOFError):
                        pass
                if time.time() > deadline:
                    raise RuntimeError("control plane: rank %d could not reach rank 0 at %s:%d+k" %
                                       (self.rank, self.addr, self.port))
                time.sleep(0.05)

    # ---- collectives (gather to rank 0, broadcast back)
    def _exchange(self, value, combine):
        if self.world == 1:
            return combine([value])
        if self.rank == 0:
            vals = [value] + [_recv(p) for p in self.peers]
            out = combine(vals)
            for p in self.peers:
                _send(p, out)
            return out
        _send(self.sock, value)
        return _recv(self.sock)

    def broadcast(self, obj):
        """rank 0's `obj` on every rank."""
        return self._exchange(obj, lambda v: v[0])

    def barrier(self):
        self._exchange(None, lambda v: None)

    def all_max(self, x):
        return self._exchange(float(x), max)

    def all_gather(self, obj):
        return self._exchange(obj, list)

    def close(self):
        for p in self.peers:
            try:
                p.close()
            except OSError:
                pass
        for s in (self.sock, self._listener):
            if s is not None:
                try:
                    s.close()
                except OSError:
                    pass
        self.peers, self.sock, self._listener = [], None, None
