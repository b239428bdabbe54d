"""CPU test of the TCP control plane (helpers/rendezvous.py) that replaced torch.distributed in train.py / bench.py:
world size 3, broadcast of a 128-byte id, barrier, max over ranks, gather."""
import multiprocessing as mp
import os
import socket
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    from sbr_b200.helpers.rendezvous import Control
    ctl = Control(rank=rank, world=world, addr="127.0.0.1", port=port, timeout=60)
    ident = ctl.broadcast(bytes(range(128)) if rank == 0 else None)
    ctl.barrier()
    mx = ctl.all_max(10.0 * rank + 1.0)
    got = ctl.all_gather(rank * rank)
    stop = [ctl.broadcast((i >= 3) if rank == 0 else None) for i in range(5)]
    ctl.barrier()
    ctl.close()
    q.put((rank, ident, mx, got, stop))


def test_control_plane_world3():
    world, port = 3, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=30)
        assert p.exitcode == 0
    for rank, ident, mx, got, stop in res:
        assert ident == bytes(range(128))
        assert mx == 21.0
        assert got == [0, 1, 4]
        assert stop == [False, False, False, True, True]


def test_single_process_is_a_no_op():
    from sbr_b200.helpers.rendezvous import Control
    ctl = Control(rank=0, world=1)
    assert ctl.broadcast("x") == "x" and ctl.all_max(2.5) == 2.5 and ctl.all_gather(7) == [7]
    ctl.barrier()
    ctl.close()
