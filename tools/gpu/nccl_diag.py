"""Step-by-step NCCL bring-up through the C ABI, logging each stage to
gpurun_out/nccl_diag_<rank>.log so a hang can be located."""
import ctypes as ct
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
logf = open(os.path.join(ROOT, 'gpurun_out', f'nccl_diag_{rank}.log'), 'w')
T0 = time.time()


def log(msg):
    logf.write(f'[{time.time() - T0:7.2f}] {msg}\n')
    logf.flush()
    os.fsync(logf.fileno())
    print(f'[rank {rank}] {msg}', flush=True)


import numpy as np
from pyfr_b200.lib import load_runtime
from pyfr_b200.comm import NCCLComm

rt = load_runtime(int(os.environ.get('LOCAL_RANK', 0)))
log(f'runtime up: {rt.device_info()}')
s = rt.new_ptr(rt.stream_create)
cs = rt.new_ptr(rt.stream_create)
comm = NCCLComm(rt, rank, world)
log('nccl comm initialised')

n = 1 << 20
a = rt.new_ptr(rt.malloc, 8*n)
b = rt.new_ptr(rt.malloc, 8*n)
h = np.full(n, float(rank + 1))
rt.memcpy(a, h.ctypes.data, 8*n)

comm.allreduce(a, 4, 1, 0, s)
rt.stream_sync(s)
rt.memcpy(h.ctypes.data, a, 8*n)
log(f'allreduce ok: {h[:2]} (expect {sum(range(1, world + 1))})')

peer = (rank + 1) % world


def xchg(stream):
    rt.nccl_group_start()
    rt.nccl_send(comm._handle, a, n, 1, peer, stream)
    rt.nccl_recv(comm._handle, b, n, 1, (rank - 1) % world, stream)
    rt.nccl_group_end()


xchg(s)
rt.stream_sync(s)
log('eager send/recv ok')

ev1, ev2 = rt.new_ptr(rt.event_create), rt.new_ptr(rt.event_create)

# point-to-point bandwidth, 32 MiB messages both ways
big = 4 << 20
a2 = rt.new_ptr(rt.malloc, 8*big)
b2 = rt.new_ptr(rt.malloc, 8*big)
for rep in range(2):
    rt.event_record(ev1, s)
    for i in range(10):
        rt.nccl_group_start()
        rt.nccl_send(comm._handle, a2, big, 1, peer, s)
        rt.nccl_recv(comm._handle, b2, big, 1, (rank - 1) % world, s)
        rt.nccl_group_end()
    rt.event_record(ev2, s)
    rt.stream_sync(s)
    ms = rt.elapsed_ms(ev1, ev2)/10
    log(f'p2p 32 MiB each way: {ms:.3f} ms -> {8*big/ms/1e6:.1f} GB/s per direction')
log('done')
logf.close()
os._exit(0)
