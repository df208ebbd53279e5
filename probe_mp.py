import os, sys, ctypes
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, '/root/repo')
from nebulasem_b200 import capi, host
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // world))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    buf = ctypes.create_string_buffer(128); assert capi.load_library().nsem_get_unique_id(buf) == 0
    uid = torch.tensor(list(buf.raw), dtype=torch.uint8, device="cuda")
dist.broadcast(uid, 0)
n = int(sys.argv[1])
s = host.Solver.synthetic_part("bubble3d", 2*n, n, n, 4, rank, world, "METIS", (2,1,1))
s.attach(local, rank, world, bytes(uid.cpu().tolist()))
print(rank, s.kernel_info, 'nBCS', s.nBCS, 'gALL', s.gALL, flush=True)
mode = sys.argv[2]
if mode == "diag":
    for it in range(20):
        s.step(1)
        d = s.diagnostics()
        if rank == 0: print('step', it, 'courant_max %.3e mass %.12e' % (d['courant_max'], d['mass']), flush=True)
else:
    s.step(3); s.sync()
    ms, _ = s.time_steps(10, per_kernel=False)
    d = s.diagnostics()
    if rank == 0: print('after 13 back-to-back: courant_max %.3e mass %.12e ms/step %.2f' % (d['courant_max'], d['mass'], ms/10), flush=True)
    ms, pk = s.time_steps(5, per_kernel=True)
    d = s.diagnostics()
    if rank == 0: print('after per-kernel 5: courant_max %.3e mass %.12e' % (d['courant_max'], d['mass']), pk/5, flush=True)
s.download(); rho,U,T,p = s.state(); nb = s.gBCSfield
print('  rank', rank, 'nonfinite real', [int((~np.isfinite(x[:nb])).sum()) for x in (rho,U,T,p)], 'ghost', [int((~np.isfinite(x[nb:])).sum()) for x in (rho,U,T,p)], flush=True)
dist.barrier(); s.close(); dist.destroy_process_group()
