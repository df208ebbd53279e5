"""Multi-partition check, launched under torchrun with one rank per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/mp_gpu_check.py [decomp]
Every rank builds its partition of the same case (C++ host decomposition), steps it on its GPU with the NCCL halo
exchange, and rank 0 compares the gathered result with the same case run as ONE partition on its own GPU.
Prints `MP_CHECK_OK` on success."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nebulasem_b200 import capi, host  # noqa: E402


def main():
    decomp = sys.argv[1] if len(sys.argv) > 1 else "METIS"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        import ctypes
        buf = ctypes.create_string_buffer(128)
        assert capi.load_library().nsem_get_unique_id(buf) == 0
        uid = torch.tensor(list(buf.raw), dtype=torch.uint8, device="cuda")
    dist.broadcast(uid, 0)
    uid_bytes = bytes(uid.cpu().tolist())
    kind, n, order, nsteps = "bubble3d", (4, 4, 4), 3, 12
    if os.environ.get("MP_CHECK_KIND"):        # e.g. "vortex": the doubly periodic isentropic vortex (CYCLIC pairs kept inside a part)
        kind = os.environ["MP_CHECK_KIND"]
        if kind == "vortex":
            n, order, nsteps = (6, 6, 1), 4, 20
    if os.environ.get("MP_CHECK_MESH"):        # e.g. "16,8,8,4,20": elements per direction, order, steps
        v = [int(x) for x in os.environ["MP_CHECK_MESH"].split(",")]
        n, order, nsteps = tuple(v[:3]), v[3], v[4]
    pxyz = {2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}.get(world, (world, 1, 1))
    case_dir = os.environ.get("MP_CHECK_CASE")     # a case directory instead (decomposition as its controls say), e.g. tests/golden/sphere/hydro-sphere
    if case_dir:
        nsteps = int(os.environ.get("MP_CHECK_STEPS", "12"))
        s = host.Solver.open_case(case_dir, 0, rank, world)
    else:
        s = host.Solver.synthetic_part(kind, *n, order, rank, world, decomp, pxyz)
    s.attach(local, rank, world, uid_bytes)
    s.step(nsteps)
    s.download()
    if rank == 0:
        print("halo transport:", s.halo_info)
    rho, U, T, p = s.state()
    nb = s.gBCSfield
    NP = s.NP
    cg = torch.tensor(s.u32("cellGlobal").astype(np.int64), device="cuda")
    ncell_global = n[0] * n[1] * n[2]
    if case_dir:
        cnt = torch.tensor([s.nBCS], dtype=torch.int64, device="cuda")
        dist.all_reduce(cnt)
        ncell_global = int(cnt.item())
    out = torch.zeros((ncell_global, NP, 5), dtype=torch.float64, device="cuda")
    loc = np.concatenate([rho[:nb, None], U[:nb], T[:nb, None]], axis=1).reshape(s.nBCS, NP, 5)
    out[cg] = torch.tensor(loc, device="cuda")
    dist.all_reduce(out)          # partitions are disjoint: the sum assembles the global field
    ok = True
    if rank == 0:
        ref = host.Solver.open_case(case_dir) if case_dir else host.Solver.synthetic(kind, *n, order)
        ref.attach(local)
        ref.step(nsteps)
        ref.download()
        r1, U1, T1, _ = ref.state()
        nb1 = ref.gBCSfield
        one = np.concatenate([r1[:nb1, None], U1[:nb1], T1[:nb1, None]], axis=1).reshape(ncell_global, NP, 5)
        got = out.cpu().numpy()
        for c, nm in enumerate(("rho", "Ux", "Uy", "Uz", "T")):
            scale = max(np.abs(one[..., c]).max(), 1e-30) if nm in ("rho",) else 1.0
            d = np.abs(got[..., c] - one[..., c]).max()
            print(f"{nm}: max |multi - single| = {d:.3e}")
        d_rho = np.linalg.norm(got[..., 0] - one[..., 0]) / np.linalg.norm(one[..., 0])
        d_th = np.linalg.norm(got[..., 0] * (got[..., 4] + 300.0) - one[..., 0] * (one[..., 4] + 300.0)) / np.linalg.norm(one[..., 0] * (one[..., 4] + 300.0))
        d_mom = np.linalg.norm(got[..., 0:1] * got[..., 1:4] - one[..., 0:1] * one[..., 1:4]) / (np.linalg.norm(one[..., 0]) * 347.0)
        print(f"rel L2: rho {d_rho:.3e} rho*theta {d_th:.3e} rho*U (scaled) {d_mom:.3e}")
        ok = d_rho <= 1e-12 and d_th <= 1e-12 and d_mom <= 1e-12 and np.isfinite(got).all()
        print("MP_CHECK_OK" if ok else "MP_CHECK_FAILED", decomp, world, flush=True)
    dist.barrier()
    s.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
