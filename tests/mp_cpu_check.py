"""World-size-2 CPU check of the partition + halo logic over gloo (no GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tests/mp_cpu_check.py
Each rank builds its partition with the C++ host library, packs the owner-side face values of its interMesh patch in
patch order (what halo_pack_kernel does on the device, ASYNC_COMM::send field.h:2283-2290) and sends them to the
peer.  Because the initial fields are continuous functions of the node coordinates, the values a rank RECEIVES must
equal its own owner-side values slot by slot -- which only holds if both sides list the shared faces in the same
order with the same face-node numbering.  Prints MP_CPU_OK."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nebulasem_b200 import host  # noqa: E402


def main():
    decomp = sys.argv[1] if len(sys.argv) > 1 else "METIS"
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    case_dir = os.environ.get("MP_CPU_CASE")       # a case directory instead of the synthetic hill (its controls name the decomposition)
    if case_dir:
        s = host.Solver.open_case(case_dir, 0, rank, world)
    else:
        s = host.Solver.synthetic_part("hill3d", 8, 2, 4, 2, rank, world, decomp, (world, 1, 1))
    peer = 1 - rank
    assert s.peers() == [peer]
    faces = s.patch_faces(f"interMesh_{rank}_{peer}")
    NPF = s.NPF
    ks = (faces.astype(np.int64)[:, None] * NPF + np.arange(NPF)[None, :]).ravel()
    FO, FN, fI = s.u32("FO").astype(np.int64), s.u32("FN").astype(np.int64), s.f64("fI")
    ks = ks[FO[ks] < s.gALL]                 # a 2-D face uses fewer than NPF slots (the rest hold the sentinel)
    ok = bool((fI[ks] == 0.5).all()) and bool((FN[ks] >= s.gBCSfield).all())      # ghost faces carry fI = 0.5 (field.cpp:257-270)
    rho, U, T, p = s.state()
    cC = s.f64("cC").reshape(-1, 3)
    mine = np.concatenate([cC[FO[ks]], T[FO[ks], None], U[FO[ks]], s.f64("p_ref")[FO[ks], None]], axis=1)
    send = torch.tensor(mine)
    recv = torch.zeros_like(send)
    ops = [dist.P2POp(dist.isend, send, peer), dist.P2POp(dist.irecv, recv, peer)]
    for r in dist.batch_isend_irecv(ops):
        r.wait()
    got = recv.numpy()
    if case_dir:
        # curved elements (the cubed sphere): the two cells place a shared node through different edges, equal to rounding only
        scale = np.abs(mine[:, :3]).max()
        ok = ok and got.shape == mine.shape and np.abs(got[:, :3] - mine[:, :3]).max() <= 1e-12 * scale and np.allclose(got[:, 3:], mine[:, 3:], rtol=1e-9, atol=1e-9)
    else:
        ok = ok and got.shape == mine.shape and np.array_equal(got[:, :3], mine[:, :3]) and np.allclose(got, mine, rtol=0, atol=1e-12)
    # the union of the partitions is the global mesh, each cell exactly once
    ncells = torch.tensor([s.nBCS])
    dist.all_reduce(ncells)
    cg = torch.zeros(int(ncells), dtype=torch.int64)
    cg[torch.tensor(s.u32("cellGlobal").astype(np.int64))] += 1
    dist.all_reduce(cg)
    ok = ok and bool((cg == 1).all())
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MP_CPU_OK" if int(flag) == 1 else "MP_CPU_FAILED", decomp, flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
