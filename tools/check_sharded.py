"""2+ GPU check (run under torchrun): the distributed sample-sort build reproduces the single-GPU build
(codes, permutation, nodes) and the sharded traversal reproduces the single-GPU result bit for bit."""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rakau_b200 as rk
from rakau_b200.distributed import ShardedTree

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
chunk = 1 << 16
per = ((N + world - 1) // world + chunk - 1) // chunk * chunk
first = min(rank * per, N); count = max(0, min(N, first + per) - first)
m, x, y, z = rk.plummer(N, first, count, chunk=chunk)
st = ShardedTree(dist, dev)
sh = [torch.from_numpy(a).to(dev) for a in (x, y, z, m)]
bi = st.build(*sh, first_index=first)
for _rep in range(int(os.environ.get('REPS', '0'))):
    bi = st.build(*sh, first_index=first)
info, out = st.acc_pot(0, 0.75)  # peer-memory exchange (pipelined copy-engine pushes)
out = [o.clone() for o in out]
print(rank, "peer memory exchange:", st._peer is not False and st._peer is not None, flush=True)
imb = st.rebalance()
out2 = [torch.zeros(N, dtype=torch.float32, device=dev) for _ in range(3)]
info2, _ = st.acc_pot(0, 0.75, out=out2)  # collective exchange (padded all-gather)
assert info["interactions"] > 0 and info2["interactions"] > 0
ok = True
if rank == 0:
    fm, fx, fy, fz = rk.plummer(N, 0, N, chunk=chunk)
    ref = rk.Octree(device=local); ref.build(fx, fy, fz, fm)
    ok &= bool((ref.codes() == st.tree.codes()).all())
    ok &= bool((ref.perm(0) == st.tree.perm(0)).all())
    a, b = ref.nodes(), st.tree.nodes()
    ok &= len(a) == len(b) and bool((a == b).all())
    ro = ref.acc_pot(0, 0.75)
    for j in range(3):
        ok &= bool((ro[j] == out[j].cpu().numpy()).all()) and bool((ro[j] == out2[j].cpu().numpy()).all())
    print("sharded build == single-GPU build:", ok, "| box", st.tree.box_size, ref.box_size, "| nodes", len(a),
          "| cost imbalance after rebalance", imb, "| build info", bi.asdict())
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
