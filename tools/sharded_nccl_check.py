"""torchrun --nproc-per-node N tools/sharded_nccl_check.py : ShardedWavelets over NCCL with the CUDA engine, checked
against the CPU oracle on rank 0 (scatter -> forward -> norm -> threshold -> inverse -> gather)."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pdwt_b200.sharded import ShardedWavelets, partition

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
B, Nr, Nc = 2 * world + 1, 512, 768
full = np.stack([(np.random.default_rng(i).standard_normal((Nr, Nc)) * 50 + 128).astype(np.float32) for i in range(B)])
S = ShardedWavelets(full if rank == 0 else None, "db7", 3)
S.forward()
n1 = S.norm1()
c5 = S.gather_coeff(5)
S.soft_threshold(10.0)
S.inverse()
rec = S.gather_image()
ok = True
if rank == 0:
    import oracle
    for i in range(B):
        O = oracle.Wavelets(full[i], "db7", 3)
        O.forward()
        ok &= abs(n1[i] - O.norm1()) <= 1e-5 * O.norm1()
        ok &= np.array_equal(c5[i].view(np.uint32), O.get_coeff(5).view(np.uint32))
        O.soft_threshold(10.0); O.inverse()
        ok &= np.array_equal(rec[i].view(np.uint32), O.get_image().view(np.uint32))
    print(f"sharded NCCL check (Layer C, device to device): world={world} B={B} blocks={partition(B, world)} native={S.native} "
          f"-> {'OK (bit-exact vs oracle)' if ok else 'MISMATCH'}", flush=True)
S.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
