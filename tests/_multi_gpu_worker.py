"""Worker of tests/test_gpu_multi.py: launched by torch.distributed.run, one rank per GPU (NCCL).  Every rank evolves
its shard of the chains; rank 0 writes the gathered records."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(out_path):
    import torch
    import torch.distributed as dist
    from time_crystal_tensor_network_b200 import engine as eng
    from time_crystal_tensor_network_b200.sharding import run_sharded_ensemble
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
    L, R, n = 14, 7, 8                      # 7 chains over 2 ranks: ragged shards (4 + 3)
    hs = np.array([eng.disorder_fields(L, 0.3, 2000 + r) for r in range(R)])
    rec = run_sharded_ensemble(L, 1.0, 1.0, hs, n, rank=rank, world_size=world, device=local, epsilon=0.12,
                               chi_max=24, mode='tebd', svd_min=1e-12, trunc_cut=1e-10)
    if rank == 0:
        np.savez(out_path, **rec)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main(sys.argv[1])
