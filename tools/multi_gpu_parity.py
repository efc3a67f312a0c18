"""Dev tool (GPU box, torchrun): BASELINE.json configs[3] on a reduced range — the index range sharded over the ranks
(one process per GPU), VRF candidates min-reduced with an NCCL all_gather, labels and the reduced nonce compared
with a single-process oracle run.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/multi_gpu_parity.py"""
import importlib, json, os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
b2 = importlib.import_module("go-spacemesh_b200")
sharding = importlib.import_module("go-spacemesh_b200.sharding")

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
commitment = b2.commitment(bytes(range(32)), bytes(range(32, 64)))
ok_all = True
for n, start, count, num_labels in ((8192, 2**37 - 3000, 3000, 2**37), (8192, 2**32 - 1500, 3001, 2048), (2, 77, 400001, 2**18)):
    diff = b2.vrf_difficulty(num_labels)
    s, k = sharding.shard_range(start, count, world, rank)
    labels, vrf = b2.labels_range(commitment, n, s, k, provider=local, vrf_difficulty_=diff)
    best = sharding.allgather_vrf(vrf, device=torch.device("cuda", local))
    sizes = [sharding.shard_range(start, count, world, r)[1] for r in range(world)]
    buf = torch.zeros((max(sizes), 16), dtype=torch.uint8, device="cuda")
    buf[:k] = torch.from_numpy(labels).cuda()
    parts = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    if rank == 0:
        from oracle import pyoracle as orc
        got = np.concatenate([p[:sz].cpu().numpy() for p, sz in zip(parts, sizes)])
        exp, found, idx, l32 = orc.c_labels_range(commitment, n, start, count, diff)
        ok = bool((got == exp).all()) and best == ((idx, l32) if found else None)
        ok_all &= ok
        print(json.dumps(dict(n=n, start=start, count=count, world=world, labels_ok=bool((got == exp).all()), vrf=str(best)[:60], vrf_ok=best == ((idx, l32) if found else None))), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok_all else 1)
