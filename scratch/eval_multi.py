"""torchrun worker (scratch): the cfg-5 eval job at N GPUs for several chunk counts, with the parity flag."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
local = int(os.environ["LOCAL_RANK"]); dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
cases = [("nccl", int(c)) for c in (sys.argv[1] if len(sys.argv) > 1 else "1,2").split(",")] + [("auto", 1)]
if os.environ.get("SP_EVAL_NO_MULTICAST") is None:
    cases.append(("unicast", 1))
for transport, chunks in cases:
    if transport == "unicast":
        os.environ["SP_EVAL_NO_MULTICAST"] = "1"
        transport = "auto"
    r = bench.eval_job_numbers(dev, world, rank, chunks=chunks, reps=9, transport=transport)
    if rank == 0:
        print(json.dumps({k: r[k] for k in ("n_gpus", "transport", "chunks_per_rank", "ms", "persons_per_s", "kept_after_nms", "table_checksum", "matches_single_device", "compact_ms")}), flush=True)
dist.barrier(); dist.destroy_process_group()
