"""torchrun probe (scratch): is torch's symmetric memory (cuMem peer mappings over NVLink) usable on this box?"""
import os, sys, time
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
local = int(os.environ["LOCAL_RANK"]); dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
try:
    buf = symm_mem.empty((world, 1024, 54), dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
    print(rank, "rendezvous ok; ptrs", [hex(p) for p in hdl.buffer_ptrs], "multicast_ptr", hex(hdl.multicast_ptr or 0),
          "signal pad", hdl.signal_pad_size, flush=True)
    buf.zero_()
    torch.cuda.synchronize(); dist.barrier()
    mine = torch.full((1024, 54), float(rank + 1), device=dev)
    for p in range(world):
        peer = hdl.get_buffer(p, (world, 1024, 54), torch.float32)
        peer[rank].copy_(mine)                      # P2P store into every rank's slot for me
    hdl.barrier(channel=0)
    torch.cuda.synchronize()
    ok = all(bool((buf[r] == r + 1).all()) for r in range(world))
    print(rank, "p2p all-gather by peer stores:", ok, flush=True)
    # timing: 7 peer copies of 2.8 MB + barrier
    big = symm_mem.empty((world, 13000, 54), dtype=torch.float32, device=dev)
    h2 = symm_mem.rendezvous(big, dist.group.WORLD)
    src = torch.randn(13000, 54, device=dev)
    peers = [h2.get_buffer(p, (world, 13000, 54), torch.float32) for p in range(world)]
    for it in range(3):
        torch.cuda.synchronize(); dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for p in range(world):
            peers[p][rank].copy_(src)
        h2.barrier(channel=0)
        b.record(); b.synchronize()
        t_p2p = a.elapsed_time(b)
        out = torch.empty((world * 13000, 54), device=dev)
        torch.cuda.synchronize(); dist.barrier()
        a.record(); dist.all_gather_into_tensor(out, src); b.record(); b.synchronize()
        t_nccl = a.elapsed_time(b)
        torch.cuda.synchronize(); dist.barrier()
        a.record(); h2.barrier(channel=0); b.record(); b.synchronize()
        t_bar = a.elapsed_time(b)
        if rank == 0:
            print("iter %d: peer-store all-gather %.3f ms, NCCL all_gather %.3f ms, symm barrier alone %.3f ms" % (it, t_p2p, t_nccl, t_bar), flush=True)
except Exception as e:
    import traceback; traceback.print_exc()
    print(rank, "symmetric memory unavailable:", repr(e), flush=True)
dist.barrier(); dist.destroy_process_group()
