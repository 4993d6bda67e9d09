"""all_to_all_single bandwidth at the sharded store's message sizes (run under torchrun)"""
import os, time, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
for rows in (6_500_000,):
    for width in (64, 1):
        x = torch.randn(rows, width, device=dev); y = torch.empty_like(x)
        splits = [rows // world] * world; splits[-1] += rows - sum(splits)
        for _ in range(3): dist.all_to_all_single(y, x, splits, splits)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): dist.all_to_all_single(y, x, splits, splits)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        remote = rows * width * 4 * (world - 1) / world
        if rank == 0: print(f"world {world} rows {rows} width {width}: {ms:.3f} ms, remote {remote/1e9:.3f} GB -> {remote/ms/1e6:.1f} GB/s per direction", flush=True)
dist.destroy_process_group()
