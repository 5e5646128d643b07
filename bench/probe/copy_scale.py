"""bench/probe/copy_scale.py -- what the host can feed: the e2e leg's copies WITHOUT any solve.
One process per GPU (torchrun, like bench.py); every rank streams the packed tick buffers of bench.py's e2e leg
(one H2D of 917,504 B and one D2H of 2,293,760 B per batch of 4,096 problems, page-locked host memory, several
streams in flight) for a fixed number of batches; rank 0 prints the aggregate rate in batches/s, GB/s and the
'solves/s' the copies alone would allow.  If this number is close to the e2e figure at N GPUs, the e2e leg is bound
by the host's memory / PCIe fabric, not by the solver or its host code."""
import json, os, sys, time
import torch
import torch.distributed as dist

H2D = 917504; D2H = 2293760; B = 4096


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    S = 8; K = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    hin = [torch.empty(H2D, dtype=torch.uint8).pin_memory() for _ in range(S)]
    hout = [torch.empty(D2H, dtype=torch.uint8).pin_memory() for _ in range(S)]
    din = [torch.empty(H2D, dtype=torch.uint8, device=dev) for _ in range(S)]
    dout = [torch.empty(D2H, dtype=torch.uint8, device=dev) for _ in range(S)]
    st = [torch.cuda.Stream() for _ in range(S)]
    def run(n):
        for j in range(n):
            s = j % S
            with torch.cuda.stream(st[s]):
                din[s].copy_(hin[s], non_blocking=True)
                hout[s].copy_(dout[s], non_blocking=True)
        torch.cuda.synchronize()
    run(200)
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter(); run(K); t = time.perf_counter() - t0
    tt = torch.tensor([t], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        tmax = float(tt.item())
        print(json.dumps(dict(probe="copy_scale", n_gpus=world, batches_per_gpu=K, seconds=tmax,
                              gbytes_per_s=world * K * (H2D + D2H) / tmax / 1e9,
                              solves_per_s_copies_alone=world * K * B / tmax)))
    if world > 1: dist.destroy_process_group()


main()
