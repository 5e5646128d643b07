"""bench/config3.py -- BASELINE config 3: one batch of 65,536 N=20 problems split contiguously over the GPUs
of one node (strong scaling).  Host buffers in, host buffers out: wall time from the first submit to the last
result gathered on every rank.  One process per GPU:
    python bench/config3.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 \
        bench/config3.py
Each rank pipelines its slice in chunks of 4,096 over K handles (mpc_b200_track_submit / _wait); there is no
collective on the solve path, only the final all-gather of the first controls.
argv: [problems [chunk [ctas_in_flight_per_gpu [max_iter]]]].  A one-shot batch is latency-bound: its wall time
is the slowest problem's (the ~0.05 % that run into the iteration cap with a backtracking line search take
~5 ms at max_iter = 100), which is why the throughput benchmark keeps many batches in flight."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from mpc_ros_b200 import capi  # noqa: E402
from mpc_ros_b200.sharding import rank_slice, gather_slices  # noqa: E402
from bench import gen_py  # noqa: E402

TOTAL = 65536
CHUNK = 4096


def main():
    total = int(sys.argv[1]) if len(sys.argv) > 1 else TOTAL
    CHUNK = int(sys.argv[2]) if len(sys.argv) > 2 else globals()["CHUNK"]
    ctas_total = int(sys.argv[3]) if len(sys.argv) > 3 else 296
    max_iter = int(sys.argv[4]) if len(sys.argv) > 4 else 100
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lo, hi = rank_slice(total, rank, world)
    n = hi - lo
    g = gen_py.problems(20261018 + 3, total)
    M = g["M"]; N = 20
    prm = capi.yaml_default_params(); prm.delay_mode = 0; prm.max_iter = max_iter
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()  # noqa: E731
    chunks = [(a, min(a + CHUNK, n)) for a in range(0, n, CHUNK)]
    K = max(1, min(len(chunks), 32))
    hs = [capi.Solver(prm, CHUNK, local) for _ in range(K)]
    for s in hs:
        s.set_option("max_ctas", max(4, min(148, -(-ctas_total // K))))
    # this rank's slice, one set of page-locked buffers per chunk (inputs and outputs stay on the host)
    bufs = []
    for a, b in chunks:
        B = b - a
        bufs.append(dict(B=B, wx=pin(g["wx"][:, lo + a:lo + b]), wy=pin(g["wy"][:, lo + a:lo + b]), pose=pin(g["pose"][:, lo + a:lo + b]),
                         vel=pin(g["vel"][:, lo + a:lo + b]), u0=torch.zeros((2, B), dtype=torch.float64).pin_memory(),
                         pred=torch.zeros((3 * N, B), dtype=torch.float64).pin_memory(),
                         stat=torch.zeros(B, dtype=torch.int32).pin_memory(), kkt=torch.zeros(B, dtype=torch.float64).pin_memory()))

    tm = {}

    def run():
        ta = time.perf_counter()
        for j, c in enumerate(bufs):
            s = hs[j % K]
            s.track_wait()
            s.track_submit_raw(c["B"], M, c["wx"].numpy(), c["wy"].numpy(), c["pose"].numpy(), c["vel"].numpy(), c["u0"].numpy(),
                               c["pred"].numpy(), status=c["stat"].numpy(), kkt=c["kkt"].numpy())
        tb = time.perf_counter()
        for s in hs:
            s.track_wait()
        tc = time.perf_counter()
        w = torch.cat([c["u0"][0] for c in bufs]).to(dev) if bufs else torch.zeros(0, dtype=torch.float64, device=dev)
        full = gather_slices(w, total, rank, world, device=dev)        # the final gather: every rank ends with all w_0
        out = full.cpu()
        tm.update(submit_ms=(tb - ta) * 1e3, wait_ms=(tc - tb) * 1e3, gather_ms=(time.perf_counter() - tc) * 1e3)
        return out

    run()                                   # warm-up (clocks, allocator, NCCL communicator)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    full = run()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    conv = sum(int(((c["stat"].numpy() == 1) & (c["kkt"].numpy() <= 1e-8)).sum()) for c in bufs)
    tt = torch.tensor([t1 - t0], dtype=torch.float64, device=dev); cc = torch.tensor([float(conv)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX); dist.all_reduce(cc, op=dist.ReduceOp.SUM)
    if rank == 0:
        print(json.dumps(dict(config=3, problems=total, n_gpus=world, wall_ms=float(tt.item()) * 1e3,
                              converged=int(cc.item()), converged_solves_per_s=float(cc.item() / tt.item()),
                              gathered=int(full.numel()), max_iter=max_iter, chunk=CHUNK, handles_per_gpu=K, rank0_breakdown=tm,
                              note="host buffers in and out, first submit to last gather complete, max over ranks")))
    for s in hs:
        s.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
