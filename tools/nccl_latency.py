#!/usr/bin/env python
"""Latency of the tiny all-gather the sharded run needs (k doubles per rank), via the library's
own NCCL binding and via torch.distributed, timed with CUDA events.  torchrun --nproc-per-node G."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from gleipnir_b200 import _lib as L
    from gleipnir_b200.sharded import NcclComm
    lib = L.load()
    comm = NcclComm()
    world, rank = comm.world, comm.rank
    for k in (4, 1000, 2000, 100000):
        send = torch.randn(k, dtype=torch.float64, device="cuda")
        recv = torch.empty(world * k, dtype=torch.float64, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        for which in ("native", "torch"):
            def once():
                if which == "native":
                    L.check(lib.gl_comm_all_gather_f64(comm.h, send.data_ptr(), recv.data_ptr(), k, st))
                else:
                    dist.all_gather_into_tensor(recv, send)
            for _ in range(20):
                once()
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            n = 500
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                once()
            e1.record()
            e1.synchronize()
            if rank == 0:
                print("all_gather k=%6d doubles/rank world=%d %-6s: %.1f us per call" %
                      (k, world, which, 1e3 * e0.elapsed_time(e1) / n), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
