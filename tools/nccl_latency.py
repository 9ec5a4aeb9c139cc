"""Latency of the exchange a single proof split across GPUs would need in EVERY sumcheck round:
an all-reduce / all-gather of the round's partial sums (3 field elements = 48 or 96 bytes) over NCCL/NVLink.
Run under torchrun with N ranks; prints one JSON line on rank 0.  Device-timed (CUDA events), max over ranks."""
import json
import os

import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
world = dist.get_world_size()
res = {"n_gpus": world}
for nbytes in (48, 96):
    x = torch.zeros(nbytes // 4, dtype=torch.int32, device="cuda")
    out = torch.zeros(world * (nbytes // 4), dtype=torch.int32, device="cuda")
    for name, fn in (("all_reduce", lambda: dist.all_reduce(x, op=dist.ReduceOp.BXOR if False else dist.ReduceOp.SUM)),
                     ("all_gather", lambda: dist.all_gather_into_tensor(out, x))):
        for _ in range(50):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        iters = 2000
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[f"{name}_{nbytes}B_us"] = round(float(t.item()), 2)
        # the same inside one CUDA graph (no per-call launch cost on the host)
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            try:
                with torch.cuda.graph(g, stream=s):
                    for _ in range(100):
                        fn()
                torch.cuda.synchronize()
                dist.barrier()
                e0.record(s)
                for _ in range(10):
                    g.replay()
                e1.record(s)
                torch.cuda.synchronize()
                t = torch.tensor([e0.elapsed_time(e1) / 1000 * 1e3], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                res[f"{name}_{nbytes}B_graph_us"] = round(float(t.item()), 2)
            except Exception as ex:  # graph capture of NCCL not available in this build
                res[f"{name}_{nbytes}B_graph_us"] = None
                res["graph_error"] = repr(ex)[:120]
if dist.get_rank() == 0:
    print(json.dumps(res), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/nccl_latency_{world}gpu.json", "w"))
torch.cuda.synchronize()
dist.barrier()
# the captured graphs still hold the communicator: leave without tearing it down (a destroy here can block)
os._exit(0)
