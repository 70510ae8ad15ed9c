#!/usr/bin/env python
"""Probe: does torch's symmetric memory (CUDA peer mappings over NVLink) rendezvous on this box?

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tools/symm_probe.py
"""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    t = symm_mem.empty(1 << 20, dtype=torch.float32, device=f"cuda:{local}")
    t.fill_(float(rank + 1))
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    print(rank, "rendezvous ok: ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal pads", [hex(p) for p in hdl.signal_pad_ptrs],
          "pad bytes", hdl.signal_pad_size, "multicast", hdl.has_multicast_support, hex(hdl.multicast_ptr) if hdl.has_multicast_support else None, flush=True)
    hdl.barrier(channel=0)
    peer = hdl.get_buffer((rank + 1) % world, t.shape, t.dtype)
    torch.cuda.synchronize()
    print(rank, "peer value", float(peer[0]), "expected", float((rank + 1) % world + 1), flush=True)
    hdl.barrier(channel=0)
    # inside a CUDA graph?
    g = torch.cuda.CUDAGraph()
    out = torch.zeros(1 << 20, device=f"cuda:{local}")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        out.copy_(peer)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            out.copy_(peer)
    g.replay()
    torch.cuda.synchronize()
    print(rank, "graph replay peer read ok", float(out[5]), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
