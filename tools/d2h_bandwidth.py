#!/usr/bin/env python
"""Host-side ceiling of the end-to-end path: device-to-host bandwidth into pinned memory, one process per GPU, all at once.

bench.py's e2e loop returns one fp32 [3,1080,1920] image (24.9 MB) per frame.  N ranks at ~830 frames/s each ask the host
for N x 20.7 GB/s of DMA writes; this tool measures what the box can take, with nothing but cudaMemcpyAsync in the loop:

  python tools/d2h_bandwidth.py                                                  # one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/d2h_bandwidth.py

Per rank: `--buffers` pinned host buffers of `--mb` MB, copies issued round-robin on one stream for `--seconds`; all ranks start
together (barrier).  Rank 0 prints one JSON line: per-rank GB/s, the aggregate, and the frames/s of 24.9 MB images it allows.
Also prints what the OS says about NUMA (nodes visible to this process, the GPU's node, allowed CPUs).
"""
import argparse
import glob
import json
import os
import time

import torch
import torch.distributed as dist


def numa_info(local):
    info = {"cpus_allowed": len(os.sched_getaffinity(0))}
    try:
        info["nodes_online"] = open("/sys/devices/system/node/online").read().strip()
    except OSError:
        info["nodes_online"] = None
    for p in ("/sys/fs/cgroup/cpuset.mems.effective", "/sys/fs/cgroup/cpuset/cpuset.effective_mems"):
        if os.path.exists(p):
            info["cpuset_mems"] = open(p).read().strip()
            break
    try:
        pr = torch.cuda.get_device_properties(local)
        dev = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        hits = glob.glob(f"/sys/bus/pci/devices/{dev}/numa_node")
        info["gpu_pci"] = dev
        info["gpu_numa_node"] = open(hits[0]).read().strip() if hits else None
    except Exception as ex:
        info["gpu_numa_node"] = repr(ex)
    return info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=float, default=24.8832)
    ap.add_argument("--buffers", type=int, default=4)
    ap.add_argument("--seconds", type=float, default=3.0)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    n = int(a.mb * 1e6 / 4)
    src = torch.rand(n, device=dev)
    host = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(a.buffers)]
    for h in host:
        h.copy_(src, non_blocking=True)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    copies = 0
    while time.perf_counter() - t0 < a.seconds:
        for h in host:
            h.copy_(src, non_blocking=True)
            copies += 1
        torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    gbs = copies * n * 4 / dt / 1e9
    row = torch.tensor([gbs], dtype=torch.float64, device=dev)
    rows = [row]
    if world > 1:
        rows = [torch.zeros_like(row) for _ in range(world)]
        dist.all_gather(rows, row)
    if rank == 0:
        per = [float(r[0]) for r in rows]
        print(json.dumps({"tool": "d2h_bandwidth", "n_gpus": world, "buffer_mb": a.mb, "per_rank_gbs": per, "aggregate_gbs": sum(per),
                          "frames_per_s_of_24.9MB_images": sum(per) * 1e9 / (3 * 1080 * 1920 * 4), "numa": numa_info(local)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
