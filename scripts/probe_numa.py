#!/usr/bin/env python
"""Pinned-memory H2D / D2H / bidirectional bandwidth of GPU 0 with the process bound to each NUMA node (where should
the host buffers of the end-to-end path live?).  python scripts/probe_numa.py"""
import glob
import os
import subprocess
import time

import torch


def cpus_of(node):
    txt = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
    out = []
    for part in txt.split(","):
        a, _, b = part.partition("-")
        out += list(range(int(a), int(b or a) + 1))
    return out


def bw(n_bytes=1 << 30, reps=5):
    h_in = torch.empty(n_bytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n_bytes, dtype=torch.uint8).pin_memory()
    h_in.fill_(1); h_out.fill_(2)   # first touch under the current affinity
    d_a = torch.empty(n_bytes, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(n_bytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for name in ("h2d", "d2h", "both"):
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(reps):
            if name in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_a.copy_(h_in, non_blocking=True)
            if name in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out.copy_(d_b, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        res[name] = round(n_bytes * reps * (2 if name == "both" else 1) / dt / 1e9, 1)
    return res


def main():
    prop = torch.cuda.get_device_properties(0)
    bus = f"{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0" if hasattr(prop, "pci_bus_id") else None
    node = None
    if bus and os.path.exists(f"/sys/bus/pci/devices/{bus}/numa_node"):
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
    nodes = sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
    print("gpu0 pci", bus, "numa_node", node, "nodes", nodes, "affinity now", len(os.sched_getaffinity(0)), "cpus")
    try:
        print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
    except Exception as e:
        print("topo failed", e)
    print("unbound:", bw())
    all_cpus = os.sched_getaffinity(0)
    for nd in nodes:
        cp = set(cpus_of(nd)) & all_cpus
        if not cp:
            continue
        os.sched_setaffinity(0, cp)
        print(f"bound to node {nd} ({len(cp)} cpus):", bw())
    os.sched_setaffinity(0, all_cpus)


if __name__ == "__main__":
    main()
