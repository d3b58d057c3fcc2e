"""Scene sharding across the GPUs of one box: independent units, no data-path collective."""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple


def rank_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) when not launched by torchrun."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def shard_round_robin(n_units: int, rank: int, world_size: int) -> List[int]:
    """Indices of the scenes / reference views owned by ``rank`` (round-robin, as SURVEY.md §8e prescribes)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size %d/%d" % (rank, world_size))
    return list(range(rank, n_units, world_size))


def shard_pairs(pairs: Sequence, rank: int, world_size: int):
    """Rows of a filter pair list owned by ``rank``: every GPU holds the full depth stack, reference views are split."""
    idx = shard_round_robin(len(pairs), rank, world_size)
    return [pairs[i] for i in idx], idx


def _parse_cpulist(text: str):
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_local_cpus(device_index: int):
    """CPUs on the NUMA node the GPU hangs off (``/sys/bus/pci/devices/<bdf>/local_cpulist``), or None if unknown."""
    import os
    import subprocess
    bdf = None
    try:
        import torch
        pr = torch.cuda.get_device_properties(device_index)
        if hasattr(pr, "pci_bus_id"):
            bdf = "%04x:%02x:%02x.0" % (getattr(pr, "pci_domain_id", 0), pr.pci_bus_id, getattr(pr, "pci_device_id", 0))
    except Exception:
        bdf = None
    if bdf is None:
        try:
            out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(device_index)],
                                 capture_output=True, text=True, timeout=10).stdout.strip()
            bdf = out.lower()[-12:] if out else None  # 00000000:1B:00.0 -> 0000:1b:00.0
        except Exception:
            bdf = None
    if not bdf:
        return None
    path = "/sys/bus/pci/devices/%s/local_cpulist" % bdf
    try:
        cpus = _parse_cpulist(open(path).read())
    except OSError:
        return None
    allowed = os.sched_getaffinity(0)
    cpus = [c for c in cpus if c in allowed]
    return cpus or None


def bind_to_gpu_numa_node(device_index: int):
    """Pin the calling process to the CPUs local to its GPU, so that the pinned host buffers it allocates afterwards
    (first touch) and its copy-engine submissions stay on the GPU's side of the socket interconnect.  With one process
    per GPU feeding 2.4 GB of features per step this is what keeps eight concurrent host->device streams from sharing
    one socket's memory controllers.  Returns the CPU list it bound to, or None when the topology is not exposed."""
    import os
    cpus = gpu_local_cpus(device_index)
    if cpus:
        try:
            os.sched_setaffinity(0, cpus)
        except OSError:
            return None
    return cpus
