"""CPU / memory affinity of a rank: run on the cores of the NUMA node its GPU hangs off, so that the page-locked host buffers it
allocates afterwards (first touch) sit next to that GPU's PCIe root. What an MPI launcher's --bind-to / a job script's numactl does
for the reference; matters for the host-memory entry points when several ranks share a multi-socket host. Best effort: any failure
(no sysfs, one NUMA node, a cpuset that excludes the node) leaves the affinity untouched."""
import os


def _cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(device_index):
    """-> (numa node or None, set of local cpus or empty set) of CUDA device `device_index` (sysfs of its PCI function)."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        with open(base + "/numa_node") as f:
            node = int(f.read().strip())
        with open(base + "/local_cpulist") as f:
            cpus = _cpulist(f.read())
        return (node if node >= 0 else None), cpus
    except Exception:
        return None, set()


def bind_to_gpu(device_index):
    """Restricts this process to the CPUs local to the GPU (when that is a proper, non-empty subset of what it may use).
    -> {"numa_node": n or None, "cpus": how many it runs on now, "bound": bool}"""
    node, local = gpu_numa_node(device_index)
    try:
        allowed = os.sched_getaffinity(0)
    except Exception:
        return {"numa_node": node, "cpus": None, "bound": False}
    target = allowed & local
    bound = False
    if target and target != allowed:
        try:
            os.sched_setaffinity(0, target)
            bound = True
        except Exception:
            bound = False
    return {"numa_node": node, "cpus": len(target if bound else allowed), "bound": bound}
