"""Host-side placement for frame egress: keep a rank's threads -- and with them the page-locked
frame buffers it allocates afterwards (first touch) -- on the NUMA node its GPU hangs off.

A D2H copy that lands in the other socket's memory crosses the inter-socket link and shares it
with every other rank doing the same; with eight ranks copying 25 MB float frames per step that
link, not PCIe, sets the end-to-end frame rate.  Call `bind_to_gpu(device)` once per process,
before the renderer allocates pinned memory.  On hosts that expose no topology (a VM with one
virtual node: /sys reports numa_node = -1 for the GPU) nothing is changed and the result says so.
"""
import ctypes as C
import os

from . import _lib as L


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def gpu_numa_node(device):
    """(pci bus id, numa node or -1) of CUDA device `device` (cudaDeviceGetPCIBusId + sysfs)."""
    buf = C.create_string_buffer(32)
    if L.load().bhr_device_pci_bus_id(int(device), buf, 32) != 0:
        return None, -1
    bus = buf.value.decode().lower()
    try:
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            return bus, int(f.read().strip())
    except (OSError, ValueError):
        return bus, -1


def bind_to_gpu(device):
    """Restrict this process to the CPUs of the GPU's NUMA node.  Returns a dict describing what was
    found and done: {"pci", "numa_node", "nodes_online", "cpus_before", "cpus_after", "bound"}."""
    bus, node = gpu_numa_node(device)
    info = {"pci": bus, "numa_node": node, "bound": False}
    try:
        before = os.sched_getaffinity(0)
    except AttributeError:
        return info
    info["cpus_before"] = len(before)
    info["cpus_after"] = len(before)
    try:
        with open("/sys/devices/system/node/online") as f:
            info["nodes_online"] = f.read().strip()
    except OSError:
        info["nodes_online"] = None
    if node < 0:
        info["note"] = "the host exposes no NUMA node for this GPU (single virtual node): placement cannot be controlled"
        return info
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read()) & before
    except OSError:
        cpus = set()
    if cpus and cpus != before:
        os.sched_setaffinity(0, cpus)
        info["bound"] = True
        info["cpus_after"] = len(cpus)
    return info
