"""NUMA placement of a rank: run on (and allocate the pinned I/O buffers from) the CPU socket the rank's GPU hangs off.

On an 8-GPU host every rank's host-buffer step writes tens of MB per step into page-locked host memory over PCIe; when
those pages live on the other socket the traffic crosses the inter-socket link and the ranks contend for it (round-1
scaling of the host-buffer path: 0.32 at 8 GPUs).  `bind_to_gpu(device)` pins the calling process to the CPUs of the GPU's
NUMA node BEFORE the engine allocates its pinned buffers (first-touch places the pages on that node).  Best effort: any
failure leaves the process unbound and is reported in the returned dict.
"""
import os
import subprocess


def _pci_bus_id(device):
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = int(visible.split(",")[device]) if visible and visible.split(",")[device].strip().isdigit() else device
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        return bus.decode() if isinstance(bus, bytes) else bus
    except Exception:
        out = subprocess.run(["nvidia-smi", "-i", str(device), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip()
        if not out:
            raise RuntimeError("no PCI bus id for device %d" % device)
        return out


def _parse_cpulist(text):
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(device):
    """(node, cpus) of the NUMA node the GPU is attached to; node < 0 when the platform does not say."""
    bus = _pci_bus_id(device).lower()
    if bus.count(":") == 2 and len(bus.split(":")[0]) == 8:        # nvml prints an 8-digit domain, sysfs uses 4
        bus = bus[4:]
    with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
        node = int(f.read().strip())
    if node < 0:
        return node, []
    with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
        return node, _parse_cpulist(f.read())


def bind_to_gpu(device):
    """Pins this process to the CPUs of `device`'s NUMA node.  Returns {"node", "cpus", "bound", "error"}."""
    info = {"node": None, "cpus": 0, "bound": False, "error": None}
    try:
        node, cpus = gpu_numa_node(device)
        info["node"] = node
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        info["cpus"] = len(allowed)
        if node >= 0 and allowed:
            os.sched_setaffinity(0, allowed)
            info["bound"] = True
    except Exception as e:                                          # noqa: BLE001 -- placement is an optimisation, never fatal
        info["error"] = "%s: %s" % (type(e).__name__, e)
    return info
