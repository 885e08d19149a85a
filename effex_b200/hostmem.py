"""Host-side placement for the host-buffer path (fx_process_host): keep a rank's threads -- and
therefore the pinned staging memory it allocates afterwards -- on the NUMA node its GPU hangs off.
With 8 ranks pulling ~54 GB/s each over PCIe, remote-socket buffers halve the aggregate rate."""
from __future__ import annotations

import os


def gpu_cpu_affinity(device_index: int) -> list[int]:
    """CPUs NVML reports as local to the GPU (empty if NVML or the query is unavailable)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            # CUDA_VISIBLE_DEVICES remaps ordinals: resolve through the PCI bus id when torch is around
            handle = None
            try:
                import torch
                if torch.cuda.is_available():
                    bus = torch.cuda.get_device_properties(device_index).pci_bus_id
                    dom = torch.cuda.get_device_properties(device_index).pci_domain_id
                    dev = torch.cuda.get_device_properties(device_index).pci_device_id
                    handle = pynvml.nvmlDeviceGetHandleByPciBusId(f"{dom:08x}:{bus:02x}:{dev:02x}.0")
            except Exception:
                handle = None
            if handle is None:
                handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            n_cpu = os.cpu_count() or 1
            words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpu + 63) // 64)
            cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
            return [c for c in cpus if c < n_cpu]
        finally:
            pynvml.nvmlShutdown()
    except Exception:
        return []


def bind_to_gpu(device_index: int) -> list[int]:
    """Restrict this process to the GPU-local CPUs (intersection with the current mask).  Returns the
    CPUs now allowed; an empty NVML answer leaves the mask untouched."""
    if not hasattr(os, "sched_setaffinity"):
        return []
    allowed = os.sched_getaffinity(0)
    local = set(gpu_cpu_affinity(device_index)) & allowed
    if local and local != allowed:
        os.sched_setaffinity(0, local)
        return sorted(local)
    return sorted(allowed)
