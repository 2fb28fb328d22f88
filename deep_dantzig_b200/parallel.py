"""Multi-GPU plumbing of the hot path: instances shard trivially, so the solve path has no
collective; torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used only for the final label/status gather and
for the gradient all-reduce of data-parallel classifier training (one flat buffer, one call per step)."""
import torch
import torch.distributed as dist


def bind_to_gpu_numa_node(device_index):
    """Pin this process (one per GPU) to the CPUs of the NUMA node its GPU hangs off, so that pinned host buffers are
    first-touched on that node and host<->device copies do not cross the socket interconnect.  Best effort: returns the
    node number, or None when the topology cannot be read (single-socket host, container without sysfs, ...)."""
    import os
    try:
        props = torch.cuda.get_device_properties(device_index)
        bdf = '%04x:%02x:%02x.0' % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        with open('/sys/bus/pci/devices/%s/numa_node' % bdf) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open('/sys/devices/system/node/node%d/cpulist' % node) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(','):
            lo, _, hi = part.partition('-')
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:        # no CUDA device, no sysfs, unknown torch attribute, ...: placement is best effort
        return None


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(num_items, rank=None, world_size=None):
    """Contiguous block [lo, hi) of the global instance index range owned by `rank` (sizes differ by at most one).
    With Philox generation the counter is the global index, so results do not depend on the number of ranks."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, rem = divmod(int(num_items), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_round_robin(items, rank=None, world_size=None):
    """Round-robin split of a work list (sweep cells x chunks, BASELINE.json config 3)."""
    if rank is None or world_size is None:
        rank, world_size = world()
    return [it for k, it in enumerate(items) if k % world_size == rank]


def gather_to_rank0(tensor, sizes=None):
    """Gather per-rank result blocks (labels / status / objective) on rank 0 along dim 0; other ranks get None.
    Blocks may differ in length by one (shard_range); they are padded to a common length for the collective."""
    rank, ws = world()
    if ws == 1:
        return tensor
    n_local = torch.tensor([tensor.shape[0]], device=tensor.device)
    counts = [torch.zeros_like(n_local) for _ in range(ws)]
    dist.all_gather(counts, n_local)
    counts = [int(c.item()) for c in counts]
    nmax = max(counts)
    pad = torch.zeros((nmax,) + tuple(tensor.shape[1:]), dtype=tensor.dtype, device=tensor.device)
    pad[: tensor.shape[0]] = tensor
    out = [torch.empty_like(pad) for _ in range(ws)] if rank == 0 else None
    dist.gather(pad, out, dst=0)
    if rank != 0:
        return None
    return torch.cat([o[:k] for o, k in zip(out, counts)], 0)


def allreduce_gradients(model):
    """Sum the gradients of all ranks with ONE collective over a flat fp32 buffer (the model has 1.2k-12k parameters,
    so the all-reduce is latency-bound: SURVEY.md section 5).  Matches the reference's sum-over-batch accumulation
    (train.py:60-66) when every rank holds a slice of the batch."""
    rank, ws = world()
    params = [q for q in model.parameters() if q.grad is not None]
    if ws == 1 or not params:
        return
    flat = torch.cat([q.grad.reshape(-1) for q in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    off = 0
    for q in params:
        k = q.numel()
        q.grad.copy_(flat[off:off + k].view_as(q.grad))
        off += k


def broadcast_parameters(model, src=0):
    rank, ws = world()
    if ws == 1:
        return
    flat = torch.cat([q.data.reshape(-1) for q in model.parameters()])
    dist.broadcast(flat, src=src)
    off = 0
    for q in model.parameters():
        k = q.numel()
        q.data.copy_(flat[off:off + k].view_as(q.data))
        off += k
