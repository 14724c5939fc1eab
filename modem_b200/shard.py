"""Frame-level data parallelism: windows are independent (one Decoder object per invocation,
decode.cc:592), so a batch is block-partitioned over ranks with no data-path collective; the only exchange is one
gather of the decoded payload bytes at the end (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(n_total, rank, world):
    """Contiguous block partition; the first n_total % world ranks get one extra window."""
    base, extra = divmod(n_total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_payload(local, n_total=None, group=None):
    """local: uint8 tensor [n_local, 5380] (CUDA for nccl, CPU for gloo).  Returns the [n_total, 5380] tensor in global
    window order on every rank (all_gather: equal shards go through all_gather_into_tensor, ragged ones are padded)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    counts = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    all_counts = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts, group=group)
    all_counts = [int(c.item()) for c in all_counts]
    mx = max(all_counts)
    if all(c == mx for c in all_counts):
        out = torch.empty((world * mx, local.shape[1]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    pad = torch.zeros((mx, local.shape[1]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, all_counts)], dim=0)
