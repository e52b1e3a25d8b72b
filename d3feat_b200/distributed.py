"""Multi-GPU: whole point-cloud fragments shard across ranks (one process per GPU); the only exchange step is
ONE all-gather of the per-fragment descriptors at the end (NCCL over NVLink/NVSwitch on the B200 box, gloo in
the CPU tests). The reference is single-process: fragments are independent units, so there is no
data-path collective inside the pyramid or the encoder.
"""
import torch
import torch.distributed as dist


def shard_fragments(n_fragments, rank, world_size):
    """Fragment f -> rank f mod world_size (round robin keeps per-rank point counts balanced for
    similarly sized fragments). Returns the list of fragment ids owned by `rank`."""
    return [f for f in range(n_fragments) if f % world_size == rank]


def all_gather_descriptors(local_desc, local_rows_per_fragment, group=None):
    """local_desc: float32[sum(rows), D] descriptors of this rank's fragments (stacked);
    local_rows_per_fragment: list[int]. Returns (desc_all [R, D], rows_all list[int], owner list[int]) with
    the fragments of rank 0 first, then rank 1, ... (each rank's own order preserved).

    Two collectives: a tiny all-gather of the row counts, then one all-gather of the descriptors padded to
    the largest per-rank row count.
    """
    if not dist.is_available() or not dist.is_initialized():
        return local_desc, list(local_rows_per_fragment), [0] * len(local_rows_per_fragment)
    world = dist.get_world_size(group)
    dev = local_desc.device
    D = local_desc.shape[1]
    n_local = len(local_rows_per_fragment)
    meta = torch.tensor([local_desc.shape[0], n_local], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    rows = [int(m[0]) for m in metas]
    nfrag = [int(m[1]) for m in metas]
    max_rows, max_frag = max(rows), max(nfrag)
    frag_rows = torch.zeros((max(max_frag, 1),), dtype=torch.int64, device=dev)
    if n_local:
        frag_rows[:n_local] = torch.tensor(local_rows_per_fragment, dtype=torch.int64, device=dev)
    frag_all = [torch.zeros_like(frag_rows) for _ in range(world)]
    dist.all_gather(frag_all, frag_rows, group=group)
    padded = torch.zeros((max(max_rows, 1), D), dtype=local_desc.dtype, device=dev)
    padded[:local_desc.shape[0]] = local_desc
    gathered = torch.empty((world, max(max_rows, 1), D), dtype=local_desc.dtype, device=dev)
    dist.all_gather_into_tensor(gathered.view(-1, D), padded, group=group) if dev.type == "cuda" else \
        dist.all_gather(list(gathered.unbind(0)), padded, group=group)
    desc_all = torch.cat([gathered[r, :rows[r]] for r in range(world)], 0)
    rows_all, owner = [], []
    for r in range(world):
        rows_all += [int(x) for x in frag_all[r][:nfrag[r]]]
        owner += [r] * nfrag[r]
    return desc_all, rows_all, owner
