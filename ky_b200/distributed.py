"""Multi-GPU plumbing: one process per GPU, samples-per-pixel split across ranks, one film reduce.

The path shards by SAMPLE INDEX: rank r renders sample indices
[r * spp / N, (r + 1) * spp / N) of EVERY pixel into a rank-local, UNCLAMPED partial film (each sample
already weighted 1/spp), the partial films are summed by one reduce (NCCL over NVLink on GPUs, gloo in the
CPU tests) and the root clamps -- after the sum, because the reference clamps after the spp-sum
(ky.cpp:3721-3726).  There is no other data-path collective.
"""
import torch
import torch.distributed as dist


def sample_range(spp, world, rank):
    """Sample indices [begin, end) of `rank`; ranges tile [0, max(1, spp)) and differ by at most one sample."""
    total = max(1, spp)
    base, extra = divmod(total, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def reduce_film(partial, dst=0, group=None, ordered=False):
    """Sums the ranks' partial films into rank `dst` and clamps there (ky.cpp:3726).  `partial` is a torch
    tensor on the device the process group's backend works with; returns it (meaningful on `dst` only).
    `ordered`: gather the partial films and add them in rank order instead of one reduce -- the sum is then the same
    bits whatever algorithm the backend's reduce would have picked (world_size films of memory on `dst`)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        root = dist.get_rank(group) == dst
        if ordered:
            parts = [torch.empty_like(partial) for _ in range(dist.get_world_size(group))] if root else None
            dist.gather(partial, parts, dst=dst, group=group)
            if root:
                partial.copy_(parts[0])
                for p in parts[1:]:
                    partial.add_(p)
        else:
            dist.reduce(partial, dst=dst, op=dist.ReduceOp.SUM, group=group)
    else:
        root = True
    if root:
        # std::clamp semantics: NaN stays NaN
        torch.clamp_(partial, 0.0, 1.0)
    return partial


def render_job(device, scene, desc, film, group=None, stream=None, ordered=False):
    """Renders this rank's share of `desc` (a whole-job kyd render desc: sample range (0, spp)) into the
    CUDA tensor `film` [h, w, 3] float32 and finishes the job with reduce_film().  `device` is a
    ky_b200.Device on the tensor's GPU with `scene` uploaded."""
    import ky_b200 as ky
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    begin, end = sample_range(desc.spp, world, rank)
    local = ky.RenderDesc.from_buffer_copy(desc)
    local.sample_begin, local.sample_end = begin, end
    local.flags = desc.flags & ~(ky.FLAG_CLAMP | ky.FLAG_ACCUMULATE)
    if end > begin:
        # (a context's own stream is a blocking stream: torch work queued on the default stream before this call -- the
        # film's zero fill -- is complete before the render touches the film; handle 0 selects that stream)
        device.render_device(local, film.data_ptr(), stream if stream is not None else torch.cuda.current_stream().cuda_stream)
    else:
        film.zero_()
    return reduce_film(film, 0, group, ordered)
