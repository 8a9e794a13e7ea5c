"""Multi-GPU frames: one process per GPU, the scene replicated, the work sharded, one reduce per frame.

The reference's only parallelism is its bucket thread pool (src/threading.cpp, src/main.cpp:273-381). Here a frame is
sharded across ranks instead:

  * Monte-Carlo frames: rank g renders the samples {s : s % world == g} of every pixel into a private FP32 sum buffer;
  * Whitted frames: rank g renders the 16-row bands {y : (y / 16) % world == g} (plus a one-row halo so the AA edge
    detector sees all eight neighbours) and leaves zeros elsewhere;

then ONE `torch.distributed.reduce` (NCCL over NVLink on GPUs, gloo in the CPU tests) sums the buffers on rank 0, which
resolves (divides by spp). Philox counters are keyed by (pixel, sample), so the N-rank image equals the 1-rank image up
to FP32 summation order. There is no other data-path collective.
"""
from . import MODE_AUTO


def render_frame(renderer, acc, width, height, spp_total=0, seed=0, rank=0, world=1, mode=MODE_AUTO, dist=None, sync=None):
    """Render one frame into `acc` (a float32 tensor of width*height*3 elements in the memory the renderer writes:
    CUDA for the product). After the call rank 0 holds the finished frame; other ranks hold their partial sums.
    `dist` is torch.distributed (initialised) when world > 1; `sync()` is called before the library touches a
    buffer the collective wrote (the library has its own stream). Returns the render stats of this rank."""
    st = renderer.render_device(acc.data_ptr(), width=width, height=height, mode=mode, spp=spp_total, seed=seed,
                                shard=(rank, world))
    if world > 1:
        dist.reduce(acc, dst=0)
        if sync is not None:
            sync()
        if st["spp_done"] > 0:
            # Monte-Carlo shards are un-normalised sums (Whitted shards are disjoint rows: spp_done == 0, nothing to divide)
            if spp_total <= 0:
                import torch
                t = torch.tensor([st["spp_done"]], dtype=torch.int64, device=acc.device)
                dist.all_reduce(t)
                spp_total = int(t.item())
            if rank == 0:
                renderer.resolve_device(acc.data_ptr(), width, height, spp_total)
    return st
