"""Multi-GPU frames with ONE PROCESS PER GPU (torchrun): the scene replicated, the work sharded, one reduce per frame.

The reference's only parallelism is its bucket thread pool (src/threading.cpp, src/main.cpp:273-381). A frame is sharded
across GPUs instead:

  * Monte-Carlo frames: GPU g renders the samples {s : s % world == g} of every pixel into a private FP32 sum buffer;
  * Whitted frames: GPU g renders the 16-row bands {y : (y / 16) % world == g} (plus a one-row halo so the AA edge
    detector sees all eight neighbours) and leaves zeros elsewhere;

then ONE reduce sums the buffers on the first GPU, which resolves (divides by spp). Philox counters are keyed by
(pixel, sample), so the N-GPU image equals the 1-GPU image up to FP32 summation order. There is no other data-path collective.

Two ways to own the GPUs, the same sharding in both:
  * one process, one context over all GPUs: `Renderer(devices=[0, 1, ...])` - the library forks a host thread per GPU and
    sums the partial frames itself (one kernel reading the peers' buffers over NVLink, or NCCL); nothing to do here;
  * one process per GPU (torchrun, what the bench driver launches): `render_frame` below - the library renders this rank's
    shard, `torch.distributed.reduce` (NCCL over NVLink on GPUs, gloo in the CPU tests) sums the buffers on rank 0.
"""
from . import MODE_AUTO


def render_frame(renderer, acc, width, height, spp_total=0, seed=0, rank=0, world=1, mode=MODE_AUTO, dist=None, sync=None,
                 on_rendered=None, flags=0):
    """Render one frame into `acc` (a float32 tensor of width*height*3 elements in the memory the renderer writes:
    CUDA for the product). After the call rank 0 holds the finished frame; other ranks hold their partial sums.
    `dist` is torch.distributed (initialised) when world > 1. `sync()` must make the collective's result visible to the
    library's own stream (e.g. torch.cuda.synchronize); when it is None the tensor's device is synchronised.
    `on_rendered()` is called once this rank's shard is rendered, before the reduce. Returns the render stats of this rank."""
    st = renderer.render_device(acc.data_ptr(), width=width, height=height, mode=mode, spp=spp_total, seed=seed,
                                shard=(rank, world), flags=flags)
    if on_rendered is not None:
        on_rendered()
    if world > 1:
        if spp_total <= 0:
            # the frame's sample count: every rank takes part, whatever its own share was (a collective that only some ranks
            # enter would hang the job)
            import torch
            t = torch.tensor([st["spp_done"]], dtype=torch.int64, device=acc.device)
            dist.all_reduce(t)
            spp_total = int(t.item())
        dist.reduce(acc, dst=0)
        if sync is not None:
            sync()
        elif acc.is_cuda:
            import torch
            torch.cuda.synchronize(acc.device)
        # Monte-Carlo shards are un-normalised sums; Whitted shards are disjoint rows (no sample passes: nothing to divide)
        if rank == 0 and spp_total > 0:
            renderer.resolve_device(acc.data_ptr(), width, height, spp_total)
    return st
