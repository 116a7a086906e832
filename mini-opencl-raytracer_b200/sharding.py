"""Multi-GPU partition of the path: one process per GPU, the scene replicated, work units
sharded, results gathered with a torch.distributed collective (NCCL over NVLink on GPUs, gloo on CPU in
the tests). Rays/pixels are independent, so the only exchange step is the gather of the framebuffer.

  ray streams ...... contiguous equal index ranges per rank, no collective (each rank keeps its hits)
  frames ........... `gid = y*W + x` (kernel_bvh.cl:394-395) split into bands of `band_rows` image rows dealt
                     round-robin to the ranks (sky and geometry rows interleave, so ranks finish together);
                     every rank renders its bands with one b2rt_execute_bands, then one all_gather of equal
                     contiguous shards + a de-interleave copy rebuilds the W*H*16-byte image on every rank.
"""
import torch
import torch.distributed as dist


def stream_range(n_items, world, rank):
    """Contiguous [lo, hi) of a ray stream for `rank` (sizes differ by at most one)."""
    lo = n_items * rank // world
    hi = n_items * (rank + 1) // world
    return lo, hi


class BandPlan:
    """Round-robin bands of image rows. The frame is padded (virtually) to a whole number of band rounds;
    padded rows are never rendered and are cut off after the gather."""

    def __init__(self, width, height, world, band_rows=8):
        if width <= 0 or height <= 0 or world <= 0 or band_rows <= 0:
            raise ValueError("width, height, world and band_rows must be positive")
        self.width, self.height, self.world, self.band_rows = width, height, world, band_rows
        self.n_bands = -(-height // band_rows)
        self.rounds = -(-self.n_bands // world)               # bands per rank (last round may be partly empty)
        self.band_pixels = band_rows * width

    def bands_of(self, rank):
        return [b for b in range(rank, self.n_bands, self.world)]

    def gid_ranges(self, rank):
        """[(gid_begin, gid_end)] this rank renders, clipped to the real frame."""
        n = self.width * self.height
        out = []
        for b in self.bands_of(rank):
            lo = b * self.band_pixels
            out.append((lo, min(lo + self.band_pixels, n)))
        return out

    def band_launch(self, rank):
        """This rank's share as (gid_begin, band_pixels, stride_pixels, n_full_bands, tail) for b2rt_execute_bands;
        `tail` is the (gid_begin, gid_end) of a last band clipped by the frame's bottom edge, or None."""
        ranges = self.gid_ranges(rank)
        full = [r for r in ranges if r[1] - r[0] == self.band_pixels]
        tail = [r for r in ranges if r[1] - r[0] != self.band_pixels]
        return rank * self.band_pixels, self.band_pixels, self.world * self.band_pixels, len(full), (tail[0] if tail else None)

    def render(self, ctx, rank):
        """Enqueue this rank's bands on a capi.Context: one strided launch sequence (+ one for a clipped last band)."""
        gid0, band, stride, n_full, tail = self.band_launch(rank)
        if n_full:
            ctx.execute_bands(gid0, band, stride, n_full)
        if tail:
            ctx.execute_range(*tail)

    def _buffer(self, name, shape, like):
        """Scratch tensors are allocated once per plan (a frame gather runs every frame)."""
        key = (name, tuple(shape), like.dtype, like.device)
        cache = self.__dict__.setdefault("_scratch", {})
        if key not in cache:
            cache[key] = torch.zeros(shape, dtype=like.dtype, device=like.device)
        return cache[key]

    def pack(self, frame, rank):
        """Own bands of a (W*H, C) frame as one contiguous (rounds*band_pixels, C) shard (zero padded). The whole
        bands are one strided copy (every world-th band of the frame), only a clipped last band is copied apart."""
        c = frame.shape[1]
        shard = self._buffer("shard%d" % rank, (self.rounds * self.band_pixels, c), frame)      # padding rows stay zero
        full = self.height // self.band_rows                      # bands not clipped by the bottom edge
        mine = len(range(rank, full, self.world))
        if mine:
            src = frame[: full * self.band_pixels].view(full, self.band_pixels, c)[rank::self.world]
            shard[: mine * self.band_pixels].view(mine, self.band_pixels, c).copy_(src)
        for i, (lo, hi) in enumerate(self.gid_ranges(rank)):
            if i >= mine:
                shard[i * self.band_pixels:i * self.band_pixels + (hi - lo)] = frame[lo:hi]
        return shard

    def unpack(self, gathered):
        """(world*rounds*band_pixels, C) all-gather result -> (W*H, C) frame (one de-interleave copy into a reused buffer)."""
        c = gathered.shape[1]
        g = gathered.view(self.world, self.rounds, self.band_pixels, c).permute(1, 0, 2, 3)   # band index = round*world + rank
        out = self._buffer("frame", (self.rounds, self.world, self.band_pixels, c), gathered)
        out.copy_(g)
        return out.view(-1, c)[: self.width * self.height]


def gather_frame(plan, frame, rank, group=None):
    """All ranks end up with the complete frame. `frame` is this rank's (W*H, C) buffer in which only its own
    bands are valid. One collective: all_gather of equal contiguous shards. The returned tensor is a buffer owned by
    `plan` and is overwritten by the next gather. `frame` is usually a zero-copy view of the library's output image, which the
    next frame's kernels (on the context's own stream) overwrite: the torch stream is therefore drained before returning, so
    that the pack copy has read `frame` and the result is complete when the caller goes on. (Harness utility: the product's
    multi-GPU path is b2rt_execute_shard / b2rt_create_multi, which need no gather.)"""
    shard = plan.pack(frame, rank)
    if plan.world == 1:
        out = plan.unpack(shard)
    else:
        gathered = plan._buffer("gathered", (plan.world * shard.shape[0], shard.shape[1]), shard)
        dist.all_gather_into_tensor(gathered, shard, group=group)
        out = plan.unpack(gathered)
    if out.is_cuda:
        torch.cuda.current_stream(out.device).synchronize()
    return out


class DeviceBuffer:
    """A torch view of device memory owned by libb2rt (e.g. the bound output buffer), no copy."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes // 4,), "typestr": "<f4", "data": (int(ptr), False), "version": 3, "strides": None}


def as_tensor(ptr, nbytes, device):
    return torch.as_tensor(DeviceBuffer(ptr, nbytes), device=device)
